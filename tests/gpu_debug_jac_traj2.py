"""debug: JAC trajectory test shape, tc vs simt per call and per trajectory (not a pytest)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from d4s import _cabi as abi
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc
from test_gpu_parity import _d4s_prior, _problem, _t
dev = torch.device("cuda:0")
m, d, n, N, steps = 4, 6, 9, 200, 10
hidden = tuple(int(v) for v in sys.argv[1:]) or (128, 256)
onet, x, cov, theta, oprior = _problem(m, d, hidden, n, N, "uniform", seed=9)
nse = nse_from_oracle(onet, dev)
prior, fn = _d4s_prior(oprior, nse, dev)
for tv in (0.3, 0.2, 0.1):
    t32 = np.float32(tv)
    ref = orc.factorized_score(onet, theta, x, np.float64(t32), oprior, None, "JAC", "uniform")
    outs = []
    for jt in (1, 0):
        abi.set_option("jac_tc", jt)
        o = nse.factorized_score(_t(theta, dev), _t(x, dev), torch.tensor(t32), fn, prior, cov_mode="JAC", prior_type="uniform")
        outs.append(orc.rel_l2(o.cpu().numpy(), ref))
    print(f"t={tv}: per-call rel err tc {outs[0]:.2e} simt {outs[1]:.2e}")
rng = np.random.default_rng(2)
z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
for st in (8, 9, 10):
    kw = dict(steps=st, eta=0.8, prior=prior, prior_type="uniform", prior_score_fn=fn, cov_mode="JAC",
              _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z[:st], dtype=torch.float32)))
    ref = orc.ddim(onet, x, z0, z[:st], st, 0.8, oprior, None, "JAC", "uniform")
    for jt in (1, 0):
        abi.set_option("jac_tc", jt)
        out = nse.ddim((N,), _t(x, dev), **kw).cpu().numpy()
        e = np.abs(out - ref).max(axis=1) / max(1.0, np.abs(ref).max())
        print(f"steps={st} jac_tc={jt}: max rel err {e.max():.2e}, samples with err > 1e-3: {(e > 1e-3).sum()} of {N}; worst {np.argsort(-e)[:5].tolist()}")
