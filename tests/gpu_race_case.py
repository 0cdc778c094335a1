"""Runs one small trajectory on the tcgen05 kernel twice (interleaved / sequential group schedule) -- a target for
compute-sanitizer --tool racecheck (not a pytest).  usage: python tests/gpu_race_case.py m d n N h1 h2 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from d4s import _cabi as abi
from d4s_helpers import nse_from_oracle
from test_gpu_parity import _d4s_prior, _problem, _t

m, d, n, N = (int(v) for v in sys.argv[1:5])
hidden = tuple(int(v) for v in sys.argv[5:]) or (64, 128)
dev = torch.device("cuda:0")
onet, x, cov, theta, oprior = _problem(m, d, hidden, n, N, "uniform", seed=3)
nse = nse_from_oracle(onet, dev)
prior, fn = _d4s_prior(oprior, nse, dev)
steps = int(os.environ.get("STEPS", "5"))
rng = np.random.default_rng(1)
z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
kw = dict(steps=steps, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn, dist_cov_est=_t(cov, dev), cov_mode="GAUSS",
          _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
abi.set_option("path", 2)
outs = []
for rep in range(int(os.environ.get("REPS", "3"))):
    outs.append(nse.ddim((N,), _t(x, dev), **kw))
abi.set_option("tc_pairing", 0)
seq = nse.ddim((N,), _t(x, dev), **kw)
torch.cuda.synchronize()
for i, o in enumerate(outs):
    d_ = (o - seq).abs()
    bad = (d_ > 0).any(dim=1).nonzero().flatten()
    print(f"run {i}: equal to the sequential schedule = {bool(torch.equal(o, seq))}; differing samples: {bad[:20].tolist()} "
          f"(of {bad.numel()}), max |diff| = {float(d_.max()):.3e}")
