"""GPU debug helper (not a pytest): one fuzz case of test_tc_per_call_random_shapes against the fp64 oracle, per sample."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from d4s import _cabi as abi
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc
import test_gpu_parity as T

fuzz_seed, want = int(sys.argv[1]), sys.argv[2]
case = [c for c in T._random_tc_cases(150, fuzz_seed) if f"m{c[0]}_d{c[1]}_h{'x'.join(map(str, c[2]))}_n{c[3]}_N{c[4]}_{c[5]}" == want][0]
m, d, hidden, n, N, pkind, seed = case
dev = torch.device("cuda:0")
onet, x, cov, theta, oprior = T._problem(m, d, hidden, n, N, pkind, seed=seed)
nse = nse_from_oracle(onet, dev)
prior, fn = T._d4s_prior(oprior, nse, dev)
th = T._t(theta[:N], dev)
tv = np.float32(0.2)
args = (th, T._t(x, dev), torch.tensor(tv), fn, prior)
kw = dict(dist_cov_est=T._t(cov, dev), cov_mode="JAC", prior_type=pkind)
abi.set_option("path", 1)
simt = nse.factorized_score(*args, **kw).cpu().numpy()
abi.set_option("path", 0)
abi.set_option("jac_tc", 1)
tc = nse.factorized_score(*args, **kw).cpu().numpy()
abi.set_option("jac_tc", 0)
ref = orc.factorized_score(onet, theta[:N], x, np.float64(tv), oprior, cov, "JAC", pkind)
ref32 = orc.factorized_score(onet.astype(np.float32), theta[:N].astype(np.float32), x.astype(np.float32), tv, oprior,
                             cov.astype(np.float32), "JAC", pkind)
nrm = np.linalg.norm(ref)
for name, o in (("simt", simt), ("tc", tc), ("oracle fp32", ref32)):
    e = np.linalg.norm(o - ref, axis=1)
    print(f"{name:12s}: rel L2 vs fp64 oracle {np.linalg.norm(o - ref) / nrm:.2e}; worst sample {int(e.argmax())} |d|={e.max():.2e} (|ref|={np.linalg.norm(ref[e.argmax()]):.2e}); "
          f"median per-sample rel {np.median(e / np.linalg.norm(ref, axis=1)):.2e}")
