"""GPU debug helper (not a pytest): the bench's JAC trajectory on the tcgen05 JAC kernel vs the fp32 SIMT kernels."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import bench
import d4s
from d4s import _cabi as abi

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
args = argparse.Namespace(task="slcp", n_obs=30, ddim_steps=steps, cov_mode="JAC", samples=N)
dev = torch.device("cuda:0")
w = bench.workload(args)
nse = bench.build_nse(w, dev)
lo, hi = torch.as_tensor(w["prior"]["low"]).to(dev), torch.as_tensor(w["prior"]["high"]).to(dev)
prior, fn = torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
x, cov = torch.as_tensor(w["x"]).to(dev), torch.as_tensor(w["cov"]).to(dev)
out = {}
for path in (1, 0):
    abi.set_option("path", path)
    abi.set_option("jac_tc", 1)
    th = nse.ddim((N,), x, steps=steps, eta=w["eta"], prior=prior, prior_type="uniform", prior_score_fn=fn,
                  dist_cov_est=cov, cov_mode="JAC", _seed=7)
    torch.cuda.synchronize()
    out[path] = th
    bad = ~torch.isfinite(th).all(1)
    print(f"path={path}: non-finite samples {int(bad.sum())} / {N}, max |theta| over finite = {float(th[~bad].abs().max()):.3e}")
abi.set_option("path", 0)
ok = torch.isfinite(out[0]).all(1) & torch.isfinite(out[1]).all(1)
d = (out[0] - out[1])[ok].abs()
print(f"tc vs simt: median |d| {float(d.median()):.2e}, 99% {float(d.flatten().quantile(0.99)):.2e}, max {float(d.max()):.2e}")
