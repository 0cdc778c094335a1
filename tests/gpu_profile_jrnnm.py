"""GPU debug helper (not a pytest): in-kernel phase counters of the trajectory kernel on a bench task shape.
usage: python tests/gpu_profile_jrnnm.py [task=jrnnm] [N=1184] [steps=20]"""
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

task = sys.argv[1] if len(sys.argv) > 1 else "jrnnm"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
args = argparse.Namespace(task=task, n_obs=30, ddim_steps=steps, cov_mode="GAUSS", samples=N)
dev = torch.device("cuda:0")
w = bench.workload(args)
nse = bench.build_nse(w, dev)
lo, hi = torch.as_tensor(w["prior"]["low"]).to(dev), torch.as_tensor(w["prior"]["high"]).to(dev)
prior, fn = torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
x, cov = torch.as_tensor(w["x"]).to(dev), torch.as_tensor(w["cov"]).to(dev)
abi.set_option("profile", 1)
abi.set_option("path", 2)
th = nse.ddim((N,), x, steps=steps, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn, dist_cov_est=cov,
              cov_mode="GAUSS", _seed=7)
torch.cuda.synchronize()
prof = abi.get_profile()
mma = {k: prof.pop(k) for k in list(prof) if k.startswith("mma_")}
tot = sum(prof.values())
items = steps * (2 if N > 148 * 4 else 1)
print(f"{task} N={N} steps={steps}: cycles per item {tot / items:.0f}:", {k: round(v / items) for k, v in prof.items()},
      "| MMA issue->commit per item:", {k: round(v / items) for k, v in mma.items()})
