"""GPU timing helper (not a pytest): batched contexts (NSE.ddim_contexts) against a Python loop of NSE.ddim calls.
usage: python tests/gpu_contexts_timing.py [B=1000] [n=30] [S=1] [steps=100]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import d4s
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
m, d = 4, 8
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
onet = orc.random_net(rng, m, d, hidden=(256, 256, 256))
nse = nse_from_oracle(onet, dev)
x = torch.as_tensor(0.5 * rng.standard_normal((B, n, d)), dtype=torch.float32, device=dev)
a = 0.2 * rng.standard_normal((B, n, m, m))
cov = torch.as_tensor(a @ a.transpose(0, 1, 3, 2) + 0.05 * np.eye(m), dtype=torch.float32, device=dev)
lo, hi = torch.full((m,), -3.46, device=dev), torch.full((m,), 3.46, device=dev)
prior = torch.distributions.Uniform(lo, hi, validate_args=False)
fn = d4s.get_vpdiff_uniform_score(lo, hi, nse)
kw = dict(steps=steps, eta=0.5, prior=prior, prior_type="uniform", prior_score_fn=fn, cov_mode="GAUSS",
          theta_clipping_range=(-3.0, 3.0), _seed=3)


def timed(f):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = f()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


dt_b, out = timed(lambda: nse.ddim_contexts((S,), x, dist_cov_est=cov, **kw))
nloop = min(B, 20)
dt_l, _ = timed(lambda: [nse.ddim((S,), x[c], dist_cov_est=cov[c], **kw) for c in range(nloop)])
print(f"B={B} contexts x n={n} obs x S={S} samples, {steps} DDIM steps (GAUSS, uniform prior, theta_dim {m}): "
      f"ddim_contexts {dt_b * 1e3:.1f} ms = {B / dt_b:.0f} contexts/s; loop of ddim calls {dt_l / nloop * 1e3:.1f} ms per "
      f"context = {nloop / dt_l:.0f} contexts/s; finite: {bool(torch.isfinite(out).all())}")
