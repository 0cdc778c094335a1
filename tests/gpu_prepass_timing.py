"""GPU helper (not a pytest): the covariance pre-pass of the GAUSS mode (sbibm_posterior_estimation.py:306-314,
`NSE.estimate_dist_cov`: 1000 draws from every single-observation posterior, 100 DDIM steps, eta 0.5) on the paired-mode
tcgen05 kernel against the fp32 SIMT kernels.  usage: python tests/gpu_prepass_timing.py [n_obs] [nsamples]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from d4s import _cabi as abi
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
onet = orc.random_net(rng, 5, 8, hidden=(256, 256, 256))
nse = nse_from_oracle(onet, dev)
x = torch.as_tensor(rng.standard_normal((n, 8)), dtype=torch.float32, device=dev)
res = {}
for name, opt in (("tcgen05 paired kernel", 1), ("fp32 SIMT kernels", 0)):
    abi.set_option("paired_tc", opt)
    for _ in range(2):
        cov = nse.estimate_dist_cov(x, nsamples=S, steps=100, eta=0.5, _seed=7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        cov = nse.estimate_dist_cov(x, nsamples=S, steps=100, eta=0.5, _seed=7)
    torch.cuda.synchronize()
    res[name] = ((time.perf_counter() - t0) / reps * 1e3, cov)
    print(f"{name:24s}: {res[name][0]:8.2f} ms per pre-pass (n_obs = {n}, {S} samples each, 100 steps; host tables included)")
abi.set_option("paired_tc", 1)
a, b = res["tcgen05 paired kernel"][1], res["fp32 SIMT kernels"][1]
print(f"max |cov_tc - cov_simt| / max |cov| = {float((a - b).abs().max() / b.abs().max()):.2e} (same Philox noise)")
