"""Probe: trajectory-kernel time on the stress shape (theta = 10, n_obs = 10^4) as a function of the observation chunk
(tile = SG samples x OC observations).   usage: python tests/gpu_obs_chunk_probe.py [N=2048] [steps=20]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import argparse

import torch

import bench
import d4s
from d4s import engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
a = argparse.Namespace(task="stress", n_obs=10000, cov_mode="GAUSS", ddim_steps=steps, samples=N)
w = bench.workload(a)
nse = bench.build_nse(w, dev)
x, cov = torch.as_tensor(w["x"]).to(dev), torch.as_tensor(w["cov"]).to(dev)
mu, pc = torch.as_tensor(w["prior"]["mean"]).to(dev), torch.as_tensor(w["prior"]["cov"]).to(dev)
prior, fn = torch.distributions.MultivariateNormal(mu, pc, validate_args=False), d4s.get_vpdiff_gaussian_score(mu, pc, nse)
spec = engine.prior_spec("gaussian", prior, fn, w["m"])
grid = engine.time_grid(steps)
ref = None
for oc in (0, 125, 64, 63, 42, 32, 16):
    st = engine.build_setup(nse, x, N, grid[:-1], grid[:-1].double() - 1.0 / steps, 1.0, spec, "GAUSS", cov, seed=7, obs_chunk=oc)
    th0 = engine.philox_normal(N, w["m"], 7, -1, 0, dev)
    th = th0.clone()
    engine.run_trajectory(st, th, None, True)
    torch.cuda.synchronize()
    th.copy_(th0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.run_trajectory(st, th, None, True)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if ref is None:
        ref = th.clone()
    d = float((th - ref).abs().max())
    print(f"obs_chunk={oc:4d}: {ms:8.3f} ms/step   max |theta - theta(default)| = {d:.2e}", flush=True)
