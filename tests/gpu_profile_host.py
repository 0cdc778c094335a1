"""GPU debug helper (not a pytest): cProfile of the host side of one end-to-end `NSE.ddim` call on a bench shape.
usage: python tests/gpu_profile_host.py [task, default slcp] [GAUSS|JAC]"""
import sys, os, cProfile, pstats, io, argparse, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import numpy as np, torch
import bench, d4s
task = sys.argv[1] if len(sys.argv) > 1 else "slcp"
mode = sys.argv[2] if len(sys.argv) > 2 else "GAUSS"
args = argparse.Namespace(task=task, n_obs=30, ddim_steps=None, cov_mode=mode, samples=1000)
dev = torch.device("cuda:0")
w = bench.workload(args)
nse = bench.build_nse(w, dev)
N, m = 1000, w["m"]
x_h = torch.as_tensor(w["x"]).pin_memory(); cov_h = torch.as_tensor(w["cov"]).pin_memory()
if w["prior_kind"] == "uniform":
    a_h, b_h = torch.as_tensor(w["prior"]["low"]).pin_memory(), torch.as_tensor(w["prior"]["high"]).pin_memory()
else:
    a_h, b_h = torch.as_tensor(w["prior"]["mean"]).pin_memory(), torch.as_tensor(w["prior"]["cov"]).pin_memory()
out_h = torch.empty(N, m).pin_memory()
def step(it):
    x = x_h.to(dev, non_blocking=True); cov = cov_h.to(dev, non_blocking=True)
    a, b = a_h.to(dev, non_blocking=True), b_h.to(dev, non_blocking=True)
    if w["prior_kind"] == "uniform":
        prior, fn = torch.distributions.Uniform(a, b, validate_args=False), d4s.get_vpdiff_uniform_score(a, b, nse)
    else:
        prior, fn = torch.distributions.MultivariateNormal(a, b, validate_args=False), d4s.get_vpdiff_gaussian_score(a, b, nse)
    th = nse.ddim((N,), x, steps=w["ddim_steps"], eta=w["eta"], prior=prior, prior_type=w["prior_kind"], prior_score_fn=fn,
                  dist_cov_est=cov, cov_mode=mode, _seed=it)
    out_h.copy_(th, non_blocking=True); torch.cuda.current_stream().synchronize()
for i in range(3): step(i)
t0=time.perf_counter()
for i in range(5): step(10+i)
print(task, mode, "e2e ms/step", (time.perf_counter()-t0)/5*1e3)
pr = cProfile.Profile(); pr.enable()
for i in range(5): step(20+i)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(32); print(s.getvalue()[:6500])
