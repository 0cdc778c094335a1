"""GPU debug helper (not a pytest): cProfile of the host side of one end-to-end `NSE.ddim` call on the bench shape."""
import sys, os, cProfile, pstats, io, argparse, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import numpy as np, torch
import bench, d4s
args = argparse.Namespace(task="slcp", n_obs=30, ddim_steps=None, cov_mode="GAUSS", samples=1000)
dev = torch.device("cuda:0")
w = bench.workload(args)
nse = bench.build_nse(w, dev)
x_h = torch.as_tensor(w["x"]).pin_memory(); cov_h = torch.as_tensor(w["cov"]).pin_memory()
lo_h = torch.as_tensor(w["prior"]["low"]).pin_memory(); hi_h = torch.as_tensor(w["prior"]["high"]).pin_memory()
out_h = torch.empty(1000, 5).pin_memory()
def step(it):
    x = x_h.to(dev, non_blocking=True); cov = cov_h.to(dev, non_blocking=True)
    lo, hi = lo_h.to(dev, non_blocking=True), hi_h.to(dev, non_blocking=True)
    prior, fn = torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
    th = nse.ddim((1000,), x, steps=1000, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn, dist_cov_est=cov, cov_mode="GAUSS", _seed=it)
    out_h.copy_(th, non_blocking=True); torch.cuda.current_stream().synchronize()
for i in range(3): step(i)
t0=time.perf_counter()
for i in range(5): step(10+i)
print("e2e ms/step", (time.perf_counter()-t0)/5*1e3)
# host-only part: time until the launch returns
pr = cProfile.Profile(); pr.enable()
for i in range(5): step(20+i)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
