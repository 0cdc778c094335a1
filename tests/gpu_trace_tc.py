"""Event timeline of the tcgen05 trajectory kernel (not a pytest).  Needs a -DD4S_TRACE build:
    python tests/build_variant.py trace -DD4S_TRACE [-DD4S_EXP=...]
    D4S_LIB_PATH=$PWD/ab/libd4s_trace.so python tests/gpu_trace_tc.py [N] [n] [steps]
Prints, for the traced items of CTA 0, when every phase ended on the fastest / slowest compute warp and when the MMA warp
issued each layer (cycles relative to the first traced event)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
path = os.path.join(ROOT, "gpurun_out", "trace.bin")
os.makedirs(os.path.dirname(path), exist_ok=True)
os.environ["D4S_TRACE_FILE"] = path
import numpy as np
import torch

import d4s
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
m, d = 5, 8
onet = orc.random_net(rng, m, d, hidden=(256, 256, 256))
nse = nse_from_oracle(onet, dev)
x = rng.standard_normal((n, d))
a = 0.2 * rng.standard_normal((n, m, m))
cov = a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)
lo, hi = torch.full((m,), -3.46, device=dev), torch.full((m,), 3.46, device=dev)
prior = torch.distributions.Uniform(lo, hi, validate_args=False)
fn = d4s.get_vpdiff_uniform_score(lo, hi, nse)
kw = dict(steps=steps, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn,
          dist_cov_est=torch.as_tensor(cov, dtype=torch.float32, device=dev), cov_mode="GAUSS",
          theta_clipping_range=(-3.0, 3.0), _seed=1)
out = nse.ddim((N,), torch.as_tensor(x, dtype=torch.float32, device=dev), **kw)
torch.cuda.synchronize()
tr = np.fromfile(path, dtype=np.int64).reshape(20, -1, 2)
names = {0: "base", 1: "layer0", 2: "helper-wait", 3: "tc-wait", 4: "epilogue", 5: "sums-final", 6: "bar", 10: "tail",
         11: "sums-tree", 20: "ITEM", 21: "->wait-last", 22: "->wait-mid"}
t0 = min(int(tr[w, 0, 1]) for w in range(20) if tr[w, 0, 1] > 0)
cw = [w for w in range(16) if tr[w, 0, 1] > 0]
nev = min(int((tr[w, :, 1] > 0).sum()) for w in cw)
print(f"compute warps: {len(cw)}, events per warp: {nev}; columns: tag, end time on the fastest / slowest warp, warp 0, spread")
prev_max = 0
for e in range(nev):
    tags = {int(tr[w, e, 0]) for w in cw}
    ts = [int(tr[w, e, 1]) - t0 for w in cw]
    tag = tags.pop() if len(tags) == 1 else -1
    print(f"  {names.get(tag, str(tag)):12s} min {min(ts):7d}  max {max(ts):7d}  warp0 {ts[0]:7d}  spread {max(ts) - min(ts):6d}  "
          f"(+{max(ts) - prev_max:6d} after the previous event's slowest warp; slowest = warp {cw[int(np.argmax(ts))]})")
    prev_max = max(ts)
w = 16
print("TMA warp: ring slot freed (= MMAs of the K-step 4 stages back completed), deltas in cycles:")
ts = [int(tr[w, e, 1]) - t0 for e in range(int((tr[w, :, 1] > 0).sum()))]
tg = [int(tr[w, e, 0]) % 10 for e in range(len(ts))]
for i in range(0, len(ts), 16):
    print("  layer", tg[i], "first at", ts[i], " deltas:", " ".join(str(ts[j] - ts[j - 1]) for j in range(max(i, 1), min(i + 16, len(ts)))))
w = 17
print("MMA warp:")
for e in range(int((tr[w, :, 1] > 0).sum())):
    tag = int(tr[w, e, 0])
    what = {3: "first MMA issued, layer", 4: "all MMAs issued, layer", 5: "operand ready seen, layer"}[tag // 10]
    print(f"  {what} {tag % 10}: {int(tr[w, e, 1]) - t0:7d}")
