"""GPU debug helper (not a pytest): in-kernel cycle counters of the tcgen05 JAC step kernel (CTA 0).
usage: python tests/gpu_debug_jac_prof.py [N] [n] [m]      (D4S_JAC_CTA_PAIR=1 profiles the CTA-pair instantiation)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from d4s import _cabi as abi
from d4s import engine
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
m = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
onet = orc.random_net(rng, m, 8, hidden=(256, 256, 256))
nse = nse_from_oracle(onet, dev)
xt = torch.as_tensor(rng.standard_normal((n, 8)), dtype=torch.float32, device=dev)
th = torch.as_tensor(rng.standard_normal((N, m)), dtype=torch.float32, device=dev)
t = torch.tensor([0.2], dtype=torch.float32)
setup = engine.build_setup(nse, xt, N, t, None, 0.0, engine.PriorSpec(abi.PRIOR_NONE), "JAC", None)
for _ in range(3):
    engine.score_partials(setup, th, 0)
torch.cuda.synchronize()
abi.set_option("profile", 1)
reps = 5
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(reps):
    engine.score_partials(setup, th, 0)
ev1.record()
torch.cuda.synchronize()
buf = (C.c_int64 * 16)()
abi.check(abi.load().d4s_get_profile(buf))
v = [int(b) for b in buf]
names = ["mma: wait operand", "mma: wait ring", "mma: issue loop", "mma: operand ready -> layer complete", "layers",
         "cw: wait last layer of prev tile", "cw: layer 0", "cw: last epilogue", "cw: wait hidden layer", "cw: mid epilogue",
         "cw: pairs", "cw: meta", "cw: end barrier"]
layers = max(1, v[4])
tiles = layers / 2
print(f"N={N} n={n} m={m}: {ev0.elapsed_time(ev1) / reps * 1e3:.1f} us per launch (profiled build, MMA warp drains per layer); "
      f"CTA 0: {tiles / reps:.1f} tiles per launch")
for i, nm in enumerate(names):
    if i == 4:
        continue
    per = v[i] / (layers if i < 4 else tiles)
    print(f"  {nm:40s} {per:9.0f} cycles per {'layer' if i < 4 else 'tile'}")
