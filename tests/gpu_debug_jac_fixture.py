"""GPU debug helper (not a pytest): JAC partial sums on a golden fixture net, tcgen05 JAC kernel vs fp32 SIMT vs fp64 oracle.
usage: python tests/gpu_debug_jac_fixture.py [case=lv_n6]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import golden_io as gio
from d4s import _cabi as abi
from d4s import engine
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

name = sys.argv[1] if len(sys.argv) > 1 else "lv_n6"
dev = torch.device("cuda:0")
g = gio.load(name)
onet = gio.oracle_net(g)
nse = nse_from_oracle(onet, dev)
x, theta = np.asarray(g["x"], np.float64), np.asarray(g["theta"], np.float64)
N, m = theta.shape
W = m + m * m
xt = torch.as_tensor(x, dtype=torch.float32, device=dev)
th = torch.as_tensor(theta, dtype=torch.float32, device=dev)
print(name, "N", N, "n", x.shape[0], "m", m)
for tval in g["t_grid"]:
    t = torch.tensor([np.float32(tval)])

    def partials(path):
        abi.set_option("path", path)
        abi.set_option("jac_tc", 1)
        setup = engine.build_setup(nse, xt, N, t, None, 0.0, engine.PriorSpec(abi.PRIOR_NONE), "JAC", None)
        engine.score_partials(setup, th, 0)
        torch.cuda.synchronize()
        nf = abi.load().d4s_partials_floats(setup.pn.handle, C.byref(setup.prob))
        return setup.keep[4][:nf].clone().reshape(N, -1, W).double().sum(1).cpu().numpy()

    simt, tc = partials(1), partials(0)
    prec, _, sc = orc.tweedies_approximation(onet, x, theta, np.float64(np.float32(tval)), None, "JAC")
    ref = np.concatenate([(prec @ sc[..., None])[..., 0].sum(1), prec.sum(1).reshape(N, -1)], axis=1)
    cond = np.linalg.cond(np.linalg.inv(prec.reshape(-1, m, m))).max()
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    print(f"  t={float(tval):.2f}: max cond(I - sigma J) {cond:9.1f}   sum P: simt {rel(simt[:, m:], ref[:, m:]):.2e}  tc {rel(tc[:, m:], ref[:, m:]):.2e}"
          f"   sum P s: simt {rel(simt[:, :m], ref[:, :m]):.2e}  tc {rel(tc[:, :m], ref[:, :m]):.2e}")
abi.set_option("path", 0)
abi.set_option("jac_tc", 0)
