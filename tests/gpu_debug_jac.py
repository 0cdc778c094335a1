"""GPU debug helper (not a pytest): partial sums of one JAC step, tcgen05 score + Jacobian kernel vs fp32 SIMT kernel
vs the fp64 oracle.  usage: python tests/gpu_debug_jac.py [m] [N] [n] [t] [hidden...]"""
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

m = int(sys.argv[1]) if len(sys.argv) > 1 else 5
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tval = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
hidden = tuple(int(v) for v in sys.argv[5:]) or (256, 256, 256)
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
d = 8
onet = orc.random_net(rng, m, d, hidden=hidden)
nse = nse_from_oracle(onet, dev)
x = rng.standard_normal((n, d))
theta = rng.standard_normal((N, m))
xt = torch.as_tensor(x, dtype=torch.float32, device=dev)
th = torch.as_tensor(theta, dtype=torch.float32, device=dev)
t = torch.tensor([tval], dtype=torch.float32)
prior = engine.PriorSpec(abi.PRIOR_NONE)
W = m + m * m


def partials(path):
    abi.set_option("path", path)
    abi.set_option("jac_tc", 1)
    setup = engine.build_setup(nse, xt, N, t, None, 0.0, prior, "JAC", None)
    engine.score_partials(setup, th, 0)
    torch.cuda.synchronize()
    nf = abi.load().d4s_partials_floats(setup.pn.handle, C.byref(setup.prob))
    ws = setup.keep[4][:nf].clone().reshape(N, -1, W).double().sum(1)
    return ws.cpu().numpy(), nf // (N * W)


simt, c1 = partials(1)
tc, c0 = partials(0)
abi.set_option("path", 0)
# oracle: per-pair score and Jacobian in fp64
prec, _, sc = orc.tweedies_approximation(onet, x, theta, np.float64(np.float32(tval)), None, "JAC")
ref = np.concatenate([(prec @ sc[..., None])[..., 0].sum(1), prec.sum(1).reshape(N, -1)], axis=1)


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


print(f"m={m} N={N} n={n} t={tval} hidden={hidden}: chunks simt={c1} tc={c0}")
print(f"  sum P s : simt-oracle {rel(simt[:, :m], ref[:, :m]):.2e}   tc-oracle {rel(tc[:, :m], ref[:, :m]):.2e}")
print(f"  sum P   : simt-oracle {rel(simt[:, m:], ref[:, m:]):.2e}   tc-oracle {rel(tc[:, m:], ref[:, m:]):.2e}")
