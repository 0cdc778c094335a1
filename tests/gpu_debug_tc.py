"""GPU debug helper (not a pytest): compares the fused tcgen05 trajectory kernel with the per-step SIMT kernels.
usage: python tests/gpu_debug_tc.py <swap 0|1> [N] [n] [steps] [hidden...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import d4s
from d4s import _cabi as abi
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

swap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
hidden = tuple(int(v) for v in sys.argv[5:]) or (256, 256, 256)
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
m, d = 5, 8
onet = orc.random_net(rng, m, d, hidden=hidden)
nse = nse_from_oracle(onet, dev)
x = rng.standard_normal((n, d))
a = 0.2 * rng.standard_normal((n, m, m))
cov = a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)
lo, hi = torch.full((m,), -3.46, device=dev), torch.full((m,), 3.46, device=dev)
prior = torch.distributions.Uniform(lo, hi, validate_args=False)
fn = d4s.get_vpdiff_uniform_score(lo, hi, nse)
z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
kw = dict(steps=steps, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn,
          dist_cov_est=torch.as_tensor(cov, dtype=torch.float32, device=dev), cov_mode="GAUSS",
          _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
xt = torch.as_tensor(x, dtype=torch.float32, device=dev)
abi.set_option("path", 1)
ref = nse.ddim((N,), xt, **kw)
torch.cuda.synchronize()
abi.set_option("path", 2)
out = nse.ddim((N,), xt, **kw)
torch.cuda.synchronize()
orc_out = orc.ddim(onet, x, z0, z, steps, 1.0, orc.UniformPrior(np.full(m, -3.46), np.full(m, 3.46)), cov, "GAUSS", "uniform")
scale = max(1.0, float(np.abs(orc_out).max()))
print(f"swap={swap} N={N} n={n} steps={steps} hidden={hidden}: |tc-simt|max={float((out - ref).abs().max()):.3e} "
      f"|simt-oracle|={np.abs(ref.cpu().numpy() - orc_out).max() / scale:.3e} "
      f"|tc-oracle|={np.abs(out.cpu().numpy() - orc_out).max() / scale:.3e} (rel to {scale:.2f})")

abi.set_option("tc_pairing", 0)  # one sample group after the other: same arithmetic, so the result must be bit-identical
out_seq = nse.ddim((N,), xt, **kw)
torch.cuda.synchronize()
abi.set_option("tc_pairing", 1)
print("interleaved vs sequential group schedule: bit-identical =", bool(torch.equal(out, out_seq)),
      "finite =", bool(torch.isfinite(out).all()))

if os.environ.get("D4S_PROFILE"):
    abi.set_option("profile", 1)
    out = nse.ddim((N,), xt, **kw)
    torch.cuda.synchronize()
    prof = abi.get_profile()
    mma = {k: prof.pop(k) for k in list(prof) if k.startswith("mma_")}
    tot = sum(prof.values())
    groups = -(-N // max(1, min(8, 128 // min(n, 128)))) 
    print("phase cycles (CTA 0):", {k: f"{v} ({100 * v / tot:.1f}%)" for k, v in prof.items()}, "total", tot, "| MMA issue->commit cycles:", mma)
