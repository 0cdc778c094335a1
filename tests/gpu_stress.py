"""Stress-shape check (BASELINE config 5 scaled down): theta=10, x=10, MLP 26->256^3->10, n_obs=10^4, Gaussian prior.
Single process: fused tcgen05 trajectory vs the fp64 oracle on a few samples, then timing at N samples.
Under torchrun (>= 2 ranks): observation-sharded run (tcgen05 kernel in export mode + NCCL all-reduce per step)
vs the single-GPU result.   usage: python tests/gpu_stress.py [N_time=2048] [steps=3] [n_obs=10000] [check_steps=steps]
(check_steps = 0 skips the oracle comparison, e.g. for long timing runs)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

import d4s
from d4s import parallel
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

N_time = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
check_steps = int(sys.argv[4]) if len(sys.argv) > 4 else steps
m, d = 10, 10
rng = np.random.default_rng(0)
onet = orc.random_net(rng, m, d, hidden=(256, 256, 256))
nse = nse_from_oracle(onet, dev)
x = rng.standard_normal((n, d))
a = 0.2 * rng.standard_normal((n, m, m))
cov = a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)
xt = torch.as_tensor(x, dtype=torch.float32, device=dev)
ct = torch.as_tensor(cov, dtype=torch.float32, device=dev)
mu, pc = torch.zeros(m, device=dev), torch.eye(m, device=dev)
prior, fn = torch.distributions.MultivariateNormal(mu, pc), d4s.get_vpdiff_gaussian_score(mu, pc, nse)
kw = dict(steps=steps, eta=1.0, prior=prior, prior_type="gaussian", prior_score_fn=fn, dist_cov_est=ct, cov_mode="GAUSS")
ok = True

Nc = 24
z0, z = rng.standard_normal((Nc, m)), rng.standard_normal((steps, Nc, m))
noise = (torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32))
single = nse.ddim((Nc,), xt, _noise=noise, **kw)
torch.cuda.synchronize()
if rank == 0 and check_steps > 0:
    t0 = time.time()
    ref = orc.ddim(onet, x, z0, z, steps, 1.0, orc.GaussianPrior(np.zeros(m), np.eye(m)), cov, "GAUSS", "gaussian")
    scale = max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(single.cpu().numpy() - ref).max()) / scale
    print(f"stress n_obs={n}: fused tcgen05 trajectory vs fp64 oracle ({Nc} samples, {steps} steps): max rel err {err:.2e} "
          f"(oracle took {time.time() - t0:.1f}s)")
    ok &= err < 1e-3
if world > 1:
    sharded = parallel.ddim_obs_sharded(nse, (Nc,), xt, _noise=noise, **kw)
    e2 = float((sharded - single).abs().max()) / max(1.0, float(single.abs().max()))
    if rank == 0:
        print(f"observation-sharded over {world} GPUs vs 1 GPU: max rel diff {e2:.2e}")
    ok &= e2 < 1e-4

# timing
def timed(fn_):
    fn_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    fn_()
    torch.cuda.synchronize()
    return time.time() - t0

dt = timed(lambda: nse.ddim((N_time,), xt, _seed=1, **kw))
pairs = N_time * n * steps
flops = pairs * 2 * (26 * 256 + 256 * 256 * 2 + 256 * 10)
if rank == 0:
    print(f"1 GPU fused: N={N_time} n={n} steps={steps}: {dt * 1e3 / steps:.1f} ms/step, {flops / dt / 1e12:.1f} algorithmic TFLOP/s, "
          f"extrapolated full stress run (N=1e5, 1000 steps): {dt / steps / N_time * 1e5 * 1000 / 60:.1f} min/GPU")
if world > 1:
    dt2 = timed(lambda: parallel.ddim_obs_sharded(nse, (N_time,), xt, _seed=1, **kw))
    if rank == 0:
        print(f"{world} GPUs observation-sharded: {dt2 * 1e3 / steps:.1f} ms/step => speed-up {dt / dt2:.2f}x "
              f"(all-reduce of {N_time * 2 * m * 4 / 1e6:.2f} MB per step)")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(int(flag))
    parallel.shutdown()
    dist.destroy_process_group()
if rank == 0:
    print("STRESS_OK" if ok else "STRESS_FAIL")
sys.exit(0 if ok else 1)
