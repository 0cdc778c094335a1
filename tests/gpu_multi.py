"""Multi-GPU check, launched with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_multi.py
(1) observation-sharded DDIM == single-GPU DDIM under replayed noise, for every transport: the in-kernel exchange over
    mapped peer memory ("p2p", GAUSS), the CUDA-graph-captured NCCL all-reduce loop ("nccl") and the eager loop;
(2) sample-sharded DDIM: concatenating the ranks' blocks == the single-GPU run with the same seed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

import d4s
from d4s import _cabi as abi
from d4s import parallel
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

rng = np.random.default_rng(0)
m, d, n, N, steps = 4, 6, 37, 64, 12
onet = orc.random_net(rng, m, d, hidden=(128, 128))
nse = nse_from_oracle(onet, dev)
x = torch.as_tensor(rng.standard_normal((n, d)), dtype=torch.float32, device=dev)
a = 0.2 * rng.standard_normal((n, m, m))
cov = torch.as_tensor(a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m), dtype=torch.float32, device=dev)
z0 = torch.as_tensor(rng.standard_normal((N, m)), dtype=torch.float32)
z = torch.as_tensor(rng.standard_normal((steps, N, m)), dtype=torch.float32)
ok = True
for pkind in ("gaussian", "uniform"):
    if pkind == "uniform":
        lo, hi = torch.full((m,), -3.0, device=dev), torch.full((m,), 3.0, device=dev)
        prior, fn = torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
    else:
        mu, pc = torch.zeros(m, device=dev), torch.eye(m, device=dev)
        prior, fn = torch.distributions.MultivariateNormal(mu, pc), d4s.get_vpdiff_gaussian_score(mu, pc, nse)
    for mode in ("GAUSS", "JAC"):
        kw = dict(steps=steps, eta=0.8, prior=prior, prior_type=pkind, prior_score_fn=fn, dist_cov_est=cov, cov_mode=mode)
        abi.set_option("path", 1)  # per-step fp32 kernels as the single-GPU reference
        ref = nse.ddim((N,), x, _noise=(z0, z), **kw)
        abi.set_option("path", 0)
        ref_tc = nse.ddim((N,), x, _noise=(z0, z), **kw)  # (GAUSS: fused tcgen05 trajectory kernel)
        for exch in (["p2p"] if mode == "GAUSS" else []) + ["nccl", "nccl_eager"]:
            out = parallel.ddim_obs_sharded(nse, (N,), x, _noise=(z0, z), exchange=exch, **kw)
            torch.cuda.synchronize()
            scale = max(1.0, float(ref.abs().max()))
            err = float((out - ref).abs().max()) / scale
            err_tc = float((out - ref_tc).abs().max()) / scale
            tol = 1e-4 if mode == "GAUSS" else 2e-2
            if rank == 0:
                print(f"obs-sharded {pkind} {mode} [{exch}]: world={world} max rel diff vs 1 GPU = {err:.2e} "
                      f"(vs the tcgen05 run {err_tc:.2e}; tol {tol})")
            ok &= err < tol
            full = out.clone()
            dist.broadcast(full, 0)
            ok &= bool(torch.equal(full, out))  # every rank holds the same theta (identical Philox / injected noise)

# a taller problem with observation chunks inside each rank's slice and Philox noise, p2p twice on the same buffers
n2, N2 = 700, 333
x2 = torch.as_tensor(rng.standard_normal((n2, d)), dtype=torch.float32, device=dev)
a2 = 0.2 * rng.standard_normal((n2, m, m))
cov2 = torch.as_tensor(a2 @ a2.transpose(0, 2, 1) + 0.05 * np.eye(m), dtype=torch.float32, device=dev)
mu, pc = torch.zeros(m, device=dev), torch.eye(m, device=dev)
kw2 = dict(steps=25, eta=1.0, prior=torch.distributions.MultivariateNormal(mu, pc), prior_type="gaussian",
           prior_score_fn=d4s.get_vpdiff_gaussian_score(mu, pc, nse), dist_cov_est=cov2, cov_mode="GAUSS")
ref2 = nse.ddim((N2,), x2, _seed=5, **kw2)
for rep in range(2):
    out2 = parallel.ddim_obs_sharded(nse, (N2,), x2, _seed=5, exchange="p2p", **kw2)
    torch.cuda.synchronize()
    err = float((out2 - ref2).abs().max()) / max(1.0, float(ref2.abs().max()))
    if rank == 0:
        print(f"obs-sharded tall (n={n2}, N={N2}) [p2p, run {rep}]: max rel diff vs 1 GPU = {err:.2e}")
    ok &= err < 1e-4

# sample sharding
lo, hi = torch.full((m,), -3.0, device=dev), torch.full((m,), 3.0, device=dev)
prior, fn = torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
kw = dict(steps=steps, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn, dist_cov_est=cov, cov_mode="GAUSS")
Ntot = 203
off, cnt = parallel.sample_shard(Ntot, rank, world)
mine = nse.ddim((cnt,), x, _seed=99, _sample_offset=off, **kw)
gathered = [torch.empty(parallel.sample_shard(Ntot, r, world)[1], m, device=dev) for r in range(world)]
dist.all_gather(gathered, mine)
full = nse.ddim((Ntot,), x, _seed=99, **kw)
same = bool(torch.equal(torch.cat(gathered), full))
if rank == 0:
    print(f"sample-sharded == single GPU: {same}")
ok &= same
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
passed = bool(int(flag))
if rank == 0:
    print("MULTI_GPU_OK" if passed else "MULTI_GPU_FAIL", flush=True)
parallel.shutdown()
dist.destroy_process_group()
sys.exit(0 if passed else 1)
