"""Multi-GPU execution of the tall-data sampler (one process per GPU, torch.distributed / NCCL for the plumbing).

Two axes:
  * posterior samples shard with NO communication: every rank runs `NSE.ddim` on its own slice of samples with
    `_sample_offset` = first global sample index, so the Philox noise of a sample does not depend on the world
    size (`sample_shard`);
  * for very tall n the OBSERVATIONS shard: rank r evaluates the score network on observations [lo_r, hi_r) only and
    produces the partial sums  sum_j s_ij, sum_j inv(Sigma_j) s_ij  (GAUSS) or  sum_j P_ij s_ij, sum_j P_ij  (JAC);
    one all-reduce (NCCL over NVLink) of that [N, C, W] buffer per DDIM step completes the sum, after which every
    rank applies prior + solve + DDIM update redundantly with identical Philox noise -- no broadcast of theta
    (`ddim_obs_sharded`).  (1 - n) and Lambda use the TOTAL number of observations.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _cabi as abi
from . import engine


def sample_shard(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """(offset, count) of the contiguous block of posterior samples owned by `rank`."""
    base, rem = divmod(n_samples, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def obs_partition(n_obs: int, world: int) -> List[slice]:
    """Balanced contiguous partition of the observations; sizes differ by at most one."""
    if world > n_obs:
        raise ValueError(f"cannot shard {n_obs} observations over {world} ranks")
    out, lo = [], 0
    base, rem = divmod(n_obs, world)
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append(slice(lo, hi))
        lo = hi
    return out


def common_obs_chunk(n_obs: int, world: int, pairs_per_tile: int) -> int:
    """Observation-chunk size that gives EVERY rank the same number of chunks C, so that the per-chunk partial-sum
    buffers [N, C, W] of all ranks line up element by element for the all-reduce."""
    n_max = -(-n_obs // world)
    c_req = -(-n_max // pairs_per_tile)
    return -(-n_max // c_req)


def ddim_obs_sharded(nse, shape, x: torch.Tensor, steps: int = 1000, eta: float = 1.0,
                     theta_clipping_range=(None, None), group: Optional["dist.ProcessGroup"] = None, _seed: int = 0,
                     _noise=None, **kwargs) -> torch.Tensor:
    """`NSE.ddim` with the observation axis sharded over the ranks of `group`; returns the full theta[N, m] on every
    rank.  `x` and `dist_cov_est` hold ALL observations on every rank (they are tiny); only the score-network
    work and the partial sums are sharded."""
    nse._require_cuda(x)
    lib = abi.require_cuda()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    opts, _, _, _ = nse._split_kwargs(dict(kwargs))
    N, m = int(shape[0]), nse.zeros.shape[0]
    n = x.shape[0]
    sl = obs_partition(n, world)[rank]
    jac = opts["cov_mode"] == "JAC"
    grid = engine.time_grid(steps)
    t, t_prev = grid[:-1], grid[:-1].to(torch.float64) - 1.0 / steps
    prior = engine.prior_spec(opts["prior_type"], opts["prior"], opts["prior_score_fn"], m)
    oc = common_obs_chunk(n, world, 64 // ((1 + m) if jac else 1))
    setup = engine.build_setup(nse, x, N, t, t_prev, float(eta), prior, opts["cov_mode"], opts["dist_cov_est"],
                               clip=theta_clipping_range, seed=_seed, obs_slice=sl, obs_chunk=oc)
    ws = setup.keep[4]
    sizes = torch.tensor([setup.prob.workspace_floats], device=x.device, dtype=torch.int64)
    lo, hi = sizes.clone(), sizes.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if int(lo) != int(hi):
        raise abi.D4sError("observation shards disagree on the partial-sum layout")
    if _noise is not None:
        z0, z = _noise
        theta = z0.to(x.device, torch.float32).contiguous().clone()
        z = z.to(x.device, torch.float32).contiguous()
    else:
        z = None
        theta = engine.philox_normal(N, m, _seed, -1, 0, x.device)
    for k in range(steps):
        mode = int(setup.step_modes[k]) if setup.step_modes is not None else setup.prob.cov_mode
        setup.prob.cov_mode = mode
        engine.score_partials(setup, theta, k)
        n_floats = lib.d4s_partials_floats(setup.pn.handle, C.byref(setup.prob))
        dist.all_reduce(ws[:n_floats], group=group)  # NCCL sum of the [N, C, W] partial sums
        engine.finalize(setup, theta, k, noise_k=None if z is None else z[k], update=True, want_score=False)
    return theta
