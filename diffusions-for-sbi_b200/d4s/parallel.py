"""Multi-GPU execution of the tall-data sampler (one process per GPU, torch.distributed / NCCL for the plumbing).

Two axes:
  * posterior samples shard with NO communication: every rank runs `NSE.ddim` on its own slice of samples with
    `_sample_offset` = first global sample index, so the Philox noise of a sample does not depend on the world
    size (`sample_shard`);
  * for very tall n the OBSERVATIONS shard: rank r evaluates the score network on observations [lo_r, hi_r) only and
    produces the partial sums  sum_j s_ij, sum_j inv(Sigma_j) s_ij  (GAUSS) or  sum_j P_ij s_ij, sum_j P_ij  (JAC);
    the sum over ranks is the one exchange step of the path, after which every rank applies prior + solve + DDIM
    update redundantly with identical Philox noise -- no broadcast of theta (`ddim_obs_sharded`).  (1 - n) and
    Lambda use the TOTAL number of observations.  Two transports:
      - "p2p" (default for GAUSS-type trajectories): the exchange happens INSIDE the fused tcgen05 trajectory kernel --
        CTAs store their sums into every peer's receive buffer through mapped peer memory (NVLink / NVSwitch), raise a
        flag, wait for the peers' flags and add the contributions in rank order.  One launch per trajectory per rank,
        no collective call, theta never leaves shared memory (`PeerExchange`, d4s_ddim_run_sharded);
      - "nccl": kernel 1 (export mode) -> `ncclAllReduce` of the [N, C, W] buffer -> kernel 3/4, every step; the whole
        loop is captured in one CUDA graph (JAC steps, nets the trajectory kernel does not take, or no peer access).
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


class PeerExchange:
    """Receive buffers of the in-kernel exchange (include/d4s.h, D4sExchange): one device allocation per rank, mapped
    into every peer of `group` with CUDA IPC.  Buffers grow on demand (collectively); the sequence counter advances by
    `steps` per trajectory on every rank alike."""

    _instances: dict = {}

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > abi.MAX_RANKS:
            raise abi.D4sError(f"in-kernel exchange supports up to {abi.MAX_RANKS} ranks (one NVSwitch box)")
        self.data_bytes = self.flag_bytes = 0
        self.local = [None, None]  # (data, flags) device pointers of this rank
        self.peers = [[None] * self.world, [None] * self.world]
        self.seq = 0

    @classmethod
    def get(cls, group=None) -> "PeerExchange":
        key = id(group) if group is not None else None
        if key not in cls._instances:
            cls._instances[key] = cls(group)
        return cls._instances[key]

    def _release(self):
        lib = abi.load()
        for kind in (0, 1):
            for r, ptr in enumerate(self.peers[kind]):
                if ptr is not None and r != self.rank:
                    lib.d4s_xchg_close(C.c_void_p(ptr))
            if self.local[kind] is not None:
                lib.d4s_xchg_free(C.c_void_p(self.local[kind]))
            self.local[kind] = None
            self.peers[kind] = [None] * self.world

    def ensure(self, data_floats: int, flag_words: int, device: torch.device):
        """Collective: (re)allocates and re-maps the buffers when the problem needs more room than they have."""
        need_d, need_f = 4 * int(data_floats), 4 * int(flag_words)
        if need_d <= self.data_bytes and need_f <= self.flag_bytes:
            return
        lib = abi.require_cuda()
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)  # nobody may still be inside a kernel that uses the old mapping
        self._release()
        self.data_bytes, self.flag_bytes = max(need_d, self.data_bytes), max(need_f, self.flag_bytes)
        handles = []
        with torch.cuda.device(device):
            for kind, nbytes in ((0, self.data_bytes), (1, self.flag_bytes)):
                ptr, h = C.c_void_p(), C.create_string_buffer(abi.IPC_HANDLE_BYTES)
                abi.check(lib.d4s_xchg_alloc(nbytes, C.byref(ptr), h))
                self.local[kind] = ptr.value
                handles.append(bytes(h.raw))
            gathered = [None] * self.world
            dist.all_gather_object(gathered, handles, group=self.group)
            for r in range(self.world):
                for kind in (0, 1):
                    if r == self.rank:
                        self.peers[kind][r] = self.local[kind]
                    else:
                        ptr = C.c_void_p()
                        abi.check(lib.d4s_xchg_open(gathered[r][kind], C.byref(ptr)))
                        self.peers[kind][r] = ptr.value
        self.seq = 0  # fresh (zeroed) flags on every rank
        dist.barrier(group=self.group)

    @classmethod
    def shutdown(cls):
        """Collective: unmaps the peers' buffers and frees the local ones.  Call it before
        `dist.destroy_process_group()` (a process must not exit while a peer still has its memory mapped and in use)."""
        for px in list(cls._instances.values()):
            if px.local[0] is not None:
                torch.cuda.synchronize()
                dist.barrier(group=px.group)
                px._release()
                px.data_bytes = px.flag_bytes = 0
        cls._instances.clear()

    def struct(self, steps: int) -> "abi.Exchange":
        xg = abi.Exchange()
        xg.world, xg.rank, xg.seq_base = self.world, self.rank, self.seq & 0xFFFFFFFF
        for r in range(self.world):
            xg.data[r], xg.flags[r] = self.peers[0][r], self.peers[1][r]
        self.seq += steps
        return xg


def _sharded_setup(nse, shape, x, steps, eta, theta_clipping_range, group, _seed, _noise, kwargs, oc_pairs):
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    opts, _, _, _ = nse._split_kwargs(dict(kwargs))
    N, m = int(shape[0]), nse.zeros.shape[0]
    n = x.shape[0]
    sl = obs_partition(n, world)[rank]
    grid = engine.time_grid(steps)
    t, t_prev = grid[:-1], grid[:-1].to(torch.float64) - 1.0 / steps
    prior = engine.prior_spec(opts["prior_type"], opts["prior"], opts["prior_score_fn"], m)
    oc = common_obs_chunk(n, world, oc_pairs) if oc_pairs else 0
    setup = engine.build_setup(nse, x, N, t, t_prev, float(eta), prior, opts["cov_mode"], opts["dist_cov_est"],
                               clip=theta_clipping_range, seed=_seed, obs_slice=sl, obs_chunk=oc)
    if _noise is not None:
        z0, z = _noise
        theta = z0.to(x.device, torch.float32).contiguous().clone()
        z = z.to(x.device, torch.float32).contiguous()
    else:
        z = None
        theta = engine.philox_normal(N, m, _seed, -1, 0, x.device)
    return setup, theta, z, opts


def p2p_eligible(nse, cov_mode: str) -> bool:
    """The in-kernel exchange rides on the fused tcgen05 trajectory kernel: GAUSS, plain Linear/ReLU net (>= 2 hidden
    layers), no theta-embedding wider than the kernel's buffers (same rule as tc_eligible() in csrc/d4s_api.cu)."""
    if cov_mode != "GAUSS":
        return False
    d = engine.net_description(nse)
    if d["out_act"] != 0 or d["affine"] is not None or float(d["out_scale"]) != 1.0:
        return False
    if len(engine._linears(d["mlp"])) < 3:
        return False
    if d["emb_theta"] is not None and any(max(l.in_features, l.out_features) > 160 for l in engine._linears(d["emb_theta"])):
        return False
    return True


def ddim_obs_sharded(nse, shape, x: torch.Tensor, steps: int = 1000, eta: float = 1.0,
                     theta_clipping_range=(None, None), group: Optional["dist.ProcessGroup"] = None, _seed: int = 0,
                     _noise=None, exchange: str = "auto", **kwargs) -> torch.Tensor:
    """`NSE.ddim` with the observation axis sharded over the ranks of `group`; returns the full theta[N, m] on every
    rank (bit-identical across ranks).  `x` and `dist_cov_est` hold ALL observations on every rank (they are tiny);
    only the score-network work and the partial sums are sharded.  `exchange`: "p2p" (in-kernel, peer memory),
    "nccl" (per-step all-reduce, CUDA-graph captured), "nccl_eager" (same without the graph) or "auto"."""
    nse._require_cuda(x)
    cov_mode = kwargs.get("cov_mode", "GAUSS")
    if exchange == "auto":
        exchange = "p2p" if p2p_eligible(nse, cov_mode) else "nccl"
    if exchange == "p2p":
        return _ddim_obs_sharded_p2p(nse, shape, x, steps, eta, theta_clipping_range, group, _seed, _noise, kwargs)
    if exchange in ("nccl", "nccl_eager"):
        return _ddim_obs_sharded_nccl(nse, shape, x, steps, eta, theta_clipping_range, group, _seed, _noise, kwargs,
                                      graph=exchange == "nccl")
    raise ValueError(f"exchange={exchange!r}: expected 'auto', 'p2p', 'nccl' or 'nccl_eager'")


class P2PTrajectory:
    """One observation-sharded trajectory on the in-kernel exchange, split into the host-side preparation (tables of this
    rank's observation slice, exchange buffers; collective) and the launch (one kernel per rank, enqueue only), so that
    callers can time the device part alone (bench.py) or re-launch from a fresh theta."""

    def __init__(self, nse, shape, x, steps=1000, eta=1.0, theta_clipping_range=(None, None), group=None, _seed=0,
                 _noise=None, **kwargs):
        lib = abi.require_cuda()
        self.setup, self.theta0, self.z, _ = _sharded_setup(nse, shape, x, steps, eta, theta_clipping_range, group,
                                                            _seed, _noise, kwargs, 0)
        self.px = PeerExchange.get(group)
        self.device, self.steps = x.device, steps
        d_fl, f_w = C.c_int64(), C.c_int64()
        abi.check(lib.d4s_xchg_layout(self.setup.pn.handle, C.byref(self.setup.prob), self.px.world, C.byref(d_fl),
                                      C.byref(f_w)))
        self.px.ensure(d_fl.value, f_w.value, x.device)
        self.exchanged_bytes_per_step = int(shape[0]) * 2 * nse.zeros.shape[0] * 4  # one rank's [N, 2m] contribution

    def launch(self, theta: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Enqueues the trajectory kernel of this rank on the current stream; every rank of the group must call it."""
        lib = abi.load()
        theta = self.theta0.clone() if theta is None else theta
        st = self.setup
        xg = self.px.struct(self.steps)
        with torch.cuda.device(self.device):
            abi.check(lib.d4s_ddim_run_sharded(st.pn.handle, C.byref(st.prob), st.n_steps, engine._ptr(st.sched),
                                               engine._ptr(st.time_feat), engine._ptr(st.mats), engine._ptr(self.z),
                                               engine._ptr(theta), C.byref(xg), engine._stream(self.device)))
        return theta


def _ddim_obs_sharded_p2p(nse, shape, x, steps, eta, clip, group, _seed, _noise, kwargs):
    traj = P2PTrajectory(nse, shape, x, steps, eta, clip, group, _seed, _noise, **kwargs)
    theta = traj.launch(traj.theta0)
    theta._d4s_keep = traj  # the kernel reads the tables after this call returns (stream order)
    return theta


class NcclTrajectory:
    """The same trajectory with a collective per step: tcgen05 / SIMT kernel 1 in export mode -> `ncclAllReduce` (sum, fp32)
    of the [N, C, W] partial sums -> kernel 3/4, for every step; with `graph=True` the whole loop is captured once in
    ONE CUDA graph and replayed (no per-step launch latency on the host)."""

    def __init__(self, nse, shape, x, steps=1000, eta=1.0, theta_clipping_range=(None, None), group=None, _seed=0,
                 _noise=None, graph=True, **kwargs):
        self.lib = abi.require_cuda()
        m = nse.zeros.shape[0]
        jac = kwargs.get("cov_mode", "GAUSS") == "JAC"
        self.setup, self.theta0, self.z, _ = _sharded_setup(nse, shape, x, steps, eta, theta_clipping_range, group,
                                                            _seed, _noise, kwargs, 64 // ((1 + m) if jac else 1))
        self.group, self.steps, self.device = group, steps, x.device
        self.theta = self.theta0.clone()
        sizes = torch.tensor([self.setup.prob.workspace_floats], device=x.device, dtype=torch.int64)
        lo, hi = sizes.clone(), sizes.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        if int(lo) != int(hi):
            raise abi.D4sError("observation shards disagree on the partial-sum layout")
        self.graph = None
        self.exchanged_bytes_per_step = 0
        if graph:
            torch.cuda.synchronize(x.device)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    self._body()
            torch.cuda.current_stream(x.device).wait_stream(side)
            self.graph = g

    def _body(self):
        st, ws = self.setup, self.setup.keep[4]
        for k in range(self.steps):
            st.prob.cov_mode = int(st.step_modes[k]) if st.step_modes is not None else st.prob.cov_mode
            engine.score_partials(st, self.theta, k)
            n_floats = self.lib.d4s_partials_floats(st.pn.handle, C.byref(st.prob))
            self.exchanged_bytes_per_step = max(self.exchanged_bytes_per_step, 4 * int(n_floats))
            dist.all_reduce(ws[:n_floats], group=self.group)  # NCCL sum of the [N, C, W] partial sums
            engine.finalize(st, self.theta, k, noise_k=None if self.z is None else self.z[k], update=True, want_score=False)

    def close(self):
        """Destroys the captured graph.  A CUDA graph that holds NCCL kernels keeps the communicator alive:
        `dist.destroy_process_group()` waits for it (observed: the process never exits), so release it first."""
        if self.graph is not None:
            torch.cuda.synchronize(self.device)
            self.graph.reset()
            self.graph = None

    def launch(self) -> torch.Tensor:
        """Runs the trajectory from theta0 (every rank must call it); returns the internal theta buffer."""
        self.theta.copy_(self.theta0)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.theta


def _ddim_obs_sharded_nccl(nse, shape, x, steps, eta, clip, group, _seed, _noise, kwargs, graph=True):
    traj = NcclTrajectory(nse, shape, x, steps, eta, clip, group, _seed, _noise, graph=graph, **kwargs)
    theta = traj.launch().clone()
    traj.close()  # (synchronises; the graph must not outlive the process group)
    return theta


def shutdown():
    """Releases everything that must not outlive the process group (exchange buffers mapped by the peers)."""
    PeerExchange.shutdown()
