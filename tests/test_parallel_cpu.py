"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: sample sharding, the observation partition and the
all-reduce of per-rank partial sums.  The kernels themselves need a GPU; here each rank emulates kernel 1 in numpy
on ITS observation shard from the same tables the kernels read, the partial sums are summed with gloo, and the
reduced result must reproduce the oracle's factorized_score on all observations."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_io as gio
from d4s import _cabi as abi
from d4s import engine, parallel
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc
from test_host_tables import _spec, emulate_finalize, emulate_scores


def test_sample_shard_covers_everything():
    for n, w in [(1000, 8), (1001, 8), (7, 2), (5, 5), (3, 1)]:
        blocks = [parallel.sample_shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (o0, c0), (o1, _) in zip(blocks[:-1], blocks[1:]):
            assert o0 + c0 == o1


def test_obs_partition_and_common_chunk():
    for n, w in [(30, 2), (31, 2), (10000, 8), (65, 2), (129, 2), (7, 7)]:
        parts = parallel.obs_partition(n, w)
        sizes = [s.stop - s.start for s in parts]
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1 and parts[0].start == 0 and parts[-1].stop == n
        for cap in (64, 10, 5):
            oc = parallel.common_obs_chunk(n, w, cap)
            assert oc <= cap
            chunks = {-(-sz // min(oc, sz)) if sz else 0 for sz in sizes}
            # same balancing rule as make_tiling() in csrc/d4s_api.cu
            cs = set()
            for sz in sizes:
                o = min(oc, sz)
                c = -(-sz // o)
                cs.add(c)
            assert len(cs) == 1, (n, w, cap, sizes, oc)
    with pytest.raises(ValueError):
        parallel.obs_partition(3, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = gio.load(name)
        onet = gio.oracle_net(g)
        nse = nse_from_oracle(onet)
        view = engine.NetView(engine._linears(nse.net))
        spec = _spec(g)
        theta = g["theta"].astype(np.float64)
        x = torch.as_tensor(g["x"])
        n, m = x.shape[0], onet.theta_dim
        cov = g["dist_cov_est"].astype(np.float64)
        Q = np.linalg.inv(cov)
        sl = parallel.obs_partition(n, world)[rank]
        t = torch.as_tensor(g["t_grid"])
        sa = engine._scalar64(nse.alpha, t.double()).sqrt()
        comb = [engine.combine_mode(spec, False, n, float(s)) for s in sa]
        rows = engine.schedule_rows(nse, t, None, 0.0, n, comb)  # (1 - n) uses the TOTAL n
        mats = engine.mats_table(rows, spec, False, n, torch.as_tensor(Q.sum(0)), m).numpy()
        tf = engine.time_feat_table(nse, view, t).numpy()
        of = engine.obs_feat_table(nse, view, x, None).numpy()[sl]  # this rank's observations only
        rows = rows.numpy()
        worst = 0.0
        for k in range(len(t)):
            s = emulate_scores(onet, nse, view, theta, x, tf[k], of, rows[k, abi.S_INV_SIGMA])
            part = np.concatenate((s.sum(1), np.einsum("jab,ijb->ia", Q[sl], s)), axis=1)  # [N, 2m] partial sums
            buf = torch.as_tensor(part)
            dist.all_reduce(buf)  # the one exchange step of the path
            tot = buf.numpy()
            out = emulate_finalize(rows[k], mats[k], m, spec, theta, tot[:, :m], tot[:, m:])
            ref = g["f64.GAUSS.factorized_score"][k]
            floor = orc.rel_l2(g["f32.GAUSS.factorized_score"][k], ref)
            worst = max(worst, orc.rel_l2(out, ref) / max(2e-6, 1e-2 * floor))
        q.put((rank, worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["slcp_n6", "lv_n6"])
def test_obs_sharded_partial_sums_allreduce_gloo(name):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(w < 1.0 for _, w in res), res
