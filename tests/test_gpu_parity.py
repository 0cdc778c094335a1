"""Parity tests proper (-m gpu): every call goes Python host -> ctypes -> libd4s.so -> sm_100a kernels and is
compared with (a) the golden vectors recorded from the unmodified reference and (b) the fp64 numpy oracle.

Tolerances (north_star): per-call factorized_score <= 1e-4 relative L2 vs the fp64 reference, widened to
3x the reference's OWN fp32-vs-fp64 error where that is larger (SURVEY.md Finding B); full GAUSS trajectories
<= 1e-3 max-abs under replayed noise."""
import os

import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import d4s_oracle as orc

pytestmark = pytest.mark.gpu

CASES = [c for c in gio.names() if not c.startswith(("langevin", "toy", "cfg", "contexts", "euler", "jacdist"))]
REL_TOL = 1e-4
TRAJ_TOL = 1e-3
# The tensor-core JAC kernel (d4s_jac_tc.cu: fp16x3 products, cross terms accumulated before the main terms, accumulator
# debias) is the default JAC path since round 2 and is held to the same bound as the fp32 SIMT kernel.


def _score_tol(mode, kernel_path, floor):
    return max(REL_TOL, 3 * floor)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _setup(name, dev):
    import d4s  # noqa: F401
    from d4s_helpers import nse_from_oracle, prior_objects
    g = gio.load(name)
    onet = gio.oracle_net(g)
    nse = nse_from_oracle(onet, dev)
    prior, fn, ptype = prior_objects(g, nse, dev)
    kw = {}
    if "kw.n" in g:
        kw["n"] = torch.as_tensor(g["kw.n"], device=dev)
    return g, onet, nse, prior, fn, ptype, kw


def _t(v, dev):
    return torch.as_tensor(np.asarray(v, np.float32), device=dev)


@pytest.mark.parametrize("name", CASES)
def test_scores_on_grid(name, dev):
    """kernel 1 with the materialising output on == tweedies_approximation(...)[2] of the reference."""
    import d4s
    g, onet, nse, prior, fn, ptype, kw = _setup(name, dev)
    for k, t in enumerate(g["t_grid"]):
        prec, mean, sc = d4s.tweedies_approximation(x=_t(g["x"], dev), theta=_t(g["theta"], dev), t=torch.tensor(t),
                                                    nse=nse, dist_cov_est=_t(g["dist_cov_est"], dev), mode="GAUSS",
                                                    n=kw.get("n"))
        assert orc.rel_l2(sc.cpu().numpy(), g["f64.scores"][k]) < 2e-5, (name, t)
        assert orc.rel_l2(prec[:2].cpu().numpy(), g["f64.GAUSS.prec"][k]) < 1e-5, (name, t)


@pytest.mark.parametrize("name", [c for c in CASES if "pf" not in c])
def test_jac_precisions(name, dev):
    """kernel 2 (forward-mode Jacobian + inverse) == inv(sigma^2/alpha (I + sigma^2 jacrev(score)))."""
    import d4s
    g, onet, nse, prior, fn, ptype, kw = _setup(name, dev)
    theta = _t(g["theta"], dev)
    for k, t in enumerate(g["t_grid"]):
        prec, _, sc = d4s.tweedies_approximation(x=_t(g["x"], dev), theta=theta, t=torch.tensor(t), nse=nse, mode="JAC")
        ref = g["f64.JAC.prec"][k]
        floor = orc.rel_l2(g["f32.JAC.prec"][k], ref)
        assert orc.rel_l2(prec[:2].cpu().numpy(), ref) < max(REL_TOL, 3 * floor), (name, t)
        assert orc.rel_l2(sc.cpu().numpy(), g["f64.scores"][k]) < 2e-5


@pytest.fixture(params=[0, 1, 2], ids=["auto", "simt", "simtjac"])
def kernel_path(request):
    """0 = automatic (tcgen05 kernels where eligible: bf16x3 trajectory kernel for GAUSS, fp16x3 kernel for JAC steps),
    1 = fp32 SIMT kernels only, 2 = automatic with the JAC steps on the fp32 SIMT kernel (jac_tc = 0; only meaningful for JAC
    tests: others skip it)."""
    from d4s import _cabi as abi
    if request.param == 2:
        if request.node.callspec.params.get("mode") != "JAC":
            pytest.skip("simtjac only differs from auto on JAC steps")
        abi.set_option("jac_tc", 0)
    abi.set_option("path", 1 if request.param == 1 else 0)
    yield request.param
    abi.set_option("path", 0)
    abi.set_option("jac_tc", 1)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["GAUSS", "JAC"])
def test_factorized_score(name, mode, dev, kernel_path):
    g, onet, nse, prior, fn, ptype, kw = _setup(name, dev)
    if f"f64.{mode}.factorized_score" not in g:
        pytest.skip("PF+JAC has no reference behaviour (SURVEY.md 8(a) a11)")
    worst = 0.0
    for k, t in enumerate(g["t_grid"]):
        out = nse.factorized_score(_t(g["theta"], dev), _t(g["x"], dev), torch.tensor(t), fn, prior,
                                   dist_cov_est=_t(g["dist_cov_est"], dev), cov_mode=mode, prior_type=ptype, **kw)
        ref = g[f"f64.{mode}.factorized_score"][k]
        floor = orc.rel_l2(g[f"f32.{mode}.factorized_score"][k], ref)
        err = orc.rel_l2(out.cpu().numpy(), ref)
        worst = max(worst, err / _score_tol(mode, kernel_path, floor))
        assert err < _score_tol(mode, kernel_path, floor), (name, mode, float(t), err, floor)
    print(f"{name} {mode}: worst err/tol = {worst:.3f}")


@pytest.mark.parametrize("name", CASES)
def test_prior_score_and_precision(name, dev):
    import d4s
    g, onet, nse, prior, fn, ptype, kw = _setup(name, dev)
    theta = _t(g["theta"], dev)
    for k, t in enumerate(g["t_grid"]):
        ref = g["f64.prior_score"][k]
        floor = orc.rel_l2(g["f32.prior_score"][k], ref)
        assert orc.rel_l2(fn(theta, torch.tensor(t)).cpu().numpy(), ref) < max(1e-5, 3 * floor), (name, t)
        if np.isfinite(g["f64.prior_prec"][k]).all():
            prec, _, _ = d4s.tweedies_approximation_prior(theta, torch.tensor(t), fn, nse)
            assert orc.rel_l2(prec.cpu().numpy(), g["f64.prior_prec"][k]) < 1e-4, (name, t)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("graph", [0, 1])
def test_trajectory_replayed_noise(name, graph, dev, monkeypatch):
    monkeypatch.setenv("D4S_CUDA_GRAPH", str(graph))
    g, onet, nse, prior, fn, ptype, kw = _setup(name, dev)
    steps, eta = int(g["traj_steps"]), float(g["traj_eta"])
    clip = tuple(float(v) for v in g["traj_clip"]) if "traj_clip" in g else (None, None)
    x = _t(g["traj_x"] if "traj_x" in g else g["x"], dev)
    z0, z = torch.as_tensor(g["traj_z0"]), torch.as_tensor(g["traj_z"])
    for mode in ("GAUSS", "JAC"):
        if f"f64.{mode}.traj" not in g:
            continue
        out = nse.ddim((z0.shape[0],), x, steps=steps, eta=eta, prior=prior, prior_type=ptype, prior_score_fn=fn,
                       dist_cov_est=_t(g["dist_cov_est"], dev), cov_mode=mode, theta_clipping_range=clip,
                       _noise=(z0, z))
        ref = g[f"f64.{mode}.traj"]
        err = np.abs(out.cpu().numpy() - ref).max()
        ref32 = np.abs(g[f"f32.{mode}.traj"] - ref).max()
        print(f"{name} {mode} graph={graph}: max|d theta| = {err:.2e} (reference fp32 vs fp64: {ref32:.2e})")
        tol = TRAJ_TOL if mode == "GAUSS" else max(TRAJ_TOL, 3 * ref32)
        assert err < tol, (name, mode, err)


# ---- oracle comparisons at shapes the fixtures do not cover -------------------------------------------------
SHAPES = [
    # m, d, hidden, n, N, prior
    (5, 8, (256, 256, 256), 30, 1000, "uniform"),   # BASELINE config 2 at full size
    (4, 20, (256, 256, 256), 30, 257, "gaussian"),  # LV shape, ragged N
    (2, 10, (256, 256, 256), 1, 33, "gaussian"),    # single observation through the tall API pieces
    (10, 10, (256, 256, 256), 70, 19, "gaussian"),  # two observation chunks
    (10, 10, (64, 128), 200, 7, "uniform"),         # many chunks, narrow ragged widths
    (1, 3, (50,), 5, 64, "gaussian"),               # m = 1, one hidden layer, padded width
    (3, 4, (96, 256, 64, 128), 9, 40, "uniform"),   # 5 linear layers
]


def _problem(m, d, hidden, n, N, pkind, seed=0):
    rng = np.random.default_rng(seed)
    onet = orc.random_net(rng, m, d, hidden=hidden)
    x = rng.standard_normal((n, d))
    a = 0.2 * rng.standard_normal((n, m, m))
    cov = a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)
    theta = rng.standard_normal((N, m))
    if pkind == "uniform":
        oprior = orc.UniformPrior(np.full(m, -3.46), np.full(m, 3.46))
    else:
        b = 0.3 * rng.standard_normal((m, m))
        oprior = orc.GaussianPrior(0.1 * rng.standard_normal(m), b @ b.T + np.eye(m))
    return onet, x, cov, theta, oprior


def _d4s_prior(oprior, nse, dev):
    import d4s
    if oprior.kind == "uniform":
        lo, hi = _t(oprior.low, dev), _t(oprior.high, dev)
        return torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
    mu, cov = _t(oprior.mean, dev), _t(oprior.cov, dev)
    return torch.distributions.MultivariateNormal(mu, cov), d4s.get_vpdiff_gaussian_score(mu, cov, nse)


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"m{s[0]}_n{s[3]}_N{s[4]}_{s[5]}")
@pytest.mark.parametrize("mode", ["GAUSS", "JAC"])
def test_factorized_score_vs_oracle(shape, mode, dev, kernel_path):
    from d4s_helpers import nse_from_oracle
    m, d, hidden, n, N, pkind = shape
    if mode == "JAC" and N * n > 4000:
        N = 4000 // n  # keep the fp64 oracle's reverse-mode Jacobian in seconds
    onet, x, cov, theta, oprior = _problem(m, d, hidden, n, N, pkind)
    theta = theta[:N]
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    for t in (0.9, 0.4, 0.2, 0.05):
        t32 = np.float32(t)
        ref = orc.factorized_score(onet, theta, x, np.float64(t32), oprior, cov, mode, pkind)
        ref32 = orc.factorized_score(onet.astype(np.float32), theta.astype(np.float32), x.astype(np.float32), t32,
                                     oprior, cov.astype(np.float32), mode, pkind)
        out = nse.factorized_score(_t(theta, dev), _t(x, dev), torch.tensor(t32), fn, prior,
                                   dist_cov_est=_t(cov, dev), cov_mode=mode, prior_type=pkind)
        err, floor = orc.rel_l2(out.cpu().numpy(), ref), orc.rel_l2(ref32, ref)
        assert err < _score_tol(mode, kernel_path, floor), (shape, mode, t, err, floor)


def test_paired_and_single_observation_ddim(dev):
    """nse.py:253-254: x[N, d] paired with the samples, x[d] and x[1, d] all use the plain score."""
    from d4s_helpers import nse_from_oracle
    rng = np.random.default_rng(3)
    onet = orc.random_net(rng, 3, 4, hidden=(64, 64))
    nse = nse_from_oracle(onet, dev)
    N, steps = 70, 12
    z0, z = rng.standard_normal((N, 3)), rng.standard_normal((steps, N, 3))
    noise = (torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32))
    xp = rng.standard_normal((N, 4))
    ref = orc.ddim_single(onet, xp, z0, z, steps, 0.5)
    out = nse.ddim((N,), _t(xp, dev), steps=steps, eta=0.5, _noise=noise)
    # a random-init net gives no contraction: |theta| grows to ~1e3, so the tolerance is relative
    assert np.abs(out.cpu().numpy() - ref).max() < TRAJ_TOL * max(1.0, np.abs(ref).max()) * 1e-1
    x1 = rng.standard_normal(4)
    ref = orc.ddim_single(onet, x1, z0, z, steps, 0.5, clip=(-1.0, 1.0))
    for xx in (x1, x1[None]):
        out = nse.ddim((N,), _t(xx, dev), steps=steps, eta=0.5, theta_clipping_range=(-1.0, 1.0), _noise=noise)
        assert np.abs(out.cpu().numpy() - ref).max() < TRAJ_TOL


def test_philox_noise_is_seeded_and_shard_invariant(dev):
    from d4s import engine
    a = engine.philox_normal(4096, 5, 1234, 7, 0, dev)
    b = engine.philox_normal(4096, 5, 1234, 7, 0, dev)
    assert torch.equal(a, b)
    lo = engine.philox_normal(1000, 5, 1234, 7, 0, dev)
    hi = engine.philox_normal(3096, 5, 1234, 7, 1000, dev)
    assert torch.equal(torch.cat((lo, hi)), a)  # sample sharding does not change any sample's noise
    c = engine.philox_normal(4096, 5, 1235, 7, 0, dev)
    assert not torch.equal(a, c)
    big = engine.philox_normal(200000, 4, 99, 0, 0, dev).double()
    assert abs(float(big.mean())) < 5e-3 and abs(float(big.var()) - 1) < 1e-2
    assert abs(float((big ** 4).mean()) - 3) < 0.1
    assert abs(float((big[:, 0] * big[:, 1]).mean())) < 5e-3


def test_ddim_seeded_runs_are_reproducible_and_shardable(dev):
    from d4s_helpers import nse_from_oracle
    onet, x, cov, theta, oprior = _problem(5, 8, (256, 256, 256), 30, 8, "uniform")
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    kw = dict(steps=30, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn,
              dist_cov_est=_t(cov, dev), cov_mode="GAUSS")
    torch.manual_seed(42)
    a = nse.ddim((512,), _t(x, dev), **kw)
    torch.manual_seed(42)
    b = nse.ddim((512,), _t(x, dev), **kw)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    lo = nse.ddim((200,), _t(x, dev), _seed=77, _sample_offset=0, **kw)
    hi = nse.ddim((312,), _t(x, dev), _seed=77, _sample_offset=200, **kw)
    full = nse.ddim((512,), _t(x, dev), _seed=77, **kw)
    assert torch.equal(torch.cat((lo, hi)), full)  # posterior samples shard over ranks with no communication
    # shards whose boundaries are NOT multiples of the samples-per-tile: a sample lands in a different slot of its tile than in
    # the full run, and must still come out bit-identical (8 ranks, 203 samples: the split tests/gpu_multi.py makes)
    from d4s import parallel
    for n_obs, pk in ((30, "uniform"), (22, "uniform"), (7, "gaussian")):
        onet2, x2, cov2, _, op2 = _problem(5, 8, (256, 256, 256), n_obs, 8, pk, seed=n_obs)
        nse2 = nse_from_oracle(onet2, dev)
        prior2, fn2 = _d4s_prior(op2, nse2, dev)
        kw2 = dict(steps=12, eta=1.0, prior=prior2, prior_type=pk, prior_score_fn=fn2, dist_cov_est=_t(cov2, dev), cov_mode="GAUSS")
        full = nse2.ddim((203,), _t(x2, dev), _seed=99, **kw2)
        parts = []
        for r in range(8):
            off, cnt = parallel.sample_shard(203, r, 8)
            parts.append(nse2.ddim((cnt,), _t(x2, dev), _seed=99, _sample_offset=off, **kw2))
        assert torch.equal(torch.cat(parts), full), (n_obs, pk, float((torch.cat(parts) - full).abs().max()))


def test_batched_inverse(dev):
    from d4s import engine
    rng = np.random.default_rng(0)
    for m in range(1, 11):
        a = rng.standard_normal((37, m, m)) + 2 * np.eye(m)
        out = engine.batched_inverse(_t(a, dev)).cpu().numpy()
        assert orc.rel_l2(out, np.linalg.inv(a.astype(np.float32).astype(np.float64))) < 1e-5


def test_full_size_trajectory_properties(dev):
    """BASELINE config 2 at full size (N=1000, n=30, 1000 steps): finite, inside the clip box, and equal to the
    observation-chunked run (the reduction over observations is exact up to fp32 summation order)."""
    from d4s import engine
    from d4s_helpers import nse_from_oracle
    onet, x, cov, theta, oprior = _problem(5, 8, (256, 256, 256), 30, 8, "uniform")
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    kw = dict(steps=1000, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn,
              dist_cov_est=_t(cov, dev), cov_mode="GAUSS", theta_clipping_range=(-3.0, 3.0), _seed=5)
    out = nse.ddim((1000,), _t(x, dev), **kw)
    assert out.shape == (1000, 5) and torch.isfinite(out).all()
    assert float(out.abs().max()) <= 3.0
    before = engine.launch_count()
    out2 = nse.ddim((1000,), _t(x, dev), **kw)
    assert torch.equal(out, out2)
    assert engine.launch_count() - before >= 2  # Philox init + the fused tcgen05 trajectory kernel
    # the per-step fp32 SIMT kernels (2 launches per DDIM step) must agree with the fused tensor-core path
    from d4s import _cabi as abi
    abi.set_option("path", 1)
    try:
        before = engine.launch_count()
        out3 = nse.ddim((1000,), _t(x, dev), **kw)
        assert engine.launch_count() - before >= 2 * 1000 + 1
    finally:
        abi.set_option("path", 0)
    # (random-init nets do not contract: the two precisions' trajectories drift apart on some samples; the pathwise
    # 1000-step gate is tests/test_gpu_long_runs.py::test_long_interleaved_trajectory_vs_fp64_oracle on shipped weights)
    frac_close = float(((out3 - out).abs() < 1e-2).float().mean())
    print(f"tcgen05 vs SIMT after 1000 steps: median |d| = {float((out3 - out).abs().median()):.2e}, "
          f"within 1e-2: {frac_close:.3f}")
    assert frac_close > 0.9


@pytest.mark.parametrize("shape", [(5, 8, (256, 256, 256), 30, 1300, "uniform"), (5, 8, (64, 128), 7, 2500, "uniform"),
                                   (4, 6, (128, 256), 12, 1190, "gaussian"), (10, 10, (256, 256), 30, 700, "gaussian"),
                                   # one short tensor-core layer, 8 samples per tile: the shape that exposed a shared-memory
                                   # race between layer 0 of the next item and the scores scratch of the previous one
                                   (3, 8, (64, 64), 9, 1500, "gaussian"),
                                   # fused per-sample sums with padding rows inside a sample's 32-row block: the lane that ends up
                                   # with a result slot of the shuffle reduction is a padding row (n = 17, 18), or not (n = 26)
                                   (5, 8, (256, 256), 17, 600, "uniform"), (4, 6, (128, 256), 18, 700, "gaussian"),
                                   (5, 8, (256, 128), 26, 900, "uniform")],
                         ids=lambda s: f"m{s[0]}_n{s[3]}_N{s[4]}_{s[5]}")
def test_tc_group_schedules_agree(shape, dev):
    """The tcgen05 trajectory kernel interleaves two sample groups per CTA when a CTA owns several (N > 148 groups):
    the interleaved schedule must be bit-identical to running the groups one after the other, and both must follow
    the fp64 oracle under replayed noise."""
    from d4s import _cabi as abi
    from d4s_helpers import nse_from_oracle
    m, d, hidden, n, N, pkind = shape
    onet, x, cov, theta, oprior = _problem(m, d, hidden, n, N, pkind, seed=3)
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    steps = 5
    rng = np.random.default_rng(1)
    z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
    kw = dict(steps=steps, eta=1.0, prior=prior, prior_type=pkind, prior_score_fn=fn, dist_cov_est=_t(cov, dev),
              cov_mode="GAUSS", _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
    abi.set_option("path", 2)
    try:
        out = nse.ddim((N,), _t(x, dev), **kw)
        abi.set_option("tc_pairing", 0)
        out_seq = nse.ddim((N,), _t(x, dev), **kw)
    finally:
        abi.set_option("tc_pairing", 1)
        abi.set_option("path", 0)
    assert torch.equal(out, out_seq)
    abi.set_option("path", 2)
    try:
        for _ in range(3):  # run-to-run determinism of the interleaved schedule
            assert torch.equal(nse.ddim((N,), _t(x, dev), **kw), out)
    finally:
        abi.set_option("path", 0)
    ref = orc.ddim(onet, x, z0, z, steps, 1.0, oprior, cov, "GAUSS", pkind)
    scale = max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(out.cpu().numpy() - ref).max()) / scale
    print(f"{shape}: max |theta - oracle| / scale = {err:.2e}")
    assert err < TRAJ_TOL, err


@pytest.mark.parametrize("shape", [(5, 8, (256, 256, 256), 30, 1000, "uniform"), (5, 8, (256, 256, 256), 30, 80, "uniform"),
                                   (4, 6, (128, 256), 12, 1184, "gaussian"), (2, 3, (64, 64), 9, 1504, "gaussian"),
                                   (5, 8, (256, 128), 26, 904, "uniform")],
                         ids=lambda s: f"m{s[0]}_n{s[3]}_N{s[4]}_{s[5]}")
def test_tc_cta_pairs_agree(shape, dev):
    """`tc_cta_pair` (opt-in): the trajectory kernel launched as clusters of two CTAs that share one cta_group::2 tensor-core
    stream (M = 256 per instruction, each CTA streams half of every weight tile).  Every row still goes through the same
    products in the same order, so the result must be bit-identical to the single-CTA kernel (even numbers of sample groups:
    the host only selects the pair kernel then)."""
    from d4s import _cabi as abi
    from d4s import engine
    from d4s_helpers import nse_from_oracle
    m, d, hidden, n, N, pkind = shape
    onet, x, cov, theta, oprior = _problem(m, d, hidden, n, N, pkind, seed=5)
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    steps = 6
    rng = np.random.default_rng(2)
    z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
    kw = dict(steps=steps, eta=1.0, prior=prior, prior_type=pkind, prior_score_fn=fn, dist_cov_est=_t(cov, dev),
              cov_mode="GAUSS", _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
    abi.set_option("path", 2)
    try:
        abi.set_option("tc_cta_pair", 0)
        out = nse.ddim((N,), _t(x, dev), **kw)
        n0 = engine.launch_count()
        abi.set_option("tc_cta_pair", 1)
        out_pair = nse.ddim((N,), _t(x, dev), **kw)
        assert engine.launch_count() - n0 == 1  # still one launch per trajectory
        for _ in range(2):
            assert torch.equal(nse.ddim((N,), _t(x, dev), **kw), out_pair)
    finally:
        abi.set_option("tc_cta_pair", 0)
        abi.set_option("path", 0)
    assert torch.isfinite(out_pair).all()
    assert torch.equal(out, out_pair)


@pytest.mark.parametrize("shape", [(5, 8, (256, 256, 256), 30, 1000), (4, 6, (128, 256), 7, 333), (2, 3, (64, 64), 40, 150),
                                   (1, 2, (64, 128), 11, 77)], ids=lambda s: f"m{s[0]}_n{s[3]}_N{s[4]}")
def test_jac_cta_pairs_agree(shape, dev):
    """`jac_cta_pair` (opt-in): the JAC step kernel launched as clusters of two CTAs on one cta_group::2 stream; the partial
    sums must be bit-identical to the single-CTA kernel (odd tile counts: the last pair's second CTA runs an empty tile)."""
    import ctypes as C
    from d4s import _cabi as abi
    from d4s import engine
    from d4s_helpers import nse_from_oracle
    m, d, hidden, n, N = shape
    rng = np.random.default_rng(11)
    onet = orc.random_net(rng, m, d, hidden=hidden)
    nse = nse_from_oracle(onet, dev)
    xt = _t(rng.standard_normal((n, d)), dev)
    th = _t(rng.standard_normal((N, m)), dev)
    t = torch.tensor([0.3], dtype=torch.float32)
    outs = []
    try:
        for pair in (0, 1, 1):
            abi.set_option("jac_cta_pair", pair)
            setup = engine.build_setup(nse, xt, N, t, None, 0.0, engine.PriorSpec(abi.PRIOR_NONE), "JAC", None)
            engine.score_partials(setup, th, 0)
            torch.cuda.synchronize()
            nf = abi.load().d4s_partials_floats(setup.pn.handle, C.byref(setup.prob))
            outs.append(setup.keep[4][:nf].clone())
    finally:
        abi.set_option("jac_cta_pair", 0)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("shape", [(5, 8, (256, 256, 256), 30, 1000), (4, 6, (128, 256), 7, 333), (2, 3, (64, 64), 40, 150),
                                   (3, 5, (256, 96, 128), 1, 500), (1, 2, (64, 128), 11, 77)],
                         ids=lambda s: f"m{s[0]}_n{s[3]}_N{s[4]}")
def test_jac_tc_partial_sums(shape, dev):
    """Kernel level: the per-sample sums [sum_j P s | sum_j P] of one JAC step from the tcgen05 score + Jacobian kernel
    against the fp32 SIMT kernel and the fp64 oracle (many tiles per CTA, ragged unit packing, every theta_dim <= 5)."""
    import ctypes as C
    from d4s import _cabi as abi
    from d4s import engine
    from d4s_helpers import nse_from_oracle
    m, d, hidden, n, N = shape
    rng = np.random.default_rng(11)
    onet = orc.random_net(rng, m, d, hidden=hidden)
    nse = nse_from_oracle(onet, dev)
    x, theta = rng.standard_normal((n, d)), rng.standard_normal((N, m))
    W = m + m * m
    t32 = np.float32(0.25)

    def partials(path):
        abi.set_option("path", path)
        abi.set_option("jac_tc", 1)
        try:
            setup = engine.build_setup(nse, _t(x, dev), N, torch.tensor([t32]), None, 0.0, engine.PriorSpec(abi.PRIOR_NONE),
                                       "JAC", None)
            engine.score_partials(setup, _t(theta, dev), 0)
            torch.cuda.synchronize()
            nf = abi.load().d4s_partials_floats(setup.pn.handle, C.byref(setup.prob))
            return setup.keep[4][:nf].clone().reshape(N, -1, W).double().sum(1).cpu().numpy()
        finally:
            abi.set_option("path", 0)

    before = engine.launch_count()
    tc = partials(0)
    assert engine.launch_count() - before == 1  # one tcgen05 launch, no SIMT kernel
    simt = partials(1)
    Nr = min(N, 2000 // n)  # the oracle's reverse-mode Jacobian on a subset keeps the test in seconds
    prec, _, sc = orc.tweedies_approximation(onet, x, theta[:Nr], np.float64(t32), None, "JAC")
    ref = np.concatenate([(prec @ sc[..., None])[..., 0].sum(1), prec.sum(1).reshape(Nr, -1)], axis=1)
    e_tc, e_simt = orc.rel_l2(tc[:Nr], ref), orc.rel_l2(simt[:Nr], ref)
    e_all = orc.rel_l2(tc, simt)
    print(f"{shape}: tc-oracle {e_tc:.2e}, simt-oracle {e_simt:.2e}, tc-simt (all samples) {e_all:.2e}")
    assert e_simt < 1e-5 and e_tc < 1e-4 and e_all < 1e-4


@pytest.mark.parametrize("pkind", ["uniform", "gaussian"])
@pytest.mark.parametrize("mode", ["GAUSS", "JAC"])
def test_ddim_contexts_vs_oracle(pkind, mode, dev):
    """Batched contexts: B independent tall problems in one launch == the reference's
    vmap(sampler) over calibration sets (jrnnm_posterior_estimation.py:286-294), checked context by context against the
    oracle under replayed noise; also equal to B separate `ddim` calls."""
    from d4s_helpers import nse_from_oracle
    m, d, B, n, S, steps = 3, 4, 5, 6, 7, 8
    rng = np.random.default_rng(5)
    onet = orc.random_net(rng, m, d, hidden=(64, 96))
    nse = nse_from_oracle(onet, dev)
    x = 0.5 * rng.standard_normal((B, n, d))
    a = 0.2 * rng.standard_normal((B, n, m, m))
    cov = a @ a.transpose(0, 1, 3, 2) + 0.05 * np.eye(m)
    if pkind == "uniform":
        oprior = orc.UniformPrior(np.full(m, -3.46), np.full(m, 3.46))
    else:
        b = 0.3 * rng.standard_normal((m, m))
        oprior = orc.GaussianPrior(0.1 * rng.standard_normal(m), b @ b.T + np.eye(m))
    prior, fn = _d4s_prior(oprior, nse, dev)
    z0, z = rng.standard_normal((B * S, m)), rng.standard_normal((steps, B * S, m))
    kw = dict(steps=steps, eta=0.8, prior=prior, prior_type=pkind, prior_score_fn=fn, cov_mode=mode)
    out = nse.ddim_contexts((S,), _t(x, dev), dist_cov_est=_t(cov, dev),
                            _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)), **kw)
    assert out.shape == (B, S, m)
    worst = 0.0
    for c in range(B):
        sl = slice(c * S, (c + 1) * S)
        ref = orc.ddim(onet, x[c], z0[sl], z[:, sl], steps, 0.8, oprior, cov[c], mode, pkind)
        scale = max(1.0, float(np.abs(ref).max()))  # a random-init net gives no contraction: relative tolerance
        one = nse.ddim((S,), _t(x[c], dev), dist_cov_est=_t(cov[c], dev),
                       _noise=(torch.as_tensor(z0[sl], dtype=torch.float32), torch.as_tensor(z[:, sl], dtype=torch.float32)),
                       **kw)
        worst = max(worst, float(np.abs(out[c].cpu().numpy() - ref).max()) / scale)
        assert float((out[c] - one).abs().max()) < 1e-4 * scale  # same kernels, per-sample instead of shared Lambda solve
    print(f"ddim_contexts {pkind} {mode}: max |theta - oracle| / scale = {worst:.2e}")
    assert worst < TRAJ_TOL


@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("contexts")])
def test_ddim_contexts_golden(name, dev):
    """`NSE.ddim_contexts` against the unmodified reference run context by context (what its vmap(sampler) computes) on
    the shipped JR-NMM net (theta / x embedding MLPs): tall GAUSS and JAC posteriors and the one-observation-per-context
    pre-pass, fp64 reference trajectories under replayed noise."""
    g, onet, nse, prior, fn, ptype, kw = _setup(name, dev)
    S, steps, eta = int(g["samples_per_context"]), int(g["traj_steps"]), float(g["traj_eta"])
    noise = (torch.as_tensor(g["traj_z0"], dtype=torch.float32), torch.as_tensor(g["traj_z"], dtype=torch.float32))
    x = _t(g["x"], dev)
    for mode in ("GAUSS", "JAC"):
        out = nse.ddim_contexts((S,), x, steps=steps, eta=eta, prior=prior, prior_type=ptype, prior_score_fn=fn,
                                dist_cov_est=_t(g["dist_cov_est"], dev), cov_mode=mode, _noise=noise)
        ref, ref32 = g[f"f64.{mode}.traj"], g[f"f32.{mode}.traj"]
        per_sample = np.abs(out.cpu().numpy() - ref).max(axis=2).ravel()
        err, floor = float(per_sample.max()), float(np.abs(ref32 - ref).max())
        print(f"{name} {mode}: max |theta - reference fp64| = {err:.2e}, median {np.median(per_sample):.2e} (reference fp32: {floor:.2e})")
        tol = max(TRAJ_TOL, 3 * floor)
        if mode == "GAUSS":
            assert err < tol, (mode, err, floor)
        else:
            # JAC trajectories are chaotic through ill-conditioned solves (the reference's own fp32 and fp64 runs diverge,
            # BASELINE.md section 4; per-call parity is the gate): the typical sample must agree, a single one may not
            assert np.median(per_sample) < tol and (per_sample > tol).mean() <= 0.1, (mode, err, floor)
    out = nse.ddim_contexts((S,), x[:, 0], steps=steps, eta=eta, _noise=noise)
    err = float(np.abs(out.cpu().numpy() - g["f64.single.traj"]).max())
    print(f"{name} single observation per context: max |theta - reference fp64| = {err:.2e}")
    assert err < TRAJ_TOL


@pytest.mark.parametrize("seed", range(8))
def test_ddim_contexts_random_shapes(seed, dev):
    """Seeded random (B, n, S, theta_dim, widths, prior, mode): one batched-contexts launch == B separate `ddim` calls."""
    from d4s import _cabi as abi
    from d4s_helpers import nse_from_oracle
    rng = np.random.default_rng(100 + seed)
    m, d = int(rng.integers(1, 8)), int(rng.integers(1, 9))
    hidden = tuple(int(rng.choice([50, 64, 128, 256])) for _ in range(int(rng.integers(1, 4))))
    B, n, S = int(rng.integers(2, 40)), int(rng.choice([2, 3, 7, 30, 70])), int(rng.choice([1, 2, 5, 33]))
    pkind, mode, steps = str(rng.choice(["uniform", "gaussian"])), str(rng.choice(["GAUSS", "JAC"])), 4
    onet = orc.random_net(rng, m, d, hidden=hidden)
    nse = nse_from_oracle(onet, dev)
    x = 0.5 * rng.standard_normal((B, n, d))
    a = 0.2 * rng.standard_normal((B, n, m, m))
    cov = a @ a.transpose(0, 1, 3, 2) + 0.05 * np.eye(m)
    if pkind == "uniform":
        oprior = orc.UniformPrior(np.full(m, -3.46), np.full(m, 3.46))
    else:
        b = 0.3 * rng.standard_normal((m, m))
        oprior = orc.GaussianPrior(0.1 * rng.standard_normal(m), b @ b.T + np.eye(m))
    prior, fn = _d4s_prior(oprior, nse, dev)
    z0, z = rng.standard_normal((B * S, m)), rng.standard_normal((steps, B * S, m))
    kw = dict(steps=steps, eta=0.8, prior=prior, prior_type=pkind, prior_score_fn=fn, cov_mode=mode)
    noise_all = (torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32))
    auto = nse.ddim_contexts((S,), _t(x, dev), dist_cov_est=_t(cov, dev), _noise=noise_all, **kw)  # tcgen05 kernel where eligible
    abi.set_option("path", 1)  # same fp32 kernels on both sides: only the batching differs
    try:
        out = nse.ddim_contexts((S,), _t(x, dev), dist_cov_est=_t(cov, dev), _noise=noise_all, **kw)
        rel_auto = (auto - out).abs().amax(dim=(1, 2)) / out.abs().amax(dim=(1, 2)).clamp_min(1.0)
        assert float(rel_auto.median()) < 1e-4 and float((rel_auto > 1e-2).float().mean()) <= 0.05, (seed, mode, float(rel_auto.max()))
        for c in range(B):
            sl = slice(c * S, (c + 1) * S)
            one = nse.ddim((S,), _t(x[c], dev), dist_cov_est=_t(cov[c], dev),
                           _noise=(torch.as_tensor(z0[sl], dtype=torch.float32), torch.as_tensor(z[:, sl], dtype=torch.float32)), **kw)
            scale = max(1.0, float(one.abs().max()))
            rel = (out[c] - one).abs().max(dim=1).values / scale
            # (per-sample Lambda solve instead of a shared inverse; JAC: see test_tc_per_call_random_shapes)
            assert float(rel.median()) < 1e-4 and float((rel > 1e-2).float().mean()) <= 0.05, (seed, c, mode, pkind, float(rel.max()))
    finally:
        abi.set_option("path", 0)


@pytest.mark.parametrize("case", [(3, 7, 3, 40, "uniform"), (4, 30, 3, 17, "gaussian"), (5, 3, 3, 130, "uniform")],
                         ids=lambda c: f"B{c[0]}_n{c[1]}_nmax{c[2]}_S{c[3]}_{c[4]}")
def test_pf_nse_ddim_contexts(case, dev):
    """`PF_NSE.ddim_contexts`: every context split into its K = ceil(n / n_max) subsets (pf_nse.py:169-197), all contexts in one
    launch == B separate `PF_NSE.ddim` calls under the same noise; K = 1 (n <= n_max) is the plain-score single-subset case."""
    from d4s_helpers import nse_from_oracle
    B, n, n_max, S, pkind = case
    m, d, steps = 4, 5, 5
    rng = np.random.default_rng(31)
    onet = orc.random_net(rng, m, d, hidden=(128, 256), n_max=n_max, pf=True)
    pf = nse_from_oracle(onet, dev)
    K = -(-n // n_max)
    x = 0.5 * rng.standard_normal((B, n, d))
    a = 0.2 * rng.standard_normal((B, K, m, m))
    cov = a @ a.transpose(0, 1, 3, 2) + 0.05 * np.eye(m)
    if pkind == "uniform":
        oprior = orc.UniformPrior(np.full(m, -3.46), np.full(m, 3.46))
    else:
        b = 0.3 * rng.standard_normal((m, m))
        oprior = orc.GaussianPrior(0.1 * rng.standard_normal(m), b @ b.T + np.eye(m))
    prior, fn = _d4s_prior(oprior, pf, dev)
    z0, z = rng.standard_normal((B * S, m)), rng.standard_normal((steps, B * S, m))
    kw = dict(steps=steps, eta=0.8, prior=prior, prior_type=pkind, prior_score_fn=fn, cov_mode="GAUSS")
    out = pf.ddim_contexts((S,), _t(x, dev), dist_cov_est=_t(cov, dev),
                           _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)), **kw)
    assert out.shape == (B, S, m) and torch.isfinite(out).all()
    for c in range(B):
        sl = slice(c * S, (c + 1) * S)
        one = pf.ddim((S,), _t(x[c], dev), dist_cov_est=_t(cov[c], dev),
                      _noise=(torch.as_tensor(z0[sl], dtype=torch.float32), torch.as_tensor(z[:, sl], dtype=torch.float32)), **kw)
        scale = max(1.0, float(one.abs().max()))
        rel = float((out[c] - one).abs().max()) / scale
        assert rel < 1e-4, (case, c, rel)


def test_estimate_dist_cov_prepass(dev):
    """sbibm_posterior_estimation.py:306-314: vmap(ddim) over the observations + torch.cov, as one launch."""
    from d4s_helpers import nse_from_oracle
    m, d, n, S, steps = 3, 4, 9, 300, 10
    rng = np.random.default_rng(6)
    onet = orc.random_net(rng, m, d, hidden=(64, 64))
    nse = nse_from_oracle(onet, dev)
    x = 0.5 * rng.standard_normal((n, d))
    z0, z = rng.standard_normal((n * S, m)), rng.standard_normal((steps, n * S, m))
    noise = (torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32))
    cov = nse.estimate_dist_cov(_t(x, dev), nsamples=S, steps=steps, eta=0.5, _noise=noise)
    assert cov.shape == (n, m, m)
    for j in (0, n - 1):
        sl = slice(j * S, (j + 1) * S)
        ref = orc.ddim_single(onet, x[j], z0[sl], z[:, sl], steps, 0.5)
        assert np.abs(cov[j].cpu().numpy() - np.cov(ref.T)).max() < 1e-3 * max(1.0, np.abs(np.cov(ref.T)).max())


@pytest.mark.parametrize("shape", [(5, 8, (256, 256, 256), 148 * 128 * 2 + 77, 1), (4, 6, (128, 256), 3 * 128 + 1, 1),
                                   (2, 3, (64, 64), 500, 1), (5, 8, (256, 256, 256), 37 * 1100, 1100), (3, 4, (64, 128, 64), 9 * 130, 130)],
                         ids=lambda s: f"m{s[0]}_L{len(s[2])}_N{s[3]}_k{s[4]}")
def test_paired_tc_kernel(shape, dev):
    """`paired_tc_kernel` (row = sample with its own observation, nse.py:253-254; k samples per observation row = the covariance
    pre-pass): against the fp32 SIMT kernels on every row and against the fp64 oracle on a few, under replayed noise; several
    tiles per CTA (items of different tiles overlap), a single tile per CTA (program order), ragged last tile, 1 .. 3 tensor-core
    layers, clipping, and the Philox path (deterministic, independent of the kernel)."""
    from d4s import _cabi as abi
    from d4s import engine
    from d4s_helpers import nse_from_oracle
    m, d, hidden, N, k = shape
    rng = np.random.default_rng(21)
    onet = orc.random_net(rng, m, d, hidden=hidden)
    nse = nse_from_oracle(onet, dev)
    steps, eta = 6, 0.5
    B = N // k
    x = 0.5 * rng.standard_normal((B, d))
    z0, z = rng.standard_normal((N, m)).astype(np.float32), rng.standard_normal((steps, N, m)).astype(np.float32)
    noise = (torch.as_tensor(z0), torch.as_tensor(z))

    def run(**kw):
        if k == 1:
            return nse.ddim((N,), _t(x, dev), steps=steps, eta=eta, **kw)
        return nse.ddim_contexts((k,), _t(x, dev), steps=steps, eta=eta, **kw).reshape(N, m)

    try:
        abi.set_option("paired_tc", 0)
        n0 = engine.launch_count()
        simt = run(_noise=noise)
        simt_launches = engine.launch_count() - n0
        abi.set_option("paired_tc", 1)
        n0 = engine.launch_count()
        tc = run(_noise=noise)
        assert engine.launch_count() - n0 == 1 and simt_launches >= 2 * steps  # one launch per trajectory
        clip_tc = run(_noise=noise, theta_clipping_range=(-0.7, 0.9))
        abi.set_option("paired_tc", 0)
        clip_simt = run(_noise=noise, theta_clipping_range=(-0.7, 0.9))
        abi.set_option("paired_tc", 1)
        ph = [run(_seed=5), run(_seed=5), run(_seed=6)]
        abi.set_option("paired_tc", 0)
        ph_simt = run(_seed=5)
    finally:
        abi.set_option("paired_tc", 1)
    scale = max(1.0, float(simt.abs().max()))
    assert torch.isfinite(tc).all()
    assert float((tc - simt).abs().max()) < TRAJ_TOL * scale, float((tc - simt).abs().max())
    # (a random-init net gives no contraction: the scores are ~`scale` large, so the unclipped coordinates of a clipped run carry
    # the same relative tolerance as the unclipped run)
    assert float((clip_tc - clip_simt).abs().max()) < TRAJ_TOL * scale and float(clip_tc.max()) <= 0.9 and float(clip_tc.min()) >= -0.7
    assert torch.equal(ph[0], ph[1]) and not torch.equal(ph[0], ph[2])
    assert float((ph[0] - ph_simt).abs().max()) < TRAJ_TOL * max(1.0, float(ph_simt.abs().max()))  # same noise stream as kernel 4
    rows = sorted({0, 1, 127, 128, N // 2, N - 2, N - 1})
    xr = x[[r // k for r in rows]]
    ref = orc.ddim_single(onet, xr, z0[rows].astype(np.float64), z[:, rows].astype(np.float64), steps, eta)
    err = np.abs(tc[rows].cpu().numpy() - ref).max() / max(1.0, np.abs(ref).max())
    print(f"{shape}: max |tc - simt| = {float((tc - simt).abs().max()):.2e}, |tc - oracle| (rel) = {err:.2e}")
    assert err < TRAJ_TOL, err


def test_jac_trajectory_fuses_plain_sum_steps(dev):
    """A JAC trajectory only needs the Jacobian inside the alpha-gate (nse.py:643).  The run of plain-sum steps before it
    goes through the fused tcgen05 trajectory kernel in one launch; the result must follow the per-step fp32 SIMT
    kernels and the fp64 oracle under replayed noise."""
    from d4s import _cabi as abi
    from d4s import engine
    from d4s_helpers import nse_from_oracle
    m, d, n, N, steps = 4, 6, 9, 200, 10
    onet, x, cov, theta, oprior = _problem(m, d, (128, 256), n, N, "uniform", seed=9)
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    rng = np.random.default_rng(2)
    z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
    kw = dict(steps=steps, eta=0.8, prior=prior, prior_type="uniform", prior_score_fn=fn, cov_mode="JAC",
              _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
    before = engine.launch_count()
    out = nse.ddim((N,), _t(x, dev), **kw)
    fused_launches = engine.launch_count() - before
    abi.set_option("path", 1)
    try:
        before = engine.launch_count()
        simt = nse.ddim((N,), _t(x, dev), **kw)
        simt_launches = engine.launch_count() - before
    finally:
        abi.set_option("path", 0)
    ref = orc.ddim(onet, x, z0, z, steps, 0.8, oprior, None, "JAC", "uniform")
    scale = max(1.0, float(np.abs(ref).max()))
    e_f, e_s = (float(np.abs(o.cpu().numpy() - ref).max()) / scale for o in (out, simt))
    print(f"JAC trajectory: fused plain-sum run {fused_launches} launches (SIMT only: {simt_launches}); "
          f"max rel err vs oracle fused {e_f:.2e}, SIMT {e_s:.2e}")
    assert simt_launches == 2 * steps and fused_launches < simt_launches - 4  # steps 0..6 are outside the gate
    assert e_f < TRAJ_TOL and e_s < TRAJ_TOL


@pytest.mark.parametrize("N,n", [(1300, 30), (40, 300), (700, 5)], ids=["paired_groups", "obs_chunks", "many_samples_per_tile"])
def test_tc_trajectory_with_theta_embedding(N, n, dev):
    """Nets whose theta goes through an embedding MLP (the JR-NMM nets): the MLP runs inside the tcgen05 trajectory
    kernel per sample group.  Interleaved / sequential group schedules must agree bit for bit, and both follow the fp32
    SIMT path (separate theta-embedding kernel) and the fp64 oracle under replayed noise."""
    from d4s import _cabi as abi
    from d4s_helpers import nse_from_oracle
    m, d, F, steps = 4, 7, 5, 5
    rng = np.random.default_rng(21)

    def mlp(widths):
        ws, bs = [], []
        for a, b in zip(widths[:-1], widths[1:]):
            k = 1 / np.sqrt(a)
            ws.append(rng.uniform(-k, k, size=(b, a)))
            bs.append(rng.uniform(-k, k, size=(b,)))
        return orc.OracleMLP(ws, bs)

    onet = orc.OracleNet(m, d, np.arange(1, F + 1) * np.pi, mlp([24 + 16 + 2 * F, 128, 256, 64, m]), mlp([m, 48, 64, 24]),
                         mlp([d, 32, 16]))
    nse = nse_from_oracle(onet, dev)
    x = 0.5 * rng.standard_normal((n, d))
    a = 0.2 * rng.standard_normal((n, m, m))
    cov = a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)
    oprior = orc.UniformPrior(np.full(m, -3.46), np.full(m, 3.46))
    prior, fn = _d4s_prior(oprior, nse, dev)
    z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
    kw = dict(steps=steps, eta=1.0, prior=prior, prior_type="uniform", prior_score_fn=fn, dist_cov_est=_t(cov, dev),
              cov_mode="GAUSS", _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
    abi.set_option("path", 2)  # the tcgen05 trajectory kernel is required: fails loudly if the net were not eligible
    try:
        out = nse.ddim((N,), _t(x, dev), **kw)
        abi.set_option("tc_pairing", 0)
        out_seq = nse.ddim((N,), _t(x, dev), **kw)
        abi.set_option("path", 1)
        simt = nse.ddim((N,), _t(x, dev), **kw)
    finally:
        abi.set_option("tc_pairing", 1)
        abi.set_option("path", 0)
    assert torch.equal(out, out_seq)
    Nr = min(N, 3000 // n + 1)
    ref = orc.ddim(onet, x, z0[:Nr], z[:, :Nr], steps, 1.0, oprior, cov, "GAUSS", "uniform")
    scale = max(1.0, float(np.abs(ref).max()))
    e_tc = float(np.abs(out[:Nr].cpu().numpy() - ref).max()) / scale
    e_simt = float(np.abs(simt[:Nr].cpu().numpy() - ref).max()) / scale
    e_all = float((out - simt).abs().max()) / max(1.0, float(simt.abs().max()))
    print(f"theta-embedding net N={N} n={n}: tc-oracle {e_tc:.2e}, simt-oracle {e_simt:.2e}, tc-simt (all samples) {e_all:.2e}")
    assert e_tc < TRAJ_TOL and e_simt < TRAJ_TOL and e_all < TRAJ_TOL


def _random_embedding_cases(count, seed):
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(count):
        m = int(rng.integers(1, 11))
        d = int(rng.integers(1, 10))
        F = int(rng.integers(1, 9))
        emb_t = tuple(int(rng.choice([16, 24, 48, 64, 100, 160])) for _ in range(int(rng.integers(1, 4))))
        emb_x = tuple(int(rng.choice([8, 16, 32])) for _ in range(int(rng.integers(1, 3))))
        hidden = tuple(int(rng.choice([64, 96, 128, 200, 256])) for _ in range(int(rng.integers(2, 4))))
        n = int(rng.choice([2, 5, 17, 30, 33, 129, 300]))
        N = int(rng.choice([1, 3, 37, 150, 601, 1500]))
        if N * n > 40000:
            N = max(1, 40000 // n)
        cases.append((m, d, F, emb_t, emb_x, hidden, n, N, str(rng.choice(["uniform", "gaussian"])), int(rng.integers(0, 1 << 30))))
    return cases


@pytest.mark.parametrize("case", _random_embedding_cases(int(os.environ.get("D4S_FUZZ_CASES", "10")), int(os.environ.get("D4S_FUZZ_SEED", "31"))),
                         ids=lambda c: f"m{c[0]}_d{c[1]}_F{c[2]}_et{'x'.join(map(str, c[3]))}_ex{'x'.join(map(str, c[4]))}_h{'x'.join(map(str, c[5]))}_n{c[6]}_N{c[7]}_{c[8]}")
def test_tc_embedding_random_shapes(case, dev):
    """Seeded random embedding nets (theta MLP 1..3 layers up to 160 wide, x MLP, 1..8 time frequencies): the in-kernel
    theta embedding of the tcgen05 trajectory kernel against the fp32 SIMT path (separate embedding kernel) under
    replayed noise, both group schedules bit-identical, and run-to-run determinism."""
    from d4s import _cabi as abi
    from d4s_helpers import nse_from_oracle
    m, d, F, emb_t, emb_x, hidden, n, N, pkind, seed = case
    rng = np.random.default_rng(seed)

    def mlp(widths):
        ws, bs = [], []
        for a, b in zip(widths[:-1], widths[1:]):
            k = 1 / np.sqrt(a)
            ws.append(rng.uniform(-k, k, size=(b, a)))
            bs.append(rng.uniform(-k, k, size=(b,)))
        return orc.OracleMLP(ws, bs)

    onet = orc.OracleNet(m, d, np.arange(1, F + 1) * np.pi, mlp([emb_t[-1] + emb_x[-1] + 2 * F, *hidden, m]),
                         mlp([m, *emb_t]), mlp([d, *emb_x]))
    nse = nse_from_oracle(onet, dev)
    x = 0.5 * rng.standard_normal((n, d))
    a = 0.2 * rng.standard_normal((n, m, m))
    cov = a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)
    oprior = (orc.UniformPrior(np.full(m, -3.46), np.full(m, 3.46)) if pkind == "uniform"
              else orc.GaussianPrior(np.zeros(m), np.eye(m) * 4.0))
    prior, fn = _d4s_prior(oprior, nse, dev)
    steps = 3
    z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
    kw = dict(steps=steps, eta=1.0, prior=prior, prior_type=pkind, prior_score_fn=fn, dist_cov_est=_t(cov, dev),
              cov_mode="GAUSS", _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
    abi.set_option("path", 2)
    try:
        out = nse.ddim((N,), _t(x, dev), **kw)
        out2 = nse.ddim((N,), _t(x, dev), **kw)
        abi.set_option("tc_pairing", 0)
        out_seq = nse.ddim((N,), _t(x, dev), **kw)
        abi.set_option("path", 1)
        simt = nse.ddim((N,), _t(x, dev), **kw)
    finally:
        abi.set_option("tc_pairing", 1)
        abi.set_option("path", 0)
    assert torch.isfinite(out).all()
    assert torch.equal(out, out2) and torch.equal(out, out_seq)
    err = float((out - simt).abs().max()) / max(1.0, float(simt.abs().max()))
    print(f"{case[:9]}: tc vs simt max rel diff {err:.2e}")
    assert err < TRAJ_TOL


def _random_tc_cases(count, seed):
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(count):
        m = int(rng.integers(1, 11))
        d = int(rng.integers(1, 12))
        hidden = tuple(int(rng.choice([50, 64, 96, 128, 200, 256])) for _ in range(int(rng.integers(2, 5))))
        n = int(rng.choice([1, 2, 5, 9, 17, 30, 33, 64, 129, 300]))
        N = int(rng.choice([1, 3, 37, 150, 601, 1500]))
        if N * n > 60000:
            N = max(1, 60000 // n)
        cases.append((m, d, hidden, n, N, str(rng.choice(["uniform", "gaussian"])), int(rng.integers(0, 1 << 30))))
    return cases


@pytest.mark.parametrize("case", _random_tc_cases(int(os.environ.get("D4S_FUZZ_CASES", "14")), int(os.environ.get("D4S_FUZZ_SEED", "2024"))), ids=lambda c: f"m{c[0]}_d{c[1]}_h{'x'.join(map(str, c[2]))}_n{c[3]}_N{c[4]}_{c[5]}")
def test_tc_trajectory_random_shapes(case, dev):
    """Seeded random shapes (theta_dim 1..10, ragged widths, 1..300 observations, 1..1500 samples, both priors): the
    tcgen05 trajectory kernel against the per-step fp32 SIMT kernels under replayed noise, and its two group schedules
    against each other."""
    from d4s import _cabi as abi
    from d4s_helpers import nse_from_oracle
    m, d, hidden, n, N, pkind, seed = case
    onet, x, cov, theta, oprior = _problem(m, d, hidden, n, N, pkind, seed=seed)
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    steps = 3
    rng = np.random.default_rng(seed + 1)
    z0, z = rng.standard_normal((N, m)), rng.standard_normal((steps, N, m))
    kw = dict(steps=steps, eta=1.0, prior=prior, prior_type=pkind, prior_score_fn=fn, dist_cov_est=_t(cov, dev),
              cov_mode="GAUSS", _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
    if n == 1:
        kw = dict(steps=steps, eta=1.0, _noise=kw["_noise"])  # one observation: the plain score (nse.py:253)
    abi.set_option("path", 1)
    try:
        simt = nse.ddim((N,), _t(x, dev), **kw)
        abi.set_option("path", 0)
        out = nse.ddim((N,), _t(x, dev), **kw)
        abi.set_option("tc_pairing", 0)
        out_seq = nse.ddim((N,), _t(x, dev), **kw)
    finally:
        abi.set_option("tc_pairing", 1)
        abi.set_option("path", 0)
    assert torch.isfinite(out).all()
    assert torch.equal(out, out_seq)
    scale = max(1.0, float(simt.abs().max()))
    err = float((out - simt).abs().max()) / scale
    print(f"{case[:6]}: tc vs simt max rel diff {err:.2e}")
    assert err < TRAJ_TOL


@pytest.mark.parametrize("case", _random_tc_cases(int(os.environ.get("D4S_FUZZ_CASES", "12")), int(os.environ.get("D4S_FUZZ_SEED", "777"))), ids=lambda c: f"m{c[0]}_d{c[1]}_h{'x'.join(map(str, c[2]))}_n{c[3]}_N{c[4]}_{c[5]}")
def test_tc_per_call_random_shapes(case, dev):
    """Per-call `factorized_score` on seeded random shapes: the tcgen05 kernels in per-step mode (GAUSS: trajectory kernel
    exporting its partial sums; JAC, theta_dim <= 5: the opt-in score + Jacobian kernel) against the fp32 SIMT kernels,
    twice (run-to-run determinism)."""
    from d4s import _cabi as abi
    from d4s_helpers import nse_from_oracle
    m, d, hidden, n, N, pkind, seed = case
    if n == 1:
        n = 2
    onet, x, cov, theta, oprior = _problem(m, d, hidden, n, N, pkind, seed=seed)
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    th = _t(theta[:N], dev)
    for mode in ("GAUSS", "JAC"):
        for tv in (0.7, 0.2):
            args = (th, _t(x, dev), torch.tensor(np.float32(tv)), fn, prior)
            kw = dict(dist_cov_est=_t(cov, dev), cov_mode=mode, prior_type=pkind)
            abi.set_option("path", 1)
            try:
                simt = nse.factorized_score(*args, **kw)
                abi.set_option("path", 0)
                abi.set_option("jac_tc", 1)
                out = nse.factorized_score(*args, **kw)
                out2 = nse.factorized_score(*args, **kw)
            finally:
                abi.set_option("path", 0)
            assert torch.equal(out, out2), (mode, tv)
            if mode == "GAUSS":
                err = float((out - simt).norm() / simt.norm())
                assert err < 1e-4, (case[:6], mode, tv, err)
            else:
                # JAC: both paths go through inv(I - sigma J) of a random net; a near-singular sample is off by percents
                # in ANY fp32 evaluation (the fp32 oracle included) and would dominate an aggregate norm, so the check is
                # per sample: the typical sample agrees to 1e-4, and only a small fraction may disagree by more than 1 %
                rel = ((out - simt).norm(dim=1) / simt.norm(dim=1).clamp_min(1e-30)).cpu().numpy()
                assert np.median(rel) < 1e-4 and (rel > 1e-2).mean() <= 0.02, (case[:6], mode, tv, float(np.median(rel)),
                                                                           float((rel > 1e-2).mean()))


def test_errors_are_loud(dev):
    import d4s
    from d4s import _cabi as abi
    nse = d4s.NSE(3, 4, hidden_features=[64, 64]).to(dev)
    with pytest.raises(abi.D4sError):
        nse.ddim((8,), torch.randn(5, 4))  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        nse.ddim((8,), torch.randn(5, 4, device=dev), steps=3, prior_type="gaussian",
                 prior=torch.distributions.MultivariateNormal(torch.zeros(3, device=dev), torch.eye(3, device=dev)),
                 cov_mode="GAUSS")  # GAUSS without dist_cov_est
    # a net the kernels cannot pack (512-wide layer) is NOT an error any more: it runs on the generic path (torch scores +
    # native kernels 3/4), still on the GPU only
    wide = d4s.NSE(3, 4, hidden_features=[512]).to(dev)
    out = wide.ddim((8,), torch.randn(1, 4, device=dev), steps=2)
    assert out.shape == (8, 3) and bool(torch.isfinite(out).all())
    with pytest.raises(abi.D4sError):
        wide.ddim((8,), torch.randn(1, 4), steps=2)


# ---- Langevin / Geffner family on the same kernels -----------------------------------------
LANGEVIN = [c for c in gio.names() if c.startswith("langevin")]


@pytest.mark.parametrize("name", LANGEVIN)
def test_langevin_family(name, dev, kernel_path):
    """annealed Langevin (F-NPSE), tamed ULA, DDIM+Langevin predictor-corrector, deterministic Geffner: golden
    trajectories of the reference under replayed noise; per-call F-NPSE score."""
    g, onet, nse, prior, fn, ptype, kw = _setup(name, dev)
    steps, lsteps = int(g["traj_steps"]), int(g["lsteps"])
    x = _t(g["x"], dev)
    xt = _t(g["traj_x"], dev) if "traj_x" in g else x
    kwt = {} if "traj_x" in g else kw
    noise = (torch.as_tensor(g["traj_z0"]), torch.as_tensor(g["traj_z"]))
    for k, t in enumerate(g["t_grid"]):
        out = nse.factorized_score_geffner(_t(g["theta"], dev), x, torch.tensor(t), fn, **kw)
        ref = g["f64.geffner_score"][k]
        floor = orc.rel_l2(g["f32.geffner_score"][k], ref)
        assert orc.rel_l2(out.cpu().numpy(), ref) < max(REL_TOL, 3 * floor), (name, t)

    def check(out, key, n_noise):
        ref = g[f"f64.{key}.traj"]
        ref32 = np.abs(g[f"f32.{key}.traj"] - ref).max()
        scale = max(1.0, np.abs(ref).max())
        err = np.abs(out.cpu().numpy() - ref).max()
        print(f"{name} {key}: max|d theta| = {err:.2e} (reference fp32 vs fp64: {ref32:.2e})")
        assert err < max(TRAJ_TOL * scale, 3 * ref32), (name, key, err)

    z0, z = noise
    out = nse.annealed_langevin_geffner((z0.shape[0],), xt, fn, steps=steps, lsteps=lsteps, tau=0.5,
                                        theta_clipping_range=(-3, 3), _noise=(z0, z[:steps * lsteps]), **kwt)
    check(out, "langevin", steps * lsteps)
    out = nse.predictor_corrector((z0.shape[0],), xt, steps=steps, prior_score_fun=fn, lsteps=lsteps, r=0.5,
                                  predictor_type="id", theta_clipping_range=(-3, 3), _noise=(z0, z[:steps * lsteps]), **kwt)
    check(out, "tamed", steps * lsteps)
    out = nse.predictor_corrector((z0.shape[0],), xt, steps=steps, prior_score_fun=fn, lsteps=lsteps, r=0.3, eta=0.7,
                                  predictor_type="ddim", _noise=(z0, z[:steps * (lsteps + 1)]), **kwt)
    check(out, "pc", steps * (lsteps + 1))
    if "f64.detgef.traj" in g:
        out = nse.deterministic_gaussian_geffner((z0.shape[0],), x, fn, steps=steps, _noise=(z0, z[:steps]))
        check(out, "detgef", steps)


# ---- toy config 1: nse_toy.NSE + FakeFNet(AnalyticGaussianEps, EpsilonNet) --------------------------------------------
TOY = [c for c in gio.names() if c.startswith("toy")]


@pytest.mark.parametrize("name", TOY)
def test_toy_gaussian_config(name, dev):
    """Per-call scores / factorized_score (GAUSS, JAC), tall trajectories and the paired covariance pre-pass of the
    Gaussian toy task against golden outputs of the reference's nse_toy.NSE (gen_gaussian_gaussian.py flow)."""
    from d4s import nse_toy
    from d4s_helpers import toy_from_golden
    g = gio.load(name)
    nse, prior_fn = toy_from_golden(g, dev)
    theta, x, cov = _t(g["theta"], dev), _t(g["x"], dev), _t(g["dist_cov_est"], dev)
    for k, t in enumerate(g["t_grid"]):
        for mode in ("GAUSS", "JAC"):
            out = nse.factorized_score(theta, x, torch.tensor(t), prior_fn, dist_cov_est=cov, cov_mode=mode)
            ref = g[f"f64.{mode}.factorized_score"][k]
            floor = orc.rel_l2(g[f"f32.{mode}.factorized_score"][k], ref)
            err = orc.rel_l2(out.cpu().numpy(), ref)
            assert err < max(REL_TOL, 3 * floor), (name, mode, float(t), err, floor)
    steps, eta = int(g["traj_steps"]), float(g["traj_eta"])
    noise = (torch.as_tensor(g["traj_z0"]), torch.as_tensor(g["traj_z"]))
    for mode in ("GAUSS", "JAC"):
        out = nse.ddim((noise[0].shape[0],), x, steps=steps, eta=eta, prior_score_fn=prior_fn, dist_cov_est=cov,
                       cov_mode=mode, _noise=noise)
        ref = g[f"f64.{mode}.traj"]
        ref32 = np.abs(g[f"f32.{mode}.traj"] - ref).max()
        err = np.abs(out.cpu().numpy() - ref).max()
        print(f"{name} {mode}: max|d theta| = {err:.2e} (reference fp32 vs fp64: {ref32:.2e})")
        assert err < max(TRAJ_TOL * max(1.0, np.abs(ref).max()), 3 * ref32), (name, mode, err)
    reps = int(g["pair_reps"])
    xp = x[None].repeat(reps, 1, 1).reshape(reps * x.shape[0], -1)
    out = nse.ddim((xp.shape[0],), xp, steps=steps, eta=0.5, _noise=(torch.as_tensor(g["pair_z0"]), torch.as_tensor(g["pair_z"])))
    ref = g["f64.pair.traj"]
    assert np.abs(out.cpu().numpy() - ref).max() < TRAJ_TOL * max(1.0, np.abs(ref).max())
    # Langevin / Geffner baselines the toy script runs next to GAUSS / JAC (nse_toy.py:474-556)
    z0 = torch.as_tensor(g["traj_z0"])
    out = nse.annealed_langevin_geffner((z0.shape[0],), x, prior_fn, steps=int(g["lang_steps"]), lsteps=int(g["lang_lsteps"]),
                                        tau=float(g["lang_tau"]), _noise=(z0, torch.as_tensor(g["lang_z"])))
    ref = g["f64.langevin.traj"]
    ref32 = np.abs(g["f32.langevin.traj"] - ref).max()
    err = np.abs(out.cpu().numpy() - ref).max()
    print(f"{name} Langevin: max|d theta| = {err:.2e} (reference fp32 vs fp64: {ref32:.2e})")
    assert err < max(TRAJ_TOL * max(1.0, np.abs(ref).max()), 3 * ref32), (name, "langevin", err)
    out = nse.deterministic_gaussian_geffner((z0.shape[0],), x, prior_fn, steps=steps, _noise=(z0, torch.as_tensor(g["det_z"])))
    ref = g["f64.det_geffner.traj"]
    ref32 = np.abs(g["f32.det_geffner.traj"] - ref).max()
    err = np.abs(out.cpu().numpy() - ref).max()
    print(f"{name} deterministic Geffner: max|d theta| = {err:.2e} (reference fp32 vs fp64: {ref32:.2e})")
    assert err < max(TRAJ_TOL * max(1.0, np.abs(ref).max()), 3 * ref32), (name, "det_geffner", err)
    # the torch definition of the module (used outside the fused path) agrees with the kernels' affine form
    tt = torch.tensor(0.3, device=dev)
    e_torch = nse(theta, x[:1].expand(theta.shape[0], -1), tt)
    prec, _, sc = __import__("d4s").tweedies_approximation(x=x[:1], theta=theta, t=tt, nse=nse, dist_cov_est=cov[:1], mode="GAUSS")
    assert orc.rel_l2((-sc[:, 0] * nse.sigma(tt)).cpu().numpy(), e_torch.detach().cpu().numpy()) < 1e-5


# ---- classifier-free guidance: prior score / precision from the net itself at x = 0 (nse.py:598-625, 532-538) ----------
@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("cfg")])
def test_cfg_pf_nse_golden(name, dev, kernel_path):
    from d4s_helpers import nse_from_oracle
    g = gio.load(name)
    nse = nse_from_oracle(gio.oracle_net(g), dev)
    theta, xs, ns = _t(g["theta"], dev), _t(g["x"], dev), _t(g["kw.n"], dev)
    cov, covp = _t(g["dist_cov_est"], dev), _t(g["dist_cov_est_prior"], dev)
    for k, t in enumerate(g["t_grid"]):
        out = nse.factorized_score(theta, xs, torch.tensor(t), None, None, dist_cov_est=cov, dist_cov_est_prior=covp,
                                   cov_mode="GAUSS", clf_free_guidance=True, n=ns)
        ref = g["f64.cfg.factorized_score"][k]
        floor = orc.rel_l2(g["f32.cfg.factorized_score"][k], ref)
        assert orc.rel_l2(out.cpu().numpy(), ref) < max(REL_TOL, 3 * floor), (name, float(t))
        out = nse.factorized_score_geffner(theta, xs, torch.tensor(t), None, clf_free_guidance=True, n=ns)
        ref = g["f64.cfg.geffner_score"][k]
        floor = orc.rel_l2(g["f32.cfg.geffner_score"][k], ref)
        assert orc.rel_l2(out.cpu().numpy(), ref) < max(REL_TOL, 3 * floor), (name, float(t))
    noise = (torch.as_tensor(g["traj_z0"]), torch.as_tensor(g["traj_z"]))
    out = nse.ddim((noise[0].shape[0],), _t(g["traj_x"], dev), steps=int(g["traj_steps"]), eta=0.8, prior=None,
                   prior_score_fn=None, dist_cov_est=cov, dist_cov_est_prior=covp, cov_mode="GAUSS",
                   clf_free_guidance=True, _noise=noise)
    ref = g["f64.cfg.traj"]
    ref32 = np.abs(g["f32.cfg.traj"] - ref).max()
    err = np.abs(out.cpu().numpy() - ref).max()
    assert err < max(TRAJ_TOL * max(1.0, np.abs(ref).max()), 3 * ref32), err


@pytest.mark.parametrize("mode", ["GAUSS", "JAC"])
def test_cfg_plain_nse_vs_oracle(mode, dev, kernel_path):
    """Plain NSE + CFG raises UnboundLocalError in the reference; the intended semantics
    (kwargs_score_prior = {}) are checked against the oracle."""
    from d4s_helpers import nse_from_oracle
    rng = np.random.default_rng(5)
    m, d, n, N = 4, 6, 9, 33
    onet = orc.random_net(rng, m, d, hidden=(128, 128, 64))
    nse = nse_from_oracle(onet, dev)
    x, theta = rng.standard_normal((n, d)), rng.standard_normal((N, m))
    a = 0.2 * rng.standard_normal((n + 1, m, m))
    covs = a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)
    for t in (0.7, 0.2):
        t32 = np.float32(t)
        ref = orc.factorized_score_cfg(onet, theta, x, np.float64(t32), covs[:n], covs[n:], mode)
        ref32 = orc.factorized_score_cfg(onet.astype(np.float32), theta.astype(np.float32), x.astype(np.float32), t32,
                                         covs[:n].astype(np.float32), covs[n:].astype(np.float32), mode)
        out = nse.factorized_score(_t(theta, dev), _t(x, dev), torch.tensor(t32), None, None, dist_cov_est=_t(covs[:n], dev),
                                   dist_cov_est_prior=_t(covs[n:], dev), cov_mode=mode, clf_free_guidance=True)
        err, floor = orc.rel_l2(out.cpu().numpy(), ref), orc.rel_l2(ref32, ref)
        assert err < max(REL_TOL, 3 * floor), (mode, t, err, floor)


def test_euler_sde_sampler_vs_oracle(dev):
    """Euler-Maruyama sampler (tall_posterior_sampler.py:217-281) driven by diffused_tall_posterior_score."""
    from functools import partial
    import d4s
    from d4s_helpers import nse_from_oracle
    onet, x, cov, theta, oprior = _problem(3, 4, (64, 64, 64), 5, 8, "gaussian", seed=7)
    nse = nse_from_oracle(onet, dev)
    prior, fn = _d4s_prior(oprior, nse, dev)
    rng = np.random.default_rng(9)
    z0, z = rng.standard_normal((8, 3)), rng.standard_normal((999, 8, 3))
    ref = orc.euler_sde_sampler(onet, x, z0, z, oprior, cov, "GAUSS", "gaussian", clip=(-4.0, 4.0))
    score_fn = partial(d4s.diffused_tall_posterior_score, prior=prior, prior_type="gaussian", prior_score_fn=fn,
                       x_obs=_t(x, dev), nse=nse, dist_cov_est=_t(cov, dev), cov_mode="GAUSS")
    out, path = d4s.euler_sde_sampler(score_fn, 8, 3, nse.beta, device=dev, theta_clipping_range=(-4.0, 4.0),
                                      _noise=(torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32)))
    assert len(path) == 1000
    err = np.abs(out.cpu().numpy() - ref).max()
    assert err < 5e-3 * max(1.0, np.abs(ref).max()), err
