"""NOTE on tolerances: the reference's "fp64" run still rounds a few scalars to fp32 (torch.eye() is
fp32 and a 0-d fp64 tensor does not promote it: vp_diffused_priors.py:65-68, tall_posterior_sampler.py:38-43),
so reference-fp64 and a true fp64 evaluation differ by ~3e-8 relative; 5e-7 is used as "equal".

Pins the numpy oracle (oracle/d4s_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz).  fp64 oracle vs fp64 reference must agree to rounding; the fp32 oracle must sit
inside the reference's own fp32-vs-fp64 noise floor (SURVEY.md Finding B)."""
import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import d4s_oracle as orc

CASES = [c for c in gio.names() if not c.startswith(("langevin", "toy", "cfg", "contexts", "euler", "jacdist"))]


def _ns(g):
    return g["kw.n"].astype(np.float64) if "kw.n" in g else None


@pytest.mark.parametrize("steps", [1, 2, 7, 25, 30, 40, 400, 1000])
def test_time_grid_matches_torch_linspace(steps):
    ref = torch.linspace(1, 0, steps + 1).numpy()
    assert np.array_equal(orc.time_grid(steps), ref)


@pytest.mark.parametrize("name", CASES)
def test_scores_and_prior_fp64(name):
    g = gio.load(name)
    net, prior, ns = gio.oracle_net(g), gio.oracle_prior(g), _ns(g)
    theta, x = g["theta"].astype(np.float64), g["x"].astype(np.float64)
    for k, t in enumerate(g["t_grid"]):
        t = np.float64(t)
        _, _, sc = orc.tweedies_approximation(net, x, theta, t, g["dist_cov_est"].astype(np.float64), "GAUSS", ns)
        assert orc.rel_l2(sc, g["f64.scores"][k]) < 1e-11
        assert orc.rel_l2(orc.prior_score(net, prior, theta, t), g["f64.prior_score"][k]) < 5e-7
        ref_pp = g["f64.prior_prec"][k]
        if np.isfinite(ref_pp).all():
            pp, _, _ = orc.tweedies_approximation_prior(net, prior, theta, t)
            assert orc.rel_l2(pp, ref_pp) < 5e-7


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["GAUSS", "JAC"])
def test_factorized_score_fp64(name, mode):
    g = gio.load(name)
    if f"f64.{mode}.factorized_score" not in g:
        pytest.skip("mode not in fixture (PF+JAC has no reference behaviour)")
    net, prior, ns = gio.oracle_net(g), gio.oracle_prior(g), _ns(g)
    theta, x = g["theta"].astype(np.float64), g["x"].astype(np.float64)
    cov = g["dist_cov_est"].astype(np.float64)
    for k, t in enumerate(g["t_grid"]):
        t = np.float64(t)
        prec, _, _ = orc.tweedies_approximation(net, x, theta[:2], t, cov, mode, ns)
        assert orc.rel_l2(prec, g[f"f64.{mode}.prec"][k]) < 5e-7, (name, mode, t)
        out = orc.factorized_score(net, theta, x, t, prior, cov, mode, str(g["prior_kind"]), ns)
        ref = g[f"f64.{mode}.factorized_score"][k]
        # JAC solves ill-conditioned systems: compare with a tolerance scaled by the reference's own
        # fp32-vs-fp64 disagreement (the per-pair precisions above are pinned to 5e-7; the solve
        # amplifies the reference's fp32-rounded scalars by the condition number of Lambda).
        floor = orc.rel_l2(g[f"f32.{mode}.factorized_score"][k], ref)
        assert orc.rel_l2(out, ref) < max(5e-7, (1e-3 if mode == "GAUSS" else 3.0) * floor), (name, mode, t)


@pytest.mark.parametrize("name", CASES)
def test_factorized_score_fp32_within_reference_noise(name):
    g = gio.load(name)
    net = gio.oracle_net(g, np.float32)
    prior, ns = gio.oracle_prior(g), _ns(g)
    theta, x = g["theta"].astype(np.float32), g["x"].astype(np.float32)
    cov = g["dist_cov_est"].astype(np.float32)
    ns32 = None if ns is None else ns.astype(np.float32)
    for k, t in enumerate(g["t_grid"]):
        out = orc.factorized_score(net, theta, x, np.float32(t), prior, cov, "GAUSS", str(g["prior_kind"]), ns32)
        assert out.dtype == np.float32
        ref64 = g["f64.GAUSS.factorized_score"][k]
        floor = orc.rel_l2(g["f32.GAUSS.factorized_score"][k], ref64)
        assert orc.rel_l2(out, ref64) < max(1e-4, 10 * floor), (name, t)


@pytest.mark.parametrize("name", CASES)
def test_trajectory_fp64(name):
    g = gio.load(name)
    net, prior, ns = gio.oracle_net(g), gio.oracle_prior(g), _ns(g)
    x = g["x"].astype(np.float64)
    clip = tuple(g["traj_clip"]) if "traj_clip" in g else (None, None)
    steps = int(g["traj_steps"])
    if steps > 400 and not g["f64.GAUSS.traj"].shape[0] <= 16:
        pytest.skip("long")
    for mode in ("GAUSS", "JAC"):
        key = f"f64.{mode}.traj"
        if key not in g:
            continue
        out = orc.ddim(net, x, g["traj_z0"], g["traj_z"], steps, float(g["traj_eta"]), prior,
                       g["dist_cov_est"].astype(np.float64), mode, str(g["prior_kind"]), clip, ns)
        err = np.abs(out - g[key]).max()
        assert err < (5e-6 if mode == "GAUSS" else 1e-4), (name, mode, err)


# ---- batched contexts: the reference run context by context = vmap(sampler) ----------------
@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("contexts")])
def test_contexts_fp64(name):
    g = gio.load(name)
    net, prior = gio.oracle_net(g), gio.oracle_prior(g)
    x, cov = g["x"].astype(np.float64), g["dist_cov_est"].astype(np.float64)
    B, S, steps, eta = x.shape[0], int(g["samples_per_context"]), int(g["traj_steps"]), float(g["traj_eta"])
    for c in range(B):
        sl = slice(c * S, (c + 1) * S)
        for mode in ("GAUSS", "JAC"):
            out = orc.ddim(net, x[c], g["traj_z0"][sl], g["traj_z"][:, sl], steps, eta, prior, cov[c], mode, str(g["prior_kind"]))
            err = np.abs(out - g[f"f64.{mode}.traj"][c]).max()
            assert err < (5e-6 if mode == "GAUSS" else 1e-4), (name, c, mode, err)
        out = orc.ddim_single(net, x[c, 0], g["traj_z0"][sl], g["traj_z"][:, sl], steps, eta)
        assert np.abs(out - g["f64.single.traj"][c]).max() < 5e-6, (name, c)


# ---- Langevin / Geffner family ---------------------------------------------------------
LANGEVIN = [c for c in gio.names() if c.startswith("langevin")]


def _lv_setup(name):
    g = gio.load(name)
    net, prior, ns = gio.oracle_net(g), gio.oracle_prior(g), _ns(g)
    steps, lsteps = int(g["traj_steps"]), int(g["lsteps"])
    return g, net, prior, ns, steps, lsteps


@pytest.mark.parametrize("name", LANGEVIN)
def test_langevin_family_fp64(name):
    g, net, prior, ns, steps, lsteps = _lv_setup(name)
    x = g["x"].astype(np.float64)
    z0, z = g["traj_z0"], g["traj_z"]
    for k, t in enumerate(g["t_grid"]):
        out = orc.factorized_score_geffner(net, g["theta"].astype(np.float64), x, np.float64(t), prior, ns)
        assert orc.rel_l2(out, g["f64.geffner_score"][k]) < 5e-7
    out = orc.annealed_langevin_geffner(net, x, z0, z, prior, steps, lsteps, 0.5, (-3, 3), ns)
    assert np.abs(out - g["f64.langevin.traj"]).max() < 5e-6
    out = orc.predictor_corrector(net, x, z0, z, prior, steps, "id", "langevin", lsteps=lsteps, r=0.5, clip=(-3, 3), ns=ns)
    assert np.abs(out - g["f64.tamed.traj"]).max() < 5e-6
    out = orc.predictor_corrector(net, x, z0, z, prior, steps, "ddim", "langevin", eta=0.7, lsteps=lsteps, r=0.3, ns=ns)
    ref = g["f64.pc.traj"]
    assert np.abs(out - ref).max() < 5e-6 * max(1.0, np.abs(ref).max())
    if "f64.detgef.traj" in g:
        out = orc.deterministic_gaussian_geffner(net, x, z0, z, prior, steps)
        ref = g["f64.detgef.traj"]
        assert np.abs(out - ref).max() < 5e-6 * max(1.0, np.abs(ref).max())


# ---- toy Gaussian-Gaussian task, nse_toy.NSE + FakeFNet (BASELINE config 1) ---------------
TOY = [c for c in gio.names() if c.startswith("toy")]


@pytest.mark.parametrize("name", TOY)
def test_toy_config_fp64(name):
    g = gio.load(name)
    net, prior = gio.oracle_net(g), gio.oracle_prior(g)
    theta, x, cov = (g[k].astype(np.float64) for k in ("theta", "x", "dist_cov_est"))
    for k, t in enumerate(g["t_grid"]):
        t = np.float64(t)
        _, _, sc = orc.tweedies_approximation(net, x, theta, t, cov, "GAUSS")
        assert orc.rel_l2(sc, g["f64.scores"][k]) < 5e-7
        for mode in ("GAUSS", "JAC"):
            out = orc.factorized_score(net, theta, x, t, prior, cov, mode, "gaussian")  # nse_toy.py:599-625
            ref = g[f"f64.{mode}.factorized_score"][k]
            floor = orc.rel_l2(g[f"f32.{mode}.factorized_score"][k], ref)
            assert orc.rel_l2(out, ref) < max(5e-7, (1e-3 if mode == "GAUSS" else 3.0) * floor), (name, mode, t)
    steps, eta = int(g["traj_steps"]), float(g["traj_eta"])
    for mode in ("GAUSS", "JAC"):
        out = orc.ddim(net, x, g["traj_z0"], g["traj_z"], steps, eta, prior, cov, mode, "gaussian")
        ref = g[f"f64.{mode}.traj"]
        assert np.abs(out - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), (name, mode)
    reps = int(g["pair_reps"])
    xp = np.tile(x[None], (reps, 1, 1)).reshape(reps * x.shape[0], -1)
    out = orc.ddim_single(net, xp, g["pair_z0"], g["pair_z"], steps, 0.5)
    ref = g["f64.pair.traj"]
    assert np.abs(out - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())
    # Langevin / Geffner baselines of the toy script (nse_toy.py:474-556; gen_gaussian_gaussian.py:205-238)
    out = orc.annealed_langevin_geffner(net, x, g["traj_z0"], g["lang_z"], prior, steps=int(g["lang_steps"]),
                                        lsteps=int(g["lang_lsteps"]), tau=float(g["lang_tau"]))
    ref = g["f64.langevin.traj"]
    assert np.abs(out - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), name
    out = orc.deterministic_gaussian_geffner(net, x, g["traj_z0"], g["det_z"], prior, steps=steps)
    ref = g["f64.det_geffner.traj"]
    assert np.abs(out - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), name


# ---- classifier-free guidance (runs in the reference for PF_NSE only) ----------------------------
@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("cfg")])
def test_cfg_branches_fp64(name):
    g = gio.load(name)
    net, ns = gio.oracle_net(g), _ns(g)
    theta, x = g["theta"].astype(np.float64), g["x"].astype(np.float64)
    cov, covp = g["dist_cov_est"].astype(np.float64), g["dist_cov_est_prior"].astype(np.float64)
    for k, t in enumerate(g["t_grid"]):
        out = orc.factorized_score_cfg(net, theta, x, np.float64(t), cov, covp, "GAUSS", ns)
        assert orc.rel_l2(out, g["f64.cfg.factorized_score"][k]) < 5e-7
        out = orc.factorized_score_geffner_cfg(net, theta, x, np.float64(t), ns)
        assert orc.rel_l2(out, g["f64.cfg.geffner_score"][k]) < 5e-7
    out = orc.ddim_cfg(net, x, g["traj_z0"], g["traj_z"], int(g["traj_steps"]), 0.8, cov, covp, "GAUSS", ns)
    assert np.abs(out - g["f64.cfg.traj"]).max() < 5e-6


def _replay_noise(seed, steps, N, m):
    rng = np.random.default_rng(seed)  # oracle/gen_golden.py:replay_noise
    return rng.standard_normal((N, m)).astype(np.float32), rng.standard_normal((steps, N, m)).astype(np.float32)


@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("euler")])
@pytest.mark.parametrize("mode", ["GAUSS", "JAC"])
def test_euler_sde_sampler_pinned_to_the_reference(name, mode):
    """oracle.euler_sde_sampler == the reference's euler_sde_sampler on diffused_tall_posterior_score
    (tall_posterior_sampler.py:158-281), fp64, replayed noise."""
    e = gio.load(name)
    g = gio.load(str(e["base"]))
    net, prior = gio.oracle_net(g), gio.oracle_prior(g)
    z0, z = _replay_noise(int(e["seed"]), 999, int(e["N"]), net.theta_dim)
    out = orc.euler_sde_sampler(net, g["x"].astype(np.float64), z0.astype(np.float64), z.astype(np.float64), prior,
                                g["dist_cov_est"].astype(np.float64), mode, str(g["prior_kind"]))
    err = np.abs(out - e[f"f64.{mode}.theta"]).max()
    floor = np.abs(e[f"f32.{mode}.theta"] - e[f"f64.{mode}.theta"]).max()
    # GAUSS is pinned pathwise (errors ~1e-6).  JAC trajectories are chaotic: the reference's own fp32 and fp64 runs end
    # O(1) apart on these fixtures (floor), so the pathwise bound is only a sanity check there;
    # the JAC gate is distributional (tests/test_gpu_long_runs.py::test_jac_distribution_gate).
    assert err < (1e-5 if mode == "GAUSS" else max(1e-5, 3 * floor)), (name, mode, err, floor)


@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("jacdist")])
def test_jac_sample_sets_are_well_formed(name):
    e = gio.load(name)
    g = gio.load(str(e["base"]))
    for tag in ("f32", "f64"):
        s = e[f"{tag}.samples"]
        assert s.shape == (int(e["N"]), int(g["theta_dim"])) and np.isfinite(s).mean() > 0.9
