"""Generic path (-m gpu): score networks that cannot be fused run with torch-evaluated scores and native kernels 3/4
(d4s_reduce_pairs, d4s_finalize_wide).  SURVEY.md section 7 hard parts (b), section 8 row a13.

  * the reference's toy driver in its ORIGINAL closure form -- `FakeFNet(real_eps_fun=<python closure>, ...)` and the
    triple-returning prior closure (gen_gaussian_gaussian.py:59-96) -- for DIM 4, 10, 16 and 32 against trajectories and
    per-call scores recorded from the reference;
  * the same (fusable) network through the generic path and through the fused kernels: identical schedule tables, identical
    noise => the two paths must agree to fp32 accuracy;
  * kernel 3a / 3b alone against numpy.
"""
import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import d4s_oracle as orc

pytestmark = pytest.mark.gpu
REL_TOL, TRAJ_TOL = 1e-4, 1e-3


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _t(v, dev):
    return torch.as_tensor(np.asarray(v, np.float32), device=dev)


def closure_toy(g, dev):
    """The toy score net and prior built the way gen_gaussian_gaussian.py:42-96 builds them: plain Python closures."""
    from d4s import nse_toy
    from d4s.tall_posterior_sampler import mean_backward, prec_matrix_backward
    from d4s.vp_diffused_priors import get_vpdiff_gaussian_score
    from d4s_helpers import _load_mlp
    m = int(g["theta_dim"])
    score_net = nse_toy.NSE(theta_dim=m, x_dim=g["x"].shape[-1])
    loc, pcov, lcov = (_t(g[k], dev) for k in ("prior.loc", "prior.cov", "lik_cov"))
    idm = torch.eye(m, device=dev)
    inv_lik, inv_prior = torch.linalg.inv(lcov.double()).float(), torch.linalg.inv(pcov.double()).float()
    posterior_cov = torch.linalg.inv((inv_lik + inv_prior).double()).float()
    inv_prior_prior = inv_prior @ loc

    def real_eps(theta, x, t):  # gen_gaussian_gaussian.py:59-72
        a = score_net.alpha(t)
        cov_diff = posterior_cov[None] * a[..., None] + (1 - a)[..., None] * idm[None]
        mean_0 = (posterior_cov @ (inv_prior_prior[:, None] + inv_lik @ x.mT)).mT
        score = -(torch.linalg.inv(cov_diff) @ (theta - (a ** 0.5) * mean_0)[..., None])[..., 0]
        return -score_net.sigma(t) * score

    eps_net = nse_toy.EpsilonNet(m)
    _load_mlp(eps_net.net, [g["net.W0"], g["net.W1"]], [g["net.b0"], g["net.b1"]])
    score_net.net = nse_toy.FakeFNet(real_eps_fun=real_eps, eps_net=eps_net, eps_net_max=float(g["eps_max"]))
    score_net = score_net.to(dev).eval()
    prior_score_fn = get_vpdiff_gaussian_score(loc, pcov, score_net)

    def prior_score(theta, t):  # gen_gaussian_gaussian.py:86-96
        mean_prior = mean_backward(theta, t, prior_score_fn, score_net)
        prec_prior = prec_matrix_backward(t, pcov, score_net).repeat(theta.shape[0], 1, 1)
        return prior_score_fn(theta, t), mean_prior, prec_prior

    return score_net, prior_score


@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("toy")])
def test_toy_closure_form_against_reference_goldens(name, dev):
    g = gio.load(name)
    nse, prior_fn = closure_toy(g, dev)
    theta, x, cov = _t(g["theta"], dev), _t(g["x"], dev), _t(g["dist_cov_est"], dev)
    for k, t in enumerate(g["t_grid"]):
        for mode in ("GAUSS", "JAC"):
            out = nse.factorized_score(theta, x, torch.tensor(t, device=dev), prior_fn, dist_cov_est=cov, cov_mode=mode)
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
        print(f"{name} closure form {mode}: max|d theta| = {err:.2e} (reference fp32 vs fp64: {ref32:.2e})")
        assert err < max(TRAJ_TOL * max(1.0, np.abs(ref).max()), 3 * ref32), (name, mode, err)
    reps = int(g["pair_reps"])
    xp = x[None].repeat(reps, 1, 1).reshape(reps * x.shape[0], -1)
    out = nse.ddim((xp.shape[0],), xp, steps=steps, eta=0.5, _noise=(torch.as_tensor(g["pair_z0"]), torch.as_tensor(g["pair_z"])))
    ref = g["f64.pair.traj"]
    assert np.abs(out.cpu().numpy() - ref).max() < TRAJ_TOL * max(1.0, np.abs(ref).max())
    z0 = torch.as_tensor(g["traj_z0"])
    out = nse.annealed_langevin_geffner((z0.shape[0],), x, prior_fn, steps=int(g["lang_steps"]), lsteps=int(g["lang_lsteps"]),
                                        tau=float(g["lang_tau"]), _noise=(z0, torch.as_tensor(g["lang_z"])))
    ref = g["f64.langevin.traj"]
    ref32 = np.abs(g["f32.langevin.traj"] - ref).max()
    assert np.abs(out.cpu().numpy() - ref).max() < max(TRAJ_TOL * max(1.0, np.abs(ref).max()), 3 * ref32)
    out = nse.deterministic_gaussian_geffner((z0.shape[0],), x, prior_fn, steps=steps, _noise=(z0, torch.as_tensor(g["det_z"])))
    ref = g["f64.det_geffner.traj"]
    ref32 = np.abs(g["f32.det_geffner.traj"] - ref).max()
    assert np.abs(out.cpu().numpy() - ref).max() < max(TRAJ_TOL * max(1.0, np.abs(ref).max()), 3 * ref32)


@pytest.mark.parametrize("name", ["slcp_n6", "lv_n6"])
@pytest.mark.parametrize("mode", ["GAUSS", "JAC"])
def test_generic_path_matches_fused_path(name, mode, dev, monkeypatch):
    """Same net, same tables, same noise: `generic.fusable` forced to False sends nse.NSE through torch scores + kernels 3/4."""
    from d4s import generic
    from d4s_helpers import nse_from_oracle, prior_objects
    g = gio.load(name)
    nse = nse_from_oracle(gio.oracle_net(g), dev)
    prior, fn, ptype = prior_objects(g, nse, dev)
    theta, x, cov = _t(g["theta"], dev), _t(g["x"], dev), _t(g["dist_cov_est"], dev)
    kw = dict(prior=prior, prior_type=ptype, prior_score_fn=fn, dist_cov_est=cov, cov_mode=mode)
    z0, z = torch.as_tensor(g["traj_z0"]), torch.as_tensor(g["traj_z"])[:12]
    fused_s = [nse.factorized_score(theta, x, torch.tensor(t), fn, prior, dist_cov_est=cov, cov_mode=mode, prior_type=ptype)
               for t in g["t_grid"]]
    fused_t = nse.ddim((z0.shape[0],), x, steps=12, eta=0.8, _noise=(z0, z), **kw)
    monkeypatch.setattr(generic, "fusable", lambda nse_: False)
    for k, t in enumerate(g["t_grid"]):
        out = nse.factorized_score(theta, x, torch.tensor(t), fn, prior, dist_cov_est=cov, cov_mode=mode, prior_type=ptype)
        ref = g[f"f64.{mode}.factorized_score"][k]
        floor = orc.rel_l2(g[f"f32.{mode}.factorized_score"][k], ref)
        err = orc.rel_l2(out.cpu().numpy(), ref)
        assert err < max(REL_TOL, 3 * floor), (name, mode, float(t), err, floor, orc.rel_l2(fused_s[k].cpu().numpy(), ref))
    gen_t = nse.ddim((z0.shape[0],), x, steps=12, eta=0.8, _noise=(z0, z), **kw)
    d = float((gen_t - fused_t).abs().max())
    print(f"{name} {mode}: generic vs fused trajectory (12 steps) max |d theta| = {d:.2e}")
    if mode == "GAUSS":  # JAC trajectories are chaotic on these fixtures (the reference's own fp32 / fp64 runs end O(1) apart,
        assert d < 1e-3  # SURVEY.md section 8(c)): JAC is gated per call above and distributionally in test_gpu_long_runs.py


@pytest.mark.parametrize("m,n,N", [(3, 7, 50), (16, 40, 33), (32, 5, 9)])
def test_reduce_pairs_and_finalize_wide_vs_numpy(m, n, N, dev):
    from d4s import _cabi as abi
    from d4s import generic
    rng = np.random.default_rng(m)
    s = rng.standard_normal((N, n, m)).astype(np.float32)
    a = 0.3 * rng.standard_normal((N, n, m, m))
    P = (a @ a.transpose(0, 1, 3, 2) + np.eye(m)).astype(np.float32)
    Q = P[0]
    part = generic.reduce_pairs(_t(s, dev), _t(P, dev), None).cpu().numpy()
    ref = np.concatenate((np.einsum("ijab,ijb->ia", P.astype(np.float64), s), P.astype(np.float64).sum(1).reshape(N, -1)), 1)
    assert orc.rel_l2(part, ref) < 1e-5
    part_g = generic.reduce_pairs(_t(s, dev), None, _t(Q, dev)).cpu().numpy()
    ref_g = np.concatenate((s.astype(np.float64).sum(1), np.einsum("jab,ijb->ia", Q.astype(np.float64), s)), 1)
    assert orc.rel_l2(part_g, ref_g) < 1e-5
    # kernel 3b: external prior, per-sample solve, no update
    theta = rng.standard_normal((N, m)).astype(np.float32)
    s0 = rng.standard_normal((N, m)).astype(np.float32)
    b = 0.2 * rng.standard_normal((N, m, m))
    P0 = (b @ b.transpose(0, 2, 1) + 0.5 * np.eye(m)).astype(np.float32)
    sched = np.zeros(abi.SCHED_STRIDE, np.float32)
    sched[abi.S_COMBINE], sched[abi.S_ONE_MINUS_N], sched[abi.S_A_OVER_S2], sched[abi.S_SIGMA] = 2, 1 - n, 0.7, 0.5
    sched[abi.S_SQRT_ALPHA], sched[abi.S_S1_WEIGHT] = 0.8, 1.0
    out = generic.finalize_wide(_t(part, dev), _t(theta, dev), True, _t(sched, dev), None, abi.PRIOR_EXTERNAL,
                                ext_s0=_t(s0, dev), ext_p0=_t(P0, dev), update=False, want_score=True).cpu().numpy()
    lam = part[:, m:].reshape(N, m, m).astype(np.float64) + (1 - n) * P0
    w = part[:, :m].astype(np.float64) + (1 - n) * np.einsum("iab,ib->ia", P0.astype(np.float64), s0)
    ref = np.linalg.solve(lam, w[..., None])[..., 0]
    cond = np.linalg.cond(lam).max()
    assert orc.rel_l2(out, ref) < 1e-6 * max(10.0, cond), (orc.rel_l2(out, ref), cond)
