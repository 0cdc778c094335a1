"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference
with the zuko/sbi stand-ins of oracle/ref_shims) on seeded inputs, in fp32 and fp64.

TEST INFRASTRUCTURE ONLY; runs in the build container (the reference does not travel to the GPU
box -- the committed .npz files do).  Usage:

    python oracle/gen_golden.py            # writes every fixture
    python oracle/gen_golden.py slcp       # only fixtures whose name contains "slcp"

Each fixture holds: the net weights (fp32 as shipped), inputs (theta, x, t grid, dist_cov_est, prior
parameters, injected noise) and the reference outputs of the section-8(a) rows:
  scores/prec of `tweedies_approximation` (GAUSS, JAC), `prior_score_fn`, `tweedies_approximation_prior`,
  `factorized_score` (GAUSS, JAC) and full `ddim` trajectories under replayed noise.
"""
from __future__ import annotations

import copy
import os
import sys
from functools import partial
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("D4S_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "ref_shims"), REF]

import torch  # noqa: E402

import nse as ref_nse  # noqa: E402
import pf_nse as ref_pf  # noqa: E402
import tall_posterior_sampler as ref_tps  # noqa: E402
import vp_diffused_priors as ref_vp  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
T_GRID = [1.0, 0.9, 0.5, 0.3, 0.29, 0.1, 0.01, 0.001]


def npy(t):
    return t.detach().cpu().numpy()


def mlp_arrays(prefix, mlp, out):
    lin = [mod for mod in mlp if isinstance(mod, torch.nn.Linear)]
    for i, l in enumerate(lin):
        out[f"{prefix}.W{i}"] = npy(l.weight).astype(np.float32)
        out[f"{prefix}.b{i}"] = npy(l.bias).astype(np.float32)
    out[f"{prefix}.L"] = np.int64(len(lin))


def net_arrays(net, out):
    mlp_arrays("net", net.net, out)
    if not isinstance(net.embedding_nn_theta, torch.nn.Identity):
        mlp_arrays("embt", net.embedding_nn_theta, out)
    if not isinstance(net.embedding_nn_x, torch.nn.Identity):
        mlp_arrays("embx", net.embedding_nn_x, out)
    out["freqs"] = npy(net.freqs).astype(np.float64)
    out["theta_dim"] = np.int64(net.zeros.shape[0])
    out["n_max"] = np.int64(getattr(net, "n_max", 1))
    out["pf"] = np.int64(isinstance(net, ref_pf.PF_NSE))


def spd_covs(gen, n, m):
    a = 0.2 * torch.randn(n, m, m, generator=gen)
    return a @ a.mT + 0.05 * torch.eye(m)


class FakeDiagNormal:
    """Replaces nse.DiagNormal so that `ddim` starts from an injected z0 (nse.py:261)."""
    z0 = None

    def __init__(self, *a, **k):
        pass

    def sample(self, shape):
        return FakeDiagNormal.z0.clone()


def run_ddim(net, x, z0, z, **kw):
    it = iter(z)
    FakeDiagNormal.z0 = z0
    with mock.patch.object(ref_nse, "DiagNormal", FakeDiagNormal), \
            mock.patch("torch.randn_like", side_effect=lambda a: next(it).to(a)):
        return net.ddim(shape=(z0.shape[0],), x=x, **kw)


def run_sampler(net, method, x, z0, z, **kw):
    """Any reference sampler under replayed noise: z0 replaces DiagNormal.sample, z[k] the k-th torch.randn_like."""
    it = iter(z)
    FakeDiagNormal.z0 = z0
    with mock.patch.object(ref_nse, "DiagNormal", FakeDiagNormal), \
            mock.patch("torch.randn_like", side_effect=lambda a: next(it).to(a)):
        return getattr(net, method)(shape=(z0.shape[0],), x=x, **kw)


def langevin_case(name, net32, x, prior_kind, prior_params, N, seed, steps, lsteps, traj_x=None, extra_kwargs=None):
    """Trajectories of the Langevin / Geffner family + per-call F-NPSE scores, fp32 and fp64."""
    out = {}
    net_arrays(net32, out)
    gen = torch.Generator().manual_seed(seed)
    m = net32.zeros.shape[0]
    theta = torch.randn(N, m, generator=gen)
    z0 = torch.randn(N, m, generator=gen)
    z = torch.randn(steps * (lsteps + 1), N, m, generator=gen)
    t_grid = [0.9, 0.5, 0.1]
    out.update(theta=npy(theta), x=npy(x), t_grid=np.asarray(t_grid, np.float32), prior_kind=np.str_(prior_kind),
               traj_z0=npy(z0), traj_z=npy(z), traj_steps=np.int64(steps), lsteps=np.int64(lsteps),
               dist_cov_est=npy(spd_covs(gen, x.shape[0], m)))
    for k, v in prior_params.items():
        out[f"prior.{k}"] = npy(v)
    extra_kwargs = extra_kwargs or {}
    for k, v in extra_kwargs.items():
        out[f"kw.{k}"] = npy(v)
    if traj_x is not None:
        out["traj_x"] = npy(traj_x)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        net = copy.deepcopy(net32).to(dt)
        prior, fn = make_prior(prior_kind, m, dt, net, **prior_params)
        kw = {k: v.to(dt) for k, v in extra_kwargs.items()}
        xx = x.to(dt)
        xt = xx if traj_x is None else traj_x.to(dt)
        kwt = kw if traj_x is None else {}
        out[f"{tag}.geffner_score"] = np.stack([
            npy(net.factorized_score_geffner(theta.to(dt), xx, torch.tensor(tv, dtype=torch.float32).to(dt), fn, **kw))
            for tv in t_grid])
        out[f"{tag}.langevin.traj"] = npy(run_sampler(net, "annealed_langevin_geffner", xt, z0.to(dt), z.to(dt),
                                                      prior_score_fn=fn, steps=steps, lsteps=lsteps, tau=0.5,
                                                      theta_clipping_range=(-3, 3), **kwt))
        out[f"{tag}.tamed.traj"] = npy(run_sampler(net, "predictor_corrector", xt, z0.to(dt), z.to(dt), steps=steps,
                                                   prior_score_fun=fn, lsteps=lsteps, r=0.5, predictor_type="id",
                                                   theta_clipping_range=(-3, 3), **kwt))
        out[f"{tag}.pc.traj"] = npy(run_sampler(net, "predictor_corrector", xt, z0.to(dt), z.to(dt), steps=steps,
                                                prior_score_fun=fn, lsteps=lsteps, r=0.3, eta=0.7, predictor_type="ddim",
                                                **kwt))
        if traj_x is None:
            out[f"{tag}.detgef.traj"] = npy(run_sampler(net, "deterministic_gaussian_geffner", xx, z0.to(dt), z.to(dt),
                                                        prior_score_fn=fn, steps=steps))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def cfg_case(name, net32, x_raw, N, seed, steps):
    """Classifier-free-guidance branches for PF_NSE (the only class for which they run in the reference):
    factorized_score(GAUSS, clf_free_guidance=True), factorized_score_geffner(clf_free_guidance=True), ddim."""
    out = {}
    net_arrays(net32, out)
    gen = torch.Generator().manual_seed(seed)
    m = net32.zeros.shape[0]
    xs, ns = net32._create_pf_data(x_raw)
    K = xs.shape[0]
    theta = torch.randn(N, m, generator=gen)
    cov_est = spd_covs(gen, K, m)
    cov_prior = spd_covs(gen, 1, m) * 3
    z0 = torch.randn(N, m, generator=gen)
    z = torch.randn(steps, N, m, generator=gen)
    t_grid = [0.9, 0.5, 0.1]
    out.update(theta=npy(theta), x=npy(xs), traj_x=npy(x_raw), dist_cov_est=npy(cov_est), dist_cov_est_prior=npy(cov_prior),
               t_grid=np.asarray(t_grid, np.float32), traj_z0=npy(z0), traj_z=npy(z), traj_steps=np.int64(steps),
               prior_kind=np.str_("cfg"))
    out["kw.n"] = npy(ns)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        net = copy.deepcopy(net32).to(dt)
        fs, gs = [], []
        for tv in t_grid:
            t = torch.tensor(tv, dtype=torch.float32).to(dt)
            fs.append(npy(net.factorized_score(theta.to(dt), xs.to(dt), t, None, None, dist_cov_est=cov_est.to(dt),
                                               dist_cov_est_prior=cov_prior.to(dt), cov_mode="GAUSS",
                                               clf_free_guidance=True, n=ns.to(dt))))
            gs.append(npy(net.factorized_score_geffner(theta.to(dt), xs.to(dt), t, None, clf_free_guidance=True,
                                                       n=ns.to(dt))))
        out[f"{tag}.cfg.factorized_score"] = np.stack(fs)
        out[f"{tag}.cfg.geffner_score"] = np.stack(gs)
        out[f"{tag}.cfg.traj"] = npy(run_ddim(net, x_raw.to(dt), z0.to(dt), z.to(dt), steps=steps, eta=0.8,
                                               prior=None, prior_score_fn=None, dist_cov_est=cov_est.to(dt),
                                               dist_cov_est_prior=cov_prior.to(dt), cov_mode="GAUSS",
                                               clf_free_guidance=True))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def toy_case(name, DIM, n_obs, eps_max, N, seed, steps, eta, lang_steps=10, lsteps_g=3):
    """Config 1: the Gaussian-Gaussian toy of gen_gaussian_gaussian.py:27-96 driven on CPU -- reference `nse_toy.NSE`,
    `FakeFNet(real_eps, EpsilonNet(DIM), eps)`, triple-returning prior score; GAUSS + JAC per-call and trajectories,
    plus the paired covariance pre-pass (gen_gaussian_gaussian.py:143-160)."""
    import embedding_nets as ref_emb
    import nse_toy as ref_toy
    from tasks.toy_examples.data_generators import Gaussian_Gaussian_mD
    out = {}
    torch.manual_seed(seed)
    means = torch.rand(DIM) * 20 - 10
    stds = torch.rand(DIM) * 25 + 0.1
    task = Gaussian_Gaussian_mD(dim=DIM, means=means, stds=stds)
    prior = task.prior
    theta_true = prior.sample(sample_shape=(1,))
    x_obs = torch.cat([task.simulator(theta_true).reshape(1, -1) for _ in range(n_obs)], dim=0)
    eps_net32 = ref_emb.EpsilonNet(DIM)
    gen = torch.Generator().manual_seed(seed + 100)
    theta = prior.loc + prior.covariance_matrix.diag().sqrt() * torch.randn(N, DIM, generator=gen) * 0.3
    cov_est = spd_covs(gen, n_obs, DIM) * 4
    z0 = torch.randn(N, DIM, generator=gen)
    z = torch.randn(steps, N, DIM, generator=gen)
    Np = 3
    z0p = torch.randn(Np * n_obs, DIM, generator=gen)
    zp = torch.randn(steps, Np * n_obs, DIM, generator=gen)
    # Langevin / Geffner baselines of the toy script (gen_gaussian_gaussian.py:205-238); drawn after the arrays above so
    # that the earlier fixture contents do not change
    z_lang = torch.randn(lang_steps * lsteps_g, N, DIM, generator=gen)
    z_det = torch.randn(steps, N, DIM, generator=gen)
    t_grid = [0.9, 0.5, 0.2, 0.05]
    mlp_arrays("net", eps_net32.net, out)
    out.update(lang_z=npy(z_lang), lang_steps=np.int64(lang_steps), lang_lsteps=np.int64(lsteps_g), lang_tau=np.float64(0.5),
               det_z=npy(z_det))
    out.update(theta=npy(theta), x=npy(x_obs), dist_cov_est=npy(cov_est), t_grid=np.asarray(t_grid, np.float32),
               prior_kind=np.str_("gaussian"), theta_dim=np.int64(DIM), eps_max=np.float64(eps_max),
               traj_z0=npy(z0), traj_z=npy(z), traj_steps=np.int64(steps), traj_eta=np.float64(eta),
               pair_z0=npy(z0p), pair_z=npy(zp), pair_reps=np.int64(Np))
    out["prior.loc"], out["prior.cov"] = npy(prior.loc), npy(prior.covariance_matrix)
    out["lik_cov"] = npy(task.simulator_cov)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        score_net = ref_toy.NSE(theta_dim=DIM, x_dim=DIM)
        idm = torch.eye(DIM, dtype=dt)
        inv_lik = torch.linalg.inv(task.simulator_cov.to(dt))
        inv_prior = torch.linalg.inv(prior.covariance_matrix.to(dt))
        posterior_cov = torch.linalg.inv(inv_lik + inv_prior)
        inv_prior_prior = inv_prior @ prior.loc.to(dt)

        def real_eps(theta, x, t):  # gen_gaussian_gaussian.py:59-72
            posterior_cov_diff = (posterior_cov[None] * score_net.alpha(t)[..., None]
                                  + (1 - score_net.alpha(t))[..., None] * idm[None])
            posterior_mean_0 = (posterior_cov @ (inv_prior_prior[:, None] + inv_lik @ x.mT)).mT
            posterior_mean_diff = (score_net.alpha(t) ** 0.5) * posterior_mean_0
            score = -(torch.linalg.inv(posterior_cov_diff) @ (theta - posterior_mean_diff)[..., None])[..., 0]
            return -score_net.sigma(t) * score

        score_net.net = ref_emb.FakeFNet(real_eps_fun=real_eps, eps_net=copy.deepcopy(eps_net32).to(dt), eps_net_max=eps_max)
        score_net.to(dt)
        prior_score_fn = ref_vp.get_vpdiff_gaussian_score(prior.loc.to(dt), prior.covariance_matrix.to(dt), score_net)

        def prior_score(theta, t):  # gen_gaussian_gaussian.py:86-96
            mean_prior_0_t = ref_tps.mean_backward(theta, t, prior_score_fn, score_net)
            prec_prior_0_t = ref_tps.prec_matrix_backward(t, prior.covariance_matrix.to(dt), score_net).repeat(
                theta.shape[0], 1, 1)
            return prior_score_fn(theta, t), mean_prior_0_t, prec_prior_0_t

        th, xx, ce = theta.to(dt), x_obs.to(dt), cov_est.to(dt)
        for mode in ("GAUSS", "JAC"):
            fs, sc = [], []
            for tv in t_grid:
                t = torch.tensor(tv, dtype=torch.float32).to(dt)
                fs.append(npy(score_net.factorized_score(th, xx, t, prior_score, dist_cov_est=ce, cov_mode=mode)))
                if mode == "GAUSS":
                    sc.append(npy(ref_toy.tweedies_approximation(x=xx, theta=th, t=t, score_fn=score_net.score, nse=score_net,
                                                                 dist_cov_est=ce, mode="GAUSS")[2]))
            out[f"{tag}.{mode}.factorized_score"] = np.stack(fs)
            if mode == "GAUSS":
                out[f"{tag}.scores"] = np.stack(sc)
            it = iter(z.to(dt))
            FakeDiagNormal.z0 = z0.to(dt)
            with mock.patch.object(ref_toy, "DiagNormal", FakeDiagNormal), \
                    mock.patch("torch.randn_like", side_effect=lambda a: next(it).to(a)):
                res = score_net.ddim(shape=(N,), x=xx, eta=eta, steps=steps, prior_score_fn=prior_score,
                                     dist_cov_est=ce, cov_mode=mode)
            out[f"{tag}.{mode}.traj"] = npy(res)
        # annealed Langevin (nse_toy.py:474-508) and the deterministic Gaussian sampler (nse_toy.py:510-556)
        it = iter(z_lang.to(dt))
        FakeDiagNormal.z0 = z0.to(dt)
        with mock.patch.object(ref_toy, "DiagNormal", FakeDiagNormal), \
                mock.patch("torch.randn_like", side_effect=lambda a: next(it).to(a)), torch.no_grad():
            res = score_net.annealed_langevin_geffner(shape=(N,), x=xx, prior_score_fn=prior_score, lsteps=lsteps_g,
                                                      tau=0.5, steps=lang_steps)
        out[f"{tag}.langevin.traj"] = npy(res)
        it = iter(z_det.to(dt))
        FakeDiagNormal.z0 = z0.to(dt)
        with mock.patch.object(ref_toy, "DiagNormal", FakeDiagNormal), \
                mock.patch("torch.randn_like", side_effect=lambda a: next(it).to(a)), torch.no_grad():
            res = score_net.deterministic_gaussian_geffner(shape=(N,), x=xx, prior_score_fn=prior_score, steps=steps)
        out[f"{tag}.det_geffner.traj"] = npy(res)
        # paired covariance pre-pass: every sample carries its own observation (gen_gaussian_gaussian.py:143-156)
        xp = xx[None, :].repeat(Np, 1, 1).reshape(Np * n_obs, -1)
        it = iter(zp.to(dt))
        FakeDiagNormal.z0 = z0p.to(dt)
        with mock.patch.object(ref_toy, "DiagNormal", FakeDiagNormal), \
                mock.patch("torch.randn_like", side_effect=lambda a: next(it).to(a)):
            res = score_net.ddim(shape=(Np * n_obs,), x=xp, steps=steps, eta=0.5)
        out[f"{tag}.pair.traj"] = npy(res)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def contexts_case(name, net32, x_ctx, prior_kind, prior_params, S, seed, steps, eta, modes=("GAUSS", "JAC")):
    """Batched contexts: what `vmap(sampler, randomness="different")(context, cov_est)` of
    jrnnm_posterior_estimation.py:286-294 computes, context by context under replayed noise, and the covariance
    pre-pass `vmap(lambda x: ddim(shape=(S,), x=x, steps, eta))` of sbibm_posterior_estimation.py:306-314 (plain
    single-observation score).  x_ctx: [B, n, d]."""
    m = net32.zeros.shape[0]
    B, n, _ = x_ctx.shape
    gen = torch.Generator().manual_seed(seed)
    cov_est = spd_covs(gen, B * n, m).reshape(B, n, m, m)
    z0 = torch.randn(B * S, m, generator=gen)
    z = torch.randn(steps, B * S, m, generator=gen)
    out = dict(x=npy(x_ctx), dist_cov_est=npy(cov_est), traj_z0=npy(z0), traj_z=npy(z), traj_steps=np.int64(steps),
               traj_eta=np.float64(eta), prior_kind=np.array(prior_kind), samples_per_context=np.int64(S))
    for k, v in prior_params.items():
        out["prior." + k] = npy(v)
    net_arrays(net32, out)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        net = copy.deepcopy(net32).to(dt)
        prior, fn = make_prior(prior_kind, m, dt, net, **prior_params)
        for mode in modes:
            res = []
            for c in range(B):
                sl = slice(c * S, (c + 1) * S)
                res.append(run_ddim(net, x_ctx[c].to(dt), z0[sl].to(dt), z[:, sl].to(dt), steps=steps, eta=eta, prior=prior,
                                    prior_type=prior_kind, prior_score_fn=fn, dist_cov_est=cov_est[c].to(dt), cov_mode=mode))
            out[f"{tag}.{mode}.traj"] = npy(torch.stack(res))  # [B, S, m]
        res = []
        for c in range(B):  # one observation per context: x[c, 0]
            sl = slice(c * S, (c + 1) * S)
            res.append(run_ddim(net, x_ctx[c, 0].to(dt), z0[sl].to(dt), z[:, sl].to(dt), steps=steps, eta=eta))
        out[f"{tag}.single.traj"] = npy(torch.stack(res))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and "traj" in k})


def make_prior(kind, m, dt, net, low=None, high=None, loc=None, cov=None):
    if kind == "uniform":
        low, high = low.to(dt), high.to(dt)
        prior = torch.distributions.Uniform(low, high, validate_args=False)
        fn = ref_vp.get_vpdiff_uniform_score(low, high, net)
    else:
        loc, cov = loc.to(dt), cov.to(dt)
        prior = torch.distributions.MultivariateNormal(loc, cov)
        fn = ref_vp.get_vpdiff_gaussian_score(loc, cov, net)
    return prior, fn


def per_call_case(name, net32, x, prior_kind, prior_params, N, seed, extra_kwargs=None, do_jac=True,
                  t_grid=T_GRID, traj=None, traj_x=None):
    """Per-call rows for fp32 and fp64 + optional trajectories."""
    out = {}
    net_arrays(net32, out)
    gen = torch.Generator().manual_seed(seed)
    m = net32.zeros.shape[0]
    n = x.shape[0]
    theta = torch.randn(N, m, generator=gen)
    cov_est = spd_covs(gen, n, m)
    out.update(theta=npy(theta), x=npy(x), dist_cov_est=npy(cov_est), t_grid=np.asarray(t_grid, np.float32),
               prior_kind=np.str_(prior_kind))
    for k, v in prior_params.items():
        out[f"prior.{k}"] = npy(v)
    extra_kwargs = extra_kwargs or {}
    for k, v in extra_kwargs.items():
        out[f"kw.{k}"] = npy(v)

    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        net = copy.deepcopy(net32).to(dt)
        prior, fn = make_prior(prior_kind, m, dt, net, **prior_params)
        th, xx, ce = theta.to(dt), x.to(dt), cov_est.to(dt)
        kw = {k: v.to(dt) for k, v in extra_kwargs.items()}
        for mode in ("GAUSS", "JAC") if do_jac else ("GAUSS",):
            fs, sc, pr = [], [], []
            for tv in t_grid:
                t = torch.tensor(tv, dtype=torch.float32).to(dt)  # the ddim grid is fp32-valued
                prec, _, score = net.tweedies_approximator(
                    x=xx, theta=th, nse=net, t=t, score_fn=partial(net.score, **kw), dist_cov_est=ce, mode=mode)
                sc.append(npy(score))
                pr.append(npy(prec[:2]))  # two samples are enough to pin the precisions
                fs.append(npy(net.factorized_score(th, xx, t, fn, prior, dist_cov_est=ce, cov_mode=mode,
                                                   prior_type=prior_kind, **kw)))
            out[f"{tag}.{mode}.factorized_score"] = np.stack(fs)
            out[f"{tag}.{mode}.prec"] = np.stack(pr)
            if mode == "GAUSS":
                out[f"{tag}.scores"] = np.stack(sc)
        ps, pp = [], []
        for tv in t_grid:
            t = torch.tensor(tv, dtype=torch.float32).to(dt)
            ps.append(npy(fn(th, t)))
            if float(net.alpha(t)) ** 0.5 > 0.5:  # the reference only evaluates it inside the gate (nse.py:643)
                pp.append(npy(ref_tps.tweedies_approximation_prior(th, t, fn, nse=net)[0]))
            else:
                pp.append(np.full((N, m, m), np.nan, npy(th).dtype))
        out[f"{tag}.prior_score"] = np.stack(ps)
        out[f"{tag}.prior_prec"] = np.stack(pp)

    if traj is not None:
        Nt, steps, eta = traj["N"], traj["steps"], traj["eta"]
        z0 = torch.randn(Nt, m, generator=gen)
        z = torch.randn(steps, Nt, m, generator=gen)
        out.update(traj_z0=npy(z0), traj_z=npy(z), traj_steps=np.int64(steps), traj_eta=np.float64(eta))
        if traj_x is not None:
            out["traj_x"] = npy(traj_x)
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            net = copy.deepcopy(net32).to(dt)
            prior, fn = make_prior(prior_kind, m, dt, net, **prior_params)
            kw = {k: v.to(dt) for k, v in extra_kwargs.items()}
            for mode in traj["modes"]:
                clip = traj.get("clip", (None, None))
                xt = x if traj_x is None else traj_x  # PF_NSE.ddim splits the raw observations itself
                kwt = kw if traj_x is None else {}
                res = run_ddim(net, xt.to(dt), z0.to(dt), z.to(dt), steps=steps, eta=eta, prior=prior,
                               prior_type=prior_kind, prior_score_fn=fn, dist_cov_est=cov_est.to(dt),
                               cov_mode=mode, theta_clipping_range=clip, **kwt)
                out[f"{tag}.{mode}.traj"] = npy(res)
        if "clip" in traj:
            out["traj_clip"] = np.asarray(traj["clip"], np.float64)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and "traj" in k})


def replay_noise(seed, steps, N, m):
    """Injected noise as a pure function of a seed (numpy PCG64: stable across versions), so that fixtures of long
    trajectories store the seed instead of [steps, N, m] floats; the tests regenerate the same arrays."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((N, m)).astype(np.float32), rng.standard_normal((steps, N, m)).astype(np.float32)


def euler_case(name, base, net32, x, prior_kind, prior_params, cov_est, N, seed):
    """`euler_sde_sampler` on `diffused_tall_posterior_score` (tall_posterior_sampler.py:158-281; the call of
    sbibm_posterior_estimation.py:347-372) under replayed noise, GAUSS and JAC, fp32 and fp64.  The net, observations,
    prior and dist_cov_est are those of the fixture `base`."""
    m = net32.zeros.shape[0]
    z0, z = replay_noise(seed, 999, N, m)
    out = dict(base=np.str_(base), seed=np.int64(seed), N=np.int64(N), checkpoints=np.asarray([100, 500, 900, 999]))
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        net = copy.deepcopy(net32).to(dt)
        prior, fn = make_prior(prior_kind, m, dt, net, **prior_params)
        for mode in ("GAUSS", "JAC"):
            score_fn = partial(ref_tps.diffused_tall_posterior_score, prior=prior, prior_type=prior_kind, prior_score_fn=fn,
                               x_obs=x.to(dt), nse=net, dist_cov_est=cov_est.to(dt), cov_mode=mode)
            it = iter(torch.as_tensor(z).to(dt))
            with mock.patch("torch.randn", side_effect=lambda *a, **k: torch.as_tensor(z0).to(dt)), \
                    mock.patch("torch.randn_like", side_effect=lambda a: next(it).to(a)):
                theta, hist = ref_tps.euler_sde_sampler(score_fn, N, dim_theta=m, beta=net.beta, device="cpu")
            out[f"{tag}.{mode}.theta"] = npy(theta)
            out[f"{tag}.{mode}.hist"] = np.stack([npy(hist[k]) for k in (100, 500, 900, 999)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and "theta" in k})


def jac_dist_case(name, base, net32, x, prior_kind, prior_params, N, seed, steps=400, eta=0.8, clip=(None, None)):
    """Reference sample SETS of a complete JAC trajectory (400 steps, eta 0.8: sbibm_posterior_estimation.py:330-334,
    617-619) in fp32 and fp64 under the same replayed noise: the distribution-level gate of SURVEY.md section 8(c).
    The net / observations / prior are those of the fixture `base`; only the samples and the seed are stored."""
    m = net32.zeros.shape[0]
    z0, z = replay_noise(seed, steps, N, m)
    out = dict(base=np.str_(base), seed=np.int64(seed), N=np.int64(N), steps=np.int64(steps), eta=np.float64(eta),
               clip=np.asarray([np.nan if c is None else c for c in clip], np.float64))
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        net = copy.deepcopy(net32).to(dt)
        prior, fn = make_prior(prior_kind, m, dt, net, **prior_params)
        res = run_ddim(net, x.to(dt), torch.as_tensor(z0).to(dt), torch.as_tensor(z).to(dt), steps=steps, eta=eta,
                       prior=prior, prior_type=prior_kind, prior_score_fn=fn, cov_mode="JAC", theta_clipping_range=clip)
        out[f"{tag}.samples"] = npy(res).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    a, b = out["f32.samples"], out["f64.samples"]
    print("wrote", name, a.shape, "nan rows:", int(np.isnan(a).any(1).sum()), int(np.isnan(b).any(1).sum()),
          "max |f32 - f64|:", float(np.nanmax(np.abs(a - b))))


def load_sbibm(task, sub="n_train_30000_bs_256_n_epochs_5000_lr_0.0001", pf=None):
    d = os.path.join(REF, "results", "sbibm", task, sub)
    if pf:
        d = os.path.join(d, pf)
    net = torch.load(os.path.join(d, "score_network.pkl"), weights_only=False, map_location="cpu")
    net.net_type = "default"
    if pf is None:
        net.tweedies_approximator = ref_tps.tweedies_approximation
    stats = torch.load(os.path.join(d, "train_means_stds_dict.pkl"), weights_only=False)
    return net.eval(), stats


def sbibm_obs(task, stats, n, log_space=False):
    x = torch.load(os.path.join(REF, "tasks", "sbibm", "data", task, "reference_observations",
                                "x_obs_100_num_1.pkl"), weights_only=False)[:n]
    if log_space:
        x = torch.log(x)
    xn = (x - stats["x_train_mean"]) / stats["x_train_std"]
    return torch.nan_to_num(xn, nan=0.0, posinf=0.0, neginf=0.0).float()  # sbibm_posterior_estimation.py:179-184


def slcp_prior(stats):
    low = (torch.full((5,), -3.0) - stats["theta_train_mean"]) / stats["theta_train_std"] * 2
    high = (torch.full((5,), 3.0) - stats["theta_train_mean"]) / stats["theta_train_std"] * 2
    return dict(low=low, high=high)  # sbibm_posterior_estimation.py:188-189


def lognormal_prior(stats, loc, scale):
    loc_n = (loc - stats["theta_train_mean"]) / stats["theta_train_std"]
    cov = torch.diag_embed(scale.square())
    cov_n = torch.diag(1 / stats["theta_train_std"]) @ cov @ torch.diag(1 / stats["theta_train_std"])
    return dict(loc=loc_n, cov=cov_n)  # sbibm_posterior_estimation.py:197-208


def main(filt=None):
    want = lambda name: filt is None or filt in name  # noqa: E731
    torch.manual_seed(0)

    if want("slcp"):
        net, st = load_sbibm("slcp")
        x = sbibm_obs("slcp", st, 6)
        per_call_case("slcp_n6", net, x, "uniform", slcp_prior(st), N=16, seed=11,
                      traj=dict(N=24, steps=40, eta=1.0, modes=("GAUSS",)))
    if want("slcp30"):
        net, st = load_sbibm("slcp")
        x = sbibm_obs("slcp", st, 30)
        per_call_case("slcp30_n30", net, x, "uniform", slcp_prior(st), N=8, seed=12, t_grid=[0.5, 0.1],
                      traj=dict(N=16, steps=1000, eta=1.0, modes=("GAUSS",)))
    if want("lv"):
        net, st = load_sbibm("lotka_volterra")
        x = sbibm_obs("lotka_volterra", st, 6, log_space=True)
        pp = lognormal_prior(st, torch.tensor([-0.125, -3.0, -0.125, -3.0]), torch.full((4,), 0.5))
        per_call_case("lv_n6", net, x, "gaussian", pp, N=16, seed=13,
                      traj=dict(N=24, steps=40, eta=1.0, modes=("GAUSS",)))
    if want("sir"):
        net, st = load_sbibm("sir")
        x = sbibm_obs("sir", st, 8, log_space=False)
        pp = lognormal_prior(st, torch.log(torch.tensor([0.4, 0.125])), torch.tensor([0.5, 0.2]))
        per_call_case("sir_n8", net, x, "gaussian", pp, N=16, seed=14,
                      traj=dict(N=24, steps=400, eta=0.8, modes=("GAUSS",)))
    if want("pf"):
        net, st = load_sbibm("slcp", sub="n_train_10000_bs_256_n_epochs_5000_lr_0.0001", pf="pf_n_max_3")
        st = dict(st, x_train_mean=st["x_train_mean"][0], x_train_std=st["x_train_std"][0])  # stats are [n_max, d]
        x = sbibm_obs("slcp", st, 8)
        xs, ns = net._create_pf_data(x)  # [3,3,8], [3,1]
        out_x = xs
        per_call_case("pf_slcp_nmax3_n8", net, out_x, "uniform", slcp_prior(st), N=12, seed=15,
                      extra_kwargs=dict(n=ns), do_jac=False, t_grid=[0.9, 0.5, 0.1, 0.01],
                      traj=dict(N=16, steps=30, eta=1.0, modes=("GAUSS",)), traj_x=x)
    if want("jrnnm"):
        d = os.path.join(REF, "results", "jrnnm", "4d", "n_train_50000_n_epochs_5000_lr_0.001")
        net = torch.load(os.path.join(d, "score_network.pkl"), weights_only=False, map_location="cpu").eval()
        net.net_type = "default"
        net.tweedies_approximator = ref_tps.tweedies_approximation
        st = torch.load(os.path.join(d, "train_means_stds_dict.pkl"), weights_only=False)
        x = torch.load(os.path.join(REF, "results", "jrnnm", "x_obs_100_[135.0, 220.0, 2000.0, 0.0].pkl"),
                       weights_only=False)[:5]
        x = ((x - st["x_train_mean"]) / st["x_train_std"]).float()  # jrnnm_posterior_estimation.py:136
        lo = torch.tensor([10.0, 50.0, 100.0, -20.0])  # tasks/jrnnm/prior.py:4-9
        hi = torch.tensor([250.0, 500.0, 5000.0, 20.0])
        low = (lo - st["theta_train_mean"]) / st["theta_train_std"] * 2  # jrnnm_posterior_estimation.py:147-148
        high = (hi - st["theta_train_mean"]) / st["theta_train_std"] * 2
        per_call_case("jrnnm4d_n5", net, x, "uniform", dict(low=low, high=high), N=10, seed=16,
                      t_grid=[0.9, 0.5, 0.29, 0.1, 0.01], traj=dict(N=12, steps=30, eta=1.0, modes=("GAUSS",)))
    if want("contexts"):
        d = os.path.join(REF, "results", "jrnnm", "4d", "n_train_50000_n_epochs_5000_lr_0.001")
        net = torch.load(os.path.join(d, "score_network.pkl"), weights_only=False, map_location="cpu").eval()
        net.net_type = "default"
        net.tweedies_approximator = ref_tps.tweedies_approximation
        st = torch.load(os.path.join(d, "train_means_stds_dict.pkl"), weights_only=False)
        x = torch.load(os.path.join(REF, "results", "jrnnm", "x_obs_100_[135.0, 220.0, 2000.0, 0.0].pkl"),
                       weights_only=False)[:12]
        x = ((x - st["x_train_mean"]) / st["x_train_std"]).float().reshape(3, 4, -1)  # 3 contexts of 4 observations
        lo = torch.tensor([10.0, 50.0, 100.0, -20.0])
        hi = torch.tensor([250.0, 500.0, 5000.0, 20.0])
        low = (lo - st["theta_train_mean"]) / st["theta_train_std"] * 2
        high = (hi - st["theta_train_mean"]) / st["theta_train_std"] * 2
        contexts_case("contexts_jrnnm4d_b3_n4", net, x, "uniform", dict(low=low, high=high), S=5, seed=51, steps=12, eta=0.8)
    if want("cfg"):
        net, st = load_sbibm("slcp", sub="n_train_10000_bs_256_n_epochs_5000_lr_0.0001", pf="pf_n_max_3")
        st = dict(st, x_train_mean=st["x_train_mean"][0], x_train_std=st["x_train_std"][0])
        cfg_case("cfg_pf_slcp_n8", net, sbibm_obs("slcp", st, 8), N=10, seed=41, steps=16)
    if want("toy"):
        toy_case("toy_gauss_d4_n5", DIM=4, n_obs=5, eps_max=1e-1, N=12, seed=31, steps=20, eta=0.5)
        toy_case("toy_gauss_d10_n30", DIM=10, n_obs=30, eps_max=1e-2, N=8, seed=32, steps=64, eta=0.2, lang_steps=200, lsteps_g=2)  # dt = 2^-6: the fp64 run of the reference is NaN when fp32(t) - dt < 0 (toy alpha has a linear term)
    if want("toywide"):  # DIM 16 / 32 of gen_gaussian_gaussian.py:24 (generic path: torch scores + native kernels 3/4)
        toy_case("toy_gauss_d16_n8", DIM=16, n_obs=8, eps_max=1e-2, N=6, seed=33, steps=16, eta=0.5, lang_steps=8, lsteps_g=2)
        toy_case("toy_gauss_d32_n4", DIM=32, n_obs=4, eps_max=1e-2, N=6, seed=34, steps=16, eta=0.5, lang_steps=8, lsteps_g=2)
    if want("langevin"):
        net, st = load_sbibm("sir")
        x = sbibm_obs("sir", st, 5, log_space=False)
        pp = lognormal_prior(st, torch.log(torch.tensor([0.4, 0.125])), torch.tensor([0.5, 0.2]))
        langevin_case("langevin_sir_n5", net, x, "gaussian", pp, N=10, seed=21, steps=12, lsteps=3)
        torch.manual_seed(4)
        net = ref_nse.NSE(theta_dim=3, x_dim=4, hidden_features=[64, 64, 64]).eval()
        gen = torch.Generator().manual_seed(7)
        x = torch.randn(6, 4, generator=gen)
        langevin_case("langevin_rand_uniform_n6", net, x, "uniform", dict(low=torch.full((3,), -2.5), high=torch.full((3,), 2.5)),
                      N=9, seed=22, steps=10, lsteps=2)
        net, st = load_sbibm("slcp", sub="n_train_10000_bs_256_n_epochs_5000_lr_0.0001", pf="pf_n_max_3")
        st = dict(st, x_train_mean=st["x_train_mean"][0], x_train_std=st["x_train_std"][0])
        x = sbibm_obs("slcp", st, 8)
        xs, ns = net._create_pf_data(x)
        langevin_case("langevin_pf_slcp_n8", net, xs, "uniform", slcp_prior(st), N=8, seed=23, steps=8, lsteps=2,
                      traj_x=x, extra_kwargs=dict(n=ns))
    if want("euler"):
        net, st = load_sbibm("slcp")
        x = sbibm_obs("slcp", st, 6)
        g0 = dict(np.load(os.path.join(OUT, "slcp_n6.npz")))
        euler_case("euler_slcp_n6", "slcp_n6", net, x, "uniform", slcp_prior(st), torch.as_tensor(g0["dist_cov_est"]), N=8, seed=61)
        net, st = load_sbibm("sir")
        x = sbibm_obs("sir", st, 8, log_space=False)
        pp = lognormal_prior(st, torch.log(torch.tensor([0.4, 0.125])), torch.tensor([0.5, 0.2]))
        g0 = dict(np.load(os.path.join(OUT, "sir_n8.npz")))
        euler_case("euler_sir_n8", "sir_n8", net, x, "gaussian", pp, torch.as_tensor(g0["dist_cov_est"]), N=8, seed=62)
    if want("jacdist"):
        net, st = load_sbibm("slcp")
        jac_dist_case("jacdist_slcp_n6", "slcp_n6", net, sbibm_obs("slcp", st, 6), "uniform", slcp_prior(st), N=256, seed=71,
                      clip=(-3.0, 3.0))
        net, st = load_sbibm("lotka_volterra")
        pp = lognormal_prior(st, torch.tensor([-0.125, -3.0, -0.125, -3.0]), torch.full((4,), 0.5))
        jac_dist_case("jacdist_lv_n6", "lv_n6", net, sbibm_obs("lotka_volterra", st, 6, log_space=True), "gaussian", pp,
                      N=256, seed=72)
        net, st = load_sbibm("sir")
        pp = lognormal_prior(st, torch.log(torch.tensor([0.4, 0.125])), torch.tensor([0.5, 0.2]))
        jac_dist_case("jacdist_sir_n8", "sir_n8", net, sbibm_obs("sir", st, 8, log_space=False), "gaussian", pp, N=256,
                      seed=73)
    if want("rand"):
        torch.manual_seed(3)
        net = ref_nse.NSE(theta_dim=3, x_dim=4, hidden_features=[64, 96]).eval()
        gen = torch.Generator().manual_seed(6)
        x = torch.randn(7, 4, generator=gen)
        per_call_case("rand_m3_n7", net, x, "gaussian", dict(loc=torch.tensor([0.3, -0.2, 0.1]),
                                                            cov=torch.tensor([[1.0, 0.2, 0.0], [0.2, 0.8, 0.1],
                                                                              [0.0, 0.1, 1.2]])),
                      N=9, seed=17, traj=dict(N=9, steps=25, eta=0.5, modes=("GAUSS", "JAC"), clip=(-3.0, 3.0)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
