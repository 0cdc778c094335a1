"""Tweedie-approximation helpers with the reference's signatures (tall_posterior_sampler.py:35-155).

These exist for API compatibility (the `nse.tweedies_approximator` hook of nse.py:588-596 and scripts that call
the helpers directly).  They *materialise* [N, n, m(, m)] tensors, which the fused sampling path never does:
`NSE.ddim` / `NSE.factorized_score` keep the per-pair quantities on chip.  The arithmetic still runs in the CUDA
kernels of libd4s -- scores and JAC precisions come from kernel 1/2 with its optional outputs switched on.
"""
from __future__ import annotations

import torch

from . import _cabi as abi
from . import engine


def _t_cpu(t):
    return torch.as_tensor(t).detach().to("cpu").reshape(1)


def mean_backward(theta, t, score_fn, nse, **kwargs):
    """Tweedie mean of p(theta_0 | theta_t): (theta + sigma^2 score) / sqrt(alpha) (tall_posterior_sampler.py:10-18).
    `score_fn(theta=, t=, **kwargs)` is any callable, e.g. a d4s.vp_diffused_priors object."""
    alpha_t, sigma_t = nse.alpha(t), nse.sigma(t)
    return (theta + sigma_t ** 2 * score_fn(theta=theta, t=t, **kwargs)) / alpha_t ** 0.5


def sigma_backward(t, dist_cov, nse):
    """Covariance of p(theta_0 | theta_t) for a Gaussian p(theta_0) = N(., dist_cov): inv(inv(Sigma) + alpha/sigma^2 I)
    (tall_posterior_sampler.py:21-32; the reference switches to a Woodbury form of the same matrix when
    sqrt(alpha) >= 0.5 -- one formula serves both here, the inverses run in fp64 in batched_inverse_kernel)."""
    return engine.batched_inverse(prec_matrix_backward(t, dist_cov, nse).reshape(-1, dist_cov.shape[-1], dist_cov.shape[-1])
                                  ).reshape(dist_cov.shape)


def prec_matrix_backward(t, dist_cov, nse):
    """inv(Sigma) + alpha/sigma^2 I for Sigma[*, m, m] (tall_posterior_sampler.py:35-43)."""
    if not dist_cov.is_cuda:
        raise abi.D4sError("d4s.prec_matrix_backward: CUDA tensors only (no CPU fallback)")
    m = dist_cov.shape[-1]
    tt = _t_cpu(t).to(torch.float64)
    aos2 = float(engine._scalar64(nse.alpha, tt) / engine._scalar64(nse.sigma, tt) ** 2)
    if m <= abi.MAX_THETA:
        inv = engine.batched_inverse(dist_cov.reshape(-1, m, m)).reshape(dist_cov.shape)
    else:  # (the register-resident inverse kernel stops at theta_dim 10; wider toy problems use cuSOLVER through torch)
        inv = torch.linalg.inv(dist_cov.to(torch.float64)).to(torch.float32)
    return inv + aos2 * torch.eye(m, device=dist_cov.device, dtype=inv.dtype)


def tweedies_approximation(x, theta, t, score_fn=None, nse=None, dist_cov_est=None, mode="JAC",
                           clip_mean_bounds=(None, None), partial_factorization=False, n=None):
    """(prec[N,n,m,m], mean[N,n,m], score[N,n,m]) of p(theta_0 | theta_t, x_j) (tall_posterior_sampler.py:46-135).

    `score_fn` is accepted for signature compatibility and ignored: the scores are those of `nse` itself
    (every call site of the reference passes `partial(nse.score, **kwargs)`).  PF_NSE subset sizes go in `n`
    (or are recovered from a `functools.partial(nse.score, n=...)`).
    """
    if mode not in ("GAUSS", "JAC"):
        raise NotImplementedError("Available methods are GAUSS, JAC")
    if n is None and score_fn is not None and getattr(score_fn, "keywords", None):
        n = score_fn.keywords.get("n")
    nse._require_cuda(theta, x)
    N, m = theta.shape
    tt = _t_cpu(t)
    from . import generic
    if not generic.fusable(nse):  # closure nets / theta_dim > 10: torch evaluates the grid (this API materialises anyway)
        toy = not hasattr(nse, "net_type")
        t_dev = torch.as_tensor(t, dtype=torch.float32, device=theta.device)
        th = theta.detach().to(torch.float32)
        t64 = tt.to(torch.float64)
        alpha_t, sigma_t = float(engine._scalar64(nse.alpha, t64)), float(engine._scalar64(nse.sigma, t64))
        if mode == "JAC":
            jac, scores = generic.jacobians_on_grid(nse, th, x, t_dev, toy)
            eye = torch.eye(m, device=theta.device)
            prec = torch.linalg.inv((sigma_t ** 2 / alpha_t) * (eye + sigma_t ** 2 * jac))
        else:
            scores = generic.scores_on_grid(nse, th, x, t_dev, toy)
            prec = prec_matrix_backward(t, dist_cov_est, nse)[None].expand(N, -1, -1, -1)
        mean = (th[:, None] + sigma_t ** 2 * scores) / math.sqrt(alpha_t)
        if clip_mean_bounds[0]:
            mean = mean.clip(*clip_mean_bounds)
        return prec, mean, scores
    setup = engine.build_setup(nse, x, N, tt, None, 0.0, engine.PriorSpec(abi.PRIOR_NONE), mode,
                               dist_cov_est if mode == "GAUSS" else None, n_sizes=n, weighted=False)
    th = theta.detach().to(torch.float32).contiguous()
    scores, prec = engine.score_partials(setup, th, 0, want_scores=True, want_prec=(mode == "JAC"))
    if mode == "GAUSS":
        prec = prec_matrix_backward(t, dist_cov_est, nse)[None].expand(N, -1, -1, -1)
    t64 = tt.to(torch.float64)
    alpha_t = float(engine._scalar64(nse.alpha, t64))
    sigma_t = float(engine._scalar64(nse.sigma, t64))
    mean = (th[:, None] + sigma_t ** 2 * scores) / math.sqrt(alpha_t)
    if clip_mean_bounds[0]:
        mean = mean.clip(*clip_mean_bounds)
    return prec, mean, scores


def tweedies_approximation_prior(theta, t, score_fn, nse, mode="vmap"):
    """(prec[N,m,m], mean[N,m], score[N,m]) for the diffused prior (tall_posterior_sampler.py:138-155).
    `score_fn` must be a d4s.vp_diffused_priors object (its parameters are needed for the closed forms)."""
    if mode != "vmap":
        raise NotImplementedError
    kind = getattr(score_fn, "kind", None)
    if kind not in ("uniform", "gaussian"):
        raise abi.D4sError("tweedies_approximation_prior needs a d4s.vp_diffused_priors score object")
    spec = score_fn._spec()
    N, m = theta.shape
    tt = _t_cpu(t).to(torch.float64)
    alpha_t = float(engine._scalar64(nse.alpha, tt))
    sigma_t = float(engine._scalar64(nse.sigma, tt))
    if kind == "uniform":
        score, jd = engine.prior_score_native(nse, theta, t, spec, want_jac=True)
        prec_diag = (alpha_t / sigma_t ** 2) / (1.0 + sigma_t ** 2 * jd)
        prec = torch.diag_embed(prec_diag)
    else:
        score = engine.prior_score_native(nse, theta, t, spec)
        eye = torch.eye(m, dtype=torch.float64)
        c_t = sigma_t ** 2 * eye + alpha_t * spec.cov
        jac = -torch.linalg.inv(c_t).T
        cov = (sigma_t ** 2 / alpha_t) * (eye + sigma_t ** 2 * jac)
        prec = torch.linalg.inv(cov).to(torch.float32).to(theta.device)[None].expand(N, -1, -1)
    mean = (theta + sigma_t ** 2 * score) / math.sqrt(alpha_t)
    return prec, mean, score


def diffused_tall_posterior_score(theta, t, prior, prior_score_fn, x_obs, nse, score_fn=None, dist_cov_est=None,
                                  cov_mode="JAC", prior_type="gaussian"):
    """Functional form of the tall score (tall_posterior_sampler.py:158-214) == `nse.factorized_score(...)`; the fused
    kernels run it in one call.  `score_fn` is accepted for signature compatibility (the scores are `nse`'s own)."""
    return nse.factorized_score(theta, x_obs, t, prior_score_fn, prior, dist_cov_est=dist_cov_est, cov_mode=cov_mode,
                                prior_type=prior_type)


def euler_sde_sampler(score_fn, nsamples, dim_theta, beta, device="cuda", debug=False, theta_clipping_range=(None, None),
                      _noise=None):
    """Euler-Maruyama integration of the backward SDE on linspace(1, 0, 1000) (tall_posterior_sampler.py:217-281):
    theta <- theta + (f - g^2 score) dt + g sqrt(|dt|) z with f = -0.5 beta(t) theta, g^2 = beta(t), dt < 0.
    `score_fn(theta, t)` is any callable (e.g. `partial(diffused_tall_posterior_score, ...)`, which itself runs on the
    fused kernels); the update runs in kernel 4.  Returns (theta, [theta_0, theta_1, ...] on the host) like the reference."""
    import ctypes as C
    if debug:
        raise NotImplementedError("debug=True returns the reference's internal tensors; not provided")
    lib = abi.require_cuda()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise abi.D4sError("d4s.euler_sde_sampler: CUDA device required (no CPU fallback)")
    time_pts = torch.linspace(1, 0, 1000)
    if _noise is not None:
        z0, z = _noise
        theta = z0.to(dev, torch.float32).contiguous().clone()
    else:
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        theta = engine.philox_normal(nsamples, dim_theta, seed, -1, 0, dev)
        z = None
    t64 = time_pts.to(torch.float64)
    dt = t64[1:] - t64[:-1]
    b = torch.as_tensor(beta(t64[:-1]), dtype=torch.float64)
    rows = torch.zeros(999, abi.SCHED_STRIDE, dtype=torch.float64)
    rows[:, abi.S_T] = t64[:-1]
    rows[:, abi.S_C_THETA] = 1.0 - 0.5 * b * dt     # theta + f dt
    rows[:, abi.S_C_SCORE] = -b * dt                # - g^2 score dt
    rows[:, abi.S_BRIDGE_STD] = (b * dt.abs()).sqrt()  # g sqrt(|dt|)
    rows[:, abi.S_CLIP] = 1.0
    sched = rows.to(torch.float32).to(dev)
    lo, hi = theta_clipping_range
    clip = (1.0, -1.0) if lo is None else (float(lo), float(hi))
    hist = torch.empty(1000, theta.shape[0], theta.shape[1], dtype=torch.float32, device=dev)  # one D2H copy at the end
    hist[0].copy_(theta)
    for i in range(999):
        score = score_fn(theta, time_pts[i].to(dev)).detach().to(torch.float32).contiguous()
        nz = None if z is None else z[i].to(dev, torch.float32).contiguous()
        with torch.cuda.device(dev):
            abi.check(lib.d4s_ddim_update(engine._ptr(theta), engine._ptr(score), theta.shape[0], theta.shape[1],
                                          engine._ptr(sched[i]), engine._ptr(nz), C.c_uint64(0 if z is not None else seed),
                                          i, 0, clip[0], clip[1], engine._stream(dev)))
        hist[i + 1].copy_(theta)
    return theta, list(hist.cpu().unbind(0))


import math  # noqa: E402  (kept at the bottom: only the helpers above use it)
