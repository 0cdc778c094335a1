"""Toy-experiment variant of the score estimator (the class of the reference's `nse_toy.py:198-625`, used by
`gen_gaussian_gaussian.py` / `gen_mixt_gauss_gaussian.py`).  Differences to `d4s.nse.NSE` that callers rely on:

* schedule `alpha(t) = exp(-0.5 (0.5 beta_d t^2 + beta_min t))`, beta in [0.1, 40] (nse_toy.py:305-339);
* `ddim(shape, x, steps=64, verbose=False, eta=1., **kwargs)`; only `x.shape[0] == shape[0]` selects the paired
  plain score (nse_toy.py:375);
* `factorized_score(theta, x_obs, t, prior_score_fn, dist_cov_est=None, cov_mode='JAC')` is always the
  precision-weighted solve and `prior_score_fn(theta, t)` returns the TRIPLE (score, mean, precision)
  (nse_toy.py:599-625, gen_gaussian_gaussian.py:86-96) -- build it with `gaussian_prior_triple(...)`;
* the network is `FakeFNet(AnalyticGaussianEps(...), EpsilonNet(DIM), eps_max)`: analytic eps + eps_max * tanh-MLP on
  [theta, x, t] with the RAW time as feature (embedding_nets.py:127-160).

The fused path handles exactly that network family: the MLP with its tanh head and the per-step affine form of the
analytic part run in the fp32 SIMT kernels (theta_dim <= 10).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
from torch import Size, Tensor

from . import _cabi as abi
from . import engine
from . import generic
from .embedding_nets import AnalyticGaussianEps, EpsilonNet, FakeFNet  # noqa: F401
from .nse import NSE as _FullNSE
from .nse import _use_graph_default
from .tall_posterior_sampler import mean_backward, prec_matrix_backward, sigma_backward  # noqa: F401  (nse_toy.py:20-50)
from .vp_diffused_priors import GaussianDiffusedScore


class GaussianPriorTriple:
    """`prior_score(theta, t) -> (score, mean_backward, precision_backward)` for a Gaussian prior, the closure of
    gen_gaussian_gaussian.py:86-96 as an object that exposes the prior parameters to the fused kernels."""
    kind = "gaussian"

    def __init__(self, mean, cov, nse):
        self.mean, self.cov = torch.as_tensor(mean), torch.as_tensor(cov)
        self.nse = nse
        self._score = GaussianDiffusedScore(mean, cov, nse)

    def _spec(self):
        return self._score._spec()

    def __call__(self, theta: Tensor, t: Tensor):
        from .tall_posterior_sampler import prec_matrix_backward
        score = self._score(theta, t)
        tt = torch.as_tensor(t).detach().to("cpu", torch.float64).reshape(1)
        a, sig = float(engine._scalar64(self.nse.alpha, tt)), float(engine._scalar64(self.nse.sigma, tt))
        mean = (theta + sig ** 2 * score) / math.sqrt(a)  # mean_backward, tall_posterior_sampler.py:10-18
        prec = prec_matrix_backward(t, self.cov.to(theta.device), self.nse).repeat(theta.shape[0], 1, 1)
        return score, mean, prec


def gaussian_prior_triple(mean, cov, nse) -> GaussianPriorTriple:
    return GaussianPriorTriple(mean, cov, nse)


class NSE(nn.Module):
    def __init__(self, theta_dim: int, x_dim: int, beta_min: float = 0.1, beta_max: float = 40, M: int = 1000,
                 eps_t: float = 1e-10, sampling_dist: str = "Uniform", **kwargs):
        super().__init__()
        self.eps_t = eps_t
        self.beta_min = beta_min
        self.beta_d = beta_max - beta_min
        self.M = M
        self.theta_dim, self.x_dim = theta_dim, x_dim
        self.net = None  # callers assign `score_net.net = FakeFNet(...)` (gen_gaussian_gaussian.py:75-77)
        self.embedding_nn_theta = nn.Identity()
        self.embedding_nn_x = nn.Identity()
        self.theta_emb_dim, self.x_emb_dim = theta_dim, x_dim
        self.register_buffer("zeros", torch.zeros(theta_dim))
        self.register_buffer("ones", torch.ones(theta_dim))
        self.register_buffer("freqs", torch.zeros(0))

    # ---- model / schedule (torch) -----------------------------------------------------------------------------
    def forward(self, theta: Tensor, x: Tensor, t: Tensor) -> Tensor:
        return self.net(theta, x, t)

    def score(self, theta, x, t):
        return -self(theta, x, t) / self.sigma(t)

    def beta(self, t: Tensor) -> Tensor:
        return self.beta_d * t + self.beta_min

    def f(self, t: Tensor) -> Tensor:
        return -0.5 * self.beta(t)

    def g(self, t: Tensor) -> Tensor:
        return torch.sqrt(self.beta(t))

    def alpha(self, t: Tensor) -> Tensor:
        return torch.exp(-0.5 * (0.5 * self.beta_d * (t ** 2) + self.beta_min * t))  # nse_toy.py:323-328

    def sigma(self, t: Tensor) -> Tensor:
        return torch.sqrt(1 - self.alpha(t) + math.exp(-16))  # nse_toy.py:334-339

    def mean_pred(self, theta: Tensor, score: Tensor, alpha_t: Tensor, **kwargs) -> Tensor:
        return (alpha_t ** (-0.5)) * (theta + (1 - alpha_t) * score)

    def bridge_mean(self, alpha_t, alpha_t_1, theta_t, theta_0, bridge_std):
        eps_hat = (theta_t - (alpha_t ** 0.5) * theta_0) / ((1 - alpha_t) ** 0.5)
        return (alpha_t_1 ** 0.5) * theta_0 + ((1 - alpha_t_1 - bridge_std ** 2) ** 0.5) * eps_hat

    # ---- what the kernels need to know about `net` ---------------------------------------------------------------
    def _d4s_net_description(self) -> dict:
        net = self.net
        if not (isinstance(net, FakeFNet) and isinstance(net.real_eps_fun, AnalyticGaussianEps)
                and isinstance(net.eps_net, EpsilonNet)):
            raise abi.D4sError("d4s.nse_toy.NSE fuses FakeFNet(AnalyticGaussianEps(...), EpsilonNet(DIM), eps_max) only; "
                               "an arbitrary Python closure as `real_eps_fun` cannot run inside a CUDA kernel")
        return dict(mlp=net.eps_net.net, emb_theta=None, out_act=1, out_scale=float(net.eps_net_max),
                    theta_cols=self.theta_dim, time_features=lambda t64: t64[..., None], time_width=1,
                    affine=net.real_eps_fun.tables)

    _require_cuda = staticmethod(_FullNSE._require_cuda)

    def _draw_seed(self) -> int:
        return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())

    # ---- samplers ---------------------------------------------------------------------------------------------------
    def ddim(self, shape: Size, x: Tensor, steps: int = 64, verbose: bool = False, eta: float = 1.0, **kwargs) -> Tensor:
        """nse_toy.py:372-387: paired plain score when `x.shape[0] == shape[0]` (the covariance pre-pass), else the
        precision-weighted tall score."""
        self._require_cuda(x)
        kw = dict(kwargs)
        noise, seed, off = kw.pop("_noise", None), kw.pop("_seed", None), kw.pop("_sample_offset", 0)
        N, m = int(shape[0]), self.theta_dim
        grid = engine.time_grid(steps)
        t, t_prev = grid[:-1], grid[:-1].to(torch.float64) - 1.0 / steps
        paired = x.shape[0] == N
        if paired:
            prior, cov_mode, cov_est = engine.PriorSpec(abi.PRIOR_NONE), "GAUSS", None
        else:
            prior_fn = kw.pop("prior_score_fn")
            cov_mode, cov_est = kw.pop("cov_mode", "JAC"), kw.pop("dist_cov_est", None)
            prior = self._prior_spec(prior_fn)
        if kw:
            raise TypeError(f"unexpected keyword arguments: {sorted(kw)}")
        if seed is None:
            seed = self._draw_seed()
        if noise is not None:
            z0, z = noise
            theta = z0.to(x.device, torch.float32).contiguous().clone()
        else:
            z = None
            theta = engine.philox_normal(N, m, seed, -1, off, x.device)
        if self._generic(prior):  # closure nets / closure priors / theta_dim > 10: torch scores + native kernels 3/4
            run = generic.GenericTall(self, x, N, t, t_prev, float(eta), prior, cov_mode, cov_est, toy=True, paired=paired,
                                      weighted=not paired, seed=seed, sample_offset=off)
            return run.run(theta, z)
        setup = engine.build_setup(self, x, N, t, t_prev, float(eta), prior, cov_mode, cov_est, paired=paired, seed=seed,
                                   sample_offset=off, weighted=not paired)
        return engine.run_trajectory(setup, theta, z, _use_graph_default())

    def _generic(self, prior) -> bool:
        """The fused kernels take FakeFNet(AnalyticGaussianEps, EpsilonNet, eps) with theta_dim <= 10 and a prior that exposes
        its parameters; everything else (the reference's `real_eps` closure, its triple-returning prior closure, DIM 16 / 32:
        gen_gaussian_gaussian.py:24, 59-72, 86-96) runs on the generic path (d4s/generic.py)."""
        return not generic.fusable(self) or not isinstance(prior, engine.PriorSpec)

    # ---- Langevin / Geffner baselines of the toy script (gen_gaussian_gaussian.py:205-238) ---------------------------
    _gammas = _FullNSE._gammas  # gamma_t = alpha(t) / alpha(t - time[-2]), alpha(t) on the last step (nse_toy.py:489-492)

    def _run_rows(self, shape: Size, x: Tensor, t_rows: Tensor, coef: dict, prior_score_fn, kw: dict) -> Tensor:
        """A trajectory whose steps are schedule rows (time + update coefficients) on the plain-sum score
        (1 - n) prior_score + sum_j score_j (nse_toy.py:503, 536-545)."""
        noise, seed, off = kw.pop("_noise", None), kw.pop("_seed", None), kw.pop("_sample_offset", 0)
        if kw:
            raise TypeError(f"unexpected keyword arguments: {sorted(kw)}")
        N, m = int(shape[0]), self.theta_dim
        if seed is None:
            seed = self._draw_seed()
        prior = self._prior_spec(prior_score_fn)
        if noise is not None:
            z0, z = noise
            theta = z0.to(x.device, torch.float32).contiguous().clone()
        else:
            z = None
            theta = engine.philox_normal(N, m, seed, -1, off, x.device)
        if self._generic(prior):
            run = generic.GenericTall(self, x, N, t_rows, None, 0.0, prior, "GAUSS", None, toy=True, weighted=False,
                                      seed=seed, sample_offset=off, coef=coef)
            return run.run(theta, z)
        setup = engine.build_setup(self, x, N, t_rows, None, 0.0, prior, "GAUSS", None,
                                   seed=seed, sample_offset=off, weighted=False, coef=coef)
        return engine.run_trajectory(setup, theta, z, _use_graph_default())

    def annealed_langevin_geffner(self, shape: Size, x: Tensor, prior_score_fn, steps: int = 64, lsteps: int = 5,
                                  tau: float = 0.5, verbose: bool = False, **kwargs) -> Tensor:
        """Annealed Langevin with the F-NPSE score (nse_toy.py:474-508): `steps` noise levels x `lsteps` updates
        theta += delta score + sqrt(2 delta) z, delta = tau (1 - gamma_t) / sqrt(gamma_t); no clipping, always the
        sum over the observations.  One fused launch."""
        self._require_cuda(x)
        if x.ndim != 2:
            raise ValueError("nse_toy.annealed_langevin_geffner expects x of shape [n_obs, x_dim] (nse_toy.py:495)")
        grid = engine.time_grid(steps)
        gamma = self._gammas(grid)
        delta = tau * (1 - gamma) / gamma.sqrt()
        rep = lambda v: v.repeat_interleave(lsteps)  # noqa: E731
        rows = steps * lsteps
        coef = dict(c_theta=torch.ones(rows, dtype=torch.float64), c_score=rep(delta), bstd=rep((2 * delta).sqrt()),
                    clip=torch.zeros(rows, dtype=torch.float64))
        return self._run_rows(shape, x, rep(grid[:-1]), coef, prior_score_fn, dict(kwargs))

    def deterministic_gaussian_geffner(self, shape: Size, x: Tensor, prior_score_fn, steps: int = 64,
                                       verbose: bool = False, **kwargs) -> Tensor:
        """Gaussian transition sampler of Geffner et al., Algorithm 2, with the continuous schedule (nse_toy.py:510-556):
        theta' = c_theta theta + var_t [S1 / sqrt(gamma) + (1 - n) prior_score] + sqrt(var_t) z."""
        self._require_cuda(x)
        if x.ndim != 2:
            raise ValueError("nse_toy.deterministic_gaussian_geffner expects x of shape [n_obs, x_dim] (nse_toy.py:531)")
        n = x.shape[0]
        grid = engine.time_grid(steps)
        g = self._gammas(grid)
        den = n - g * (n - 1)
        var = (1 - g) / den  # nse_toy.py:548
        coef = dict(c_theta=(n / g.sqrt() - (n - 1) * g.sqrt()) / den, c_score=var, s1w=1 / g.sqrt(), bstd=var.sqrt())
        return self._run_rows(shape, x, grid[:-1], coef, prior_score_fn, dict(kwargs))

    def _prior_spec(self, prior_score_fn):
        """`engine.PriorSpec` of a `gaussian_prior_triple(...)` object (closed forms inside the kernels), or the callable
        itself when it is an opaque closure like gen_gaussian_gaussian.py:86-96 (evaluated by torch, generic path)."""
        if getattr(prior_score_fn, "kind", None) == "gaussian" and hasattr(prior_score_fn, "_spec"):
            return prior_score_fn._spec()
        if callable(prior_score_fn):
            return prior_score_fn
        raise abi.D4sError("nse_toy.NSE needs a prior score function: `d4s.nse_toy.gaussian_prior_triple(mean, cov, nse)` "
                           "or a callable (theta, t) -> (score, mean, precision)")

    def factorized_score(self, theta, x_obs, t, prior_score_fn, dist_cov_est=None, cov_mode="JAC") -> Tensor:
        """nse_toy.py:599-625: always the precision-weighted combination (no alpha-gate; Gaussian prior triple)."""
        self._require_cuda(theta, x_obs)
        spec = self._prior_spec(prior_score_fn)
        tt = torch.as_tensor(t).detach().to("cpu").reshape(1)
        if self._generic(spec):
            run = generic.GenericTall(self, x_obs, theta.shape[0], tt, None, 0.0, spec, cov_mode, dist_cov_est, toy=True,
                                      always_weighted=True)
            return run.step(0, theta.detach().to(torch.float32).contiguous(), update=False, want_score=True)
        setup = engine.build_setup(self, x_obs, theta.shape[0], tt, None, 0.0, spec, cov_mode, dist_cov_est)
        th = theta.detach().to(torch.float32).contiguous()
        engine.score_partials(setup, th, 0)
        return engine.finalize(setup, th, 0)


class NSELoss(nn.Module):
    """Denoising score matching loss of the toy module (nse_toy.py:628-660 declares it; same objective as nse.py:708-751).
    Pure torch; present for import compatibility of the reference's toy drivers."""

    def __init__(self, estimator: NSE, eps_t: float = 1e-5):
        super().__init__()
        self.eps_t = eps_t
        self.estimator = estimator

    def forward(self, theta: Tensor, x: Tensor, **kwargs) -> Tensor:
        t = torch.rand(theta.shape[0], dtype=theta.dtype, device=theta.device) * (1 - self.eps_t) + self.eps_t
        noise = torch.randn_like(theta)
        theta_t = self.estimator.alpha(t).sqrt()[:, None] * theta + self.estimator.sigma(t)[:, None] * noise
        return (self.estimator(theta_t, x, t[:, None], **kwargs) - noise).square().mean()
