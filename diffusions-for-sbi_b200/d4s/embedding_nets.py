"""Toy score-network pieces of the reference's section-4.1/4.2 experiments (embedding_nets.py:127-160):
`EpsilonNet` (random perturbation MLP with a tanh head on [theta, x, t]) and `FakeFNet` (analytic eps + eps_max *
perturbation).  `AnalyticGaussianEps` is the closed-form noise predictor of the Gaussian-Gaussian toy task
(gen_gaussian_gaussian.py:59-72 builds it as a Python closure; a closure cannot be fused, so it is a module here
that also hands its per-step affine form  eps = M_t theta - G_t x - g_t  to the CUDA kernels)."""
from __future__ import annotations

import torch
import torch.nn as nn


class EpsilonNet(nn.Module):
    def __init__(self, DIM: int):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(2 * DIM + 1, 5 * DIM), nn.ReLU(), nn.Linear(5 * DIM, DIM), nn.Tanh())

    def forward(self, theta, x, t):
        return self.net(torch.cat((theta, x, t), dim=-1))


class AnalyticGaussianEps(nn.Module):
    """eps(theta, x, t) = sigma_t (alpha_t S_post + (1 - alpha_t) I)^-1 (theta - sqrt(alpha_t) S_post (S_pr^-1 mu_pr +
    S_lik^-1 x)) for a Gaussian prior N(mu_pr, S_pr) and likelihood N(theta, S_lik); S_post = (S_lik^-1 + S_pr^-1)^-1."""

    def __init__(self, prior_mean, prior_cov, lik_cov, nse):
        super().__init__()
        self.__dict__["nse"] = nse  # not a submodule (avoids a reference cycle in .to()/.cuda())
        pm, pc, lc = (torch.as_tensor(v, dtype=torch.float64).cpu() for v in (prior_mean, prior_cov, lik_cov))
        inv_lik, inv_prior = torch.linalg.inv(lc), torch.linalg.inv(pc)
        self.register_buffer("post_cov", torch.linalg.inv(inv_lik + inv_prior))
        self.register_buffer("inv_lik", inv_lik)
        self.register_buffer("inv_prior_mu", inv_prior @ pm)

    def tables(self, t64: torch.Tensor) -> torch.Tensor:
        """[len(t), m*m + m*d + m] fp64 rows [M_t | G_t | g_t] with eps = M_t theta - G_t x - g_t."""
        from . import engine
        a = engine._scalar64(self.nse.alpha, t64)
        sig = engine._scalar64(self.nse.sigma, t64)
        pc, il, ipm = self.post_cov.double().cpu(), self.inv_lik.double().cpu(), self.inv_prior_mu.double().cpu()
        m = pc.shape[0]
        eye = torch.eye(m, dtype=torch.float64)
        M = sig[:, None, None] * torch.linalg.inv(a[:, None, None] * pc + (1 - a)[:, None, None] * eye)
        G = a.sqrt()[:, None, None] * (M @ pc @ il)
        g = a.sqrt()[:, None] * (M @ pc @ ipm)
        return torch.cat((M.reshape(len(t64), -1), G.reshape(len(t64), -1), g), dim=1)

    def forward(self, theta, x, t):
        """torch evaluation (any device/dtype), same broadcasting as the reference closure: theta[B,m], x[B,d], t[B,1]."""
        nse = self.nse
        dt = theta.dtype
        pc, il, ipm = self.post_cov.to(theta), self.inv_lik.to(theta), self.inv_prior_mu.to(theta)
        a = nse.alpha(t)
        cov = pc[None] * a[..., None] + (1 - a)[..., None] * torch.eye(pc.shape[0], dtype=dt, device=theta.device)[None]
        mean0 = (pc @ (ipm[:, None] + il @ x.mT)).mT
        score = -(torch.linalg.inv(cov) @ (theta - (a ** 0.5) * mean0)[..., None])[..., 0]
        return -nse.sigma(t) * score


class FakeFNet(nn.Module):
    """real_eps_fun(theta, x, t) + eps_net_max * eps_net(theta, x, t)   (embedding_nets.py:141-160)."""

    def __init__(self, real_eps_fun, eps_net, eps_net_max):
        super().__init__()
        self.real_eps_fun = real_eps_fun
        self.eps_net = eps_net
        self.eps_net_max = eps_net_max

    def forward(self, theta, x, t):
        if t.ndim == 0:
            t = t[None, None].repeat(theta.shape[0], 1)
        return self.real_eps_fun(theta, x, t) + self.eps_net_max * self.eps_net(theta, x, t)
