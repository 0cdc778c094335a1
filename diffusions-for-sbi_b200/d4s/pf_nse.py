"""`PF_NSE`: partially factorised score estimator (pf_nse.py:15-233 of the reference).

The network sees a *set* of up to `n_max` observations: masked mean of the embedded set elements plus a one-hot
encoding of the set size, concatenated as [theta, x_agg, temb, onehot(n)] (pf_nse.py:137-164).  At sampling time
the n observations are split into ceil(n/n_max) zero-padded subsets (pf_nse.py:169-197) and the tall-data
machinery of `NSE` runs over the K subsets.  For the fused kernels a subset is just one more "observation" whose
layer-1 partial is W1[:, x] x_agg + W1[:, onehot] e_n -- computed once per trajectory by engine.obs_feat_table.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Size, Tensor

from .nn import MLP
from .nse import NSE


class PF_NSE(NSE):
    def __init__(self, theta_dim: int, x_dim: int, n_max: int, freqs: int = 3,
                 build_net: Callable[..., nn.Module] = MLP, embedding_nn_theta: nn.Module = nn.Identity(),
                 embedding_nn_x: nn.Module = nn.Identity(), **kwargs) -> None:
        super().__init__(theta_dim=theta_dim, x_dim=x_dim, freqs=freqs, build_net=build_net, net_type="default",
                         embedding_nn_theta=embedding_nn_theta, embedding_nn_x=embedding_nn_x, **kwargs)
        self.n_max = n_max
        self.net = build_net(self.theta_emb_dim + self.x_emb_dim + 2 * freqs + n_max, theta_dim, **kwargs)

    # ---- set handling (pf_nse.py:78-115) ----------------------------------------------------------------
    def _create_mask(self, n: Tensor) -> Tensor:
        idx = torch.arange(self.n_max, device=n.device).to(n)
        return (idx < n).to(n).unsqueeze(-1)  # (*, n_max, 1)

    def set_mask(self, mask: Tensor) -> None:
        emb = self.embedding_nn_x
        if hasattr(emb, "set_mask"):
            emb.set_mask(mask)
        elif hasattr(emb, "__iter__"):
            for layer in emb:
                if hasattr(layer, "set_mask"):
                    layer.set_mask(mask)

    def aggregate(self, x: Tensor, n: Tensor) -> Tensor:
        return x.sum(dim=-2) / n if n.sum() != 0 else x.mean(dim=-2)

    def forward(self, theta: Tensor, x: Tensor, t: Tensor, n: Tensor) -> Tensor:
        """eps(theta, {x^j}, t, n): theta[*, m], x[*, n_max, d], n[*, 1] (pf_nse.py:117-164)."""
        mask = self._create_mask(n)
        self.set_mask(mask)
        ang = self.freqs * t[..., None]
        temb = torch.cat((ang.cos(), ang.sin()), dim=-1)
        th = self.embedding_nn_theta(theta)
        xe = mask * self.embedding_nn_x(mask * x)
        x_agg = self.aggregate(xe, n)
        if not n.eq(0).all():
            hot = F.one_hot(n.long().squeeze(-1) - 1, num_classes=self.n_max).to(th.dtype)
        else:
            hot = torch.zeros((n.shape[0], self.n_max), dtype=th.dtype, device=th.device)
        batch = torch.broadcast_shapes(th.shape[:-1], x_agg.shape[:-1], temb.shape[:-1], hot.shape[:-1])
        feats = [v.expand(batch + v.shape[-1:]) for v in (th, x_agg, temb, hot)]
        return self.net(torch.cat(feats, dim=-1))

    def score(self, theta: Tensor, x: Tensor, t: Tensor, n: Tensor, **kwargs) -> Tensor:
        return -self(theta, x, t, n) / self.sigma(t)

    def _create_pf_data(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        """x[n, d] -> xs[K, n_max, d] (zero padded), ns[K, 1] with K = ceil(n / n_max) (pf_nse.py:169-197)."""
        if x.ndim == 1:
            x = x[None]
        n_obs, d = x.shape
        K = -(-n_obs // self.n_max)
        xs = x.new_zeros(K * self.n_max, d)
        xs[:n_obs] = x
        ns = x.new_full((K, 1), float(self.n_max))
        if n_obs % self.n_max:
            ns[-1, 0] = float(n_obs % self.n_max)
        return xs.reshape(K, self.n_max, d), ns

    # ---- samplers: reshape the context, then defer to NSE (pf_nse.py:199-233) ---------------------------------
    def ddim(self, shape: Size, x: Tensor, steps: int = 64, eta: float = 1.0, verbose: bool = False,
             theta_clipping_range=(None, None), **kwargs) -> Tensor:
        xs, ns = self._create_pf_data(x)
        if "n" in kwargs:
            if tuple(kwargs["n"].shape) != tuple(ns.shape):
                raise AssertionError(f"n.shape = {tuple(ns.shape)} != kwargs['n'].shape = {tuple(kwargs['n'].shape)}")
            ns = kwargs.pop("n")
        return super().ddim(shape=shape, x=xs, steps=steps, eta=eta, verbose=verbose,
                            theta_clipping_range=theta_clipping_range, n=ns, **kwargs)

    def ddim_contexts(self, shape: Size, x: Tensor, steps: int = 64, eta: float = 1.0, verbose: bool = False,
                      theta_clipping_range=(None, None), **kwargs) -> Tensor:
        """`vmap(lambda x_c, cov_c: self.ddim(shape, x_c, ..., dist_cov_est=cov_c))(x, dist_cov_est)` as one trajectory launch
        (see `NSE.ddim_contexts`): x[B, n, d] -> every context is split into its K = ceil(n / n_max) subsets like `ddim` does
        (pf_nse.py:169-197); dist_cov_est is [B, K, m, m] (one covariance per subset, as for `ddim`)."""
        if x.ndim != 3:
            raise ValueError("PF_NSE.ddim_contexts: x must be [B, n, d]")
        parts = [self._create_pf_data(xc) for xc in x]
        xs = torch.stack([p[0] for p in parts])  # [B, K, n_max, d]
        ns = torch.stack([p[1] for p in parts])  # [B, K, 1]
        return super().ddim_contexts(shape=shape, x=xs, steps=steps, eta=eta, verbose=verbose,
                                     theta_clipping_range=theta_clipping_range, n=ns, **kwargs)

    def predictor_corrector(self, shape: Size, x: Tensor, steps: int = 64, verbose: bool = False,
                            predictor_type="ddim", corrector_type="langevin", theta_clipping_range=(None, None),
                            **kwargs) -> Tensor:
        xs, ns = self._create_pf_data(x)  # pf_nse.py:235-262
        return super().predictor_corrector(shape=shape, x=xs, steps=steps, verbose=verbose, predictor_type=predictor_type,
                                           corrector_type=corrector_type, theta_clipping_range=theta_clipping_range,
                                           n=ns, **kwargs)

    def annealed_langevin_geffner(self, shape: Size, x: Tensor, prior_score_fn, prior_type=None, steps: int = 400,
                                  lsteps: int = 5, tau: float = 0.5, theta_clipping_range=(None, None),
                                  verbose: bool = False, **kwargs) -> Tensor:
        xs, ns = self._create_pf_data(x)  # pf_nse.py:264-297 (PF-NPSE of Geffner et al.)
        return super().annealed_langevin_geffner(shape=shape, x=xs, prior_score_fn=prior_score_fn, prior_type=prior_type,
                                                 steps=steps, lsteps=lsteps, tau=tau,
                                                 theta_clipping_range=theta_clipping_range, verbose=verbose, n=ns, **kwargs)
