"""Generic path of the tall-data sampler ("arbitrary nse.net"; section 8 row a13).

Score networks the fused kernels cannot take -- a Python closure inside `FakeFNet` (gen_gaussian_gaussian.py:59-72), modules
other than Linear/ReLU stacks, theta_dim in (10, 32] (gen_gaussian_gaussian.py:24 runs DIM 16 and 32) -- are evaluated by
torch (eager, on the GPU) on the (sample x observation) grid, exactly the way the reference forms them
(tall_posterior_sampler.py:60-121, nse_toy.py:53-92).  Everything after the scores stays native:

    kernel 3a  d4s_reduce_pairs    reduction over observations of the precision-weighted scores
    kernel 3b  d4s_finalize_wide   prior term, Lambda solve (LU, fp64), and
    kernel 4                        the DDIM / Langevin update with the same Philox keying as the fused kernels

so a generic-path trajectory uses the same schedule tables, the same noise and the same update arithmetic as a fused one.
CUDA tensors only; there is no CPU fallback here either.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch
from torch.func import jacrev, vmap

from . import _cabi as abi
from . import engine


def fusable(nse) -> bool:
    """True when libd4s can pack `nse.net` (Linear/ReLU stack, optional tanh head / analytic affine part, theta_dim <= 10)."""
    try:
        d = engine.net_description(nse)
        lin = engine._linears(d["mlp"], allow_tanh_head=d["out_act"] == 1)
        if d["emb_theta"] is not None:
            engine._linears(d["emb_theta"])
        return 2 <= len(lin) <= abi.MAX_LAYERS and lin[-1].out_features <= abi.MAX_THETA and \
            all(l.out_features <= abi.MAX_WIDTH for l in lin[:-1])
    except (abi.D4sError, AttributeError, TypeError):
        return False


# ---- torch evaluation of the network on the grid ----------------------------------------------------------------------
def scores_on_grid(nse, theta: torch.Tensor, x: torch.Tensor, t: torch.Tensor, toy: bool, **kw) -> torch.Tensor:
    """s[N, n, m] = score(theta_i, x_j, t).  `toy` selects the flattened-batch call of nse_toy.py:57-64 (FakeFNet takes
    [B, m], [B, d], [B, 1]); otherwise the broadcasting call of tall_posterior_sampler.py:86-89."""
    N, n = theta.shape[0], x.shape[0]
    with torch.no_grad():
        if toy:
            th = theta[:, None, :].expand(N, n, -1).reshape(N * n, -1)
            xx = x[None].expand(N, n, -1).reshape(N * n, -1)
            return nse.score(theta=th, x=xx, t=t[None, None].expand(N * n, 1)).reshape(N, n, -1)
        return nse.score(theta[:, None, :], x[None], t, **kw)


def jacobians_on_grid(nse, theta: torch.Tensor, x: torch.Tensor, t: torch.Tensor, toy: bool):
    """(J[N, n, m, m] = d s_ij / d theta_i, s[N, n, m]) by `vmap(vmap(jacrev))` like tall_posterior_sampler.py:60-85 /
    nse_toy.py:66-74."""
    if toy:
        def single(th, xx):
            s = nse.score(theta=th[None, :], x=xx[None, :], t=t[None, None])[0]
            return s, s
    else:
        def single(th, xx):
            s = nse.score(th[None, :], xx[None, :], t[None])[0]
            return s, s
    per_x = vmap(jacrev(single, has_aux=True), in_dims=(None, 0))
    jac, s = vmap(per_x, in_dims=(0, None))(theta, x)
    return jac.detach(), s.detach()


# ---- native kernels 3 / 4 ----------------------------------------------------------------------------------------------
def reduce_pairs(scores: torch.Tensor, pair_prec: Optional[torch.Tensor], obs_prec: Optional[torch.Tensor],
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = abi.require_cuda()
    N, n, m = scores.shape
    scores = scores.to(torch.float32).contiguous()
    if pair_prec is not None:
        pair_prec = pair_prec.to(torch.float32).contiguous()
    W = m + m * m if pair_prec is not None else 2 * m
    if out is None:
        out = torch.empty(N, W, dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        abi.check(lib.d4s_reduce_pairs(engine._ptr(scores), engine._ptr(pair_prec), engine._ptr(obs_prec), None, N, n, m,
                                       engine._ptr(out), engine._stream(scores.device)))
    return out


def finalize_wide(part: torch.Tensor, theta: torch.Tensor, layout_jac: bool, sched_row: torch.Tensor,
                  mats_row: Optional[torch.Tensor], prior_kind: int, low=None, high=None, ext_s0=None, ext_p0=None,
                  noise=None, seed: int = 0, step_index: int = 0, sample_offset: int = 0, clip=(1.0, -1.0),
                  update: bool = True, want_score: bool = False) -> Optional[torch.Tensor]:
    lib = abi.require_cuda()
    N, m = theta.shape
    score = torch.empty_like(theta) if want_score else None
    with torch.cuda.device(theta.device):
        abi.check(lib.d4s_finalize_wide(engine._ptr(part), N, m, 1 if layout_jac else 0, engine._ptr(sched_row),
                                        engine._ptr(mats_row), prior_kind, engine._ptr(low), engine._ptr(high),
                                        engine._ptr(ext_s0), engine._ptr(ext_p0), engine._ptr(noise),
                                        C.c_uint64(seed & (2 ** 64 - 1)), step_index, sample_offset, clip[0], clip[1],
                                        1 if update else 0, engine._ptr(theta), engine._ptr(score),
                                        engine._stream(theta.device)))
    return score


# ---- host driver --------------------------------------------------------------------------------------------------------
class GenericTall:
    """Tables + per-step driver of one generic-path trajectory / call.  `prior` is an `engine.PriorSpec` (closed-form
    priors: same step matrices as the fused path) or a callable `prior_closure(theta, t) -> (score, mean, precision)`
    (the toy scripts' triple, nse_toy.py:618), which is evaluated by torch every step and handed to kernel 3b."""

    def __init__(self, nse, x: torch.Tensor, n_samples: int, t: torch.Tensor, t_prev: Optional[torch.Tensor], eta: float,
                 prior, cov_mode: str, dist_cov_est: Optional[torch.Tensor], toy: bool, paired: bool = False,
                 weighted: bool = True, clip=(None, None), seed: int = 0, sample_offset: int = 0,
                 coef: Optional[dict] = None, always_weighted: bool = False, score_kwargs: Optional[dict] = None):
        if not x.is_cuda:
            raise abi.D4sError("d4s generic path: CUDA tensors only (no CPU fallback)")
        if cov_mode not in ("GAUSS", "JAC"):
            raise NotImplementedError("Available methods are GAUSS, JAC")
        self.nse, self.toy, self.paired = nse, toy, paired
        self.x = x if x.ndim > 1 else x[None]
        self.N, self.m = n_samples, int(nse.zeros.shape[0])
        if self.m > abi.MAX_THETA_WIDE:
            raise abi.D4sError(f"theta_dim {self.m} > {abi.MAX_THETA_WIDE} is not supported")
        dev = x.device
        self.dev = dev
        n = self.x.shape[0]
        self.n = 1 if paired else n
        self.jac = cov_mode == "JAC"
        self.closure = prior if callable(prior) and not isinstance(prior, engine.PriorSpec) else None
        spec = engine.PriorSpec(abi.PRIOR_NONE) if self.closure is not None else prior
        self.spec = spec
        if paired or not weighted:
            weighted = False
        t = t.detach().to("cpu").reshape(-1)
        self.t = t
        sa = engine._scalar64(nse.alpha, t.to(torch.float64)).sqrt()
        if self.closure is not None:  # toy scripts: always the weighted solve (nse_toy.py:619-625)
            comb = [abi.COMBINE_PER_SAMPLE if weighted else abi.COMBINE_SUM] * t.numel()
        else:
            c_hi, c_lo = (engine.combine_mode(spec, self.jac, self.n, v, weighted) for v in (1.0, 0.0))
            comb = torch.where(sa > 0.5, c_hi, c_lo).tolist()
            if always_weighted and weighted:
                comb = [abi.COMBINE_PER_SAMPLE if (self.jac or spec.kind != abi.PRIOR_GAUSSIAN) else abi.COMBINE_SHARED] * t.numel()
        self.comb = comb
        q_dev = sum_q = None
        if not self.jac and weighted:
            if dist_cov_est is None:
                raise ValueError("cov_mode='GAUSS' needs dist_cov_est (tall_posterior_sampler.py:125)")
            q64 = engine._inv_cov64(dist_cov_est)
            sum_q = q64.sum(dim=0)
            q_dev = engine._dev32_cached(q64, dev)
        self.q_dev = q_dev
        rows = engine.schedule_rows(nse, t, None if t_prev is None else t_prev.detach().to("cpu").reshape(-1), eta, self.n,
                                    comb, coef)
        mats = engine.mats_table(rows, spec, self.jac, self.n, sum_q, self.m)
        self.sched, self.mats = engine._dev32(rows, dev), engine._dev32(mats, dev)
        self.sig2_over_a = (rows[:, abi.S_SIGMA] ** 2 / rows[:, abi.S_ALPHA]).tolist()
        self.sig2 = (rows[:, abi.S_SIGMA] ** 2).tolist()
        self.t_dev = t.to(torch.float32).to(dev)
        self.low = self.high = None
        if spec.kind == abi.PRIOR_UNIFORM:
            self.low, self.high = engine._dev32(spec.low, dev), engine._dev32(spec.high, dev)
        lo, hi = clip
        self.clip = (1.0, -1.0) if lo is None else (float(lo), float(hi))
        self.seed, self.sample_offset = seed, sample_offset
        self.score_kwargs = score_kwargs or {}
        self.eye = torch.eye(self.m, device=dev)

    def step(self, k: int, theta: torch.Tensor, noise_k: Optional[torch.Tensor] = None, update: bool = True,
             want_score: bool = False) -> Optional[torch.Tensor]:
        nse, t = self.nse, self.t_dev[k]
        comb = self.comb[k]
        layout_jac = False
        if self.paired:  # plain score of the paired observation (nse.py:253, nse_toy.py:375)
            with torch.no_grad():
                if self.toy:
                    s = nse.score(theta=theta, x=self.x, t=t[None, None].expand(theta.shape[0], 1))
                else:
                    s = nse.score(theta, self.x, t, **self.score_kwargs)
            part = torch.cat((s.to(torch.float32), torch.zeros_like(s, dtype=torch.float32)), dim=1).contiguous()
        elif self.jac and comb != abi.COMBINE_SUM:
            jac, s = jacobians_on_grid(nse, theta, self.x, t, self.toy)
            cov = self.sig2_over_a[k] * (self.eye + self.sig2[k] * jac)  # tall_posterior_sampler.py:91-93
            part = reduce_pairs(s, torch.linalg.inv(cov), None)
            layout_jac = True
        else:
            s = scores_on_grid(nse, theta, self.x, t, self.toy, **self.score_kwargs)
            part = reduce_pairs(s, None, self.q_dev if comb != abi.COMBINE_SUM else None)
        kind, ext_s0, ext_p0 = self.spec.kind, None, None
        if self.closure is not None and not self.paired:
            triple = self.closure(theta, t)
            if isinstance(triple, (tuple, list)):
                ext_s0 = triple[0].detach().to(torch.float32).contiguous()
                ext_p0 = triple[2].detach().to(torch.float32).contiguous() if len(triple) > 2 else None
            else:
                ext_s0 = triple.detach().to(torch.float32).contiguous()
            kind = abi.PRIOR_EXTERNAL
        return finalize_wide(part, theta, layout_jac, self.sched[k], self.mats[k], kind, self.low, self.high, ext_s0,
                             ext_p0, noise_k, self.seed, k, self.sample_offset, self.clip, update, want_score)

    def run(self, theta: torch.Tensor, noise: Optional[torch.Tensor]) -> torch.Tensor:
        if noise is not None:
            noise = noise.to(theta.device, torch.float32).contiguous()
        for k in range(self.t.numel()):
            self.step(k, theta, None if noise is None else noise[k])
        return theta
