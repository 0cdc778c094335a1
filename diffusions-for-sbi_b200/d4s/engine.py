"""Host side of the tall-data sampler: turns an `NSE`-like module + the reference's keyword arguments
into the device tables libd4s consumes, and drives the C ABI.

What is computed here (once per trajectory / per call, in fp64 on the host, then cast to fp32):
  * the DDIM schedule table (nse.py:258-259, 283-298) from the module's own `alpha`/`sigma`
    (callers monkey-patch them, gen_gaussian_gaussian.py:46-50);
  * the separable layer-1 partials: per observation `obs_feat` and per step `time_feat`
    (SURVEY.md section 7 identity 1; PF_NSE layout pf_nse.py:163-164);
  * the per-step small matrices of the precision-weighted combination (nse.py:628-654):
    Lambda^-1 (GAUSS + Gaussian prior), the (1-n) P0 C^-1 prior map, sqrt(alpha) mu0.
Everything that touches a posterior sample runs in the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _cabi as abi


# ---------------------------------------------------------------------------------------------------------
# prior description
# ---------------------------------------------------------------------------------------------------------
@dataclass
class PriorSpec:
    kind: int  # abi.PRIOR_*
    low: Optional[torch.Tensor] = None  # uniform: [m] fp64 cpu
    high: Optional[torch.Tensor] = None
    mean: Optional[torch.Tensor] = None  # gaussian: [m], [m, m] fp64 cpu
    cov: Optional[torch.Tensor] = None   # covariance of the diffused prior SCORE (prior_score_fn's own, vp_diffused_priors.py:65)
    cov_prec: Optional[torch.Tensor] = None  # covariance behind the backward precision P0 (`prior.covariance_matrix`, nse.py:630)


def _cpu64(v) -> torch.Tensor:
    return torch.as_tensor(v).detach().to("cpu", torch.float64)


def prior_spec(prior_type: Optional[str], prior, prior_score_fn, m: int) -> PriorSpec:
    """Reads the prior parameters from the score-function object (d4s.vp_diffused_priors factories expose
    `.kind/.low/.high/.mean/.cov`) or, failing that, from the torch distribution the reference passes as
    `prior=` (sbibm_posterior_estimation.py:190,205)."""
    if prior_type is None:
        return PriorSpec(abi.PRIOR_NONE)
    fn_kind = getattr(prior_score_fn, "kind", None)
    if prior_type == "gaussian":
        if fn_kind == "gaussian":
            mean, cov = _cpu64(prior_score_fn.mean), _cpu64(prior_score_fn.cov)
        elif prior is not None and hasattr(prior, "covariance_matrix"):
            mean, cov = _cpu64(prior.loc), _cpu64(prior.covariance_matrix)
        else:
            raise ValueError("gaussian prior: pass a d4s.vp_diffused_priors score fn or a MultivariateNormal `prior`")
        cov_prec = cov
        if prior is not None and hasattr(prior, "covariance_matrix"):
            cov_prec = _cpu64(prior.covariance_matrix)  # nse.py:630 reads the precision from `prior`, the score from the fn
        return PriorSpec(abi.PRIOR_GAUSSIAN, mean=mean.reshape(m), cov=cov.reshape(m, m), cov_prec=cov_prec.reshape(m, m))
    if prior_type == "uniform":
        if fn_kind == "uniform":
            low, high = _cpu64(prior_score_fn.low), _cpu64(prior_score_fn.high)
        elif prior is not None and hasattr(prior, "low"):
            low, high = _cpu64(prior.low), _cpu64(prior.high)
        elif prior is not None and hasattr(prior, "base_dist"):
            low, high = _cpu64(prior.base_dist.low), _cpu64(prior.base_dist.high)
        else:
            raise ValueError("uniform prior: pass a d4s.vp_diffused_priors score fn or a Uniform `prior`")
        return PriorSpec(abi.PRIOR_UNIFORM, low=low.reshape(m).contiguous(), high=high.reshape(m).contiguous())
    raise NotImplementedError(f"prior_type={prior_type!r}: only 'gaussian' and 'uniform' have a fused diffused score")


# ---------------------------------------------------------------------------------------------------------
# packed net cache
# ---------------------------------------------------------------------------------------------------------
def pad_width(h: int) -> int:
    """Hidden widths are zero padded to 64 / 128 / 256 (same rule as pad_width() in csrc/d4s_api.cu)."""
    for w in (64, 128, 256):
        if h <= w:
            return w
    raise abi.D4sError(f"hidden width {h} > {abi.MAX_WIDTH} is not supported")


class NetView:
    """Host-side (fp64) view of the layer-0 blocks of `nse.net`; needs no CUDA device."""

    def __init__(self, linears):
        L = len(linears)
        if not 2 <= L <= abi.MAX_LAYERS:
            raise abi.D4sError(f"score MLP must have 2..{abi.MAX_LAYERS} Linear layers, got {L}")
        self.linears = linears
        self.m = linears[-1].out_features
        self.h1 = linears[0].out_features
        self.h1p = pad_width(self.h1)
        self.w1 = linears[0].weight.detach().to("cpu", torch.float64)  # [H1, in]
        self.b1 = linears[0].bias.detach().to("cpu", torch.float64)


class PackedNet(NetView):
    """Device-resident packed copy of `nse.net` (d4s_net_create) + the host view."""

    def __init__(self, linears, theta_cols: int, device: torch.device, emb_linears=(), out_act: int = 0,
                 out_scale: float = 1.0):
        super().__init__(linears)
        lib = abi.require_cuda()
        L = len(linears)
        desc = abi.NetDesc()
        desc.n_layers = L
        desc.theta_cols = theta_cols
        keep = []
        desc.dims[0] = linears[0].in_features
        for l, lin in enumerate(linears):
            w = lin.weight.detach().to("cpu", torch.float32).contiguous()
            b = lin.bias.detach().to("cpu", torch.float32).contiguous()
            keep += [w, b]
            desc.dims[l + 1] = lin.out_features
            desc.weight[l] = w.data_ptr()
            desc.bias[l] = b.data_ptr()
        desc.out_act = int(out_act)
        desc.out_scale = float(out_scale)
        desc.emb_layers = len(emb_linears)
        for l, lin in enumerate(emb_linears):
            w = lin.weight.detach().to("cpu", torch.float32).contiguous()
            b = lin.bias.detach().to("cpu", torch.float32).contiguous()
            keep += [w, b]
            desc.emb_dims[l], desc.emb_dims[l + 1] = lin.in_features, lin.out_features
            desc.emb_weight[l] = w.data_ptr()
            desc.emb_bias[l] = b.data_ptr()
        handle = C.c_void_p()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            abi.check(lib.d4s_net_create(C.byref(handle), C.byref(desc), C.c_void_p(stream)))
        del keep
        self.handle = handle
        self.device = device
        assert self.h1p == lib.d4s_net_h1_padded(handle)
        self._finalizer = weakref.finalize(self, lib.d4s_net_free, handle)


_net_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _linears(mlp, allow_tanh_head: bool = False) -> list:
    mods = list(mlp.modules())[1:] if isinstance(mlp, torch.nn.Sequential) else list(mlp.children())
    lin = []
    for i, mod in enumerate(mods):
        if isinstance(mod, torch.nn.Linear):
            lin.append(mod)
        elif isinstance(mod, (torch.nn.ReLU, torch.nn.Identity)):
            continue
        elif allow_tanh_head and isinstance(mod, torch.nn.Tanh) and i == len(mods) - 1:
            continue
        else:
            raise abi.D4sError(f"unsupported module in score MLP: {type(mod).__name__} (Linear/ReLU stacks only)")
    return lin


def net_description(nse) -> dict:
    """What the kernels need to know about `nse.net`.  Modules may override it with `_d4s_net_description()`
    (the toy `nse_toy.NSE` + `FakeFNet` does: tanh head, analytic affine part, raw-t time feature)."""
    if hasattr(nse, "_d4s_net_description"):
        return nse._d4s_net_description()
    emb = nse.embedding_nn_theta
    return dict(mlp=nse.net, emb_theta=None if isinstance(emb, torch.nn.Identity) else emb, out_act=0, out_scale=1.0,
                theta_cols=int(getattr(nse, "theta_emb_dim", nse.zeros.shape[0])), time_features=None, affine=None)


def packed_net(nse, device: torch.device) -> PackedNet:
    d = net_description(nse)
    lin = _linears(d["mlp"], allow_tanh_head=d["out_act"] == 1)
    emb_lin = [] if d["emb_theta"] is None else _linears(d["emb_theta"])
    key = tuple((p.data_ptr(), p._version) for l in lin + emb_lin for p in (l.weight, l.bias)) + (
        str(device), d["out_act"], float(d["out_scale"]))
    hit = _net_cache.get(nse)
    if hit is not None and hit[0] == key:
        return hit[1]
    pn = PackedNet(lin, d["theta_cols"], device, emb_lin, d["out_act"], d["out_scale"])
    _net_cache[nse] = (key, pn)
    return pn


# ---------------------------------------------------------------------------------------------------------
# schedule + tables (fp64 host)
# ---------------------------------------------------------------------------------------------------------
def time_grid(steps: int) -> torch.Tensor:
    """`torch.linspace(1, 0, steps + 1)` in fp32 (nse.py:258); its values are exactly what the reference feeds
    to alpha/sigma/the time embedding, in fp32 and in fp64 runs alike."""
    return torch.linspace(1, 0, steps + 1)


def _scalar64(fn, t64: torch.Tensor) -> torch.Tensor:
    out = fn(t64)
    return torch.as_tensor(out, dtype=torch.float64).reshape(t64.shape)


def schedule_rows(nse, t: torch.Tensor, t_prev: Optional[torch.Tensor], eta: float, n_total: int,
                  combine: Sequence[int], coef: Optional[dict] = None) -> torch.Tensor:
    """[len(t), 16] fp64 schedule rows.  t_prev = t - dt gives the DDIM update coefficients; `coef` overrides update
    columns by name (c_theta, c_score, bstd, s1w, tame, clip) for the Langevin-family samplers."""
    t64 = t.to(torch.float64)
    a = _scalar64(nse.alpha, t64)
    sig = _scalar64(nse.sigma, t64)
    rows = torch.zeros(t64.numel(), abi.SCHED_STRIDE, dtype=torch.float64)
    rows[:, abi.S_T] = t64
    rows[:, abi.S_ALPHA] = a
    rows[:, abi.S_SIGMA] = sig
    rows[:, abi.S_INV_SIGMA] = 1.0 / sig
    rows[:, abi.S_A_OVER_S2] = a / sig ** 2
    rows[:, abi.S_SQRT_ALPHA] = a.sqrt()
    rows[:, abi.S_COMBINE] = torch.as_tensor(list(combine), dtype=torch.float64)
    rows[:, abi.S_ONE_MINUS_N] = float(1 - n_total)
    rows[:, abi.S_S1_WEIGHT] = 1.0
    rows[:, abi.S_CLIP] = 1.0
    if t_prev is not None:
        a1 = _scalar64(nse.alpha, t_prev.to(torch.float64))
        bs = eta * (((1 - a1) / (1 - a)) * (1 - a / a1)).clamp_min(0).sqrt()  # nse.py:285-287
        c_theta = (a1 / a).sqrt()  # sqrt(alpha_{t-1}) * alpha_t^-1/2      (nse.py:205, 219)
        # est_noise = (theta - sqrt(a) theta0) / sqrt(1-a) = -sqrt(1-a) * score   (nse.py:218, exact algebra)
        c_score = c_theta * (1 - a) - (1 - a1 - bs ** 2).clamp_min(0).sqrt() * (1 - a).sqrt()
        rows[:, abi.S_C_THETA] = c_theta
        rows[:, abi.S_C_SCORE] = c_score
        rows[:, abi.S_BRIDGE_STD] = bs
    if coef:
        cols = dict(c_theta=abi.S_C_THETA, c_score=abi.S_C_SCORE, bstd=abi.S_BRIDGE_STD, s1w=abi.S_S1_WEIGHT,
                    tame=abi.S_TAME, clip=abi.S_CLIP)
        for name, val in coef.items():
            rows[:, cols[name]] = torch.as_tensor(val, dtype=torch.float64).reshape(-1)
    return rows


def time_embedding64(nse, t64: torch.Tensor) -> torch.Tensor:
    fn = net_description(nse)["time_features"]
    if fn is not None:  # e.g. the toy net feeds the raw time (embedding_nets.py:138)
        return fn(t64)
    ang = nse.freqs.detach().to("cpu", torch.float64) * t64[..., None]  # nse.py:125
    return torch.cat((ang.cos(), ang.sin()), dim=-1)  # nse.py:126


def layer1_blocks(nse, pn: NetView):
    """Column blocks of the first Linear: [theta | x | temb | (one-hot n, PF_NSE only)]."""
    e_t = int(getattr(nse, "theta_emb_dim", pn.m))
    e_x = int(getattr(nse, "x_emb_dim"))
    f2 = int(net_description(nse).get("time_width", 2 * nse.freqs.numel()))
    cols = pn.w1.shape[1]
    n_hot = cols - e_t - e_x - f2
    if n_hot < 0:
        raise abi.D4sError("first Linear is narrower than theta+x+time features")
    return (slice(0, e_t), slice(e_t, e_t + e_x), slice(e_t + e_x, e_t + e_x + f2),
            slice(e_t + e_x + f2, cols), n_hot)


def time_feat_table(nse, pn: NetView, t: torch.Tensor) -> torch.Tensor:
    """[len(t), H1p] fp64: W1[:, t-block] temb(t) + b1, zero padded.  Cached on the packed-net view (which is rebuilt
    whenever a weight changes) per time grid: it does not depend on the observations, so repeated trajectories on the
    same grid reuse it."""
    tc = t.detach().to("cpu", torch.float64).contiguous()
    freqs = getattr(nse, "freqs", None)
    key = (tc.numel(), hash(tc.numpy().tobytes()),
           None if freqs is None else tuple(freqs.detach().to("cpu", torch.float64).reshape(-1).tolist()))
    cache = pn.__dict__.setdefault("_tf_cache", {})
    hit = cache.get(key)
    if hit is not None:
        return hit
    out = _time_feat_table(nse, pn, t)
    if len(cache) > 16:
        cache.clear()
    cache[key] = out
    return out


def _time_feat_table(nse, pn: NetView, t: torch.Tensor) -> torch.Tensor:
    _, _, tcol, _, _ = layer1_blocks(nse, pn)
    temb = time_embedding64(nse, t.to(torch.float64))
    out = torch.zeros(t.numel(), pn.h1p, dtype=torch.float64)
    out[:, :pn.h1] = temb @ pn.w1[:, tcol].T + pn.b1
    return out


def obs_feat_table(nse, pn: NetView, x: torch.Tensor, n_sizes: Optional[torch.Tensor]) -> torch.Tensor:
    """[n, H1p] fp64 layer-1 partial of every observation (or PF_NSE subset).

    NSE:    W1[:, x-block] e_x(x_j)                                                  (nse.py:130-136)
    PF_NSE: W1[:, x-block] masked-mean_j e_x(x^j) + W1[:, one-hot block] onehot(n)  (pf_nse.py:137-164)
    """
    _, xcol, _, hcol, n_hot = layer1_blocks(nse, pn)
    emb = nse.embedding_nn_x
    with torch.no_grad():
        if n_sizes is None:
            xe = emb(x).detach().to("cpu", torch.float64)  # [n, e_x]
            feat = xe @ pn.w1[:, xcol].T
        else:
            ns = n_sizes.detach().to("cpu", torch.float64).reshape(-1, 1)  # [K, 1]
            K, n_max = x.shape[0], x.shape[1]
            mask = (torch.arange(n_max, dtype=torch.float64)[None, :] < ns).to(x.dtype).unsqueeze(-1).to(x.device)
            x = x.reshape(K, n_max, -1)
            if hasattr(nse, "set_mask"):
                nse.set_mask(mask)
            xe = (mask * emb(mask * x)).detach().to("cpu", torch.float64)  # pf_nse.py:141,154
            if float(ns.sum()) != 0:
                x_agg = xe.sum(dim=-2) / ns  # pf_nse.py:109-111
            else:
                x_agg = xe.mean(dim=-2)
            feat = x_agg @ pn.w1[:, xcol].T
            if not bool((ns == 0).all()):
                onehot = torch.nn.functional.one_hot(ns.long().reshape(-1) - 1, num_classes=n_hot).to(torch.float64)
                feat = feat + onehot @ pn.w1[:, hcol].T  # pf_nse.py:158-164
    out = torch.zeros(feat.shape[0], pn.h1p, dtype=torch.float64)
    out[:, :pn.h1] = feat
    return out


def combine_mode(prior: PriorSpec, cov_jac: bool, n_total: int, sqrt_alpha: float, weighted: bool = True) -> int:
    """Which branch of nse.py:627-654 a step takes (decided on the host: it only depends on t)."""
    if not weighted or prior.kind == abi.PRIOR_NONE:
        return abi.COMBINE_SUM
    if prior.kind == abi.PRIOR_GAUSSIAN:
        return abi.COMBINE_PER_SAMPLE if cov_jac else abi.COMBINE_SHARED
    if sqrt_alpha > 0.5 and n_total > 1:  # nse.py:643
        return abi.COMBINE_PER_SAMPLE
    return abi.COMBINE_SUM


def mats_table(rows: torch.Tensor, prior: PriorSpec, cov_jac: bool, n_total: int, sum_q: Optional[torch.Tensor],
               m: int, w_sum: Optional[float] = None) -> torch.Tensor:
    """[steps, 2 m^2 + m] fp64: Lmat | G | mu_t  (layout of D4sStep.mats)."""
    S = rows.shape[0]
    eye = torch.eye(m, dtype=torch.float64)
    a = rows[:, abi.S_ALPHA]
    sig2 = rows[:, abi.S_SIGMA] ** 2
    aos2 = rows[:, abi.S_A_OVER_S2]
    comb = rows[:, abi.S_COMBINE].to(torch.int64)
    one_minus_n = float(1 - n_total)
    lmat = torch.zeros(S, m, m, dtype=torch.float64)
    g = torch.zeros(S, m, m, dtype=torch.float64)
    mu = torch.zeros(S, m, dtype=torch.float64)
    base = torch.zeros(S, m, m, dtype=torch.float64)
    if not cov_jac and sum_q is not None:
        # sum_j w_j [inv(Sigma_j) + alpha/sigma^2 I]; w_j = 1 except for the classifier-free-guidance pseudo-observation
        base = sum_q[None] + ((n_total if w_sum is None else w_sum) * aos2)[:, None, None] * eye
    if prior.kind == abi.PRIOR_GAUSSIAN:
        cov0, mean0 = prior.cov, prior.mean
        c_t = sig2[:, None, None] * eye + a[:, None, None] * cov0  # vp_diffused_priors.py:65-68
        c_inv = torch.linalg.inv(c_t)
        cov_p = prior.cov_prec if prior.cov_prec is not None else cov0
        p0 = torch.linalg.inv(cov_p)[None] + aos2[:, None, None] * eye  # prec_matrix_backward(prior cov), nse.py:630
        mu = a.sqrt()[:, None] * mean0
        weighted = comb != abi.COMBINE_SUM
        # score_prior = -(theta - mu_t) C^-1 (row-vector form, vp_diffused_priors.py:71) = -C^-T (theta - mu_t)
        g_w = -one_minus_n * (p0 @ c_inv.transpose(-1, -2))
        g_s = -one_minus_n * c_inv.transpose(-1, -2)
        g = torch.where(weighted[:, None, None], g_w, g_s)
        lam = one_minus_n * p0 + base
        if cov_jac:
            lmat = one_minus_n * p0
        else:
            shared = comb == abi.COMBINE_SHARED
            safe = torch.where(shared[:, None, None], lam, eye.expand(S, m, m))
            lmat = torch.where(shared[:, None, None], torch.linalg.inv(safe), lam)
    else:
        lmat = base  # uniform / none: per-sample prior precision is added on the device
        shared = comb == abi.COMBINE_SHARED  # no analytic prior (classifier-free guidance): Lambda is the base itself
        if bool(shared.any()):
            safe = torch.where(shared[:, None, None], base, eye.expand(S, m, m))
            lmat = torch.where(shared[:, None, None], torch.linalg.inv(safe), base)
    return torch.cat((lmat.reshape(S, -1), g.reshape(S, -1), mu), dim=1)


# ---------------------------------------------------------------------------------------------------------
# caches of the observation-dependent tables (per-call API: factorized_score in a user loop, euler_sde_sampler)
# ---------------------------------------------------------------------------------------------------------
class _LRU:
    """Tiny LRU keyed by tensor identity.  Every entry keeps a strong reference to the tensors its key was derived from,
    so their storage cannot be freed and handed to another tensor while the entry lives (a key is (data_ptr, _version,
    shape, ...): an in-place update bumps `_version`, and a live reference rules out address reuse)."""

    def __init__(self, cap: int):
        self.cap, self.d = cap, {}

    def get(self, key):
        hit = self.d.get(key)
        if hit is not None:
            self.d[key] = self.d.pop(key)  # most recent last
            return hit[0]
        return None

    def put(self, key, value, refs=()):
        self.d[key] = (value, refs)
        while len(self.d) > self.cap:
            self.d.pop(next(iter(self.d)))

    def clear(self):
        self.d.clear()


def tensor_key(t):
    """Identity of a tensor's current contents (None-safe)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        return ("py", repr(t))
    return (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), str(t.dtype), str(t.device))


_qinv_cache = _LRU(8)     # dist_cov_est -> inv (fp64, host)
_obsfeat_cache = _LRU(8)  # (packed net, x, n_sizes) -> layer-0 observation partials (fp64, host)
_dev_cache = _LRU(64)     # host fp64 table -> fp32 device copy (keyed by the host tensor's identity)
_setup_cache = _LRU(1200)  # whole per-call setups (NSE.factorized_score with the same inputs and the same t)
_tables_cache = _LRU(8)    # VALUES of everything the schedule / step matrices depend on -> their device copies


def clear_caches():
    for c in (_qinv_cache, _obsfeat_cache, _dev_cache, _setup_cache, _tables_cache):
        c.clear()


def _bytes(v) -> Optional[bytes]:
    return None if v is None else torch.as_tensor(v).detach().to("cpu", torch.float64).contiguous().numpy().tobytes()


def _sched_mats_dev(nse, t, t_prev, eta, n_total, comb, coef, prior: PriorSpec, cov_jac, sum_q, m, w_sum, device):
    """Device copies of the schedule rows and the per-step matrices of a trajectory.  A `ddim` call on fresh input tensors
    (new x / dist_cov_est every call) still repeats these tables whenever the grid, eta, n, the prior and sum_j inv(Sigma_j) are
    the same; building them is most of the host time of a call.  The key holds the VALUES they depend on -- alpha / sigma are
    evaluated on the grid every time (callers monkey-patch them), the prior's parameters and sum_q enter as bytes -- never a
    tensor identity, so a hit cannot be stale."""
    t64 = t.to(torch.float64)
    key = (_bytes(_scalar64(nse.alpha, t64)), _bytes(_scalar64(nse.sigma, t64)),
           None if t_prev is None else _bytes(_scalar64(nse.alpha, t_prev.to(torch.float64))), _bytes(t64), float(eta), int(n_total),
           tuple(int(c) for c in comb), None if not coef else tuple((k, _bytes(v)) for k, v in sorted(coef.items())),
           prior.kind, _bytes(prior.low), _bytes(prior.high), _bytes(prior.mean), _bytes(prior.cov), _bytes(prior.cov_prec),
           bool(cov_jac), _bytes(sum_q), int(m), None if w_sum is None else float(w_sum), str(device))
    hit = _tables_cache.get(key)
    if hit is None:
        rows = schedule_rows(nse, t, t_prev, eta, n_total, comb, coef)
        mats = mats_table(rows, prior, cov_jac, n_total, sum_q, m, w_sum)
        hit = (_dev32(rows, device), _dev32(mats, device))
        _tables_cache.put(key, hit, ())
    return hit


def _inv_cov64(cov: torch.Tensor) -> torch.Tensor:
    key = tensor_key(cov)
    hit = _qinv_cache.get(key)
    if hit is None:
        hit = torch.linalg.inv(cov.detach().to("cpu", torch.float64))
        _qinv_cache.put(key, hit, (cov,))
    return hit


def _obs_feat_cached(nse, pn, x, n_sizes):
    emb_key = tuple((q.data_ptr(), q._version) for q in nse.embedding_nn_x.parameters())  # not part of the packed net
    key = (id(pn), tensor_key(x), tensor_key(n_sizes), emb_key)
    hit = _obsfeat_cache.get(key)
    if hit is None:
        hit = obs_feat_table(nse, pn, x, n_sizes)
        _obsfeat_cache.put(key, hit, (pn, x, n_sizes))
    return hit


def _dev32_cached(t64: torch.Tensor, device) -> torch.Tensor:
    """fp32 device copy of a host table that is itself cached (its identity is stable across calls)."""
    key = (tensor_key(t64), str(device))
    hit = _dev_cache.get(key)
    if hit is None:
        hit = _dev32(t64, device)
        _dev_cache.put(key, hit, (t64,))
    return hit


# ---------------------------------------------------------------------------------------------------------
# launching
# ---------------------------------------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _dev32(t64: torch.Tensor, device) -> torch.Tensor:
    return t64.to(torch.float32).contiguous().to(device, non_blocking=False)


@dataclass
class TallSetup:
    """Device tables of one trajectory / one call (everything d4s_ddim_run needs)."""
    pn: PackedNet
    prob: abi.Problem
    sched: torch.Tensor
    time_feat: torch.Tensor
    mats: torch.Tensor
    step_modes: Optional[np.ndarray]
    keep: tuple  # tensors referenced by raw pointers in `prob`
    n_steps: int
    affine: Optional[torch.Tensor] = None  # [steps, m*m + m*d + m] analytic part of eps (toy nets)


def build_setup(nse, x: torch.Tensor, n_samples: int, t: torch.Tensor, t_prev: Optional[torch.Tensor], eta: float,
                prior: PriorSpec, cov_mode: str, dist_cov_est: Optional[torch.Tensor], paired: bool = False,
                n_sizes: Optional[torch.Tensor] = None, clip=(None, None), seed: int = 0, sample_offset: int = 0,
                weighted: bool = True, obs_slice: Optional[slice] = None, obs_chunk: int = 0,
                coef: Optional[dict] = None, cfg: Optional[dict] = None, contexts: int = 0) -> TallSetup:
    """`obs_slice` restricts the kernels to a shard of the observations (observation sharding over ranks):
    (1-n) and Lambda still use the total n, the partial sums are all-reduced by the caller.

    `paired = k >= 1` (True == 1): every sample has its own observation, row i // k of `x`, and the plain score is used
    (nse.py:253); k = S with x[B, d] is the covariance pre-pass (S draws from each of B single-observation posteriors).

    `contexts = B > 0`: B independent tall problems in one launch (replaces `vmap(sampler)` over calibration sets,
    jrnnm_posterior_estimation.py:286-294).  `x` is [B * n, d] and `dist_cov_est` [B * n, m, m], context-major;
    samples [c S, (c+1) S) with S = n_samples / B belong to context c."""
    lib = abi.require_cuda()
    if not x.is_cuda:
        raise abi.D4sError("d4s: observations must live on a CUDA device (no CPU fallback)")
    device = x.device
    pn = packed_net(nse, device)
    m = pn.m
    if cov_mode not in ("GAUSS", "JAC"):
        raise NotImplementedError("Available methods are GAUSS, JAC")
    cov_jac = cov_mode == "JAC"
    x2 = x if x.ndim > 1 else x[None]
    n_total = x2.shape[0]
    if contexts:
        if paired or cfg is not None or obs_slice is not None:
            raise NotImplementedError("batched contexts with paired / classifier-free guidance / sharded inputs")
        if x2.shape[0] % contexts or n_samples % contexts:
            raise ValueError("batched contexts: x must be [B * n, d] and n_samples a multiple of B")
        n_total = x2.shape[0] // contexts  # observations of ONE context: the n of (1 - n) and of Lambda
    if paired:
        weighted = False
    sum_q = q_dev = ctx_qsum = None
    # classifier-free guidance (nse.py:598-625, 532-538): the prior score/precision come from the net itself at x = 0
    # (PF_NSE: empty set, n = 0).  That is the tall score with one more pseudo-observation of weight (1 - n).
    w64 = None
    if cfg is not None:
        if paired or obs_slice is not None:
            raise NotImplementedError("classifier-free guidance with paired contexts / observation sharding")
        w64 = torch.cat((torch.ones(n_total, dtype=torch.float64), torch.tensor([1.0 - n_total], dtype=torch.float64)))
        prior = PriorSpec(abi.PRIOR_NONE)
    if not cov_jac and weighted and not paired:
        if dist_cov_est is None:
            raise ValueError("cov_mode='GAUSS' needs dist_cov_est (tall_posterior_sampler.py:125)")
        q64 = _inv_cov64(dist_cov_est)  # [n, m, m] fp64 on the host, cached per dist_cov_est tensor
        q_cached = cfg is None
        if cfg is not None:
            if cfg.get("prior_cov") is None:
                raise ValueError("clf_free_guidance with cov_mode='GAUSS' needs dist_cov_est_prior (nse.py:613)")
            qp = torch.linalg.inv(cfg["prior_cov"].detach().to("cpu", torch.float64).reshape(1, m, m))
            q64 = torch.cat((q64, qp))
            sum_q = (w64[:, None, None] * q64).sum(dim=0)
        elif contexts:  # per-context sum_j inv(Sigma_cj) goes to the device; the step matrices carry no observation term
            ctx_qsum = q64.reshape(contexts, n_total, m, m).sum(dim=1)
            sum_q = torch.zeros(m, m, dtype=torch.float64)
        else:
            sum_q = q64.sum(dim=0)
    t = t.detach().to("cpu").reshape(-1)
    sa = _scalar64(nse.alpha, t.to(torch.float64)).sqrt()
    # branch of nse.py:627-654 per step; it only depends on sqrt(alpha) > 0.5, so two evaluations cover every step
    c_hi, c_lo = (combine_mode(prior, cov_jac, n_total, v, weighted) for v in (1.0, 0.0))
    if contexts:  # Lambda differs from context to context: no shared inverse, every weighted step solves per sample
        c_hi, c_lo = (abi.COMBINE_PER_SAMPLE if c == abi.COMBINE_SHARED else c for c in (c_hi, c_lo))
    comb = torch.where(sa > 0.5, c_hi, c_lo).tolist()
    if cfg is not None and weighted:  # nse.py:622-625: always the weighted solve
        comb = [abi.COMBINE_PER_SAMPLE if cov_jac else abi.COMBINE_SHARED] * len(comb)
    sched_dev, mats_dev = _sched_mats_dev(nse, t, None if t_prev is None else t_prev.detach().to("cpu").reshape(-1), eta, n_total,
                                          comb, coef, prior, cov_jac, sum_q, m, None if w64 is None else float(w64.sum()), device)
    tf = time_feat_table(nse, pn, t)
    of = _obs_feat_cached(nse, pn, x2, n_sizes)
    of_cached = cfg is None and obs_slice is None
    if cfg is not None:  # layer-0 partial of the pseudo-observation: x = 0 (PF_NSE: masked empty set, no one-hot)
        if n_sizes is None:
            x0 = torch.zeros((1,) + tuple(x2.shape[1:]), dtype=x2.dtype, device=x2.device)
            of = torch.cat((of, obs_feat_table(nse, pn, x0, None)))
        else:
            of = torch.cat((of, torch.zeros(1, of.shape[1], dtype=of.dtype)))
    if obs_slice is not None:
        of = of[obs_slice]
        if sum_q is not None:
            q64 = q64[obs_slice]
            q_cached = False
    if sum_q is not None:
        q_dev = _dev32_cached(q64, device) if q_cached else _dev32(q64, device)
    of_dev = _dev32_cached(of, device) if of_cached else _dev32(of, device)
    low_dev = high_dev = None
    if prior.kind == abi.PRIOR_UNIFORM:
        low_dev, high_dev = _dev32(prior.low, device), _dev32(prior.high, device)

    prob = abi.Problem()
    prob.n_samples = n_samples
    prob.n_obs = of.shape[0] // contexts if contexts else of.shape[0]
    prob.ctx_samples = n_samples // contexts if contexts else 0
    qsum_dev = None if ctx_qsum is None else _dev32(ctx_qsum, device)
    prob.ctx_qsum = 0 if qsum_dev is None else qsum_dev.data_ptr()
    prob.theta_dim = m
    prob.cov_mode = abi.COV_JAC if cov_jac else abi.COV_GAUSS
    prob.prior_kind = prior.kind
    prob.paired = int(paired)  # k >= 1: samples per observation row (True == 1: one row per sample)
    prob.obs_chunk = obs_chunk
    prob.n_obs_total = n_total
    prob.obs_feat = of_dev.data_ptr()
    prob.obs_prec = 0 if q_dev is None else q_dev.data_ptr()
    w_dev = None if w64 is None else _dev32(w64, device)
    prob.obs_weight = 0 if w_dev is None else w_dev.data_ptr()
    prob.prior_low = 0 if low_dev is None else low_dev.data_ptr()
    prob.prior_high = 0 if high_dev is None else high_dev.data_ptr()
    lo, hi = clip
    prob.clip_lo, prob.clip_hi = (1.0, -1.0) if lo is None else (float(lo), float(hi))
    prob.seed = seed & (2 ** 64 - 1)
    prob.sample_offset = sample_offset
    # JAC trajectories skip the Jacobian on SUM steps (nse.py:642-643 discards it there)
    step_modes = None
    if cov_jac and any(c == abi.COMBINE_SUM for c in comb):
        step_modes = np.asarray([abi.COV_GAUSS if c == abi.COMBINE_SUM else abi.COV_JAC for c in comb], np.int32)
    ws_floats = 0
    for mode in {prob.cov_mode} | (set(step_modes.tolist()) if step_modes is not None else set()):
        q = abi.Problem.from_buffer_copy(prob)
        q.cov_mode = mode
        ws_floats = max(ws_floats, lib.d4s_workspace_floats(pn.handle, C.byref(q)))
    if ws_floats <= 0:
        raise abi.D4sError("d4s_workspace_floats failed")
    ws = torch.empty(ws_floats, dtype=torch.float32, device=device)
    prob.workspace = ws.data_ptr()
    prob.workspace_floats = ws_floats
    affine_fn = net_description(nse)["affine"]
    affine_dev = x_raw = None
    if affine_fn is not None:  # closed-form part of eps: per-row [M_t | G_t | g_t] (gen_gaussian_gaussian.py:59-72)
        affine_dev = _dev32(affine_fn(t.to(torch.float64)), device)
        x_raw = x2.detach().to(torch.float32).contiguous()
        if obs_slice is not None:
            x_raw = x_raw[obs_slice].contiguous()
        prob.obs_raw = x_raw.data_ptr()
        prob.x_dim = x_raw.shape[-1]
    return TallSetup(pn, prob, sched_dev, _dev32_cached(tf, device), mats_dev, step_modes,
                     (of_dev, q_dev, low_dev, high_dev, ws, x_raw, w_dev, qsum_dev), t.numel(), affine_dev)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def philox_normal(n_rows: int, m: int, seed: int, step_index: int, sample_offset: int, device) -> torch.Tensor:
    """z[i, k] from the generator the fused update uses (d4s_philox_normal)."""
    lib = abi.require_cuda()
    out = torch.empty(n_rows, m, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        abi.check(lib.d4s_philox_normal(_ptr(out), n_rows, m, C.c_uint64(seed & (2 ** 64 - 1)), step_index,
                                        sample_offset, _stream(device)))
    return out


def run_trajectory(setup: TallSetup, theta: torch.Tensor, noise: Optional[torch.Tensor], use_graph: bool) -> torch.Tensor:
    """theta[N, m] (initial draw) is updated in place through every DDIM step and returned."""
    lib = abi.require_cuda()
    assert theta.is_cuda and theta.dtype == torch.float32 and theta.is_contiguous()
    if noise is not None:
        noise = noise.to(theta.device, torch.float32).contiguous()
        assert noise.shape == (setup.n_steps,) + tuple(theta.shape)
    modes = None
    if setup.step_modes is not None:
        modes = setup.step_modes.ctypes.data_as(C.POINTER(C.c_int32))
    with torch.cuda.device(theta.device):
        abi.check(lib.d4s_ddim_run(setup.pn.handle, C.byref(setup.prob), setup.n_steps, _ptr(setup.sched),
                                   _ptr(setup.time_feat), _ptr(setup.mats), _ptr(noise), _ptr(theta), modes,
                                   _ptr(setup.affine), 1 if use_graph else 0, _stream(theta.device)))
    return theta


def step_struct(setup: TallSetup, k: int, noise_k: Optional[torch.Tensor], update: bool) -> abi.Step:
    if noise_k is None:  # (the struct only holds pointers into the setup's own tables: reusable)
        cache = setup.__dict__.setdefault("_steps", {})
        hit = cache.get((k, update))
        if hit is None:
            hit = cache[(k, update)] = _step_struct(setup, k, None, update)
        return hit
    return _step_struct(setup, k, noise_k, update)


def _step_struct(setup: TallSetup, k: int, noise_k: Optional[torch.Tensor], update: bool) -> abi.Step:
    st = abi.Step()
    st.sched = setup.sched[k].data_ptr()
    st.time_feat = setup.time_feat[k].data_ptr()
    st.mats = setup.mats[k].data_ptr()
    st.noise = 0 if noise_k is None else noise_k.data_ptr()
    st.step_index = k
    st.update = 1 if update else 0
    st.affine = 0 if setup.affine is None else setup.affine[k].data_ptr()
    return st


def score_partials(setup: TallSetup, theta: torch.Tensor, k: int = 0, want_scores=False, want_prec=False):
    """Kernel 1(+2) alone: fills the workspace partial sums; optionally materialises scores / JAC precisions."""
    lib = abi.require_cuda()
    N, m = theta.shape
    n = setup.prob.n_obs if not setup.prob.paired else 1
    scores = torch.empty(N, n, m, dtype=torch.float32, device=theta.device) if want_scores else None
    prec = None
    if want_prec and setup.prob.cov_mode == abi.COV_JAC:
        prec = torch.empty(N, n, m, m, dtype=torch.float32, device=theta.device)
    st = step_struct(setup, k, None, False)
    with torch.cuda.device(theta.device):
        abi.check(lib.d4s_score_partials(setup.pn.handle, C.byref(setup.prob), C.byref(st), _ptr(theta), _ptr(scores),
                                         _ptr(prec), _stream(theta.device)))
    return scores, prec


def finalize(setup: TallSetup, theta: torch.Tensor, k: int = 0, noise_k: Optional[torch.Tensor] = None,
             update: bool = False, want_score: bool = True) -> Optional[torch.Tensor]:
    """Kernel 3(+4) alone: reduces the workspace, applies prior + solve; returns the total score [N, m]."""
    lib = abi.require_cuda()
    out = torch.empty_like(theta) if want_score else None
    st = step_struct(setup, k, noise_k, update)
    with torch.cuda.device(theta.device):
        abi.check(lib.d4s_finalize(setup.pn.handle, C.byref(setup.prob), C.byref(st), _ptr(theta), _ptr(out),
                                   _stream(theta.device)))
    return out


def batched_inverse(a: torch.Tensor) -> torch.Tensor:
    lib = abi.require_cuda()
    a = a.to(torch.float32).contiguous()
    if not a.is_cuda:
        raise abi.D4sError("d4s.batched_inverse: CUDA tensors only")
    m = a.shape[-1]
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        abi.check(lib.d4s_batched_inverse(_ptr(a), _ptr(out), a.numel() // (m * m), m, _stream(a.device)))
    return out


def prior_score_native(nse, theta: torch.Tensor, t, spec: PriorSpec, want_jac: bool = False):
    """Diffused prior score (and, for the uniform prior, its derivative) through d4s_prior_score."""
    if not theta.is_cuda:
        raise abi.D4sError("d4s prior scores need CUDA tensors (no CPU fallback)")
    lib = abi.require_cuda()
    m = theta.shape[-1]
    tt = torch.as_tensor(t).detach().to("cpu").reshape(1)
    rows = schedule_rows(nse, tt, None, 0.0, 0, [abi.COMBINE_SUM])  # n = 0 => (1 - n) = 1
    mats = mats_table(rows, spec, False, 0, None, m)
    dev = theta.device
    th = theta.detach().to(torch.float32).contiguous()
    sched, mats = rows.to(torch.float32).to(dev), mats.to(torch.float32).to(dev)
    low = high = None
    if spec.kind == abi.PRIOR_UNIFORM:
        low, high = spec.low.to(torch.float32).to(dev), spec.high.to(torch.float32).to(dev)
    out = torch.empty_like(th)
    jac = torch.empty_like(th) if (want_jac and spec.kind == abi.PRIOR_UNIFORM) else None
    with torch.cuda.device(dev):
        abi.check(lib.d4s_prior_score(spec.kind, _ptr(th), th.shape[0], m, _ptr(sched), _ptr(mats), _ptr(low), _ptr(high),
                                      _ptr(out), _ptr(jac), _stream(dev)))
    return (out, jac) if want_jac else out


def launch_count() -> int:
    return int(abi.load().d4s_launch_count())
