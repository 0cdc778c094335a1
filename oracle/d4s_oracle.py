"""CPU oracle for the tall-data diffusion posterior sampler -- TEST INFRASTRUCTURE ONLY.

A numpy restatement (fp32 or fp64, chosen by `dtype`) of the reference algorithm on the hot
path `NSE.ddim -> NSE.factorized_score -> tweedies_approximation (+ vp_diffused_priors,
PF_NSE)` of JuliaLinhart/diffusions-for-sbi.  Every function cites the reference file:line it
restates (paths relative to the reference checkout).  It deliberately keeps the reference's
*computational structure* (full MLP on every (sample, observation) pair, per-step inverses,
materialised [N,n,m,m] precisions, N batched solves) so that timing it is a fair stand-in for
the reference's own CPU path.

Rules: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module, and only as the checker / the CPU baseline.
The product path (diffusions-for-sbi_b200/) never imports it.

Pinning: the reference ships no golden vectors for this path ("parity unpinned" by its own
tests).  The oracle is pinned instead against outputs of the *unmodified
reference itself*, run in the build container with the stand-ins of oracle/ref_shims
(oracle/gen_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

SQRT2 = math.sqrt(2.0)
LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)

# "numpy" (default; what the parity tests use) or "torch": route the MLP GEMM+bias+ReLU of
# OracleMLP.__call__ through torch's CPU ops.  numpy's elementwise ops are single-threaded, which makes
# the numpy MLP ~2-3x slower than the reference's ATen path on 8 cores; bench.py's CPU-baseline legs
# switch this on so the reported baseline is not an unfairly slow strawman.
GEMM_BACKEND = "numpy"

try:  # vectorised erf; scipy is in the image, math.erf is the fallback
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


# --------------------------------------------------------------------------------------
# network description
# --------------------------------------------------------------------------------------
@dataclass
class OracleMLP:
    """Linear/ReLU stack without trailing activation (zuko.nn.MLP as used at nse.py:87-90)."""

    weights: List[np.ndarray]  # each [out, in]
    biases: List[np.ndarray]  # each [out]

    def astype(self, dtype) -> "OracleMLP":
        return OracleMLP([w.astype(dtype) for w in self.weights], [b.astype(dtype) for b in self.biases])

    def __call__(self, h: np.ndarray) -> np.ndarray:
        L = len(self.weights)
        lead = h.shape[:-1]
        h = h.reshape(-1, h.shape[-1])  # one [rows, in] GEMM per layer, as torch's vmap'd nn.Linear does
        if GEMM_BACKEND == "torch":  # timing legs only: the reference's own multi-threaded ATen CPU path
            import torch
            ht = torch.from_numpy(np.ascontiguousarray(h))
            for l, (w, b) in enumerate(zip(self.weights, self.biases)):
                ht = torch.addmm(torch.from_numpy(b), ht, torch.from_numpy(w).t())
                if l < L - 1:
                    ht = ht.relu_()
            h = ht.numpy()
            return h.reshape(lead + h.shape[-1:])
        for l, (w, b) in enumerate(zip(self.weights, self.biases)):
            h = h @ w.T
            h += b
            if l < L - 1:
                np.maximum(h, 0, out=h)
        return h.reshape(lead + h.shape[-1:])

    def forward_with_masks(self, h: np.ndarray):
        """Forward pass that also returns the ReLU masks D_l = (z_l > 0) of every hidden layer."""
        masks = []
        L = len(self.weights)
        lead = h.shape[:-1]
        h = h.reshape(-1, h.shape[-1])
        for l, (w, b) in enumerate(zip(self.weights, self.biases)):
            h = h @ w.T + b
            if l < L - 1:
                masks.append((h > 0).reshape(lead + h.shape[-1:]))
                h = np.maximum(h, 0)
        return h.reshape(lead + h.shape[-1:]), masks

    def vjp_rows(self, masks, in_cols: slice) -> np.ndarray:
        """Rows of d(out)/d(in[in_cols]) for every batch element, by reverse accumulation
        (what `jacrev` does at tall_posterior_sampler.py:76-85): G <- G diag(D_l) W_l.
        masks[l] has shape [..., H_l]; returns [..., out, n_in_cols]."""
        L = len(self.weights)
        batch = masks[0].shape[:-1] if masks else ()
        g = np.broadcast_to(self.weights[-1], batch + self.weights[-1].shape)  # [..., out, H_{L-1}]
        for l in range(L - 2, -1, -1):
            g = g * masks[l][..., None, :]
            w = self.weights[l]
            if l == 0:
                w = w[:, in_cols]
            lead = g.shape[:-1]
            g = (g.reshape(-1, g.shape[-1]) @ w).reshape(lead + (w.shape[-1],))
        return g


@dataclass
class OracleNet:
    """State of an `NSE` / `PF_NSE` module (nse.py:23-104, pf_nse.py:15-76) as plain arrays."""

    theta_dim: int
    x_dim: int
    freqs: np.ndarray  # [F] = arange(1, F+1) * pi  (nse.py:100)
    net: OracleMLP
    emb_theta: Optional[OracleMLP] = None  # nn.Identity when None (nse.py:31-32)
    emb_x: Optional[OracleMLP] = None
    n_max: int = 1  # >1 => PF_NSE input layout [theta, x_agg, temb, onehot(n)] (pf_nse.py:163-164)
    pf: bool = False
    # VP-SDE schedule (nse.py:163-170); the toy class overrides it (nse_toy.py:323-339)
    schedule: str = "vp16"
    beta_min: float = 0.1
    beta_d: float = 39.9
    dtype: type = np.float64
    # toy FakeFNet (embedding_nets.py:141-160): eps = real_eps + eps_max * tanh(MLP([theta, x, t])) with the analytic
    # Gaussian-posterior eps of gen_gaussian_gaussian.py:59-72; keys post_cov, inv_lik, inv_prior_mu, eps_max
    toy: Optional[dict] = None

    def astype(self, dtype) -> "OracleNet":
        return OracleNet(self.theta_dim, self.x_dim, self.freqs.astype(dtype), self.net.astype(dtype),
                         None if self.emb_theta is None else self.emb_theta.astype(dtype),
                         None if self.emb_x is None else self.emb_x.astype(dtype),
                         self.n_max, self.pf, self.schedule, self.beta_min, self.beta_d, dtype,
                         None if self.toy is None else {k: (np.asarray(v, dtype) if k != "eps_max" else v)
                                                        for k, v in self.toy.items()})

    # ---- schedule: nse.py:163-170 / nse_toy.py:323-339 --------------------------------
    def alpha(self, t):
        t = np.asarray(t, dtype=self.dtype)
        if self.schedule == "vp16":
            return np.exp(-16 * t ** 2)  # nse.py:163
        log_alpha = self.dtype(0.5 * self.beta_d) * (t ** 2) + self.dtype(self.beta_min) * t  # nse_toy.py:327
        return np.exp(self.dtype(-0.5) * log_alpha)

    def sigma(self, t):
        return np.sqrt(1 - self.alpha(t) + self.dtype(math.exp(-16)))  # nse.py:170


def time_embedding(net: OracleNet, t) -> np.ndarray:
    """nse.py:125-126: t <- freqs * t[..., None]; cat(cos, sin)."""
    t = np.asarray(t, dtype=net.dtype)
    a = net.freqs * t[..., None]
    return np.concatenate((np.cos(a), np.sin(a)), axis=-1)


def _embed(mlp: Optional[OracleMLP], v: np.ndarray) -> np.ndarray:
    return v if mlp is None else mlp(v)


# --------------------------------------------------------------------------------------
# a1/a2: NSE.forward, NSE.score (nse.py:112-145)
# --------------------------------------------------------------------------------------
def nse_forward(net: OracleNet, theta: np.ndarray, x: np.ndarray, t) -> np.ndarray:
    """eps(theta, x, t); theta[*, m], x[*, d] broadcastable batch dims, t scalar or [*]."""
    temb = time_embedding(net, t)
    th = _embed(net.emb_theta, theta)
    xe = _embed(net.emb_x, x)
    batch = np.broadcast_shapes(th.shape[:-1], xe.shape[:-1], temb.shape[:-1])  # nse.py:133
    th = np.broadcast_to(th, batch + th.shape[-1:])
    xe = np.broadcast_to(xe, batch + xe.shape[-1:])
    temb = np.broadcast_to(temb, batch + temb.shape[-1:])
    return net.net(np.concatenate((th, xe, temb), axis=-1))  # nse.py:136


def nse_score(net: OracleNet, theta, x, t) -> np.ndarray:
    if net.toy is not None:  # paired rows: FakeFNet on (theta_i, x_i)
        D = net.dtype
        xx = np.broadcast_to(x, theta.shape[:-1] + x.shape[-1:])
        inp = np.concatenate((theta, xx, np.full(theta.shape[:-1] + (1,), t, dtype=D)), -1)
        M, G, g = toy_affine(net, t)
        eps = theta @ M.T - (xx @ G.T + g) + D(net.toy["eps_max"]) * np.tanh(net.net(inp))
        return -eps / net.sigma(t)
    return -nse_forward(net, theta, x, t) / net.sigma(t)  # nse.py:144-145


# --------------------------------------------------------------------------------------
# a11: PF_NSE (pf_nse.py:78-197)
# --------------------------------------------------------------------------------------
def pf_create_data(net: OracleNet, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """pf_nse.py:169-197: split x[n,d] into ceil(n/n_max) zero-padded subsets + their sizes."""
    n_obs = x.shape[0] if x.ndim > 1 else 1
    k, r = divmod(n_obs, net.n_max)
    xs = [x[i * net.n_max:(i + 1) * net.n_max] for i in range(k)]
    ns = [np.full((1,), net.n_max, dtype=x.dtype) for _ in range(k)]
    if r > 0:
        pad = np.zeros((net.n_max - r, x.shape[-1]), dtype=x.dtype)
        xs.append(np.concatenate((x[n_obs - r:], pad), axis=0))
        ns.append(np.full((1,), r, dtype=x.dtype))
    return np.stack(xs), np.stack(ns)


def pf_forward(net: OracleNet, theta: np.ndarray, xs: np.ndarray, t, ns: np.ndarray) -> np.ndarray:
    """pf_nse.py:117-164 for theta[N,1,m], xs[1,K,n_max,d] (or [K,n_max,d]), ns[K,1] -> eps[N,K,m]."""
    if xs.ndim == 3:
        xs = xs[None]
    K = xs.shape[1]
    mask = (np.arange(net.n_max)[None, :] < ns.reshape(K, 1)).astype(net.dtype)[..., None]  # [K,n_max,1] pf_nse.py:84
    xm = mask * xs  # pf_nse.py:141
    xe = _embed(net.emb_x, xm)
    xe = mask * xe  # pf_nse.py:154
    if ns.sum() != 0:
        x_agg = xe.sum(axis=-2) / ns.reshape(K, 1)  # pf_nse.py:109-111
    else:
        x_agg = xe.mean(axis=-2)
    if not (ns == 0).all():
        onehot = np.eye(net.n_max, dtype=net.dtype)[ns.reshape(K).astype(np.int64) - 1]  # pf_nse.py:158-159
    else:
        onehot = np.zeros((K, net.n_max), dtype=net.dtype)
    temb = time_embedding(net, t)
    th = _embed(net.emb_theta, theta)
    batch = np.broadcast_shapes(th.shape[:-1], x_agg.shape[:-1], temb.shape[:-1], onehot.shape[:-1])
    parts = [np.broadcast_to(a, batch + a.shape[-1:]) for a in (th, x_agg, temb, onehot)]
    return net.net(np.concatenate(parts, axis=-1))  # pf_nse.py:163-164


def pair_inputs(net: OracleNet, theta: np.ndarray, x: np.ndarray, t, ns=None):
    """Builds the [N, n, in] input grid of the final MLP for the tall case."""
    N = theta.shape[0]
    temb = time_embedding(net, t)
    th = _embed(net.emb_theta, theta)  # [N, e_theta]
    if net.pf:
        xs = x if x.ndim == 3 else x[0]
        K = xs.shape[0]
        mask = (np.arange(net.n_max)[None, :] < ns.reshape(K, 1)).astype(net.dtype)[..., None]
        xe = mask * _embed(net.emb_x, mask * xs)
        x_agg = xe.sum(axis=-2) / ns.reshape(K, 1) if ns.sum() != 0 else xe.mean(axis=-2)
        if not (ns == 0).all():
            onehot = np.eye(net.n_max, dtype=net.dtype)[ns.reshape(K).astype(np.int64) - 1]
        else:
            onehot = np.zeros((K, net.n_max), dtype=net.dtype)
        n = K
        parts = [np.broadcast_to(th[:, None], (N, n, th.shape[-1])),
                 np.broadcast_to(x_agg[None], (N, n, x_agg.shape[-1])),
                 np.broadcast_to(temb, (N, n, temb.shape[-1])),
                 np.broadcast_to(onehot[None], (N, n, onehot.shape[-1]))]
    else:
        xe = _embed(net.emb_x, x)
        n = xe.shape[0]
        parts = [np.broadcast_to(th[:, None], (N, n, th.shape[-1])),
                 np.broadcast_to(xe[None], (N, n, xe.shape[-1])),
                 np.broadcast_to(temb, (N, n, temb.shape[-1]))]
    return np.concatenate(parts, axis=-1), th.shape[-1]


def toy_affine(net: OracleNet, t):
    """Analytic eps of the Gaussian toy task (gen_gaussian_gaussian.py:59-72) in affine form:
    eps = sigma inv(alpha S_post + (1-alpha) I) (theta - sqrt(alpha) S_post (S_pr^-1 mu + S_lik^-1 x)) = M theta - G x - g."""
    D = net.dtype
    a, sig = net.alpha(t), net.sigma(t)
    pc, il, ipm = (net.toy[k].astype(D) for k in ("post_cov", "inv_lik", "inv_prior_mu"))
    m = pc.shape[0]
    M = sig * np.linalg.inv(a * pc + (1 - a) * np.eye(m, dtype=D))
    G = (a ** 0.5) * (M @ pc @ il)
    g = (a ** 0.5) * (M @ pc @ ipm)
    return M, G, g


def pair_eps(net: OracleNet, theta, x, t, ns=None, want_jac=False):
    """eps[N, n, m] on the (sample, observation) grid (and d eps / d theta [N, n, m, m] when asked)."""
    if net.toy is None:
        inp, e_theta = pair_inputs(net, theta, x, t, ns)
        if not want_jac:
            return net.net(inp)
        eps, masks = net.net.forward_with_masks(inp)
        jac = net.net.vjp_rows(masks, slice(0, e_theta))  # [N,n,m,e_theta]
        emb_jac = _theta_jacobian_of_embedding(net, theta)
        if emb_jac is not None:
            jac = jac @ emb_jac[:, None]  # chain through the theta embedding
        return eps, jac
    # toy: FakeFNet.forward (embedding_nets.py:155-160) with the raw time as third input block
    D = net.dtype
    N, n, m = theta.shape[0], x.shape[0], theta.shape[1]
    tt = np.full((N, n, 1), t, dtype=D)
    inp = np.concatenate((np.broadcast_to(theta[:, None], (N, n, m)), np.broadcast_to(x[None], (N, n, x.shape[1])), tt), -1)
    M, G, g = toy_affine(net, t)
    real = (theta @ M.T)[:, None, :] - (x @ G.T + g)[None]
    z, masks = net.net.forward_with_masks(inp)
    th = np.tanh(z)
    eps = real + D(net.toy["eps_max"]) * th
    if not want_jac:
        return eps
    jz = net.net.vjp_rows(masks, slice(0, m))  # d z / d theta
    jac = D(net.toy["eps_max"]) * (1 - th ** 2)[..., None] * jz + M
    return eps, jac


# --------------------------------------------------------------------------------------
# a7: prec_matrix_backward (tall_posterior_sampler.py:35-43)
# --------------------------------------------------------------------------------------
def prec_matrix_backward(net: OracleNet, t, dist_cov: np.ndarray) -> np.ndarray:
    alpha_t, sigma_t = net.alpha(t), net.sigma(t)
    eye = np.eye(dist_cov.shape[-1], dtype=net.dtype)
    return np.linalg.inv(dist_cov) + (alpha_t / sigma_t ** 2) * eye


# --------------------------------------------------------------------------------------
# a6: tweedies_approximation (tall_posterior_sampler.py:46-135)
# --------------------------------------------------------------------------------------
def _theta_jacobian_of_embedding(net: OracleNet, theta: np.ndarray) -> Optional[np.ndarray]:
    """d e_theta(theta_i) / d theta_i, [N, e, m]; None for nn.Identity."""
    if net.emb_theta is None:
        return None
    _, masks = net.emb_theta.forward_with_masks(theta)
    return net.emb_theta.vjp_rows(masks, slice(None))


def tweedies_approximation(net: OracleNet, x, theta, t, dist_cov_est=None, mode="JAC", ns=None):
    """Returns (prec[N,n,m,m], mean[N,n,m], score[N,n,m]).

    GAUSS (tall_posterior_sampler.py:95-128): scores on the (theta_i, x_j) grid, precision
    inv(Sigma_j) + alpha/sigma^2 I repeated over samples.
    JAC (tall_posterior_sampler.py:60-94): J_ij = d s_ij / d theta_i by reverse accumulation,
    cov = sigma^2/alpha (I + sigma^2 J), prec = inv(cov).
    """
    alpha_t, sigma_t = net.alpha(t), net.sigma(t)
    N, m = theta.shape
    if mode == "GAUSS":
        eps = pair_eps(net, theta, x, t, ns)
        score = -eps / sigma_t
        prec = prec_matrix_backward(net, t, dist_cov_est)  # [n,m,m]
        prec = np.tile(prec[None], (N, 1, 1, 1))  # tall_posterior_sampler.py:128 (materialised)
    elif mode == "JAC":
        eps, jac_eps = pair_eps(net, theta, x, t, ns, want_jac=True)
        score = -eps / sigma_t
        jac_score = -jac_eps / sigma_t
        cov = (sigma_t ** 2 / alpha_t) * (np.eye(m, dtype=net.dtype) + (sigma_t ** 2) * jac_score)  # :91-93
        prec = np.linalg.inv(cov)  # :94
    else:
        raise NotImplementedError("Available methods are GAUSS, JAC")
    mean = 1 / (alpha_t ** 0.5) * (theta[:, None] + sigma_t ** 2 * score)  # :132
    return prec, mean, score


# --------------------------------------------------------------------------------------
# a9/a10: diffused prior scores (vp_diffused_priors.py:5-74)
# --------------------------------------------------------------------------------------
def _norm_cdf(u):
    return 0.5 * (1 + _erf(u / SQRT2))


def _norm_pdf(u):
    return np.exp(-0.5 * u * u - LOG_SQRT_2PI)


@dataclass
class UniformPrior:
    low: np.ndarray
    high: np.ndarray
    kind: str = "uniform"


@dataclass
class GaussianPrior:
    mean: np.ndarray
    cov: np.ndarray
    kind: str = "gaussian"


def vpdiff_uniform_score(net: OracleNet, prior: UniformPrior, theta, t, with_jac_diag=False):
    """vp_diffused_priors.py:20-46: f'/(f + 1e-6) per dimension.  With `with_jac_diag` also returns
    the derivative of that expression w.r.t. theta (the reference obtains it by jacrev at
    tall_posterior_sampler.py:147; it is diagonal because the score is separable)."""
    a, b = prior.low.astype(net.dtype), prior.high.astype(net.dtype)
    s, sig = net.alpha(t) ** 0.5, net.sigma(t)
    ub, ua = (b * s - theta) / sig, (a * s - theta) / sig
    f = (_norm_cdf(ub) - _norm_cdf(ua)) / s
    pb, pa = _norm_pdf(ub), _norm_pdf(ua)
    fp = -1 / sig * (pb - pa) / s
    score = fp / (f + net.dtype(1e-6))
    if not with_jac_diag:
        return score.astype(net.dtype)
    # d/dtheta: du/dtheta = -1/sig ; dPhi(u)/dtheta = -phi(u)/sig ; dphi(u)/dtheta = u phi(u)/sig
    fpp = -(ub * pb - ua * pa) / (sig ** 2 * s)
    den = f + net.dtype(1e-6)
    jac = fpp / den - fp * fp / den ** 2
    return score.astype(net.dtype), jac.astype(net.dtype)


def vpdiff_gaussian_score(net: OracleNet, prior: GaussianPrior, theta, t):
    """vp_diffused_priors.py:55-72: -(theta - sqrt(alpha) mu) @ inv(sigma^2 I + alpha Sigma)."""
    s, sig = net.alpha(t) ** 0.5, net.sigma(t)
    m = theta.shape[-1]
    loc = s * prior.mean.astype(net.dtype)
    cov = sig ** 2 * np.eye(m, dtype=net.dtype) + s ** 2 * prior.cov.astype(net.dtype)
    return (-(theta - loc) @ np.linalg.inv(cov)).astype(net.dtype)


def prior_score(net, prior, theta, t):
    if prior.kind == "uniform":
        return vpdiff_uniform_score(net, prior, theta, t)
    return vpdiff_gaussian_score(net, prior, theta, t)


# --------------------------------------------------------------------------------------
# a8: tweedies_approximation_prior (tall_posterior_sampler.py:138-155)
# --------------------------------------------------------------------------------------
def tweedies_approximation_prior(net: OracleNet, prior, theta, t):
    alpha_t, sigma_t = net.alpha(t), net.sigma(t)
    m = theta.shape[-1]
    if prior.kind == "uniform":
        score, jd = vpdiff_uniform_score(net, prior, theta, t, with_jac_diag=True)
        jac = jd[..., None] * np.eye(m, dtype=net.dtype)
    else:
        score = vpdiff_gaussian_score(net, prior, theta, t)
        s, sig = alpha_t ** 0.5, sigma_t
        cov0 = sig ** 2 * np.eye(m, dtype=net.dtype) + s ** 2 * prior.cov.astype(net.dtype)
        jac = np.broadcast_to(-np.linalg.inv(cov0).T, theta.shape[:1] + (m, m))
    mean = 1 / (alpha_t ** 0.5) * (theta + sigma_t ** 2 * score)
    cov = (sigma_t ** 2 / alpha_t) * (np.eye(m, dtype=net.dtype) + (sigma_t ** 2) * jac)
    return np.linalg.inv(cov), mean, score


# --------------------------------------------------------------------------------------
# a5: NSE.factorized_score (nse.py:544-656)
# --------------------------------------------------------------------------------------
def factorized_score(net: OracleNet, theta, x_obs, t, prior, dist_cov_est=None, cov_mode="GAUSS",
                     prior_type="gaussian", ns=None) -> np.ndarray:
    n_obs = x_obs.shape[0]  # nse.py:587 (= K subsets for PF_NSE)
    prec_0_t, _, scores = tweedies_approximation(net, x_obs, theta, t, dist_cov_est, cov_mode, ns)
    p_score = prior_score(net, prior, theta, t)
    if prior_type == "gaussian":  # nse.py:628-638
        prec_prior = prec_matrix_backward(net, t, prior.cov.astype(net.dtype))  # [m,m]
        prec_score_prior = (prec_prior @ p_score[..., None])[..., 0]
        prec_score_post = (prec_0_t @ scores[..., None])[..., 0]
        lda = prec_prior * (1 - n_obs) + prec_0_t.sum(axis=1)
        weighted = prec_score_prior + (prec_score_post - prec_score_prior[:, None]).sum(axis=1)
        return np.linalg.solve(lda, weighted[..., None])[..., 0]
    total = (1 - n_obs) * p_score + scores.sum(axis=1)  # nse.py:642
    if (net.alpha(t) ** 0.5 > 0.5) and (n_obs > 1):  # nse.py:643
        prec_prior, _, _ = tweedies_approximation_prior(net, prior, theta, t)  # [N,m,m]
        prec_score_prior = (prec_prior @ p_score[..., None])[..., 0]
        prec_score_post = (prec_0_t @ scores[..., None])[..., 0]
        lda = prec_prior * (1 - n_obs) + prec_0_t.sum(axis=1)
        weighted = prec_score_prior + (prec_score_post - prec_score_prior[:, None]).sum(axis=1)
        total = np.linalg.solve(lda, weighted[..., None])[..., 0]
    return total


# --------------------------------------------------------------------------------------
# a15: classifier-free guidance branches (nse.py:598-625, 532-538)
# --------------------------------------------------------------------------------------
def _cfg_context(net: OracleNet, x_obs, ns):
    """x_ = zeros_like(x_obs[0][None]) and, for PF_NSE, n = zeros_like(n[0][None]) (nse.py:599-606).  For a plain NSE the
    reference leaves `kwargs_score_prior` unassigned (UnboundLocalError); the intended value {} is used here."""
    x0 = np.zeros_like(x_obs[:1])
    ns0 = None if ns is None else np.zeros((1, 1), dtype=ns.dtype)
    return x0, ns0


def factorized_score_cfg(net: OracleNet, theta, x_obs, t, dist_cov_est=None, dist_cov_est_prior=None, cov_mode="GAUSS",
                         ns=None):
    n_obs = x_obs.shape[0]
    prec, _, scores = tweedies_approximation(net, x_obs, theta, t, dist_cov_est, cov_mode, ns)
    x0, ns0 = _cfg_context(net, x_obs, ns)
    prec_p, _, s_p = tweedies_approximation(net, x0, theta, t, dist_cov_est_prior, cov_mode, ns0)  # nse.py:607-615
    ps_prior = (prec_p @ s_p[..., None])[..., 0][:, 0, :]
    ps_post = (prec @ scores[..., None])[..., 0]
    lda = prec_p[:, 0] * (1 - n_obs) + prec.sum(axis=1)
    weighted = ps_prior + (ps_post - ps_prior[:, None]).sum(axis=1)
    return np.linalg.solve(lda, weighted[..., None])[..., 0]  # nse.py:625


def factorized_score_geffner_cfg(net: OracleNet, theta, x, t, ns=None):
    """nse.py:532-541 with clf_free_guidance=True."""
    n_obs = x.shape[0]
    scores = -pair_eps(net, theta, x, t, ns) / net.sigma(t)
    x0, ns0 = _cfg_context(net, x, ns)
    s_p = -pair_eps(net, theta, x0, t, ns0) / net.sigma(t)
    return (1 - n_obs) * s_p[:, 0] + scores.sum(axis=1)


def ddim_cfg(net: OracleNet, x, z0, z, steps, eta, dist_cov_est, dist_cov_est_prior, cov_mode="GAUSS", ns=None):
    D = net.dtype
    time = time_grid(steps).astype(D)
    theta = z0.astype(D)
    for k in range(steps):
        score = factorized_score_cfg(net, theta, x, time[k], dist_cov_est, dist_cov_est_prior, cov_mode, ns)
        theta = ddim_step(net, theta, score, time[k], 1 / steps, eta, z[k].astype(D))
    return theta


# --------------------------------------------------------------------------------------
# a3/a4: NSE.ddim, ddim_step, mean_pred, bridge_mean (nse.py:199-299)
# --------------------------------------------------------------------------------------
def time_grid(steps: int) -> np.ndarray:
    """torch.linspace(1, 0, steps+1) in fp32 (nse.py:258).  torch evaluates start + step*i for the first
    half and end - step*(steps-i) for the second half, with step = (end-start)/steps in fp32 and a fused
    multiply-add (emulated here by forming the exact product/sum in fp64 and rounding once)."""
    n = steps + 1
    step = np.float64(np.float32((np.float32(0.0) - np.float32(1.0)) / np.float32(steps)))
    i = np.arange(n, dtype=np.float64)
    half = n // 2
    lo = 1.0 + step * i
    hi = 0.0 - step * (steps - i)
    return np.where(i < half, lo, hi).astype(np.float32)


def ddim_step(net: OracleNet, theta, score, t, dt, eta, z):
    """nse.py:281-298 with the noise z injected instead of torch.randn_like."""
    D = net.dtype
    alpha_t = net.alpha(t)
    alpha_t_1 = net.alpha(np.asarray(t, dtype=D) - D(dt))
    bridge_std = D(eta) * ((((1 - alpha_t_1) / (1 - alpha_t)) * (1 - alpha_t / alpha_t_1)) ** 0.5)  # :285-287
    theta_0 = (alpha_t ** (-0.5)) * (theta + (1 - alpha_t) * score)  # mean_pred nse.py:205-206
    est_noise = (theta - (alpha_t ** 0.5) * theta_0) / ((1 - alpha_t) ** 0.5)  # bridge_mean :218
    mean = (alpha_t_1 ** 0.5) * theta_0 + ((1 - alpha_t_1 - bridge_std ** 2) ** 0.5) * est_noise  # :219-221
    return (mean + z * bridge_std).astype(D)


def ddim(net: OracleNet, x, z0, z, steps, eta, prior=None, dist_cov_est=None, cov_mode="GAUSS",
         prior_type="gaussian", clip=(None, None), ns=None, return_path=False):
    """nse.py:223-267 (tall branch) with injected initial draw z0[N,m] and per-step noise z[steps,N,m]."""
    D = net.dtype
    time = time_grid(steps).astype(D)  # `.to(x)` casts the fp32 grid to x's dtype
    dt = 1 / steps
    theta = z0.astype(D)
    path = []
    for k in range(steps):
        t = time[k]
        score = factorized_score(net, theta, x, t, prior, dist_cov_est, cov_mode, prior_type, ns)
        theta = ddim_step(net, theta, score, t, dt, eta, z[k].astype(D))
        if clip[0] is not None:
            theta = np.clip(theta, clip[0], clip[1])
        if return_path:
            path.append(theta.copy())
    return (theta, np.stack(path)) if return_path else theta


def ddim_single(net: OracleNet, x, z0, z, steps, eta, clip=(None, None)):
    """nse.py:253-254 branch: plain `score` (one observation broadcast, or paired rows)."""
    D = net.dtype
    time = time_grid(steps).astype(D)
    theta = z0.astype(D)
    for k in range(steps):
        t = time[k]
        score = nse_score(net, theta, x, t)
        theta = ddim_step(net, theta, score, t, 1 / steps, eta, z[k].astype(D))
        if clip[0] is not None:
            theta = np.clip(theta, clip[0], clip[1])
    return theta


# --------------------------------------------------------------------------------------
# a16: Langevin / Geffner family (nse.py:301-542, 658-705)
# --------------------------------------------------------------------------------------
def factorized_score_geffner(net: OracleNet, theta, x, t, prior, ns=None):
    """nse.py:479-542 (analytic prior): (1 - n) prior_score + sum_j score_j."""
    n_obs = x.shape[0] if x.ndim > 1 else 1
    scores = -pair_eps(net, theta, x if x.ndim > 1 else x[None], t, ns) / net.sigma(t)
    return (1 - n_obs) * prior_score(net, prior, theta, t) + scores.sum(axis=1)


def _gamma(net: OracleNet, time, i, steps):
    D = net.dtype
    t = time[i]
    if i < steps - 1:
        return net.alpha(t) / net.alpha(np.asarray(t, dtype=D) - time[-2])  # nse.py:456-457
    return net.alpha(t)


def annealed_langevin_geffner(net: OracleNet, x, z0, z, prior, steps=400, lsteps=5, tau=0.5, clip=(None, None), ns=None):
    """nse.py:418-475 with injected noise z[steps*lsteps, N, m]."""
    D = net.dtype
    time = time_grid(steps).astype(D)
    theta = z0.astype(D)
    k = 0
    single = x.ndim == 1 or x.shape[0] == 1
    for i in range(steps):
        t = time[i]
        gamma_t = _gamma(net, time, i, steps)
        delta = D(tau) * (1 - gamma_t) / (gamma_t ** 0.5)
        for _ in range(lsteps):
            zz = z[k].astype(D)
            k += 1
            if single:
                score = nse_score(net, theta, x, t)
            else:
                score = factorized_score_geffner(net, theta, x, t, prior, ns)
            theta = theta + delta * score + ((2 * delta) ** 0.5) * zz
        if clip[0] is not None:
            theta = np.clip(theta, clip[0], clip[1])
    return theta.astype(D)


def deterministic_gaussian_geffner(net: OracleNet, x, z0, z, prior, steps=1000, clip=(None, None)):
    """nse.py:658-705 with injected noise z[steps, N, m]."""
    D = net.dtype
    time = time_grid(steps).astype(D)
    theta = z0.astype(D)
    n = x.shape[0] if x.ndim > 1 else 1
    for i in range(steps):
        t = time[i]
        g = _gamma(net, time, i, steps)
        scores = -pair_eps(net, theta, x, t) / net.sigma(t)
        p_score = prior_score(net, prior, theta, t)
        mu = 1 / g ** 0.5 * (n * theta + (1 - g) * scores.sum(axis=1)) - (n - 1) * g ** 0.5 * theta  # :688
        mu = mu * (1 / (n - g * (n - 1)))
        var = (1 - g) / (n - g * (n - 1))
        mu = mu + var * (1 - n) * p_score
        theta = mu + z[i].astype(D) * var ** 0.5
        if clip[0] is not None:
            theta = np.clip(theta, clip[0], clip[1])
    return theta.astype(D)


def predictor_corrector(net: OracleNet, x, z0, z, prior, steps, predictor="ddim", corrector="langevin", eta=1.0, lsteps=5,
                        r=0.5, clip=(None, None), dist_cov_est=None, cov_mode="GAUSS", prior_type="gaussian", ns=None):
    """nse.py:340-416 (tall branch) with injected noise, consumed in call order: predictor draw, then one per
    corrector step.  corrector 'langevin' uses the F-NPSE score (nse.py:381), corrector 'id' the GAUSS/JAC score."""
    D = net.dtype
    time = time_grid(steps).astype(D)
    dt = 1 / steps
    theta = z0.astype(D)
    k = 0

    def score_fun(th, t):
        if corrector == "langevin":
            return factorized_score_geffner(net, th, x, t, prior, ns)
        return factorized_score(net, th, x, t, prior, dist_cov_est, cov_mode, prior_type, ns)

    for i in range(steps):
        t = time[i]
        if predictor == "ddim":
            s = score_fun(theta, t)
            theta = ddim_step(net, theta, s, t, dt, eta, z[k].astype(D))
            k += 1
        if corrector == "langevin":
            tc = np.asarray(t, dtype=D) - D(dt)
            for _ in range(lsteps):
                zz = z[k].astype(D)
                k += 1
                g = score_fun(theta, tc)
                eps = D(r) * (net.alpha(tc) ** 0.5) * (net.sigma(tc) ** 2)  # nse.py:335
                tamed = (eps / (1 + eps * np.linalg.norm(g, axis=-1)))[..., None]
                theta = theta + tamed * g + ((2 * eps) ** 0.5) * zz
        if clip[0] is not None:
            theta = np.clip(theta, clip[0], clip[1])
    return theta.astype(D)


def euler_sde_sampler(net: OracleNet, x, z0, z, prior, dist_cov_est=None, cov_mode="GAUSS", prior_type="gaussian",
                      clip=(None, None)):
    """tall_posterior_sampler.py:217-281 driven by diffused_tall_posterior_score (:158-214), injected noise z[999,N,m]."""
    D = net.dtype
    time_pts = np.linspace(np.float32(1), np.float32(0), 1000, dtype=np.float32).astype(D)
    theta = z0.astype(D)
    for i in range(999):
        t = time_pts[i]
        dt = time_pts[i + 1] - t
        beta = 32 * t if net.schedule == "vp16" else D(net.beta_d) * t + D(net.beta_min)  # nse.py:147-149
        score = factorized_score(net, theta, x, t, prior, dist_cov_est, cov_mode, prior_type)
        drift = -0.5 * beta * theta - beta * score
        theta = theta + drift * dt + beta ** 0.5 * z[i].astype(D) * abs(dt) ** 0.5
        if clip[0] is not None:
            theta = np.clip(theta, clip[0], clip[1])
    return theta.astype(D)


# --------------------------------------------------------------------------------------
# helpers used by tests / bench (not part of the reference)
# --------------------------------------------------------------------------------------
def random_net(rng: np.random.Generator, m, d, hidden=(256, 256, 256), freqs=3, n_max=1, pf=False,
               dtype=np.float64) -> OracleNet:
    """Random-init net of a task's shape (nn.Linear default init: U(-1/sqrt(in), 1/sqrt(in)))."""
    n_in = m + d + 2 * freqs + (n_max if pf else 0)
    widths = [n_in, *hidden, m]
    ws, bs = [], []
    for a, b in zip(widths[:-1], widths[1:]):
        k = 1 / math.sqrt(a)
        ws.append(rng.uniform(-k, k, size=(b, a)).astype(dtype))
        bs.append(rng.uniform(-k, k, size=(b,)).astype(dtype))
    return OracleNet(m, d, (np.arange(1, freqs + 1) * math.pi).astype(dtype), OracleMLP(ws, bs),
                     n_max=n_max, pf=pf, dtype=dtype)


def rel_l2(a, b) -> float:
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
