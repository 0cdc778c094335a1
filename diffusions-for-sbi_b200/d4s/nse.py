"""`NSE`: the reference's neural score estimator API (nse.py:23-299, 544-656 of diffusions-for-sbi) with the
tall-data sampling path executed by the sm_100a kernels of libd4s.

Kept verbatim from the reference: constructor signature, `forward/score`, the VP-SDE schedule functions
`alpha/sigma/beta/f/g`, `mean_pred/bridge_mean`, and the signatures + keyword contract of `ddim`,
`ddim_step` and `factorized_score` (kwargs `prior, prior_type, prior_score_fn, dist_cov_est, cov_mode,
n`, forwarded through `ddim(**kwargs)` exactly like nse.py:254-256).

Different by design: `ddim` and `factorized_score` do not evaluate the network pair by pair in torch; they
hand the whole trajectory / call to the C ABI (engine.py).  They require CUDA tensors and raise otherwise:
there is no CPU or torch fallback on this path.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Optional

import torch
import torch.nn as nn
from torch import Size, Tensor

from . import _cabi as abi
from . import engine
from . import generic
from .nn import MLP


def _use_graph_default() -> bool:
    return os.environ.get("D4S_CUDA_GRAPH", "1") != "0"


class NSE(nn.Module):
    """Noise-parametrised score network eps(theta, x, t) for the VP SDE with beta(t) = 32 t (nse.py:23-104)."""

    def __init__(self, theta_dim: int, x_dim: int, freqs: int = 3, build_net: Callable[..., nn.Module] = MLP,
                 net_type: str = "default", embedding_nn_theta: nn.Module = nn.Identity(),
                 embedding_nn_x: nn.Module = nn.Identity(), **kwargs) -> None:
        super().__init__()
        if net_type != "default":
            raise NotImplementedError("net_type='fnet' is a dead configuration of the reference (SURVEY.md section 2)")
        self.embedding_nn_theta = embedding_nn_theta
        self.embedding_nn_x = embedding_nn_x
        self.theta_emb_dim, self.x_emb_dim = self.get_theta_x_embedding_dim(theta_dim, x_dim)
        self.net_type = net_type
        self.n_max = 1
        self.net = build_net(self.theta_emb_dim + self.x_emb_dim + 2 * freqs, theta_dim, **kwargs)
        self.register_buffer("freqs", torch.arange(1, freqs + 1) * math.pi)
        self.register_buffer("zeros", torch.zeros(theta_dim))
        self.register_buffer("ones", torch.ones(theta_dim))

    # The reference stores its Tweedie hook as an instance attribute (nse.py:98) and scripts re-assign it after
    # unpickling (sbibm_posterior_estimation.py:602).  It is accepted and ignored: the fused kernels never
    # materialise the [N, n, m, m] tensors the hook returns.  Use d4s.tweedies_approximation for the hook API.
    @property
    def tweedies_approximator(self):
        from .tall_posterior_sampler import tweedies_approximation
        return self.__dict__.get("_tweedies_approximator", tweedies_approximation)

    @tweedies_approximator.setter
    def tweedies_approximator(self, fn):
        self.__dict__["_tweedies_approximator"] = fn

    def get_theta_x_embedding_dim(self, theta_dim: int, x_dim: int):
        with torch.no_grad():
            e_t = self.embedding_nn_theta(torch.ones(1, theta_dim)).shape[-1]
            e_x = self.embedding_nn_x(torch.ones(1, x_dim)).shape[-1]
        return e_t, e_x

    # ---- model definition (torch; used for training and ad-hoc evaluation, not by the samplers) -----------
    def forward(self, theta: Tensor, x: Tensor, t: Tensor) -> Tensor:
        """eps(theta, x, t) with theta[*, m], x[*, d], t[*] broadcast over the batch dims (nse.py:112-136)."""
        ang = self.freqs * t[..., None]
        temb = torch.cat((ang.cos(), ang.sin()), dim=-1)
        th = self.embedding_nn_theta(theta)
        xe = self.embedding_nn_x(x)
        batch = torch.broadcast_shapes(th.shape[:-1], xe.shape[:-1], temb.shape[:-1])
        feats = [v.expand(batch + v.shape[-1:]) for v in (th, xe, temb)]
        return self.net(torch.cat(feats, dim=-1))

    def score(self, theta: Tensor, x: Tensor, t: Tensor, **kwargs) -> Tensor:
        return -self(theta, x, t) / self.sigma(t)  # nse.py:144-145

    # ---- VP SDE: d theta = -0.5 beta(t) theta dt + sqrt(beta(t)) dW (nse.py:147-170) -----------------------
    def beta(self, t: Tensor) -> Tensor:
        return 32 * t

    def f(self, t: Tensor) -> Tensor:
        return -0.5 * self.beta(t)

    def g(self, t: Tensor) -> Tensor:
        return torch.sqrt(self.beta(t))

    def alpha(self, t: Tensor) -> Tensor:
        return torch.exp(-16 * t ** 2)

    def sigma(self, t: Tensor) -> Tensor:
        return torch.sqrt(1 - self.alpha(t) + math.exp(-16))

    def mean_pred(self, theta: Tensor, score: Tensor, alpha_t: Tensor, **kwargs) -> Tensor:
        return (alpha_t ** (-0.5)) * (theta + (1 - alpha_t) * score)  # nse.py:199-207

    def bridge_mean(self, alpha_t: Tensor, alpha_t_1: Tensor, theta_t: Tensor, theta_0: Tensor, bridge_std: float) -> Tensor:
        eps_hat = (theta_t - (alpha_t ** 0.5) * theta_0) / ((1 - alpha_t) ** 0.5)  # nse.py:209-221
        return (alpha_t_1 ** 0.5) * theta_0 + ((1 - alpha_t_1 - bridge_std ** 2) ** 0.5) * eps_hat

    # ---- helpers ------------------------------------------------------------------------------------------
    @staticmethod
    def _require_cuda(*tensors):
        for v in tensors:
            if isinstance(v, Tensor) and not v.is_cuda:
                raise abi.D4sError("diffusions-for-sbi_b200 samplers need CUDA tensors (sm_100a); there is no CPU "
                                   "fallback -- move the module and its inputs to the GPU")
        abi.require_cuda()

    def _is_single_context(self, shape: Size, x: Tensor) -> bool:
        # nse.py:253: paired rows, a single 1-D observation, or one observation row use the plain score
        return x.shape[0] == shape[0] or x.ndim == 1 or x.shape[0] == 1

    def _split_kwargs(self, kwargs: dict):
        kw = dict(kwargs)
        opts = dict(
            prior=kw.pop("prior", None),
            prior_type=kw.pop("prior_type", "gaussian"),
            prior_score_fn=kw.pop("prior_score_fn", None),
            dist_cov_est=kw.pop("dist_cov_est", None),
            dist_cov_est_prior=kw.pop("dist_cov_est_prior", None),
            cov_mode=kw.pop("cov_mode", "GAUSS"),
            clf_free_guidance=kw.pop("clf_free_guidance", False),
            n_sizes=kw.pop("n", None),
        )
        noise = kw.pop("_noise", None)  # (z0[N,m], z[steps,N,m]) injected noise for replay tests
        seed = kw.pop("_seed", None)
        sample_offset = kw.pop("_sample_offset", 0)
        if kw:
            raise TypeError(f"unexpected keyword arguments: {sorted(kw)}")
        return opts, noise, seed, sample_offset

    @staticmethod
    def _cfg(opts):
        """Classifier-free guidance (nse.py:598-625): prior score and precision from the net itself at x = 0."""
        return dict(prior_cov=opts["dist_cov_est_prior"]) if opts["clf_free_guidance"] else None

    def _draw_seed(self) -> int:
        # one draw from torch's global CPU generator, so torch.manual_seed(...) makes runs reproducible
        return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())

    # ---- samplers -----------------------------------------------------------------------------------------
    def ddim(self, shape: Size, x: Tensor, steps: int = 1000, eta: float = 1.0, verbose: bool = False,
             theta_clipping_range=(None, None), **kwargs) -> Tensor:
        """DDIM sampler adapted to n context observations (nse.py:223-267); returns theta[shape[0], m] on x.device.

        The whole `steps`-long loop is enqueued by one C-ABI call (`d4s_ddim_run`): ONE launch of the fused tcgen05
        trajectory kernel for GAUSS-type runs (and for the runs of plain-sum steps of a JAC trajectory), ONE launch of the
        paired-mode tcgen05 kernel when every sample has its own observation; JAC steps inside the alpha-gate (nse.py:643)
        run the tcgen05 score + Jacobian kernel and kernel 3/4 per step, toy nets the fp32 SIMT kernels (captured in a
        CUDA graph when the steps are short).
        """
        self._require_cuda(x)
        opts, noise, seed, sample_offset = self._split_kwargs(kwargs)
        N = int(shape[0])
        m = self.zeros.shape[0]
        grid = engine.time_grid(steps)  # nse.py:258
        t, t_prev = grid[:-1], grid[:-1].to(torch.float64) - 1.0 / steps  # nse.py:259, 284
        single = self._is_single_context(shape, x)
        if single:
            paired = x.ndim > 1 and x.shape[0] == N and N > 1
            prior = engine.PriorSpec(abi.PRIOR_NONE)
            weighted = False
        else:
            paired = False
            prior = (engine.PriorSpec(abi.PRIOR_NONE) if opts["clf_free_guidance"]
                     else engine.prior_spec(opts["prior_type"], opts["prior"], opts["prior_score_fn"], m))
            weighted = True
        if seed is None:
            seed = self._draw_seed()
        cfg = None if single else self._cfg(opts)
        if cfg is not None:
            prior = engine.PriorSpec(abi.PRIOR_NONE)
        if noise is not None:
            z0, z = noise
            theta = z0.to(x.device, torch.float32).contiguous().clone()
        else:
            z = None
            theta = engine.philox_normal(N, m, seed, -1, sample_offset, x.device)  # nse.py:261
        if not generic.fusable(self):  # nets libd4s cannot pack (other modules, theta_dim > 10): torch scores + kernels 3/4
            if cfg is not None or opts["n_sizes"] is not None:
                raise NotImplementedError("generic path: classifier-free guidance / PF_NSE subsets need a fusable net")
            run = generic.GenericTall(self, x, N, t, t_prev, float(eta), prior, opts["cov_mode"], opts["dist_cov_est"],
                                      toy=False, paired=paired, weighted=weighted, clip=theta_clipping_range, seed=seed,
                                      sample_offset=sample_offset)
            return run.run(theta, z)
        setup = engine.build_setup(self, x, N, t, t_prev, float(eta), prior, opts["cov_mode"], opts["dist_cov_est"],
                                   paired=paired, n_sizes=opts["n_sizes"], clip=theta_clipping_range, seed=seed,
                                   sample_offset=sample_offset, weighted=weighted, cfg=cfg)
        return engine.run_trajectory(setup, theta, z, _use_graph_default())

    def ddim_contexts(self, shape: Size, x: Tensor, steps: int = 1000, eta: float = 1.0, verbose: bool = False,
                      theta_clipping_range=(None, None), **kwargs) -> Tensor:
        """`torch.func.vmap(lambda x_c, cov_c: self.ddim(shape, x_c, ..., dist_cov_est=cov_c))(x, dist_cov_est)` as ONE
        trajectory launch (the custom kernels have no vmap rule; jrnnm_posterior_estimation.py:286-294,
        sbibm_posterior_estimation.py:306-314).

        x: [B, n, d] (B independent contexts of n observations each) or [B, d] (one observation per context: the plain
        single-observation posterior, i.e. the covariance pre-pass).  dist_cov_est: [B, n, m, m] for GAUSS.
        Returns theta[B, shape[0], m].  Same keyword contract as `ddim`; `_noise` is (z0[B*S, m], z[steps, B*S, m])."""
        self._require_cuda(x)
        opts, noise, seed, sample_offset = self._split_kwargs(kwargs)
        if opts["clf_free_guidance"]:
            raise NotImplementedError("ddim_contexts: classifier-free guidance is not batched")
        S, m = int(shape[0]), self.zeros.shape[0]
        B = x.shape[0]
        N = B * S
        grid = engine.time_grid(steps)
        t, t_prev = grid[:-1], grid[:-1].to(torch.float64) - 1.0 / steps
        if seed is None:
            seed = self._draw_seed()
        pf = opts["n_sizes"] is not None  # PF_NSE: x is [B, K, n_max, d], `n` the subset sizes [B, K, 1]
        if x.ndim == 2 or x.shape[1] == 1:  # one observation (PF_NSE: one subset) per context: plain score, sample i reads row i // S
            xr = x.reshape((B,) + tuple(x.shape[2:])) if pf else x.reshape(B, -1)
            setup = engine.build_setup(self, xr.contiguous(), N, t, t_prev, float(eta),
                                       engine.PriorSpec(abi.PRIOR_NONE), opts["cov_mode"], None, paired=S,
                                       n_sizes=opts["n_sizes"].reshape(B, 1) if pf else None,
                                       clip=theta_clipping_range, seed=seed, sample_offset=sample_offset, weighted=False)
        else:
            n = x.shape[1]
            prior = engine.prior_spec(opts["prior_type"], opts["prior"], opts["prior_score_fn"], m)
            cov = opts["dist_cov_est"]
            if cov is not None:
                cov = cov.reshape(B * n, m, m)
            # PF_NSE: a subset is one "observation" of its context
            xs = x.reshape((B * n,) + tuple(x.shape[2:])) if pf else x.reshape(B * n, -1)
            ns = opts["n_sizes"].reshape(B * n, 1) if pf else None
            setup = engine.build_setup(self, xs, N, t, t_prev, float(eta), prior, opts["cov_mode"], cov, n_sizes=ns,
                                       clip=theta_clipping_range, seed=seed, sample_offset=sample_offset, contexts=B)
        if noise is not None:
            z0, z = noise
            theta = z0.to(x.device, torch.float32).contiguous().clone()
        else:
            z = None
            theta = engine.philox_normal(N, m, seed, -1, sample_offset, x.device)
        return engine.run_trajectory(setup, theta, z, _use_graph_default()).reshape(B, S, m)

    def estimate_dist_cov(self, x: Tensor, nsamples: int = 1000, steps: int = 100, eta: float = 0.5, **kwargs) -> Tensor:
        """Covariance pre-pass of the GAUSS mode (sbibm_posterior_estimation.py:306-314): `nsamples` draws from every
        single-observation posterior p(theta | x_j) in one launch, then their covariance.  x: [n, d]; returns [n, m, m]."""
        samples = self.ddim_contexts((nsamples,), x, steps=steps, eta=eta, **kwargs)  # [n, nsamples, m]
        c = samples - samples.mean(dim=1, keepdim=True)
        return c.transpose(1, 2) @ c / (nsamples - 1)  # == vmap(lambda s: torch.cov(s.mT))

    def ddim_step(self, theta: Tensor, x: Tensor, t: Tensor, score_fun: Callable[[Tensor, Tensor, Tensor], Tensor],
                  dt: float, eta: float, **kwargs) -> Tensor:
        """One DDIM step with a caller-supplied score function (nse.py:269-299); the update runs in kernel 4."""
        self._require_cuda(theta)
        score = score_fun(theta, x, t).detach()
        noise = kwargs.pop("_noise", None)
        if noise is not None:  # the kernel reads raw fp32 device memory
            noise = torch.as_tensor(noise).detach().to(theta.device, torch.float32).contiguous()
            if tuple(noise.shape) != tuple(theta.shape):
                raise ValueError(f"_noise must have the shape of theta {tuple(theta.shape)}, got {tuple(noise.shape)}")
        tt = torch.as_tensor(t).detach().to("cpu").reshape(1)
        rows = engine.schedule_rows(self, tt, tt.to(torch.float64) - dt, float(eta), 1, [abi.COMBINE_SUM])
        sched = rows.to(torch.float32).to(theta.device)
        out = theta.detach().to(torch.float32).contiguous().clone()
        sc = score.to(torch.float32).contiguous()
        lib = abi.require_cuda()
        import ctypes as C
        seed = self._draw_seed() if noise is None else 0
        with torch.cuda.device(theta.device):
            abi.check(lib.d4s_ddim_update(engine._ptr(out), engine._ptr(sc), out.shape[0], out.shape[1],
                                          engine._ptr(sched), engine._ptr(noise), C.c_uint64(seed), 0, 0, 1.0, -1.0,
                                          engine._stream(theta.device)))
        return out

    def factorized_score(self, theta, x_obs, t, prior_score_fn, prior, dist_cov_est=None, dist_cov_est_prior=None,
                         cov_mode="GAUSS", prior_type="gaussian", clf_free_guidance=False, **kwargs) -> Tensor:
        """Precision-weighted tall-data score ("GAUSS"/"JAC", nse.py:544-656), one fused call: kernel 1(+2) then 3."""
        self._require_cuda(theta, x_obs)
        n_sizes = kwargs.pop("n", None)
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        m = self.zeros.shape[0]
        # Every table of the call (layer-0 partials of the observations, inv(Sigma_j), schedule row, step matrices) is a
        # function of (weights, x_obs, dist_cov_est, prior, t): a loop that calls this with the same inputs -- the
        # reference's euler_sde_sampler, user-written samplers -- reuses them instead of rebuilding them on the host.
        tt = torch.as_tensor(t).detach().to("cpu").reshape(1)
        if not generic.fusable(self):
            if clf_free_guidance or n_sizes is not None:
                raise NotImplementedError("generic path: classifier-free guidance / PF_NSE subsets need a fusable net")
            run = generic.GenericTall(self, x_obs, theta.shape[0], tt, None, 0.0,
                                      engine.prior_spec(prior_type, prior, prior_score_fn, m), cov_mode, dist_cov_est, toy=False)
            return run.step(0, theta.detach().to(torch.float32).contiguous(), update=False, want_score=True)
        pn = engine.packed_net(self, x_obs.device)
        key = ("fs", id(pn), engine.tensor_key(x_obs), engine.tensor_key(dist_cov_est), engine.tensor_key(dist_cov_est_prior),
               int(theta.shape[0]), float(tt), cov_mode, prior_type, id(prior_score_fn), id(prior), bool(clf_free_guidance),
               engine.tensor_key(n_sizes), tuple((q.data_ptr(), q._version) for q in self.embedding_nn_x.parameters()),
               abi.options_key())
        setup = engine._setup_cache.get(key)
        if setup is None:
            cfg = dict(prior_cov=dist_cov_est_prior) if clf_free_guidance else None
            spec = (engine.PriorSpec(abi.PRIOR_NONE) if clf_free_guidance
                    else engine.prior_spec(prior_type, prior, prior_score_fn, m))
            setup = engine.build_setup(self, x_obs, theta.shape[0], tt, None, 0.0, spec, cov_mode, dist_cov_est,
                                       n_sizes=n_sizes, cfg=cfg)
            if setup.step_modes is not None:
                setup.prob.cov_mode = int(setup.step_modes[0])
            engine._setup_cache.put(key, setup, (pn, x_obs, dist_cov_est, dist_cov_est_prior, prior_score_fn, prior, n_sizes))
        th = theta.detach().to(torch.float32).contiguous()
        engine.score_partials(setup, th, 0)
        return engine.finalize(setup, th, 0)

    # ---- Langevin / Geffner family (nse.py:301-542, 658-705): same kernels, SUM combination, other update rules ----
    def _geffner_prior(self, prior_score_fun, m):
        """F-NPSE uses the analytic diffused prior score only (nse.py:540): read its parameters off the d4s object."""
        kind = getattr(prior_score_fun, "kind", None)
        if kind not in ("uniform", "gaussian"):
            raise abi.D4sError("the Langevin/Geffner samplers need a d4s.vp_diffused_priors score object "
                               "(get_vpdiff_uniform_score / get_vpdiff_gaussian_score) as prior score function")
        return prior_score_fun._spec()

    def _run_rows(self, shape, x, t_rows, coef, prior, weighted, opts, clip, noise, seed, sample_offset, cfg=None):
        """Runs a trajectory whose steps are arbitrary schedule rows (time + update coefficients per row)."""
        N, m = int(shape[0]), self.zeros.shape[0]
        if seed is None:
            seed = self._draw_seed()
        setup = engine.build_setup(self, x, N, t_rows, None, 0.0, prior, opts["cov_mode"] if weighted else "GAUSS",
                                   opts["dist_cov_est"] if weighted else None, n_sizes=opts["n_sizes"], clip=clip,
                                   seed=seed, sample_offset=sample_offset, weighted=weighted, coef=coef, cfg=cfg)
        if noise is not None:
            z0, z = noise
            theta = z0.to(x.device, torch.float32).contiguous().clone()
        else:
            z = None
            theta = engine.philox_normal(N, m, seed, -1, sample_offset, x.device)
        return engine.run_trajectory(setup, theta, z, _use_graph_default())

    def _gammas(self, grid: Tensor):
        """gamma_t = alpha(t) / alpha(t - time[-2]) for all but the last step, alpha(t) for the last (nse.py:456-459)."""
        t64 = grid[:-1].to(torch.float64)
        a = engine._scalar64(self.alpha, t64)
        a_prev = engine._scalar64(self.alpha, t64 - float(grid[-2]))
        gamma = a / a_prev
        gamma[-1] = a[-1]
        return gamma

    def factorized_score_geffner(self, theta: Tensor, x: Tensor, t: Tensor, prior_score_fun=None,
                                 clf_free_guidance: bool = False, **kwargs) -> Tensor:
        """F-NPSE score (1 - n) prior_score + sum_j score_j (nse.py:479-542), one fused call."""
        self._require_cuda(theta, x)
        m = self.zeros.shape[0]
        spec = engine.PriorSpec(abi.PRIOR_NONE) if clf_free_guidance else self._geffner_prior(prior_score_fun, m)
        tt = torch.as_tensor(t).detach().to("cpu").reshape(1)
        x2 = x if x.ndim > 1 else x[None]
        setup = engine.build_setup(self, x2, theta.shape[0], tt, None, 0.0, spec, "GAUSS", None,
                                   n_sizes=kwargs.get("n"), weighted=False, cfg={} if clf_free_guidance else None)
        th = theta.detach().to(torch.float32).contiguous()
        engine.score_partials(setup, th, 0)
        return engine.finalize(setup, th, 0)

    def annealed_langevin_geffner(self, shape: Size, x: Tensor, prior_score_fn, prior_type: Optional[str] = None,
                                  steps: int = 400, lsteps: int = 5, tau: float = 0.5,
                                  theta_clipping_range=(None, None), verbose: bool = False, **kwargs) -> Tensor:
        """Annealed Langevin dynamics with the F-NPSE score (nse.py:418-475): `steps` noise levels x `lsteps`
        updates theta += delta * score + sqrt(2 delta) z, clipped once per level.  One fused launch."""
        self._require_cuda(x)
        opts, noise, seed, off = self._split_kwargs(kwargs)
        m = self.zeros.shape[0]
        grid = engine.time_grid(steps)
        gamma = self._gammas(grid)
        delta = tau * (1 - gamma) / gamma.sqrt()  # nse.py:461
        single = x.ndim == 1 or x.shape[0] == 1  # nse.py:465
        cfg = {} if (opts["clf_free_guidance"] and not single) else None
        prior = engine.PriorSpec(abi.PRIOR_NONE) if (single or cfg is not None) else self._geffner_prior(prior_score_fn, m)
        rep = lambda v: v.repeat_interleave(lsteps)  # noqa: E731
        clip = torch.zeros(steps, lsteps, dtype=torch.float64)
        clip[:, -1] = 1.0
        coef = dict(c_theta=torch.ones(steps * lsteps, dtype=torch.float64), c_score=rep(delta),
                    bstd=rep((2 * delta).sqrt()), clip=clip.reshape(-1))
        return self._run_rows(shape, x, rep(grid[:-1]), coef, prior, False, opts, theta_clipping_range, noise, seed, off, cfg)

    def deterministic_gaussian_geffner(self, shape: Size, x: Tensor, prior_score_fn, steps: int = 1000,
                                       verbose: bool = False, theta_clipping_range=(None, None), **kwargs) -> Tensor:
        """Gaussian transition sampler of Geffner et al. (Algorithm 2) with the continuous schedule (nse.py:658-705):
        theta' = c_theta theta + var_t [S1 / sqrt(gamma) + (1 - n) prior_score] + sqrt(var_t) z."""
        self._require_cuda(x)
        opts, noise, seed, off = self._split_kwargs(kwargs)
        m = self.zeros.shape[0]
        n = x.shape[0] if x.ndim > 1 else 1
        grid = engine.time_grid(steps)
        g = self._gammas(grid)
        den = n - g * (n - 1)
        var = (1 - g) / den  # nse.py:692
        coef = dict(c_theta=(n / g.sqrt() - (n - 1) * g.sqrt()) / den, c_score=var, s1w=1 / g.sqrt(), bstd=var.sqrt())
        prior = self._geffner_prior(prior_score_fn, m)
        x2 = x if x.ndim > 1 else x[None]
        return self._run_rows(shape, x2, grid[:-1], coef, prior, False, opts, theta_clipping_range, noise, seed, off)

    def predictor_corrector(self, shape: Size, x: Tensor, steps: int = 1000, verbose: bool = False,
                            predictor_type: str = "ddim", corrector_type: str = "langevin",
                            theta_clipping_range=(None, None), **kwargs) -> Tensor:
        """Predictor-corrector sampler (nse.py:340-416): DDIM or identity predictor at t, `lsteps` tamed-ULA corrector
        steps (nse.py:301-338) or identity at t - dt.  kwargs as in the reference: `prior_score_fun` (F-NPSE score,
        corrector 'langevin') or the `factorized_score` kwargs (corrector 'id'), `eta`, `lsteps`, `r`, `n`."""
        self._require_cuda(x)
        if predictor_type not in ("ddim", "id") or corrector_type not in ("langevin", "id"):
            raise NotImplementedError("predictor_type in {'ddim','id'}, corrector_type in {'langevin','id'}")
        kw = dict(kwargs)
        eta = float(kw.pop("eta", 1.0))
        lsteps, r = int(kw.pop("lsteps", 0)), float(kw.pop("r", 0.0))
        prior_score_fun = kw.pop("prior_score_fun", None)
        opts, noise, seed, off = self._split_kwargs(kw)
        N, m = int(shape[0]), self.zeros.shape[0]
        single = self._is_single_context(shape, x)
        if corrector_type == "langevin" and lsteps < 1:
            raise TypeError("corrector_type='langevin' needs lsteps and r")
        grid = engine.time_grid(steps)
        t64 = grid[:-1].to(torch.float64)
        dt = 1.0 / steps
        per = (1 if predictor_type == "ddim" else 0) + (lsteps if corrector_type == "langevin" else 0)
        if per == 0:
            raise NotImplementedError("identity predictor with identity corrector does nothing")
        T = torch.zeros(steps, per, dtype=torch.float64)
        cols = {k: torch.zeros(steps, per, dtype=torch.float64) for k in ("c_theta", "c_score", "bstd", "tame", "clip")}
        j = 0
        if predictor_type == "ddim":
            rows = engine.schedule_rows(self, grid[:-1], t64 - dt, eta, 1, [0] * steps)
            T[:, 0] = t64
            cols["c_theta"][:, 0] = rows[:, abi.S_C_THETA]
            cols["c_score"][:, 0] = rows[:, abi.S_C_SCORE]
            cols["bstd"][:, 0] = rows[:, abi.S_BRIDGE_STD]
            j = 1
        if corrector_type == "langevin":
            tc = t64 - dt  # nse.py:413: corrector at t - dt
            eps = r * engine._scalar64(self.alpha, tc).sqrt() * engine._scalar64(self.sigma, tc) ** 2  # nse.py:335
            T[:, j:] = tc[:, None]
            cols["c_theta"][:, j:] = 1.0
            cols["tame"][:, j:] = eps[:, None]
            cols["bstd"][:, j:] = (2 * eps).sqrt()[:, None]
        cols["clip"][:, -1] = 1.0
        coef = {k: v.reshape(-1) for k, v in cols.items()}
        if single:
            prior, weighted = engine.PriorSpec(abi.PRIOR_NONE), False
            paired = x.ndim > 1 and x.shape[0] == N and N > 1
            if paired:
                raise NotImplementedError("predictor_corrector with paired contexts")
        elif corrector_type == "langevin":
            prior, weighted = self._geffner_prior(prior_score_fun, m), False  # nse.py:381
        else:
            prior, weighted = engine.prior_spec(opts["prior_type"], opts["prior"], opts["prior_score_fn"], m), True
        return self._run_rows(shape, x if x.ndim > 1 else x[None], T.reshape(-1), coef, prior, weighted, opts,
                              theta_clipping_range, noise, seed, off)

    def langevin_corrector(self, theta: Tensor, x: Tensor, t: Tensor, score_fun, lsteps: int, r: float, **kwargs) -> Tensor:
        """`lsteps` tamed-ULA steps with a caller-supplied score function (nse.py:301-338); update in kernel 4."""
        self._require_cuda(theta)
        import ctypes as C
        lib = abi.require_cuda()
        tt = torch.as_tensor(t).detach().to("cpu", torch.float64).reshape(1)
        eps = float(r * engine._scalar64(self.alpha, tt).sqrt() * engine._scalar64(self.sigma, tt) ** 2)
        rows = engine.schedule_rows(self, tt, None, 0.0, 1, [abi.COMBINE_SUM],
                                    dict(c_theta=[1.0], tame=[eps], bstd=[math.sqrt(2 * eps)], clip=[0.0]))
        sched = rows.to(torch.float32).to(theta.device)
        out = theta.detach().to(torch.float32).contiguous().clone()
        seed = self._draw_seed()
        for k in range(lsteps):
            g = score_fun(out, x, t).detach().to(torch.float32).contiguous()
            with torch.cuda.device(theta.device):
                abi.check(lib.d4s_ddim_update(engine._ptr(out), engine._ptr(g), out.shape[0], out.shape[1],
                                              engine._ptr(sched), None, C.c_uint64(seed), k, 0, 1.0, -1.0,
                                              engine._stream(theta.device)))
        return out

    # ---- stand-alone prior score (used by d4s.vp_diffused_priors objects) -------------------------------------
    def _prior_score_native(self, theta: Tensor, t: Tensor, spec: "engine.PriorSpec", want_jac: bool = False):
        return engine.prior_score_native(self, theta, t, spec, want_jac)


class NSELoss(nn.Module):
    """Noise-parametrised denoising score matching loss (nse.py:708-751): t ~ U(0, 1), eps ~ N(0, I),
    theta_t = sqrt(alpha(t)) theta + sigma(t) eps, loss = mean ||estimator(theta_t, x, t) - eps||^2.  Pure torch
    (training is outside the sampling hot path); kept so that the reference drivers' `from nse import NSE, NSELoss`
    resolves against the drop-in modules."""

    def __init__(self, estimator: NSE):
        super().__init__()
        self.estimator = estimator

    def forward(self, theta: Tensor, x: Tensor, **kwargs) -> Tensor:
        t = torch.rand(theta.shape[0], dtype=theta.dtype, device=theta.device)
        noise = torch.randn_like(theta)
        theta_t = self.estimator.alpha(t).sqrt()[:, None] * theta + self.estimator.sigma(t)[:, None] * noise
        return (self.estimator(theta_t, x, t, **kwargs) - noise).square().mean()
