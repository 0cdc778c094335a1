"""Loading the reference's pickled score networks (`torch.save(score_network, ...)`,
sbibm_posterior_estimation.py:136-139) into the d4s classes.

The pickles name `nse.NSE` / `pf_nse.PF_NSE`, `zuko.nn.MLP`, `zuko.nn.Linear`, torch modules and (for PF models) a
`functools.partial` of `tall_posterior_sampler.tweedies_approximation`.  A custom unpickler maps those names onto this
package, so neither the reference checkout nor zuko has to be importable; attributes missing from old pickles are
patched like the drivers do (sbibm_posterior_estimation.py:601-602).
"""
from __future__ import annotations

import pickle

import torch

from . import nn as d4nn
from . import nse as d4nse
from . import pf_nse as d4pf
from . import tall_posterior_sampler as d4tps

_CLASS_MAP = {
    ("nse", "NSE"): d4nse.NSE,
    ("pf_nse", "PF_NSE"): d4pf.PF_NSE,
    ("zuko.nn", "MLP"): d4nn.MLP,
    ("zuko.nn", "Linear"): d4nn.Linear,
    ("tall_posterior_sampler", "tweedies_approximation"): d4tps.tweedies_approximation,
    ("tall_posterior_sampler", "tweedies_approximation_prior"): d4tps.tweedies_approximation_prior,
    ("tall_posterior_sampler", "prec_matrix_backward"): d4tps.prec_matrix_backward,
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        hit = _CLASS_MAP.get((module, name))
        if hit is not None:
            return hit
        return super().find_class(module, name)


class _PickleModule:
    """The `pickle_module` interface torch.load expects."""
    __name__ = "d4s.io"
    Unpickler = _Unpickler
    load = staticmethod(pickle.load)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    loads = staticmethod(pickle.loads)


def load_score_network(path: str, device="cpu"):
    """Returns a `d4s.NSE` / `d4s.PF_NSE` holding the weights of a reference `score_network.pkl`."""
    net = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PickleModule)
    if not isinstance(net, d4nse.NSE):
        raise TypeError(f"{path} does not hold an NSE / PF_NSE module (got {type(net).__name__})")
    net.__dict__.pop("tweedies_approximator", None)  # stored instance attribute of the reference; ours is a property
    if not hasattr(net, "net_type"):
        net.net_type = "default"
    if not hasattr(net, "n_max"):
        net.n_max = 1
    if not hasattr(net, "theta_emb_dim"):
        net.theta_emb_dim, net.x_emb_dim = net.get_theta_x_embedding_dim(net.zeros.shape[0], _x_dim(net))
    return net.to(device).eval()


def _x_dim(net) -> int:
    first = next(m for m in net.net.modules() if isinstance(m, torch.nn.Linear))
    emb = net.embedding_nn_x
    if isinstance(emb, torch.nn.Identity):
        extra = 2 * net.freqs.numel() + net.zeros.shape[0] + (net.n_max if isinstance(net, d4pf.PF_NSE) else 0)
        return first.in_features - extra
    return next(m for m in emb.modules() if isinstance(m, torch.nn.Linear)).in_features


# ---- normalisation / log-space wrappers of the reference drivers -----------------------------------------------------
# The score networks are trained on z-scored (theta, x) (log-transformed first for Lotka-Volterra / SIR), so every driver
# wraps the sampler in the same three steps (sbibm_posterior_estimation.py:176-214, 385-388;
# jrnnm_posterior_estimation.py:136-160): normalise the observations and the prior, sample in the normalised space,
# map the samples back.
def normalize_context(context: torch.Tensor, x_train_mean, x_train_std, x_log_space: bool = False) -> torch.Tensor:
    """(log) z-scoring of the observations; NaN / inf from zero-variance summary features become 0
    (sbibm_posterior_estimation.py:179-184)."""
    if x_log_space:
        context = torch.log(context)
    out = (context - x_train_mean) / x_train_std
    return torch.nan_to_num(out, nan=0.0, posinf=0.0, neginf=0.0)


def normalize_prior(prior, prior_type: str, theta_train_mean, theta_train_std, score_network, device,
                    theta_log_space: bool = False):
    """(prior_norm, prior_score_fn_norm) in the normalised theta space (sbibm_posterior_estimation.py:186-212).
    Uniform bounds are mapped with the drivers' factor 2 (`(low - mean) / std * 2`); a LogNormal prior with
    `theta_log_space` is the Gaussian of its base distribution."""
    from .vp_diffused_priors import get_vpdiff_gaussian_score, get_vpdiff_uniform_score
    if prior_type == "uniform":
        base = getattr(prior, "base_dist", prior)  # BoxUniform = Independent(Uniform)
        low = ((base.low - theta_train_mean) / theta_train_std * 2).to(device)
        high = ((base.high - theta_train_mean) / theta_train_std * 2).to(device)
        return (torch.distributions.Uniform(low, high, validate_args=False),
                get_vpdiff_uniform_score(low, high, score_network))
    if prior_type == "gaussian":
        if theta_log_space:
            loc, cov = prior.base_dist.loc, torch.diag_embed(prior.base_dist.scale.square())
        else:
            loc, cov = prior.loc, prior.covariance_matrix
        inv_std = torch.diag(1 / theta_train_std)
        loc_n = ((loc - theta_train_mean) / theta_train_std).to(device)
        cov_n = (inv_std @ cov @ inv_std).to(device)
        return (torch.distributions.MultivariateNormal(loc_n, cov_n, validate_args=False),
                get_vpdiff_gaussian_score(loc_n, cov_n, score_network))
    raise NotImplementedError(f"prior_type={prior_type!r}")


def unnormalize_samples(samples: torch.Tensor, theta_train_mean, theta_train_std, theta_log_space: bool = False) -> torch.Tensor:
    """Back to the original parameter space (sbibm_posterior_estimation.py:385-388)."""
    out = samples.detach().cpu() * theta_train_std + theta_train_mean
    return torch.exp(out) if theta_log_space else out


def sample_tall_posterior(score_network, context, nsamples, steps, prior, prior_type, cov_mode, theta_train_mean,
                          theta_train_std, x_train_mean, x_train_std, sampler_type: str = "ddim", clip: bool = False,
                          theta_log_space: bool = False, x_log_space: bool = False, clf_free_guidance: bool = False,
                          device="cuda:0") -> torch.Tensor:
    """The sampling body of the reference's `run_sample_sgm` (sbibm_posterior_estimation.py:144-388) on the fused kernels:
    normalise, estimate the GAUSS covariances with the single-observation pre-pass (one launch), run `ddim` (or the
    Euler-Maruyama sampler on `diffused_tall_posterior_score`), un-normalise.  Returns samples[nsamples, m] on the host
    in the ORIGINAL parameter space."""
    from functools import partial
    from .tall_posterior_sampler import diffused_tall_posterior_score, euler_sde_sampler
    dev = torch.device(device)
    net = score_network.to(dev)
    ctx = normalize_context(context, x_train_mean, x_train_std, x_log_space).to(dev, torch.float32)
    prior_n, fn_n = normalize_prior(prior, prior_type, theta_train_mean, theta_train_std, net, dev, theta_log_space)
    rng = (-3, 3) if clip else (None, None)
    cov_est = cov_prior = None
    if cov_mode == "GAUSS":  # sbibm_posterior_estimation.py:304-322
        cov_est = net.estimate_dist_cov(ctx, nsamples=nsamples, steps=100, eta=0.5)
        if clf_free_guidance:
            cov_prior = net.estimate_dist_cov(torch.zeros_like(ctx[:1]), nsamples=nsamples, steps=100, eta=0.5)
    if sampler_type == "ddim":
        eta = 1 if steps == 1000 else 0.8 if steps == 400 else 0.5  # sbibm_posterior_estimation.py:330-334
        samples = net.ddim(shape=(nsamples,), x=ctx, eta=eta, steps=steps, theta_clipping_range=rng, prior=prior_n,
                           prior_type=prior_type, prior_score_fn=fn_n, clf_free_guidance=clf_free_guidance,
                           dist_cov_est=cov_est, dist_cov_est_prior=cov_prior, cov_mode=cov_mode)
    else:
        score_fn = partial(diffused_tall_posterior_score, prior=prior_n, prior_type=prior_type, prior_score_fn=fn_n,
                           x_obs=ctx, nse=net, dist_cov_est=cov_est, cov_mode=cov_mode)
        samples, _ = euler_sde_sampler(score_fn, nsamples, dim_theta=theta_train_mean.shape[-1], beta=net.beta,
                                       device=str(dev), theta_clipping_range=rng)
    if bool(torch.isnan(samples).any()):
        raise FloatingPointError(f"NaN in samples: {int(torch.isnan(samples).sum())}")
    return unnormalize_samples(samples, theta_train_mean, theta_train_std, theta_log_space)
