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
