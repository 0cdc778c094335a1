"""Builds d4s modules from golden fixtures / oracle nets (test helper)."""
import numpy as np
import torch

import d4s
from d4s import nn as d4nn


def _load_mlp(mlp: torch.nn.Sequential, weights, biases):
    lin = [mod for mod in mlp if isinstance(mod, torch.nn.Linear)]
    assert len(lin) == len(weights)
    with torch.no_grad():
        for l, w, b in zip(lin, weights, biases):
            assert tuple(l.weight.shape) == tuple(w.shape), (l.weight.shape, w.shape)
            l.weight.copy_(torch.as_tensor(np.asarray(w, np.float32)))
            l.bias.copy_(torch.as_tensor(np.asarray(b, np.float32)))


def _embed_from(omlp):
    if omlp is None:
        return torch.nn.Identity()
    widths = [omlp.weights[0].shape[1]] + [w.shape[0] for w in omlp.weights]
    mod = d4nn.MLP(widths[0], widths[-1], hidden_features=widths[1:-1])
    _load_mlp(mod, omlp.weights, omlp.biases)
    return mod


def nse_from_oracle(onet, device="cpu"):
    """d4s.NSE / d4s.PF_NSE with the weights of an oracle net."""
    hidden = [w.shape[0] for w in onet.net.weights[:-1]]
    F = len(onet.freqs)
    kw = dict(freqs=F, hidden_features=hidden, embedding_nn_theta=_embed_from(onet.emb_theta),
              embedding_nn_x=_embed_from(onet.emb_x))
    if onet.pf:
        mod = d4s.PF_NSE(onet.theta_dim, onet.x_dim, onet.n_max, **kw)
    else:
        mod = d4s.NSE(onet.theta_dim, onet.x_dim, **kw)
    _load_mlp(mod.net, onet.net.weights, onet.net.biases)
    with torch.no_grad():
        mod.freqs.copy_(torch.as_tensor(np.asarray(onet.freqs, np.float32)))
    return mod.to(device).eval()


def prior_objects(g, nse, device):
    """(prior distribution, d4s prior score object, prior_type) like the reference drivers build them."""
    kind = str(g["prior_kind"])
    if kind == "uniform":
        low = torch.as_tensor(g["prior.low"], device=device)
        high = torch.as_tensor(g["prior.high"], device=device)
        return (torch.distributions.Uniform(low, high, validate_args=False),
                d4s.get_vpdiff_uniform_score(low, high, nse), "uniform")
    loc = torch.as_tensor(g["prior.loc"], device=device)
    cov = torch.as_tensor(g["prior.cov"], device=device)
    return (torch.distributions.MultivariateNormal(loc, cov), d4s.get_vpdiff_gaussian_score(loc, cov, nse), "gaussian")


def toy_from_golden(g, device="cpu"):
    """(d4s.nse_toy.NSE with FakeFNet(AnalyticGaussianEps, EpsilonNet, eps_max), prior triple) of a toy fixture."""
    from d4s import nse_toy
    m = int(g["theta_dim"])
    nse = nse_toy.NSE(theta_dim=m, x_dim=g["x"].shape[-1])
    eps_net = nse_toy.EpsilonNet(m)
    _load_mlp(eps_net.net, [g["net.W0"], g["net.W1"]], [g["net.b0"], g["net.b1"]])
    real = nse_toy.AnalyticGaussianEps(g["prior.loc"], g["prior.cov"], g["lik_cov"], nse)
    nse.net = nse_toy.FakeFNet(real_eps_fun=real, eps_net=eps_net, eps_net_max=float(g["eps_max"]))
    nse = nse.to(device).eval()
    prior_fn = nse_toy.gaussian_prior_triple(torch.as_tensor(g["prior.loc"], device=device),
                                             torch.as_tensor(g["prior.cov"], device=device), nse)
    return nse, prior_fn
