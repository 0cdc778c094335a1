"""Runs the UNMODIFIED reference (staged under oracle/_ref by oracle/make_ref.sh) on a synthetic bench workload.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by `bench.py --impl reference`, by bench.py's `cpu_baseline` leg and by
tests; never by the product path.  The reference modules are loaded under their own names (`nse`, `pf_nse`,
`tall_posterior_sampler`, `vp_diffused_priors`) from oracle/_ref with the zuko / sbi stand-ins beside them; nothing of
this repository's kernels or host code is on that path.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "nse.py"))


def _modules():
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `sh oracle/make_ref.sh` where the reference checkout exists")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import nse as ref_nse  # noqa: E402
    import pf_nse as ref_pf  # noqa: E402
    import vp_diffused_priors as ref_vp  # noqa: E402
    for mod in (ref_nse, ref_pf, ref_vp):
        assert os.path.dirname(os.path.abspath(mod.__file__)) == REF_DIR, mod.__file__
    return ref_nse, ref_pf, ref_vp


def build(w: dict, device: str = "cpu"):
    """Reference `NSE` / `PF_NSE` of the workload's architecture with the workload's weights, plus prior objects."""
    import torch
    ref_nse, ref_pf, ref_vp = _modules()
    from zuko.nn import MLP  # the stand-in staged next to the reference (oracle/_ref/zuko)

    def load(mlp, ws, bs):
        lin = [mod for mod in mlp if isinstance(mod, torch.nn.Linear)]
        assert len(lin) == len(ws)
        with torch.no_grad():
            for l, wt, b in zip(lin, ws, bs):
                l.weight.copy_(torch.as_tensor(wt))
                l.bias.copy_(torch.as_tensor(b))

    kw = dict(freqs=w["freqs"], hidden_features=list(w["hidden"]))
    if w["emb"]:
        e = w["emb"]
        kw["embedding_nn_theta"] = MLP(w["m"], e[-1], hidden_features=list(e[:-1]))
        kw["embedding_nn_x"] = MLP(w["d"], e[-1], hidden_features=list(e[:-1]))
        load(kw["embedding_nn_theta"], *w["emb_theta"])
        load(kw["embedding_nn_x"], *w["emb_x"])
    net = (ref_pf.PF_NSE(w["m"], w["d"], w["n_max"], **kw) if w["n_max"] > 1 else ref_nse.NSE(w["m"], w["d"], **kw))
    load(net.net, w["weights"], w["biases"])
    net = net.to(device).eval()
    dev = torch.device(device)
    if w["prior_kind"] == "uniform":
        lo, hi = torch.as_tensor(w["prior"]["low"]).to(dev), torch.as_tensor(w["prior"]["high"]).to(dev)
        prior = torch.distributions.Independent(torch.distributions.Uniform(lo, hi, validate_args=False), 1)
        fn = ref_vp.get_vpdiff_uniform_score(lo, hi, net)
    else:
        mu, pc = torch.as_tensor(w["prior"]["mean"]).to(dev), torch.as_tensor(w["prior"]["cov"]).to(dev)
        prior = torch.distributions.MultivariateNormal(mu, pc, validate_args=False)
        fn = ref_vp.get_vpdiff_gaussian_score(mu, pc, net)
    return net, prior, fn


def ddim_seconds(net, prior, fn, w: dict, n_samples: int, ddim_steps: int, cov_mode: str, device: str = "cpu") -> float:
    """Wall time of ONE complete `NSE.ddim` call of the reference (nse.py:223-267) with `ddim_steps` steps."""
    import torch
    dev = torch.device(device)
    x = torch.as_tensor(w["x"]).to(dev)
    cov = torch.as_tensor(w["cov"]).to(dev)
    eta = w["eta"]
    kwargs = dict(prior=prior, prior_type=w["prior_kind"], prior_score_fn=fn, dist_cov_est=cov, cov_mode=cov_mode)
    if w["n_max"] > 1:  # PF_NSE: subsets of n_max observations (pf_nse.py:199-233 builds them itself)
        kwargs.pop("dist_cov_est")
        kwargs["dist_cov_est"] = cov
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    out = net.ddim((n_samples,), x, steps=ddim_steps, eta=eta, **kwargs)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    assert out.shape == (n_samples, w["m"]) and bool(torch.isfinite(out).all()), "reference produced non-finite samples"
    return dt
