"""Minimal stand-in for `zuko==1.1.0` (absent from this image, no network).

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference modules under
/root/reference (nse.py, pf_nse.py, ...) can be imported in the build container to
generate golden vectors (oracle/gen_golden.py).  Written from zuko's public API, not from
its source: only `nn.MLP`/`nn.Linear` (Linear+ReLU stack), `utils.broadcast` and
`distributions.DiagNormal` carry arithmetic on the sampling path.
"""
from . import distributions, nn, transforms, utils  # noqa: F401
