"""Trajectory-level gates (-m gpu) beyond the short golden trajectories:

  * a 1000-step, N = 640 GAUSS trajectory on the shipped SLCP weights (n_obs = 30: the bench shape, 160 sample groups on
    148 CTAs => the interleaved two-group schedule) under replayed noise, compared on a subset of samples with the fp64
    oracle run on exactly those samples (samples do not interact);
  * `euler_sde_sampler` against trajectories recorded from the unmodified reference (oracle/gen_golden.py euler_case);
  * the distribution-level JAC gate of SURVEY.md section 8(c): complete 400-step JAC trajectories on the shipped SLCP /
    Lotka-Volterra / SIR weights vs sample sets recorded from the reference in fp32 and fp64 -- sliced Wasserstein
    distance to the fp64 set must lie within the spread of the reference's own two precisions.
"""
from functools import partial

import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import d4s_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _t(v, dev):
    return torch.as_tensor(np.asarray(v, np.float32), device=dev)


def _base(name, dev):
    from d4s_helpers import nse_from_oracle, prior_objects
    g = gio.load(name)
    onet = gio.oracle_net(g)
    nse = nse_from_oracle(onet, dev)
    prior, fn, ptype = prior_objects(g, nse, dev)
    return g, onet, nse, prior, fn, ptype


def replay_noise(seed, steps, N, m):
    """Same generator as oracle/gen_golden.py:replay_noise (numpy PCG64)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((N, m)).astype(np.float32), rng.standard_normal((steps, N, m)).astype(np.float32)


def test_long_interleaved_trajectory_vs_fp64_oracle(dev):
    g, onet, nse, prior, fn, ptype = _base("slcp30_n30", dev)
    N, steps, m = 640, 1000, 5
    z0, z = replay_noise(1234, steps, N, m)
    out = nse.ddim((N,), _t(g["x"], dev), steps=steps, eta=1.0, prior=prior, prior_type=ptype, prior_score_fn=fn,
                   dist_cov_est=_t(g["dist_cov_est"], dev), cov_mode="GAUSS",
                   _noise=(torch.as_tensor(z0), torch.as_tensor(z))).cpu().numpy()
    assert np.isfinite(out).all()
    # CTA 0 interleaves groups 0 and 148 (samples 0-3 and 592-595); 300-303 sits in a CTA of its own pair; 636-639 is the
    # ragged last group
    idx = np.array([0, 1, 2, 3, 300, 301, 592, 593, 594, 595, 638, 639])
    ref = orc.ddim(onet, g["x"].astype(np.float64), z0[idx].astype(np.float64), z[:, idx].astype(np.float64), steps, 1.0,
                   gio.oracle_prior(g), g["dist_cov_est"].astype(np.float64), "GAUSS", ptype)
    err = np.abs(out[idx] - ref).max()
    print(f"1000-step interleaved trajectory vs fp64 oracle on {len(idx)} samples: max |d theta| = {err:.2e}")
    assert err < 1e-3, err


@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("euler")])
@pytest.mark.parametrize("mode", ["GAUSS", "JAC"])
def test_euler_sde_sampler_golden(name, mode, dev):
    import d4s
    e = gio.load(name)
    g, onet, nse, prior, fn, ptype = _base(str(e["base"]), dev)
    N, m = int(e["N"]), onet.theta_dim
    z0, z = replay_noise(int(e["seed"]), 999, N, m)
    score_fn = partial(d4s.diffused_tall_posterior_score, prior=prior, prior_type=ptype, prior_score_fn=fn,
                       x_obs=_t(g["x"], dev), nse=nse, dist_cov_est=_t(g["dist_cov_est"], dev), cov_mode=mode)
    out, hist = d4s.euler_sde_sampler(score_fn, N, m, nse.beta, device=dev, _noise=(torch.as_tensor(z0), torch.as_tensor(z)))
    assert len(hist) == 1000
    ref, ref32 = e[f"f64.{mode}.theta"], e[f"f32.{mode}.theta"]
    floor = np.abs(ref32 - ref).max()
    err = np.abs(out.cpu().numpy() - ref).max()
    ck = np.stack([hist[int(k)].numpy() for k in e["checkpoints"]])
    err_h = np.abs(ck - e[f"f64.{mode}.hist"]).max()
    print(f"{name} {mode}: max |d theta| = {err:.2e} (checkpoints {err_h:.2e}; reference fp32 vs fp64: {floor:.2e})")
    assert np.isfinite(out.cpu().numpy()).all()
    if mode == "GAUSS":  # pinned pathwise (the reference's fp32 and fp64 runs agree to ~1e-6 here)
        tol = max(1e-3, 3 * floor)
        assert err < tol and err_h < max(tol, 3 * np.abs(e[f"f32.{mode}.hist"] - e[f"f64.{mode}.hist"]).max()), (err, err_h, floor)
    else:  # JAC: the reference's own two precisions end O(1) apart after 999 steps (floor): reported, bounded loosely
        assert err < 10 * max(floor, 0.1), (err, floor)


def sliced_wasserstein(a, b, n_proj=512, seed=0):
    """Sliced Wasserstein-2 distance between two equally sized sample sets (numpy; the metric of sbibm_results.py:277)."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((a.shape[1], n_proj))
    d /= np.linalg.norm(d, axis=0, keepdims=True)
    pa, pb = np.sort(a @ d, axis=0), np.sort(b @ d, axis=0)
    return float(np.sqrt(((pa - pb) ** 2).mean()))


def trimmed(a, keep):
    """Outlier-robust aggregation (the reference drops the runs above its 0.99 quantile, sbibm_results.py:230-242): here
    the `keep` samples closest to the component-wise median are kept, NaN rows dropped first."""
    a = a[np.isfinite(a).all(1)]
    r = np.linalg.norm(a - np.median(a, axis=0), axis=1)
    return a[np.argsort(r)[:keep]]


@pytest.mark.parametrize("name", [c for c in gio.names() if c.startswith("jacdist")])
@pytest.mark.parametrize("kernel", ["default", "simt"])
def test_jac_distribution_gate(name, kernel, dev):
    from d4s import _cabi as abi
    e = gio.load(name)
    g, onet, nse, prior, fn, ptype = _base(str(e["base"]), dev)
    N, steps, m = int(e["N"]), int(e["steps"]), onet.theta_dim
    z0, z = replay_noise(int(e["seed"]), steps, N, m)
    clip = (None, None) if np.isnan(e["clip"]).any() else tuple(float(c) for c in e["clip"])
    abi.set_option("path", 1 if kernel == "simt" else 0)
    try:
        out = nse.ddim((N,), _t(g["x"], dev), steps=steps, eta=float(e["eta"]), prior=prior, prior_type=ptype,
                       prior_score_fn=fn, cov_mode="JAC", theta_clipping_range=clip,
                       _noise=(torch.as_tensor(z0), torch.as_tensor(z))).cpu().numpy()
    finally:
        abi.set_option("path", 0)
    r32, r64 = e["f32.samples"], e["f64.samples"]
    keep = int(0.99 * min(len(trimmed(x_, N)) for x_ in (out, r32, r64)))
    a, b, c = trimmed(out, keep), trimmed(r32, keep), trimmed(r64, keep)
    spread = sliced_wasserstein(b, c)           # the reference against itself: fp32 vs fp64
    ours = sliced_wasserstein(a, c)
    ours32 = sliced_wasserstein(a, b)
    scale = float(np.std(c, axis=0).mean())
    path = float(np.median(np.abs(out - r64)[np.isfinite(out).all(1) & np.isfinite(r64).all(1)]))
    print(f"{name} [{kernel}]: sW(ours, ref64) = {ours:.4f}, sW(ours, ref32) = {ours32:.4f}, sW(ref32, ref64) = {spread:.4f}; "
          f"posterior std {scale:.3f}; median pathwise |d theta| vs ref64 {path:.2e}; kept {keep} of {N}")
    assert np.isfinite(out).all(1).mean() >= min(np.isfinite(r32).all(1).mean(), np.isfinite(r64).all(1).mean()) - 0.02
    # inside the spread of the reference's own precisions (plus a floor for sets that coincide pathwise)
    assert min(ours, ours32) <= max(1.5 * spread, 0.02 * scale), (ours, ours32, spread, scale)
