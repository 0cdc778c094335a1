"""CPU tests of the host logic (d4s.engine): schedule, layer-1 partials, per-step matrices and the branch
selection of nse.py:627-654.  The device arithmetic of kernels 1-4 is emulated here in numpy from the SAME
tables the kernels read, and compared with the oracle -- so a wrong table is caught without a GPU."""
import numpy as np
import pytest
import torch

import golden_io as gio
from d4s import _cabi as abi
import d4s
from d4s import engine
from d4s_helpers import nse_from_oracle
from oracle import d4s_oracle as orc

CASES = [c for c in gio.names() if not c.startswith(("langevin", "toy", "cfg", "contexts", "euler", "jacdist"))]


def _spec(g):
    if str(g["prior_kind"]) == "uniform":
        return engine.PriorSpec(abi.PRIOR_UNIFORM, low=torch.as_tensor(g["prior.low"], dtype=torch.float64),
                                high=torch.as_tensor(g["prior.high"], dtype=torch.float64))
    return engine.PriorSpec(abi.PRIOR_GAUSSIAN, mean=torch.as_tensor(g["prior.loc"], dtype=torch.float64),
                            cov=torch.as_tensor(g["prior.cov"], dtype=torch.float64))


def emulate_scores(onet, nse, view, theta, x, tf_row, of, inv_sigma):
    """kernel 1: separable layer 0 + hidden layers + output layer on the (sample x obs) grid."""
    feat = theta if onet.emb_theta is None else onet.emb_theta(theta)  # kernel 0: theta-embedding MLP
    A = view.w1[:, :feat.shape[-1]].numpy()  # theta block
    z = (feat @ A.T)[:, None, :] + of[None, :, :view.h1] + tf_row[None, None, :view.h1]
    h = np.maximum(z, 0)
    ws, bs = onet.net.weights, onet.net.biases
    for l in range(1, len(ws) - 1):
        h = np.maximum(h @ ws[l].T + bs[l], 0)
    eps = h @ ws[-1].T + bs[-1]
    return -eps * inv_sigma


def emulate_finalize(row, mats, m, prior, theta, S1, S2, sumP=None):
    """kernel 3: prior term + combine, from one schedule row and one mats row."""
    comb = int(row[abi.S_COMBINE])
    aos2, omn = row[abi.S_A_OVER_S2], row[abi.S_ONE_MINUS_N]
    lmat = mats[:m * m].reshape(m, m)
    G = mats[m * m:2 * m * m].reshape(m, m)
    mu = mats[2 * m * m:]
    N = theta.shape[0]
    g = np.zeros_like(theta)
    lam = np.broadcast_to(lmat, (N, m, m)).copy()
    if sumP is not None:
        lam = lam + sumP
    if prior.kind == abi.PRIOR_GAUSSIAN:
        g = (theta - mu) @ G.T
    elif prior.kind == abi.PRIOR_UNIFORM:
        onet = orc.OracleNet(m, 1, np.zeros(1), None)
        up = orc.UniformPrior(prior.low.numpy(), prior.high.numpy())
        s, sig = row[abi.S_SQRT_ALPHA], row[abi.S_SIGMA]
        ub, ua = (up.high * s - theta) / sig, (up.low * s - theta) / sig
        f = (orc._norm_cdf(ub) - orc._norm_cdf(ua)) / s
        pb, pa = orc._norm_pdf(ub), orc._norm_pdf(ua)
        fp = -(pb - pa) / (sig * s)
        den = f + 1e-6
        s0 = fp / den
        if comb == abi.COMBINE_PER_SAMPLE:
            fpp = -(ub * pb - ua * pa) / (sig ** 2 * s)
            jac = fpp / den - fp ** 2 / den ** 2
            p0 = aos2 / (1 + sig ** 2 * jac)
            lam[:, np.arange(m), np.arange(m)] += omn * p0
            g = omn * p0 * s0
        else:
            g = omn * s0
    if comb == abi.COMBINE_SUM:
        return g + S1
    w = (S2 if sumP is not None else S2 + aos2 * S1) + g
    if comb == abi.COMBINE_SHARED:
        return w @ lmat.T
    return np.linalg.solve(lam, w[..., None])[..., 0]


@pytest.mark.parametrize("name", CASES)
def test_tables_reproduce_factorized_score_gauss(name):
    g = gio.load(name)
    onet, oprior = gio.oracle_net(g), gio.oracle_prior(g)
    nse = nse_from_oracle(onet)
    view = engine.NetView(engine._linears(nse.net))
    spec = _spec(g)
    theta = g["theta"].astype(np.float64)
    x = torch.as_tensor(g["x"])
    ns = torch.as_tensor(g["kw.n"]) if "kw.n" in g else None
    cov = g["dist_cov_est"].astype(np.float64)
    n = x.shape[0]
    m = onet.theta_dim
    Q = np.linalg.inv(cov)
    t = torch.as_tensor(g["t_grid"])
    sa = engine._scalar64(nse.alpha, t.double()).sqrt()
    comb = [engine.combine_mode(spec, False, n, float(s)) for s in sa]
    rows = engine.schedule_rows(nse, t, None, 0.0, n, comb)
    mats = engine.mats_table(rows, spec, False, n, torch.as_tensor(Q.sum(0)), m).numpy()
    tf = engine.time_feat_table(nse, view, t).numpy()
    of = engine.obs_feat_table(nse, view, x, ns).numpy()
    rows = rows.numpy()
    for k in range(len(t)):
        s = emulate_scores(onet, nse, view, theta, x, tf[k], of, rows[k, abi.S_INV_SIGMA])
        assert orc.rel_l2(s, g["f64.scores"][k]) < 1e-6
        S1 = s.sum(1)
        S2 = np.einsum("jab,ijb->ia", Q, s)
        out = emulate_finalize(rows[k], mats[k], m, spec, theta, S1, S2)
        ref = g["f64.GAUSS.factorized_score"][k]
        floor = orc.rel_l2(g["f32.GAUSS.factorized_score"][k], ref)
        assert orc.rel_l2(out, ref) < max(2e-6, 1e-2 * floor), (name, float(t[k]))


@pytest.mark.parametrize("name", [c for c in CASES if "pf" not in c])
def test_tables_reproduce_factorized_score_jac(name):
    g = gio.load(name)
    onet, oprior = gio.oracle_net(g), gio.oracle_prior(g)
    nse = nse_from_oracle(onet)
    spec = _spec(g)
    theta = g["theta"].astype(np.float64)
    x = g["x"].astype(np.float64)
    n, m = x.shape[0], onet.theta_dim
    t = torch.as_tensor(g["t_grid"])
    sa = engine._scalar64(nse.alpha, t.double()).sqrt()
    comb = [engine.combine_mode(spec, True, n, float(s)) for s in sa]
    rows = engine.schedule_rows(nse, t, None, 0.0, n, comb)
    mats = engine.mats_table(rows, spec, True, n, None, m).numpy()
    rows = rows.numpy()
    for k in range(len(t)):
        prec, _, s = orc.tweedies_approximation(onet, x, theta, np.float64(t[k]), None, "JAC")
        V = np.einsum("ijab,ijb->ia", prec, s)
        out = emulate_finalize(rows[k], mats[k], m, spec, theta, s.sum(1), V, sumP=prec.sum(1))
        ref = g["f64.JAC.factorized_score"][k]
        floor = orc.rel_l2(g["f32.JAC.factorized_score"][k], ref)
        assert orc.rel_l2(out, ref) < max(2e-6, 3.0 * floor), (name, float(t[k]))


@pytest.mark.parametrize("steps,eta", [(40, 1.0), (400, 0.8), (25, 0.5), (1000, 1.0)])
def test_ddim_coefficients_match_reference_update(steps, eta):
    """theta' = c_theta theta + c_score score + bridge_std z  ==  nse.py:281-298 (oracle.ddim_step)."""
    onet = orc.random_net(np.random.default_rng(0), 3, 2, hidden=(8,))
    nse = nse_from_oracle(onet)
    grid = engine.time_grid(steps)
    assert np.array_equal(grid.numpy(), orc.time_grid(steps))
    t = grid[:-1]
    rows = engine.schedule_rows(nse, t, t.double() - 1.0 / steps, eta, 5, [0] * steps).numpy()
    rng = np.random.default_rng(1)
    theta, score, z = rng.standard_normal((3, 7, 3))
    for k in [0, 1, steps // 2, steps - 2, steps - 1]:
        ref = orc.ddim_step(onet, theta, score, np.float64(t[k]), 1 / steps, eta, z)
        out = rows[k, abi.S_C_THETA] * theta + rows[k, abi.S_C_SCORE] * score + rows[k, abi.S_BRIDGE_STD] * z
        assert np.abs(out - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), k
    assert rows[-1, abi.S_BRIDGE_STD] == 0.0  # last step is deterministic (alpha(t - dt) = 1)


def test_combine_mode_gate():
    uni = engine.PriorSpec(abi.PRIOR_UNIFORM)
    gau = engine.PriorSpec(abi.PRIOR_GAUSSIAN)
    assert engine.combine_mode(gau, False, 30, 0.1) == abi.COMBINE_SHARED
    assert engine.combine_mode(gau, True, 30, 0.1) == abi.COMBINE_PER_SAMPLE
    assert engine.combine_mode(uni, False, 30, 0.51) == abi.COMBINE_PER_SAMPLE
    assert engine.combine_mode(uni, False, 30, 0.5) == abi.COMBINE_SUM  # strict > (nse.py:643)
    assert engine.combine_mode(uni, True, 1, 0.9) == abi.COMBINE_SUM
    assert engine.combine_mode(engine.PriorSpec(abi.PRIOR_NONE), False, 1, 0.9) == abi.COMBINE_SUM


def test_driver_normalisation_wrappers_cpu():
    """d4s.io.normalize_context / normalize_prior / unnormalize_samples restate sbibm_posterior_estimation.py:179-212,
    385-388; checked against the formulas written out by hand (and against the golden fixtures' normalised priors)."""
    from d4s import io
    g = torch.Generator().manual_seed(0)
    x = torch.rand(7, 4, generator=g) + 0.5
    mean, std = torch.tensor([0.1, 0.2, 0.3, 0.4]), torch.tensor([1.0, 2.0, 0.0, 0.5])  # a zero std -> inf/nan -> 0
    out = io.normalize_context(x, mean, std, x_log_space=True)
    ref = (torch.log(x) - mean) / std
    assert torch.equal(out[:, [0, 1, 3]], ref[:, [0, 1, 3]]) and bool((out[:, 2] == 0).all())
    tm, ts = torch.tensor([0.5, -1.0]), torch.tensor([2.0, 0.25])
    nse = d4s.NSE(2, 4, hidden_features=[64, 64])
    uni = torch.distributions.Independent(torch.distributions.Uniform(torch.tensor([-3.0, -3.0]), torch.tensor([3.0, 3.0])), 1)
    pn, fn = io.normalize_prior(uni, "uniform", tm, ts, nse, "cpu")
    assert torch.allclose(pn.low, (torch.tensor([-3.0, -3.0]) - tm) / ts * 2) and fn.kind == "uniform"
    assert torch.allclose(torch.as_tensor(fn.high), (torch.tensor([3.0, 3.0]) - tm) / ts * 2)
    logn = torch.distributions.LogNormal(torch.tensor([-0.125, -3.0]), torch.tensor([0.5, 0.5]))
    pn, fn = io.normalize_prior(torch.distributions.Independent(logn, 1), "gaussian", tm, ts, nse, "cpu", theta_log_space=True)
    assert torch.allclose(pn.loc, (torch.tensor([-0.125, -3.0]) - tm) / ts)
    assert torch.allclose(pn.covariance_matrix, torch.diag(torch.tensor([0.25, 0.25]) / ts ** 2)) and fn.kind == "gaussian"
    s = torch.randn(5, 2, generator=g)
    assert torch.allclose(io.unnormalize_samples(s, tm, ts, theta_log_space=True), torch.exp(s * ts + tm))
    assert torch.allclose(io.unnormalize_samples(s, tm, ts), s * ts + tm)


def test_schedule_and_matrix_tables_are_cached_by_value():
    """engine._sched_mats_dev: the device copies of the schedule rows / step matrices are reused when everything they depend
    on has the same VALUES (fresh tensors with equal contents: hit), and rebuilt when any of it changes -- a monkey-patched
    alpha, another eta, another prior parameter, another sum of precisions (never a stale hit)."""
    rng = np.random.default_rng(0)
    onet = orc.random_net(rng, 3, 4, hidden=(64, 64))
    nse = nse_from_oracle(onet, "cpu")
    m, n, steps = 3, 7, 20
    grid = engine.time_grid(steps)
    t, t_prev = grid[:-1], grid[:-1].to(torch.float64) - 1.0 / steps
    comb = [abi.COMBINE_SHARED] * steps
    b = rng.standard_normal((m, m))
    cov0 = torch.as_tensor(b @ b.T + np.eye(m))
    sum_q = torch.as_tensor(np.eye(m) * 5.0)

    def spec(scale=1.0):
        return engine.PriorSpec(abi.PRIOR_GAUSSIAN, mean=torch.zeros(m, dtype=torch.float64), cov=scale * cov0.clone())

    def tables(**over):
        kw = dict(eta=1.0, prior=spec(), sum_q=sum_q.clone())
        kw.update(over)
        return engine._sched_mats_dev(nse, t.clone(), t_prev.clone(), kw["eta"], n, list(comb), None, kw["prior"], False,
                                      kw["sum_q"], m, None, "cpu")

    engine.clear_caches()
    s0, m0 = tables()
    s1, m1 = tables()
    assert s1 is s0 and m1 is m0  # equal values in fresh tensors: the same device copies
    rows = engine.schedule_rows(nse, t, t_prev, 1.0, n, comb)
    assert torch.equal(s0, rows.to(torch.float32))
    assert torch.equal(m0, engine.mats_table(rows, spec(), False, n, sum_q, m).to(torch.float32))
    for over in (dict(eta=0.5), dict(prior=spec(1.5)), dict(sum_q=sum_q * 2.0)):
        s2, m2 = tables(**over)
        assert s2 is not s0 and (not torch.equal(s2, s0) or not torch.equal(m2, m0)), over
    old_alpha = nse.alpha
    try:
        nse.alpha = lambda tt: old_alpha(tt) ** 1.1  # callers monkey-patch the schedule
        s3, _ = tables()
        assert s3 is not s0 and not torch.equal(s3, s0)
    finally:
        del nse.alpha
    s4, m4 = tables()
    assert torch.equal(s4, s0) and torch.equal(m4, m0)
