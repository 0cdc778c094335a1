"""Loads tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import glob
import os

import numpy as np

from oracle import d4s_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def _mlp(g, prefix, dtype):
    if f"{prefix}.L" not in g:
        return None
    L = int(g[f"{prefix}.L"])
    return orc.OracleMLP([g[f"{prefix}.W{i}"].astype(dtype) for i in range(L)],
                         [g[f"{prefix}.b{i}"].astype(dtype) for i in range(L)])


def oracle_net(g, dtype=np.float64):
    x_dim = g["x"].shape[-1]
    if "eps_max" in g:  # toy FakeFNet fixture (gen_golden.toy_case)
        pc, lc, mu = (g[k].astype(np.float64) for k in ("prior.cov", "lik_cov", "prior.loc"))
        inv_lik, inv_prior = np.linalg.inv(lc), np.linalg.inv(pc)
        toy = dict(post_cov=np.linalg.inv(inv_lik + inv_prior).astype(dtype), inv_lik=inv_lik.astype(dtype),
                   inv_prior_mu=(inv_prior @ mu).astype(dtype), eps_max=float(g["eps_max"]))
        return orc.OracleNet(int(g["theta_dim"]), x_dim, np.zeros(0, dtype), _mlp(g, "net", dtype), schedule="toy",
                             dtype=dtype, toy=toy)
    return orc.OracleNet(int(g["theta_dim"]), x_dim, g["freqs"].astype(dtype), _mlp(g, "net", dtype),
                         _mlp(g, "embt", dtype), _mlp(g, "embx", dtype), n_max=int(g["n_max"]),
                         pf=bool(int(g["pf"])), dtype=dtype)


def oracle_prior(g):
    if str(g["prior_kind"]) == "cfg":
        return None
    if str(g["prior_kind"]) == "uniform":
        return orc.UniformPrior(g["prior.low"].astype(np.float64), g["prior.high"].astype(np.float64))
    return orc.GaussianPrior(g["prior.loc"].astype(np.float64), g["prior.cov"].astype(np.float64))
