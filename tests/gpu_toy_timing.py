"""GPU timing helper (not a pytest): BASELINE config 1, the Gaussian-Gaussian toy of gen_gaussian_gaussian.py
(`nse_toy.NSE` + `FakeFNet(analytic eps, EpsilonNet, eps)`), timed the way the reference times it
(gen_gaussian_gaussian.py:141-267: 1000 samples, wall clock around the sampler, GAUSS including the 100-step covariance
pre-pass with 1000 samples per observation) next to the published numbers of BASELINE.md section 1.
usage: python tests/gpu_toy_timing.py [dim=10] [n_obs=32] [eps=1e-2]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from d4s import nse_toy

DIM = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n_obs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
eps_max = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-2
N = 1000
dev = torch.device("cuda:0")
torch.manual_seed(0)
# task of tasks/toy_examples/data_generators.py:Gaussian_Gaussian_mD with the script's random means / stds
means, stds = torch.rand(DIM) * 20 - 10, torch.rand(DIM) * 25 + 0.1
prior_loc, prior_cov = means, torch.diag(stds ** 2)
rho = 0.8
lik_cov = (1 - rho) * torch.eye(DIM) + rho * torch.ones(DIM, DIM)  # equicorrelated simulator noise
theta_true = prior_loc + stds * torch.randn(DIM)
x_obs = (theta_true + torch.distributions.MultivariateNormal(torch.zeros(DIM), lik_cov).sample((n_obs,))).to(dev)

nse = nse_toy.NSE(theta_dim=DIM, x_dim=DIM)
eps_net = nse_toy.EpsilonNet(DIM)
real = nse_toy.AnalyticGaussianEps(prior_loc.numpy(), prior_cov.numpy(), lik_cov.numpy(), nse)
nse.net = nse_toy.FakeFNet(real_eps_fun=real, eps_net=eps_net, eps_net_max=eps_max)
nse = nse.to(dev).eval()
prior_fn = nse_toy.gaussian_prior_triple(prior_loc.to(dev), prior_cov.to(dev), nse)


def cov_prepass():
    """gen_gaussian_gaussian.py:143-160: 1000 single-observation posterior samples per observation, 100 steps."""
    xp = x_obs[:, None].repeat(1, N, 1).reshape(n_obs * N, -1)
    s = nse.ddim((n_obs * N,), xp, steps=100, eta=0.5).reshape(n_obs, N, DIM)
    c = s - s.mean(1, keepdim=True)
    return c.transpose(1, 2) @ c / (N - 1)


def timed(f, reps=3):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), out


t_cov, cov = timed(cov_prepass)
published = {"GAUSS": {1000: 4.77, 400: 2.04, 150: 0.90, 50: 0.45}, "JAC": {1000: 8.18, 400: 3.26, 150: 1.22, 50: 0.41}}
rows = []
for mode in ("GAUSS", "JAC"):
    for steps in (1000, 400):
        eta = 1.0 if steps == 1000 else 0.8
        t, th = timed(lambda: nse.ddim((N,), x_obs, steps=steps, eta=eta, prior_score_fn=prior_fn, dist_cov_est=cov,
                                       cov_mode=mode))
        total = t + (t_cov if mode == "GAUSS" else 0.0)
        pub = published[mode][steps] if (DIM, n_obs) == (10, 32) else None
        rows.append({"mode": mode, "steps": steps, "sampler_s": round(t, 4), "cov_prepass_s": round(t_cov, 4) if mode == "GAUSS" else 0.0,
                     "total_s": round(total, 4), "samples_per_s": round(N / total, 1), "published_s": pub,
                     "vs_published": None if pub is None else round(pub / total, 1), "finite": bool(torch.isfinite(th).all())})
# the Langevin (F-NPSE, 5 inner steps, tau 0.5) and deterministic Geffner baselines of the same script (:205-238)
published_l = {1000: 16.65, 400: 6.65}
for steps in (1000, 400):
    for mode, f in (("Langevin", lambda: nse.annealed_langevin_geffner((N,), x_obs, prior_fn, lsteps=5, tau=0.5, steps=steps)),
                    ("DET_GEF", lambda: nse.deterministic_gaussian_geffner((N,), x_obs, prior_fn, steps=steps))):
        t, th = timed(f)
        pub = published_l[steps] if (mode == "Langevin" and (DIM, n_obs) == (10, 32)) else None
        rows.append({"mode": mode, "steps": steps, "sampler_s": round(t, 4), "cov_prepass_s": 0.0, "total_s": round(t, 4),
                     "samples_per_s": round(N / t, 1), "published_s": pub,
                     "vs_published": None if pub is None else round(pub / t, 1), "finite": bool(torch.isfinite(th).all())})
print(json.dumps({"workload": f"toy Gaussian-Gaussian dim={DIM}, n_obs={n_obs}, eps={eps_max}, 1000 samples; "
                              "published = BASELINE.md section 1 (unspecified CUDA GPU, torch eager)", "rows": rows}))
