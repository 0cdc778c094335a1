"""Observation sharding (-m gpu): SURVEY.md section 8(e), BASELINE config 5 (n_obs = 10^4, theta = 10).

  * 1 GPU: the stress shape through the fused tcgen05 trajectory kernel against the fp64 oracle (replayed noise);
  * 1 GPU: the in-kernel exchange with world = 1 (the kernel publishes to and reads from its OWN receive buffer) must
    reproduce the plain trajectory bit for bit -- this covers the exchange code path on a single-GPU box;
  * >= 2 GPUs: tests/gpu_multi.py under torchrun (in-kernel peer-memory exchange, NCCL graph, NCCL eager vs 1 GPU).
The multi-rank kernels wait for one another, so they are never run as several processes on one GPU."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import d4s_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _stress(dev, n, m=10, d=10, hidden=(256, 256, 256), seed=0):
    import d4s
    from d4s_helpers import nse_from_oracle
    rng = np.random.default_rng(seed)
    onet = orc.random_net(rng, m, d, hidden=hidden)
    nse = nse_from_oracle(onet, dev)
    x = rng.standard_normal((n, d))
    a = 0.2 * rng.standard_normal((n, m, m))
    cov = a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)
    mu, pc = torch.zeros(m, device=dev), torch.eye(m, device=dev)
    prior, fn = torch.distributions.MultivariateNormal(mu, pc), d4s.get_vpdiff_gaussian_score(mu, pc, nse)
    kw = dict(prior=prior, prior_type="gaussian", prior_score_fn=fn,
              dist_cov_est=torch.as_tensor(cov, dtype=torch.float32, device=dev), cov_mode="GAUSS")
    return rng, onet, nse, x, cov, kw


def test_stress_shape_vs_oracle(dev):
    """n_obs = 10^4, theta = 10 (BASELINE config 5), 16 samples, 3 DDIM steps under replayed noise vs the fp64 oracle."""
    n, m, Nc, steps = 10000, 10, 16, 3
    rng, onet, nse, x, cov, kw = _stress(dev, n)
    z0, z = rng.standard_normal((Nc, m)), rng.standard_normal((steps, Nc, m))
    noise = (torch.as_tensor(z0, dtype=torch.float32), torch.as_tensor(z, dtype=torch.float32))
    out = nse.ddim((Nc,), torch.as_tensor(x, dtype=torch.float32, device=dev), steps=steps, eta=1.0, _noise=noise, **kw)
    ref = orc.ddim(onet, x, z0, z, steps, 1.0, orc.GaussianPrior(np.zeros(m), np.eye(m)), cov, "GAUSS", "gaussian")
    err = float(np.abs(out.cpu().numpy() - ref).max()) / max(1.0, float(np.abs(ref).max()))
    assert err < 1e-3, err


def _single_rank_group():
    import torch.distributed as dist
    if not dist.is_initialized():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    return dist


@pytest.mark.parametrize("n,N,pkind", [(300, 70, "gaussian"), (30, 1300, "uniform"), (7, 50, "uniform")],
                         ids=["obs_chunks", "paired_groups", "small"])
def test_in_kernel_exchange_world1_is_bit_identical(n, N, pkind, dev):
    import d4s
    from d4s import parallel
    _single_rank_group()
    m = 5
    rng, onet, nse, x, cov, kw = _stress(dev, n, m=m, d=8, seed=3)
    if pkind == "uniform":
        lo, hi = torch.full((m,), -3.0, device=dev), torch.full((m,), 3.0, device=dev)
        kw.update(prior=torch.distributions.Uniform(lo, hi, validate_args=False), prior_type="uniform",
                  prior_score_fn=d4s.get_vpdiff_uniform_score(lo, hi, nse))
    xt = torch.as_tensor(x, dtype=torch.float32, device=dev)
    steps = 12
    ref = nse.ddim((N,), xt, steps=steps, eta=0.8, _seed=11, **kw)
    for rep in range(2):  # twice: the second trajectory reuses the buffers with an advanced sequence number
        out = parallel.ddim_obs_sharded(nse, (N,), xt, steps=steps, eta=0.8, _seed=11, exchange="p2p", **kw)
        torch.cuda.synchronize()
        assert torch.equal(out, ref), float((out - ref).abs().max())


def test_multi_gpu_obs_and_sample_sharding():
    """tests/gpu_multi.py under torchrun when the box has >= 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run(["timeout", "900", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(here, "gpu_multi.py")],
                       capture_output=True, text=True, timeout=960)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout
