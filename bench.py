#!/usr/bin/env python
"""Benchmark of the tall-data DDIM posterior sampler (BASELINE.json metric: posterior samples/s, DDIM GAUSS/JAC).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--cov-mode GAUSS|JAC] [--n-obs 30]

A bench "step" is ONE complete `NSE.ddim` trajectory (default: BASELINE config 2 = sbibm SLCP shape, theta=5,
x=8, MLP 19->256^3->5, n_obs=30, 1000 posterior samples per GPU, 1000 DDIM steps, eta=1, GAUSS, uniform prior;
synthetic observations and random-init weights).  Rank r of N samples its own 1000 posterior samples
(samples shard with no communication => "scaling": "weak"); value = all ranks' samples / max-over-ranks time.

value        device-resident: tables/inputs in HBM, CUDA events around d4s_ddim_run only.
e2e          NSE.ddim(...) through the public API with HOST inputs (pinned), H2D + table build + D2H inside.
roofline     dominant kernel: the fused tcgen05 trajectory kernel (GAUSS; one launch per trajectory) or the fp32 SIMT
             score kernel (JAC): algorithmic flops per launch / launch duration vs the measured bf16 tensor peak.
cpu_baseline the oracle port (numpy + torch-CPU GEMMs, all host threads) on a bounded sample of DDIM steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

SHAPES = {  # BASELINE.json configs
    "slcp": dict(m=5, d=8, hidden=(256, 256, 256), prior="uniform"),
    "lotka_volterra": dict(m=4, d=20, hidden=(256, 256, 256), prior="gaussian"),
    "sir": dict(m=2, d=10, hidden=(256, 256, 256), prior="gaussian"),
    "stress": dict(m=10, d=10, hidden=(256, 256, 256), prior="gaussian"),
    # config 4: Jansen-Rit nets (theta / x embedding MLPs 64-64-64 -> 32, 32 time frequencies; probed from the shipped
    # weights) and the PF_NSE partial factorisation (n_max = 3) on the SLCP shape
    "jrnnm": dict(m=4, d=33, hidden=(128, 256, 128), prior="uniform", freqs=32, emb=(64, 64, 64, 32)),
    "pf_slcp": dict(m=5, d=8, hidden=(256, 256, 256), prior="uniform", n_max=3),
}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE traj_tc_kernel launch (a whole 1000-step trajectory) of the default
# workload, from the ncu --set full capture summarised in profiles/r02_ncu_full_traj_tc.txt (3.126 MB + 0.129 MB): weights,
# observation features and per-step tables are read once and stay in L2; the kernel is not HBM-bound.  NOT measured in the
# bench run itself (a number printed under a profiler is never a bench value): the roofline record names this source.
TRAJ_TC_DRAM_BYTES = 3125760 + 128768


def is_default_workload(args, w) -> bool:
    return (args.task == "slcp" and args.cov_mode == "GAUSS" and args.n_obs == 30 and w["ddim_steps"] == 1000
            and args.samples == 1000)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--task", default="slcp", choices=sorted(SHAPES))
    ap.add_argument("--cov-mode", default="GAUSS", choices=["GAUSS", "JAC"])
    ap.add_argument("--n-obs", type=int, default=30)
    ap.add_argument("--samples", type=int, default=1000, help="posterior samples per GPU")
    ap.add_argument("--ddim-steps", type=int, default=None, help="default 1000 (GAUSS) / 400 (JAC)")
    ap.add_argument("--cpu-sample-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", default="cpu", help="--impl reference: 'cpu' (default) or 'cuda' (torch eager on the B200)")
    ap.add_argument("--ref-ddim-steps", type=int, default=None,
                    help="--impl reference: DDIM steps of the complete trajectory timed per bench step "
                         "(default 50 GAUSS / 20 JAC on the CPU, the full trajectory on cuda)")
    ap.add_argument("--stress-samples", type=int, default=8192, help="posterior samples of the observation-sharded record")
    ap.add_argument("--stress-steps", type=int, default=100, help="DDIM steps of the observation-sharded record")
    ap.add_argument("--stress-n-obs", type=int, default=10000)
    ap.add_argument("--no-obs-sharded", action="store_true", help="skip the observation-sharded stress record")
    ap.add_argument("--no-sweep", action="store_true", help="skip the n_obs sweep and the JAC sub-record (1 GPU only)")
    ap.add_argument("--jac-simt", action="store_true",
                    help="JAC steps on the fp32 SIMT kernel instead of the default tcgen05 score + Jacobian kernel (fp16x3)")
    ap.add_argument("--jac-tc", action="store_true", help="(accepted for older scripts: the tcgen05 JAC kernel is the default)")
    return ap.parse_args()


def workload(args):
    """Synthetic inputs of the task's shape: seeded, identical on every rank."""
    sh = SHAPES[args.task]
    m, d, n = sh["m"], sh["d"], args.n_obs
    rng = np.random.default_rng(0)
    F, emb, n_max = sh.get("freqs", 3), sh.get("emb"), sh.get("n_max", 1)  # freqs = 3 by default (nse.py:27)

    def init_mlp(widths):  # nn.Linear default init: U(-1/sqrt(in), 1/sqrt(in))
        ws, bs = [], []
        for a_, b_ in zip(widths[:-1], widths[1:]):
            k_ = 1.0 / np.sqrt(a_)
            ws.append(rng.uniform(-k_, k_, size=(b_, a_)).astype(np.float32))
            bs.append(rng.uniform(-k_, k_, size=(b_,)).astype(np.float32))
        return ws, bs

    n_in = (emb[-1] if emb else m) + (emb[-1] if emb else d) + 2 * F + (n_max if n_max > 1 else 0)
    weights, biases = init_mlp([n_in, *sh["hidden"], m])
    emb_theta = init_mlp([m, *emb]) if emb else None
    emb_x = init_mlp([d, *emb]) if emb else None
    x = np.random.default_rng(1).standard_normal((n, d)).astype(np.float32)
    n_units = -(-n // n_max)  # observations, or PF_NSE subsets of n_max observations (pf_nse.py:41-60)
    a = 0.2 * np.random.default_rng(2).standard_normal((n_units, m, m))
    cov = (a @ a.transpose(0, 2, 1) + 0.05 * np.eye(m)).astype(np.float32)
    if sh["prior"] == "uniform":
        prior = dict(low=np.full(m, -3.46, np.float32), high=np.full(m, 3.46, np.float32))
    else:
        prior = dict(mean=np.zeros(m, np.float32), cov=np.eye(m, dtype=np.float32))
    ddim_steps = args.ddim_steps or (1000 if args.cov_mode == "GAUSS" else 400)
    eta = 1.0 if ddim_steps == 1000 else 0.8 if ddim_steps == 400 else 0.5  # sbibm_posterior_estimation.py:330-334
    return dict(weights=weights, biases=biases, x=x, cov=cov, prior=prior, m=m, d=d, n=n, ddim_steps=ddim_steps,
                eta=eta, prior_kind=sh["prior"], hidden=sh["hidden"], freqs=F, emb=emb, emb_theta=emb_theta, emb_x=emb_x,
                n_max=n_max, n_in=n_in, n_units=n_units)


def build_nse(w, dev):
    """d4s.NSE of the task's architecture holding the synthetic weights."""
    import torch
    import d4s
    from d4s import nn as d4nn

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
        kw["embedding_nn_theta"] = d4nn.MLP(w["m"], e[-1], hidden_features=list(e[:-1]))
        kw["embedding_nn_x"] = d4nn.MLP(w["d"], e[-1], hidden_features=list(e[:-1]))
        load(kw["embedding_nn_theta"], *w["emb_theta"])
        load(kw["embedding_nn_x"], *w["emb_x"])
    nse = d4s.PF_NSE(w["m"], w["d"], w["n_max"], **kw) if w["n_max"] > 1 else d4s.NSE(w["m"], w["d"], **kw)
    load(nse.net, w["weights"], w["biases"])
    return nse.to(dev).eval()


def oracle_objects(w):
    """The same net / prior as oracle objects -- used ONLY by the CPU-baseline legs."""
    from oracle import d4s_oracle as orc
    mlp = lambda wb: None if wb is None else orc.OracleMLP([a.copy() for a in wb[0]], [b.copy() for b in wb[1]])
    onet = orc.OracleNet(w["m"], w["d"], (np.arange(1, w["freqs"] + 1) * np.pi).astype(np.float32),
                         mlp((w["weights"], w["biases"])), mlp(w["emb_theta"]), mlp(w["emb_x"]), n_max=w["n_max"],
                         pf=w["n_max"] > 1, dtype=np.float32)
    if w["prior_kind"] == "uniform":
        oprior = orc.UniformPrior(w["prior"]["low"].astype(np.float64), w["prior"]["high"].astype(np.float64))
    else:
        oprior = orc.GaussianPrior(w["prior"]["mean"].astype(np.float64), w["prior"]["cov"].astype(np.float64))
    return orc, onet, oprior


def flops_per_pair(w):
    """Score MLP on one (sample, observation unit) pair; the embedding nets run once per sample / observation."""
    dims = [w["n_in"], *w["hidden"], w["m"]]
    return 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))


def config_dict(args, w, world):
    return {"workload": f"sbibm {args.task} shape: theta={w['m']}, x={w['d']}, MLP {w['n_in']}->"
                        f"{'x'.join(map(str, w['hidden']))}->{w['m']}, n_obs={w['n']}, {args.samples} posterior "
                        f"samples per GPU, DDIM {w['ddim_steps']} steps eta={w['eta']}, {args.cov_mode}, "
                        f"{w['prior_kind']} prior; one bench step = one full trajectory",
            "n_obs": w["n"], "samples_per_gpu": args.samples, "ddim_steps": w["ddim_steps"], "cov_mode": args.cov_mode,
            "jac_kernel": ("fp32 SIMT (--jac-simt)" if getattr(args, "jac_simt", False) else "tcgen05 fp16x3") if args.cov_mode == "JAC" else None,
            "parallelism": f"samples sharded over {world} GPU(s), no collective",
            "l2": "256 MiB buffer written between timed trajectories (L2 flush)"}


# ---------------------------------------------------------------------------------------------------------
# CPU baseline = oracle port on the host cores (bounded sample of DDIM steps)
# ---------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(args, w, n_ddim_steps, repeats=1):
    """posterior samples/s of the oracle port, extrapolated from `n_ddim_steps` DDIM steps spread over the
    trajectory (the cost per step is t-independent except for the alpha-gate of nse.py:643)."""
    import torch
    orc, onet, oprior = oracle_objects(w)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc.GEMM_BACKEND = "torch"
    N = args.samples
    rng = np.random.default_rng(3)
    theta = rng.standard_normal((N, w["m"])).astype(np.float32)
    grid = orc.time_grid(w["ddim_steps"])[:-1]
    idx = np.linspace(0, len(grid) - 1, n_ddim_steps).round().astype(int)
    z = rng.standard_normal((N, w["m"])).astype(np.float32)

    x_or, ns = (orc.pf_create_data(onet, w["x"]) if w["n_max"] > 1 else (w["x"], None))

    def one(tv):
        s = orc.factorized_score(onet, theta, x_or, np.float32(tv), oprior, w["cov"], args.cov_mode, w["prior_kind"], ns)
        return orc.ddim_step(onet, theta, s, np.float32(tv), 1 / w["ddim_steps"], w["eta"], z)

    one(grid[idx[0]])  # warm-up
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for k in idx:
            one(grid[k])
        dt = (time.perf_counter() - t0) / len(idx)
        best = dt if best is None else min(best, dt)
    rate = N / (best * w["ddim_steps"])
    return rate, cores, best


def reference_rate(args, w, ref_steps, device="cpu", repeats=1):
    """posterior samples/s of the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.sh): ONE complete
    `NSE.ddim` call with `ref_steps` DDIM steps on linspace(1, 0, ref_steps + 1) -- the same share of steps inside the
    alpha-gate of nse.py:643 as the full trajectory -- timed as a whole.  The rate is normalised to the workload's
    trajectory length: samples * (ref_steps / ddim_steps) / seconds.  Returns (rate, cores, seconds of the call)."""
    import torch
    from oracle import ref_runner
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    key = (device, args.task, args.n_obs)
    if key not in reference_rate.cache:
        net, prior, fn = ref_runner.build(w, device)
        ref_runner.ddim_seconds(net, prior, fn, w, args.samples, 2, args.cov_mode, device)  # warm-up (thread pools, vmap)
        reference_rate.cache[key] = (net, prior, fn)
    net, prior, fn = reference_rate.cache[key]
    best = min(ref_runner.ddim_seconds(net, prior, fn, w, args.samples, ref_steps, args.cov_mode, device)
               for _ in range(repeats))
    return args.samples * (ref_steps / w["ddim_steps"]) / best, cores, best


reference_rate.cache = {}


def cpu_baseline_record(args, w):
    """The reference's CPU path on the box's host cores (bounded sample, ~10-30 s): the staged unmodified reference when
    present (kind "reference"), else the oracle port (kind "port")."""
    from oracle import ref_runner
    N, n = args.samples, w["n"]
    if ref_runner.available():
        ref_steps = 100 if args.cov_mode == "GAUSS" else 40
        ref_steps = min(ref_steps, w["ddim_steps"])
        rate, cores, secs = reference_rate(args, w, ref_steps, "cpu", repeats=2)
        return {"value": rate, "unit": "samples/s", "cores": cores, "kind": "reference",
                "sample": f"unmodified reference NSE.ddim (oracle/_ref) on the host cores: best of 2 complete {ref_steps}-step "
                          f"trajectories at N={N}, n_obs={n}; rate normalised to the {w['ddim_steps']}-step trajectory",
                "ms_per_ddim_step": secs / ref_steps * 1e3, "seconds_timed": secs}
    rate, cores, dt = cpu_oracle_rate(args, w, args.cpu_sample_steps)
    return {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{args.cpu_sample_steps} of {w['ddim_steps']} DDIM steps at N={N}, n_obs={n}, extrapolated "
                      f"linearly; oracle port with torch-CPU GEMMs (oracle/_ref not staged)",
            "ms_per_ddim_step": dt * 1e3}


def run_reference(args):
    """Reference arm: the reference's own `NSE.ddim` (torch, CPU by default; `--ref-device cuda` = torch eager on the
    B200) on this arm's config.  One bench step = one complete reference trajectory of `--ref-ddim-steps` steps."""
    w = workload(args)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_runner
    t_all = time.perf_counter()
    if ref_runner.available():
        dev = args.ref_device
        ref_steps = args.ref_ddim_steps or ((50 if args.cov_mode == "GAUSS" else 20) if dev == "cpu" else w["ddim_steps"])
        ref_steps = min(ref_steps, w["ddim_steps"])
        for _ in range(min(args.warmup, 1)):
            reference_rate(args, w, 2, dev)
        rates, secs = [], []
        for _ in range(args.steps):
            r, cores, dt = reference_rate(args, w, ref_steps, dev)
            rates.append(r)
            secs.append(dt)
        kind = "reference"
        sample = (f"unmodified reference NSE.ddim (oracle/_ref, torch {'CPU, all host threads' if dev == 'cpu' else 'eager on ' + dev}): "
                  f"one complete {ref_steps}-step trajectory per bench step (N={args.samples}, n_obs={w['n']}); value = samples x "
                  f"({ref_steps}/{w['ddim_steps']}) / seconds, i.e. normalised to the {w['ddim_steps']}-step trajectory")
        ms_per_step = float(np.median(secs)) * 1e3
    else:  # the staged reference is missing (make_ref.sh was not run where /root/reference exists): oracle port
        rates, secs = [], []
        n_s = max(2, args.cpu_sample_steps // 2)
        for _ in range(min(args.warmup, 1)):
            cpu_oracle_rate(args, w, 2)
        for _ in range(args.steps):
            r, cores, dt = cpu_oracle_rate(args, w, n_s)
            rates.append(r)
            secs.append(dt * n_s)
        kind = "port"
        ref_steps = n_s
        sample = (f"oracle port (numpy + torch-CPU GEMMs): {n_s} of {w['ddim_steps']} DDIM steps per bench step "
                  f"(N={args.samples}, n_obs={w['n']}), rate normalised to the full trajectory")
        ms_per_step = float(np.median(secs)) * 1e3
    value = float(np.median(rates))
    line = {"impl": "reference", "metric": "posterior samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, w, 1),
            "ref_ddim_steps_per_bench_step": ref_steps, "ref_device": args.ref_device,
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (recipe line of B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# observation-sharded stress record (BASELINE config 5): the path WITH an exchange step
# ---------------------------------------------------------------------------------------------------------
def obs_sharded_record(args, dev, rank, world):
    """theta = 10, n_obs = 10^4, MLP 26->256^3->10, Gaussian prior, GAUSS: every rank evaluates its slice of the
    observations for ALL `--stress-samples` samples; the per-step sum over ranks of the [N, 2m] partial sums happens inside
    the fused tcgen05 trajectory kernel over mapped peer memory ("p2p").  Reported: ms per DDIM step (CUDA events, max over
    ranks) for p2p, for the CUDA-graph-captured kernel -> ncclAllReduce -> kernel loop and for the single-GPU trajectory
    measured in the SAME run on rank 0, the bare NCCL all-reduce latency of the exchanged buffer, and the parity of the
    sharded result against the single-GPU one.  The reference cannot run this shape at all (it materialises
    [N, n, m, m]: 3.3 TB at these sizes, tall_posterior_sampler.py:114-128)."""
    import copy
    import torch
    import torch.distributed as dist
    from d4s import engine, parallel
    import d4s

    a2 = copy.copy(args)
    a2.task, a2.n_obs, a2.cov_mode, a2.ddim_steps = "stress", args.stress_n_obs, "GAUSS", args.stress_steps
    w = workload(a2)
    N, m, n, steps = args.stress_samples, w["m"], w["n"], args.stress_steps
    nse = build_nse(w, dev)
    x = torch.as_tensor(w["x"]).to(dev)
    cov = torch.as_tensor(w["cov"]).to(dev)
    mu, pc = torch.as_tensor(w["prior"]["mean"]).to(dev), torch.as_tensor(w["prior"]["cov"]).to(dev)
    prior, fn = torch.distributions.MultivariateNormal(mu, pc, validate_args=False), d4s.get_vpdiff_gaussian_score(mu, pc, nse)
    kw = dict(eta=1.0, prior=prior, prior_type="gaussian", prior_score_fn=fn, dist_cov_est=cov, cov_mode="GAUSS")
    flops_step = float(N) * n * flops_per_pair(w)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(launch):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def single_gpu_ms(n_steps):
        grid = engine.time_grid(n_steps)
        spec = engine.prior_spec("gaussian", prior, fn, m)
        st = engine.build_setup(nse, x, N, grid[:-1], grid[:-1].double() - 1.0 / n_steps, 1.0, spec, "GAUSS", cov, seed=7)
        th0 = engine.philox_normal(N, m, 7, -1, 0, dev)
        th = th0.clone()
        engine.run_trajectory(st, th, None, True)  # warm-up
        torch.cuda.synchronize()
        th.copy_(th0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.run_trajectory(st, th, None, True)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / n_steps

    rec = {"workload": f"stress: theta={m}, x={w['d']}, MLP {w['n_in']}->256x256x256->{m}, n_obs={n}, {N} posterior samples "
                       f"(all of them on every rank), DDIM {steps} steps, GAUSS, gaussian prior",
           "n_gpus": world, "n_obs": n, "samples": N, "ddim_steps": steps}
    if world == 1:
        ms = single_gpu_ms(steps)
        rec.update({"transport": "none (1 GPU: plain fused trajectory kernel)", "ms_per_ddim_step": ms,
                    "samples_per_s": N / (ms * steps * 1e-3), "algorithmic_tflops": flops_step / (ms * 1e-3) / 1e12,
                    "allreduce_bytes_per_step": 0})
        return rec
    # ---- parity: sharded vs single GPU on a small seeded problem (same Philox noise on both sides) ----------
    Nc, sc = 64, 8
    ref = nse.ddim((Nc,), x, steps=sc, _seed=3, **kw)
    out = parallel.ddim_obs_sharded(nse, (Nc,), x, steps=sc, _seed=3, exchange="p2p", **kw)
    out_n = parallel.ddim_obs_sharded(nse, (Nc,), x, steps=sc, _seed=3, exchange="nccl", **kw)
    torch.cuda.synchronize()
    scale = max(1.0, float(ref.abs().max()))
    par = torch.tensor([float((out - ref).abs().max()) / scale, float((out_n - ref).abs().max()) / scale], device=dev)
    same = out.clone()
    dist.broadcast(same, 0)
    ident = torch.tensor([1.0 if torch.equal(same, out) else 0.0], device=dev)
    dist.all_reduce(par, op=dist.ReduceOp.MAX)
    dist.all_reduce(ident, op=dist.ReduceOp.MIN)
    # ---- p2p: one launch per rank, exchange inside the kernel ------------------------------------------------
    traj = parallel.P2PTrajectory(nse, (N,), x, steps=steps, _seed=7, **kw)
    warm = parallel.P2PTrajectory(nse, (N,), x, steps=3, _seed=7, **kw)
    warm.launch()
    ms_p2p = timed(lambda: traj.launch()) / steps
    # ---- NCCL: kernel -> all-reduce -> kernel per step, whole loop in one CUDA graph --------------------------
    ns = max(10, steps // 5)
    ntraj = parallel.NcclTrajectory(nse, (N,), x, steps=ns, _seed=7, graph=True, **kw)
    ntraj.launch()
    ms_nccl = timed(lambda: ntraj.launch()) / ns
    ntraj.close()  # a graph holding NCCL kernels must not outlive the process group
    etraj = parallel.NcclTrajectory(nse, (N,), x, steps=ns, _seed=7, graph=False, **kw)
    etraj.launch()
    ms_eager = timed(lambda: etraj.launch()) / ns
    # bare all-reduce of the exchanged buffer
    buf = torch.zeros(ntraj.exchanged_bytes_per_step // 4, device=dev)
    for _ in range(5):
        dist.all_reduce(buf)
    reps = 50

    def ar():
        for _ in range(reps):
            dist.all_reduce(buf)
    us_ar = timed(ar) / reps * 1e3
    # single-GPU time of the same problem, measured in this run on rank 0 (the other ranks wait)
    single = torch.zeros(1, dtype=torch.float64, device=dev)
    if rank == 0:
        single[0] = single_gpu_ms(max(5, steps // 10))
    dist.all_reduce(single, op=dist.ReduceOp.MAX)
    ms_single = float(single)
    rec.update({"transport": "p2p: in-kernel exchange over mapped peer memory (NVLink), one launch per trajectory and rank",
                "ms_per_ddim_step": ms_p2p, "samples_per_s": N / (ms_p2p * steps * 1e-3),
                "algorithmic_tflops": flops_step / (ms_p2p * 1e-3) / 1e12,
                "single_gpu_ms_per_ddim_step": ms_single, "speedup_vs_single_gpu": ms_single / ms_p2p,
                "efficiency_vs_single_gpu": ms_single / ms_p2p / world,
                "exchanged_bytes_per_step_per_rank": traj.exchanged_bytes_per_step * world,
                "comm_nranks": world,
                "nccl_graph": {"ms_per_ddim_step": ms_nccl, "steps_timed": ns, "efficiency_vs_single_gpu": ms_single / ms_nccl / world,
                               "allreduce_bytes_per_step": ntraj.exchanged_bytes_per_step,
                               "launches_per_step": "kernel 1 (export) + ncclAllReduce + kernel 3/4, one CUDA graph per trajectory"},
                "nccl_eager_ms_per_ddim_step": ms_eager,
                "nccl_allreduce_alone_us": us_ar,
                "kernel_ms_per_step_vs_allreduce": {"kernel_ms": ms_p2p, "allreduce_ms": us_ar * 1e-3},
                "parity_vs_single_gpu": {"samples": Nc, "steps": sc, "max_rel_diff_p2p": float(par[0]),
                                         "max_rel_diff_nccl": float(par[1]), "identical_on_all_ranks": bool(float(ident) == 1.0)}})
    return rec


# ---------------------------------------------------------------------------------------------------------
# the metric's other axes: samples/s vs n_obs for GAUSS and JAC (sbibm_posterior_estimation.py:21 N_OBS_LIST)
# ---------------------------------------------------------------------------------------------------------
def trajectory_rate(args, dev, n_obs, cov_mode, reps=2, e2e=False):
    """(device-resident samples/s, ms per trajectory, launches per trajectory, e2e samples/s or None) of one config of the
    default task: tables resident, CUDA events around d4s_ddim_run, L2 flushed by the caller's 256 MiB buffer."""
    import copy
    import torch
    import d4s
    from d4s import engine
    a2 = copy.copy(args)
    a2.n_obs, a2.cov_mode, a2.ddim_steps = n_obs, cov_mode, None
    w = workload(a2)
    N, m = args.samples, w["m"]
    nse = build_nse(w, dev)
    x = torch.as_tensor(w["x"]).to(dev)
    cov = torch.as_tensor(w["cov"]).to(dev)
    if w["prior_kind"] == "uniform":
        lo, hi = torch.as_tensor(w["prior"]["low"]).to(dev), torch.as_tensor(w["prior"]["high"]).to(dev)
        prior, fn = torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
    else:
        mu, pc = torch.as_tensor(w["prior"]["mean"]).to(dev), torch.as_tensor(w["prior"]["cov"]).to(dev)
        prior, fn = torch.distributions.MultivariateNormal(mu, pc, validate_args=False), d4s.get_vpdiff_gaussian_score(mu, pc, nse)
    kw = dict(steps=w["ddim_steps"], eta=w["eta"], prior=prior, prior_type=w["prior_kind"], prior_score_fn=fn,
              dist_cov_est=cov, cov_mode=cov_mode)
    xs, kwn = x, {}
    if w["n_max"] > 1:
        xs, ns = nse._create_pf_data(x)
        kwn = dict(n=ns)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    # the public call once (warm-up; also the e2e figure), then the resident-table launches
    nse.ddim((N,), xs, _seed=1, **kw, **kwn)
    torch.cuda.synchronize()
    e2e_rate = None
    if e2e:
        flush.fill_(1.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = nse.ddim((N,), xs, _seed=2, **kw, **kwn).cpu()
        e2e_rate = N / (time.perf_counter() - t0)
        assert bool(torch.isfinite(out).all())
    single = n_obs == 1
    spec = engine.PriorSpec(0) if single else engine.prior_spec(w["prior_kind"], prior, fn, m)
    grid = engine.time_grid(w["ddim_steps"])
    setup = engine.build_setup(nse, xs, N, grid[:-1], grid[:-1].double() - 1.0 / w["ddim_steps"], w["eta"], spec, cov_mode,
                               None if single else cov, n_sizes=kwn.get("n"), seed=5, weighted=not single)
    th0 = engine.philox_normal(N, m, 5, -1, 0, dev)
    th = th0.clone()
    ms = []
    l0 = engine.launch_count()
    for k in range(reps + 1):
        th.copy_(th0)
        flush.fill_(float(k))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.run_trajectory(setup, th, None, True)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    launches = (engine.launch_count() - l0) // (reps + 1)
    assert bool(torch.isfinite(th).all())
    best = float(np.mean(ms[1:]))
    return N / (best * 1e-3), best, int(launches), e2e_rate, w


def per_call_record(args, dev):
    """`NSE.factorized_score` through the public per-call API at the bench shape (what `euler_sde_sampler` or a
    user-written sampler calls once per step): steady-state wall time per call over a loop that revisits 20 time points
    (tables cached per (x, dist_cov_est, prior, t)), and the first-visit cost of a new t."""
    import torch
    import d4s
    w = workload(args)
    N, m = args.samples, w["m"]
    nse = build_nse(w, dev)
    x, cov = torch.as_tensor(w["x"]).to(dev), torch.as_tensor(w["cov"]).to(dev)
    if w["prior_kind"] == "uniform":
        lo, hi = torch.as_tensor(w["prior"]["low"]).to(dev), torch.as_tensor(w["prior"]["high"]).to(dev)
        prior, fn = torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
    else:
        mu, pc = torch.as_tensor(w["prior"]["mean"]).to(dev), torch.as_tensor(w["prior"]["cov"]).to(dev)
        prior, fn = torch.distributions.MultivariateNormal(mu, pc, validate_args=False), d4s.get_vpdiff_gaussian_score(mu, pc, nse)
    theta = torch.randn(N, m, device=dev)
    ts = [torch.tensor(0.05 + 0.045 * k) for k in range(20)]
    out = {}
    for mode in ("GAUSS", "JAC"):
        call = lambda t: nse.factorized_score(theta, x, t, fn, prior, dist_cov_est=cov, cov_mode=mode,  # noqa: E731
                                              prior_type=w["prior_kind"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in ts:
            call(t)
        torch.cuda.synchronize()
        first = (time.perf_counter() - t0) / len(ts)
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            for t in ts:
                call(t)
        torch.cuda.synchronize()
        out[mode] = {"steady_us": (time.perf_counter() - t0) / (reps * len(ts)) * 1e6, "first_visit_us": first * 1e6}
    out["note"] = (f"N={N}, n_obs={w['n']}; steady = tables of the (x, cov, prior, t) combination cached on the device, "
                   "loop not synchronised per call; first_visit = new t (schedule row + step matrices built on the host)")
    return out


def prepass_record(args, dev):
    """Covariance pre-pass of the GAUSS pipeline (sbibm_posterior_estimation.py:306-314: `vmap(ddim)` over the observations,
    1000 draws each, 100 steps, eta 0.5, then `torch.cov`) through `NSE.estimate_dist_cov`: one launch of the paired-mode
    tcgen05 trajectory kernel; wall time per call, host tables included."""
    import torch
    from d4s.engine import launch_count
    w = workload(args)
    nse = build_nse(w, dev)
    x = torch.as_tensor(w["x"]).to(dev)
    for _ in range(2):
        nse.estimate_dist_cov(x, nsamples=1000, steps=100, eta=0.5, _seed=11)
    torch.cuda.synchronize()
    reps = 5
    l0 = launch_count()
    t0 = time.perf_counter()
    for k in range(reps):
        cov = nse.estimate_dist_cov(x, nsamples=1000, steps=100, eta=0.5, _seed=12 + k)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    return {"ms_per_call": ms, "n_obs": int(x.shape[0]), "samples_per_observation": 1000, "ddim_steps": 100, "eta": 0.5,
            "samples_per_s": x.shape[0] * 1000 / (ms * 1e-3), "launches_per_call": (launch_count() - l0) / reps,
            "finite": bool(torch.isfinite(cov).all()),
            "note": "NSE.estimate_dist_cov = vmap(ddim) over the observations + torch.cov; paired_tc_kernel (tcgen05), one launch "
                    "per trajectory + the Philox initial draw; wall clock per call incl. host tables"}


def n_obs_sweep(args, dev):
    """samples/s of complete trajectories for n_obs in {1, 8, 14, 22, 30} x {GAUSS (1000 steps, eta 1), JAC (400 steps,
    eta 0.8)} on the default task's shape, N = --samples (sbibm_posterior_estimation.py:21, 330-334, 617-619)."""
    out = {"n_obs": [1, 8, 14, 22, 30], "unit": "samples/s", "samples": args.samples,
           "note": "device-resident tables, CUDA events around d4s_ddim_run, mean of 2 trajectories after 1 warm-up; "
                   "launches = kernel launches per trajectory (1 = fused tcgen05 trajectory kernel)"}
    for mode in ("GAUSS", "JAC"):
        rates, launches = {}, {}
        for n in out["n_obs"]:
            r, _, l, _, _ = trajectory_rate(args, dev, n, mode)
            rates[str(n)] = r
            launches[str(n)] = l
        out[mode] = rates
        out[mode + "_launches"] = launches
    return out


def jac_record(args, dev, peaks):
    """The JAC half of the metric on the default task (n_obs = --n-obs, 400 steps, eta 0.8): value, e2e, roofline of the
    Jacobian step kernel, and the reference's CPU path on the same config."""
    import copy
    import torch
    from d4s import engine
    rate, ms, launches, e2e_rate, w = trajectory_rate(args, dev, args.n_obs, "JAC", reps=2, e2e=True)
    a2 = copy.copy(args)
    a2.cov_mode, a2.ddim_steps = "JAC", None
    N, m = args.samples, w["m"]
    # steps inside the alpha-gate evaluate the Jacobian (forward-mode: 1 + m rows per pair); the others are plain-sum steps
    t = engine.time_grid(w["ddim_steps"])[:-1].double()
    gate = int((torch.exp(-16 * t ** 2).sqrt() > 0.5).sum()) if args.n_obs > 1 else 0
    flops = float(N) * w["n_units"] * flops_per_pair(w) * ((1 + m) * gate + (w["ddim_steps"] - gate))
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = flops / (ms * 1e-3) / 1e12
    from d4s import _cabi as abi
    simt_rate = None
    if not getattr(args, "jac_simt", False):  # the same trajectory with the Jacobian steps on the fp32 SIMT kernel, for comparison
        abi.set_option("jac_tc", 0)
        try:
            simt_rate = trajectory_rate(args, dev, args.n_obs, "JAC", reps=1)[0]
        finally:
            abi.set_option("jac_tc", 1)
    rec = {"value": rate, "unit": "samples/s", "ms_per_trajectory": ms, "ddim_steps": w["ddim_steps"], "eta": w["eta"],
           "gpu_launches_per_trajectory": launches, "jacobian_steps": gate,
           "jac_kernel": "fp32 SIMT (--jac-simt)" if getattr(args, "jac_simt", False) else
                         "jac_tc_kernel (tcgen05 kind::f16, fp16x3 split, cross terms accumulated first, forward-mode tangent rows)",
           "fp32_simt_jac_kernel_samples_per_s": simt_rate,
           "e2e": {"value": e2e_rate, "unit": "samples/s"},
           "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "flops_per_trajectory": flops,
                        "note": "algorithmic flops of the whole JAC trajectory (forward-mode tangents counted on the "
                                "Jacobian steps only) / trajectory time; kernels: traj_tc_kernel on the plain-sum steps, "
                                "the JAC step kernel + finalize_kernel inside the alpha-gate"}}
    if not args.no_cpu_baseline:
        rec["cpu_baseline"] = cpu_baseline_record(a2, workload(a2))
    return rec


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import d4s
    from d4s import _cabi as abi
    from d4s import engine

    w = workload(args)
    N, m, n = args.samples, w["m"], w["n"]
    if getattr(args, "jac_simt", False):
        abi.set_option("jac_tc", 0)
    nse = build_nse(w, dev)
    # host (pinned) copies of the inputs for the e2e leg
    x_h = torch.as_tensor(w["x"]).pin_memory()
    cov_h = torch.as_tensor(w["cov"]).pin_memory()
    if w["prior_kind"] == "uniform":
        lo_h = torch.as_tensor(w["prior"]["low"]).pin_memory()
        hi_h = torch.as_tensor(w["prior"]["high"]).pin_memory()
    else:
        mu_h = torch.as_tensor(w["prior"]["mean"]).pin_memory()
        pc_h = torch.as_tensor(w["prior"]["cov"]).pin_memory()

    def make_prior():
        if w["prior_kind"] == "uniform":
            lo, hi = lo_h.to(dev, non_blocking=True), hi_h.to(dev, non_blocking=True)
            return torch.distributions.Uniform(lo, hi, validate_args=False), d4s.get_vpdiff_uniform_score(lo, hi, nse)
        mu, pc = mu_h.to(dev, non_blocking=True), pc_h.to(dev, non_blocking=True)
        return (torch.distributions.MultivariateNormal(mu, pc, validate_args=False),
                d4s.get_vpdiff_gaussian_score(mu, pc, nse))

    out_h = torch.empty(N, m).pin_memory()
    seed0 = 1234

    def e2e_step(it):
        """The call a user makes: host inputs -> NSE.ddim -> samples back on the host."""
        x = x_h.to(dev, non_blocking=True)
        cov = cov_h.to(dev, non_blocking=True)
        prior, fn = make_prior()
        th = nse.ddim((N,), x, steps=w["ddim_steps"], eta=w["eta"], prior=prior, prior_type=w["prior_kind"],
                      prior_score_fn=fn, dist_cov_est=cov, cov_mode=args.cov_mode, _seed=seed0 + it,
                      _sample_offset=rank * N)
        out_h.copy_(th, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_h

    # device-resident leg: tables built once, only d4s_ddim_run in the timed region
    x_d, cov_d = x_h.to(dev), cov_h.to(dev)
    prior, fn = make_prior()
    spec = engine.prior_spec(w["prior_kind"], prior, fn, m)
    grid = engine.time_grid(w["ddim_steps"])
    xs_d, ns_d = nse._create_pf_data(x_d) if w["n_max"] > 1 else (x_d, None)  # PF_NSE: subsets of n_max observations
    setup = engine.build_setup(nse, xs_d, N, grid[:-1], grid[:-1].double() - 1.0 / w["ddim_steps"], w["eta"], spec,
                               args.cov_mode, cov_d, n_sizes=ns_d, seed=seed0, sample_offset=rank * N)
    theta0 = engine.philox_normal(N, m, seed0, -1, rank * N, dev)
    theta = theta0.clone()
    h2d = sum(t.numel() * 4 for t in (x_h, cov_h)) + sum(t.numel() * 4 for t in (setup.sched, setup.time_feat, setup.mats)) \
        + sum(t.numel() * 4 for t in setup.keep[:4] if t is not None)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn_step, k):
        flush.fill_(float(k))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn_step(k)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    def dev_step(k):
        theta.copy_(theta0)
        engine.run_trajectory(setup, theta, None, True)

    for k in range(args.warmup):
        dev_step(k)
        e2e_step(k)
    barrier()
    clocks = Clocks(local)
    clocks.start()
    l0 = engine.launch_count()
    barrier()
    ms_dev = [timed(dev_step, k) for k in range(args.steps)]
    barrier()
    launches = engine.launch_count() - l0
    ms_e2e = []
    prof = None
    if os.environ.get("D4S_BENCH_PROFILE_E2E"):  # host-side profile of the end-to-end steps, to stderr
        import cProfile
        prof = cProfile.Profile()
    for k in range(args.steps):
        flush.fill_(float(k))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if prof:
            prof.enable()
        e2e_step(k)
        if prof:
            prof.disable()
        ms_e2e.append((time.perf_counter() - t0) * 1e3)
    if prof and rank == 0:
        import io
        import pstats
        buf = io.StringIO()
        pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(25)
        sys.stderr.write(f"e2e ms per step: {ms_e2e}\n" + buf.getvalue()[:5000])
    barrier()
    clk = clocks.stop()
    assert torch.isfinite(theta).all(), "non-finite samples"

    # per-step fp32 SIMT kernels (the JAC path, and the GAUSS fallback): average launch durations with CUDA events
    reps = 200
    engine.score_partials(setup, theta, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps):
        engine.score_partials(setup, theta, (k * 7) % w["ddim_steps"])
    e1.record()
    e1.synchronize()
    k1_ms = e0.elapsed_time(e1) / reps
    e0.record()
    for k in range(reps):
        engine.finalize(setup, theta, (k * 7) % w["ddim_steps"])
    e1.record()
    e1.synchronize()
    k3_ms = e0.elapsed_time(e1) / reps
    fused = launches <= 2 * args.steps  # the fused tcgen05 trajectory kernel ran (one launch per trajectory)

    obs_rec = None
    if not args.no_obs_sharded:
        try:
            obs_rec = obs_sharded_record(args, dev, rank, world)
        except Exception as exc:  # reported, never silently dropped
            obs_rec = {"error": f"{type(exc).__name__}: {exc}"}
    tot_dev = torch.tensor([sum(ms_dev), sum(ms_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_dev, op=dist.ReduceOp.MAX)
    tot_dev = tot_dev.cpu().tolist()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        rpp = (1 + m) if args.cov_mode == "JAC" else 1
        units = w["n_units"]  # score-net evaluations per sample and step: observations, or PF_NSE subsets
        flops_step = float(N) * units * flops_per_pair(w) * rpp  # algorithmic flops of one DDIM step
        ms_step = tot_dev[0] / args.steps
        dims = [w["n_in"], *w["hidden"], m]
        hidden_flops = 2.0 * sum(a * b for a, b in zip(dims[1:-2], dims[2:-1])) * N * units  # layers on the tensor core
        n_k1 = w["ddim_steps"] if setup.step_modes is None else int((setup.step_modes == 1).sum())
        if fused:
            # dominant kernel = traj_tc_kernel: ONE launch per trajectory; its duration is the timed region itself.
            peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
            src = ("measured (MEASURED_PEAKS.json bf16_tflops_sustained: the kernel runs for a whole trajectory)"
                   if peaks else "fallback 1400 sustained")
            launch_ms = float(np.mean(ms_dev))
            ach = flops_step * w["ddim_steps"] / (launch_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                        "traffic": TRAJ_TC_DRAM_BYTES if is_default_workload(args, w) else None,
                        "traffic_source": "profiles/r02_ncu_full_traj_tc.txt (a separate ncu --set full capture of this command's "
                                          "launch, NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum)",
                        "kernel": "traj_tc_kernel (tcgen05 kind::f16, bf16x3 split, fp32 TMEM accumulators)",
                        "peak_source": src, "flops_per_launch": flops_step * w["ddim_steps"], "launch_ms": launch_ms,
                        "share_of_step": launch_ms / ms_step,
                        "note": "achieved counts ALGORITHMIC flops (2*in*out per pair and layer); the bf16x3 split issues 3 "
                                "MMAs per product, so the tensor pipe executes 3x the hidden-layer flops",
                        "tensor_pipe_executed_tflops": 3.0 * hidden_flops * w["ddim_steps"] / (launch_ms * 1e-3) / 1e12,
                        "simt_score_kernel_launch_ms": k1_ms, "simt_finalize_launch_ms": k3_ms}
        else:
            peak_tf = peaks.get("bf16_tflops", 1590.0)
            src = "measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)" if peaks else "fallback 1590"
            ach = flops_step / (k1_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                        "traffic": None,
                        "kernel": ("jac_tc_kernel (tcgen05 kind::f16, fp16x3 split, forward-mode tangent rows)"
                                   if (args.cov_mode == "JAC" and not getattr(args, "jac_simt", False))
                                   else "score_simt_kernel (fp32 SIMT; forward-mode tangents in JAC)"),
                        "peak_source": src, "flops_per_launch": flops_step, "launch_ms": k1_ms,
                        "finalize_launch_ms": k3_ms,
                        # launches of this kernel per trajectory = steps inside the alpha-gate (the others run fused)
                        "launches_per_trajectory": int(n_k1),
                        "share_of_step": min(1.0, k1_ms * n_k1 / ms_step)}
        value = N * world / (ms_step * 1e-3)
        e2e_val = N * world / (tot_dev[1] / args.steps * 1e-3)
        line = {"metric": "posterior samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, w, world),
                "clocks": clk,
                "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(N * m * 4), "ms_per_step": tot_dev[1] / args.steps,
                        "ms_steps_rank0": [round(v, 3) for v in ms_e2e]},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "us_per_ddim_step": ms_step * 1e3 / w["ddim_steps"]}
        if obs_rec is not None:
            line["obs_sharded"] = obs_rec
        if world == 1 and not args.no_sweep and args.task == "slcp" and args.cov_mode == "GAUSS":
            line["sweep"] = n_obs_sweep(args, dev)
            line["jac"] = jac_record(args, dev, peaks)
            line["per_call_us"] = per_call_record(args, dev)
            line["prepass"] = prepass_record(args, dev)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_record(args, w)
        print(json.dumps(line))
    if world > 1:
        from d4s import parallel
        parallel.shutdown()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
