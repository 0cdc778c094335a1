"""Same-box A/B of kernel builds.  usage: python tests/gpu_ab.py [--rounds 2] [--tasks slcp,lotka_volterra] lib1.so lib2.so ...
Every (lib, task) pair is one fresh process of bench.py (device-resident value, 5 trajectories after 2 warm-ups), rounds
alternate between the builds so that clock / thermal drift hits all of them alike."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
rounds, tasks = 2, ["slcp", "lotka_volterra"]
while args and args[0].startswith("--"):
    if args[0] == "--rounds":
        rounds = int(args[1])
    elif args[0] == "--tasks":
        tasks = args[1].split(",")
    args = args[2:]
libs = args
res = {(l, t): [] for l in libs for t in tasks}
for r in range(rounds):
    for t in tasks:
        for l in libs:
            env = dict(os.environ, D4S_LIB_PATH=os.path.join(ROOT, l))
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--task", t, "--steps", "5", "--warmup", "2",
                                  "--no-cpu-baseline", "--no-obs-sharded", "--no-sweep"], env=env, capture_output=True, text=True)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                res[(l, t)].append(d["value"])
            except Exception:
                print("FAILED", l, t, out.stderr[-600:], flush=True)
                res[(l, t)].append(float("nan"))
for t in tasks:
    base = None
    for l in libs:
        v = res[(l, t)]
        mean = sum(v) / len(v)
        base = base or mean
        print(f"{t:16s} {l:40s} {mean:9.0f} samples/s  ({(mean / base - 1) * 100:+5.1f} %)  runs: {' '.join(f'{x:.0f}' for x in v)}", flush=True)
