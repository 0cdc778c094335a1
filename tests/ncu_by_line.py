"""Maps the SASS-level source page of an ncu report back to source lines (not a pytest).
usage: python tests/ncu_by_line.py <report.ncu-rep> <kernel cubin> <function substring> [top N]
  cubin: cuobjdump -xelf all libd4s.so -> d4s_traj_tc.sm_100a.cubin (same build as the profiled one, -lineinfo)
Prints, per source line (file:line of the innermost inlined frame nvdisasm -g reports), executed warp instructions, stall
samples and the dominant stall reasons; profiles/r02_ncu_lines_*.txt are outputs of this script."""
import collections
import csv
import io
import re
import subprocess
import sys

rep, cubin, func = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60

within = int(sys.argv[sys.argv.index("--within") + 1]) if "--within" in sys.argv else None  # only code inlined at this line
outer = "--outer" in sys.argv  # attribute to the OUTERMOST frame (call site in the kernel body) instead of the innermost
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
addr_line = {}
chain, in_func, fresh = [], False, True
for ln in dis.splitlines():
    if ln.startswith("\t.global") or ln.startswith(".text.") or ".section\t.text." in ln:
        in_func = func in ln
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        fresh = True
        if in_func and chain:
            if within is not None and chain[-1][1] != within:
                continue
            addr_line[int(m.group(1), 16)] = (chain[-1] if outer else chain[0], m.group(2).strip())

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
base = None
tot_inst = tot_samp = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    base = a if base is None else base
    if within is not None and (a - base) not in addr_line:
        continue
    key, sass = addr_line.get(a - base, (("?", 0), ""))
    inst = int(r[col["Instructions Executed"]] or 0)
    samp = int(r[col["# Samples"]] or 0)
    c = agg[key]
    c["inst"] += inst
    c["samp"] += samp
    c["conf"] += int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
    for s in stall_cols:
        c[s] += int(r[col[s]] or 0)
    tot_inst += inst
    tot_samp += samp
print(f"total warp instructions {tot_inst}, samples {tot_samp}")
print(f"{'line':>28s} {'inst %':>7s} {'samp %':>7s} {'smem excess':>12s}  top stalls")
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = sorted(((c[s], s[6:]) for s in stall_cols), reverse=True)[:3]
    print(f"{key[0] + ':' + str(key[1]):>28s} {100 * c['inst'] / max(tot_inst, 1):7.2f} {100 * c['samp'] / max(tot_samp, 1):7.2f} "
          f"{c['conf']:12d}  " + ", ".join(f"{n}:{100 * v / max(c['samp'], 1):.0f}%" for v, n in st))
if "--ranges" in sys.argv:
    # aggregate by ranges of lines given as name=lo-hi,... after --ranges
    spec = sys.argv[sys.argv.index("--ranges") + 1]
    print("\nby range:")
    for item in spec.split(","):
        name, rng = item.split("=")
        lo, hi_ = (int(v) for v in rng.split("-"))
        inst = sum(c["inst"] for k, c in agg.items() if k[0].startswith("d4s_traj_tc") and lo <= k[1] <= hi_)
        samp = sum(c["samp"] for k, c in agg.items() if k[0].startswith("d4s_traj_tc") and lo <= k[1] <= hi_)
        print(f"  {name:24s} inst {100 * inst / tot_inst:6.2f} %  samples {100 * samp / tot_samp:6.2f} %")
