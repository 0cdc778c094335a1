"""Turns gpurun_out ncu artefacts into the small text summaries committed under profiles/.
usage: python profiles/summarize.py launches <launches.csv> | full <prof.ncu-rep> [kernel-regex]"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) > vi:
            agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, gpu__time_duration.sum (ns); cold-cache + serialised => compare SHARES")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} n={len(v):5d} avg_ns={sum(v) / len(v):12.1f} share={sum(v) / tot:.3f}")


def full(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print(f"# {path}: ncu --set full, {len(data)} launch(es)")
    for r in data:
        if pat and pat not in r[ki]:
            continue
        print(f"## {r[ki][:90]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:85s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
