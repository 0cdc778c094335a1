"""SASS evidence of the tcgen05 / TMEM / bulk-copy instructions in the built libd4s.so (cuobjdump -sass, sm_100a).
usage: python profiles/sass_evidence.py [lib] > profiles/rNN_sass_*.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "diffusions-for-sbi_b200/d4s/libd4s.so"
KEEP = re.compile(r"^(UTCHMMA|UTCBAR|UTCATOMSWS|UBLKCP|LDTM|STTM|SYNCS|USETMAXREG|UCGABAR|FFMA2|FMUL2|BAR\.|MEMBAR|CCTL|REDUX|SHFL|LDL|STL)")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print("# SASS evidence of the tcgen05 / TMEM / bulk-copy instructions in the shipped libd4s.so (cuobjdump -sass, sm_100a).")
print("# UTCHMMA = tcgen05.mma (gdesc = shared-memory operand descriptor, tmem = tensor-memory operand: the TS form; .2CTA = cta_group::2),")
print("# LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit (.2CTA.MULTICAST = to both CTAs of a pair), UBLKCP = cp.async.bulk (1-D TMA),")
print("# SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, UTCATOMSWS = tcgen05.alloc / dealloc, UCGABAR = barrier.cluster, LDL / STL = local memory.")
func, counts, example, total = None, None, None, 0


def flush():
    if func and any(k in func for k in ("traj_tc_kernel", "jac_tc_kernel")):
        print(f"\n## {func}: {total} SASS instructions")
        for k, c in sorted(counts.items(), key=lambda kv: -kv[1]):
            print(f"  {c:6d} x {k:38s} e.g. {example[k]}")


for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        flush()
        func, counts, example, total = m.group(1), collections.Counter(), {}, 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)\s*(.*?);", ln)
    if m and func:
        total += 1
        op = m.group(1)
        if KEEP.match(op):
            counts[op] += 1
            example.setdefault(op, (op + " " + m.group(2)).strip()[:110])
flush()
