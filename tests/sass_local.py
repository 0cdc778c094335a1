"""Lists the local-memory instructions (LDL / STL: spills and dynamically indexed arrays) of a kernel with the source line of
their outermost frame (not a pytest).  usage: python tests/sass_local.py <lib.so> [kernel substring] [min line] [max line]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib = os.path.abspath(sys.argv[1])
func = sys.argv[2] if len(sys.argv) > 2 else "traj_tc_kernel"
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if "traj_tc" in f or func.split("_kernel")[0] in f]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, cub[0])], capture_output=True, text=True).stdout
chain, fresh, infunc = [], True, False
sites = collections.Counter()
for ln in dis.splitlines():
    if ".section\t.text." in ln or ln.startswith(".text."):
        infunc = func in ln and "$" not in ln.split(func)[-1]
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        fresh = True
        if infunc and re.search(r"\b(LDL|STL)", m.group(2)) and chain and lo <= chain[-1][1] <= hi:
            sites[(chain[-1][1], chain[0][1], re.sub(r"\s+", " ", m.group(2).strip()))] += 1
for (outer, inner, sass), c in sorted(sites.items()):
    print(f"outer line {outer:5d}  inner {inner:5d}  x{c}  {sass}")
print(len(sites), "sites")
