"""Builds an experimental variant of libd4s.so next to the default one (A/B runs: tests/gpu_ab.py with D4S_LIB_PATH).
usage: python tests/build_variant.py <name> [nvcc flags ...]     e.g.  python tests/build_variant.py e121 -DD4S_EXP=121"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusions-for-sbi_b200"))
from d4s import _build

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "ab", f"libd4s_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
print(_build.build(force=True, extra_flags=flags, out=out, obj_dir=os.path.join("/tmp", "d4s_obj_" + name), verbose="-v" in flags))
