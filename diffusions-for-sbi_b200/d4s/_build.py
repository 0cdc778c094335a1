"""Builds libd4s.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB = os.path.join(HERE, "libd4s.so")
SOURCES = ["d4s_api.cu", "d4s_score_simt.cu", "d4s_finalize.cu", "d4s_traj_tc.cu", "d4s_jac_tc.cu", "d4s_paired_tc.cu", "d4s_generic.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "-cudart", "static"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "d4s.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None, obj_dir=None):
    """Compiles every CUDA source of the package into d4s/libd4s.so.  Returns the library path.
    `extra_flags` / `out` / `obj_dir` build an experimental variant next to it (A/B runs with D4S_LIB_PATH)."""
    if out is None and not force and not needs_build():
        return LIB
    objs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + list(extra_flags)
    if obj_dir:
        os.makedirs(obj_dir, exist_ok=True)
    def compile_one(src):
        obj = os.path.join(obj_dir or CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", os.path.join(CSRC, src), "-o", obj]
        return src, obj, subprocess.run(cmd, capture_output=True, text=True)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:  # one nvcc per source, side by side
        results = list(pool.map(compile_one, SOURCES))
    for src, obj, r in results:
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    target = out or LIB
    cmd = [_nvcc(), "-shared", "-cudart", "static", "-o", target, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return target


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
