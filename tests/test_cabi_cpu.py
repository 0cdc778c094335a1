"""CPU checks of the C-ABI library: it loads, exports every symbol include/d4s.h declares, reports errors
instead of falling back, and the ctypes structures match the C layout."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from d4s import _cabi
    return _cabi.load()


def test_every_declared_symbol_is_exported(lib):
    from d4s import _cabi
    hdr = open(os.path.join(ROOT, "include", "d4s.h")).read()
    declared = sorted(set(re.findall(r"D4S_API [\w\s\*]+?\b(d4s_\w+)\(", hdr)))
    assert declared and sorted(_cabi.EXPORTS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (d4s_\w+)", out))
    assert set(declared) <= exported
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_constants(lib):
    from d4s import _cabi
    hdr = open(os.path.join(ROOT, "include", "d4s.h")).read()
    consts = dict(re.findall(r"#define (D4S_\w+) (-?\d+)", hdr))
    assert int(consts["D4S_ABI_VERSION"]) == _cabi.ABI_VERSION == lib.d4s_abi_version()
    assert int(consts["D4S_SCHED_STRIDE"]) == _cabi.SCHED_STRIDE
    assert int(consts["D4S_MAX_THETA"]) == _cabi.MAX_THETA
    for name in ("T", "ALPHA", "SIGMA", "INV_SIGMA", "A_OVER_S2", "C_THETA", "C_SCORE", "BRIDGE_STD", "SQRT_ALPHA",
                 "COMBINE", "ONE_MINUS_N"):
        assert int(consts["D4S_S_" + name]) == getattr(_cabi, "S_" + name)
    assert lib.d4s_mats_floats(5) == 2 * 25 + 5


def test_struct_layouts_match_c(lib, tmp_path):
    """Compiles a tiny C program against include/d4s.h and compares sizeof/offsetof with ctypes."""
    from d4s import _cabi
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "d4s.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(D4sNetDesc), offsetof(D4sNetDesc, weight),'
                   'sizeof(D4sProblem), offsetof(D4sProblem, obs_feat), offsetof(D4sProblem, seed),'
                   'offsetof(D4sProblem, workspace), sizeof(D4sStep), offsetof(D4sStep, step_index),'
                   'offsetof(D4sProblem, clip_lo));return 0;}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(_cabi.NetDesc), _cabi.NetDesc.weight.offset, C.sizeof(_cabi.Problem),
            _cabi.Problem.obs_feat.offset, _cabi.Problem.seed.offset, _cabi.Problem.workspace.offset,
            C.sizeof(_cabi.Step), _cabi.Step.step_index.offset, _cabi.Problem.clip_lo.offset]
    assert got == want


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly (never route through the oracle or torch)."""
    import torch
    import d4s
    from d4s import _cabi
    nse = d4s.NSE(3, 4, hidden_features=[64, 64])
    with pytest.raises(_cabi.D4sError):
        nse.ddim((8,), torch.randn(5, 4), steps=2)
    with pytest.raises(_cabi.D4sError):
        nse.factorized_score(torch.randn(8, 3), torch.randn(5, 4), torch.tensor(0.5), None, None)
    if lib.d4s_device_count() == 0:
        handle = C.c_void_p()
        desc = _cabi.NetDesc()
        desc.n_layers = 2
        desc.dims[0], desc.dims[1], desc.dims[2] = 7, 64, 3
        desc.theta_cols = 3
        w0, b0 = (C.c_float * (64 * 7))(), (C.c_float * 64)()
        w1, b1 = (C.c_float * (3 * 64))(), (C.c_float * 3)()
        desc.weight[0], desc.bias[0] = C.addressof(w0), C.addressof(b0)
        desc.weight[1], desc.bias[1] = C.addressof(w1), C.addressof(b1)
        rc = lib.d4s_net_create(C.byref(handle), C.byref(desc), None)
        assert rc == -2 and b"no CUDA device" in lib.d4s_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "diffusions-for-sbi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle_", ""), f"{f} mentions the oracle"


def test_dropin_resolves_every_hot_path_import_of_the_reference_drivers():
    """The exact hot-path import lines of the reference drivers (sbibm_posterior_estimation.py:6,13,14,
    jrnnm_posterior_estimation.py:5,14,16,17, sbibm_pf_npse.py:5,6,13, gen_gaussian_gaussian.py:11-13,15,
    gen_mixt_gauss_gaussian.py:7,8,12) against the drop-in directory (no reference checkout needed)."""
    import subprocess
    import sys
    lines = [
        "from nse import NSE, NSELoss",
        "from tall_posterior_sampler import diffused_tall_posterior_score, euler_sde_sampler, tweedies_approximation",
        "from vp_diffused_priors import get_vpdiff_gaussian_score, get_vpdiff_uniform_score",
        "from zuko.nn import MLP",
        "from nse import NSELoss",
        "from pf_nse import PF_NSE",
        "from embedding_nets import EpsilonNet, FakeFNet",
        "from nse_toy import NSE",
        "from tall_posterior_sampler import mean_backward, prec_matrix_backward",
        "from nse_toy import mean_backward, prec_matrix_backward",
        "from nse_toy import NSELoss as ToyLoss",
        "from tall_posterior_sampler import sigma_backward",
    ]
    dropin = os.path.join(ROOT, "diffusions-for-sbi_b200", "dropin")
    code = f"import sys; sys.path.insert(0, {dropin!r})\n" + "\n".join(lines) + "\nprint('IMPORTS_OK')"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "IMPORTS_OK" in r.stdout, r.stderr[-2000:]
