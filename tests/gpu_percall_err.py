"""Per-call factorized_score error of the GAUSS tensor-core path vs the fp64 oracle on the golden fixtures (not a pytest)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusions-for-sbi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import golden_io as gio
from d4s import _cabi as abi
from oracle import d4s_oracle as orc
from test_gpu_parity import _setup, _t
dev = torch.device("cuda:0")
for name in ("slcp30_n30", "lv_n6", "sir_n8", "slcp_n6", "rand_m3_n7"):
    g, onet, nse, prior, fn, ptype, kw = _setup(name, dev)
    errs = []
    for path in (2, 1):
        abi.set_option("path", path)
        e = []
        for k, t in enumerate(g["t_grid"]):
            try:
                out = nse.factorized_score(_t(g["theta"], dev), _t(g["x"], dev), torch.tensor(t), fn, prior, dist_cov_est=_t(g["dist_cov_est"], dev), cov_mode="GAUSS", prior_type=ptype, **kw)
                e.append(orc.rel_l2(out.cpu().numpy(), g["f64.GAUSS.factorized_score"][k]))
            except Exception as ex:
                e.append(float("nan"))
        errs.append(e)
    abi.set_option("path", 0)
    print(name, "tc:", " ".join(f"{v:.1e}" for v in errs[0]), "| simt:", " ".join(f"{v:.1e}" for v in errs[1]))
