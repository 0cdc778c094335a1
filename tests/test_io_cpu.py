"""Loading the reference's shipped pickles into the d4s classes (needs /root/reference; skipped where it is absent,
e.g. on the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import golden_io as gio

REF = os.environ.get("D4S_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "results")), reason="reference checkout not present")


def test_load_shipped_nse_matches_golden_weights():
    import d4s
    from d4s import io
    p = os.path.join(REF, "results/sbibm/slcp/n_train_30000_bs_256_n_epochs_5000_lr_0.0001/score_network.pkl")
    net = io.load_score_network(p)
    assert type(net) is d4s.NSE and net.theta_emb_dim == 5 and net.x_emb_dim == 8
    g = gio.load("slcp_n6")
    lin = [m for m in net.net if isinstance(m, torch.nn.Linear)]
    for i, l in enumerate(lin):
        assert np.array_equal(l.weight.detach().numpy(), g[f"net.W{i}"])
    # the torch definition of the module reproduces the reference's scores (fp32)
    th, x = torch.as_tensor(g["theta"]), torch.as_tensor(g["x"])
    s = net.score(th[:, None], x[None, :], torch.tensor(g["t_grid"][2])).detach().numpy()
    assert np.abs(s - g["f32.scores"][2]).max() < 1e-4 * np.abs(g["f32.scores"][2]).max()


def test_load_shipped_pf_nse_and_jrnnm():
    import d4s
    from d4s import io
    p = os.path.join(REF, "results/sbibm/slcp/n_train_10000_bs_256_n_epochs_5000_lr_0.0001/pf_n_max_3/score_network.pkl")
    net = io.load_score_network(p)
    assert type(net) is d4s.PF_NSE and net.n_max == 3
    p = os.path.join(REF, "results/jrnnm/4d/n_train_50000_n_epochs_5000_lr_0.001/score_network.pkl")
    net = io.load_score_network(p)
    assert type(net) is d4s.NSE and net.theta_emb_dim == 32 and net.x_emb_dim == 32


def test_dropin_directory_resolves_reference_module_names():
    code = ("import sys; sys.path.insert(0, %r); import nse, pf_nse, tall_posterior_sampler, vp_diffused_priors, nse_toy, "
            "embedding_nets, torch; import d4s; assert nse.NSE is d4s.NSE and pf_nse.PF_NSE is d4s.PF_NSE; "
            "net = torch.load(%r, weights_only=False, map_location='cpu'); assert type(net).__module__ == 'd4s.nse', type(net); "
            "print('DROPIN_OK')") % (os.path.join(ROOT, "diffusions-for-sbi_b200", "dropin"),
                                    os.path.join(REF, "results/sbibm/sir/n_train_30000_bs_256_n_epochs_5000_lr_0.0001/score_network.pkl"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "DROPIN_OK" in r.stdout, r.stderr[-2000:]
