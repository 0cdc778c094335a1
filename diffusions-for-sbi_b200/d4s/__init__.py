"""diffusions-for-sbi_b200: B200-native (sm_100a) tall-data diffusion posterior sampler.

Host-side mirror of the reference's sampling API (`NSE.ddim`, `NSE.factorized_score`, `PF_NSE`,
`tweedies_approximation`, `vp_diffused_priors`) over the C ABI of libd4s.so.  No CPU fallback.
"""
from . import _cabi  # noqa: F401
from .nse import NSE, NSELoss  # noqa: F401
from .pf_nse import PF_NSE  # noqa: F401
from .tall_posterior_sampler import (diffused_tall_posterior_score, euler_sde_sampler, mean_backward,  # noqa: F401
                                     prec_matrix_backward, sigma_backward, tweedies_approximation,
                                     tweedies_approximation_prior)
from .vp_diffused_priors import get_vpdiff_gaussian_score, get_vpdiff_uniform_score  # noqa: F401

__all__ = ["NSE", "NSELoss", "PF_NSE", "tweedies_approximation", "tweedies_approximation_prior", "prec_matrix_backward",
           "mean_backward", "sigma_backward",
           "diffused_tall_posterior_score", "euler_sde_sampler",
           "get_vpdiff_gaussian_score", "get_vpdiff_uniform_score"]
