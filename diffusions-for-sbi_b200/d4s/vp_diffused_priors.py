"""Closed-form scores of VP-diffused priors, API of the reference's `vp_diffused_priors.py`.

`get_vpdiff_uniform_score(a, b, nse)` / `get_vpdiff_gaussian_score(mean, cov, nse)` return callables
`fn(theta, t) -> [N, m]` like the reference's closures (vp_diffused_priors.py:5-74), but as objects that also
expose `.kind` and the prior parameters so that the tall sampler can evaluate them inside its fused kernel.
Calling the object evaluates the score on the GPU through libd4s (kernel 3 with zero observations).
"""
import torch

from . import _cabi as abi
from . import engine


class _DiffusedPriorScore:
    kind = None

    def __init__(self, nse):
        self.nse = nse

    def _spec(self) -> "engine.PriorSpec":
        raise NotImplementedError

    def __call__(self, theta: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        lead = theta.shape[:-1]
        th = theta.reshape(-1, theta.shape[-1])
        out = engine.prior_score_native(self.nse, th, t, self._spec())
        return out.reshape(*lead, theta.shape[-1])


class UniformDiffusedScore(_DiffusedPriorScore):
    """score_d = f'_d / (f_d + 1e-6), f_d = [Phi(u_b) - Phi(u_a)] / sqrt(alpha)   (vp_diffused_priors.py:29-46)."""
    kind = "uniform"

    def __init__(self, a, b, nse):
        super().__init__(nse)
        self.low, self.high = torch.as_tensor(a), torch.as_tensor(b)

    def _spec(self):
        return engine.PriorSpec(abi.PRIOR_UNIFORM, low=engine._cpu64(self.low).reshape(-1),
                                high=engine._cpu64(self.high).reshape(-1))


class GaussianDiffusedScore(_DiffusedPriorScore):
    """score = -(theta - sqrt(alpha) mu) (sigma^2 I + alpha Sigma)^-1   (vp_diffused_priors.py:55-72)."""
    kind = "gaussian"

    def __init__(self, mean, cov, nse):
        super().__init__(nse)
        self.mean, self.cov = torch.as_tensor(mean), torch.as_tensor(cov)

    def _spec(self):
        m = self.mean.numel()
        return engine.PriorSpec(abi.PRIOR_GAUSSIAN, mean=engine._cpu64(self.mean).reshape(m),
                                cov=engine._cpu64(self.cov).reshape(m, m))


def get_vpdiff_uniform_score(a, b, nse):
    return UniformDiffusedScore(a, b, nse)


def get_vpdiff_gaussian_score(mean, cov, nse):
    return GaussianDiffusedScore(mean, cov, nse)
