import torch
from torch.distributions import Independent, Uniform


class BoxUniform(Independent):
    def __init__(self, low, high, reinterpreted_batch_ndims=1, device=None):
        super().__init__(Uniform(torch.as_tensor(low, dtype=torch.float32),
                                 torch.as_tensor(high, dtype=torch.float32), validate_args=False),
                         reinterpreted_batch_ndims)
