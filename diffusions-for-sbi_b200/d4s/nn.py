"""Minimal network builders with the constructor signature the reference expects from `zuko.nn.MLP`
(nse.py:11,28: `build_net(in_features, out_features, **kwargs)`)."""
from typing import Sequence

import torch.nn as nn


class Linear(nn.Linear):
    pass


class MLP(nn.Sequential):
    """Linear -> ReLU -> ... -> Linear (no trailing activation)."""

    def __init__(self, in_features: int, out_features: int, hidden_features: Sequence[int] = (64, 64),
                 activation=None, normalize: bool = False, **kwargs):
        if normalize:
            raise NotImplementedError("normalize=True is not supported by the fused score kernel")
        act = nn.ReLU if activation is None else activation
        widths = [in_features, *hidden_features, out_features]
        mods = []
        for i, (a, b) in enumerate(zip(widths[:-1], widths[1:])):
            mods.append(Linear(a, b, **kwargs))
            if i < len(widths) - 2:
                mods.append(act())
        super().__init__(*mods)
        self.in_features, self.out_features = in_features, out_features
