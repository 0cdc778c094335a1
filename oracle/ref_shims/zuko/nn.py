import torch.nn as tnn


class Linear(tnn.Linear):
    """zuko.nn.Linear: a plain affine layer (extra zuko kwargs are not used by the reference)."""


class MLP(tnn.Sequential):
    """zuko.nn.MLP(in, out, hidden_features=(64, 64), activation=None -> ReLU, normalize=False).

    Linear/activation stack with no trailing activation.  Shipped pickles restore `_modules`
    directly, so this constructor only matters for freshly built nets.
    """

    def __init__(self, in_features, out_features, hidden_features=(64, 64), activation=None,
                 normalize=False, **kwargs):
        if activation is None:
            activation = tnn.ReLU
        layers = []
        widths = [in_features, *hidden_features, out_features]
        for a, b in zip(widths[:-1], widths[1:]):
            layers.append(Linear(a, b, **kwargs))
            if normalize:
                layers.append(tnn.LayerNorm(b))
            layers.append(activation())
        layers = layers[:-2] if normalize else layers[:-1]
        super().__init__(*layers)
        self.in_features = in_features
        self.out_features = out_features
