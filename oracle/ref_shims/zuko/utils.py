import torch


def broadcast(*tensors, ignore=0):
    """Broadcast all but the last `ignore` dims of every tensor to a common batch shape."""
    if isinstance(ignore, int):
        ignore = [ignore] * len(tensors)
    batch = [t.shape[: t.dim() - i] for t, i in zip(tensors, ignore)]
    common = torch.broadcast_shapes(*batch)
    out = []
    for t, i in zip(tensors, ignore):
        out.append(t.expand(common + t.shape[t.dim() - i:]))
    return tuple(out)
