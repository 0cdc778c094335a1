class FreeFormJacobianTransform:  # only named by NSE.flow, which is off the sampling path
    def __init__(self, *a, **k):
        raise NotImplementedError("FreeFormJacobianTransform is out of scope of the shim")
