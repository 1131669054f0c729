"""Import shim (oracle scaffolding only). Starfish.transforms.extinct is only reached when the
model has an "Av" parameter; the synthetic oracle models never set it.  A_lambda = 0 is exact for Av=0."""
import numpy as np


def _zero(wave, a_v, r_v=3.1, unit="aa"):
    if a_v != 0:
        raise NotImplementedError("extinction shim only supports Av == 0")
    return np.zeros_like(np.asarray(wave, dtype=float))


ccm89 = odonnell94 = calzetti00 = fitzpatrick99 = _zero


def fm07(wave, a_v, unit="aa"):
    return _zero(wave, a_v)
