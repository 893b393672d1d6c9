"""Constants of the simulation -- same names, values and derivation order as the reference's
Constants.py:7-46 (imported everywhere there as `gv`).

The drop-in modules read these attributes at CALL time, exactly like the reference modules read
`gv.*`; when a module called `Constants` is already importable (the reference's own file next to
2DGrowth.py) the drop-ins use THAT one, so the driver's values flow through unchanged
(see runtime.constants()).

`configure()` is the programmatic form of the reference's six positional argv overrides
(Constants.py:7-13) and additionally exposes Hs, Ha, temperature and flux, which the reference
hard-codes (Constants.py:33,34,45,46).
"""
import sys

import numpy as np


def _derive(ns):
    # Constants.py:28-46, operation order kept (nbins_x = int(ceil(L/bin_size)) is fp-fragile)
    ns["E_as"] = np.sqrt(ns["E_a"] * ns["E_s"])
    ns["r_as"] = (ns["r_a"] + ns["r_s"]) / 2
    ns["sqrt3"] = np.sqrt(3)
    ns["L"] = ns["W"] * ns["r_s"]
    ns["Dmin"] = -(ns["Hs"]) * ns["r_s"] * ns["sqrt3"] / 2
    ns["Dmax"] = ns["Ha"] * ns["r_s"]
    ns["D"] = ns["Dmax"] - ns["Dmin"]
    ns["bin_size"] = 3 * ns["r_s"]
    ns["nbins_x"] = int(np.ceil(ns["L"] / ns["bin_size"]))
    ns["nbins_y"] = int(np.ceil(ns["D"] / ns["bin_size"]))
    ns["boltzmann"] = 8.617e-5
    ns["beta"] = 1.0 / ns["temperature"] / ns["boltzmann"] / 4
    return ns


def configure(E_a=2.083, E_s=2.083, r_a=2.7, r_s=2.7, W=18, folder="_Images/", Hs=10, Ha=10, temperature=300,
              deposition_rate=1.0, target=None):
    """Set the module attributes (of this module, or of `target`) from the primary constants."""
    ns = dict(E_a=float(E_a), E_s=float(E_s), r_a=float(r_a), r_s=float(r_s), W=int(W), folder=folder,
              Hs=int(Hs), Ha=int(Ha), temperature=temperature, deposition_rate=float(deposition_rate))
    _derive(ns)
    mod = target if target is not None else sys.modules[__name__]
    for k, v in ns.items():
        setattr(mod, k, v)
    return mod


# Constants.py:7-25: six positional overrides or the Cu-Cu defaults
if len(sys.argv) > 6 and __name__ == "Constants":
    configure(float(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
else:
    configure()
