"""pse_b200 — B200-native Positively Split Ewald (PSE) Brownian dynamics.

Drop-in for the hot path of the HOOMD plugin stochasticHydroTools/PSE: `integrate.PSEv1` (alias
`integrate.PSE`), `shear_function.*`, `variant.shear_variant` keep the plugin's signatures
(PSEv1/integrate.py, PSEv1/shear_function.py, PSEv1/variant.py); HOOMD itself is replaced by the
minimal particle/box shim in `pse_b200.system`.  All numerics run in libpse_b200.so (CUDA, sm_100a)
behind the C ABI of include/pse_b200.h; there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the native library is missing)
from . import engine, integrate, pair, shear_function, system, variant  # noqa: F401
