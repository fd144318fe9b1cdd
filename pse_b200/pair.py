"""Pair-force providers on the engine's neighbour list.

The reference integrates whatever HOOMD's pair potentials accumulated in `net_force` before
`Stokes::integrateStepOne` (PSEv1/Stokes.cc:447,457); the example script runs without any
(examples/run.py).  These classes stand in for `hoomd.md.pair.lj` / `hoomd.md.pair.dpd_conservative` of
HOOMD 2.3.3 for the standalone `pse_b200.system.System`: constructing one registers it with the current
system, and every step the integrator's engine evaluates it into `net_force` (x, y, z, energy)."""
from . import _lib
from . import system as _system


class _pair:
    kind = None

    def __init__(self, r_cut=0.0, epsilon=1.0, sigma=2.0, system=None):
        self.r_cut, self.epsilon, self.sigma = float(r_cut), float(epsilon), float(sigma)
        self.system = system if system is not None else _system.current()
        self.system.forces.append(self)
        self.enabled = True

    def disable(self):
        self.enabled = False

    def enable(self):
        self.enabled = True

    def compute(self, engine, pos, out, accumulate):
        return engine.pair_force(pos, self.kind, self.epsilon, self.sigma, self.r_cut, out=out, accumulate=accumulate)


class lj(_pair):
    """U = 4 eps [(sigma/r)^12 - (sigma/r)^6] for r < r_cut (no shift), `hoomd.md.pair.lj` default mode."""
    kind = _lib.PSE_PAIR_LJ

    def __init__(self, r_cut, epsilon=1.0, sigma=2.0, system=None):
        if not r_cut > 0:
            raise RuntimeError("pair.lj: r_cut must be positive")
        super().__init__(r_cut, epsilon, sigma, system)


class wca(_pair):
    """Purely repulsive LJ: cut at 2^(1/6) sigma and shifted by +eps.  sigma = 2a = 2 by default (touching spheres)."""
    kind = _lib.PSE_PAIR_WCA

    def __init__(self, epsilon=1.0, sigma=2.0, system=None):
        super().__init__(0.0, epsilon, sigma, system)


class harmonic(_pair):
    """F = A (1 - r / r_cut) rhat (`hoomd.md.pair.dpd_conservative`): a soft contact repulsion, r_cut = 2a by default."""
    kind = _lib.PSE_PAIR_HARMONIC

    def __init__(self, A=1.0, r_cut=2.0, system=None):
        if not r_cut > 0:
            raise RuntimeError("pair.harmonic: r_cut must be positive")
        super().__init__(r_cut, A, 0.0, system)
