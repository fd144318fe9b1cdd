"""`integrate.PSEv1` — the plugin's integrator with the reference's signature and behaviour
(PSEv1/integrate.py:15-123), running on the B200 engine behind the C ABI instead of HOOMD + `_PSEv1.Stokes`.

    import pse_b200 as PSEv1
    sysdef = PSEv1.system.set_current(PSEv1.system.System(positions, PSEv1.system.Box(L)))
    PSEv1.integrate.mode_standard(dt=1e-3)
    pse = PSEv1.integrate.PSEv1(group=sysdef.all(), seed=1, T=1.0, xi=0.5, error=1e-3, function_form=ff)
    sysdef.run(1000)
"""
import math

from . import engine as _engine
from . import shear_function
from . import system as _system


def mode_standard(dt):
    """`hoomd.md.integrate.mode_standard(dt)`: sets the time step of the current system."""
    _system.current().dt = float(dt)


class _constant_variant:
    def __init__(self, v):
        self.v = float(v)

    def getValue(self, timestep):
        return self.v


def _setup_variant_input(T):
    """`hoomd.variant._setup_variant_input`: a number or anything with get_value/getValue(timestep)."""
    if hasattr(T, "getValue"):
        return T
    if hasattr(T, "get_value"):
        class _W:
            def __init__(self, t): self.t = t
            def getValue(self, ts): return self.t.get_value(ts)
        return _W(T)
    return _constant_variant(T)


class PSEv1:
    """One-step overdamped integration with RPY hydrodynamic interactions.

    group, T, seed, xi, error, function_form, max_strain, nlist_type: as PSEv1/integrate.py:32.
    `nlist_type` ("cell" | "tree" | "stencil") selects a HOOMD builder in the reference; all three map to the
    engine's cell-list builder here, invalid names raise as in the reference (:76-78)."""

    def __init__(self, group, T, seed=0, xi=0.5, error=0.001, function_form=None, max_strain=0.5, nlist_type="cell"):
        sysdef = group.system
        if sysdef.dt is None:
            raise RuntimeError("Error creating Stokes: call integrate.mode_standard(dt) first")
        self.system = sysdef
        self.T = _setup_variant_input(T)
        self.rcut = math.sqrt(-math.log(error)) / xi  # integrate.py:47
        import torch
        if not torch.cuda.is_available():  # integrate.py:51-53
            raise RuntimeError("Error creating Stokes")
        if nlist_type.upper() not in ("CELL", "TREE", "STENCIL"):
            raise RuntimeError("Error constructing neighborlist")
        b = sysdef.box
        cfg = _engine.make_config(sysdef.N, (b.Lx, b.Ly, b.Lz), xi=xi, error=error, max_strain=max_strain,
                                  T=self.T.getValue(sysdef.timestep), dt=sysdef.dt, seed=seed, xy=b.xy, r_buff=0.4)
        self.cpp_method = _engine.Engine(cfg)          # Stokes(...) + setParams()
        self.max_strain = max_strain
        if function_form is not None:                  # setShear (integrate.py:90-94)
            self.function_form = function_form
        else:
            self.function_form = shear_function.steady(dt=0)
        sysdef.integrator = self

    # -- reference methods
    def set_params(self, T=None, function_form=None, max_strain=0.5):
        if T is not None:
            self.T = _setup_variant_input(T)
        if function_form is not None:
            self.function_form = function_form
            self.max_strain = max_strain

    def stop_shear(self, max_strain=0.5):
        self.function_form = shear_function.steady(dt=0)
        self.max_strain = max_strain

    # -- driven by System.run (HOOMD's IntegratorTwoStep::update)
    def set_tilt(self, xy):
        self.cpp_method.set_tilt(xy)

    def integrate_step(self, timestep):
        """Stokes::integrateStepOne (PSEv1/Stokes.cc:429-523); step two is empty (:528-530)."""
        s = self.system
        self.cpp_method.set_temperature(self.T.getValue(timestep))
        rate = self.function_form.get_shear_rate(timestep)
        return self.cpp_method.step(s.pos, s.image, s.net_force, timestep, shear_rate=rate, vel=s.vel)


PSE = PSEv1  # BASELINE.json spells the class `PSEv1.integrate.PSE`
