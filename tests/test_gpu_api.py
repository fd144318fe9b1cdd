"""GPU: the plugin-facing Python API (`integrate.PSEv1`, shear functions, box_resize variant) drives the same C ABI
calls as the raw engine — examples/run.py-shaped run (simple cubic lattice of 1000 spheres, L = 64, sine shear)."""
import math

import numpy as np
import pytest

import pse_b200 as PSEv1
from pse_b200 import engine as E

pytestmark = pytest.mark.gpu


def _sc_lattice(n, L):
    a = L / n
    g = (np.arange(n) + 0.5) * a - L / 2
    return np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)


def test_runpy_shaped_simulation_matches_raw_engine(cuda):
    import torch
    N, L, dt = 1000, 64.0, 1e-3
    pos = _sc_lattice(10, L)
    s = PSEv1.system.set_current(PSEv1.system.System(pos, PSEv1.system.Box(L)))
    PSEv1.integrate.mode_standard(dt=dt)
    ff = PSEv1.shear_function.sine(dt=dt, shear_rate=1.0, shear_freq=1.0)
    pse = PSEv1.integrate.PSEv1(group=s.all(), seed=1, T=1.0, xi=0.5, error=1e-3, function_form=ff)
    assert (pse.cpp_method.params.Nx, pse.cpp_method.params.P) == (64, 6)        # SURVEY.md §8 row 1'
    assert pse.rcut == pytest.approx(math.sqrt(-math.log(1e-3)) / 0.5)
    PSEv1.system.box_resize(s, xy=PSEv1.variant.shear_variant(ff, total_timestep=100))
    s.run(5)
    assert s.timestep == 5 and bool(torch.isfinite(s.pos).all())
    # the same five steps through the raw engine
    cfg = E.make_config(N, L, T=1.0, dt=dt, seed=1)
    eng = E.Engine(cfg)
    p4 = np.zeros((N, 4), dtype=np.float32); p4[:, :3] = pos
    p = torch.from_numpy(p4).cuda(); im = torch.zeros((N, 3), dtype=torch.int32, device="cuda"); F = torch.zeros((N, 4), device="cuda")
    for t in range(5):
        eng.set_tilt(ff.get_strain(t) - math.floor(ff.get_strain(t) + 0.5))      # wrapped strain in [-0.5, 0.5)
        eng.set_temperature(1.0)
        eng.step(p, im, F, t, shear_rate=ff.get_shear_rate(t))
    assert float((p - s.pos).abs().max()) < 1e-5
    assert float((s.pos[:, :3] - torch.from_numpy(pos).cuda()).abs().max()) > 1e-3   # the particles did move
    # set_params / stop_shear keep the reference's signatures
    pse.set_params(T=0.5)
    pse.stop_shear()
    s.run(1)
    assert pse.function_form.get_shear_rate(6) == 0


def test_temperature_variant_and_zero_T(cuda):
    import torch
    N, L = 512, 40.0
    pos = _sc_lattice(8, L)
    s = PSEv1.system.set_current(PSEv1.system.System(pos, PSEv1.system.Box(L)))
    PSEv1.integrate.mode_standard(dt=1e-3)

    class Ramp:  # a variant: T(t)
        def get_value(self, t): return 0.0 if t < 2 else 1.0
    pse = PSEv1.integrate.PSE(group=s.all(), T=Ramp(), seed=3)
    before = s.pos.clone()
    s.run(2)                                   # T = 0 and F = 0: nothing moves
    assert torch.equal(before, s.pos)
    s.run(1)                                   # T = 1: Brownian motion
    assert float((s.pos - before).abs().max()) > 1e-3
    with pytest.raises(RuntimeError):
        PSEv1.integrate.PSEv1(group=s.all(), T=1.0, nlist_type="bogus")
