"""GPU: the plugin-facing Python API (`integrate.PSEv1`, shear functions, box_resize variant) drives the same C ABI
calls as the raw engine — examples/run.py-shaped run (simple cubic lattice of 1000 spheres, L = 64, sine shear)."""
import math

import numpy as np
import pytest

import pse_b200 as PSEv1
from pse_b200 import engine as E
from tests import util

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


def _pair_oracle(pos, L, xy, kind, eps, sigma, rcut):
    """O(N^2) numpy pair forces with HOOMD's minimum-image convention (float64): (F, per-particle energy)."""
    x = pos[:, :3].astype(np.float64)
    d = x[:, None, :] - x[None, :, :]
    img = np.rint(d[..., 2] / L); d[..., 2] -= L * img
    img = np.rint(d[..., 1] / L); d[..., 1] -= L * img; d[..., 0] -= L * xy * img
    img = np.rint(d[..., 0] / L); d[..., 0] -= L * img
    r2 = (d * d).sum(-1)
    np.fill_diagonal(r2, np.inf)
    if kind == "wca":
        rcut = 2.0 ** (1.0 / 6.0) * sigma
    m = r2 < rcut * rcut
    r2m = np.where(m, r2, 1.0)
    if kind == "harmonic":
        r = np.sqrt(r2m)
        fr = eps * (1.0 / r - 1.0 / rcut); u = eps * (rcut - r) - eps * (rcut * rcut - r2m) / (2 * rcut)
    else:
        s6 = (sigma * sigma / r2m) ** 3
        fr = 24 * eps * s6 * (2 * s6 - 1) / r2m; u = 4 * eps * s6 * (s6 - 1) + (eps if kind == "wca" else 0.0)
    fr = np.where(m, fr, 0.0); u = np.where(m, u, 0.0)
    return (fr[..., None] * d).sum(1), 0.5 * u.sum(1)


@pytest.mark.parametrize("xy", [0.0, 0.3])
def test_pair_forces_match_numpy(cuda, xy):
    """pse_pair_force (stand-in for the HOOMD pair potentials behind net_force) against an O(N^2) sum."""
    import torch
    from pse_b200 import _lib
    from pse_b200 import engine as E
    N, L = 1500, util.box_length(1500, 0.25)
    pos_np = util.lattice_positions(N, L, 3)
    pos = torch.from_numpy(pos_np).cuda()
    eng = E.Engine(E.make_config(N, L, xy=xy, T=1.0, dt=1e-3, seed=1))
    for kind, code, eps, sigma, rcut in (("wca", _lib.PSE_PAIR_WCA, 1.3, 2.7, 0.0), ("lj", _lib.PSE_PAIR_LJ, 0.7, 1.9, 4.5),
                                         ("harmonic", _lib.PSE_PAIR_HARMONIC, 25.0, 0.0, 2.6)):
        F = eng.pair_force(pos, code, epsilon=eps, sigma=sigma, r_cut=rcut).cpu().numpy().astype(np.float64)
        Fo, Uo = _pair_oracle(pos_np, L, xy, kind, eps, sigma, rcut)
        assert np.abs(Fo).max() > 0
        assert np.abs(F[:, :3] - Fo).max() < 2e-5 * np.abs(Fo).max(), kind
        assert np.abs(F[:, 3] - Uo).max() < 1e-4 * max(np.abs(Uo).max(), 1e-30), kind   # fp32 sum of mixed-sign pair energies
        assert np.abs(F[:, :3].sum(0)).max() < 1e-3 * np.abs(Fo).max()       # Newton's third law on a full list
    with pytest.raises(E.PSEError):
        eng.pair_force(pos, _lib.PSE_PAIR_LJ, r_cut=50.0)                      # beyond the list


def test_pair_provider_drives_a_run_and_restart_roundtrip(cuda, tmp_path):
    """A WCA suspension run through the HOOMD-shaped API, saved and reloaded: the reloaded system continues where the original would have."""
    import torch
    import pse_b200 as PSEv1
    N, L = 2000, util.box_length(2000, 0.3)

    def make(system):
        PSEv1.system.set_current(system)
        PSEv1.integrate.mode_standard(dt=1e-3)
        PSEv1.pair.wca(epsilon=1.0, sigma=2.6)   # wider than the spheres so that the jittered lattice feels it from step 0
        return PSEv1.integrate.PSEv1(group=system.all(), seed=5, T=1.0, xi=0.5, error=1e-3)

    s = PSEv1.system.System(util.lattice_positions(N, L, 1), PSEv1.system.Box(L))
    make(s)
    s.run(5)
    assert float(s.net_force[:, :3].abs().max()) > 0                            # the provider filled net_force
    s.save(tmp_path / "restart")
    s.run(3)
    s2 = PSEv1.system.System.load(tmp_path / "restart")
    assert s2.timestep == 5
    pse2 = make(s2)
    pse2.cpp_method.lanczos_m = s2.restart_lanczos_m
    s2.run(3)
    # (bit for bit with PSE_WAVE=v1; the default spreading merges its windows with float reductions whose order across
    # blocks is not fixed, so two runs agree to round-off)
    d = (s.pos[:, :3] - s2.pos[:, :3]).abs()
    d = torch.minimum(d, L - d)
    assert float(d.max()) < 2e-5
