"""GPU parity suite (-m gpu): the CUDA path behind the C ABI against
  (1) the reference's own CUDA kernels compiled unmodified (oracle/_ref/libpse_ref.so),
  (2) the CPU restatement (oracle/pse_oracle.c) and its dense double-precision Ewald sum,
  (3) size-independent properties at the BASELINE.json headline size (N = 1M).
Tolerances: integer outputs (neighbour list, particle->grid index) bit-exact; deterministic M.F and
Brownian displacements for identical random vectors 1e-5 relative (BASELINE.json north_star), measured
as |a-b|_2/|b|_2 and max|a-b|/max|b|; accuracy vs dense Ewald 3x the requested error."""
import math

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-5


class System:
    def __init__(self, N, L, xy=0.0, seed=0, lattice=False, ref_pi=True, error=1e-3, xi=0.5, T=1.0, dt=1e-3, want_ref=True):
        import torch
        from oracle import oraclewrap as O
        from oracle import refwrap
        from pse_b200 import _lib
        from pse_b200 import engine as E
        self.torch, self.E = torch, E
        self.N, self.L, self.T, self.dt = N, L, T, dt
        self.cfg = E.make_config(N, L, xy=xy, flags=_lib.PSE_FLAG_REF_PI if ref_pi else 0, T=T, dt=dt, seed=1, error=error, xi=xi)
        self.eng = E.Engine(self.cfg)
        self.p = self.eng.params
        Lx = L if np.isscalar(L) else L[0]
        self.pos_np = util.lattice_positions(N, Lx, seed) if lattice else util.random_positions(N, Lx, seed)
        if not np.isscalar(L):
            self.pos_np[:, 1] *= L[1] / L[0]; self.pos_np[:, 2] *= L[2] / L[0]
        self.F_np = util.random_forces(N, seed + 1)
        self.pos = torch.from_numpy(self.pos_np).cuda(); self.F = torch.from_numpy(self.F_np).cuda()
        self.eng.build_neighbors(self.pos)
        self.nl = self.eng.neighbor_list()
        self.orc = O.Oracle(N, L, xi=xi, error=error, xy=xy, ref_pi=ref_pi)
        self.ref = None
        if want_ref and refwrap.available():
            self.ref = refwrap.Reference(self.cfg, self.p, E.ewald_table(self.cfg))
            self.ref.set_neighbors(*self.nl)

    def nl_np(self):
        return [t.cpu().numpy().view(np.uint32) for t in self.nl]

    def noise(self, seed=5, neutralise=True):
        torch, p = self.torch, self.p
        g = torch.Generator(device="cuda"); g.manual_seed(seed)
        G = p.Nx * p.Ny * p.Nz
        up = torch.rand((self.N, 3), device="cuda", generator=g); ug = torch.rand((G, 6), device="cuda", generator=g)
        if neutralise:  # nodes the reference visits twice with a data race: u = 0.5 -> exactly 0
            v = ug.view(p.Nx, p.Ny, p.Nz, 6)
            if p.Nz % 2 == 0: v[:, :, p.Nz // 2, :] = 0.5
            if p.Ny % 2 == 0: v[:, p.Ny // 2, 0, :] = 0.5
        return up, ug


def close(a, b, tol=TOL):
    l2, mx = util.rel_err(a.detach().cpu().numpy() if hasattr(a, "detach") else a, b.detach().cpu().numpy() if hasattr(b, "detach") else b)
    assert l2 < tol and mx < tol, (l2, mx)


@pytest.fixture(scope="module", params=[0.0, 0.3], ids=["ortho", "sheared"])
def cfg1(request, cuda):
    """BASELINE.json config 1: N = 1000, phi = 0.1, error 1e-3 (36^3 grid), i.i.d. positions incl. overlaps."""
    return System(1000, util.box_length(1000, 0.1), xy=request.param)


# ---------------------------------------------------------------- bit-exact integer outputs
def test_neighbor_list_bit_exact_vs_bruteforce(cfg1):
    nn, head, nl = cfg1.nl_np()
    onn, ohead, onl = cfg1.orc.neighbors(cfg1.pos_np, cfg1.p.rcut + 0.4, brute=True)
    assert np.array_equal(nn, onn) and np.array_equal(head, ohead) and np.array_equal(nl, onl)


def test_grid_index_bit_exact(cfg1):
    gi = cfg1.eng.grid_index(cfg1.pos).cpu().numpy()
    assert np.array_equal(gi, cfg1.orc.grid_index(cfg1.pos_np))
    _check_index_against_reference_spread(cfg1, gi)


def _check_index_against_reference_spread(s, gi):
    """The reference has no index output; its Spread kernel with prefac = 1, expfac = 0 deposits exactly
    F on each of the P^3 support nodes (PSEv1/Mobility.cu:241-246), i.e. an integer histogram."""
    if s.ref is None:
        pytest.skip("reference library not built")
    torch, p, N = s.torch, s.p, s.N
    ids = (np.arange(N) % 1021 + 1).astype(np.float64)
    Fh = torch.zeros_like(s.F); Fh[:, 0] = 1.0; Fh[:, 1] = torch.from_numpy(ids).float().cuda()
    gX, gY, _ = s.ref.spread(s.pos, Fh, P=p.P, prefac=1.0, expfac=0.0)
    hist = np.zeros(p.Nx * p.Ny * p.Nz); hid = np.zeros_like(hist)
    g = gi.astype(np.int64)
    for tx in range(p.P):
        for ty in range(p.P):
            for tz in range(p.P):
                lin = (((g[:, 0] + tx) % p.Nx) * p.Ny + (g[:, 1] + ty) % p.Ny) * p.Nz + (g[:, 2] + tz) % p.Nz
                hist += np.bincount(lin, minlength=hist.size); hid += np.bincount(lin, weights=ids, minlength=hist.size)
    assert np.array_equal(hist, gX[:, 0].cpu().numpy()) and np.array_equal(hid, gY[:, 0].cpu().numpy())


# ---------------------------------------------------------------- deterministic operators
def test_mreal_parity(cfg1):
    U = cfg1.eng.mreal(cfg1.pos, cfg1.F)
    cfg1.orc.set_neighbors(*cfg1.nl_np())
    close(U, cfg1.orc.mreal(cfg1.pos_np, cfg1.F_np))
    if cfg1.ref: close(U, cfg1.ref.mreal(cfg1.pos, cfg1.F))
    assert float(U[:, 3].abs().max()) == 0.0  # .w written as 0 (PSEv1/Mobility.cu:632)


def test_mwave_parity(cfg1):
    U = cfg1.eng.mwave(cfg1.pos, cfg1.F)
    close(U, cfg1.orc.mwave(cfg1.pos_np, cfg1.F_np))
    if cfg1.ref: close(U, cfg1.ref.mwave(cfg1.pos, cfg1.F))


def test_mobility_parity_and_repeatability(cfg1):
    U = cfg1.eng.mobility(cfg1.pos, cfg1.F)
    if cfg1.ref: close(U, cfg1.ref.mobility(cfg1.pos, cfg1.F))
    close(U, cfg1.eng.mreal(cfg1.pos, cfg1.F) + cfg1.eng.mwave(cfg1.pos, cfg1.F), 2e-7)
    close(cfg1.eng.mobility(cfg1.pos, cfg1.F), U, 1e-6)


@pytest.mark.parametrize("xy", [0.0, 0.3])
def test_mobility_within_ewald_error_of_dense(cuda, xy):
    from oracle import oraclewrap as O
    N = 400
    L = util.box_length(N, 0.1)
    for error in (1e-3, 1e-4):
        s = System(N, L, xy=xy, seed=3, ref_pi=False, error=error, want_ref=False)
        Ud = O.dense_mobility(s.pos_np[:, :3], s.F_np[:, :3], L, xy=xy)
        U = s.eng.mobility(s.pos, s.F).cpu().numpy()[:, :3].astype(np.float64)
        err = np.linalg.norm(U - Ud) / np.linalg.norm(Ud)
        assert err < 3 * error, (error, err)
        # the reference's 2*pi typo moves the answer by more than the parity tolerance but far less than `error`
        s2 = System(N, L, xy=xy, seed=3, ref_pi=True, error=error, want_ref=False)
        U2 = s2.eng.mobility(s2.pos, s2.F).cpu().numpy()[:, :3].astype(np.float64)
        d = np.linalg.norm(U2 - U) / np.linalg.norm(U)
        assert 1e-6 < d < 3e-4, d


# ---------------------------------------------------------------- Brownian parts, identical random vectors
def test_velocity_parity_injected_noise(cfg1):
    s = cfg1
    up, ug = s.noise()
    psi = s.torch.zeros_like(s.F); psi[:, :3] = s.torch.from_numpy((up.cpu().numpy() * np.float32(2 * 1.73205080757) - np.float32(1.73205080757))).cuda()
    s.orc.set_neighbors(*s.nl_np())
    # real-space Brownian term alone (Lanczos), same starting m on all sides
    s.eng.lanczos_m = 2
    Ue, m = s.eng.velocity(s.pos, s.F, timestep=3, u_particles=up, u_grid=ug, parts=4)
    psi_np = np.zeros((s.N, 4), dtype=np.float32); a = np.float32(1.73205080757)
    psi_np[:, :3] = np.float32(2) * a * up.cpu().numpy() - a
    Uo, mo, _ = s.orc.lanczos(s.pos_np, psi_np, s.T, s.dt, m_in=2)
    assert abs(m - mo) <= 1
    close(Ue, Uo, 1e-3 if m != mo else 2e-5)   # one extra iteration changes the result at the `error` level
    # wave-space Brownian term alone
    Uw, _ = s.eng.velocity(s.pos, s.F, timestep=3, u_particles=up, u_grid=ug, parts=2)
    nf = math.sqrt(2.0 * s.T / s.dt / s.p.quadW)
    close(Uw, s.orc.mwave(s.pos_np, s.F_np, do_det=False, u_grid=ug.cpu().numpy(), noise_fac=nf))
    if s.ref:
        s.ref.set_noise_tables(up, ug)
        try:
            s.eng.lanczos_m = 2; s.ref.m_lanczos = 2
            Ue, m = s.eng.velocity(s.pos, s.F, timestep=3, u_particles=up, u_grid=ug, parts=7)
            Ur = s.ref.velocity(s.pos, s.F, s.T, s.dt, 3)
            assert m == s.ref.m_lanczos
            close(Ue, Ur)
        finally:
            s.ref.set_noise_tables(None, None)


def test_zero_temperature_skips_brownian(cfg1):
    s = cfg1
    s.eng.set_temperature(0.0)
    try:
        U, _ = s.eng.velocity(s.pos, s.F, timestep=1, parts=7)  # PSEv1/Brownian.cu:855,885
        close(U, s.eng.mobility(s.pos, s.F), 1e-6)  # spreading atomics make repeated runs differ at 1e-7
    finally:
        s.eng.set_temperature(s.T)


def test_step_parity_engine_rng_odd_grid(cuda):
    """Three full BD steps with the engine's own Philox streams against gpu_stokes_step_one: 75^3 grid (odd sizes
    have no doubly-visited nodes), steady shear rate 0.5, tilted box."""
    import torch
    s = System(20000, 77.0, xy=0.1, seed=2, lattice=True)
    if s.ref is None:
        pytest.skip("reference library not built")
    assert s.p.Nx == 75
    pe, pr = s.pos.clone(), s.pos.clone()
    ie = torch.zeros((s.N, 3), dtype=torch.int32, device="cuda"); ir = ie.clone()
    vel = torch.zeros_like(s.F); vel[:, 3] = 1.0
    acc = torch.zeros((s.N, 3), device="cuda")
    s.eng.lanczos_m = 2; s.ref.m_lanczos = 2
    for t in range(3):
        s.eng.build_neighbors(pr); s.ref.set_neighbors(*s.eng.neighbor_list())
        s.ref.step(pr, vel, acc, ir, s.F, s.T, s.dt, t, shear_rate=0.5)
        s.eng.step(pe, ie, s.F, t, shear_rate=0.5)
        assert s.eng.lanczos_m == s.ref.m_lanczos
        assert torch.equal(ie, ir)
        assert float((pe - pr).abs().max()) < 2e-5 * 1.0   # positions O(40), per-step displacement O(0.1)
    assert float((pe - s.pos).abs().max()) > 1e-2


def test_brownian_covariance_is_mobility(cuda):
    """<u u^T> dt / (2T) -> M over many seeds (fluctuation-dissipation), small system, engine RNG."""
    import torch
    N = 12
    L = 13.0
    s = System(N, L, seed=4, want_ref=False, ref_pi=False)
    Z = torch.zeros_like(s.F)
    M = np.zeros((3 * N, 3 * N))
    for c in range(3 * N):
        e = torch.zeros_like(s.F); e[c // 3, c % 3] = 1
        M[:, c] = s.eng.mobility(s.pos, e).cpu().numpy()[:, :3].reshape(-1)
    nsamp = 6000
    C = np.zeros_like(M)
    us = []
    for t in range(nsamp):
        U, _ = s.eng.velocity(s.pos, Z, timestep=t, parts=6)
        us.append(U[:, :3].reshape(-1))
    Uall = torch.stack(us).double().cpu().numpy()
    C = Uall.T @ Uall / nsamp * s.dt / (2 * s.T)
    assert abs(np.trace(C) / np.trace(M) - 1) < 0.03
    assert np.linalg.norm(C - M) / np.linalg.norm(M) < 0.12   # ~ sqrt(2*dim/nsamp) sampling error
    assert abs(Uall.mean()) < 5 * Uall.std() / math.sqrt(Uall.size)


# ---------------------------------------------------------------- edge cases
@pytest.mark.parametrize("error,P", [(1e-2, 4), (3e-3, 5), (1e-4, 8)])
def test_other_support_sizes_and_noncubic_box(cuda, error, P):
    s = System(600, (40.0, 36.0, 44.0), xy=0.2, seed=6, error=error)
    assert s.p.P == P
    nn, head, nl = s.nl_np()
    onn, ohead, onl = s.orc.neighbors(s.pos_np, s.p.rcut + 0.4, brute=True)
    assert np.array_equal(nn, onn) and np.array_equal(nl, onl)
    gi = s.eng.grid_index(s.pos).cpu().numpy()
    assert np.array_equal(gi, s.orc.grid_index(s.pos_np))
    _check_index_against_reference_spread(s, gi)
    U = s.eng.mobility(s.pos, s.F)
    if s.ref: close(U, s.ref.mobility(s.pos, s.F))
    s.orc.set_neighbors(nn, head, nl)
    close(U, s.orc.mreal(s.pos_np, s.F_np) + s.orc.mwave(s.pos_np, s.F_np))


def test_boundary_overlap_and_coincident_particles(cuda):
    import torch
    N, L = 64, 30.0
    s = System(N, L, seed=8)
    pos = s.pos_np.copy()
    h = np.float32(L / 2)
    pos[0, :3] = (-h, -h, -h)                    # exactly on the lower corner: fraction 0
    pos[1, :3] = (np.nextafter(h, np.float32(0)),) * 3   # last float below the upper corner
    pos[2, :3] = (1.0, 2.0, 3.0); pos[3, :3] = (1.0, 2.0, 3.0)          # coincident: r < dr, skipped (PSEv1/Mobility.cu:652)
    pos[4, :3] = (5.0, 5.0, 5.0); pos[5, :3] = (5.5, 5.0, 5.0)          # overlapping: r < 2a branch of the table
    pos[6, :3] = (h - np.float32(0.01), 0, 0); pos[7, :3] = (-h + np.float32(0.01), 0, 0)  # neighbours across the boundary
    p = torch.from_numpy(pos).cuda()
    s.eng.build_neighbors(p)
    nn, head, nl = [t.cpu().numpy().view(np.uint32) for t in s.eng.neighbor_list()]
    onn, ohead, onl = s.orc.neighbors(pos, s.p.rcut + 0.4, brute=True)
    assert np.array_equal(nn, onn) and np.array_equal(nl, onl)
    assert 7 in nl[head[6]: head[6] + nn[6]] and 3 in nl[head[2]: head[2] + nn[2]]
    assert np.array_equal(s.eng.grid_index(p).cpu().numpy(), s.orc.grid_index(pos))
    U = s.eng.mobility(p, s.F)
    assert bool(torch.isfinite(U).all())
    s.ref.set_neighbors(*s.eng.neighbor_list()) if s.ref else None
    if s.ref: close(U, s.ref.mobility(p, s.F))


def test_single_particle_periodic_self_mobility(cuda):
    import torch
    L = 20.0
    s = System(1, L, ref_pi=False, want_ref=False)
    pos = torch.zeros((1, 4), device="cuda"); F = torch.tensor([[1.0, 0, 0, 0]], device="cuda")
    U = s.eng.mobility(pos, F)
    assert abs(float(U[0, 0]) - (1 - 2.837297 / L + 4 * math.pi / 3 / L**3)) < 3e-3
    assert s.eng.stats()["nnz"] == 0


def test_stale_list_is_rebuilt_and_host_step_matches_device_step(cuda):
    import torch
    s = System(3000, util.box_length(3000, 0.15), seed=9, lattice=True, want_ref=False)
    b0 = s.eng.stats()["nlist_builds"]
    moved = s.pos.clone(); moved[:, 0] += 0.05
    s.eng.mreal(moved, s.F)
    assert s.eng.stats()["nlist_builds"] == b0            # within the buffer: list kept
    moved[:, 0] += 3.0; moved[:, 0] = (moved[:, 0] + s.L / 2) % s.L - s.L / 2
    far = moved.clone(); far[::2, 1] += 1.0; far[:, 1] = (far[:, 1] + s.L / 2) % s.L - s.L / 2
    U = s.eng.mreal(far, s.F)
    assert s.eng.stats()["nlist_builds"] == b0 + 1        # displaced beyond r_buff/2: rebuilt
    s2 = System(3000, s.L, seed=9, lattice=True, want_ref=False)
    close(U, s2.eng.mreal(far, s.F), 1e-6)
    # host-buffer entry point == device entry point
    pd = s.pos.clone(); im = torch.zeros((s.N, 3), dtype=torch.int32, device="cuda")
    s.eng.lanczos_m = 2; s2.eng.lanczos_m = 2
    s.eng.step(pd, im, s.F, 11, shear_rate=0.3)
    ph = s.pos_np.copy(); ih = np.zeros((s.N, 3), dtype=np.int32)
    s2.eng.step_host(ph, ih, s.F_np, 11, shear_rate=0.3)
    assert np.abs(ph - pd.cpu().numpy()).max() < 1e-5 and np.array_equal(ih, im.cpu().numpy())


def test_pipelined_host_steps_match_device_steps(cuda):
    """pse_step_host_async / pse_wait (device-resident state, forces up and state down on copy streams beside the compute,
    double-buffered host arrays) against the same steps through the device entry point: same trajectories, including
    list rebuilds, a state re-upload in the middle (PSE_HOST_STATE_IN), prefetched forces (pse_host_prefetch_forces) and a velocity download."""
    import torch
    N, L = 20000, util.box_length(20000, 0.2)
    a = System(N, L, seed=4, lattice=True, want_ref=False)
    b = System(N, L, seed=4, lattice=True, want_ref=False)
    pd = a.pos.clone(); im = torch.zeros((N, 3), dtype=torch.int32, device="cuda"); vd = torch.zeros_like(a.F)
    hp = [a.pos_np.copy(), a.pos_np.copy()]; hi = [np.zeros((N, 3), dtype=np.int32) for _ in range(2)]
    hv = np.zeros((N, 4), dtype=np.float32)
    Fs = [util.random_forces(N, 50 + t) for t in range(8)]
    a.eng.lanczos_m = 3; b.eng.lanczos_m = 3
    for t in range(8):
        a.eng.step(pd, im, torch.from_numpy(Fs[t]).cuda(), t, shear_rate=0.1, vel=vd)
        if t == 4:       # the host edits the state: it must be taken from the host arrays again
            b.eng.wait()
            hp[t & 1][:] = hp[(t - 1) & 1]; hi[t & 1][:] = hi[(t - 1) & 1]
        if t == 5:                               # from step 6 on the forces come from prefetches, issued one call ahead (queue of two)
            b.eng.prefetch_forces(Fs[6])
        if t == 6:
            b.eng.prefetch_forces(Fs[7])
        b.eng.step_host_async(hp[t & 1], hi[t & 1], None if t >= 6 else Fs[t], t, shear_rate=0.1, vel_np=hv if t == 7 else None, state_in=(t == 4))
    b.eng.wait()
    assert a.eng.stats()["nlist_builds"] >= 2
    # (spreading merges tile windows with floating-point reductions in arbitrary order: equal to round-off, not bitwise)
    assert np.abs(hp[7 & 1] - pd.cpu().numpy()).max() < 2e-5 and np.array_equal(hi[7 & 1], im.cpu().numpy())
    assert np.abs(hv[:, :3] - vd.cpu().numpy()[:, :3]).max() < 2e-4 * np.abs(hv[:, :3]).max()
    a.eng.close(); b.eng.close()


def test_operator_reapplied_at_fixed_positions_reuses_static_work(cuda):
    """M applied again at bit-identical positions keeps the pruned list / wave-space binning / factor rows of the previous call
    (fewer launches), any movement - or a new tilt - redoes them; results equal the engine that never reuses (PSE_REUSE=0)."""
    import os
    import torch
    s = System(20000, util.box_length(20000, 0.2), seed=6, lattice=True, want_ref=False)
    G = torch.from_numpy(util.random_forces(s.N, 77)).cuda()
    U1 = s.eng.mobility(s.pos, s.F).clone()
    l0 = s.eng.stats()["kernel_launches"]
    U2 = s.eng.mobility(s.pos, G).clone()
    l1 = s.eng.stats()["kernel_launches"]
    moved = s.pos.clone(); moved[5, 1] += 1e-4
    U3 = s.eng.mobility(moved, G).clone()
    l2 = s.eng.stats()["kernel_launches"]
    assert l1 - l0 < l2 - l1                       # the second call skipped the position-only kernels, the third did not
    os.environ["PSE_REUSE"] = "0"
    try:
        t = System(20000, s.L, seed=6, lattice=True, want_ref=False)
    finally:
        del os.environ["PSE_REUSE"]
    close(U1, t.eng.mobility(s.pos, s.F), 2e-6); close(U2, t.eng.mobility(s.pos, G), 2e-6); close(U3, t.eng.mobility(moved, G), 2e-6)
    s.eng.set_tilt(0.1)
    t.eng.set_tilt(0.1)
    close(s.eng.mobility(moved, G), t.eng.mobility(moved, G), 2e-6)   # same positions, new box: nothing may be reused
    s.eng.close(); t.eng.close()


def test_tiled_wave_path_is_bitwise_reproducible_and_matches_scatter_path(cuda):
    """The tile-owned spreading has a fixed summation order (no atomics): repeated runs are bitwise equal; the
    fallback scatter path (PSE_WAVE_TILED=0, used for tiny grids / P > 10) gives the same answer to round-off."""
    import os
    import torch
    os.environ["PSE_WAVE"] = "v1"
    try:
        s = System(20000, util.box_length(20000, 0.2), xy=0.2, seed=12, want_ref=False)
    finally:
        del os.environ["PSE_WAVE"]
    a = s.eng.mwave(s.pos, s.F); b = s.eng.mwave(s.pos, s.F)
    assert torch.equal(a, b)
    # default path (spread2: window merged with vector reductions, summation order across blocks not fixed): round-off level
    s3 = System(20000, s.L, xy=0.2, seed=12, want_ref=False)
    c = s3.eng.mwave(s3.pos, s3.F); d = s3.eng.mwave(s3.pos, s3.F)
    close(c, d, 1e-6); close(c, a, 1e-6)
    os.environ["PSE_WAVE_TILED"] = "0"
    try:
        s2 = System(20000, s.L, xy=0.2, seed=12, want_ref=False)
    finally:
        del os.environ["PSE_WAVE_TILED"]
    close(a, s2.eng.mwave(s2.pos, s2.F), 1e-5)  # tiled path uses ex2-based exponentials


@pytest.mark.parametrize("xy", [0.0, 0.25, 0.5, -0.5])
def test_config4_sheared_suspension_parity(cuda, xy):
    """BASELINE.json config 4: N = 100k, phi = 0.3 steady-shear suspension, deformed box with tilt swept over the
    wrapped-strain range of the variant ([-0.5, 0.5]): integer outputs exact, M.F vs the reference."""
    s = System(100000, util.box_length(100000, 0.3), xy=xy, seed=21, lattice=True)
    assert s.p.Nx == 108
    gi = s.eng.grid_index(s.pos).cpu().numpy()
    assert np.array_equal(gi, s.orc.grid_index(s.pos_np))
    nn, head, nl = s.nl_np()
    onn, ohead, onl = s.orc.neighbors(s.pos_np, s.p.rcut + 0.4, brute=False)
    assert np.array_equal(nn, onn) and np.array_equal(nl, onl)
    if s.ref is None:
        pytest.skip("reference library not built")
    close(s.eng.mobility(s.pos, s.F), s.ref.mobility(s.pos, s.F))
    # one sheared step: affine advection v.x += rate * y (PSEv1/Stokes.cu:168)
    import torch
    pe, pr = s.pos.clone(), s.pos.clone()
    ie = torch.zeros((s.N, 3), dtype=torch.int32, device="cuda"); ir = ie.clone()
    vel = torch.zeros_like(s.F); vel[:, 3] = 1.0
    s.eng.set_temperature(0.0)
    s.ref.step(pr, vel, torch.zeros((s.N, 3), device="cuda"), ir, s.F, 0.0, s.dt, 0, shear_rate=1.0)
    s.eng.step(pe, ie, s.F, 0, shear_rate=1.0)
    assert torch.equal(ie, ir) and float((pe - pr).abs().max()) < 1e-5


def test_config2_parity(cuda):
    """BASELINE.json config 2: N = 100k, phi = 0.2, error 1e-3 (125^3 grid, odd: no Nyquist planes): integer outputs exact,
    M.F and the full velocity (identical random vectors, same Lanczos m) against the reference kernels."""
    s = System(100000, util.box_length(100000, 0.2), seed=31, lattice=True)
    assert (s.p.Nx, s.p.P) == (125, 6)
    assert np.array_equal(s.eng.grid_index(s.pos).cpu().numpy(), s.orc.grid_index(s.pos_np))
    nn, head, nl = s.nl_np()
    onn, ohead, onl = s.orc.neighbors(s.pos_np, s.p.rcut + 0.4, brute=False)
    assert np.array_equal(nn, onn) and np.array_equal(nl, onl)
    assert abs(nl.size / s.N - 36.2) < 3.5   # SURVEY.md §8 table (ideal-gas estimate; the jittered lattice differs by a few)
    if s.ref is None:
        pytest.skip("reference library not built")
    close(s.eng.mreal(s.pos, s.F), s.ref.mreal(s.pos, s.F))
    close(s.eng.mwave(s.pos, s.F), s.ref.mwave(s.pos, s.F))
    up, ug = s.noise()
    s.ref.set_noise_tables(up, ug)
    try:
        s.eng.lanczos_m = 2; s.ref.m_lanczos = 2
        Ue, m = s.eng.velocity(s.pos, s.F, timestep=9, u_particles=up, u_grid=ug, parts=7)
        Ur = s.ref.velocity(s.pos, s.F, s.T, s.dt, 9)
        assert m == s.ref.m_lanczos
        close(Ue, Ur)
    finally:
        s.ref.set_noise_tables(None, None)


# ---------------------------------------------------------------- properties at the headline size
@pytest.fixture(scope="module")
def big(cuda):
    """BASELINE.json config 3: N = 1,000,000, phi = 0.3, error 1e-3 (240^3 grid)."""
    return System(1000000, util.box_length(1000000, 0.3), lattice=True, seed=0)


def test_full_size_integer_outputs(big):
    gi = big.eng.grid_index(big.pos).cpu().numpy()
    assert np.array_equal(gi, big.orc.grid_index(big.pos_np))
    _check_index_against_reference_spread(big, gi)
    nn, head, nl = big.nl_np()
    onn, ohead, onl = big.orc.neighbors(big.pos_np, big.p.rcut + 0.4, brute=False)
    assert np.array_equal(nn, onn) and np.array_equal(nl, onl)
    assert abs(nl.size / big.N - 54.3) < 1.5


def test_full_size_properties(big):
    torch = big.torch
    F, G = big.F, torch.from_numpy(util.random_forces(big.N, 77)).cuda()
    for op in (big.eng.mreal, big.eng.mwave, big.eng.mobility):
        MF, MG = op(big.pos, F), op(big.pos, G)
        a = float((G[:, :3].double() * MF[:, :3].double()).sum()); b = float((F[:, :3].double() * MG[:, :3].double()).sum())
        assert abs(a - b) < 1e-4 * max(abs(a), abs(b))            # symmetry
        assert float((F[:, :3].double() * MF[:, :3].double()).sum()) > 0   # each half positive definite
        lin = op(big.pos, 2.0 * F - 0.5 * G)
        close(lin, 2.0 * MF - 0.5 * MG, 2e-6)                     # linearity


def test_full_size_parity_with_reference_kernels(big):
    if big.ref is None:
        pytest.skip("reference library not built")
    close(big.eng.mobility(big.pos, big.F), big.ref.mobility(big.pos, big.F))
    up, ug = big.noise()
    big.ref.set_noise_tables(up, ug)
    try:
        big.eng.lanczos_m = 2; big.ref.m_lanczos = 2
        Ue, m = big.eng.velocity(big.pos, big.F, timestep=3, u_particles=up, u_grid=ug, parts=7)
        Ur = big.ref.velocity(big.pos, big.F, big.T, big.dt, 3)
        assert m == big.ref.m_lanczos
        close(Ue, Ur)
    finally:
        big.ref.set_noise_tables(None, None)


# ---------------------------------------------------------------- own FFT passes vs cuFFT
@pytest.mark.parametrize("N,phi,xy,Lfac", [(1000, 0.1, 0.0, (1, 1, 1)), (3000, 0.05, 0.3, (1, 1, 1)), (4000, 0.1, -0.2, (1.0, 1.3, 0.8)),
                                           (20000, 0.2, 0.0, (1, 1, 1))])
def test_own_fft_matches_cufft(cuda, monkeypatch, N, phi, xy, Lfac):
    """fft.cuh (shared-memory passes, scaling fused into the x pass) against the cuFFT R2C/C2R + scale_kernel pipeline:
    deterministic wave part and the wave-space noise for identical uniforms; grids with different radix mixes."""
    import torch
    from pse_b200 import engine as E
    L0 = util.box_length(N, phi)
    L = tuple(L0 * f for f in Lfac)
    cfg = E.make_config(N, L if Lfac != (1, 1, 1) else L0, xy=xy, T=1.0, dt=1e-3, seed=3)
    pos_np = util.random_positions(N, L0, 2)
    pos_np[:, 1] *= Lfac[1]; pos_np[:, 2] *= Lfac[2]
    pos = torch.from_numpy(pos_np).cuda(); F = torch.from_numpy(util.random_forces(N, 4)).cuda()
    out = {}
    for mode in ("own", "cufft"):
        monkeypatch.setenv("PSE_FFT", mode)
        eng = E.Engine(cfg)
        p = eng.params
        g = torch.Generator(device="cuda"); g.manual_seed(11)
        ug = torch.rand((p.Nx * p.Ny * p.Nz, 6), device="cuda", generator=g)
        out[mode] = (eng.mwave(pos, F).clone(), eng.velocity(pos, F, timestep=5, u_grid=ug, parts=2)[0].clone(), (p.Nx, p.Ny, p.Nz))
        eng.close()
    print("grid", out["own"][2])
    close(out["own"][0], out["cufft"][0], 2e-6)
    close(out["own"][1], out["cufft"][1], 2e-6)


# ---------------------------------------------------------------- BASELINE.json config 5 (largest): N = 8M, phi = 0.4, error 1e-4
def test_config5_eight_million_properties(cuda):
    """The largest BASELINE.json configuration at the reference's sizing rule (xi = 0.45 -> 432^3 grid, P = 8; xi = 0.5 would need
    576^3 > the 512^3 cap of PSEv1/Stokes.cc:203): M.F and the full velocity (identical random vectors, same Lanczos m) against the
    reference's own kernels at this size, plus size-independent properties (symmetry, positive definiteness, linearity) and one
    full BD step."""
    import torch
    from pse_b200 import engine as E
    free, _ = torch.cuda.mem_get_info()
    if free < 80 * 2**30:
        pytest.skip("needs ~60 GB of device memory")
    N, phi = 8000000, 0.4
    L = util.box_length(N, phi)
    cfg = E.make_config(N, L, xi=0.45, error=1e-4, T=1.0, dt=1e-3, seed=7)
    eng = E.Engine(cfg)
    p = eng.params
    assert (p.Nx, p.Ny, p.Nz, p.P) == (432, 432, 432, 8)
    pos = torch.from_numpy(util.lattice_positions(N, L, 0)).cuda()
    F = torch.from_numpy(util.random_forces(N, 1)).cuda(); G = torch.from_numpy(util.random_forces(N, 2)).cuda()
    MF, MG = eng.mobility(pos, F).clone(), eng.mobility(pos, G).clone()
    a = float((G[:, :3].double() * MF[:, :3].double()).sum()); b = float((F[:, :3].double() * MG[:, :3].double()).sum())
    assert abs(a - b) < 1e-4 * max(abs(a), abs(b))
    assert float((F[:, :3].double() * MF[:, :3].double()).sum()) > 0
    close(eng.mobility(pos, 2.0 * F - 0.5 * G), 2.0 * MF - 0.5 * MG, 3e-6)
    st = eng.stats()
    assert abs(st["nnz"] / N - 145.8) < 8.0                          # <nbrs> within r_cut + 0.4 at phi = 0.4 (145.8 for an ideal gas; the jittered lattice gives 140)
    # the reference's own kernels at this size (gpu_stokes_Mobility_wrap, PSEv1/Mobility.cu:729-782, and the full velocity with
    # identical random vectors, PSEv1/Brownian.cu:772-923): ~3.2 GB of complex grids + k table and a 12.9 GB Krylov basis
    from oracle import refwrap
    if refwrap.available():
        ref = refwrap.Reference(cfg_ref_pi(cfg), p, E.ewald_table(cfg))
        engp = E.Engine(cfg_ref_pi(cfg))          # same 2*pi constant as the reference (PSE_FLAG_REF_PI)
        engp.build_neighbors(pos)
        ref.set_neighbors(*engp.neighbor_list())
        close(engp.mobility(pos, F), ref.mobility(pos, F))
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        up = torch.rand((N, 3), device="cuda", generator=g); ug = torch.rand((p.Nx * p.Ny * p.Nz, 6), device="cuda", generator=g)
        v = ug.view(p.Nx, p.Ny, p.Nz, 6); v[:, :, p.Nz // 2, :] = 0.5; v[:, p.Ny // 2, 0, :] = 0.5   # doubly-visited nodes
        ref.set_noise_tables(up, ug)
        try:
            engp.lanczos_m = 2; ref.m_lanczos = 2
            Ue, m5 = engp.velocity(pos, F, timestep=3, u_particles=up, u_grid=ug, parts=7)
            Ur = ref.velocity(pos, F, 1.0, 1e-3, 3)
            assert m5 == ref.m_lanczos
            close(Ue, Ur)
        finally:
            ref.set_noise_tables(None, None)
        del ref, up, ug, v, Ue, Ur
        engp.close()
        torch.cuda.empty_cache()
    img = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    p0 = pos.clone()
    m = eng.step(pos, img, F, 0)
    torch.cuda.synchronize()
    d = (pos[:, :3] - p0[:, :3]).abs()
    d = torch.minimum(d, L - d)
    assert 2 <= m <= 20 and torch.isfinite(pos).all() and 0.01 < float(d.max()) < 1.0   # Brownian step ~ sqrt(2 kT M dt) ~ 0.05
    eng.close()


def cfg_ref_pi(cfg):
    import copy
    from pse_b200 import _lib
    c = copy.copy(cfg)
    c.flags = cfg.flags | _lib.PSE_FLAG_REF_PI
    return c


# ---------------------------------------------------------------- multi-GPU logic on ONE device: virtual ranks
@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("g,xy,comm", [(1, 0.0, "peer"), (2, 0.0, "peer"), (3, 0.3, "peer"), (2, 0.3, "coll")])
def test_sharded_step_virtual_ranks(cuda, monkeypatch, g, xy, comm):
    """the slab decomposition of the WHOLE step (own-row neighbour list / pruning / SpMV / Lanczos with vector
    halo rows and the three-word (double) all-reduce, own-plane spreading with halo-plane reduction, FFT passes + transposes + fused x
    pass, halo fetch, interpolation, velocity all-gather) run as g virtual ranks on one GPU (pse_local_world: the collectives
    become device copies + a host barrier inside the library) against the single-domain engine: every operator, injected
    noise, and three full steps with list rebuilds.  "peer": the default transport (kernels reading the other ranks' buffers
    behind a device-side flag barrier, peer.cuh); "coll": the packed collectives (NCCL between processes)."""
    import torch
    monkeypatch.setenv("PSE_COMM", comm)
    from pse_b200 import engine as E, sharded as S
    N, L = 30000, util.box_length(30000, 0.2)
    cfg = E.make_config(N, L, xy=xy, T=1.0, dt=1e-3, seed=1)
    pos = torch.from_numpy(util.lattice_positions(N, L, 4)).cuda(); F = torch.from_numpy(util.random_forces(N, 5)).cuda()
    single = E.Engine(cfg)
    p = single.params
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    up = torch.rand((N, 3), device="cuda", generator=gen); ug = torch.rand((p.Nx * p.Ny * p.Nz, 6), device="cuda", generator=gen)
    ref = {"mf": single.mobility(pos, F).clone(), "mreal": single.mreal(pos, F).clone(), "mwave": single.mwave(pos, F).clone()}
    single.lanczos_m = 4
    ref["vel"], ref["m"] = single.velocity(pos, F, 7, up, ug)
    ref["wn"], _ = single.velocity(pos, F, 7, up, ug, parts=2)
    ps, img = pos.clone(), torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    for t in range(3):
        single.step(ps, img, F, t)
    torch.cuda.synchronize()
    lw = S.LocalWorld(cfg, g)

    def work(r, e):
        out = {"mf": e.mobility(pos, F), "mreal": e.mreal(pos, F), "mwave": e.mwave(pos, F)}
        e.lanczos_m = 4
        out["vel"], out["m"] = e.velocity(pos, F, 7, up, ug)
        out["wn"], _ = e.velocity(pos, F, 7, up, ug, parts=2)
        q, im = pos.clone(), torch.zeros((N, 3), dtype=torch.int32, device="cuda")
        for t in range(3):
            e.step(q, im, F, t)
        out["pos"], out["img"] = q, im
        out["info"] = e.shard_info().as_dict()
        return out
    outs = lw.run(work)
    torch.cuda.synchronize()
    for r, o in enumerate(outs):
        for k in ("mf", "mreal", "mwave", "vel", "wn"):
            l2, mx = util.rel_err(o[k].cpu().numpy(), ref[k].cpu().numpy())
            assert l2 < 5e-6 and mx < 1e-5, (g, r, k, l2, mx)
        assert o["m"] == ref["m"]
        d = (o["pos"][:, :3] - ps[:, :3]).abs().max().item()
        assert d < 2e-5 and torch.equal(o["img"], img), (g, r, d)
        assert torch.equal(o["pos"], outs[0]["pos"])        # replicated state stays bitwise identical across ranks
        assert torch.equal(o["vel"], outs[0]["vel"])
    infos = [o["info"] for o in outs]
    assert infos[0]["row0"] == 0 and infos[-1]["row1"] == N and infos[0]["x0"] == 0 and infos[-1]["x1"] == p.Nx
    for r in range(g - 1):
        assert infos[r]["row1"] == infos[r + 1]["row0"] and infos[r]["x1"] == infos[r + 1]["x0"]
    if g > 1:
        assert all(i["collectives"] > 0 and i["bytes_sent"] > 0 for i in infos)
        assert all(i["buffer_planes"] < p.Nx for i in infos)   # a rank holds its slab of the grid only
    lw.close(); single.close()


# ---------------------------------------------------------------- physics known-answers of SURVEY.md §4
def test_xi_independence(cuda):
    """"Changing value will not affect results, only speed" (examples/run.py:50): M.F at xi = 0.3, 0.5, 0.8 agrees within `error`."""
    import torch
    from pse_b200 import engine as E
    N, L = 2000, util.box_length(2000, 0.1)
    pos = torch.from_numpy(util.lattice_positions(N, L, 6)).cuda(); F = torch.from_numpy(util.random_forces(N, 7)).cuda()
    U = {}
    for xi in (0.3, 0.5, 0.8):
        eng = E.Engine(E.make_config(N, L, xi=xi, error=1e-3, T=1.0, dt=1e-3, seed=1))
        U[xi] = eng.mobility(pos, F).clone()
        eng.close()
    for xi in (0.3, 0.8):
        l2, mx = util.rel_err(U[xi].cpu().numpy(), U[0.5].cpu().numpy())
        assert l2 < 6e-3 and mx < 6e-3, (xi, l2, mx)   # each side is within 3x error of the dense sum (test above)


def test_two_spheres_far_field(cuda):
    """Two spheres d = 4a apart in a large periodic box: the pair mobility is the free-space RPY tensor minus the periodic
    background 2.837297 a/L (same constant as the self term) up to O(d^2/L^3)."""
    import torch
    from pse_b200 import engine as E
    L, d = 80.0, 4.0
    pos = torch.tensor([[-d / 2, 0.3, -0.2, 0.0], [d / 2, 0.3, -0.2, 0.0]], dtype=torch.float32, device="cuda")
    eng = E.Engine(E.make_config(2, L, xi=0.5, error=1e-3, T=1.0, dt=1e-3, seed=1))
    par = 3.0 / (2 * d) - 1.0 / d**3 - 2.837297 / L      # along the line of centres
    perp = 3.0 / (4 * d) + 1.0 / (2 * d**3) - 2.837297 / L
    for axis, expect in ((0, par), (1, perp), (2, perp)):
        F = torch.zeros((2, 4), dtype=torch.float32, device="cuda"); F[1, axis] = 1.0
        U = eng.mobility(pos, F)
        assert abs(float(U[0, axis]) - expect) < 3e-3, (axis, float(U[0, axis]), expect)
        assert abs(float(U[1, axis]) - (1 - 2.837297 / L + 4 * math.pi / 3 / L**3)) < 3e-3
    eng.close()


# ---------------------------------------------------------------- round-2 parity gaps ("next round" item 1)
def test_config1_dense_ewald_error_at_N1000(cuda):
    """BASELINE.json config 1 at its own size (N = 1000, phi = 0.1): M.F against the dense double-precision Ewald sum.
    north_star: "both sides must stay within the requested Ewald error".  Measured on a B200 (relative L2 error / `error`):
    1.08 (ortho, 1e-3), 1.68 (xy = 0.3, 1e-3), 1.90 and 1.94 (1e-4) - the reference's parameter rules (PSEv1/Stokes.cc:129-236)
    size r_cut, k_max and the Gaussian support so that EACH truncation is about `error`; their sum lands between one and two
    times `error`, and the reference's own kernels, run here on the same inputs, land on the same value (asserted below).
    So the bound asserted is 2 x `error` for both sides."""
    from oracle import oraclewrap as O
    N = 1000
    L = util.box_length(N, 0.1)
    got = {}
    for xy in (0.0, 0.3):
        for error in (1e-3, 1e-4):
            s = System(N, L, xy=xy, seed=3, ref_pi=False, error=error, want_ref=False)
            Ud = O.dense_mobility(s.pos_np[:, :3], s.F_np[:, :3], L, xy=xy)
            U = s.eng.mobility(s.pos, s.F).cpu().numpy()[:, :3].astype(np.float64)
            err = float(np.linalg.norm(U - Ud) / np.linalg.norm(Ud))
            s.eng.close()
            sr = System(N, L, xy=xy, seed=3, ref_pi=True, error=error)   # the reference's kernels (with their 2*pi constant)
            err_ref = None
            if sr.ref is not None:
                Ur = sr.ref.mobility(sr.pos, sr.F).cpu().numpy()[:, :3].astype(np.float64)
                err_ref = float(np.linalg.norm(Ur - Ud) / np.linalg.norm(Ud))
            sr.eng.close()
            got[(xy, error)] = (err, err_ref)
    print("dense-Ewald relative L2 error at N = 1000 (engine, reference kernels):",
          {k: (f"{v[0]:.2e}", None if v[1] is None else f"{v[1]:.2e}") for k, v in got.items()})
    for (xy, error), (err, err_ref) in got.items():
        assert err < ERR_MULT * error, (xy, error, err)
        if err_ref is not None:
            assert err_ref < (ERR_MULT + 0.2) * error and abs(err - err_ref) < 0.35 * error, (xy, error, err, err_ref)


ERR_MULT = 2.0   # see the docstring above for the measured multiples


def test_gpu_lanczos_matches_dense_sqrtm(cuda):
    """The GPU Lanczos path (pse_velocity, parts = 4, injected psi) against sqrt(2T/dt) sqrtm(M_real) psi with M_real assembled
    column by column from pse_mreal and its square root from a dense symmetric eigendecomposition (float64).  Lanczos stops
    at a relative step norm of `error` (PSEv1/Brownian.cu:606), so agreement is expected at the `error` level, and a tighter
    `error` (which also moves r_cut, i.e. M_real itself) must converge further against ITS dense square root."""
    import torch
    N = 300
    L = util.box_length(N, 0.2)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    up = torch.rand((N, 3), device="cuda", generator=g)
    a = np.float32(1.73205080757)
    psi = (np.float32(2) * a * up.cpu().numpy() - a).astype(np.float64).reshape(-1)
    got = {}
    for error, tol in ((1e-3, 2e-3), (1e-5, 5e-5)):
        s = System(N, L, seed=13, lattice=True, want_ref=False, error=error)
        M = np.zeros((3 * N, 3 * N))
        for c in range(3 * N):
            e = torch.zeros_like(s.F); e[c // 3, c % 3] = 1
            M[:, c] = s.eng.mreal(s.pos, e).cpu().numpy()[:, :3].reshape(-1).astype(np.float64)
        assert np.abs(M - M.T).max() < 2e-6
        M = 0.5 * (M + M.T)
        lam, W = np.linalg.eigh(M)
        assert lam.min() > 0   # "positively split": the real-space part alone is positive definite
        exact = math.sqrt(2 * s.T / s.dt) * (W @ (np.sqrt(lam) * (W.T @ psi)))
        s.eng.lanczos_m = 2
        U, m = s.eng.velocity(s.pos, s.F, timestep=1, u_particles=up, parts=4)
        err = np.linalg.norm(U.cpu().numpy()[:, :3].astype(np.float64).reshape(-1) - exact) / np.linalg.norm(exact)
        print(f"error = {error}: Lanczos m = {m}, relative error vs dense sqrtm = {err:.2e}")
        assert 2 <= m <= 40 and err < tol, (error, m, err)
        got[error] = (m, err)
        s.eng.close()
    assert got[1e-5][0] > got[1e-3][0] and got[1e-5][1] < got[1e-3][1]


def test_tilt_flip_through_max_strain(cuda):
    """a steady-shear run whose wrapped strain (`variant.shear_variant`, PSEv1/VariantShearFunction.cc:34-43)
    jumps from +max_strain to -max_strain.  box_resize re-images the particles after the jump (HOOMD's updater does); on the
    steps around the flip the neighbour list equals brute force, the particle->grid index is bit-exact, every particle sits
    inside the primary cell of the new box, and M.F matches the reference kernels and the dense Ewald sum."""
    import torch
    import pse_b200 as PSEv1
    from oracle import oraclewrap as O
    from oracle import refwrap
    from pse_b200 import _lib
    from pse_b200 import engine as E
    N, L, dt = 1500, util.box_length(1500, 0.15), 0.02
    pos0 = util.lattice_positions(N, L, 17)
    s = PSEv1.system.set_current(PSEv1.system.System(pos0, PSEv1.system.Box(L)))
    PSEv1.integrate.mode_standard(dt=dt)
    ff = PSEv1.shear_function.steady(dt=dt, shear_rate=2.5)          # strain advances by 0.05 per step: flip after step 10
    pse = PSEv1.integrate.PSEv1(group=s.all(), seed=2, T=1e-3, xi=0.5, error=1e-3, function_form=ff)
    var = PSEv1.variant.shear_variant(ff, total_timestep=1000)
    PSEv1.system.box_resize(s, xy=var)
    eng = pse.cpp_method
    F = torch.from_numpy(util.random_forces(N, 5)).cuda()
    s.set_forces(0.05 * F)
    tilts = []
    flipped = False
    for t in range(14):
        s.run(1)
        xy = float(var.get_value(t))           # the tilt step t was taken with
        tilts.append(xy)
        just_flipped = len(tilts) > 1 and tilts[-1] < tilts[-2] - 0.5
        flipped |= just_flipped
        if not (just_flipped or abs(xy) > 0.4):
            continue
        # state after the step, in the box of tilt xy
        pos_np = s.pos.cpu().numpy()
        orc = O.Oracle(N, L, xy=xy, ref_pi=True)
        eng.build_neighbors(s.pos)
        nn, head, nl = [a.cpu().numpy().view(np.uint32) for a in eng.neighbor_list()]
        onn, ohead, onl = orc.neighbors(pos_np, eng.params.rcut + 0.4, brute=True)
        assert np.array_equal(nn, onn) and np.array_equal(nl, onl), (t, xy)
        assert np.array_equal(eng.grid_index(s.pos).cpu().numpy(), orc.grid_index(pos_np)), (t, xy)
        frac_x = (pos_np[:, 0] - xy * pos_np[:, 1]) / L + 0.5
        assert frac_x.min() > -1e-3 and frac_x.max() < 1 + 1e-3, (t, xy, frac_x.min(), frac_x.max())
        U = eng.mobility(s.pos, F)
        if just_flipped:
            Ud = O.dense_mobility(pos_np[:, :3].astype(np.float64), F.cpu().numpy()[:, :3].astype(np.float64), L, xy=xy)
            err = np.linalg.norm(U.cpu().numpy()[:, :3] - Ud) / np.linalg.norm(Ud)
            assert err < 3e-3, (t, xy, err)
        if refwrap.available():
            cfg = E.make_config(N, L, xy=xy, flags=_lib.PSE_FLAG_REF_PI, T=1.0, dt=dt, seed=1)
            e2 = E.Engine(cfg)
            ref = refwrap.Reference(cfg, e2.params, E.ewald_table(cfg))
            e2.build_neighbors(s.pos); ref.set_neighbors(*e2.neighbor_list())
            close(e2.mobility(s.pos, F), ref.mobility(s.pos, F))
            e2.close()
    assert flipped and min(tilts) < -0.4 and max(tilts) > 0.4, tilts
