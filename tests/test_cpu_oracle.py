"""CPU suite (-m "not gpu"): pins the oracle (oracle/pse_oracle.c) and the host side of the engine
against golden vectors generated from the reference's own source expressions (tests/golden/,
tests/golden/make_golden.py), against physics known-answers the reference implies, and
checks that the C-ABI library loads and exports every symbol include/pse_b200.h declares.
No GPU compute is called here."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from oracle import oraclewrap as O
from pse_b200 import _lib
from pse_b200 import engine as E
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------- C ABI surface
def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pse_b200.h")).read()
    declared = set(re.findall(r"\b(pse_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pse_engine", "pse_shear"}
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = E.make_config(1000, 34.7)
    h = ctypes.c_void_p()
    rc = _lib.lib.pse_create(ctypes.byref(cfg), None, ctypes.byref(h))
    assert rc == _lib.PSE_ENODEVICE and not h.value
    assert b"no CPU fallback" in _lib.lib.pse_last_error(None)
    with pytest.raises(E.PSEError):
        E.Engine(cfg)


# ---------------------------------------------------------------- parameters (Stokes::setParams)
# expected values: SURVEY.md §8 table (float32 restatement of PSEv1/Stokes.cc:129-236,309-310)
CONFIGS = [
    (1000, 0.1, 1e-3, 0.5, dict(N3=36, P=6, kmax=3, ewald_n=5255, gaussm=4.22, eta=0.4703, rcut=5.25652)),
    (100000, 0.2, 1e-3, 0.5, dict(N3=125, P=6, kmax=3, ewald_n=5255, gaussm=4.22, eta=0.5295, rcut=5.25652)),
    (1000000, 0.3, 1e-3, 0.5, dict(N3=240, P=6, kmax=3, ewald_n=5255, gaussm=4.22, eta=0.5088, rcut=5.25652)),
    (100000, 0.3, 1e-3, 0.5, dict(N3=108, P=6, kmax=3, ewald_n=5255, gaussm=4.22, eta=0.5413, rcut=5.25652)),
    (8000000, 0.4, 1e-4, 0.45, dict(N3=432, P=8, kmax=3, ewald_n=6743, gaussm=4.99, eta=0.5340, rcut=6.74412)),
]


@pytest.mark.parametrize("N,phi,err,xi,exp", CONFIGS)
def test_params_match_survey_and_oracle(N, phi, err, xi, exp):
    L = util.box_length(N, phi)
    p = E.derive_params(E.make_config(N, L, xi=xi, error=err))
    o = O.Oracle(N, L, xi=xi, error=err).prm
    assert (p.Nx, p.Ny, p.Nz) == (exp["N3"],) * 3
    assert p.P == exp["P"] and p.kmax == exp["kmax"] and p.ewald_n == exp["ewald_n"]
    assert abs(p.gaussm - exp["gaussm"]) < 5e-3 and abs(p.eta - exp["eta"]) < 1e-4 and abs(p.rcut - exp["rcut"]) < 1e-5
    for k in ("Nx", "Ny", "Nz", "P", "kmax", "ewald_n", "rcut", "dr", "gaussm", "eta", "hx", "hy", "hz", "self", "quadW",
              "prefac", "expfac"):
        assert getattr(p, k) == getattr(o, k), k  # bit-identical between engine host code and oracle


def test_grid_cap_and_literal_runpy():
    # literal examples/run.py: L = 64 -> 64^3 grid (SURVEY.md §8 row 1')
    p = E.derive_params(E.make_config(1000, 64.0))
    assert (p.Nx, p.P) == (64, 6) and abs(p.eta - 0.5054) < 1e-4
    # 576^3 at xi = 0.5 for config 5 exceeds the reference's 512^3 cap (PSEv1/Stokes.cc:203-214)
    cfg = E.make_config(8000000, util.box_length(8000000, 0.4), xi=0.5, error=1e-4)
    out = _lib.pse_params()
    assert _lib.lib.pse_derive_params(ctypes.byref(cfg), ctypes.byref(out)) == _lib.PSE_EGRID
    cfg.flags = _lib.PSE_FLAG_LIFT_GRID_CAP
    assert _lib.lib.pse_derive_params(ctypes.byref(cfg), ctypes.byref(out)) == 0 and out.Nx == 576
    assert p.seed_hashed == ((((0 * 0x12345677 + 0x12345) & 0xFFFFFFFF) ^ (0x12345 >> 16)) * 0x45679) & 0xFFFFFFFF


def test_self_mobility_pin():
    p = E.derive_params(E.make_config(1000, 34.7))
    assert abs(p.self - 0.335617124) < 1e-7


# ---------------------------------------------------------------- real-space table
@pytest.mark.parametrize("xi", [0.5, 0.3, 0.8])
def test_table_matches_reference_expressions(xi):
    g = np.load(os.path.join(GOLD, f"ewald_table_xi{xi}.npz"))
    n = int(g["ewald_n"])
    o = O.Oracle(1000, 80.0, xi=xi)
    t = E.ewald_table(E.make_config(1000, 80.0, xi=xi))
    assert o.prm.ewald_n == n and t.shape == (n + 1, 4)
    for tab in (o.table, t):
        # bit-exact where the double-precision expressions are well conditioned; below r = 0.02 the reference's
        # own double arithmetic cancels ~1e8-sized terms and float-level differences are expected
        assert np.array_equal(tab[20:, :2], g["fg32"][20:])
        assert np.allclose(tab[:20, :2], g["fg32"][:20], rtol=5e-6, atol=0)
        assert np.array_equal(tab[:-1, 2:], tab[1:, :2]) and np.all(tab[-1, 2:] == 0)  # PSEv1/Stokes.cc:414-420


def test_table_known_answers():
    o = O.Oracle(1000, 80.0, xi=0.5)
    f2, g2 = o.real_fg(2.0, 0.5)
    assert abs(f2 + 0.00452117243) < 1e-10 and abs(g2 - 0.0829571233) < 1e-10
    # branches agree at contact
    fa, ga = o.real_fg(2.0 - 1e-9, 0.5); fb, gb = o.real_fg(2.0 + 1e-9, 0.5)
    assert abs(fa - fb) < 1e-8 and abs(ga - gb) < 1e-8
    # small-xi limit: free-space RPY minus 3 xi / sqrt(pi)
    xi = 0.02
    for r in (1.0, 3.0):
        f, g = o.real_fg(r, xi)
        frpy = 3 / (4 * r) * (1 + 2 / (3 * r * r)) if r > 2 else 1 - 9 * r / 32
        grpy = 3 / (4 * r) * (2 - 4 / (3 * r * r)) if r > 2 else 1 - 3 * r / 16
        c = 3 * xi / math.sqrt(math.pi)
        assert abs(f - (frpy - c)) < 1e-4 and abs(g - (grpy - c)) < 1e-4  # O(xi^3) corrections


def test_table_against_independent_quadrature():
    """f = f_RPY - f_wave, g = g_RPY - g_wave with the wave part integrated numerically from its Fourier
    definition: (1/2pi^2) int k^2 B(k) [j0 - j1/x | 2 j1/x] dk,
    B = (6 pi / k^2)(1 + k^2/4xi^2) exp(-k^2/4xi^2) sinc^2(k)."""
    from scipy import integrate
    o = O.Oracle(1000, 80.0, xi=0.5)
    xi = 0.5
    for r in (0.7, 1.9, 2.4, 4.0):
        def B(k):
            return 6 * math.pi / k**2 * (1 + k**2 / (4 * xi**2)) * math.exp(-k**2 / (4 * xi**2)) * (math.sin(k) / k) ** 2
        def j0(x): return math.sin(x) / x
        def j1(x): return math.sin(x) / x**2 - math.cos(x) / x
        fw = integrate.quad(lambda k: k * k * B(k) * (j0(k * r) - j1(k * r) / (k * r)), 1e-9, 40, limit=400)[0] / (2 * math.pi**2)
        gw = integrate.quad(lambda k: k * k * B(k) * (2 * j1(k * r) / (k * r)), 1e-9, 40, limit=400)[0] / (2 * math.pi**2)
        frpy = 3 / (4 * r) + 1 / (2 * r**3) if r >= 2 else 1 - 9 * r / 32
        grpy = 3 / (2 * r) - 1 / r**3 if r >= 2 else 1 - 3 * r / 16
        f, g = o.real_fg(r, xi)
        assert abs(f - (frpy - fw)) < 1e-8 and abs(g - (grpy - gw)) < 1e-8


# ---------------------------------------------------------------- shear functions
def _mk_shear():
    L = _lib.lib
    dt = 1e-3
    arr = lambda *a: (ctypes.c_double * len(a))(*a)
    f = [L.pse_shear_create(1, arr(1.5), 1, 10, dt), L.pse_shear_create(2, arr(2.0, 3.0), 2, 10, dt),
         L.pse_shear_create(3, arr(0.1, 1.0, 50.0, 2.0), 4, 10, dt), L.pse_shear_create(4, arr(2.0, 0.5), 2, 10, dt)]
    f.append(L.pse_shear_create_windowed(f[2], f[3]))
    f.append(L.pse_shear_create(0, None, 0, 0, dt))
    return f


def test_shear_functions_match_reference_classes():
    rows = np.load(os.path.join(GOLD, "shear_functions.npz"))["rows"]
    f = _mk_shear()
    L = _lib.lib
    for k, t, rate, strain, off in rows:
        h = f[int(k)]
        assert L.pse_shear_offset(h) == int(off)
        r, s = L.pse_shear_rate(h, int(t)), L.pse_shear_strain(h, int(t))
        assert r == pytest.approx(rate, rel=1e-13, abs=1e-13, nan_ok=True) and s == pytest.approx(strain, rel=1e-13, abs=1e-13, nan_ok=True)  # t < offset wraps unsigned -> inf/nan on both sides


def test_shear_variant_wraps_strain():
    L = _lib.lib
    h = L.pse_shear_create(1, (ctypes.c_double * 1)(1.0), 1, 5, 1e-2)  # strain = (t-5)*0.01
    v = lambda t: L.pse_shear_variant_value(h, 200, -0.5, 0.5, t)
    assert v(0) == 0 and v(5) == 0
    assert v(30) == pytest.approx(0.25) and v(56) == pytest.approx(-0.49) and v(155) == pytest.approx(-0.5)
    assert v(205) == v(1000) == pytest.approx(2.0 - 2.0)  # end value = wrap(strain(offset + total)) = wrap(2.0) = 0


# ---------------------------------------------------------------- host Lanczos pieces
def test_tridiag_sqrt_against_numpy():
    rng = np.random.default_rng(0)
    for m in (1, 2, 5, 17, 60, 100):
        a = 1 + rng.random(m); b = 0.3 * rng.random(max(m - 1, 1))
        T = np.diag(a) + np.diag(b[: m - 1], 1) + np.diag(b[: m - 1], -1)
        w, V = np.linalg.eigh(T)
        ref = (V * np.sqrt(w)) @ V[0]
        c = np.zeros(m)
        dp = lambda x: x.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        assert _lib.lib.pse_test_tridiag_sqrt_e1(m, dp(a), dp(b), dp(c)) == 0
        assert np.allclose(c, ref, rtol=1e-11, atol=1e-12)
    a = np.array([1.0, -2.0]); b = np.array([0.1]); c = np.zeros(2)
    assert _lib.lib.pse_test_tridiag_sqrt_e1(2, dp(a), dp(b), dp(c)) == _lib.PSE_EEIGEN  # not SPD (reference exits)


def test_philox_known_answers():
    def ph(c, k):
        c = np.array(c, dtype=np.uint32); k = np.array(k, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        ctypes.CDLL(_lib.LIB_PATH).pse_test_philox4x32(vp(c), vp(k), vp(o))
        return list(o)
    assert ph([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert ph([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert ph([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


# ---------------------------------------------------------------- oracle physics
def _system(N, phi, seed=0, xy=0.0, lattice=False, **kw):
    L = util.box_length(N, phi)
    o = O.Oracle(N, L, xy=xy, **kw)
    pos = util.lattice_positions(N, L, seed) if lattice else util.random_positions(N, L, seed)
    F = util.random_forces(N, seed + 1)
    return L, o, pos, F


def test_oracle_neighbor_list_brute_equals_cells():
    for xy in (0.0, 0.37):
        L, o, pos, F = _system(700, 0.15, xy=xy)
        a = [x.copy() for x in o.neighbors(pos, o.prm.rcut + 0.4, brute=True)]
        b = o.neighbors(pos, o.prm.rcut + 0.4, brute=False)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        nn, head, nl = a
        for i in (0, 13, 699):  # rows ascending, symmetric
            row = nl[head[i]: head[i] + nn[i]]
            assert np.all(np.diff(row.astype(np.int64)) > 0)
            for j in row[:5]:
                assert i in nl[head[j]: head[j] + nn[j]]


@pytest.mark.parametrize("xy", [0.0, 0.3])
def test_oracle_mobility_within_ewald_error_of_dense(xy):
    """Both halves of the split, summed, agree with a dense double-precision Ewald sum within `error`
    (exact pi on both sides) — and the answer does not depend on xi (examples/run.py:50)."""
    N, phi = 160, 0.1
    L = util.box_length(N, phi)
    pos = util.random_positions(N, L, 3); F = util.random_forces(N, 4)
    Ud = O.dense_mobility(pos[:, :3], F[:, :3], L, xy=xy)
    # "within the requested Ewald error": the reference truncates where the kernels have decayed to ~error
    # (f(rcut) = -1.3e-3 at error = 1e-3), so the bound is a small multiple of it, and
    # tightening `error` must tighten the result
    for xi, error in ((0.5, 1e-3), (0.4, 1e-3), (0.5, 1e-4)):
        o = O.Oracle(N, L, xi=xi, xy=xy, error=error, ref_pi=False)
        o.neighbors(pos, o.prm.rcut, brute=True)
        U = o.mreal(pos, F)[:, :3].astype(np.float64) + o.mwave(pos, F)[:, :3]
        err = np.linalg.norm(U - Ud) / np.linalg.norm(Ud)
        assert err < 3 * error, (xi, error, err)
    # dense oracle is split-independent to near round-off
    Ud2 = O.dense_mobility(pos[:, :3], F[:, :3], L, xy=xy, xi_d=0.45)
    assert np.linalg.norm(Ud - Ud2) / np.linalg.norm(Ud) < 1e-9


def test_periodic_self_mobility_of_one_sphere():
    # simple-cubic lattice of one sphere per cell: M = 1 - 2.837297 (a/L) + (4 pi/3)(a/L)^3   
    for L in (20.0, 30.0):
        pos = np.zeros((1, 3)); F = np.array([[1.0, 0.0, 0.0]])
        U = O.dense_mobility(pos, F, L, xi_d=0.5)
        assert abs(U[0, 0] - (1 - 2.837297 / L + 4 * math.pi / 3 / L**3)) < 2e-7
        o = O.Oracle(1, L, ref_pi=False)
        o.set_neighbors(np.zeros(1), np.zeros(1), np.zeros(1))
        p4 = np.zeros((1, 4), dtype=np.float32); F4 = np.array([[1, 0, 0, 0]], dtype=np.float32)
        u = o.mreal(p4, F4)[0, 0] + o.mwave(p4, F4)[0, 0]
        assert abs(u - U[0, 0]) < 3e-3  # truncation error of the split at error = 1e-3


def test_oracle_operators_symmetric_positive():
    L, o, pos, F = _system(300, 0.1, seed=5)
    G = util.random_forces(300, 9)
    o.neighbors(pos, o.prm.rcut + 0.4)
    for op in (o.mreal, o.mwave):
        a = np.sum(G[:, :3].astype(np.float64) * op(pos, F)[:, :3]); b = np.sum(F[:, :3].astype(np.float64) * op(pos, G)[:, :3])
        assert abs(a - b) < 2e-4 * max(abs(a), abs(b), 1.0)
        assert np.sum(F[:, :3].astype(np.float64) * op(pos, F)[:, :3]) > 0  # both halves positive: the "positive split"


def test_oracle_lanczos_matches_dense_sqrt():
    N = 60
    L, o, pos, _ = _system(N, 0.1, seed=7)
    o.neighbors(pos, o.prm.rcut + 0.4)
    M = np.zeros((3 * N, 3 * N))
    for c in range(3 * N):  # dense M_real by applying the SpMV to unit vectors
        e = np.zeros((N, 4), dtype=np.float32); e[c // 3, c % 3] = 1
        M[:, c] = o.mreal(pos, e)[:, :3].reshape(-1)
    assert np.allclose(M, M.T, atol=1e-6)
    w, V = np.linalg.eigh(0.5 * (M + M.T))
    assert w.min() > 0
    rng = np.random.default_rng(1)
    psi = np.zeros((N, 4), dtype=np.float32); psi[:, :3] = (rng.random((N, 3)) * 2 - 1) * math.sqrt(3)
    T, dt = 1.0, 1e-3
    ref = math.sqrt(2 * T / dt) * ((V * np.sqrt(w)) @ (V.T @ psi[:, :3].reshape(-1).astype(np.float64)))
    U, m, sn = o.lanczos(pos, psi, T, dt, m_in=2)
    err = np.linalg.norm(U[:, :3].reshape(-1) - ref) / np.linalg.norm(ref)
    assert 2 <= m <= 30 and sn <= 1e-3 and err < 2e-3, (m, sn, err)


def test_oracle_wave_noise_is_real_and_scaled():
    # pure noise field: Hermitian by construction -> velocity independent of forces, scales with sqrt(T/dt)
    N = 50
    L, o, pos, F = _system(N, 0.05, seed=2)
    G = o.prm.Nx * o.prm.Ny * o.prm.Nz
    ug = np.random.default_rng(3).random((G, 6), dtype=np.float32)
    a = o.mwave(pos, F, do_det=False, u_grid=ug, noise_fac=1.0)
    b = o.mwave(pos, 0 * F, do_det=False, u_grid=ug, noise_fac=2.0)
    assert np.allclose(2 * a, b, rtol=1e-5, atol=1e-7) and np.abs(a).max() > 0


def test_oracle_integrate_wraps_and_shears():
    o = O.Oracle(2, 20.0, xy=0.25)
    pos = np.array([[9.99, 1.0, -9.99, 0], [0, 9.995, 0, 0]], dtype=np.float32)
    vel = np.array([[20.0, 0, -20.0, 0], [0, 10.0, 0, 0]], dtype=np.float32)
    img = np.zeros((2, 3), dtype=np.int32)
    o.integrate(pos, img, vel, 1e-3, shear_rate=2.0)
    # particle 0: x advanced by (20 + 2*1)*1e-3 past hi + xy*y -> wrapped by -L; z wrapped by +L
    assert img[0].tolist() == [0, 0, -1] or img[0].tolist() == [1, 0, -1]
    assert img[1].tolist()[1] == 1 and pos[1, 1] < -9.9 and abs(pos[1, 0] - (2.0 * 9.995 * 1e-3 - 20 * 0.25)) < 1e-4


@pytest.mark.parametrize("xi,error", [(0.5, 1e-3), (0.5, 1e-4), (0.8, 1e-3), (0.45, 1e-4), (0.3, 1e-3)])
def test_constant_bank_polynomials_match_closed_forms(xi, error):
    """f(r), g(r) = exp(-xi^2 (r-2)^2) P8(A/r + B) used by the SpMV for r >= 2a (real.cuh TableCheb): the host fit reports its own
    float-evaluated error; here it is re-evaluated independently against the oracle's closed forms (oracle/pse_oracle.c:97)."""
    rcut = math.sqrt(-math.log(error)) / xi
    orc = O.Oracle(100, 40.0, xi=xi, error=error)
    out = (ctypes.c_float * 21)(); err = ctypes.c_double()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.pse_test_fit_rpy_cheb.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)]
    assert lib.pse_test_fit_rpy_cheb(xi, rcut, out, ctypes.byref(err)) == 0
    c = np.array(out[:], dtype=np.float64)
    A, B, ne, cf, cg = c[0], c[1], c[2], c[3:12], c[12:21]
    r = np.linspace(2.0, rcut, 1501)
    t = A / r + B
    assert t.min() > -1.0001 and t.max() < 1.0001
    e = np.exp2(ne * (r - 2.0) ** 2)
    f = np.polynomial.polynomial.polyval(t, cf) * e
    g = np.polynomial.polynomial.polyval(t, cg) * e
    fo, go = np.array([orc.real_fg(x, xi) for x in r]).T
    ef, eg = np.abs(f - fo).max() / np.abs(fo).max(), np.abs(g - go).max() / np.abs(go).max()
    assert ef < 3e-6 and eg < 3e-6, (ef, eg)   # same bound as the engine applies before trusting the fit
    assert err.value < 3e-6   # the engine keeps the reference table when this check fails


@pytest.mark.parametrize("N", [36, 45, 64, 72, 108, 125, 135, 240, 432, 576])
def test_inplace_fft_schedule_reproduces_numpy(N):
    """Host logic of fft.cuh: a numpy transcription of the in-place decimation-in-frequency passes (radices from the engine's
    planner) leaves X[k] at pos_of[k]; the adjoint passes (conjugate twiddles first) undo it up to the factor N."""
    lib = ctypes.CDLL(_lib.LIB_PATH)
    radix = (ctypes.c_int * 12)(); npass = ctypes.c_int(); pos = (ctypes.c_uint16 * N)()
    assert lib.pse_test_fft_plan(N, radix, ctypes.byref(npass), pos) == 0
    rad = list(radix[: npass.value]); pos = np.array(pos[:])
    assert int(np.prod(rad)) == N and sorted(pos.tolist()) == list(range(N))
    rng = np.random.default_rng(N)
    x = rng.normal(size=N) + 1j * rng.normal(size=N)
    w = np.exp(-2j * np.pi * np.arange(N) / N)
    a = x.copy(); n = N
    for r in rad:                                   # forward: butterfly, then twiddle w_n^{j p}
        m = n // r
        for b in range(N // n):
            for j in range(m):
                idx = b * n + j + m * np.arange(r)
                y = np.array([sum(a[idx[q]] * np.exp(-2j * np.pi * p * q / r) for q in range(r)) for p in range(r)])
                a[idx] = y * w[(j * np.arange(r) * (N // n)) % N]
        n = m
    X = np.fft.fft(x)
    assert np.allclose(a[pos], X, rtol=1e-10, atol=1e-10)
    n = 1
    for r in reversed(rad):                         # adjoint: conjugate twiddle, then conjugate butterfly
        n *= r; m = n // r
        for b in range(N // n):
            for j in range(m):
                idx = b * n + j + m * np.arange(r)
                y = a[idx] * np.conj(w[(j * np.arange(r) * (N // n)) % N])
                a[idx] = np.array([sum(y[p] * np.exp(2j * np.pi * p * q / r) for p in range(r)) for q in range(r)])
    assert np.allclose(a, N * x, rtol=1e-10, atol=1e-9)
