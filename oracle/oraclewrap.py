"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/_ref/libpse_oracle.so, the CPU (C + OpenMP)
restatement of the reference algorithm (oracle/pse_oracle.c).  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg only."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "_ref", "libpse_oracle.so")


class orc_config(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("Lx", ctypes.c_float), ("Ly", ctypes.c_float), ("Lz", ctypes.c_float),
                ("xy", ctypes.c_float), ("xi", ctypes.c_float), ("error", ctypes.c_float), ("max_strain", ctypes.c_float),
                ("ref_pi", ctypes.c_int)]


class orc_params(ctypes.Structure):
    _fields_ = [("Nx", ctypes.c_int), ("Ny", ctypes.c_int), ("Nz", ctypes.c_int), ("P", ctypes.c_int), ("kmax", ctypes.c_int),
                ("ewald_n", ctypes.c_int), ("rcut", ctypes.c_float), ("dr", ctypes.c_float), ("gaussm", ctypes.c_float),
                ("eta", ctypes.c_float), ("hx", ctypes.c_float), ("hy", ctypes.c_float), ("hz", ctypes.c_float),
                ("self", ctypes.c_float), ("quadW", ctypes.c_float), ("prefac", ctypes.c_float), ("expfac", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def available():
    return os.path.exists(ORACLE_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(ORACLE_LIB)
        _lib.orc_real_fg.restype = None
        _lib.orc_real_fg.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        _lib.orc_dense_mobility.argtypes = [ctypes.c_int] + [ctypes.c_double] * 4 + [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_double] * 3 + [ctypes.c_void_p]
        _lib.orc_set_spteqr.argtypes = [ctypes.c_void_p]
    return _lib


_blas = None


def use_lapacke(on=True):
    """Route the Lanczos tridiagonal solve through LAPACKE_spteqr as the reference does (PSEv1/Brownian.cu:540), bound from
    the OpenBLAS that ships with scipy (the library the compiled reference links, oracle/Makefile).  Returns True if bound."""
    global _blas
    if not on:
        lib().orc_set_spteqr(None)
        return False
    if _blas is None:
        import glob
        import scipy
        cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
        if not cands:
            return False
        _blas = ctypes.CDLL(cands[0])
    try:
        fn = ctypes.cast(_blas.scipy_LAPACKE_spteqr, ctypes.c_void_p)
    except AttributeError:
        return False
    lib().orc_set_spteqr(fn)
    return True


def _f(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    def __init__(self, N, L, xi=0.5, error=1e-3, max_strain=0.5, xy=0.0, ref_pi=True):
        Lx, Ly, Lz = (L, L, L) if np.isscalar(L) else L
        self.cfg = orc_config(N=N, Lx=Lx, Ly=Ly, Lz=Lz, xy=xy, xi=xi, error=error, max_strain=max_strain, ref_pi=1 if ref_pi else 0)
        self.prm = orc_params()
        rc = lib().orc_derive_params(ctypes.byref(self.cfg), ctypes.byref(self.prm))
        if rc != 0:
            raise RuntimeError(f"orc_derive_params failed: {rc}")
        self.N = N
        self.table = np.zeros((self.prm.ewald_n + 1, 4), dtype=np.float32)
        lib().orc_table(ctypes.byref(self.prm), ctypes.c_float(xi), _f(self.table))
        self.nn = self.head = self.nl = None

    def real_fg(self, r, xi):
        f, g = ctypes.c_double(), ctypes.c_double()
        lib().orc_real_fg(float(r), float(xi), ctypes.byref(f), ctypes.byref(g))
        return f.value, g.value

    def neighbors(self, pos4, rlist, brute=True):
        N = self.N
        fn = lib().orc_nlist_bruteforce if brute else lib().orc_nlist_cells
        nn = np.zeros(N, dtype=np.uint32); head = np.zeros(N, dtype=np.uint32)
        nnz = ctypes.c_size_t(0)
        fn(ctypes.byref(self.cfg), _f(pos4), ctypes.c_float(rlist), _f(nn), _f(head), None, ctypes.c_size_t(0), ctypes.byref(nnz))
        nl = np.zeros(max(nnz.value, 1), dtype=np.uint32)
        rc = fn(ctypes.byref(self.cfg), _f(pos4), ctypes.c_float(rlist), _f(nn), _f(head), _f(nl), ctypes.c_size_t(nl.size), ctypes.byref(nnz))
        assert rc == 0
        self.nn, self.head, self.nl = nn, head, nl[: nnz.value]
        return self.nn, self.head, self.nl

    def set_neighbors(self, nn, head, nl):
        self.nn, self.head, self.nl = (np.ascontiguousarray(a, dtype=np.uint32) for a in (nn, head, nl))

    def grid_index(self, pos4):
        out = np.zeros((self.N, 3), dtype=np.int32)
        lib().orc_grid_index(ctypes.byref(self.cfg), ctypes.byref(self.prm), _f(pos4), _f(out))
        return out

    def mreal(self, pos4, F4):
        U = np.zeros((self.N, 4), dtype=np.float32)
        lib().orc_mreal(ctypes.byref(self.cfg), ctypes.byref(self.prm), _f(self.table), _f(pos4), _f(F4), _f(self.nn), _f(self.head), _f(self.nl), _f(U))
        return U

    def mwave(self, pos4, F4, do_det=True, u_grid=None, noise_fac=0.0):
        U = np.zeros((self.N, 4), dtype=np.float32)
        lib().orc_mwave(ctypes.byref(self.cfg), ctypes.byref(self.prm), _f(pos4), _f(F4), _f(U), ctypes.c_int(1 if do_det else 0),
                        None if u_grid is None else _f(u_grid), ctypes.c_float(noise_fac))
        return U

    def lanczos(self, pos4, psi4, T, dt, m_in=2):
        U = np.zeros((self.N, 4), dtype=np.float32)
        m = ctypes.c_int(m_in); sn = ctypes.c_float(0)
        rc = lib().orc_lanczos(ctypes.byref(self.cfg), ctypes.byref(self.prm), _f(self.table), _f(pos4), _f(psi4), _f(self.nn), _f(self.head),
                               _f(self.nl), ctypes.c_float(T), ctypes.c_float(dt), ctypes.byref(m), _f(U), ctypes.byref(sn))
        assert rc == 0, rc
        return U, m.value, sn.value

    def velocity(self, pos4, F4, T, dt, u_particles=None, u_grid=None, m_in=2):
        U = np.zeros((self.N, 4), dtype=np.float32)
        m = ctypes.c_int(m_in)
        rc = lib().orc_velocity(ctypes.byref(self.cfg), ctypes.byref(self.prm), _f(self.table), _f(pos4), _f(F4), _f(self.nn), _f(self.head),
                                _f(self.nl), ctypes.c_float(T), ctypes.c_float(dt), None if u_particles is None else _f(u_particles),
                                None if u_grid is None else _f(u_grid), ctypes.byref(m), _f(U))
        assert rc == 0, rc
        return U, m.value

    def integrate(self, pos4, image3, vel4, dt, shear_rate=0.0):
        lib().orc_integrate(ctypes.byref(self.cfg), _f(pos4), _f(image3), _f(vel4), ctypes.c_float(dt), ctypes.c_float(shear_rate))


def dense_mobility(pos3, F3, L, xy=0.0, xi_d=None, tol=1e-12):
    """Double-precision dense Ewald RPY mobility U = M F (exact pi), the accuracy oracle."""
    pos3 = np.ascontiguousarray(pos3, dtype=np.float64); F3 = np.ascontiguousarray(F3, dtype=np.float64)
    N = len(pos3)
    Lx, Ly, Lz = (L, L, L) if np.isscalar(L) else L
    s = np.sqrt(-np.log(tol))
    if xi_d is None:
        xi_d = s / (0.5 * min(Lx, Ly, Lz) * 0.98)   # real-space sum converged inside the minimum image
    rc_d, kc_d = s / xi_d * 1.3, 2 * s * xi_d * 1.3
    U = np.zeros((N, 3), dtype=np.float64)
    lib().orc_dense_mobility(N, float(Lx), float(Ly), float(Lz), float(xy), _f(pos3), _f(F3), float(xi_d), float(rc_d), float(kc_d), _f(U))
    return U
