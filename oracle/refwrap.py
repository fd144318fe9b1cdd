"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/_ref/libpse_ref.so, i.e. the
reference's own CUDA kernels (compiled unmodified from /root/reference/PSEv1/*.cu by
oracle/Makefile) driven by oracle/ref_harness.cu.

Only tests/, __graft_entry__.smoke() and bench.py (--impl reference) may import this module.
The product (pse_b200/) never does.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libpse_ref.so")


class pse_ref_params(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int),
        ("Lx", ctypes.c_float), ("Ly", ctypes.c_float), ("Lz", ctypes.c_float), ("xy", ctypes.c_float),
        ("xi", ctypes.c_float), ("eta", ctypes.c_float), ("rcut", ctypes.c_float), ("dr", ctypes.c_float),
        ("ewald_n", ctypes.c_int), ("self", ctypes.c_float),
        ("Nx", ctypes.c_int), ("Ny", ctypes.c_int), ("Nz", ctypes.c_int), ("P", ctypes.c_int),
        ("hx", ctypes.c_float), ("hy", ctypes.c_float), ("hz", ctypes.c_float),
        ("error", ctypes.c_float),
    ]


def available():
    return os.path.exists(REF_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(REF_LIB)
        for n in ("pse_ref_setgridk", "pse_ref_spread", "pse_ref_contract", "pse_ref_mreal", "pse_ref_mwave",
                  "pse_ref_mobility", "pse_ref_velocity", "pse_ref_lanczos", "pse_ref_step", "pse_ref_set_noise_tables"):
            getattr(_lib, n).restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class Reference:
    """Holds the arrays Stokes owns in the reference (gridk, gridX/Y/Z, table; PSEv1/Stokes.cc:260-268,
    :325-326) and calls the reference drivers on caller-provided particle arrays."""

    def __init__(self, cfg, params, table_np, device="cuda"):
        import torch
        self.torch = torch
        p = params
        self.N = cfg.N
        self.cfg = cfg
        self.prm = pse_ref_params(N=cfg.N, Lx=cfg.box.Lx, Ly=cfg.box.Ly, Lz=cfg.box.Lz, xy=cfg.box.xy, xi=cfg.xi, eta=p.eta,
                                  rcut=p.rcut, dr=p.dr, ewald_n=p.ewald_n, self=p.self, Nx=p.Nx, Ny=p.Ny, Nz=p.Nz, P=p.P,
                                  hx=p.hx, hy=p.hy, hz=p.hz, error=cfg.error)
        self.G = p.Nx * p.Ny * p.Nz
        self.table = torch.from_numpy(table_np).to(device).contiguous()
        self.gridk = torch.zeros((self.G, 4), dtype=torch.float32, device=device)
        self.gX = torch.zeros((self.G, 2), dtype=torch.float32, device=device)
        self.gY = torch.zeros_like(self.gX)
        self.gZ = torch.zeros_like(self.gX)
        self.nlist = None
        self.m_lanczos = 2  # PSEv1/Stokes.cc:132
        self.seed_hashed = p.seed_hashed

    def set_tilt(self, xy):
        self.prm.xy = xy

    def set_neighbors(self, n_neigh, headlist, nlist):
        self.nlist = (n_neigh, headlist, nlist)

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"reference {what} failed with code {rc}")

    def set_noise_tables(self, u_particles, u_grid):
        self._ck(lib().pse_ref_set_noise_tables(_p(u_particles), _p(u_grid)), "set_noise_tables")

    def spread(self, pos, F, P=None, prefac=None, expfac=None, zero=True):
        if zero:
            self.gX.zero_(); self.gY.zero_(); self.gZ.zero_()
        P = self.prm.P if P is None else P
        self._ck(lib().pse_ref_spread(ctypes.byref(self.prm), _p(pos), _p(F), _p(self.gX), _p(self.gY), _p(self.gZ),
                                      ctypes.c_int(P), ctypes.c_float(prefac), ctypes.c_float(expfac)), "spread")
        return self.gX, self.gY, self.gZ

    def mreal(self, pos, F):
        U = self.torch.zeros_like(F)
        nn, head, nl = self.nlist
        self._ck(lib().pse_ref_mreal(ctypes.byref(self.prm), _p(pos), _p(U), _p(F), _p(self.table), _p(nn), _p(nl), _p(head)),
                 "mreal")
        return U

    def mwave(self, pos, F):
        U = self.torch.zeros_like(F)
        self._ck(lib().pse_ref_mwave(ctypes.byref(self.prm), _p(pos), _p(U), _p(F), _p(self.gridk), _p(self.gX), _p(self.gY),
                                     _p(self.gZ)), "mwave")
        return U

    def mobility(self, pos, F):
        U = self.torch.zeros_like(F)
        nn, head, nl = self.nlist
        self._ck(lib().pse_ref_mobility(ctypes.byref(self.prm), _p(pos), _p(U), _p(F), _p(self.table), _p(self.gridk),
                                        _p(self.gX), _p(self.gY), _p(self.gZ), _p(nn), _p(nl), _p(head)), "mobility")
        return U

    def velocity(self, pos, F, T, dt, timestep):
        U = self.torch.zeros_like(F)
        nn, head, nl = self.nlist
        m = ctypes.c_int(self.m_lanczos)
        self._ck(lib().pse_ref_velocity(ctypes.byref(self.prm), _p(pos), _p(U), _p(F), _p(self.table), _p(self.gridk),
                                        _p(self.gX), _p(self.gY), _p(self.gZ), _p(nn), _p(nl), _p(head), ctypes.c_float(T),
                                        ctypes.c_float(dt), ctypes.c_uint(timestep & 0xFFFFFFFF),
                                        ctypes.c_uint(self.seed_hashed), ctypes.byref(m)), "velocity")
        self.m_lanczos = m.value
        return U

    def lanczos(self, psi, pos, T, dt):
        U = self.torch.zeros_like(psi)
        nn, head, nl = self.nlist
        m = ctypes.c_int(self.m_lanczos)
        self._ck(lib().pse_ref_lanczos(ctypes.byref(self.prm), _p(psi), _p(pos), _p(U), _p(self.table), _p(nn), _p(nl), _p(head),
                                       ctypes.c_float(T), ctypes.c_float(dt), ctypes.byref(m)), "lanczos")
        self.m_lanczos = m.value
        return U

    def step(self, pos, vel, accel, image, F, T, dt, timestep, shear_rate=0.0, sync=True):
        nn, head, nl = self.nlist
        m = ctypes.c_int(self.m_lanczos)
        self._ck(lib().pse_ref_step(ctypes.byref(self.prm), _p(pos), _p(vel), _p(accel), _p(image), _p(F), _p(self.table),
                                    _p(self.gridk), _p(self.gX), _p(self.gY), _p(self.gZ), _p(nn), _p(nl), _p(head),
                                    ctypes.c_float(T), ctypes.c_float(dt), ctypes.c_uint(timestep & 0xFFFFFFFF),
                                    ctypes.c_uint(self.seed_hashed), ctypes.byref(m), ctypes.c_float(shear_rate),
                                    ctypes.c_int(1 if sync else 0)), "step")
        self.m_lanczos = m.value
        return m.value
