"""Thin object wrapper over the C ABI (include/pse_b200.h) for device arrays held by torch.

PyTorch is used for device memory and streams only; every numerical operation goes through
libpse_b200.so.  Tensors must be CUDA float32 [N,4] (positions, forces, velocities) or int32
[N,3] (images), contiguous — the reference's Scalar4 / int3 layouts (PSEv1/Stokes.cc:436-470).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import lib


class PSEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pse_b200 error {code}: {msg}")
        self.code = code


def make_config(N, L, xi=0.5, error=1e-3, max_strain=0.5, T=1.0, dt=1e-3, seed=0, xy=0.0, flags=0, r_buff=0.4):
    Lx, Ly, Lz = (L, L, L) if np.isscalar(L) else L
    return _lib.pse_config(N=int(N), box=_lib.pse_box(Lx, Ly, Lz, xy), xi=xi, error=error, max_strain=max_strain, T=T,
                           dt=dt, seed=int(seed) & 0xFFFFFFFF, flags=flags, r_buff=r_buff)


def derive_params(cfg):
    """Stokes::setParams without a GPU (PSEv1/Stokes.cc:129-319)."""
    p = _lib.pse_params()
    rc = lib.pse_derive_params(ctypes.byref(cfg), ctypes.byref(p))
    if rc not in (_lib.PSE_OK,):
        raise PSEError(rc, "pse_derive_params failed")
    return p


def ewald_table(cfg):
    """Real-space table as float32 [ewald_n+1, 4] (PSEv1/Stokes.cc:322-422)."""
    p = derive_params(cfg)
    out = np.zeros((p.ewald_n + 1, 4), dtype=np.float32)
    rc = lib.pse_ewald_table(ctypes.byref(cfg), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    if rc != _lib.PSE_OK:
        raise PSEError(rc, "pse_ewald_table failed")
    return out


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _check4(t, N, name):
    import torch
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (N, 4)):
        raise ValueError(f"{name} must be a contiguous CUDA float32 tensor of shape ({N}, 4)")


class Engine:
    """Owns one pse_engine handle."""

    def __init__(self, cfg, stream=None):
        import torch
        if not torch.cuda.is_available():
            raise PSEError(_lib.PSE_ENODEVICE, "no CUDA device: the PSE hot path has no CPU fallback")
        self.cfg = cfg
        self.N = cfg.N
        self._h = ctypes.c_void_p()
        s = ctypes.c_void_p(stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream)
        rc = lib.pse_create(ctypes.byref(cfg), s, ctypes.byref(self._h))
        if rc != _lib.PSE_OK:
            raise PSEError(rc, lib.pse_last_error(None).decode())
        self.params = _lib.pse_params()
        lib.pse_get_params(self._h, ctypes.byref(self.params))

    def close(self):
        if getattr(self, "_h", None) and lib is not None:  # (module globals are gone at interpreter shutdown)
            lib.pse_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc != _lib.PSE_OK:
            raise PSEError(rc, lib.pse_last_error(self._h).decode())

    # -- configuration
    def set_box(self, Lx, Ly, Lz, xy):
        b = _lib.pse_box(Lx, Ly, Lz, xy)
        self._ck(lib.pse_set_box(self._h, ctypes.byref(b)))

    def set_tilt(self, xy):
        b = self.cfg.box
        self.set_box(b.Lx, b.Ly, b.Lz, xy)
        self.cfg.box.xy = xy

    def wrap_positions(self, pos, image=None):
        """Re-image every particle into the current (possibly re-tilted) box, in place."""
        _check4(pos, self.N, "pos")
        self._ck(lib.pse_wrap_positions(self._h, _ptr(pos), _ptr(image)))

    def set_temperature(self, T):
        self._ck(lib.pse_set_temperature(self._h, T))

    @property
    def lanczos_m(self):
        return lib.pse_get_lanczos_m(self._h)

    @lanczos_m.setter
    def lanczos_m(self, m):
        self._ck(lib.pse_set_lanczos_m(self._h, int(m)))

    def stats(self):
        s = _lib.pse_stats()
        self._ck(lib.pse_get_stats(self._h, ctypes.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def set_profiling(self, on=True):
        self._ck(lib.pse_set_profiling(self._h, 1 if on else 0))

    def profile(self):
        """{phase: (total_ms, spans)} accumulated since set_profiling(True) (CUDA events on the engine stream)."""
        n = lib.pse_get_profile(self._h, None, None, 0)
        ms = (ctypes.c_double * n)(); calls = (ctypes.c_uint64 * n)()
        lib.pse_get_profile(self._h, ms, calls, n)
        return {lib.pse_profile_phase_name(i).decode(): (ms[i], calls[i]) for i in range(n)}

    # -- bit-exact outputs
    def build_neighbors(self, pos):
        _check4(pos, self.N, "pos")
        self._ck(lib.pse_build_neighbors(self._h, _ptr(pos)))

    def neighbor_list(self):
        """(n_neigh[N], headlist[N], nlist[nnz]) uint32 stored as int32 tensors, reference layout."""
        import torch
        nnz = ctypes.c_size_t()
        self._ck(lib.pse_neighbor_list(self._h, None, None, None, 0, ctypes.byref(nnz)))
        dev = torch.device("cuda", torch.cuda.current_device())
        nn = torch.empty(self.N, dtype=torch.int32, device=dev)
        head = torch.empty(self.N, dtype=torch.int32, device=dev)
        nl = torch.empty(max(nnz.value, 1), dtype=torch.int32, device=dev)
        self._ck(lib.pse_neighbor_list(self._h, _ptr(nn), _ptr(head), _ptr(nl), nl.numel(), ctypes.byref(nnz)))
        return nn, head, nl[: nnz.value]

    def grid_index(self, pos):
        import torch
        _check4(pos, self.N, "pos")
        out = torch.empty((self.N, 3), dtype=torch.int32, device=pos.device)
        self._ck(lib.pse_grid_index(self._h, _ptr(pos), _ptr(out)))
        return out

    # -- operators
    def _op(self, fn, pos, F):
        import torch
        _check4(pos, self.N, "pos"); _check4(F, self.N, "F")
        U = torch.zeros_like(F)   # (.w of the output is left untouched by the engine, as the reference does with vel.w)
        self._ck(fn(self._h, _ptr(pos), _ptr(F), _ptr(U)))
        return U

    def mreal(self, pos, F):
        return self._op(lib.pse_mreal, pos, F)

    def mwave(self, pos, F):
        return self._op(lib.pse_mwave, pos, F)

    def mobility(self, pos, F):
        return self._op(lib.pse_mobility, pos, F)

    def pair_force(self, pos, kind, epsilon=1.0, sigma=2.0, r_cut=0.0, out=None, accumulate=False):
        """Pair forces (F.xyz, per-particle energy) on the engine's neighbour list: the stand-in for the HOOMD pair
        potentials whose net_force the reference integrates (PSEv1/Stokes.cc:447,457)."""
        import torch
        _check4(pos, self.N, "pos")
        F = out if out is not None else torch.zeros((self.N, 4), dtype=torch.float32, device=pos.device)
        prm = _lib.pse_pair_params(int(kind), float(epsilon), float(sigma), float(r_cut))
        self._ck(lib.pse_pair_force(self._h, _ptr(pos), ctypes.byref(prm), _ptr(F), 1 if accumulate else 0))
        return F

    def velocity(self, pos, F, timestep=0, u_particles=None, u_grid=None, parts=7):
        import torch
        _check4(pos, self.N, "pos"); _check4(F, self.N, "F")
        U = torch.zeros_like(F)
        m = ctypes.c_int(0)
        self._ck(lib.pse_velocity(self._h, _ptr(pos), _ptr(F), _ptr(U), int(timestep) & 0xFFFFFFFF, _ptr(u_particles),
                                  _ptr(u_grid), parts, ctypes.byref(m)))
        return U, m.value

    def step(self, pos, image, F, timestep, shear_rate=0.0, vel=None):
        _check4(pos, self.N, "pos"); _check4(F, self.N, "F")
        m = ctypes.c_int(0)
        self._ck(lib.pse_step(self._h, _ptr(pos), _ptr(image), _ptr(F), _ptr(vel), int(timestep) & 0xFFFFFFFF,
                              float(shear_rate), ctypes.byref(m)))
        return m.value

    def step_host(self, pos_np, image_np, F_np, timestep, shear_rate=0.0, vel_np=None):
        """Same step through host numpy buffers (float32 [N,4], int32 [N,3]); updates them in place."""
        m = ctypes.c_int(0)
        vp = lambda a: ctypes.c_void_p(0 if a is None else a.ctypes.data)
        self._ck(lib.pse_step_host(self._h, vp(pos_np), vp(image_np), vp(F_np), vp(vel_np), int(timestep) & 0xFFFFFFFF,
                                   float(shear_rate), ctypes.byref(m)))
        return m.value

    def step_host_async(self, pos_np, image_np, F_np, timestep, shear_rate=0.0, vel_np=None, state_in=False, state_out=True):
        """Pipelined host step: the state stays on the device between calls, forces go up and the new state comes down on
        copy streams beside the compute.  `pos_np` / `image_np` are OUTPUTS (inputs too with state_in=True / on the first
        call); they are valid after wait() - alternate two sets of host arrays to keep the pipeline full."""
        m = ctypes.c_int(0)
        vp = lambda a: ctypes.c_void_p(0 if a is None else a.ctypes.data)
        self._ck(lib.pse_step_host_async(self._h, vp(pos_np), vp(image_np), vp(F_np), vp(vel_np), int(timestep) & 0xFFFFFFFF,
                                         float(shear_rate), (_lib.PSE_HOST_STATE_IN if state_in else 0) | (0 if state_out else _lib.PSE_HOST_NO_STATE_OUT),
                                         ctypes.byref(m)))
        return m.value

    def wait(self):
        self._ck(lib.pse_wait(self._h))

    def prefetch_forces(self, F_np):
        """Start uploading the forces of the NEXT step_host_async call (pass F_np=None there to consume them)."""
        self._ck(lib.pse_host_prefetch_forces(self._h, ctypes.c_void_p(F_np.ctypes.data)))
