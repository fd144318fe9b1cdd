"""ctypes binding of the C ABI declared in include/pse_b200.h.

The shared library is built in-tree (pse_b200/libpse_b200.so) by __graft_entry__.build() /
`make -C pse_b200/csrc`.  There is no fallback: if the library is missing, importing this
module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSE_B200_LIB") or os.path.join(_HERE, "libpse_b200.so")  # override: tuning variants of the same ABI

PSE_OK = 0
PSE_EINVAL, PSE_ENODEVICE, PSE_ECUDA, PSE_EGRID, PSE_ENOMEM, PSE_EEIGEN, PSE_ECAPACITY = -1, -2, -3, -4, -5, -6, -7
PSE_FLAG_REF_PI = 1
PSE_FLAG_LIFT_GRID_CAP = 2
PSE_HOST_STATE_IN = 1
PSE_HOST_NO_STATE_OUT = 2


class pse_box(ctypes.Structure):
    _fields_ = [("Lx", ctypes.c_float), ("Ly", ctypes.c_float), ("Lz", ctypes.c_float), ("xy", ctypes.c_float)]


class pse_config(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_uint32),
        ("box", pse_box),
        ("xi", ctypes.c_float),
        ("error", ctypes.c_float),
        ("max_strain", ctypes.c_float),
        ("T", ctypes.c_float),
        ("dt", ctypes.c_float),
        ("seed", ctypes.c_uint32),
        ("flags", ctypes.c_uint32),
        ("r_buff", ctypes.c_float),
    ]


class pse_params(ctypes.Structure):
    _fields_ = [
        ("Nx", ctypes.c_int), ("Ny", ctypes.c_int), ("Nz", ctypes.c_int),
        ("P", ctypes.c_int), ("kmax", ctypes.c_int), ("ewald_n", ctypes.c_int),
        ("rcut", ctypes.c_float), ("dr", ctypes.c_float), ("gaussm", ctypes.c_float), ("eta", ctypes.c_float),
        ("hx", ctypes.c_float), ("hy", ctypes.c_float), ("hz", ctypes.c_float),
        ("self", ctypes.c_float), ("quadW", ctypes.c_float), ("prefac", ctypes.c_float), ("expfac", ctypes.c_float),
        ("seed_hashed", ctypes.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class pse_stats(ctypes.Structure):
    _fields_ = [
        ("nnz", ctypes.c_uint64), ("nnz_active", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64), ("fft_execs", ctypes.c_uint64), ("graph_launches", ctypes.c_uint64),
        ("nlist_builds", ctypes.c_uint64), ("lanczos_m", ctypes.c_int), ("lanczos_stepnorm", ctypes.c_float),
    ]


class pse_shard_info(ctypes.Structure):
    _fields_ = [
        ("rank", ctypes.c_int), ("world", ctypes.c_int), ("x0", ctypes.c_int), ("x1", ctypes.c_int),
        ("y0", ctypes.c_int), ("y1", ctypes.c_int), ("halo_left", ctypes.c_int), ("halo_right", ctypes.c_int),
        ("buffer_planes", ctypes.c_int), ("layer0", ctypes.c_int), ("layer1", ctypes.c_int), ("halo_layers", ctypes.c_int),
        ("row0", ctypes.c_uint32), ("row1", ctypes.c_uint32),
        ("a2a_send_bytes", ctypes.c_uint64 * 16), ("a2a_recv_bytes", ctypes.c_uint64 * 16),
        ("bytes_sent", ctypes.c_uint64), ("collectives", ctypes.c_uint64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["a2a_send_bytes"] = list(self.a2a_send_bytes)[: self.world]
        d["a2a_recv_bytes"] = list(self.a2a_recv_bytes)[: self.world]
        return d


# every symbol include/pse_b200.h declares: name -> (restype, argtypes)
_vp, _u32, _f, _i, _d = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float, ctypes.c_int, ctypes.c_double
_cfgp, _prmp = ctypes.POINTER(pse_config), ctypes.POINTER(pse_params)
class pse_pair_params(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("epsilon", ctypes.c_float), ("sigma", ctypes.c_float), ("r_cut", ctypes.c_float)]


PSE_PAIR_LJ, PSE_PAIR_WCA, PSE_PAIR_HARMONIC = 0, 1, 2

SYMBOLS = {
    "pse_derive_params": (_i, [_cfgp, _prmp]),
    "pse_ewald_table": (_i, [_cfgp, ctypes.POINTER(_f)]),
    "pse_shear_create": (_vp, [_i, ctypes.POINTER(_d), _i, _u32, _d]),
    "pse_shear_create_windowed": (_vp, [_vp, _vp]),
    "pse_shear_rate": (_d, [_vp, _u32]),
    "pse_shear_strain": (_d, [_vp, _u32]),
    "pse_shear_offset": (_u32, [_vp]),
    "pse_shear_destroy": (None, [_vp]),
    "pse_shear_variant_value": (_d, [_vp, _u32, _d, _d, _u32]),
    "pse_create": (_i, [_cfgp, _vp, ctypes.POINTER(_vp)]),
    "pse_destroy": (None, [_vp]),
    "pse_last_error": (ctypes.c_char_p, [_vp]),
    "pse_get_params": (_i, [_vp, _prmp]),
    "pse_set_box": (_i, [_vp, ctypes.POINTER(pse_box)]),
    "pse_wrap_positions": (_i, [_vp, _vp, _vp]),
    "pse_set_temperature": (_i, [_vp, _f]),
    "pse_set_lanczos_m": (_i, [_vp, _i]),
    "pse_get_lanczos_m": (_i, [_vp]),
    "pse_build_neighbors": (_i, [_vp, _vp]),
    "pse_neighbor_list": (_i, [_vp, _vp, _vp, _vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "pse_grid_index": (_i, [_vp, _vp, _vp]),
    "pse_mreal": (_i, [_vp, _vp, _vp, _vp]),
    "pse_mwave": (_i, [_vp, _vp, _vp, _vp]),
    "pse_mobility": (_i, [_vp, _vp, _vp, _vp]),
    "pse_velocity": (_i, [_vp, _vp, _vp, _vp, _u32, _vp, _vp, _u32, ctypes.POINTER(_i)]),
    "pse_step": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _f, ctypes.POINTER(_i)]),
    "pse_step_host": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _f, ctypes.POINTER(_i)]),
    "pse_step_host_async": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _f, _u32, ctypes.POINTER(_i)]),
    "pse_wait": (_i, [_vp]),
    "pse_host_prefetch_forces": (_i, [_vp, _vp]),
    "pse_pair_force": (_i, [_vp, _vp, ctypes.POINTER(pse_pair_params), _vp, _i]),
    "pse_get_stats": (_i, [_vp, ctypes.POINTER(pse_stats)]),
    "pse_shard_plan": (_i, [_cfgp, _i, _i, ctypes.POINTER(pse_shard_info)]),
    "pse_comm_unique_id": (_i, [ctypes.POINTER(ctypes.c_uint8)]),
    "pse_local_world_create": (_vp, [_i]),
    "pse_local_world_destroy": (None, [_vp]),
    "pse_shard_init": (_i, [_vp, _i, _i, ctypes.POINTER(ctypes.c_uint8), _vp]),
    "pse_shard_get_info": (_i, [_vp, ctypes.POINTER(pse_shard_info)]),
    "pse_set_profiling": (_i, [_vp, _i]),
    "pse_get_profile": (_i, [_vp, ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_uint64), _i]),
    "pse_profile_phase_name": (ctypes.c_char_p, [_i]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the PSE hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.pse_test_tridiag_sqrt_e1.restype = _i
    lib.pse_test_tridiag_sqrt_e1.argtypes = [_i, ctypes.POINTER(_d), ctypes.POINTER(_d), ctypes.POINTER(_d)]
    return lib


lib = load()
