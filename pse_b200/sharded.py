"""Multi-GPU deterministic mobility U = M F: slab decomposition over the ranks of one node.

One process per GPU (torch.distributed, NCCL over NVLink).  Particle data are replicated — every rank passes
the same positions and forces — and the work is sharded (include/pse_b200.h, pse_shard_*): x-slabs of the
Fourier grid with two all-to-all transposes around the k-space pass, one neighbour exchange of P-1 halo planes,
contiguous row ranges of the real-space SpMV, and a final all-reduce of the partial velocities.

The reference is single-GPU (PSEv1/Stokes.cc:104); this is new work.  The host-side index
logic (split sizes, peers) is pure Python and is exercised on CPU with gloo in tests/test_host_api.py.
"""
import ctypes

from . import _lib
from ._lib import lib
from .engine import Engine, PSEError, _check4, _ptr


def split_sizes(info):
    """(send, recv) element counts per peer of the forward all-to-all, in floats; the way back swaps them."""
    w = info.world
    return [int(info.a2a_send_floats[q]) for q in range(w)], [int(info.a2a_recv_floats[q]) for q in range(w)]


def plan(cfg, rank, world):
    """The decomposition of `cfg` over `world` ranks as seen by `rank` (host only, no GPU)."""
    info = _lib.pse_shard_info()
    rc = lib.pse_shard_plan(ctypes.byref(cfg), rank, world, ctypes.byref(info))
    if rc != _lib.PSE_OK:
        raise PSEError(rc, "pse_shard_plan failed (more ranks than x-tiles of the Fourier grid?)")
    return info


def halo_peers(rank, world):
    """(destination of my first P-1 planes, source of the planes that follow my slab): periodic in x."""
    return (rank - 1) % world, (rank + 1) % world


class ShardedMobility:
    def __init__(self, cfg, group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch, self.group = dist, torch, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.eng = Engine(cfg)
        self.N = cfg.N
        self.info = _lib.pse_shard_info()
        rc = lib.pse_shard_setup(self.eng._h, self.rank, self.world, ctypes.byref(self.info))
        if rc != _lib.PSE_OK:
            raise PSEError(rc, lib.pse_last_error(self.eng._h).decode())
        self.send_sizes, self.recv_sizes = split_sizes(self.info)
        dev = torch.device("cuda", torch.cuda.current_device())
        f32 = dict(dtype=torch.float32, device=dev)
        self.buf_a = torch.empty(max(sum(self.send_sizes), 1), **f32)   # x-slab side of the transposes
        self.buf_b = torch.empty(max(sum(self.recv_sizes), 1), **f32)   # y-slab side
        self.halo_out = torch.empty(int(self.info.halo_floats), **f32)
        self.halo_in = torch.empty(int(self.info.halo_floats), **f32)

    def _ck(self, rc):
        if rc != _lib.PSE_OK:
            raise PSEError(rc, lib.pse_last_error(self.eng._h).decode())

    def mobility(self, pos, F):
        """U = M F; `pos`, `F` identical on every rank; the result is complete on every rank."""
        torch, dist = self.torch, self.dist
        _check4(pos, self.N, "pos"); _check4(F, self.N, "F")
        h = self.eng._h
        U = torch.empty_like(F)
        self._ck(lib.pse_shard_fwd(h, _ptr(pos), _ptr(F), _ptr(self.buf_a)))
        na, nb = sum(self.send_sizes), sum(self.recv_sizes)
        dist.all_to_all_single(self.buf_b[:nb], self.buf_a[:na], self.recv_sizes, self.send_sizes, group=self.group)
        self._ck(lib.pse_shard_kspace(h, _ptr(self.buf_b), _ptr(self.buf_b)))
        dist.all_to_all_single(self.buf_a[:na], self.buf_b[:nb], self.send_sizes, self.recv_sizes, group=self.group)
        self._ck(lib.pse_shard_inv(h, _ptr(self.buf_a), _ptr(self.halo_out)))
        if self.world > 1:
            dst, src = halo_peers(self.rank, self.world)
            ops = [dist.P2POp(dist.isend, self.halo_out, dst, self.group), dist.P2POp(dist.irecv, self.halo_in, src, self.group)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        else:
            self.halo_in.copy_(self.halo_out)
        self._ck(lib.pse_shard_finish(h, _ptr(self.halo_in), _ptr(U)))
        dist.all_reduce(U, group=self.group)
        return U
