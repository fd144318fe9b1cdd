"""Multi-GPU PSE: the whole Brownian-dynamics step slab-decomposed over the GPUs of one node.

One process per GPU.  After `pse_shard_init` (include/pse_b200.h) every operator of the engine keeps its signature:
each rank passes the same particle arrays and receives the same complete result.  Inside, a rank owns a contiguous range
of x layers of cells (its particles: neighbour list, pruning, SpMV rows, Lanczos vectors, spreading, interpolation) and
the x planes of the Fourier grid those layers cover; what crosses a slab face is exchanged by collectives the C++ engine
issues itself on its stream through the NCCL C API (vector halo rows per Lanczos product, one three-word (double) all-reduce per
iteration, grid halo planes, two all-to-all transposes, one all-gather of the velocities).  Python only hands over the
128-byte NCCL unique id, broadcast with torch.distributed.

The reference is single-GPU (PSEv1/Stokes.cc:104); this is new work.  The host-side decomposition
(`plan`) needs no GPU and is exercised on CPU with gloo in tests/test_host_api.py.
"""
import ctypes
import threading

from . import _lib
from ._lib import lib
from .engine import Engine, PSEError


def plan(cfg, rank, world):
    """The static decomposition of `cfg` over `world` ranks as seen by `rank` (host only, no GPU)."""
    info = _lib.pse_shard_info()
    rc = lib.pse_shard_plan(ctypes.byref(cfg), rank, world, ctypes.byref(info))
    if rc != _lib.PSE_OK:
        raise PSEError(rc, lib.pse_last_error(None).decode())
    return info


def ring_peers(rank, world):
    """(left, right) neighbours of a slab: periodic in x."""
    return (rank - 1) % world, (rank + 1) % world


def nccl_unique_id(group=None, device=None):
    """128-byte NCCL id created by rank 0 of `group` and broadcast to the others (uint8 tensor -> ctypes array)."""
    import torch
    import torch.distributed as dist
    buf = (ctypes.c_uint8 * 128)()
    if dist.get_rank(group) == 0:
        rc = lib.pse_comm_unique_id(buf)
        if rc != _lib.PSE_OK:
            raise PSEError(rc, lib.pse_last_error(None).decode())
    cuda = dist.get_backend(group) == "nccl"
    t = torch.tensor(list(buf), dtype=torch.uint8, device=device if cuda else "cpu")
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return (ctypes.c_uint8 * 128)(*t.cpu().tolist())


class ShardedEngine(Engine):
    """Engine that is rank `rank` of `world`: same methods as Engine, every collective issued inside the C++ library."""

    def __init__(self, cfg, group=None, stream=None, rank=None, world=None, local_world=None):
        super().__init__(cfg, stream)
        if local_world is not None:          # virtual ranks inside one process (tests): rank / world given explicitly
            self.rank, self.world = int(rank), int(world)
            rc = lib.pse_shard_init(self._h, self.rank, self.world, None, local_world)
        else:
            import torch
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            uid = nccl_unique_id(group, torch.device("cuda", torch.cuda.current_device())) if self.world > 1 else None
            rc = lib.pse_shard_init(self._h, self.rank, self.world, uid, None)
        self._ck(rc)

    def shard_info(self):
        info = _lib.pse_shard_info()
        self._ck(lib.pse_shard_get_info(self._h, ctypes.byref(info)))
        return info


ShardedMobility = ShardedEngine   # round-1 name (deterministic M.F only); the whole step is sharded now


class LocalWorld:
    """`world` virtual ranks on ONE GPU: one engine and one host thread per rank, collectives replaced by device copies and a
    host barrier inside the library (pse_local_world).  Parity-test vehicle for the multi-rank code path."""

    def __init__(self, cfg, world):
        import torch
        self.world = world
        self._w = ctypes.c_void_p(lib.pse_local_world_create(world))
        if not self._w:
            raise PSEError(_lib.PSE_EINVAL, "pse_local_world_create failed")
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        self.engines = []
        errs = [None] * world

        def make(r):
            try:
                self.engines_by_rank[r] = ShardedEngine(cfg, stream=self.streams[r], rank=r, world=world, local_world=self._w)
            except Exception as ex:  # noqa: BLE001
                errs[r] = ex
        self.engines_by_rank = [None] * world
        for r in range(world):   # (pse_shard_init itself has no collective in a local world)
            make(r)
        if any(errs):
            raise next(e for e in errs if e)
        self.engines = self.engines_by_rank

    def run(self, fn):
        """fn(rank, engine) on every rank concurrently (the collectives inside block until all ranks arrive)."""
        import torch
        torch.cuda.synchronize()
        out, errs = [None] * self.world, [None] * self.world

        def work(r):
            try:
                with torch.cuda.stream(self.streams[r]):
                    out[r] = fn(r, self.engines[r])
                self.streams[r].synchronize()
            except Exception as ex:  # noqa: BLE001
                errs[r] = ex
        th = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for ex in errs:
            if ex:
                raise ex
        return out

    def close(self):
        for e in self.engines:
            if e is not None:
                e.close()
        self.engines = []
        if self._w:
            lib.pse_local_world_destroy(self._w)
            self._w = None
