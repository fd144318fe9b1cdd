"""Host-side mirror of the plugin API (no GPU): shear functions, variant, argument validation, and the host logic of the
slab-decomposed engine over gloo (world_size 2 and 3)."""
import math
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import pse_b200 as PSEv1
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Sys:
    def __init__(self, t=0):
        self.timestep = t

    def getCurrentTimeStep(self):
        return self.timestep


@pytest.fixture
def at_step_10():
    PSEv1.system._current = _Sys(10)
    yield
    PSEv1.system._current = None


def test_shear_function_classes_mirror_reference(at_step_10):
    sf = PSEv1.shear_function
    dt = 1e-3
    s = sf.steady(dt=dt, shear_rate=2.0)
    assert s.get_offset() == 10 and s.get_shear_rate(500) == 2.0 and s.get_strain(510) == pytest.approx(2.0 * 500 * dt)
    z = sf.steady(dt=0)                                     # "no shear" default of integrate.py:93
    assert z.get_shear_rate(99) == 0 and z.get_strain(99) == 0
    f = sf.sine(dt=dt, shear_rate=1.0, shear_freq=2.0, zero=4)
    assert f.get_offset() == 4
    assert f.get_shear_rate(4 + 125) == pytest.approx(math.cos(2 * 2 * 3.1415926536 * 0.125))
    c = sf.chirp(dt=dt, amplitude=0.1, omega_0=1.0, omega_f=10.0, periodT=2.0)
    assert c.get_strain(10) == 0 and abs(c.get_strain(1500)) <= 0.1
    w = sf.tukey_window(dt=dt, periodT=2.0, tukey_param=0.5)
    assert w.get_strain(10) == 0 and w.get_strain(10 + 1000) == 1 and 0 < w.get_strain(10 + 100) < 1
    ww = sf.windowed(c, w)
    t = 10 + 300
    assert ww.get_strain(t) == pytest.approx(c.get_strain(t) * w.get_strain(t))
    assert ww.get_shear_rate(t) == pytest.approx(c.get_shear_rate(t) * w.get_strain(t) + c.get_strain(t) * w.get_shear_rate(t))
    assert ww.get_offset() == c.get_offset()


def test_shear_function_validation_errors(at_step_10):
    sf = PSEv1.shear_function
    for bad in (lambda: sf.sine(dt=1e-3, shear_rate=0, shear_freq=1), lambda: sf.sine(dt=1e-3, shear_rate=1, shear_freq=-1),
                lambda: sf.tukey_window(dt=1e-3, periodT=1.0, tukey_param=0), lambda: sf.tukey_window(dt=1e-3, periodT=1.0, tukey_param=1.5),
                lambda: sf.steady(dt=1e-3, zero=-1), lambda: sf.steady(dt=1e-3, zero=11)):
        with pytest.raises(RuntimeError):
            bad()


def test_shear_variant_wraps_into_max_strain(at_step_10):
    f = PSEv1.shear_function.steady(dt=1e-2, shear_rate=1.0, zero=5)
    v = PSEv1.variant.shear_variant(f, total_timestep=200, max_strain=0.5)
    assert v.get_value(0) == 0 and v.get_value(30) == pytest.approx(0.25) and v.get_value(56) == pytest.approx(-0.49)
    assert v.get_value(10**6) == pytest.approx(0.0)
    with pytest.raises(RuntimeError):
        PSEv1.variant.shear_variant(f, total_timestep=0)


def test_integrator_refuses_to_run_without_cuda():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    s = PSEv1.system.System.__new__(PSEv1.system.System)     # no device arrays without a GPU
    s.N, s.box, s.dt, s.timestep = 10, PSEv1.system.Box(30.0), 1e-3, 0
    g = PSEv1.system.Group(s)
    with pytest.raises(RuntimeError):                         # PSEv1/integrate.py:51-53
        PSEv1.integrate.PSEv1(group=g, T=1.0)
    assert PSEv1.integrate.PSE is PSEv1.integrate.PSEv1


def test_nccl_id_hand_over_world_size_2_gloo(tmp_path):
    """The only thing the host layer does for the slab-decomposed engine (pse_b200/sharded.py): rank 0 creates the 128-byte
    NCCL unique id through the C ABI (pse_comm_unique_id binds the NCCL the process already loaded) and broadcasts it; every
    rank must end up with the same bytes.  Ring neighbours are mutual.  (CPU, gloo.)"""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from pse_b200 import sharded as S
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = bytes(S.nccl_unique_id())
        assert len(uid) == 128 and any(uid)
        got = [None] * world
        dist.all_gather_object(got, uid)
        assert got[0] == got[1]
        l, r = S.ring_peers(rank, world)
        assert l == r == 1 - rank                # two ranks: both neighbours are the other rank
        dist.barrier(); dist.destroy_process_group()
        open(os.path.join({str(tmp_path)!r}, f"ok_{{rank}}"), "w").write("ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert (tmp_path / "ok_0").exists() and (tmp_path / "ok_1").exists()


def test_shard_plan_all_to_all_is_consistent_world_3_gloo(tmp_path):
    """Slab decomposition of the multi-GPU step (pse_b200/sharded.py): the static plan every rank derives on its own must
    tile the grid and the cell layers, the per-peer transpose sizes must agree pairwise and drive a real (gloo, CPU)
    all_to_all_single there and back, ring neighbours must be mutual; too many ranks are refused on every rank alike."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from pse_b200 import engine as E, sharded as S
        from tests import util
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        cfg = E.make_config(100000, util.box_length(100000, 0.2))       # 125^3 grid, P = 6: uneven slabs
        info = S.plan(cfg, rank, world).as_dict()
        allinfo = [None] * world
        dist.all_gather_object(allinfo, info)
        assert allinfo[0]["x0"] == 0 and allinfo[-1]["x1"] == 125 and allinfo[0]["y0"] == 0 and allinfo[-1]["y1"] == 125
        assert allinfo[0]["layer0"] == 0
        for r in range(world - 1):
            assert allinfo[r]["x1"] == allinfo[r + 1]["x0"] and allinfo[r]["y1"] == allinfo[r + 1]["y0"]
            assert allinfo[r]["layer1"] == allinfo[r + 1]["layer0"]
        for r in range(world):
            i = allinfo[r]
            assert i["x1"] - i["x0"] >= max(i["halo_left"], i["halo_right"])          # halos reach the adjacent slab only
            assert i["x1"] - i["x0"] + i["halo_left"] + i["halo_right"] <= i["buffer_planes"] < 125
            for q in range(world):
                assert i["a2a_send_bytes"][q] == allinfo[q]["a2a_recv_bytes"][r]    # what r sends to q is what q expects from r
        send = [n // 4 for n in info["a2a_send_bytes"]]; recv = [n // 4 for n in info["a2a_recv_bytes"]]
        a = torch.cat([torch.full((n,), float(rank * 100 + q)) for q, n in enumerate(send)])
        b = torch.empty(sum(recv))
        dist.all_to_all_single(b, a, recv, send)
        off = 0
        for r, n in enumerate(recv):
            assert bool((b[off:off + n] == r * 100 + rank).all()); off += n
        back = torch.empty(sum(send))
        dist.all_to_all_single(back, b, send, recv)                     # the way back swaps the roles
        assert torch.equal(back, a)
        peers = [None] * world
        dist.all_gather_object(peers, S.ring_peers(rank, world))
        assert all(peers[peers[r][0]][1] == r and peers[peers[r][1]][0] == r for r in range(world))
        try:
            S.plan(E.make_config(1000, 34.7), rank, 8)                  # 36^3 grid: 8 slabs would be thinner than their halos
            raise SystemExit("expected an error")
        except E.PSEError:
            pass
        dist.barrier(); dist.destroy_process_group()
        open(os.path.join({str(tmp_path)!r}, f"ok_{{rank}}"), "w").write("ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29534")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
                          "--master-port", "29534", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-3000:]
    assert all((tmp_path / f"ok_{r}").exists() for r in range(3))


def test_shard_plan_headline_configs_on_eight_ranks():
    """The static decomposition of BASELINE.json configs[2] (240^3) and configs[4] (432^3, P = 8) over 8 ranks (host only): slabs tile the
    grid, every rank's local real-space buffer is a fraction of the grid (no rank allocates a full one), the transposes balance."""
    from pse_b200 import engine as E, sharded as S
    for N, phi, xi, error, grid in ((1000000, 0.3, 0.5, 1e-3, 240), (8000000, 0.4, 0.45, 1e-4, 432)):
        cfg = E.make_config(N, util.box_length(N, phi), xi=xi, error=error, r_buff=0.8)
        infos = [S.plan(cfg, r, 8).as_dict() for r in range(8)]
        assert infos[0]["x0"] == 0 and infos[-1]["x1"] == grid and infos[0]["y0"] == 0 and infos[-1]["y1"] == grid
        for r in range(7):
            assert infos[r]["x1"] == infos[r + 1]["x0"] and infos[r]["layer1"] == infos[r + 1]["layer0"]
        for r, i in enumerate(infos):
            nown = i["x1"] - i["x0"]
            assert abs(nown - grid / 8) <= 0.15 * grid / 8 and nown >= max(i["halo_left"], i["halo_right"])   # whole layers of cells: 9 or 10 of 79
            assert i["buffer_planes"] <= nown + i["halo_left"] + i["halo_right"] + 16 and i["buffer_planes"] < grid // 3
            assert sum(i["a2a_send_bytes"]) == sum(infos[q]["a2a_recv_bytes"][r] for q in range(8))
    with pytest.raises(E.PSEError):
        S.plan(E.make_config(1000000, util.box_length(1000000, 0.3)), 0, 17)      # more ranks than the communicator supports


def test_local_world_handles_and_abi_guards_without_gpu():
    """Host-only pieces of the slab-decomposed ABI: the in-process world handle, and null / range checks that must not need a device."""
    import ctypes
    from pse_b200 import _lib
    lib = _lib.lib
    w = lib.pse_local_world_create(3)
    assert w
    lib.pse_local_world_destroy(w)
    assert not lib.pse_local_world_create(0) and not lib.pse_local_world_create(17)
    info = _lib.pse_shard_info()
    assert lib.pse_shard_init(None, 0, 2, None, None) == _lib.PSE_EINVAL
    assert lib.pse_shard_get_info(None, ctypes.byref(info)) == _lib.PSE_EINVAL
    assert lib.pse_host_prefetch_forces(None, None) == _lib.PSE_EINVAL and lib.pse_wait(None) == _lib.PSE_EINVAL
    cfg = PSEv1.engine.make_config(100000, util.box_length(100000, 0.2))
    assert lib.pse_shard_plan(ctypes.byref(cfg), 3, 3, ctypes.byref(info)) == _lib.PSE_EINVAL      # rank outside the world


def test_system_save_load_roundtrip_cpu(tmp_path):
    """Restart file: positions, images, step counter, box, forces survive a save/load (no GPU needed)."""
    from pse_b200 import system as S
    rng = np.random.default_rng(3)
    s = S.System(rng.uniform(-5, 5, (50, 3)).astype(np.float32), S.Box(10.0, 12.0, 14.0, xy=0.25), device="cpu")
    s.timestep = 1234
    s.image[:, 0] = 2
    s.set_forces(rng.normal(size=(50, 3)).astype(np.float32))
    s.save(tmp_path / "r")
    t = S.System.load(tmp_path / "r", device="cpu")
    assert t.timestep == 1234 and (t.box.Lx, t.box.Ly, t.box.Lz, t.box.xy) == (10.0, 12.0, 14.0, 0.25)
    assert np.array_equal(t.pos.numpy(), s.pos.numpy()) and np.array_equal(t.image.numpy(), s.image.numpy())
    assert np.array_equal(t.net_force.numpy(), s.net_force.numpy())


def test_pair_provider_validation():
    from pse_b200 import pair, system as S
    s = S.set_current(S.System(np.zeros((4, 3), dtype=np.float32), S.Box(10.0), device="cpu"))
    with pytest.raises(RuntimeError):
        pair.lj(r_cut=0.0)
    w = pair.wca()
    assert s.forces == [w] and w.enabled
    w.disable()
    assert not w.enabled
