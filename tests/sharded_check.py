"""Multi-GPU check (run under torchrun, one rank per GPU): the slab-decomposed engine against the single-GPU engine.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_check.py [N] [phi] [xi] [error]

Checks M.F, a velocity evaluation with injected noise and three full steps (positions, images), that every rank holds the
same bits, and prints device times of M.F and of a full step for both engines.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pse_b200 import engine as E, sharded as S
from tests import util

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
xi = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
error = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-3
cases = [(int(sys.argv[1]) if len(sys.argv) > 1 else 200000, float(sys.argv[2]) if len(sys.argv) > 2 else 0.3, 0.0)]
if len(sys.argv) <= 3:
    cases.append((60000, 0.2, 0.3))


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for N, phi, xy in cases:
    L = util.box_length(N, phi)
    cfg = E.make_config(N, L, xy=xy, T=1.0, dt=1e-3, seed=1, xi=xi, error=error, r_buff=0.8)
    pos = torch.from_numpy(util.lattice_positions(N, L, 0)).cuda(); F = torch.from_numpy(util.random_forces(N, 1)).cuda()
    sh = S.ShardedEngine(cfg)
    single = E.Engine(cfg)
    p = single.params
    res = {}
    for name, eng in (("sharded", sh), ("single", single)):
        gen = torch.Generator(device="cuda"); gen.manual_seed(3)
        up = torch.rand((N, 3), device="cuda", generator=gen); ug = torch.rand((p.Nx * p.Ny * p.Nz, 6), device="cuda", generator=gen)
        mf = eng.mobility(pos, F).clone()
        eng.lanczos_m = 5
        vel, m = eng.velocity(pos, F, 7, up, ug)
        del up, ug
        q, img = pos.clone(), torch.zeros((N, 3), dtype=torch.int32, device="cuda")
        for t in range(3):
            eng.step(q, img, F, t)
        torch.cuda.synchronize()
        res[name] = (mf, vel.clone(), m, q, img)
    e_mf = util.rel_err(res["sharded"][0].cpu().numpy(), res["single"][0].cpu().numpy())
    e_v = util.rel_err(res["sharded"][1].cpu().numpy(), res["single"][1].cpu().numpy())
    dpos = float((res["sharded"][3][:, :3] - res["single"][3][:, :3]).abs().max())
    same_img = bool(torch.equal(res["sharded"][4], res["single"][4]))
    chk = torch.stack([res["sharded"][0].double().sum(), res["sharded"][3].double().sum()]); allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    same = all(bool(torch.equal(a, allchk[0])) for a in allchk)
    qs, ims = pos.clone(), torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    q1, im1 = pos.clone(), torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    ts = [100]
    def step_sh():
        sh.step(qs, ims, F, ts[0]); ts[0] += 1
    def step_1():
        single.step(q1, im1, F, ts[0]); ts[0] += 1
    t_mf_sh = timeit(lambda: sh.mobility(pos, F)); t_mf_1 = timeit(lambda: single.mobility(pos, F))
    t_st_sh = timeit(step_sh, 20); t_st_1 = timeit(step_1, 20)
    info = sh.shard_info().as_dict()
    if os.environ.get("PSE_PROF"):
        sh.set_profiling(True)
        for _ in range(10):
            step_sh()
        prof = sh.profile(); sh.set_profiling(False)
        if rank == 0:
            print("phases (us/step, sharded rank 0):", {k: round(v[0] / 10 * 1e3, 1) for k, v in prof.items() if v[1]}, flush=True)
    if rank == 0:
        print(json.dumps({"N": N, "grid": int(p.Nx), "P": int(p.P), "xy": xy, "world": world, "mf_rel_l2": e_mf[0], "mf_rel_max": e_mf[1],
                          "vel_rel_l2": e_v[0], "vel_rel_max": e_v[1], "m": [res["sharded"][2], res["single"][2]], "pos_maxdiff_3steps": dpos,
                          "images_equal": same_img, "identical_on_all_ranks": same, "mf_us_sharded": t_mf_sh, "mf_us_single": t_mf_1,
                          "step_us_sharded": t_st_sh, "step_us_single": t_st_1, "halo_planes": [info["halo_left"], info["halo_right"]],
                          "buffer_planes": info["buffer_planes"], "halo_layers": info["halo_layers"]}), flush=True)
    # positions: a few units in the last place of a coordinate of size L / 2 (1.5e-5 at the 438-wide box of config 5)
    ok &= e_mf[0] < 5e-6 and e_mf[1] < 1e-5 and e_v[0] < 5e-6 and e_v[1] < 1e-5 and dpos < max(2e-5, 3e-7 * L) and same_img and same and res["sharded"][2] == res["single"][2]
    sh.close(); single.close()
    del sh, single, res
    torch.cuda.empty_cache()
dist.barrier()
if rank == 0: print("SHARDED_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
