"""Strong scaling of the slab-decomposed deterministic mobility M.F (one suspension over all ranks of one node).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29520 tests/sharded_scaling.py

Prints one JSON line per configuration on rank 0: time per M.F (max over ranks, device-synchronised) and, when the
single-GPU engine of the same configuration fits and PSE_SCALING_CHECK != 0, the agreement with it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pse_b200 import engine as E, sharded as S
from tests import util

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
check = os.environ.get("PSE_SCALING_CHECK", "1") != "0"
configs = [(8000000, 0.4, 0.45, 1e-4, "BASELINE.json configs[4]: N=8M, phi=0.4, error 1e-4, xi=0.45 (432^3, P=8)"),
           (1000000, 0.3, 0.5, 1e-3, "BASELINE.json configs[2]: N=1M, phi=0.3, error 1e-3 (240^3, P=6)")]
for N, phi, xi, error, name in configs:
    L = util.box_length(N, phi)
    cfg = E.make_config(N, L, T=1.0, dt=1e-3, seed=1, xi=xi, error=error)
    pos = torch.from_numpy(util.lattice_positions(N, L, 0)).cuda(); F = torch.from_numpy(util.random_forces(N, 1)).cuda()
    sm = S.ShardedMobility(cfg)
    U = sm.mobility(pos, F)

    def timeit(fn, n):
        fn(); torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n * 1e6

    t_sh = timeit(lambda: sm.mobility(pos, F), 10)
    line = {"config": name, "N": N, "world": world, "mf_us_sharded": t_sh, "grid": int(sm.eng.params.Nx), "P": int(sm.eng.params.P)}
    if check and rank == 0:
        single = E.Engine(cfg)
        Uref = single.mobility(pos, F)
        torch.cuda.synchronize()
        l2, mx = util.rel_err(U.cpu().numpy(), Uref.cpu().numpy())
        single.mobility(pos, F); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            single.mobility(pos, F)
        torch.cuda.synchronize()
        line.update({"mf_us_single": (time.perf_counter() - t0) / 5 * 1e6, "rel_l2_vs_single": l2, "max_vs_single": mx})
        line["speedup"] = line["mf_us_single"] / t_sh
        del single
    dist.barrier()
    if rank == 0:
        print("SCALING " + json.dumps(line), flush=True)
    del sm
    torch.cuda.empty_cache()
dist.destroy_process_group()
