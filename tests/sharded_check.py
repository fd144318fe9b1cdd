"""Multi-GPU check (run under torchrun, one rank per GPU): slab-decomposed M.F against the single-GPU engine.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_check.py [N] [phi]
"""
import os, sys, time
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
for N, phi, xy in cases:
    L = util.box_length(N, phi)
    cfg = E.make_config(N, L, xy=xy, T=1.0, dt=1e-3, seed=1, xi=xi, error=error)
    pos = torch.from_numpy(util.lattice_positions(N, L, 0)).cuda(); F = torch.from_numpy(util.random_forces(N, 1)).cuda()
    sm = S.ShardedMobility(cfg)
    U = sm.mobility(pos, F)
    single = E.Engine(cfg)
    Uref = single.mobility(pos, F)
    torch.cuda.synchronize()
    l2, mx = util.rel_err(U.cpu().numpy(), Uref.cpu().numpy())
    # every rank must hold the same complete result
    chk = U.double().sum().reshape(1); allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    same = all(float(a) == float(allchk[0]) for a in allchk)
    def timeit(fn, n=10):
        fn(); torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); dist.barrier()
        return (time.perf_counter() - t0) / n * 1e6
    t_sh = timeit(lambda: sm.mobility(pos, F)); t_1 = timeit(lambda: single.mobility(pos, F))
    if rank == 0:
        print(f"N={N} grid={single.params.Nx} P={single.params.P} xy={xy} world={world}: sharded vs single rel L2 {l2:.2e} max {mx:.2e} identical_on_all_ranks={same} "
              f"| M.F sharded {t_sh:.0f} us, single GPU {t_1:.0f} us", flush=True)
    ok &= l2 < 2e-6 and mx < 5e-6 and same
    del sm, single
dist.barrier()
if rank == 0: print("SHARDED_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
