"""Driver for compute-sanitizer (memcheck / racecheck / initcheck) on the small config: two BD steps, all kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pse_b200 import engine as E
from tests import util

N = 3000
L = util.box_length(N, 0.15)
for xy, graph in ((0.0, "1"), (0.3, "0")):
    os.environ["PSE_GRAPH"] = graph
    eng = E.Engine(E.make_config(N, L, xy=xy, T=1.0, dt=1e-3, seed=2))
    pos = torch.from_numpy(util.random_positions(N, L, 0)).cuda(); F = torch.from_numpy(util.random_forces(N, 1)).cuda()
    img = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    for t in range(3):
        m = eng.step(pos, img, F, t, shear_rate=0.5)
    U = eng.mobility(pos, F)
    nn, head, nl = eng.neighbor_list()
    torch.cuda.synchronize()
    print("ok", xy, m, float(U.abs().max()), int(nn.sum()), eng.stats()["graph_launches"])
    eng.close()

# pipelined host entry point (copy streams, control-word kernels)
import numpy as np
eng = E.Engine(E.make_config(N, L, T=1.0, dt=1e-3, seed=2))
hp = [util.random_positions(N, L, 0) for _ in range(2)]; hi = [np.zeros((N, 3), dtype=np.int32) for _ in range(2)]
hF = util.random_forces(N, 1)
for t in range(4):
    eng.step_host_async(hp[t & 1], hi[t & 1], hF, t, state_in=(t == 0))
eng.wait()
print("ok host-async", float(np.abs(hp[1]).max()))
eng.close()

# slab-decomposed step as two virtual ranks (peer-memory kernels with plain pointers, host barrier)
from pse_b200 import sharded as S
N = 12000   # (two slabs need a grid of at least ~60 planes)
L = util.box_length(N, 0.15)
cfg = E.make_config(N, L, xy=0.2, T=1.0, dt=1e-3, seed=2)
lw = S.LocalWorld(cfg, 2)
pos = torch.from_numpy(util.random_positions(N, L, 0)).cuda(); F = torch.from_numpy(util.random_forces(N, 1)).cuda()
def work(r, e):
    q, im = pos.clone(), torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    for t in range(2):
        m = e.step(q, im, F, t, shear_rate=0.5)
    return m, float(e.mobility(q, F).abs().max())
print("ok sharded", lw.run(work))
lw.close()
