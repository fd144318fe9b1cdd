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
