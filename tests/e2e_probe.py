"""Timing probe (not a test): device steps vs the host entry points at the headline config."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pse_b200 import engine as E
from tests import util
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = util.box_length(N, 0.3)
cfg = E.make_config(N, L, T=1.0, dt=1e-3, seed=1, r_buff=0.8)
stream = torch.cuda.Stream()
eng = E.Engine(cfg, stream=stream)
pos_np, F_np = util.lattice_positions(N, L, 0), util.random_forces(N, 100)
pos = torch.from_numpy(pos_np).cuda(); F = torch.from_numpy(F_np).cuda(); img = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
t = 0
with torch.cuda.stream(stream):
    for _ in range(5):
        eng.step(pos, img, F, t); t += 1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        eng.step(pos, img, F, t); t += 1
    torch.cuda.synchronize(); print(f"device      {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/step")
h_F = torch.from_numpy(F_np).pin_memory().numpy()
hp = [torch.from_numpy(pos.cpu().numpy()).pin_memory().numpy() for _ in range(2)]
hi = [torch.from_numpy(img.cpu().numpy()).pin_memory().numpy() for _ in range(2)]
for label in ("async", "async again", "sync"):
    eng.step_host_async(hp[0], hi[0], h_F, t, state_in=True); t += 1; eng.wait()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(20):
        if label == "sync": eng.step_host(hp[0], hi[0], h_F, t)
        else: eng.step_host_async(hp[i & 1], hi[i & 1], h_F, t)
        t += 1
    eng.wait(); torch.cuda.synchronize(); print(f"{label:11s} {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/step  builds={eng.stats()['nlist_builds']}")
