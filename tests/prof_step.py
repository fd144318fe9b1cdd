"""Profiling driver (not a test): a few BD steps at the headline config for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pse_b200 import engine as E, _lib
from tests import util

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
phi = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = util.box_length(N, phi)
xi = float(os.environ.get("PSE_XI", 0.5)); error = float(os.environ.get("PSE_ERROR", 1e-3))
cfg = E.make_config(N, L, xi=xi, error=error, T=1.0, dt=1e-3, seed=1, flags=int(os.environ.get("PSE_FLAGS", 0)))
eng = E.Engine(cfg)
pos = torch.from_numpy(util.lattice_positions(N, L, 0)).cuda()
F = torch.from_numpy(util.random_forces(N, 1)).cuda()
img = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
eng.lanczos_m = 5
m = eng.step(pos, img, F, 0)
torch.cuda.synchronize()
noprof = bool(os.environ.get("PSE_NOPROF"))
if not noprof: eng.set_profiling(True)
import time
t0 = time.time()
for t in range(1, steps + 1):
    m = eng.step(pos, img, F, t)
torch.cuda.synchronize()
wall = (time.time() - t0) / steps * 1e3
if noprof:
    print(f"N={N} steps={steps} m={m} wall/step={wall:.3f} ms (no profiling)  {eng.stats()}")
    sys.exit(0)
prof = eng.profile()
tot = sum(v[0] for v in prof.values())
print(f"N={N} steps={steps} m={m} wall/step={wall:.3f} ms  sum(phases)/step={tot/steps:.3f} ms  {eng.stats()}")
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    if n: print(f"  {k:14s} {ms/steps:8.3f} ms/step  {n/steps:5.1f} spans/step  {ms/n*1e3:9.1f} us/span")
