"""Diagnostic (not a test): parity numbers at the headline size against the reference kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util
from tests.test_gpu_parity import System

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
big = System(N, util.box_length(N, 0.3), lattice=True, seed=0)
for name, fe, fr in [("mreal", big.eng.mreal, big.ref.mreal), ("mwave", big.eng.mwave, big.ref.mwave), ("mobility", big.eng.mobility, big.ref.mobility)]:
    print(name, util.rel_err(fe(big.pos, big.F).cpu().numpy(), fr(big.pos, big.F).cpu().numpy()))
up, ug = big.noise()
big.ref.set_noise_tables(up, ug)
for parts in (2, 4, 7):
    big.eng.lanczos_m = 2; big.ref.m_lanczos = 2
    Ue, m = big.eng.velocity(big.pos, big.F, timestep=3, u_particles=up, u_grid=ug, parts=parts)
    if parts == 7:
        Ur = big.ref.velocity(big.pos, big.F, big.T, big.dt, 3)
    elif parts == 4:
        psi = torch.zeros_like(big.F); psi[:, :3] = (up * 2 - 1) * 1.73205080757
        Ur = big.ref.lanczos(psi, big.pos, big.T, big.dt)
    else:
        Z = torch.zeros_like(big.F)
        full0 = big.ref.velocity(big.pos, Z, big.T, big.dt, 3)
        psi = torch.zeros_like(big.F); psi[:, :3] = (up * 2 - 1) * 1.73205080757
        big.ref.m_lanczos = 2
        Ur = full0 - big.ref.lanczos(psi, big.pos, big.T, big.dt)
        Ue, m = big.eng.velocity(big.pos, Z, timestep=3, u_particles=up, u_grid=ug, parts=2)
    print("parts", parts, "m", m, big.ref.m_lanczos, util.rel_err(Ue.cpu().numpy(), Ur.cpu().numpy()))
