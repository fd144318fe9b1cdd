"""Scratch driver for the first GPU run: engine vs reference kernels on configs 1 and 2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pse_b200 import engine as E, _lib
from oracle import refwrap
from tests import util

def run(N, phi, seed=0, xy=0.0, lattice=False):
    L = util.box_length(N, phi)
    cfg = E.make_config(N, L, xy=xy, flags=_lib.PSE_FLAG_REF_PI, T=1.0, dt=1e-3, seed=1)
    eng = E.Engine(cfg)
    p = eng.params
    print(f"--- N={N} phi={phi} L={L:.3f} grid={p.Nx} P={p.P} eta={p.eta:.4f} xy={xy}")
    pos_np = util.lattice_positions(N, L, seed) if lattice else util.random_positions(N, L, seed)
    pos = torch.from_numpy(pos_np).cuda(); F = torch.from_numpy(util.random_forces(N, seed + 1)).cuda()
    t0 = time.time(); eng.build_neighbors(pos); torch.cuda.synchronize(); print("build", time.time() - t0, eng.stats())
    nn, head, nl = eng.neighbor_list()
    ref = refwrap.Reference(cfg, p, E.ewald_table(cfg))
    ref.set_neighbors(nn, head, nl)
    # grid index vs reference spread histogram
    gi = eng.grid_index(pos).cpu().numpy().astype(np.int64)
    ones = torch.zeros_like(F); ones[:, 0] = 1.0; ones[:, 1] = (torch.arange(N, device="cuda") % 1021 + 1).float()
    gX, gY, gZ = ref.spread(pos, ones, P=p.P, prefac=1.0, expfac=0.0)
    hist = np.zeros(p.Nx * p.Ny * p.Nz); histid = np.zeros_like(hist)
    ids = (np.arange(N) % 1021 + 1).astype(np.float64)
    for tx in range(p.P):
        for ty in range(p.P):
            for tz in range(p.P):
                ix = (gi[:, 0] + tx) % p.Nx; iy = (gi[:, 1] + ty) % p.Ny; iz = (gi[:, 2] + tz) % p.Nz
                lin = (ix * p.Ny + iy) * p.Nz + iz
                np.add.at(hist, lin, 1.0); np.add.at(histid, lin, ids)
    print("grid index: count mismatch nodes", int((hist != gX[:, 0].cpu().numpy()).sum()), "id-sum mismatch", int((histid != gY[:, 0].cpu().numpy()).sum()))
    for name, fe, fr in [("mreal", eng.mreal, ref.mreal), ("mwave", eng.mwave, ref.mwave), ("mobility", eng.mobility, ref.mobility)]:
        a = fe(pos, F); b = fr(pos, F); torch.cuda.synchronize()
        print(name, "rel L2 / max:", util.rel_err(a.cpu().numpy(), b.cpu().numpy()))
    # velocity with injected noise
    G = p.Nx * p.Ny * p.Nz
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    up = torch.rand((N, 3), device="cuda", generator=g); ug = torch.rand((G, 6), device="cuda", generator=g)
    # neutralise the nodes the reference visits twice: u = 0.5 maps to exactly 0
    ug3 = ug.view(p.Nx, p.Ny, p.Nz, 6)
    if p.Nz % 2 == 0: ug3[:, :, p.Nz // 2, :] = 0.5
    if p.Ny % 2 == 0: ug3[:, p.Ny // 2, 0, :] = 0.5
    ref.set_noise_tables(up, ug)
    for parts, nm in [(2, "wave noise"), (4, "lanczos"), (7, "full velocity")]:
        eng.lanczos_m = 2; ref.m_lanczos = 2
        a, m = eng.velocity(pos, F, timestep=3, u_particles=up, u_grid=ug, parts=parts)
        if parts == 7:
            b = ref.velocity(pos, F, 1.0, 1e-3, 3); mr = ref.m_lanczos
        elif parts == 4:
            a3 = torch.sqrt(torch.tensor(3.0)).item()
            psi = torch.zeros_like(F); psi[:, :3] = (up * 2 - 1) * 1.73205080757
            b = ref.lanczos(psi, pos, 1.0, 1e-3); mr = ref.m_lanczos
        else:
            # wave noise only: reference velocity with F = 0 minus lanczos part is messy; use T>0, F=0 full minus lanczos
            Z = torch.zeros_like(F)
            ref.m_lanczos = 2
            full0 = ref.velocity(pos, Z, 1.0, 1e-3, 3)
            psi = torch.zeros_like(F); psi[:, :3] = (up * 2 - 1) * 1.73205080757
            ref.m_lanczos = 2
            lz = ref.lanczos(psi, pos, 1.0, 1e-3)
            b = full0 - lz; mr = -1
            a, m = eng.velocity(pos, Z, timestep=3, u_particles=up, u_grid=ug, parts=2)
        torch.cuda.synchronize()
        print(nm, "m eng/ref", m, mr, "rel L2 / max:", util.rel_err(a.cpu().numpy(), b.cpu().numpy()))
    ref.set_noise_tables(None, None)
    # philox (non-injected) parity
    eng.lanczos_m = 2; ref.m_lanczos = 2
    a, m = eng.velocity(pos, F, timestep=7, parts=7); b = ref.velocity(pos, F, 1.0, 1e-3, 7)
    print("philox full velocity m", m, ref.m_lanczos, util.rel_err(a.cpu().numpy(), b.cpu().numpy()))
    # timing
    for name, fn in [("eng mobility", lambda: eng.mobility(pos, F)), ("ref mobility", lambda: ref.mobility(pos, F)),
                     ("eng velocity", lambda: eng.velocity(pos, F, timestep=9)), ("ref velocity", lambda: ref.velocity(pos, F, 1.0, 1e-3, 9))]:
        fn(); torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5): fn()
        torch.cuda.synchronize(); print(name, (time.time() - t0) / 5 * 1e3, "ms")
    print(eng.stats())

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "mid"
    run(1000, 0.1)
    run(1000, 0.1, xy=0.3)
    if mode in ("mid", "big"):
        run(100000, 0.2)
        run(100000, 0.3, xy=0.25, lattice=True)
    if mode == "big": run(1000000, 0.3, lattice=True)
