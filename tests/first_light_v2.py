"""GPU first light for the spread2 / interp2 kernels (run by hand under gpurun, not a pytest file):
each variant of the wave path against the round-1 tile-owned kernels on the same inputs."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util  # noqa: E402


def engine(cfg, **env):
    from pse_b200 import engine as E
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return E.Engine(cfg)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def main():
    from pse_b200 import engine as E
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    cases = [  # N, phi, error, xy, Lfac
        (20000, 0.2, 1e-3, 0.0, (1, 1, 1)),
        (20000, 0.2, 1e-3, 0.3, (1, 1, 1)),
        (30000, 0.25, 1e-4, 0.2, (1, 1, 1)),       # P = 8
        (100000, 0.2, 1e-3, 0.0, (1, 1, 1)),       # 125^3: odd grid -> scalar reductions, partial tiles
        (6000, 0.1, 1e-3, -0.2, (1.0, 1.3, 0.8)),  # non-cubic
        (20000, 0.2, 4e-4, 0.1, (1, 1, 1)),        # P = 7
    ]
    ok = True
    for N, phi, error, xy, Lfac in cases:
        L0 = util.box_length(N, phi)
        L = tuple(L0 * f for f in Lfac) if Lfac != (1, 1, 1) else L0
        cfg = E.make_config(N, L, xy=xy, error=error, T=1.0, dt=1e-3, seed=3)
        pos_np = util.random_positions(N, L0, 2)
        pos_np[:, 1] *= Lfac[1]; pos_np[:, 2] *= Lfac[2]
        pos = torch.from_numpy(pos_np).cuda(); F = torch.from_numpy(util.random_forces(N, 4)).cuda()
        base = engine(cfg, PSE_WAVE="v1")
        p = base.params
        Ub = base.mwave(pos, F).clone()
        variants = [("v2", dict())]
        for name, env in variants:
            eng = engine(cfg, **env)
            U = eng.mwave(pos, F)
            torch.cuda.synchronize()
            l2, mx = util.rel_err(U.cpu().numpy(), Ub.cpu().numpy())
            Uv, _ = eng.velocity(pos, F, timestep=4, parts=3)
            Ubv, _ = base.velocity(pos, F, timestep=4, parts=3)
            l2v, mxv = util.rel_err(Uv.cpu().numpy(), Ubv.cpu().numpy())
            good = l2 < 2e-6 and mx < 2e-6 and l2v < 2e-6 and mxv < 2e-6
            ok &= good
            print(f"N={N} grid={p.Nx}x{p.Ny}x{p.Nz} P={p.P} xy={xy} {name:22s} mwave l2={l2:.2e} max={mx:.2e}  det+noise l2={l2v:.2e} max={mxv:.2e} {'OK' if good else 'FAIL'}",
                  flush=True)
            eng.close()
        base.close()
    print("FIRST LIGHT", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
