"""Minimal standalone particle / box shim that stands in for the HOOMD objects the reference plugin is
handed (`hoomd.context.current.system_definition`, `group.all()`, `ParticleData`, `BoxDim`;
PSEv1/integrate.py:58-86, PSEv1/Stokes.cc:436-470).  It owns the device arrays in the reference's
layouts and a step counter; it does not try to be HOOMD."""
import numpy as np


class Box:
    """Periodic box centred at the origin, sheared in xy (HOOMD BoxDim subset)."""

    def __init__(self, Lx, Ly=None, Lz=None, xy=0.0):
        self.Lx = float(Lx); self.Ly = float(Lx if Ly is None else Ly); self.Lz = float(Lx if Lz is None else Lz)
        self.xy = float(xy)


class Group:
    """`hoomd.group.all()` stand-in.  The reference only works for the group of all particles."""

    def __init__(self, system):
        self.system = system
        self.cpp_group = self


class System:
    """Particle data on one GPU: positions float4 (x,y,z,type), net_force float4 (F, pe), image int3."""

    def __init__(self, positions, box, device="cuda"):
        import torch
        pos = np.asarray(positions, dtype=np.float32)
        if pos.ndim != 2 or pos.shape[1] not in (3, 4):
            raise ValueError("positions must be [N,3] or [N,4]")
        N = pos.shape[0]
        p4 = np.zeros((N, 4), dtype=np.float32); p4[:, : pos.shape[1]] = pos
        self.N = N
        self.box = box if isinstance(box, Box) else Box(*np.atleast_1d(box))
        self.pos = torch.from_numpy(p4).to(device)
        self.net_force = torch.zeros((N, 4), dtype=torch.float32, device=device)
        self.vel = torch.zeros((N, 4), dtype=torch.float32, device=device)
        self.image = torch.zeros((N, 3), dtype=torch.int32, device=device)
        self.timestep = 0
        self.dt = None          # set by integrate.mode_standard
        self.integrator = None  # set by integrate.PSEv1
        self.updaters = []
        self.forces = []        # pair-force providers (pse_b200.pair); their sum replaces net_force every step
        self.constant_force = None  # set_forces(): an external force added to the providers' sum

    def all(self):
        return Group(self)

    def getCurrentTimeStep(self):
        return self.timestep

    def set_forces(self, F):
        """External (constant) force per particle, [N,3] or [N,4]; kept across steps."""
        import torch
        F = torch.as_tensor(F, dtype=torch.float32, device=self.pos.device)
        self.net_force[:, : F.shape[1]] = F
        self.constant_force = self.net_force.clone()   # kept whether or not a pair provider is registered yet

    def _compute_forces(self):
        """HOOMD's ForceCompute pass: net_force = external force + sum of the enabled pair providers."""
        active = [f for f in self.forces if f.enabled]
        if not active:
            return
        eng = self.integrator.cpp_method
        if self.constant_force is None:
            self.net_force.zero_()
        else:
            self.net_force.copy_(self.constant_force)
        for f in active:
            f.compute(eng, self.pos, self.net_force, True)

    # -- restart / trajectory: positions, images, step counter, box, Lanczos m; the RNG is stateless
    def save(self, path):
        import numpy as np
        m = self.integrator.cpp_method.lanczos_m if self.integrator is not None and hasattr(self.integrator, "cpp_method") else 0
        np.savez(path, pos=self.pos.cpu().numpy(), image=self.image.cpu().numpy(), timestep=self.timestep,
                 box=np.array([self.box.Lx, self.box.Ly, self.box.Lz, self.box.xy]), lanczos_m=m,
                 net_force=self.net_force.cpu().numpy())

    @classmethod
    def load(cls, path, device="cuda"):
        import numpy as np, torch
        d = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        b = d["box"]
        s = cls(d["pos"], Box(float(b[0]), float(b[1]), float(b[2]), float(b[3])), device=device)
        s.image = torch.from_numpy(d["image"]).to(device)
        s.net_force = torch.from_numpy(d["net_force"]).to(device)
        s.timestep = int(d["timestep"])
        s.restart_lanczos_m = int(d["lanczos_m"])
        return s

    def run(self, nsteps):
        """`hoomd.run(n)`: updaters (box tilt) then the integrator, once per step."""
        if self.integrator is None:
            raise RuntimeError("no integration method set")
        for _ in range(int(nsteps)):
            for u in self.updaters:
                u(self.timestep)
            self._compute_forces()
            self.integrator.integrate_step(self.timestep)
            self.timestep += 1


class box_resize:
    """`hoomd.update.box_resize(xy=variant)` stand-in: sets the box tilt from a variant every step
    (the plugin itself never changes the tilt)."""

    def __init__(self, system, xy):
        self.system, self.xy = system, xy
        system.updaters.append(self)

    def __call__(self, timestep):
        val = float(self.xy.get_value(timestep))
        changed = val != self.system.box.xy
        self.system.box.xy = val
        if self.system.integrator is not None:
            self.system.integrator.set_tilt(val)
            # HOOMD's BoxResizeUpdater re-images the particles after every box change; the wrapped-strain variant jumps
            # from +max_strain to -max_strain (PSEv1/VariantShearFunction.cc:34-43), which leaves up to an eighth of the
            # particles outside the primary cell of the new box
            if changed:
                self.system.integrator.cpp_method.wrap_positions(self.system.pos, self.system.image)


_current = None


def set_current(system):
    global _current
    _current = system
    return system


def current():
    if _current is None:
        raise RuntimeError("no current system: call pse_b200.system.set_current(System(...)) first")
    return _current
