#!/usr/bin/env python3
"""bench.py — throughput of one full PSE Brownian-dynamics step (BASELINE.json metric) on N GPUs of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (real-space RPY + wave-space spread/FFT/scale/interpolate, deterministic and
stochastic parts, Lanczos, Euler update) over one synthetic suspension.
Workload at every N: BASELINE.json configs[2] — ONE suspension of 1,000,000 spheres, phi = 0.3, error 1e-3, xi = 0.5,
kT = 1, dt = 1e-3.  N > 1 slab-decomposes that one suspension over the N ranks (strong scaling; pse_b200/sharded.py);
the throughput of N independent replicas is reported as a side key, never as `value`.  With 8 ranks (or --config5) a
`config5` block times BASELINE.json configs[4] (N = 8M, phi = 0.4, error 1e-4, 432^3) the same way.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline (+ per-kernel table), cpu_baseline, e2e,
gpu_launches, clocks, phases, ms_per_step_median.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import util  # noqa: E402  (synthetic suspensions shared with the tests)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N", type=int, default=1000000)
    ap.add_argument("--phi", type=float, default=0.3)
    ap.add_argument("--error", type=float, default=1e-3)
    ap.add_argument("--xi", type=float, default=0.5)
    # Verlet buffer of the engine's list.  Measured at config 3 (profiles/r2_notes.md): 1.6 gives 3 % more steps/s (rebuilds every ~10
    # steps instead of ~3) but 13 % slower stand-alone M.F calls (every call prunes the longer rows); 0.8 balances the two metrics.
    ap.add_argument("--r-buff", type=float, default=0.8)
    ap.add_argument("--ref-r-buff", type=float, default=0.4, help="buffer of the list handed to the reference arm (HOOMD 2.x default r_buff)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-N", type=int, default=1000000)
    ap.add_argument("--config5", action="store_true", help="also time BASELINE.json configs[4] (N = 8M, 432^3); default with 8 ranks")
    ap.add_argument("--no-replicas", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        # NVML in-process when available (a spawned nvidia-smi takes driver locks and was seen to stall the reference
        # arm's per-step cudaMalloc/cudaFree by up to a second); nvidia-smi otherwise
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
            while not self.stop_flag:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])
                time.sleep(0.04)
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(args):
    """CPU port (oracle/pse_oracle.c, C + OpenMP: cell-list real-space sum with the same table, same
    spread/FFT/scale/interpolate, Lanczos with the tridiagonal solve through LAPACKE_spteqr as PSEv1/Brownian.cu:540)
    timed on this box's host cores AT THE HEADLINE SIZE: the deterministic M.F and one full velocity evaluation + Euler
    update, same N / phi / xi / error.  Bounded: two operator evaluations, 10-30 s of CPU work."""
    from oracle import oraclewrap as O
    Ns = min(args.cpu_sample_N, args.N)
    L = util.box_length(Ns, args.phi)
    lapacke = O.use_lapacke(True)
    o = O.Oracle(Ns, L, xi=args.xi, error=args.error, ref_pi=False)
    pos, F = util.lattice_positions(Ns, L, 0), util.random_forces(Ns, 1)
    rng = np.random.default_rng(2)
    G = o.prm.Nx * o.prm.Ny * o.prm.Nz
    up, ug = rng.random((Ns, 3), dtype=np.float32), rng.random((G, 6), dtype=np.float32)
    t0 = time.perf_counter()
    o.neighbors(pos, o.prm.rcut + args.r_buff, brute=False)
    t_nl = time.perf_counter() - t0
    t0 = time.perf_counter()
    o.mreal(pos, F); o.mwave(pos, F)
    t_mf = time.perf_counter() - t0
    t0 = time.perf_counter()
    U, m = o.velocity(pos, F, 1.0, 1e-3, u_particles=up, u_grid=ug, m_in=2)
    img = np.zeros((Ns, 3), dtype=np.int32)
    o.integrate(pos, img, U, 1e-3)
    t_step = time.perf_counter() - t0
    cores = O.lib().orc_num_threads()
    scale = Ns / args.N
    return {"value": (1.0 / t_step) * scale, "unit": "steps/s", "cores": cores, "kind": "port", "mf_us": t_mf * 1e6 / scale,
            "sample": f"at N={Ns}, phi={args.phi}, grid {o.prm.Nx}^3 on {cores} host threads: deterministic M.F {t_mf:.2f} s; 1 full step (M.F + Brownian, "
                      f"Lanczos m={m}, tridiagonal solve via {'LAPACKE_spteqr (scipy OpenBLAS)' if lapacke else 'Jacobi (LAPACKE not found)'}) {t_step:.2f} s "
                      f"(+{t_nl:.2f} s neighbour list, not counted)" + ("" if Ns == args.N else f"; scaled by {Ns}/{args.N}")}


# per-kernel roofline table: algorithmic bytes over the CUDA-event time of the phase, and the non-HBM unit the
# committed ncu captures show each kernel is actually bound by (profiles/r2_summary.md)
BOUND_BY = {
    "spread": "shared-memory data pipe (conflict-free residue-owned RMW) + red.global vector merge",
    "interp": "shared-memory data pipe / L2 window staging",
    "fft_r2c": "instruction issue (shared-memory radix <= 5 butterflies)",
    "scale": "instruction issue (x FFT both ways + B(k) projection in one kernel)",
    "fft_c2r": "instruction issue (shared-memory radix <= 5 butterflies)",
    "lanczos_spmv": "instruction issue + L1 gathers (92 instructions per pair)",
    "prune": "L1 gathers", "wave_bin": "HBM (W records written once, read by spread and interp)",
}


def kernel_table(phases, N, G, nnz, nnz_stored, P, m, peak):
    rs = 12 + ((P * P + P) + 3) // 4 * 4   # floats per W record
    alg = {
        "spread": 4.0 * rs * N + 12.0 * G,                 # W records in, three grids out (zero fill is part of the phase)
        "interp": 4.0 * rs * N + 12.0 * G + 16.0 * N,      # W records + grids in, velocities out
        "fft_r2c": 2 * 24.0 * G - 12.0 * G,                # z pass (12G in, ~12G out) + y pass (12G + 12G)
        "scale": 24.0 * G,                                 # fused x pass: spectrum in, spectrum out
        "fft_c2r": 2 * 24.0 * G - 12.0 * G,
        "lanczos_spmv": 56.0 * N + 4.0 * nnz,
        "prune": 16.0 * N + 4.0 * nnz_stored + 4.0 * nnz,
        "wave_bin": 16.0 * N + 4.0 * rs * N,
    }
    out = {}
    for k, b in alg.items():
        if k in phases and phases[k]["us_per_launch"]:
            us = phases[k]["ms_per_step"] * 1e3 / max(phases[k]["launches_per_step"], 1e-9) if k == "lanczos_spmv" else phases[k]["ms_per_step"] * 1e3
            out[k] = {"us": us, "algorithmic_MB": b / 1e6, "GBps": b / us / 1e3, "frac": b / us / 1e3 / peak, "bound_by": BOUND_BY.get(k)}
    return out


def main():
    args = parse()
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the PSE hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.impl == "reference" and rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    from pse_b200 import engine as E

    N, phi = args.N, args.phi
    L = util.box_length(N, phi)
    T, dt = 1.0, 1e-3
    K, W = args.steps, max(args.warmup, 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(eng, stream, pos, img, F, step0, n):
        """n steps, one CUDA event after each on the engine's stream: (total ms, per-step ms list, last Lanczos m)."""
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record(stream)
        m = 0
        for i in range(n):
            m = eng.step(pos, img, F, step0 + i)
            evs[i + 1].record(stream)
        torch.cuda.synchronize()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
        return evs[0].elapsed_time(evs[n]), per, m

    line = {"metric": "BD steps/sec", "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (jittered FCC lattice, N(0,1) forces)"}

    # ------------------------------------------------------------------ reference arm (rank 0 only; same single suspension)
    if args.impl == "reference":
        from oracle import refwrap
        if not refwrap.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpse_ref.so not built (reference sources absent at build time)"}))
            return
        # the list the reference kernels walk is the BUFFERED list (they test the cutoff per pair, PSEv1/Mobility.cu:652), so its
        # buffer is the reference's own default (HOOMD 2.x nlist r_buff = 0.4), not this engine's tuning
        cfg = E.make_config(N, L, xi=args.xi, error=args.error, T=T, dt=dt, seed=1, r_buff=args.ref_r_buff)
        eng = E.Engine(cfg)
        p = eng.params
        pos_np, F_np = util.lattice_positions(N, L, seed=0), util.random_forces(N, seed=100)
        pos = torch.from_numpy(pos_np).cuda(); F = torch.from_numpy(F_np).cuda()
        img = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
        # the reference plugin's own kernels (compiled unmodified for sm_100a); HOOMD's neighbour list is not in the
        # reference tree, so the list comes from the engine's builder, rebuilt every step OUTSIDE the timed spans
        cfg_r = E.make_config(N, L, xi=args.xi, error=args.error, T=T, dt=dt, seed=1, r_buff=args.ref_r_buff, flags=1)
        ref = refwrap.Reference(cfg_r, p, E.ewald_table(cfg_r))
        vel = torch.zeros_like(F); vel[:, 3] = 1.0
        acc = torch.zeros((N, 3), device="cuda")
        h_pos = torch.from_numpy(pos_np.copy()).pin_memory(); h_F = torch.from_numpy(F_np).pin_memory()
        h_img = torch.zeros((N, 3), dtype=torch.int32).pin_memory()

        def ref_steps(n, t0, e2e):
            per = []
            for t in range(t0, t0 + n):
                eng.build_neighbors(pos); ref.set_neighbors(*eng.neighbor_list())
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if e2e:
                    pos.copy_(h_pos, non_blocking=True); F.copy_(h_F, non_blocking=True); img.copy_(h_img, non_blocking=True)
                ref.step(pos, vel, acc, img, F, T, dt, t, sync=False)
                if e2e:
                    h_pos.copy_(pos, non_blocking=True); h_img.copy_(img, non_blocking=True)
                b.record(); torch.cuda.synchronize()
                per.append(a.elapsed_time(b))
            return per
        ref_steps(W, 0, False)
        sampler = ClockSampler(local); sampler.start()
        per = ref_steps(K, W, False)
        h_pos.copy_(pos); h_img.copy_(img)
        per_e2e = ref_steps(max(K // 2, 1), W + K, True)
        clocks = sampler.summary()
        ms = sum(per)
        line.update({"impl": "reference", "value": K / (ms * 1e-3), "ms_per_step": ms / K, "ms_per_step_median": float(np.median(per)),
                     "value_from_median": 1e3 / float(np.median(per)), "clocks": clocks,
                     "config": {"workload": f"PSE BD step, ONE suspension of N={N} spheres, phi={phi}, error={args.error}, xi={args.xi}, grid {p.Nx}^3, P={p.P} "
                                            f"(BASELINE.json configs[2]); the reference is single-GPU (PSEv1/Stokes.cc:104): rank 0 runs it on one GPU at every --gpus"},
                     "e2e": {"value": len(per_e2e) / (sum(per_e2e) * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": 44 * N, "d2h_bytes_per_step": 28 * N,
                             "value_from_median": 1e3 / float(np.median(per_e2e))},
                     "gpu_launches": 0, "lanczos_m": ref.m_lanczos,
                     "cpu_baseline": {"value": K / (ms * 1e-3), "unit": "steps/s", "cores": 0, "kind": "reference",
                                      "sample": "PSE has no CPU path (PSEv1/integrate.py:51-53): this arm runs the reference's own CUDA kernels "
                                                "(PSEv1/*.cu compiled unmodified for sm_100a, oracle/_ref) on one GPU through gpu_stokes_step_one; "
                                                "neighbour-list construction (HOOMD, external to the plugin) is excluded from its timed spans; the mean includes "
                                                "the reference's per-step cudaMalloc/cudaFree of the Krylov basis (PSEv1/Brownian.cu:414), the median is robust to its outliers"}})
        print(json.dumps(line))
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ our arm
    cfg = E.make_config(N, L, xi=args.xi, error=args.error, T=T, dt=dt, seed=1, r_buff=args.r_buff)
    stream = torch.cuda.Stream()          # the engine's launching stream; all events below are recorded on it
    if world > 1:
        from pse_b200 import sharded as S
        eng = S.ShardedEngine(cfg, stream=stream)
    else:
        eng = E.Engine(cfg, stream=stream)
    p = eng.params
    pos_np = util.lattice_positions(N, L, seed=0)     # the same suspension on every rank
    F_np = util.random_forces(N, seed=100)
    pos = torch.from_numpy(pos_np).cuda(); F = torch.from_numpy(F_np).cuda()
    img = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    G = p.Nx * p.Ny * p.Nz
    line["config"] = {"workload": f"PSE BD step, ONE suspension of N={N} spheres, phi={phi}, error={args.error}, xi={args.xi}, kT=1, dt=1e-3, "
                                  f"grid {p.Nx}x{p.Ny}x{p.Nz}, P={p.P}, r_cut={p.rcut:.4f}, r_buff={args.r_buff} (BASELINE.json configs[2])",
                      "parallelism": "single GPU" if world == 1 else
                      f"x-slab decomposition of the one suspension over {world} ranks: own-row neighbour list / SpMV / Lanczos with vector halo rows, own-plane "
                      f"spreading + FFT with two transposes, velocity all-gather; exchanges are peer-memory kernels over NVLink behind device-side flag barriers",
                      "l2_policy": f"working set per step ({(24 * G + 100 * 16 * N) / 1e6:.0f} MB grids+basis) exceeds the 126 MB L2; no explicit flush"}

    step_no = 0
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(W):
            eng.step(pos, img, F, step_no); step_no += 1
    barrier()
    s0 = eng.stats()
    sampler = ClockSampler(local); sampler.start()
    with torch.cuda.stream(stream):
        ms, per, m = timed_steps(eng, stream, pos, img, F, step_no, K)
    step_no += K
    barrier()
    ms = max_over_ranks(ms)
    s1 = eng.stats()
    launches = int(s1["kernel_launches"] - s0["kernel_launches"])

    # ------------------------------------------------------------------ per-phase device times (same workload, CUDA events)
    # long enough (~1 s) for >= 20 clock samples; the sampler keeps running through it
    eng.set_profiling(True)
    KP = int(min(400, max(K, 1000.0 / max(ms / K, 0.1))))
    with torch.cuda.stream(stream):
        for _ in range(KP):
            eng.step(pos, img, F, step_no); step_no += 1
    prof = eng.profile()
    eng.set_profiling(False)
    clocks = sampler.summary()
    st = eng.stats()
    nnz, nnz_stored = st["nnz_active"], st["nnz"]  # the SpMV walks the pruned rows (pairs inside r_cut); per rank when sharded
    phases = {k: {"ms_per_step": v[0] / KP, "us_per_launch": (v[0] / v[1] * 1e3) if v[1] else None, "launches_per_step": v[1] / KP}
              for k, v in prof.items() if v[1]}
    peak, peak_src = measured_peak()
    line.update({"value": K / (ms * 1e-3), "ms_per_step": ms / K, "ms_per_step_median": float(np.median(per)), "clocks": clocks,
                 "gpu_launches": launches, "phases": phases, "lanczos_m": m, "nnz": int(nnz), "nnz_stored": int(nnz_stored),
                 "nlist_builds_in_timed_region": int(s1["nlist_builds"] - s0["nlist_builds"])})

    if world == 1:
        # dominant kernel: the real-space SpMV inside the Lanczos iteration (m per step) — algorithmic bytes B_spmv = 56 N + 4 nnz
        dom = "lanczos_spmv" if "lanczos_spmv" in phases else "spmv"
        b_spmv = 56.0 * N + 4.0 * nnz
        # the first of the m Lanczos products of a step also multiplies the forces (dual right-hand side: +16 N read, +16 N
        # written), which replaces the separate deterministic SpMV; averaged over the m launches the phase timer sees
        if os.environ.get("PSE_SPMV_DUAL", "1") != "0" and dom == "lanczos_spmv":
            b_spmv += 32.0 * N / max(m, 1)
        t_dom = phases[dom]["us_per_launch"] * 1e-6
        traffic = None   # DRAM bytes of one launch of that kernel from the committed `ncu --set full` capture of the same workload
        for cap_name in ("r2_top_kernels.json", "r1_top_kernels.json"):
            try:
                caps = json.load(open(os.path.join(ROOT, "profiles", cap_name)))
                tr = [c["dram__bytes_read.sum"] + c["dram__bytes_write.sum"] for c in caps if c["kernel"].startswith("void spmv_kernel<4, 1, 2, 1, 0>")]
                if tr and N == 1000000:
                    traffic = 1e6 * sum(tr) / len(tr)  # the capture reports Mbyte
                    break
            except Exception:
                pass
        roof = {"bound": "hbm", "kernel": "spmv_kernel<4,LANCZOS,POLY,PRUNED>", "achieved": b_spmv / t_dom / 1e9, "peak": peak, "unit": "GB/s",
                "frac": b_spmv / t_dom / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_spmv, "us_per_launch": t_dom * 1e6, "share_of_step": phases[dom]["ms_per_step"] / (ms / K)}
        b_step = (120.0 * G + 64.0 * N) + (m + 1) * (56.0 * N + 4.0 * nnz) + 64.0 * N * m + 16.0 * N * (m + 1) + 72.0 * N
        roof["step"] = {"algorithmic_bytes": b_step, "achieved": b_step / (ms / K * 1e-3) / 1e9, "frac": b_step / (ms / K * 1e-3) / 1e9 / peak,
                        "formula": "B_step = 120G + 64N + (m+1)(56N + 4nnz) + 64Nm + 16N(m+1) + 72N (SURVEY.md §8d)"}
        roof["kernels"] = kernel_table(phases, N, G, nnz, nnz_stored, p.P, m, peak)
        line["roofline"] = roof

    # ------------------------------------------------------------------ end to end through the host-buffer C ABI entry points
    # pipelined (pse_step_host_async): state device-resident, per step 16 N bytes of forces up and 28 N bytes of positions +
    # images down, both inside the timed region, double-buffered host arrays; and the synchronous form (state up and down, wait)
    h_F = torch.from_numpy(F_np).pin_memory()
    h_pos = [torch.from_numpy(pos.cpu().numpy()).pin_memory() for _ in range(2)]
    h_img = [torch.from_numpy(img.cpu().numpy()).pin_memory() for _ in range(2)]
    hf = h_F.numpy()
    hp, hi = [t.numpy() for t in h_pos], [t.numpy() for t in h_img]
    KE = max(K, 1)
    eng.step_host_async(hp[0], hi[0], hf, step_no, state_in=True); step_no += 1
    eng.wait()
    barrier()
    t0 = time.perf_counter()
    use_prefetch = os.environ.get("BENCH_PREFETCH", "0") != "0"
    if use_prefetch:
        eng.prefetch_forces(hf)
    for i in range(KE):   # slab-decomposed: every rank uploads the forces, rank 0 downloads the result
        if use_prefetch:  # forces of step i+1 go up while step i computes (pse_host_prefetch_forces; measured no better than uploading beside the head)
            eng.prefetch_forces(hf)
        eng.step_host_async(hp[i & 1], hi[i & 1], None if use_prefetch else hf, step_no, state_out=(rank == 0)); step_no += 1
    eng.wait()
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    assert np.isfinite(hp[(KE - 1) & 1]).all()
    KS = max(K // 2, 1)
    eng.step_host_async(hp[0], hi[0], None if use_prefetch else hf, step_no); step_no += 1     # (untimed: every rank's host arrays hold the common state again)
    eng.wait()
    barrier()
    t0 = time.perf_counter()
    for _ in range(KS):
        eng.step_host(hp[0], hi[0], hf, step_no); step_no += 1
    torch.cuda.synchronize()
    t_sync = max_over_ranks(time.perf_counter() - t0)
    line["e2e"] = {"value": KE / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": 16 * N, "d2h_bytes_per_step": 28 * N,
                   "api": "pse_step_host_async + pse_wait (C ABI, pinned host buffers, device-resident state: every step 16 N bytes of forces go up beside the "
                          "position-only head of the step and 28 N bytes of positions + images come down while the next step computes)" + ("; forces uploaded on every rank, result downloaded on rank 0" if world > 1 else ""),
                   "synchronous": {"value": KS / t_sync, "h2d_bytes_per_step": 44 * N, "d2h_bytes_per_step": 28 * N,
                                   "api": "pse_step_host (positions + images + forces in, positions + images out, blocking)"}}

    # steady shear (the plugin's main use case): the tilt moves every step, so the captured step graph is not replayed and the
    # step is issued eagerly; same suspension, shear rate 1, tilt advanced by rate * dt per step as box_resize would
    if world == 1:
        qs, ims = pos.clone(), img.clone()
        xy = 0.0
        with torch.cuda.stream(stream):
            for _ in range(3):
                xy += dt; eng.set_tilt(xy); eng.step(qs, ims, F, step_no, shear_rate=1.0); step_no += 1
            torch.cuda.synchronize()
            a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(K):
                xy += dt; eng.set_tilt(xy); eng.step(qs, ims, F, step_no, shear_rate=1.0); step_no += 1
            b0.record(stream)
            torch.cuda.synchronize()
        line["sheared"] = {"value": K / (a0.elapsed_time(b0) * 1e-3), "unit": "steps/s",
                           "note": "same suspension under steady shear (rate 1, tilt changing every step: eager launches instead of graph replay)"}
        eng.set_tilt(0.0)
        del qs, ims

    # deterministic M.F time (second half of the BASELINE metric): `mf_us` with positions that change from call to call (two
    # configurations alternate, a displacement of 1e-3 radii: everything position-dependent is redone every call), and
    # `mf_us_fixed_positions` for the operator applied again at an unchanged configuration (iterative solvers), where the
    # pruned list, the wave-space binning and the Gaussian factor rows of the previous call are reused
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pos2 = pos.clone(); pos2[:, 0] += 1e-3
    with torch.cuda.stream(stream):
        eng.mobility(pos, F); eng.mobility(pos2, F)
        barrier()
        a.record(stream)
        for i in range(10):
            eng.mobility(pos2 if i & 1 else pos, F)
        b.record(stream)
        torch.cuda.synchronize()
        line["mf_us"] = max_over_ranks(a.elapsed_time(b) / 10 * 1e3)
        eng.mobility(pos, F)
        barrier()
        a.record(stream)
        for _ in range(10):
            eng.mobility(pos, F)
        b.record(stream)
    torch.cuda.synchronize()
    line["mf_us_fixed_positions"] = max_over_ranks(a.elapsed_time(b) / 10 * 1e3)
    del pos2
    if world == 1:
        b_mf = 120.0 * G + 64.0 * N + 56.0 * N + 4.0 * nnz
        line["roofline"]["mf"] = {"us": line["mf_us"], "algorithmic_MB": b_mf / 1e6, "frac": b_mf / (line["mf_us"] * 1e-6) / 1e9 / peak}
    else:
        info = eng.shard_info().as_dict()
        line["shard"] = {k: info[k] for k in ("x0", "x1", "halo_left", "halo_right", "buffer_planes", "halo_layers", "row0", "row1")}
        line["shard"]["collectives_per_step"] = None
        c0, b0 = info["collectives"], info["bytes_sent"]
        with torch.cuda.stream(stream):
            eng.step(pos, img, F, step_no); step_no += 1
        torch.cuda.synchronize()
        info = eng.shard_info().as_dict()
        line["shard"]["collectives_per_step"] = int(info["collectives"] - c0)
        line["shard"]["bytes_sent_per_step_rank0"] = int(info["bytes_sent"] - b0)

    # ------------------------------------------------------------------ side keys
    eng.close()
    del eng
    torch.cuda.empty_cache()
    if world > 1 and not args.no_replicas:
        # N independent replicas (no data-path collective): reported separately, never as scaling
        cfg_r = E.make_config(N, L, xi=args.xi, error=args.error, T=T, dt=dt, seed=1 + rank, r_buff=args.r_buff)
        er = E.Engine(cfg_r, stream=stream)
        pr = torch.from_numpy(util.lattice_positions(N, L, seed=rank)).cuda()
        with torch.cuda.stream(stream):
            for t in range(W):
                er.step(pr, img, F, t)
            barrier()
            ms_r, _, _ = timed_steps(er, stream, pr, img, F, W, K)
        barrier()
        ms_r = max_over_ranks(ms_r)
        line["replicas"] = {"value": world * K / (ms_r * 1e-3), "unit": "steps/s", "note": f"{world} independent suspensions, one per GPU; not scaling"}
        er.close(); del er, pr
        torch.cuda.empty_cache()
    if args.config5 or world >= 8:
        try:
            line["config5"] = config5_block(torch, dist, E, world, rank, stream, max_over_ranks, barrier)
        except Exception as ex:  # reported, never required
            line["config5"] = {"unavailable": str(ex)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args)
        except Exception as ex:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def config5_block(torch, dist, E, world, rank, stream, max_over_ranks, barrier):
    """BASELINE.json configs[4]: N = 8M, phi = 0.4, error 1e-4 (xi = 0.45 -> 432^3 grid, P = 8), one suspension over all ranks."""
    N, phi, xi, error = 8000000, 0.4, 0.45, 1e-4
    L = util.box_length(N, phi)
    cfg = E.make_config(N, L, xi=xi, error=error, T=1.0, dt=1e-3, seed=1, r_buff=0.8)
    if world > 1:
        from pse_b200 import sharded as S
        eng = S.ShardedEngine(cfg, stream=stream)
    else:
        eng = E.Engine(cfg, stream=stream)
    p = eng.params
    pos = torch.from_numpy(util.lattice_positions(N, L, seed=0)).cuda(); F = torch.from_numpy(util.random_forces(N, seed=100)).cuda()
    img = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for t in range(3):
            eng.step(pos, img, F, t)
        barrier()
        a.record(stream)
        for t in range(3, 13):
            m = eng.step(pos, img, F, t)
        b.record(stream)
        torch.cuda.synchronize()
        ms = max_over_ranks(a.elapsed_time(b) / 10)
        eng.mobility(pos, F)
        barrier()
        a.record(stream)
        for _ in range(5):
            eng.mobility(pos, F)
        b.record(stream)
        torch.cuda.synchronize()
    mf = max_over_ranks(a.elapsed_time(b) / 5 * 1e3)
    out = {"workload": f"N={N}, phi={phi}, error={error}, xi={xi}, grid {p.Nx}^3, P={p.P} (BASELINE.json configs[4]), one suspension over {world} rank(s)",
           "ms_per_step": ms, "steps_per_s": 1e3 / ms, "mf_us": mf, "lanczos_m": m}
    eng.close()
    return out


if __name__ == "__main__":
    main()
