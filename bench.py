#!/usr/bin/env python3
"""bench.py — throughput of one full PSE Brownian-dynamics step (BASELINE.json metric) on N GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (real-space RPY + wave-space spread/FFT/scale/interpolate,
deterministic and stochastic parts, Lanczos, Euler update) over one synthetic suspension.
Workload at every N: BASELINE.json configs[2] — 1,000,000 spheres, phi = 0.3, error 1e-3, xi = 0.5,
kT = 1, dt = 1e-3 — per GPU (weak scaling: independent replicas).

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline, cpu_baseline, e2e,
gpu_launches, clocks, phases.
"""
import argparse
import ctypes
import json
import math
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
    ap.add_argument("--r-buff", type=float, default=0.8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-N", type=int, default=100000)
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
                time.sleep(0.04)  # the timed region of the default run is ~80 ms; denser polling costs ~2 % of throughput
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
    spread/FFT/scale/interpolate, Lanczos) timed on this box's host cores on a bounded sample: one full
    velocity evaluation + Euler update at N_s particles, same phi / xi / error; steps/s scaled by N_s/N
    (the algorithm is O(N) at fixed density)."""
    from oracle import oraclewrap as O
    Ns = min(args.cpu_sample_N, args.N)
    L = util.box_length(Ns, args.phi)
    o = O.Oracle(Ns, L, xi=args.xi, error=args.error, ref_pi=False)
    pos, F = util.lattice_positions(Ns, L, 0), util.random_forces(Ns, 1)
    rng = np.random.default_rng(2)
    G = o.prm.Nx * o.prm.Ny * o.prm.Nz
    up, ug = rng.random((Ns, 3), dtype=np.float32), rng.random((G, 6), dtype=np.float32)
    t0 = time.perf_counter()
    o.neighbors(pos, o.prm.rcut + args.r_buff, brute=False)
    t_nl = time.perf_counter() - t0
    t0 = time.perf_counter()
    U, m = o.velocity(pos, F, 1.0, 1e-3, u_particles=up, u_grid=ug, m_in=2)
    img = np.zeros((Ns, 3), dtype=np.int32)
    o.integrate(pos, img, U, 1e-3)
    t_step = time.perf_counter() - t0
    cores = O.lib().orc_num_threads()
    return {"value": (1.0 / t_step) * Ns / args.N, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"1 full step (M.F + Brownian, Lanczos m={m}) at N={Ns}, phi={args.phi}, grid {o.prm.Nx}^3: {t_step:.2f} s "
                      f"(+{t_nl:.2f} s neighbour list, not counted); scaled by {Ns}/{args.N}"}


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
    cfg = E.make_config(N, L, xi=args.xi, error=args.error, T=T, dt=dt, seed=1 + rank, r_buff=args.r_buff)
    stream = torch.cuda.Stream()          # the engine's launching stream; all events below are recorded on it
    eng = E.Engine(cfg, stream=stream)
    p = eng.params
    pos_np = util.lattice_positions(N, L, seed=rank)
    F_np = util.random_forces(N, seed=100 + rank)
    pos = torch.from_numpy(pos_np).cuda(); F = torch.from_numpy(F_np).cuda()
    img = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    G = p.Nx * p.Ny * p.Nz
    K, W = args.steps, max(args.warmup, 0)
    workload = {"workload": f"PSE BD step, N={N} spheres, phi={phi}, error={args.error}, xi={args.xi}, kT=1, dt=1e-3, grid {p.Nx}x{p.Ny}x{p.Nz}, "
                            f"P={p.P}, r_cut={p.rcut:.4f}, r_buff={args.r_buff} (BASELINE.json configs[2])",
                "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (no data-path collective)",
                "l2_policy": f"working set per step ({(24 * G + 100 * 16 * N) / 1e6:.0f} MB grids+basis) exceeds the 126 MB L2; no explicit flush"}

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

    line = {"metric": "BD steps/sec", "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (jittered FCC lattice, N(0,1) forces)",
            "config": workload}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        from oracle import refwrap
        if not refwrap.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpse_ref.so not built (reference sources absent at build time)"}))
            return
        # the reference plugin's own kernels (compiled unmodified for sm_100a); HOOMD's neighbour list is not in the
        # reference tree, so the list comes from the engine's builder, rebuilt every step OUTSIDE the timed spans
        cfg_r = E.make_config(N, L, xi=args.xi, error=args.error, T=T, dt=dt, seed=1, r_buff=args.r_buff, flags=1)
        ref = refwrap.Reference(cfg_r, p, E.ewald_table(cfg_r))
        vel = torch.zeros_like(F); vel[:, 3] = 1.0
        acc = torch.zeros((N, 3), device="cuda")
        h_pos = torch.from_numpy(pos_np.copy()).pin_memory(); h_F = torch.from_numpy(F_np).pin_memory()
        h_img = torch.zeros((N, 3), dtype=torch.int32).pin_memory()

        def ref_steps(n, t0, e2e):
            total = 0.0
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
                total += a.elapsed_time(b)
                if os.environ.get("PSE_BENCH_VERBOSE"):
                    print(f"[ref step {t}] {a.elapsed_time(b):.2f} ms  m={ref.m_lanczos}  free={torch.cuda.mem_get_info()[0] >> 20} MiB", file=sys.stderr)
            return total
        ref_steps(W, 0, False)
        sampler = ClockSampler(local); sampler.start()
        ms = ref_steps(K, W, False)
        clocks = sampler.summary()
        h_pos.copy_(pos); h_img.copy_(img)
        ms_e2e = ref_steps(max(K // 2, 1), W + K, True)
        line.update({"impl": "reference", "value": K / (ms * 1e-3), "ms_per_step": ms / K, "clocks": clocks,
                     "e2e": {"value": max(K // 2, 1) / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": 44 * N, "d2h_bytes_per_step": 28 * N},
                     "gpu_launches": 0, "lanczos_m": ref.m_lanczos,
                     "cpu_baseline": {"value": K / (ms * 1e-3), "unit": "steps/s", "cores": 0, "kind": "reference",
                                      "sample": "PSE has no CPU path (PSEv1/integrate.py:51-53): this arm runs the reference's own CUDA kernels "
                                                "(PSEv1/*.cu compiled unmodified for sm_100a, oracle/_ref) on the same GPU through gpu_stokes_step_one; "
                                                "neighbour-list construction (HOOMD, external to the plugin) is excluded from its timed spans"}})
        print(json.dumps(line))
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ our arm: device-resident steps
    step_no = 0
    for _ in range(W):
        eng.step(pos, img, F, step_no); step_no += 1
    barrier()
    s0 = eng.stats()
    sampler = ClockSampler(local); sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(K):
        m = eng.step(pos, img, F, step_no); step_no += 1
    b.record(stream)
    barrier()
    clocks = sampler.summary()
    ms = max_over_ranks(a.elapsed_time(b))
    s1 = eng.stats()
    launches = int(s1["kernel_launches"] - s0["kernel_launches"])

    # ------------------------------------------------------------------ per-phase device times (same workload, CUDA events)
    eng.set_profiling(True)
    KP = max(min(K, 10), 1)
    for _ in range(KP):
        eng.step(pos, img, F, step_no); step_no += 1
    prof = eng.profile()
    eng.set_profiling(False)
    st = eng.stats()
    nnz, nnz_stored = st["nnz_active"], st["nnz"]  # the SpMV walks the pruned rows (pairs inside r_cut)
    phases = {k: {"ms_per_step": v[0] / KP, "us_per_launch": (v[0] / v[1] * 1e3) if v[1] else None, "launches_per_step": v[1] / KP}
              for k, v in prof.items() if v[1]}
    peak, peak_src = measured_peak()
    # dominant kernel: the real-space SpMV inside the Lanczos iteration (m per step) — algorithmic bytes B_spmv = 56 N + 4 nnz
    dom = "lanczos_spmv" if "lanczos_spmv" in phases else "spmv"
    b_spmv = 56.0 * N + 4.0 * nnz
    # the first of the m Lanczos products of a step also multiplies the forces (dual right-hand side: +16 N read, +16 N
    # written), which replaces the separate deterministic SpMV; averaged over the m launches the phase timer sees
    dual = os.environ.get("PSE_SPMV_DUAL", "1") != "0" and dom == "lanczos_spmv"
    if dual:
        b_spmv += 32.0 * N / max(m, 1)
    t_dom = phases[dom]["us_per_launch"] * 1e-6
    # DRAM bytes of one launch of that kernel from the committed `ncu --set full` capture of the same workload (profiles/)
    traffic = None
    try:
        caps = json.load(open(os.path.join(ROOT, "profiles", "r1_top_kernels.json")))
        tr = [c["dram__bytes_read.sum"] + c["dram__bytes_write.sum"] for c in caps if c["kernel"].startswith("void spmv_kernel<4, 1, 2, 1, 0>")]
        if tr and abs(N - 1000000) < 1:
            traffic = 1e6 * sum(tr) / len(tr)  # the capture reports Mbyte
    except Exception:
        traffic = None
    roof = {"bound": "hbm", "kernel": "spmv_kernel<4,LANCZOS,POLY,PRUNED>", "achieved": b_spmv / t_dom / 1e9, "peak": peak, "unit": "GB/s",
            "frac": b_spmv / t_dom / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": b_spmv, "us_per_launch": t_dom * 1e6, "share_of_step": phases[dom]["ms_per_step"] / (ms / K)}
    b_step = (120.0 * G + 64.0 * N) + (m + 1) * (56.0 * N + 4.0 * nnz) + 64.0 * N * m + 16.0 * N * (m + 1) + 72.0 * N
    roof["step"] = {"algorithmic_bytes": b_step, "achieved": b_step / (ms / K * 1e-3) / 1e9, "frac": b_step / (ms / K * 1e-3) / 1e9 / peak,
                    "formula": "B_step = 120G + 64N + (m+1)(56N + 4nnz) + 64Nm + 16N(m+1) + 72N (SURVEY.md §8d)"}

    # ------------------------------------------------------------------ end to end through the host-buffer C ABI entry point
    import torch as _t
    h_pos = _t.from_numpy(pos.cpu().numpy()).pin_memory(); h_F = _t.from_numpy(F_np).pin_memory()
    h_img = _t.from_numpy(img.cpu().numpy()).pin_memory()
    hp, hf, hi = h_pos.numpy(), h_F.numpy(), h_img.numpy()
    KE = max(K // 2, 1)
    eng.step_host(hp, hi, hf, step_no); step_no += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(KE):
        eng.step_host(hp, hi, hf, step_no); step_no += 1
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)

    line.update({"value": world * K / (ms * 1e-3), "ms_per_step": ms / K, "clocks": clocks, "gpu_launches": launches,
                 "e2e": {"value": world * KE / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": 44 * N, "d2h_bytes_per_step": 28 * N,
                         "api": "pse_step_host (C ABI, pinned host buffers: pos+image+force in, pos+image out)"},
                 "roofline": roof, "phases": phases, "lanczos_m": m, "nnz": int(nnz), "nnz_stored": int(nnz_stored), "nlist_builds_in_timed_region": int(s1["nlist_builds"] - s0["nlist_builds"]),
                 "mf_us": None})
    # deterministic M.F time (second half of the BASELINE metric)
    barrier()
    eng.mobility(pos, F)
    a.record(stream)
    for _ in range(5):
        eng.mobility(pos, F)
    b.record(stream); torch.cuda.synchronize()
    line["mf_us"] = a.elapsed_time(b) / 5 * 1e3
    # multi-GPU: the deterministic M.F additionally runs slab-decomposed over all ranks (one suspension, strong scaling)
    if world > 1:
        try:
            from pse_b200 import sharded as S
            cfg_s = E.make_config(N, L, xi=args.xi, error=args.error, T=T, dt=dt, seed=1, r_buff=args.r_buff)
            pos_s = torch.from_numpy(util.lattice_positions(N, L, seed=0)).cuda(); F_s = torch.from_numpy(util.random_forces(N, seed=100)).cuda()
            del eng
            torch.cuda.empty_cache()
            sm = S.ShardedMobility(cfg_s)
            sm.mobility(pos_s, F_s); sm.mobility(pos_s, F_s)
            barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                sm.mobility(pos_s, F_s)
            torch.cuda.synchronize()
            line["mf_us_sharded"] = max_over_ranks(time.perf_counter() - t0) / 5 * 1e6
            line["config"]["mf_sharded"] = f"one suspension over {world} ranks: x-slab spreading/FFT with 2 all-to-all transposes, halo exchange, row-sharded SpMV, all-reduce"
        except Exception as ex:
            line["mf_us_sharded"] = None
            line["config"]["mf_sharded"] = f"unavailable: {ex}"
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args)
        except Exception as ex:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
