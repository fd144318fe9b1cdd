"""Writes profiles/r1_summary.md from the committed bench lines and ncu summaries (run after summarize_ncu.py)."""
import json, os
HERE = os.path.dirname(os.path.abspath(__file__))
L = lambda f: json.load(open(os.path.join(HERE, f)))
d, r, ls, top = L("r1_bench_ours.json"), L("r1_bench_ref.json"), L("r1_launch_shares.json"), L("r1_top_kernels.json")
out = []
out.append("# Round 1 — measured on one B200 (gpurun), N = 1,000,000, phi = 0.3, error 1e-3, xi = 0.5, kT = 1, dt = 1e-3\n")
out.append("Files: `r1_bench_ours.json` / `r1_bench_ref.json` (the two `bench.py` arms, NOT under a profiler), `r1_launches.csv` (ncu\n"
           "`--metrics gpu__time_duration.sum --clock-control none` of `bench.py --steps 2 --warmup 3`) and its per-kernel shares\n"
           "`r1_launch_shares.json`, `r1_top_kernels.json` (ncu `--set full` of the top kernels, key metrics + stall shares), both written by\n"
           "`summarize_ncu.py`; `r1_sanitizer.txt` (compute-sanitizer memcheck / racecheck of a step); this file by `make_summary.py`.\n"
           "ncu times are cold-cache and serialised: compare SHARES with `phases` of the bench line, not absolutes.\n")
out.append("## bench.py (CUDA events, no profiler)\n")
out.append("| arm | steps/s | ms/step | e2e steps/s | Lanczos m |\n|---|---|---|---|---|")
out.append(f"| ours (engine RNG, r_buff 0.8, list rebuilt {d['nlist_builds_in_timed_region']}x in {d['steps']} steps) | {d['value']:.1f} | {d['ms_per_step']:.3f} | {d['e2e']['value']:.1f} | {d['lanczos_m']} |")
out.append(f"| reference kernels (PSEv1/*.cu unmodified, sm_100a; neighbour list excluded; steady state 25 ms/step, the mean carries the outliers of its per-step cudaMalloc/cudaFree, `Brownian.cu:414-760`) | {r['value']:.1f} | {r['ms_per_step']:.3f} | {r['e2e']['value']:.1f} | {r['lanczos_m']} |\n")
rf = d["roofline"]
out.append(f"Deterministic M.F: {d['mf_us']:.0f} us.  CPU port ({d['cpu_baseline']['cores']} threads): {d['cpu_baseline']['value']:.3f} steps/s ({d['cpu_baseline']['sample']}).\n")
tr = f"{rf['traffic'] / 1e6:.0f} MB" if rf.get("traffic") else "n/a"
out.append(f"Roofline (HBM, peak {rf['peak']} GB/s {rf['peak_source']}): dominant kernel {rf['kernel']}: {rf['us_per_launch']:.1f} us/launch, algorithmic "
           f"{rf['algorithmic_bytes_per_launch'] / 1e6:.1f} MB -> {rf['achieved']:.0f} GB/s = {100 * rf['frac']:.1f}% of peak, {100 * rf['share_of_step']:.0f}% of the step "
           f"(ncu DRAM traffic of the same launch: {tr}); whole step {rf['step']['algorithmic_bytes'] / 1e9:.2f} GB algorithmic -> {rf['step']['achieved']:.0f} GB/s = {100 * rf['step']['frac']:.1f}%.\n")
out.append("## Per-phase device time (bench.py `phases`, CUDA events on the engine stream, profiling mode = branches serial)\n")
out.append("| phase | ms/step | launches/step | us/launch | share |\n|---|---|---|---|---|")
for k, v in sorted(d["phases"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    out.append(f"| {k} | {v['ms_per_step']:.3f} | {v['launches_per_step']:.1f} | {v['us_per_launch']:.1f} | {100 * v['ms_per_step'] / d['ms_per_step']:.1f}% |")
out.append("\n(`scale` = x forward FFT + k-space scaling + x inverse FFT in one kernel; `fft_r2c` = z + y forward passes; `fft_c2r` = y + z inverse passes; "
           "`wave_bin` includes the per-call Gaussian factor rows.)\n")
out.append(f"## ncu launch list, one step without a list rebuild ({ls['launches']} launches, {ls['step_total_us_serialised']:.0f} us serialised)\n")
out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
for k in ls["kernels"][:22]:
    out.append(f"| `{k['kernel'][:70]}` | {k['launches']} | {k['total_us']:.1f} | {100 * k['share']:.1f}% |")
out.append("\n## ncu --set full, top kernels (`r1_top_kernels.json`; N = 1M step, one launch each)\n")
out.append("| kernel | us | DRAM r+w MB | DRAM % | issue active % | L1TEX % | L1 hit % | regs | dyn smem KB | smem wavefronts (conflicts) M | top stalls (% of warp-active) |\n|---|---|---|---|---|---|---|---|---|---|---|")
seen = set()
for k in top:
    name = k["kernel"].split("(")[0]
    if name in seen:
        continue
    seen.add(name)
    g = lambda m: k.get(m, 0)
    out.append(f"| `{name[:40]}` | {g('gpu__time_duration.sum'):.0f} | {g('dram__bytes_read.sum') + g('dram__bytes_write.sum'):.0f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
               f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g('l1tex__t_sector_hit_rate.pct'):.0f} | "
               f"{g('launch__registers_per_thread'):.0f} | {g('launch__shared_mem_per_block_dynamic'):.0f} | {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6:.0f} "
               f"({g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:.0f}) | {', '.join(f'{a} {b}' for a, b in k['stall_pct_of_warp_active'].items())} |")
out.append(open(os.path.join(HERE, "r1_notes.md")).read())
open(os.path.join(HERE, "r1_summary.md"), "w").write("\n".join(out))
print("wrote r1_summary.md")
