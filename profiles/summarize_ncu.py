"""Turn ncu reports into the JSON / CSV summaries kept under profiles/ (run here, on the CPU box, after gpurun).

  python profiles/summarize_ncu.py top  gpurun_out/r1_top.ncu-rep   profiles/r1_top_kernels.json
  python profiles/summarize_ncu.py list gpurun_out/r1_launches.csv  profiles/r1_launch_shares.json
"""
import csv, json, subprocess, sys, collections

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def top(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = {"kernel": d["Kernel Name"][:90]}
        for m in KEEP:
            if m in d and d[m] not in ("", "n/a"):
                k[m] = float(d[m].replace(",", "")); k[m + "_unit"] = units[hdr.index(m)]
        st = {h.split("issue_stalled_")[1].replace("_per_warp_active.pct", ""): float(v) for h, v in d.items()
              if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith("_per_warp_active.pct") and v not in ("", "n/a")}
        k["stall_pct_of_warp_active"] = {a: round(b, 1) for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:5]}
        res.append(k)
    json.dump(res, open(out, "w"), indent=1)
    for k in res:
        print(f"{k['kernel'][:48]:48s} {k.get('gpu__time_duration.sum', 0):9.1f} us  dram {k.get('dram__bytes_read.sum', 0) + k.get('dram__bytes_write.sum', 0):8.1f} MB"
              f"  issue {k.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  l1tex {k.get('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}%")


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    data = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv]
    idx = [i for i, (k, v) in enumerate(data) if k.startswith("integrate_kernel")]
    seg = data[idx[-2] + 1: idx[-1] + 1]  # the last complete step
    agg = collections.OrderedDict()
    for k, v in seg:
        a = agg.setdefault(k[:70], [0, 0.0]); a[0] += 1; a[1] += v / 1e3
    tot = sum(v for _, v in seg) / 1e3
    res = {"step_total_us_serialised": tot, "launches": len(seg),
           "kernels": [{"kernel": k, "launches": c, "total_us": round(v, 1), "share": round(v / tot, 4)} for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(res, open(out, "w"), indent=1)
    print(f"one step: {len(seg)} launches, {tot:.1f} us serialised")
    for k in res["kernels"][:14]:
        print(f"  {k['kernel'][:56]:56s} {k['launches']:3d} {k['total_us']:9.1f} {100 * k['share']:5.1f}%")


if __name__ == "__main__":
    {"top": top, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
