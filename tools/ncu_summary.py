"""Summarise gpurun_out ncu artefacts into small text files for profiles/ (run here, no GPU needed).
   python tools/ncu_summary.py <tag>      reads gpurun_out/<tag>_launches.csv and <tag>_prof_tile.ncu-rep"""
import csv, io, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel share of the step -------------------------------------------------
lp = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp, errors="ignore") if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ik])
        v = float(r[iv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    unit = rows[1][hdr.index("Metric Unit")] if "Metric Unit" in hdr else "ns"
    with open(os.path.join(out_dir, f"{tag}_launch_list.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, first {len(rows)-1} launches of\n"
                f"# `python bench.py --steps 2 --warmup 1 --no-cpu` (cold-cache, serialised: compare SHARES)\n"
                f"# unit: {unit}\n# kernel | launches | total | share | min | mean | max\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k} | {a[0]} | {a[1]:.0f} | {100*a[1]/tot:.2f}% | {a[2]:.0f} | {a[1]/a[0]:.0f} | {a[3]:.0f}\n")
    print(open(os.path.join(out_dir, f"{tag}_launch_list.txt")).read())

# ---- full capture: curated raw metrics ------------------------------------------------------------
rp = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_tile.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "inst_executed", "thread_inst_executed",
            "smsp__average_warp_latency_per_inst_issued.ratio"] + \
           [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(os.path.join(out_dir, f"{tag}_tile_kernel_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tile -s 2 -c 2\n"
                f"# command: python bench.py --steps 1 --warmup 1 --no-cpu --iters 120   (1080p, w=3; one launch = 120 sweeps)\n"
                f"# one column per captured launch; ncu flushes caches between replays, so DRAM bytes are cold-cache\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"{w} [{units[i]}] : " + " | ".join(r[i] for r in rows[2:]) + "\n")
    print(open(os.path.join(out_dir, f"{tag}_tile_kernel_ncu.txt")).read())
