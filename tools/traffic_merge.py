"""Merge tools/traffic_probe.py's plain timings with the ncu CSV of the same command into
profiles/<tag>_traffic.json (+ a readable table).  No GPU needed.
    python tools/traffic_merge.py <plain.jsonl> <ncu.csv> <tag>"""
import csv, io, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
plain, ncu_csv, tag = sys.argv[1:4]
cases = [json.loads(l) for l in open(plain) if l.startswith("{")]
lines = [l for l in open(ncu_csv, errors="ignore") if l.startswith('"')]
rows = list(csv.reader(io.StringIO("".join(lines))))
hdr = rows[0]
iid, ik, im, iv, iu = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
launches = {}
for r in rows[1:]:
    if "k_jacobi_tile" not in r[ik]:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu].lower()
    if u.startswith("kbyte"): v *= 1e3
    elif u.startswith("mbyte"): v *= 1e6
    elif u.startswith("gbyte"): v *= 1e9
    elif u in ("us", "usecond"): v *= 1e3
    elif u in ("ms", "msecond"): v *= 1e6
    launches.setdefault(int(r[iid]), {})[r[im]] = v
order = [launches[i] for i in sorted(launches)]
PEAK = 6455.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out, pos = {}, 0
table = ["workload | window | k | sweeps/launch | DRAM read MB | DRAM write MB | B per pixel-iteration | L2 hit % | "
         "plain time ms | Gpix-it/s | DRAM GB/s (bytes / plain time) | frac of measured HBM peak | algorithmic 32 B / DRAM B"]
for c in cases:
    n = c["tile_launches"]
    grp = order[pos:pos + n]; pos += n
    if len(grp) < n:
        break
    rd = sum(g["dram__bytes_read.sum"] for g in grp); wr = sum(g["dram__bytes_write.sum"] for g in grp)
    hit = sum(g["lts__t_sector_hit_rate.pct"] for g in grp) / n
    pixit = c["H"] * c["W"] * c["batch"] * c["sweeps"]
    bpp = (rd + wr) / pixit
    gbs = (rd + wr) / (c["iterate_ms"] * 1e-3) / 1e9
    key = f'{c["workload"]}_w{c["window"]}_k{c["k"]}'
    out[key] = {"dram_bytes_per_pixel_iteration": bpp, "dram_read_bytes": rd, "dram_write_bytes": wr, "l2_hit_pct": hit,
                "sweeps": c["sweeps"], "launches": n, "plain_iterate_ms": c["iterate_ms"], "gpixit_s": c["gpixit_s"],
                "dram_gbs": gbs, "dram_frac_of_measured_peak": gbs / PEAK, "algorithmic_over_dram": 32.0 / bpp}
    table.append(f'{c["workload"]} | {c["window"]} | {c["k"]} | {c["sweeps"]} | {rd/1e6:.1f} | {wr/1e6:.1f} | {bpp:.3f} | {hit:.1f} | '
                 f'{c["iterate_ms"]:.3f} | {c["gpixit_s"]:.0f} | {gbs:.0f} | {gbs/PEAK:.3f} | {32.0/bpp:.1f}')
out["_comment"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none "
                   "on `python tools/traffic_probe.py` (one fused launch per row; ncu flushes caches before each "
                   "replay, so L2-resident rows show the cold first touch); times are from the same command WITHOUT ncu")
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", f"{tag}_traffic_table.txt"), "w").write("\n".join(table) + "\n")
print("\n".join(table))
