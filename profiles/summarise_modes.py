"""Aggregates an `ncu --csv` log of profiles/compare_modes.py into the three counters the north star names,
per schedule (megakernel = render_kernel*, wavefront = wf_*): warp execution efficiency (active threads per
executed warp instruction / 32, instruction-weighted), FP32 (FMA) pipe utilisation (time-weighted), and L2 / HBM
bytes per ray.  usage: summarise_modes.py <ncu.csv> <rays per frame>"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
rays = float(sys.argv[2])
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per_launch = defaultdict(dict)
for r in rows[1:]:
    try:
        per_launch[(r[iid], r[ik])][r[im]] = float(r[iv].replace(",", ""))
    except ValueError:
        pass
agg = defaultdict(lambda: defaultdict(float))
for (_, name), m in per_launch.items():
    mode = "megakernel" if "render_kernel" in name else "wavefront" if "wf_" in name else None
    if mode is None:
        continue
    a = agg[mode]
    t = m.get("gpu__time_duration.sum", 0.0)
    inst = m.get("smsp__inst_executed.sum", 0.0)
    a["launches"] += 1
    a["time_ns"] += t
    a["inst"] += inst
    a["thread_inst"] += inst * m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
    a["fma_pct_x_time"] += t * m.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a["l2_bytes"] += m.get("lts__t_bytes.sum", 0.0)
    a["dram_bytes"] += m.get("dram__bytes.sum", 0.0)
print(f"{'schedule':12s} {'launches':>8s} {'GPU time ms':>12s} {'warp exec eff':>14s} {'FMA pipe %':>11s} {'warp-instr/ray':>15s} {'L2 B/ray':>10s} {'HBM B/ray':>10s}")
for mode, a in agg.items():
    print(f"{mode:12s} {int(a['launches']):8d} {a['time_ns'] / 1e6:12.2f} {a['thread_inst'] / max(1, a['inst']) / 32:14.3f} "
          f"{a['fma_pct_x_time'] / max(1, a['time_ns']):11.2f} {a['inst'] / rays:15.1f} {a['l2_bytes'] / rays:10.1f} {a['dram_bytes'] / rays:10.2f}")
