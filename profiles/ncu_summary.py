"""Summarise an .ncu-rep (read here, without a GPU): headline metrics + the hottest SASS lines.
    python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_elapsed.avg", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print(f"  {k:64s} {d[k]:>16s} {units[hdr.index(k)]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, hi in enumerate(starts):
    end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0] != "Address"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print(f"\n=== kernel #{si}: {rows[hi - 1][1] if hi else ''}")
    print(f"SASS lines: {len(data)}, warp-stall samples: {tot}")
    agg = {}
    for r in data:
        for h in stalls:
            agg[h] = agg.get(h, 0) + int(r[ix[h]] or 0)
    print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    print(f"top {top_n} SASS lines by samples")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top_n]:
        s = {h[6:]: int(r[ix[h]] or 0) for h in stalls}
        best = ", ".join(f"{k}={v}" for k, v in sorted(s.items(), key=lambda kv: -kv[1])[:2] if v)
        print(f"  {r[ix['Address']][-5:]} {r[ix['# Samples']]:>6s} x{r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:64]:64s} {best}")
