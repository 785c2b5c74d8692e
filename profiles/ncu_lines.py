"""Per-CUDA-source-line totals of an .ncu-rep (needs -lineinfo and --import-source on): instructions executed and
warp-stall samples per line of the kernel's source, top N by instructions.
    python profiles/ncu_lines.py gpurun_out/<name>.ncu-rep [top_n] [file-substring]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, ix = None, None, None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr) if h not in ix} if False else {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
        continue
    if r[ix["Address"]] != "-":          # SASS rows repeat the line's numbers
        continue
    key = (cur_file.split("/")[-1], int(r[0]))
    inst = int(r[ix["Instructions Executed"]] or 0)
    smp = int(r[ix["# Samples"]] or 0)
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += inst
    a[1] += smp
tot_i = sum(v[0] for v in agg.values())
tot_s = sum(v[1] for v in agg.values())
print(f"total instructions {tot_i}, samples {tot_s}")
for (f, ln), (inst, smp, text) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    if want and want not in f:
        continue
    print(f"{f}:{ln:5d} {inst:12d} {100 * inst / max(tot_i, 1):5.1f}%  samples {smp:7d} {100 * smp / max(tot_s, 1):5.1f}%  {text.strip()[:90]}")
