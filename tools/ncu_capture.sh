#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel-regex> <launch-count> <command...>
# ncu --set full capture of the kernels matching <kernel-regex>; the .ncu-rep stays in /tmp on the GPU box (tens of MB
# each: gpurun copies back at most 64 MiB), only the text summaries land in gpurun_out/:
#   <name>_summary.txt   the per-kernel metric table of tools/ncu_summary.py (duration, DRAM bytes, tensor-pipe activity, stalls)
#   <name>_source.csv    the hottest source lines (--page source), first kernel instance
name=$1; regex=$2; count=$3; shift 3
rep=/tmp/${name}.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:${regex}" -c "${count}" -f -o /tmp/${name} "$@" > gpurun_out/${name}_ncu.log 2>&1
tail -2 gpurun_out/${name}_ncu.log
python tools/ncu_summary.py kernels ${rep} gpurun_out/${name}_summary.txt
ncu -i ${rep} --page source --csv --print-source sass 2>/dev/null | head -4000 > /tmp/${name}_source_all.csv
python - <<PY
import csv, sys
rows = list(csv.reader(open("/tmp/${name}_source_all.csv", errors="replace")))
hdr = None
out = []
for r in rows:
    if hdr is None:
        if any("Sampling" in c or "Samples" in c for c in r):
            hdr = r
        continue
    out.append(r)
if hdr:
    idx = [i for i, c in enumerate(hdr) if "Samples" in c or "Sampling" in c]
    key = idx[0] if idx else 0
    def val(r):
        try: return float(r[key].replace(",", ""))
        except Exception: return 0.0
    out.sort(key=val, reverse=True)
    with open("gpurun_out/${name}_source.csv", "w") as f:
        w = csv.writer(f); w.writerow(hdr)
        for r in out[:60]: w.writerow(r)
PY
# the same samples per CUDA source line (needs -lineinfo, which build.py passes)
ncu -i ${rep} --page source --csv --print-source cuda 2>/dev/null | head -3000 > /tmp/${name}_cuda_all.csv
python - <<PY
import csv
rows = list(csv.reader(open("/tmp/${name}_cuda_all.csv", errors="replace")))
hdr, out = None, []
for r in rows:
    if hdr is None:
        if any("Samples" in c for c in r):
            hdr = r
        continue
    out.append(r)
if hdr:
    key = [i for i, c in enumerate(hdr) if "Samples" in c][0]
    def val(r):
        try: return float(r[key].replace(",", ""))
        except Exception: return 0.0
    tot = sum(val(r) for r in out) or 1.0
    keep = [i for i, c in enumerate(hdr) if c in ("Source", "# Samples", "Instructions Executed", "Warp Stall Sampling (All Samples)") or c.startswith("stall_")]
    out.sort(key=val, reverse=True)
    with open("gpurun_out/${name}_cuda_lines.csv", "w") as f:
        w = csv.writer(f); w.writerow(["share"] + [hdr[i] for i in keep])
        for r in out[:70]: w.writerow(["%.3f" % (val(r) / tot)] + [r[i] if i < len(r) else "" for i in keep])
PY
ls -la gpurun_out/${name}_* | cut -c30-
