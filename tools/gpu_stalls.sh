#!/bin/bash
# ncu --set full of one kernel family; returns every warp-stall metric and the per-SASS-line sampling table.
# usage: tools/gpu_stalls.sh NAME KERNEL_REGEX driver-args...
mkdir -p gpurun_out
name=$1; rx=$2; shift 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s ${SKIP:-2} -c ${COUNT:-1} -f -o gpurun_out/st_$name python tools/prof_driver.py "$@" > gpurun_out/st_$name.log 2>&1
echo "$name rc=$?"
ncu -i gpurun_out/st_$name.ncu-rep --page raw --csv > gpurun_out/st_${name}_raw.csv 2>/dev/null
python - "$name" <<'P'
import csv, sys
name = sys.argv[1]
rows = list(csv.reader(open("gpurun_out/st_%s_raw.csv" % name)))
hdr, units = rows[0], rows[1]
with open("gpurun_out/st_%s_stalls.txt" % name, "w") as f:
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write("== %s\n" % d.get("Kernel Name", "")[:160])
        for k in hdr:
            if "stall" in k or "issue" in k or "pipe" in k and "pct" in k or k.startswith("launch__") or "duration" in k or "bank" in k or "wavefront" in k:
                f.write("  %-100s %s %s\n" % (k, d[k], units[hdr.index(k)]))
P
ncu -i gpurun_out/st_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/st_${name}_sass.csv 2>/dev/null
gzip -f gpurun_out/st_${name}_sass.csv
rm -f gpurun_out/st_${name}_raw.csv
ls -la gpurun_out | head -30
