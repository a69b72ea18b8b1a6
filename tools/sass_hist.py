"""Dynamic opcode histogram of a kernel from `ncu --page source --csv --print-source sass` output (gz ok).
usage: python tools/sass_hist.py gpurun_out/st_x_sass.csv.gz  -> executed warp-instructions and stall samples per opcode"""
import csv, gzip, sys, collections
f = gzip.open(sys.argv[1], "rt") if sys.argv[1].endswith(".gz") else open(sys.argv[1])
rows = list(csv.reader(f))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {k: i for i, k in enumerate(hdr)}
ex, smp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.rstrip(";")
    key = op
    n = int(r[ix["Instructions Executed"]] or 0)
    ex[key] += n; tot += n
    smp[key] += int(r[ix["# Samples"]] or 0)
stot = sum(smp.values())
print("total warp-instructions %d, samples %d" % (tot, stot))
for k, v in ex.most_common(45):
    print("%-28s %12d  %5.1f%%   samples %5.1f%%" % (k, v, 100.0 * v / tot, 100.0 * smp[k] / max(stot, 1)))
