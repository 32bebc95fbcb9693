"""Stall samples and executed instructions per CUDA source line of an .ncu-rep captured with --import-source on:
python scripts/ncu_lines.py file.ncu-rep [min share]"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
cur = None; hdr = None; agg = collections.OrderedDict()
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= iex or not r[0].isdigit(): continue     # per-line summary rows only
    a = agg.setdefault((cur, int(r[0])), [0, 0, r[1]])
    try:
        a[0] += int(r[isamp]); a[1] += int(r[iex])
    except ValueError:      # a source line with embedded quotes confuses the csv export
        pass
tot = sum(a[0] for a in agg.values()); tote = sum(a[1] for a in agg.values())
print("samples", tot, "warp instructions", tote)
for (f, ln), a in agg.items():
    if a[0] > tot * thr or a[1] > tote * thr * 1.5:
        print("%-14s %5d samp %.3f inst %.3f | %s" % (f, ln, a[0] / tot, a[1] / tote, a[2].strip()[:105]))
