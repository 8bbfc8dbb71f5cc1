"""Top stall-sample SASS instructions of a kernel from an .ncu-rep (source page), run here without a GPU.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[0][1] if rows[0] else "")
hdr = rows[h]
si, ii, ai = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
tot = sum(float(r[si] or 0) for r in body) or 1.0
toti = sum(float(r[ii] or 0) for r in body)
print("total samples %d, warp instructions executed %d" % (tot, toti))
order = sorted(range(len(body)), key=lambda k: -float(body[k][si] or 0))[:top]
for k in sorted(order):
    r = body[k]
    print("%5.1f%%  #%-4d exec %10s  %s" % (100 * float(r[si]) / tot, k, r[ii], r[ai].strip()[:100]))
