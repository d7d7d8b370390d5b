"""Dev: instruction and sample share of every source line of an ncu report, in line order."""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == 'Line No'][0]
hdr = rows[hi]
ie = hdr.index('Instructions Executed'); sa = hdr.index('# Samples')
def f(x):
    try: return float(x)
    except: return 0.0
data = [r for r in rows[hi+1:] if len(r) > ie and r[0] not in ('', 'Line No')]
tot = sum(f(r[ie]) for r in data); ts = sum(f(r[sa]) for r in data)
print("total inst %.4g samples %d" % (tot, ts))
for r in data:
    if 100*f(r[ie])/tot >= thr or 100*f(r[sa])/ts >= thr:
        print("%5s inst %5.2f%% samp %5.2f%% | %s" % (r[0], 100*f(r[ie])/tot, 100*f(r[sa])/ts, r[1].rstrip()[:110]))
