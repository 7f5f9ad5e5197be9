"""Where the warp-stall samples of an .ncu-rep land: python scripts/ncu_hot.py rep [top]  (SASS view with -lineinfo source lines)"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "Address": hdr = i; break
H = rows[hdr]; ix = {h: i for i, h in enumerate(H)}
print("columns:", [h for h in H[:8]])
data = rows[hdr + 1:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data if len(r) > ix["# Samples"])
stall_cols = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = collections.Counter()
for h in stall_cols:
    agg[h] = sum(int(r[ix[h]] or 0) for r in data if len(r) > ix[h])
print("stall totals:", [(k, v) for k, v in agg.most_common(8)])
# by instruction
srt = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))
for r in srt[:top]:
    st = sorted([(int(r[ix[h]] or 0), h[6:]) for h in stall_cols], reverse=True)[:2]
    print(f'{int(r[ix["# Samples"]]):6d} {100*int(r[ix["# Samples"]])/tot:5.1f}%  {r[ix["Source"]].strip()[:70]:70s} {st}')
