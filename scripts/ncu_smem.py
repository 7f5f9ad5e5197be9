"""Shared-memory wavefronts per SASS instruction of an .ncu-rep: python scripts/ncu_smem.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hdr]; ix = {h: i for i, h in enumerate(H)}
data = rows[hdr + 1:]
W, E, I = ix["L1 Wavefronts Shared"], ix["L1 Wavefronts Shared Excessive"], ix["L1 Wavefronts Shared Ideal"]
tw = sum(int(r[W] or 0) for r in data); te = sum(int(r[E] or 0) for r in data)
print("total wavefronts", tw, "excessive", te, f"({100*te/tw:.1f}%)")
for n, r in enumerate(data):
    r.append(n)
srt = sorted(data, key=lambda r: -int(r[W] or 0))
for r in srt[:top]:
    print(f'{r[-1]:5d} {int(r[W]):10d} excess {int(r[E] or 0):10d} ideal {int(r[I] or 0):10d} exec {int(r[ix["Instructions Executed"]]):9d}  {r[ix["Source"]].strip()[:60]}')
