"""Executed-instruction mix by opcode of an .ncu-rep (SASS page): python scripts/ncu_opmix.py rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; ix = {h: i for i, h in enumerate(H)}
cnt = collections.Counter(); tot = 0
for r in rows[hdr + 1:]:
    if len(r) <= ix["Instructions Executed"]: continue
    n = int(r[ix["Instructions Executed"]] or 0)
    toks = r[ix["Source"]].split()
    op = next((t for t in toks if not t.startswith("@")), "?").split(".")[0]
    if op in ("LDS", "STS", "LDG", "STG"): op = next(t for t in toks if not t.startswith("@")).split(".")[0] + "." + (".".join(next(t for t in toks if not t.startswith("@")).split(".")[1:]) or "32")
    cnt[op] += n; tot += n
print("total warp instructions", tot)
for op, n in cnt.most_common(top):
    print(f"{op:24s} {n:12d} {100*n/tot:5.1f}%")
