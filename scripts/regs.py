"""Register / spill / shared-memory usage of every kernel in an object file: python scripts/regs.py file.o [filter]"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "--dump-resource-usage", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
name = None
for line in out.splitlines():
    line = line.strip()
    if line.startswith("Function"):
        name = subprocess.run(["c++filt", line.split()[1].rstrip(":")], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|hg::|dev::", "", name).split("(")[0]
    elif line.startswith("REG") and name and flt in name:
        print(f"{name[:90]:90s} {line}")
