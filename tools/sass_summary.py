"""SASS instruction counts per kernel of the built library (cuobjdump -sass; no GPU needed):
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fem_elastoplasticity_b200", "libfem_b200.so")
COLS = ["UBLKCP", "UTMALDG", "SYNCS", "DFMA", "DMUL", "DADD", "F2F", "LDS", "LDG", "LDG.256", "STG", "STG.256"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
counts, order, cur, k = {}, [], None, 0
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[k]
        k += 1
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_\.]+)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        counts[cur]["total"] += 1
        counts[cur][base] += 1
        if base in ("LDG", "STG") and ".256" in op:
            counts[cur][base + ".256"] += 1
print("SASS instruction counts of the shipped kernels (cuobjdump -sass fem_elastoplasticity_b200/libfem_b200.so, sm_100a), round 2; tools/sass_summary.py.")
print("UBLKCP = cp.async.bulk (bulk async copy global -> shared), UTMALDG = cp.async.bulk.tensor (TMA tensor copy), SYNCS = mbarrier ops,")
print("LDG.256 / STG.256 = 256-bit global loads / stores (ld/st.global.v4.f64; counted inside LDG / STG too).\n")
print(f"{'kernel':100s} {'total':>6s} " + " ".join(f"{c:>7s}" for c in COLS))
for name in sorted(order):
    c = counts[name]
    print(f"{name[:100]:100s} {c['total']:6d} " + " ".join(f"{c[x]:7d}" for x in COLS))
