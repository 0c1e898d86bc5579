"""Static resource usage (registers, stack, shared, local = spills) of every kernel in libessentials_b200.so,
aggregated over template instantiations:  python scripts/kernel_resources.py > profiles/rNN_kernel_resources.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "--dump-resource-usage", os.path.join(ROOT, "essentials_b200", "libessentials_b200.so")],
                     capture_output=True, text=True).stdout
rows, cur = [], None
for line in txt.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        rows.append((cur,) + tuple(int(x) for x in m.groups()))
        cur = None
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
agg = {}
for (_, reg, stack, shared, local), name in zip(rows, names):
    base = re.sub(r"<.*", "", name).replace("void ", "").split("(")[0]
    a = agg.setdefault(base, [0, 0, 0, 0, 0])
    a[0] += 1
    a[1], a[2], a[3], a[4] = max(a[1], reg), max(a[2], stack), max(a[3], shared), max(a[4], local)
print(f"{'kernel (max over its instantiations)':72s} {'inst':>4s} {'REG':>4s} {'STACK':>5s} {'SHARED':>6s} {'LOCAL':>5s}")
for k, (n, reg, st, sh, lo) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} {n:4d} {reg:4d} {st:5d} {sh:6d} {lo:5d}")
