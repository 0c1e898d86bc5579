"""Static resource usage of every kernel in the built objects (cuobjdump -res-usage): registers, shared memory,
stack/local bytes (spills), per kernel template, max over its instantiations. No GPU needed.
    python scripts/resource_usage.py > profiles/r02_resource_usage.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = collections.defaultdict(lambda: {"n": 0, "reg": 0, "smem": 0, "stack": 0, "local": 0})
for obj in sorted(glob.glob(os.path.join(ROOT, "essentials_b200", "csrc", "build", "*.o"))):
    text = subprocess.run("cuobjdump -res-usage %s | c++filt" % obj, shell=True, capture_output=True, text=True).stdout
    name = None
    for line in text.splitlines():
        m = re.match(r"\s*Function (?:void )?([\w:]+?)(?:<|\()", line)
        if m:
            name = m.group(1)
            for prefix in ("gunrock::operators::advance::kernels::", "gunrock::operators::", "gunrock::"):
                if name.startswith(prefix):
                    name = name[len(prefix):]
                    break
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            r = rows[name]
            r["n"] += 1
            r["reg"] = max(r["reg"], int(m.group(1)))
            r["stack"] = max(r["stack"], int(m.group(2)))
            r["smem"] = max(r["smem"], int(m.group(3)))
            r["local"] = max(r["local"], int(m.group(4)))
            name = None
print("# cuobjdump -res-usage over essentials_b200/csrc/build/*.o (sm_100a); max over each kernel's instantiations")
print("%-58s %6s %5s %8s %6s %6s" % ("kernel", "insts", "regs", "smem B", "stack", "local"))
for name, r in sorted(rows.items(), key=lambda kv: -kv[1]["reg"]):
    print("%-58s %6d %5d %8d %6d %6d" % (name[:58], r["n"], r["reg"], r["smem"], r["stack"], r["local"]))
