#!/usr/bin/env python
"""One table out of the `-Xptxas -v` logs of csrc/build/: kernel, registers, spill bytes, static shared memory, stack frame.
Usage: tools/ptxas_summary.py hip-bvh-construction_b200/csrc/build/*.ptxas.log  (or `make -C csrc ptxas-summary`)."""
import re
import subprocess
import sys


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


rows = []
for path in sys.argv[1:]:
    cur = None
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            cur = {"file": path.split("/")[-1].replace(".ptxas.log", ""), "name": m.group(1), "regs": 0, "spill": 0, "smem": 0, "stack": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["spill"] = int(m.group(1)), int(m.group(2)) + int(m.group(3))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = int(s.group(1)) if s else 0
names = demangle([r["name"] for r in rows])
print(f"{'file':16s} {'regs':>4s} {'spill':>5s} {'stack':>5s} {'smem':>6s}  kernel")
for r in rows:
    nm = re.sub(r"\(.*", "", names.get(r["name"], r["name"]))
    nm = nm.replace("void ", "")
    print(f"{r['file']:16s} {r['regs']:4d} {r['spill']:5d} {r['stack']:5d} {r['smem']:6d}  {nm}")
