#!/usr/bin/env python
"""Where does ptxas spill?  Count local-memory loads/stores (LDL/STL) per function and source line of a -lineinfo object.
usage: spill_lines.py file.o [function-substring]"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj = os.path.abspath(sys.argv[1])
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", obj], cwd=td, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    fn, line = "?", "?"
    cnt = collections.Counter()
    for l in txt.splitlines():
        m = re.match(r"\s*(?:\.text\.)?(_Z\w+|\$\w+\$\w+):", l)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = f"{os.path.basename(m.group(1))}:{m.group(2)}"
            continue
        m = re.search(r"\b(STL|LDL)\b", l)
        if m and want in fn:
            cnt[(fn[-60:], line, m.group(1))] += 1
    for (f, ln, op), n in sorted(cnt.items(), key=lambda kv: -kv[1])[:60]:
        print(f"{n:4d} {op} {ln:28s} {f}")


if __name__ == "__main__":
    main()
