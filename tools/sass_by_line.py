#!/usr/bin/env python
"""Static SASS instruction histogram by source line for one kernel of a built library (no GPU needed):

    python tools/sass_by_line.py pg_embedding_b200/libpgemb_b200.so 'search_kernelILi1ELb0ELi4' [--top 40]

Extracts the cubin (cuobjdump), disassembles it with line info (nvdisasm -g; the libraries are built with -lineinfo) and
counts instructions per (file, line).  Static counts, not executed counts -- but together with the known trip counts of a
hop they show where the instructions of the traversal are (used for DESIGN.md section 6's per-hop budget)."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    tmp = tempfile.mkdtemp(prefix="sass_by_line_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    in_fn, cur, hist, ops = False, ("?", 0), collections.Counter(), collections.defaultdict(collections.Counter)
    total = 0
    for line in out.splitlines():
        if line.startswith("//--------------------- .text."):
            in_fn = pat in line
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m:
            hist[cur] += 1
            ops[cur][m.group(1).split(".")[0]] += 1
            total += 1
    print(f"{total} instructions in kernels matching {pat!r}")
    for (f, ln), c in hist.most_common(top):
        print(f"{c:6d}  {f}:{ln:<5d} {dict(ops[(f, ln)].most_common(4))}")


if __name__ == "__main__":
    main()
