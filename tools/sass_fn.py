"""SASS of one kernel of libbwq.so: python tools/sass_fn.py <substring of the mangled name> [--dump]
prints the opcode histogram (and the listing with --dump)."""
import collections
import re
import subprocess
import sys

lib = "ml_qem_b200/lib/libbwq.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if sys.argv[1] not in name:
        continue
    ops = collections.Counter()
    lines = [l for l in b.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
        if m:
            ops[m.group(2)] += 1
    print(name, "instructions", len(lines))
    print(" ".join(f"{k}:{v}" for k, v in ops.most_common(40)))
    if "--dump" in sys.argv:
        print("\n".join(lines))
