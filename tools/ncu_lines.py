"""Per-CUDA-source-line instruction and stall-sample totals of one profiled launch in an .ncu-rep:
python tools/ncu_lines.py rep [launch_index]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# a launch = run of blocks; a new launch starts when a "Function Name" row follows rows of another launch and addresses restart
launches = []; cur_rows = None; seen_addr = None
h = None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        h = r; continue
    launches.append(r)
ix = {}
for i, x in enumerate(h): ix.setdefault(x, i)
# split launches by detecting address re-occurrence count
per_launch = collections.defaultdict(lambda: collections.OrderedDict())
count = collections.Counter(); cur = None
for r in launches:
    if len(r) != len(h): continue
    if r[0] != "":
        cur = (int(r[0]), r[1].strip()); continue
    a = r[2]
    if not a.startswith("0x"): continue
    key = (a, cur)
    k = count[key]; count[key] += 1
    per_launch[k][key] = r
# the cuda,sass view lists every launch once per file block; index k = k-th occurrence
L = per_launch[want]
per = collections.OrderedDict(); tot_i = tot_s = 0; seen = set()
for (a, cur), r in L.items():
    if a in seen: continue
    seen.add(a)
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    e = per.setdefault(cur, [0, 0, collections.Counter()])
    e[0] += n; e[1] += s; tot_i += n; tot_s += s
    op = r[3].split()
    if op: e[2][(op[1] if op[0].startswith("@") else op[0]).split(".")[0]] += n
print("launches seen", len(per_launch), "total inst", tot_i, "samples", tot_s)
for (ln, src), (n, s, ops) in sorted(per.items(), key=lambda kv: -kv[1][0])[:50]:
    print(f"{ln:4d} {100*n/max(1,tot_i):5.1f}% inst {100*s/max(1,tot_s):5.1f}% smp  {src[:80]:80s} | " + " ".join(f"{k}:{v*100//max(1,n)}" for k, v in ops.most_common(5)))
