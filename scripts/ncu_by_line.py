"""Attribute an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv) to CUDA source lines using
nvdisasm -g line info of the same cubin.  usage: ncu_by_line.py src.csv dis.txt mangled_kernel [top]"""
import csv, re, sys, collections
src_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: collect (file,line) per instruction in order for the kernel
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
loc = []
cur = ("?", 0)
inl = None
for l in lines[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", l):
        loc.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
assert len(body) == len(loc), (len(body), len(loc))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = [0, 0, 0, 0, 0]
for r, lc in zip(body, loc):
    v = [int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), int(r[ix["stall_long_sb"]] or 0),
         int(r[ix["stall_barrier"]] or 0), int(r[ix["stall_wait"]] or 0)]
    for k in range(5):
        agg[lc][k] += v[k]; tot[k] += v[k]
print("total samples %d inst %d long_sb %d barrier %d wait %d" % tuple(tot))
print("%-28s %8s %6s %12s %6s %8s %8s" % ("file:line", "samples", "%", "inst", "%", "long_sb", "barrier"))
for lc, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %8d %6.2f %12d %6.2f %8d %8d" % (f"{lc[0]}:{lc[1]}", v[0], 100 * v[0] / tot[0], v[1], 100 * v[1] / tot[1], v[2], v[3]))
# by file
byf = collections.defaultdict(lambda: [0, 0])
for lc, v in agg.items():
    byf[lc[0]][0] += v[0]; byf[lc[0]][1] += v[1]
for f, v in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print("FILE %-28s samples %6.2f%% inst %6.2f%%" % (f, 100 * v[0] / tot[0], 100 * v[1] / tot[1]))
