"""Development tool: per-source-line executed warp instructions of one kernel.
nvdisasm -g line info of the cubin joined (by instruction order) with the ncu source page (csv).
usage: line_attrib.py <nvdisasm -g -c output> <ncu --page source --csv output> <mangled kernel name> <source file> <units>"""
import collections, csv, re, sys
dis, ncu, kern, srcfile, units = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
cur = None; instr = []; on = False
for ln in open(dis):
    if ln.startswith(".text."):
        on = ln.strip().rstrip(":") == ".text." + kern
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        instr.append((cur, m.group(2).strip()))
rows = list(csv.reader(open(ncu)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
h = rows[hi[0] + 1]; data = rows[hi[0] + 2:]
iE, iN = h.index("Instructions Executed"), h.index("# Samples")
print("instructions: nvdisasm %d, ncu %d" % (len(instr), len(data)))
agg = collections.defaultdict(lambda: [0.0, 0.0])
for i in range(min(len(instr), len(data))):
    a = agg[instr[i][0]]
    a[0] += float(data[i][iE] or 0); a[1] += float(data[i][iN] or 0)
ts = sum(v[1] for v in agg.values())
src = open(srcfile).read().split("\n")
base = srcfile.split("/")[-1]
tot = 0
for key, v in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    f, l = key if key else ("?", 0)
    if v[0] / units < 0.5:
        continue
    tot += v[0] / units
    text = src[l - 1].strip()[:100] if f == base and l - 1 < len(src) else ""
    print("%-14s %4d %7.1f %5.1f%%  %s" % (f, l, v[0] / units, 100 * v[1] / ts, text))
print("total listed %.1f" % tot)
