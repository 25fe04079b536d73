"""Aggregate an ncu source-page CSV by the CUDA source FUNCTION each SASS instruction's innermost line belongs to.
Usage: ncu_by_func.py src.csv nvdisasm.txt kernelSubstring csrcDir"""
import collections, csv, os, re, sys, bisect
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
kern = sys.argv[3]; src = sys.argv[4]
starts = {}
for fn in os.listdir(src):
    ls = open(os.path.join(src, fn)).read().splitlines(); st = []
    for i, l in enumerate(ls, 1):
        m = re.match(r'^(?:template\s*<[^>]*>\s*)?(?:MLB_\w+|static|__global__|extern "C"|__device__|inline)[\w\s\*&<>:,]*?\b(\w+)\s*\(', l)
        if m and not l.startswith("#"): st.append((i, m.group(1)))
    starts[fn] = st
def funcOf(k):
    if not k: return "?"
    st = starts.get(k[0])
    if not st: return k[0]
    j = bisect.bisect_right([s[0] for s in st], k[1]) - 1
    return "%s:%s" % (k[0].split(".")[0][4:], st[j][1]) if j >= 0 else k[0]
lines = []; inside = False; cur = None
for line in open(sys.argv[2]):
    if line.startswith(".text."):
        inside = kern in line; continue
    if line.startswith("\t.section") and inside and ".text." not in line: inside = False
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line): lines.append((cur, line))
n = min(len(data), len(lines))
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
agg = collections.defaultdict(collections.Counter)
for i in range(n):
    a = agg[funcOf(lines[i][0])]; sass = data[i][ix["Source"]]
    ex = f(data[i], "Instructions Executed")
    a["samples"] += f(data[i], "# Samples"); a["exec"] += ex
    op = sass.split()[0] if not sass.strip().startswith("@") else sass.split()[1]
    if op.startswith(("DFMA", "DADD", "DMUL", "DSETP", "MUFU.RCP64", "MUFU.RSQ64")): a["fp64"] += ex
    if op.startswith("LDL"): a["ldl"] += ex
    if op.startswith("STL"): a["stl"] += ex
    a["long_sb"] += f(data[i], "stall_long_sb")
tot = sum(a["samples"] for a in agg.values()); totx = sum(a["exec"] for a in agg.values())
print("%-34s %7s %7s %7s %7s %7s %7s" % ("function", "smp%", "exec%", "fp64%", "ldl%", "stl%", "longsb%"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:45]:
    print("%-34s %6.2f%% %6.2f%% %6.1f%% %6.1f%% %6.1f%% %6.1f%%" % (k, 100 * a["samples"] / tot, 100 * a["exec"] / totx, 100 * a["fp64"] / max(a["exec"], 1),
          100 * a["ldl"] / max(a["exec"], 1), 100 * a["stl"] / max(a["exec"], 1), 100 * a["long_sb"] / max(a["samples"], 1)))
