"""Join an ncu source-page CSV (SASS rows, in program order) with `nvdisasm --print-line-info` of the SAME binary to get
stall samples and executed instructions per CUDA source line (innermost inlined frame and outermost function).
Usage: ncu_by_line.py src.csv nvdisasm.txt kernelSubstring [top]"""
import collections, csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
kern = sys.argv[3]; top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = []; inside = False; cur = None
for line in open(sys.argv[2]):
    if line.startswith(".text."):
        inside = kern in line; continue
    if line.startswith("\t.section") and inside and ".text." not in line: inside = False
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line): lines.append(cur)
print("sass rows", len(data), "disasm instr", len(lines))
n = min(len(data), len(lines))
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for i in range(n):
    a = agg[lines[i]]
    a["samples"] += f(data[i], "# Samples"); a["exec"] += f(data[i], "Instructions Executed"); a["thr"] += f(data[i], "Thread Instructions Executed")
    for s in stalls: a[s] += f(data[i], s)
tot = sum(a["samples"] for a in agg.values()); totx = sum(a["exec"] for a in agg.values())
print("%-28s %7s %7s %6s | top stalls" % ("line", "smp%", "exec%", "thr/w"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((s, a[s]) for s in stalls), key=lambda x: -x[1])[:3]
    print("%-28s %6.2f%% %6.2f%% %6.1f | %s" % ("%s:%s" % k if k else "?", 100 * a["samples"] / tot, 100 * a["exec"] / totx, a["thr"] / max(a["exec"], 1),
                                             ", ".join("%s %.0f%%" % (s[6:], 100 * v / max(a["samples"], 1)) for s, v in st)))
