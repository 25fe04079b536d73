"""Local-memory traffic and long-scoreboard stalls per source line (innermost inlined frame), from an ncu source-page CSV and `nvdisasm -gi`
of the same binary. Usage: ncu_local_by_line.py src.csv nvdisasm_gi.txt kernelSubstring [top]"""
import collections, csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; d = rows[2:]
kern = sys.argv[3]; top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
frames = []; inside = False; chain = []; done = True
for line in open(sys.argv[2]):
    if line.startswith(".text."):
        inside = kern in line; chain = []; continue
    if line.startswith("\t.section") and inside and ".text." not in line: inside = False
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        fr = (os.path.basename(m.group(1)), int(m.group(2)))
        if done: chain = [fr]; done = False
        else: chain.append(fr)
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
        frames.append(tuple(chain)); done = True
def val(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
agg = collections.defaultdict(collections.Counter)
for fr, r in zip(frames, d):
    mine = [f for f in fr if f[0].startswith("mlb_")]
    key = mine[0] if mine else (fr[0] if fr else ("?", 0))
    a = agg[key]; ex = val(r, "Instructions Executed")
    s = r[ix["Source"]].split(); op = s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?")
    a["ex"] += ex; a["smp"] += val(r, "# Samples"); a["lsb"] += val(r, "stall_long_sb")
    if op.startswith("LDL"): a["ldl"] += ex
    if op.startswith("STL"): a["stl"] += ex
T = sum(a["smp"] for a in agg.values()); X = sum(a["ex"] for a in agg.values()); LS = sum(a["lsb"] for a in agg.values())
LD = sum(a["ldl"] for a in agg.values()); ST = sum(a["stl"] for a in agg.values())
print("warp instr %.3g, LDL %.3g (%.1f%%), STL %.3g (%.1f%%)" % (X, LD, 100 * LD / X, ST, 100 * ST / X))
print("%-24s %7s %7s %8s %8s %8s" % ("line", "smp%", "exec%", "lsb%", "ldl%", "stl%"))
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1]["lsb"]))[:top]:
    print("%-24s %6.2f%% %6.2f%% %7.2f%% %7.2f%% %7.2f%%" % ("%s:%d" % k, 100 * a["smp"] / T, 100 * a["ex"] / X, 100 * a["lsb"] / LS, 100 * a["ldl"] / LD, 100 * a["stl"] / ST))
