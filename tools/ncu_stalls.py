"""Summarise an ncu source-page CSV (ncu -i rep --page source --csv): stall reasons, hot SASS regions, no_inst stalls vs preceding opcode."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s in stalls: tot[s] += f(r, s)
S = sum(tot.values())
print("stall samples total", S)
for s, v in tot.most_common(): print("  %-26s %6.2f%%" % (s, 100 * v / S))
inst = sum(f(r, "Instructions Executed") for r in data)
print("warp instructions executed", inst)
# no_inst by opcode of the instruction and by previous opcode
byop = collections.Counter(); byprev = collections.Counter(); execop = collections.Counter()
prev = None
for r in data:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op = m.group(1) if m else "?"
    byop[op] += f(r, "stall_no_inst"); execop[op] += f(r, "Instructions Executed")
    if prev: byprev[prev] += f(r, "stall_no_inst")
    prev = op
print("no_inst samples by opcode at the stalled pc:", [(k, int(v)) for k, v in byop.most_common(12)])
print("no_inst samples by PREVIOUS opcode:", [(k, int(v)) for k, v in byprev.most_common(12)])
print("executed warp-instr by opcode:", [(k, "%.1f%%" % (100 * v / inst)) for k, v in execop.most_common(25)])
# hot regions: 256-instruction windows
win = 256
agg = []
for i in range(0, len(data), win):
    chunk = data[i:i + win]
    agg.append((sum(f(r, "# Samples") for r in chunk), sum(f(r, "stall_no_inst") for r in chunk), sum(f(r, "Instructions Executed") for r in chunk), i))
print("top windows by samples (samples, no_inst, executed, start index):")
for a in sorted(agg, reverse=True)[:15]: print("  ", [int(x) for x in a])
