"""Attribute an ncu source-page CSV to the ENCLOSING major function of every SASS instruction, through the inline chains of
`nvdisasm -gi` (innermost-line attribution, tools/ncu_by_func.py, files everything inlined under the helper it came from: half the kernel
ends up in `operator+`). Usage: ncu_by_owner.py src.csv nvdisasm_gi.txt kernelSubstring csrcDir"""
import bisect, collections, csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; d = rows[2:]
kern, src = sys.argv[3], sys.argv[4]
MAJOR = {"flightLoop", "flightLanes", "chuteLanes", "finishLane", "integrateStep", "stateDerivative", "appliedForce", "finSetForce", "noseConeForce", "bodyTubeForce", "transitionForce", "airProperties",
         "aeroSetup", "rocketInertia", "rocketTimeStep", "triggerEvents", "laneInit", "pinkValue", "turbulenceAt", "atmosphere", "integrateLanes",
         "endDetector", "finStripTransonic", "finStripSubsonic", "cylinderSkinFriction", "forceFromCdCN", "meanWindAt", "gravityForceGlobal",
         "mtTwist", "mtInitByArray", "pinkInit", "runControlSystem", "ndInterpolate", "recoveryForce"}
starts = {}
for fn in os.listdir(src):
    st = []
    for i, l in enumerate(open(os.path.join(src, fn)).read().splitlines(), 1):
        m = re.match(r'^(?:template\s*<[^>]*>\s*)?(?:MLB_\w+|static|__global__|extern "C"|__device__|inline)[\w\s\*&<>:,]*?\b(\w+)\s*\(', l)
        if m and not l.startswith("#"): st.append((i, m.group(1)))
    starts[fn] = st
def funcOf(k):
    st = starts.get(k[0])
    if not st: return k[0]
    j = bisect.bisect_right([s[0] for s in st], k[1]) - 1
    return st[j][1] if j >= 0 else k[0]
owners = []; inside = False; chain = []; done = True; outer = None
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
        outer = (os.path.basename(m.group(3)), int(m.group(4))) if m.group(3) else None
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
        frames = list(chain) + ([outer] if chain and outer else [])
        owners.append(next((funcOf(f) for f in frames if funcOf(f) in MAJOR), "?")); done = True
n = min(len(owners), len(d))
agg = collections.defaultdict(collections.Counter)
def val(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
for o, r in zip(owners[:n], d[:n]):
    a = agg[o]; ex = val(r, "Instructions Executed")
    a["smp"] += val(r, "# Samples"); a["ex"] += ex; a["lsb"] += val(r, "stall_long_sb"); a["bar"] += val(r, "stall_barrier")
    a["thr"] += val(r, "Thread Instructions Executed"); a["wait"] += val(r, "stall_wait")
    s = r[ix["Source"]].split(); op = s[1] if s[0].startswith("@") else s[0]
    if op.startswith("LDL"): a["ldl"] += ex
    if op.startswith("STL"): a["stl"] += ex
T = sum(a["smp"] for a in agg.values()); X = sum(a["ex"] for a in agg.values()); LS = sum(a["lsb"] for a in agg.values())
print("sass rows %d, disassembled %d" % (len(d), len(owners)))
print("%-24s %7s %7s %14s %9s %7s %7s %9s %7s" % ("enclosing function", "smp%", "exec%", "long_sb share", "barrier%", "ldl%", "stl%", "thr/inst", "wait%"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:26]:
    print("%-24s %6.1f%% %6.1f%% %13.1f%% %8.1f%% %6.1f%% %6.1f%% %9.1f %6.1f%%" % (k, 100 * a["smp"] / T, 100 * a["ex"] / X, 100 * a["lsb"] / LS, 100 * a["bar"] / max(a["smp"], 1),
                                                               100 * a["ldl"] / max(a["ex"], 1), 100 * a["stl"] / max(a["ex"], 1), a["thr"] / max(a["ex"], 1), 100 * a["wait"] / max(a["smp"], 1)))
