"""Static SASS instruction count per source line for one kernel (needs -lineinfo). Usage: sass_by_line.py <lib.so> <kernel substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
lib, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cnt = collections.Counter(); cur = None; inside = False; ops = collections.Counter()
for line in txt.splitlines():
    if line.startswith(".text."):
        inside = kern in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        cnt[cur] += 1
        ops[m.group(1).split(".")[0]] += 1
print("total instructions", sum(cnt.values()))
byfile = collections.Counter()
for (f, l), c in cnt.items():
    byfile[f] += c
print(byfile.most_common(12))
print(ops.most_common(25))
for (f, l), c in cnt.most_common(top):
    print(c, f, l)
