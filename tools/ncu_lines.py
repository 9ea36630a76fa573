"""Per-source-line warp-stall samples of one kernel: joins the SASS page of an ncu report with nvdisasm's line table.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex mangled_name_substring [top]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
si, ii = h.index("# Samples"), h.index("Instructions Executed")
sass = [(r[1].strip(), int(r[si] or 0), int(r[ii] or 0)) for r in rows[hdr + 1:] if len(r) > si and r[0].startswith("0x")]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "artgpn_b200", "libartgpn_b200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# locate the function's text section
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and mangled in l)
lines, cur, inl = [], None, None
for l in dis[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
n = min(len(lines), len(sass))
print("sass instructions: ncu %d, nvdisasm %d" % (len(sass), len(lines)))
agg = {}
for k in range(n):
    key = lines[k]
    a = agg.setdefault(key, [0, 0])
    a[0] += sass[k][1]
    a[1] += sass[k][2]
tot = sum(a[0] for a in agg.values()) or 1
src_cache = {}
def src(key):
    if key is None:
        return ""
    f = os.path.join(ROOT, "artgpn_b200", "csrc", key[0])
    if f not in src_cache:
        src_cache[f] = open(f).read().splitlines() if os.path.exists(f) else []
    L = src_cache[f]
    return L[key[1] - 1].strip()[:100] if 0 < key[1] <= len(L) else ""
print("total samples %d" % tot)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6d %5.1f%%  inst %9d  %s:%s | %s" % (a[0], 100.0 * a[0] / tot, a[1], key[0] if key else "?", key[1] if key else "?", src(key)))
