"""Map ncu SASS-level stall samples back to CUDA source lines (needs -lineinfo): python tools/ncu_lines.py report.ncu-rep kernel_substr"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, kname = sys.argv[1], sys.argv[2]
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bioslam_b200", "libbioslam_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
start = [i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and kname in l][0]
line_of_off = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        if line_of_off: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of_off[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (os.environ.get("NCU_IMPORT_ARGS", "").split()), capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# a report with several launches repeats the two header rows per launch: keep launch NCU_LAUNCH (default 0)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(starts) > 1:
    w = int(os.environ.get("NCU_LAUNCH", "0"))
    rows = rows[starts[w]:(starts[w + 1] if w + 1 < len(starts) else len(rows))]
hdr = rows[1]
ia, isamp = hdr.index("Address"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = collections.Counter(); aggst = collections.defaultdict(collections.Counter); total = 0
for r in rows[2:]:
    off = int(r[ia], 16) - base
    n = int(r[isamp])
    total += n
    key = line_of_off.get(off, (None, ""))[0]
    agg[key] += n
    for i in stall_cols:
        v = int(r[i])
        if v: aggst[key][hdr[i]] += v
print("total samples", total)
srcs = {}
for key, n in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
    txt = ""
    if key:
        f = os.path.join(os.path.dirname(so), "csrc", key[0])
        if os.path.exists(f):
            srcs.setdefault(f, open(f).read().splitlines())
            txt = srcs[f][key[1] - 1].strip()[:90]
    st = ", ".join(f"{k[6:]}:{v}" for k, v in aggst[key].most_common(3))
    print(f"{n:6d} {100*n/total:5.1f}%  {key}  [{st}]  {txt}")
