"""Executed instructions of one opcode (or stall samples with OPCODE = '@samples') per inlined call chain (needs -lineinfo):
python tools/ncu_inst_lines.py report.ncu-rep kernel_substr OPCODE [top] [so_path]"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, kname, opc = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
so = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bioslam_b200", "libbioslam_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin") and "host_capi" not in f][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and kname in l][0]
chain_of_off = {}
chain = []
fresh = True
short = {"kernels_solve.cuh": "S", "kernels_sweep_pipe.cuh": "P", "cuda_compat.h": "C", "kernels_local.cuh": "L", "bioslam_b200.cu": "B"}
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        if chain_of_off: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        f = os.path.basename(m.group(1))
        fr = short.get(f, f) + ":" + m.group(2)
        if not chain or chain[-1] != fr: chain.append(fr)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        chain_of_off[int(m.group(1), 16)] = tuple(chain)
        fresh = True
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia = hdr.index("Address"); isrc = hdr.index("Source")
ie = hdr.index("# Samples") if opc == "@samples" else [i for i, h in enumerate(hdr) if "Instructions Executed" in h][0]
base = int(rows[2][ia], 16)
agg = collections.Counter(); total = 0
for r in rows[2:]:
    if opc != "@samples" and opc not in r[isrc]: continue
    off = int(r[ia], 16) - base
    n = int(r[ie]); total += n
    ch = chain_of_off.get(off, ())
    # drop the innermost helper frames (cuda_compat.h) and the outermost kernel frame
    key = " < ".join(c for c in ch if not c.startswith("C:"))
    agg[key] += n
print("total", opc, ":", total)
for key, n in agg.most_common(top):
    print(f"{n:12d} {100*n/total:5.1f}%  {key}")
