"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py launches.csv [--solve]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit'); gi = H.index('Grid Size')
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    ms = v / 1e6 if u.startswith('ns') else (v / 1e3 if u.startswith('us') else v)
    seq.append((r[ki].split('(')[0], r[gi], ms))
tot = collections.defaultdict(float); cnt = collections.Counter()
for k, g, ms in seq: tot[k] += ms; cnt[k] += 1
T = sum(tot.values())
print("kernel,launches,total_ms,share,avg_ms")
for k, v in sorted(tot.items(), key=lambda x: -x[1]): print("%s,%d,%.3f,%.3f,%.4f" % (k, cnt[k], v, v / T, v / cnt[k]))
if "--solve" in sys.argv:
    idx = [i for i, s in enumerate(seq) if 'chain_sweep_kernel' in s[0] and s[1].startswith('(148')]
    i0 = idx[-1]
    print("# one solve (last):")
    j = i0
    while j < len(seq) and any(t in seq[j][0] for t in ('chain_', 'top_')):
        print("#  %-32s %-14s %8.3f ms" % seq[j]); j += 1
    print("#  solve total %.3f ms" % sum(s[2] for s in seq[i0:j]))
