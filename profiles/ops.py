"""Per-opcode totals of an ncu source-page csv export: python profiles/ops.py gpurun_out/<name>"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1] + "_source.csv")))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in rows[2:]:
    if len(r) < len(hdr): continue
    op = [o for o in r[ix['Source']].strip().split() if not o.startswith('@')][0]
    a = agg[op]; a[0] += int(r[ix['Instructions Executed']]); a[1] += int(r[ix['L1 Wavefronts Shared']] or 0)
    a[2] += int(r[ix['L1 Tag Requests Global']] or 0); a[3] += int(r[ix['# Samples']] or 0)
tot = sum(a[3] for a in agg.values())
print("total inst", sum(a[0] for a in agg.values()))
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][3])[:int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    print(f"  {op:30s} inst {a[0]:>10d} shwave {a[1]:>10d} tagreq {a[2]:>10d} samples {100*a[3]/tot:5.1f}%")
