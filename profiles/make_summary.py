"""Regenerates profiles/r01_ncu_summary.md from the raw exports (ncu -i X.ncu-rep --page raw --csv) and the launch list."""
import collections
import csv


def raw(path):
    rows = list(csv.reader(open(path)))
    return [dict(zip(rows[0], r)) for r in rows[2:]]


def fnum(x):
    return float(x.replace(',', ''))


ALG = {'k_fused_assemble': 1739.1, 'k_local_assemble': None, 'k_segmented_reduce': None, 'k_spmv': 341.5,
       'k_cg_update': 6 * 8 * 1.728, 'k_cg_direction': 3 * 8 * 1.728}
out = ["# Round 1 ncu evidence (1x B200, n=119: 10,110,954 tets, 1,728,000 dofs, nnz 25,575,838)\n",
       "All captures: `ncu --set full --clock-control none --import-source on` (or `--metrics gpu__time_duration.sum` for the launch list) around",
       "`python bench.py --steps 2 --warmup 1 --min-warmup-s 0 --no-cpu-baseline --e2e-steps 1` under `gpurun`. Numbers printed by a run under ncu are never bench values;",
       "the bench values quoted in DESIGN.md come from separate un-profiled runs of `python bench.py`. Raw one-row-per-launch exports: `r01_ncu_*_raw.csv`;",
       "launch list: `r01_launches_final.csv` (first path of the round: `r01_launches_v0.*`). Regenerate with `python profiles/make_summary.py`.\n",
       "## Per-kernel summary (final kernels of the round; the two-kernel rows were captured before the last local-matrix tweaks)\n",
       "| kernel | time us | DRAM read MB | DRAM write MB | traffic MB | algorithmic MB | DRAM % of ncu peak | LSU wavefront % | warps active % | regs | L1 hit % | L2 hit % |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for f in ['fused', 'twokernel', 'cg']:
    for d in raw(f'profiles/r01_ncu_{f}_raw.csv'):
        short = d['Kernel Name'].split('(')[0].replace('void ', '')
        t = fnum(d['gpu__time_duration.sum'])
        tu = t if t > 5 else t * 1e3
        rd, wr = fnum(d['dram__bytes_read.sum']), fnum(d['dram__bytes_write.sum'])
        a = ALG[[k for k in ALG if k in short][0]]
        lsu = d.get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', '0')
        out.append(f"| `{short}` | {tu:.1f} | {rd:.1f} | {wr:.1f} | {rd + wr:.1f} | {'%.1f' % a if a else '-'} | "
                   f"{fnum(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']):.1f} | {fnum(lsu):.1f} | "
                   f"{fnum(d['sm__warps_active.avg.pct_of_peak_sustained_active']):.1f} | {d['launch__registers_per_thread']} | "
                   f"{fnum(d['l1tex__t_sector_hit_rate.pct']):.1f} | {fnum(d['lts__t_sector_hit_rate.pct']):.1f} |")
lines = [l for l in open('profiles/r01_launches_final.csv') if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row['Kernel Name']
    k = k[:k.index('(')] if '(' in k else k
    v, u = fnum(row['Metric Value']), row['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v
    agg.setdefault(k, []).append(v)
out.append("\n## Launch list of the same command (`--metrics gpu__time_duration.sum`, cold-cache, serialised: shares, not absolutes)\n")
out.append("| kernel | launches | mean us | total ms |\n|---|---:|---:|---:|")
for k, v in agg.items():
    out.append(f"| `{k[:80]}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / 1e3:.2f} |")
mean = lambda key: (lambda v: sum(v) / len(v))([x for k, v in agg.items() if key in k for x in v])
sp, up, di = mean('k_spmv_sell<1'), mean('k_cg_update'), mean('k_cg_direction')
out.append(f"\nAssembly step = one `k_fused_assemble` launch ({mean('k_fused_assemble'):.1f} us under ncu; bench.py measures 0.444 ms with CUDA events).")
out.append(f"CG iteration = `k_spmv_sell<1,c16>` {sp:.1f} us ({100 * sp / (sp + up + di):.0f} %) + `k_cg_update` {up:.1f} us + "
           f"`k_cg_direction` {di:.1f} us = {sp + up + di:.1f} us (bench: 92.3 us/iter with CUDA events).")
open('profiles/r01_ncu_summary.md', 'w').write("\n".join(out) + "\n")
print("\n".join(out[8:15]))
print(out[-2])
print(out[-1])
