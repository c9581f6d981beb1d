"""Regenerates profiles/r02_ncu_summary.md from the raw ncu exports (ncu -i X.ncu-rep --page raw --csv) and the launch
list committed beside it.

    python profiles/make_summary_r02.py [bench.json]

The optional argument is a bench.py JSON line of an UN-profiled run; its event-timed numbers are quoted next to the ncu
durations (a number printed under ncu is never a bench value)."""
import collections
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def raw(name):
    p = os.path.join(HERE, name)
    if not os.path.exists(p):
        return []
    rows = list(csv.reader(open(p)))
    return [(dict(zip(rows[0], r)), dict(zip(rows[0], rows[1]))) for r in rows[2:]]


def fnum(x):
    return float(x.replace(',', ''))


def scaled(d, u, key):
    v, unit = fnum(d[key]), u[key]
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
                "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1.0)


# capture -> (workload text, algorithmic MB per launch)
CAPS = [
    ("r02_ncu_persist_c4_raw.csv", "C4: 3D P1 stiffness, 10,110,954 tets (persistent kernel, the default)", 172 * 10110954 / 1e6),
    ("r02_ncu_fused_plain_c4_raw.csv", "C4, plain fused kernel (FDB_FUSED_PERSIST=0; captured before the persistent kernel existed)", 172 * 10110954 / 1e6),
    ("r02_ncu_fused_c2_raw.csv", "C2: 2D P1 stiffness, 3,998,792 triangles", 112 * 3998792 / 1e6),
    ("r02_ncu_persist_c3_raw.csv", "C3: 2D P2 ADR (non-symmetric), 2,000,000 triangles (persistent kernel, the default)", 400 * 2000000 / 1e6),
    ("r02_ncu_fused_c3_raw.csv", "C3, plain fused kernel (FDB_FUSED_PERSIST_P2=0)", 400 * 2000000 / 1e6),
    ("r02_ncu_fused_p2tet_raw.csv", "C5-sized slab: 3D P2 stiffness, n=76, 2,633,856 tets", 663 * 2633856 / 1e6),
    ("r02_ncu_cg_raw.csv", "C4 CG iteration (1,728,000 dofs, nnz 25,575,838)", None),
    ("r02_ncu_rowfill_raw.csv", "C4 pattern build, row-wise (setup)", None),
]
ALG_CG = {"k_spmv_sell": 341.5, "k_cg_update": 6 * 8 * 1.728, "k_cg_direction": 3 * 8 * 1.728}

out = ["# Round 2 ncu evidence (1x B200)\n",
       "Full captures: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n>` around "
       "`tools/ab_assembly.py` / `tools/solver_ab.py` (the same library calls `bench.py` makes), under `gpurun` "
       "(`tools/r2_call1.sh`, `tools/r2_call18.sh`). Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` around "
       "`python bench.py --steps 2 --warmup 1 --min-warmup-s 0 --no-cpu-baseline --e2e-steps 1 --no-extra` "
       "(`r02_launches.csv`). Numbers printed by a run under ncu are never bench values; per-launch times under ncu are "
       "cold-cache and serialised, so shares are comparable, absolutes are not. Raw one-row-per-launch exports: `r02_ncu_*_raw.csv`. "
       "Regenerate with `python profiles/make_summary_r02.py`.\n",
       "## Per-kernel summary\n",
       "| workload | kernel | time us | DRAM read MB | DRAM write MB | traffic MB | algorithmic MB | traffic / algorithmic | DRAM % of peak | LSU wavefront % | warps active % | regs | grid x block | L2 hit % |",
       "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|"]
for fname, what, alg in CAPS:
    for d, u in raw(fname):
        short = d['Kernel Name'].split('(')[0].replace('void ', '')
        t = scaled(d, u, 'gpu__time_duration.sum')
        rd, wr = scaled(d, u, 'dram__bytes_read.sum'), scaled(d, u, 'dram__bytes_write.sum')
        a = alg
        if a is None:
            a = next((v for k, v in ALG_CG.items() if k in short), None)
        lsu = d.get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', '0')
        out.append(f"| {what} | `{short}` | {t:.1f} | {rd:.1f} | {wr:.1f} | {rd + wr:.1f} | {'%.1f' % a if a else '-'} | "
                   f"{'%.2f' % ((rd + wr) / a) if a else '-'} | "
                   f"{fnum(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']):.1f} | {fnum(lsu):.1f} | "
                   f"{fnum(d['sm__warps_active.avg.pct_of_peak_sustained_active']):.1f} | {d['launch__registers_per_thread']} | "
                   f"{d['launch__grid_size']} x {d['launch__block_size']} | {fnum(d['lts__t_sector_hit_rate.pct']):.1f} |")

ll = os.path.join(HERE, 'r02_launches.csv')
if os.path.exists(ll):
    lines = [l for l in open(ll) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row['Kernel Name']
        k = k[:k.index('(')] if '(' in k else k
        v, u = fnum(row['Metric Value']), row['Metric Unit']
        v = v / 1e3 if u in ('ns', 'nsecond') else v * 1e3 if u in ('ms', 'msecond') else v
        agg.setdefault(k.replace('void ', ''), []).append(v)
    out.append("\n## Launch list of the bench command (cold-cache, serialised: shares, not absolutes)\n")
    out.append("| kernel | launches | mean us | total ms |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / 1e3:.2f} |")

    def mean(key):
        v = [x for k, vs in agg.items() if key in k for x in vs]
        return sum(v) / len(v) if v else float('nan')
    sp, up, di = mean('k_spmv_sell<1'), mean('k_cg_update'), mean('k_cg_direction')
    out.append(f"\nTimed region of `bench.py` (one step) = one `k_fused_persist` launch: {mean('k_fused_persist'):.1f} us under ncu "
               f"= 100 % of the step's kernels.")
    out.append(f"CG iteration = `k_spmv_sell<1,c16>` {sp:.1f} us ({100 * sp / (sp + up + di):.0f} %) + `k_cg_update` {up:.1f} us + "
               f"`k_cg_direction` {di:.1f} us = {sp + up + di:.1f} us under ncu.")
if len(sys.argv) > 1:
    b = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    r, s, e = b['roofline'], b['solve'], b['e2e']
    out.append(f"\n## Un-profiled bench run of the same build (`python bench.py`, CUDA events)\n")
    out.append(f"- assembly step {b['ms_per_step']:.4f} ms (kernel {r['ms_kernel_1']:.4f} ms) -> {b['value'] / 1e9:.2f} G tets/s, "
               f"{r['achieved']:.0f} GB/s algorithmic = **{r['frac']:.3f}** of {r['peak']:.0f} GB/s; clocks {b['clocks']}")
    out.append(f"- SpMV {b['spmv']['ms'] * 1e3:.1f} us ({b['spmv']['roofline']['frac']:.3f}); CG {s['iters']} iterations, "
               f"{s['seconds'] * 1e3:.2f} ms, {s['us_per_iter']:.1f} us/iter ({s['roofline']['frac']:.3f})")
    out.append(f"- e2e {e['seconds_per_step'] * 1e3:.2f} ms per `discretize_operator` call from host arrays: {e['breakdown']}")
open(os.path.join(HERE, 'r02_ncu_summary.md'), 'w').write("\n".join(out) + "\n")
print("\n".join(out[4:20]))
