"""Pretty-print the key numbers of bench.py JSON lines read from stdin."""
import json
import sys

for l in sys.stdin:
    l = l.strip()
    if l.startswith('[fdb]'):
        print(l)
    elif l.startswith('{'):
        d = json.loads(l)
        r = d.get('roofline', {})
        s = d.get('solve', {})
        e = d.get('e2e', {})
        print(f"  N={d.get('n_gpus')} ms_per_step={d['ms_per_step']:.4f} value={d['value']/1e9:.2f} G el/s frac={r.get('frac', 0):.3f} "
              f"k1={r.get('ms_kernel_1', 0):.3f} k2={r.get('ms_kernel_2', 0):.3f} setup={d.get('setup_s', 0):.3f}s "
              f"forcing={d.get('ms_forcing', 0):.3f}ms | CG {s.get('iters')} it {s.get('seconds', 0)*1e3:.2f} ms "
              f"{s.get('us_per_iter', 0):.1f} us/it frac={s.get('roofline', {}).get('frac', 0):.3f} | spmv "
              f"{d.get('spmv', {}).get('ms', 0)*1e3:.1f} us frac={d.get('spmv', {}).get('roofline', {}).get('frac', 0):.3f} | e2e "
              f"{e.get('value', 0)/1e6:.1f} M el/s ({e.get('seconds_per_step', 0)*1e3:.1f} ms) | clocks {d.get('clocks')}"
              + (f" | cpu {d['cpu_baseline']['value']:.0f} el/s" if 'cpu_baseline' in d else ''))
    elif 'rror' in l:
        print(l[:300])
