import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('[fdb]'): print(l)
    elif l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(f"  ms_per_step={d['ms_per_step']:.4f} value={d['value']/1e9:.2f} G/s frac={r['frac']:.3f} local={r['ms_local']:.3f} reduce={r['ms_reduce']:.3f} setup={d['setup_s']:.3f} forcing={d['ms_forcing']:.3f}", ('cg_us=%.1f'%d['solve']['us_per_iter']) if 'solve' in d else '')
    elif 'rror' in l: print(l[:300])
