#!/usr/bin/env python
"""Per-kernel table of a bench.py JSON line:  python tools/kt.py gpurun_out/bench.json [...]"""
import json, sys
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"== {path}: {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.4f} ms/step  e2e {(d.get('e2e') or {}).get('value')}")
    for k, v in sorted((d.get('kernels') or {}).items(), key=lambda kv: -kv[1]['ms_total']):
        print(f"   {k:20s} {v['us_per_launch']:8.1f} us x {v['launches'] / d['steps']:.0f}/step  share {v.get('share', 0):.3f}")
