#!/usr/bin/env python
"""Prints the kernel sequence of one step between t0 and t1 ms (all streams) from a chrome trace."""
import gzip, json, sys
path, t0, t1 = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
d = json.load(gzip.open(path) if path.endswith('.gz') else open(path))
ev = [e for e in d['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
ev.sort(key=lambda e: e['ts'])
n = len(ev) // 3
step = ev[n:2 * n]
s0 = step[0]['ts']
prev = None
for e in step:
    t = (e['ts'] - s0) / 1e3
    if t0 <= t < t1:
        g = e.get('args', {}).get('grid', '')
        print(f"{t:7.3f} s{e['args'].get('stream')} dur={e['dur']:6.1f} grid={str(g):14s} {e['name'].replace('void ','').replace('at::native::','').replace('stcat::','')[:100]}")
