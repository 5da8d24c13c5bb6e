#!/usr/bin/env python
"""Reads a torch-profiler chrome trace (bench.py STCAT_TRACE=...) and prints one step's timeline in windows."""
import gzip, json, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/trace.json.gz'
W = float(sys.argv[2]) if len(sys.argv) > 2 else 500.0
d = json.load(gzip.open(path) if path.endswith('.gz') else open(path))
ev = [e for e in d['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
ev.sort(key=lambda e: e['ts'])
n = len(ev) // 3
step = ev[n:2 * n]
s0 = step[0]['ts']
print('kernels/step', n, 'span ms', (step[-1]['ts'] + step[-1]['dur'] - s0) / 1e3)
win = collections.defaultdict(lambda: collections.defaultdict(float))
nwin = collections.Counter()
for e in step:
    w = int((e['ts'] - s0) // W)
    nm = e['name'].replace('void ', '').replace('stcat::', '').replace('at::native::', '')[:34]
    win[w][nm] += e['dur']; nwin[w] += 1
for w in sorted(win):
    top = sorted(win[w].items(), key=lambda x: -x[1])[:3]
    print(f'{w*W/1e3:5.1f}ms n={nwin[w]:4d} busy={sum(win[w].values()):7.0f}us | ' + ' | '.join(f'{n}:{t:.0f}' for n, t in top))
