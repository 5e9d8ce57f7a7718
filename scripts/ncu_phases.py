#!/usr/bin/env python
"""warp-stall samples of the rollout kernel aggregated by kernel phase. usage: ncu_phases.py rep"""
import re, sys, collections, subprocess
rep = sys.argv[1]
txt = subprocess.run([sys.executable, 'scripts/ncu_lines.py', rep, '600'], capture_output=True, text=True).stdout
if len(sys.argv) > 2 and sys.argv[2] == 'WORK':
    src = open('recogym_b200/csrc/rg_rollout_fast.cu').read().split('\n')
else:
    src = subprocess.run(['git', 'show', (sys.argv[2] if len(sys.argv) > 2 else 'HEAD') + ':recogym_b200/csrc/rg_rollout_fast.cu'],
                         capture_output=True, text=True).stdout.split('\n')
def find(s):
    return next(i + 1 for i, l in enumerate(src) if s in l)
marks = [('resample(tier2/3)', find('template <int NB, int PP, bool T2>')), ('setup', find('rollout_cta_body(const Params &p)')),
         ('A:compact/shift', find('---- A: compact the live slots')), ('B:pass1', find('---- B: pass 1')),
         ('C:pass2a', find('---- C: pass 2a')), ('C:pass2b', find('---- C: pass 2b')),
         ('C2:resample-call', find('---- C2: re-sampling')), ('D:advance', find('---- D: advance the users')),
         ('end', find('---- drain the prefetched tiles'))]
agg = collections.defaultdict(lambda: [0.0, 0.0])
for ln in txt.split('\n'):
    m = re.match(r'\s*([\d.]+)% smp\s+([\d.]+)% inst\s+(\S+):(\d+)', ln)
    if not m: continue
    f, l = m.group(3), int(m.group(4))
    name = f
    if f == 'rg_rollout_fast.cu':
        name = 'pre'
        for n, start in marks:
            if l >= start: name = n
    agg[name][0] += float(m.group(1)); agg[name][1] += float(m.group(2))
print(txt.split('\n')[0])
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print('%-24s %6.2f%% smp %6.2f%% inst' % (k, v[0], v[1]))
