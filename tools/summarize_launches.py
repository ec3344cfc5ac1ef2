"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (last of the profiled steps)."""
import collections
import csv
import sys

path, nsteps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
data = [(r[ik], float(r[iv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > iv]
# a training step ends with the fused optimizer launch (adam_pack_kernel; adam_batched_kernel before round 2c): take the
# launches between the last two of them
ends = [i for i, (k, _) in enumerate(data) if 'adam_pack_kernel' in k] or \
    [i for i, (k, _) in enumerate(data) if 'adam_batched_kernel' in k]
if len(ends) >= 2:
    step = data[ends[-2] + 1:ends[-1] + 1]
else:
    step = data[len(data) - len(data) // nsteps:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in step:
    k = k.split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    if k.startswith('at::'):
        k = 'torch:' + k[:72]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print('launches in one step: %d   sum of kernel durations: %.2f ms (ncu: serialised, cold caches)' % (len(step), tot / 1e6))
print('%6s %11s %6s  %s' % ('count', 'total us', 'share', 'kernel'))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%6d %11.1f %5.1f%%  %s' % (c, t / 1e3, 100 * t / tot, k))
