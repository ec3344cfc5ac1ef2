import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/unet-zoo_b200')
import torch
from tests.test_fusion_gpu import _phiseg_step
def cmp(a, b, tag):
    rows = []
    for n in a:
        den = float(b[n].norm())
        if den < 1e-9: continue
        rows.append((float((a[n] - b[n]).norm()) / den, n, den))
    rows.sort(reverse=True)
    import statistics
    print(tag, 'median', statistics.median(r[0] for r in rows), 'worst:')
    for r in rows[:6]: print('   %.3e  %s  (norm %.3e)' % r)
for B in (4, 12):
    l_u1, g_u1, _ = _phiseg_step(False, False, B=B)
    l_u2, g_u2, _ = _phiseg_step(False, False, B=B)
    l_f, g_f, _ = _phiseg_step(False, True, B=B)
    l_d, g_d, _ = _phiseg_step(True, False, B=B)
    print('B', B, 'losses', l_u1, l_u2, l_f, l_d)
    cmp(g_u1, g_u2, 'unfused vs unfused')
    cmp(g_f, g_u1, 'fused vs unfused')
    cmp(g_d, g_u1, 'det vs unfused')
    cmp(g_f, g_d, 'fused vs det')
