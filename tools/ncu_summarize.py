"""ncu raw-page CSVs (tools/gpu_probe_r2b.sh) -> profiles/r02_ncu_kernels.md + profiles/r02_ncu_summary.json
    python tools/ncu_summarize.py gpurun_out/r2b_probe_conv_raw.csv gpurun_out/r2b_probe_mem_raw.csv"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csv.field_size_limit(10 ** 9)
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1.0, 'ns': 1e-3, 'ms': 1e3, 'cycle': 1.0, '': 1.0}
# algorithmic bytes / FLOPs of the probe launches (tools/ncu_probe.py, batch 12): kernel-name substring + grid -> figure
B = 12


def algorithmic(name, grid):
    def conv(cin, cout, h):
        return {'flops': 2.0 * B * h * h * cin * cout * 9, 'bytes': B * h * h * (cin + cout) * 2.0 + 9 * cin * cout * 2.0}
    table = {
        ('conv_tc2_kernel<64', '(148, 1, 1)'): conv(128, 128, 128),
        ('conv_tc2_kernel<64', '(48, 1, 1)'): conv(192, 192, 16),
        ('conv_tc2_kernel<32', '(296, 1, 1)'): conv(32, 32, 128),
        ('conv_tc_kernel', '(1, 4, 1)'): conv(192, 192, 2),
    }
    for (n, g), v in table.items():
        if n in name and g == grid:
            return v
    npx = {'c32': B * 128 * 128 * 32, 'c128': B * 128 * 128 * 128}
    if 'bn_apply_train' in name:
        return {'bytes_per_elem': 4}
    if 'bn_bwd_reduce' in name:
        return {'bytes_per_elem': 4}
    if 'bn_bwd_apply_train' in name:
        return {'bytes_per_elem': 6}
    if 'eval_sample_stats' in name:
        return {'bytes': sum(100 * 2 * (128 >> l) ** 2 * 4 for l in range(5)) + 100 * 2048 + 2 * 2 * 16384 * 4 * 10}
    return {}


def main(paths):
    rows_out = []
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]

        def col(suffix):
            for i, h in enumerate(hdr):
                if h == suffix or h.endswith(suffix):
                    return i
            return None
        c = {k: col(v) for k, v in dict(name='Kernel Name', grid='Grid Size', t='gpu__time_duration.sum',
                                        rd='dram__bytes_read.sum', wr='dram__bytes_write.sum',
                                        hmma='sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
                                        cyc='sm__cycles_elapsed.max', regs='launch__registers_per_thread',
                                        l2hit='lts__t_sector_hit_rate.pct', smem='launch__shared_mem_per_block_dynamic').items()}
        seen = {}
        for r in rows[2:]:
            def val(k):
                i = c[k]
                if i is None or r[i] in ('', 'no data', 'n/a'):
                    return None
                try:
                    return float(r[i]) * UNIT.get(units[i], 1.0)
                except ValueError:
                    return None
            name = r[c['name']].replace('void <unnamed>::', '').split('(CUtensorMap')[0].split('(const')[0][:60]
            key = (name, r[c['grid']])
            seen[key] = dict(kernel=name, grid=r[c['grid']], us=val('t'), dram_read=val('rd'), dram_write=val('wr'),
                             tensor_active=(val('hmma') / 4.0 / val('cyc')) if val('hmma') is not None and val('cyc') else None,
                             regs=val('regs'), l2_hit_pct=val('l2hit'))        # the LAST (warm) instance of every (kernel, grid)
        rows_out += list(seen.values())
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    hbm, tf = peaks.get('hbm_gbs', 6650.0), peaks.get('bf16_tflops', 1590.0)
    md = ['| kernel | grid | us (ncu, cold, serialised) | DRAM read MB | DRAM write MB | DRAM GB/s (%% of %.0f) | algorithmic | tensor pipe active | regs |' % hbm,
          '|---|---|---|---|---|---|---|---|---|']
    for r in rows_out:
        alg = algorithmic(r['kernel'], r['grid'])
        traffic = (r['dram_read'] or 0) + (r['dram_write'] or 0)
        gbs = traffic / (r['us'] * 1e-6) / 1e9 if r['us'] else 0
        a = ''
        if 'flops' in alg:
            a = '%.1f GFLOP (%.0f TFLOP/s = %.2f of burst peak), %.1f MB' % (alg['flops'] / 1e9, alg['flops'] / (r['us'] * 1e-6) / 1e12,
                                                                           alg['flops'] / (r['us'] * 1e-6) / 1e12 / tf, alg['bytes'] / 1e6)
            r['algorithmic_flops'], r['algorithmic_bytes'] = alg['flops'], alg['bytes']
        elif 'bytes' in alg:
            a = '%.1f MB' % (alg['bytes'] / 1e6)
            r['algorithmic_bytes'] = alg['bytes']
        md.append('| `%s` | %s | %.1f | %.1f | %.1f | %.0f (%.0f %%) | %s | %s | %d |' % (
            r['kernel'], r['grid'], r['us'], (r['dram_read'] or 0) / 1e6, (r['dram_write'] or 0) / 1e6, gbs, 100 * gbs / hbm, a,
            ('%.2f' % r['tensor_active']) if r['tensor_active'] else '-', int(r['regs'] or 0)))
        r['dram_gb_per_s'] = gbs
    dom = next((r for r in rows_out if 'conv_tc2_kernel<64' in r['kernel'] and r['grid'] == '(148, 1, 1)'), None)
    out = {'source': 'ncu --set full --clock-control none, one warm launch per kernel (tools/gpu_probe_r2b.sh); cold-cache, '
                     'serialised durations', 'kernels': rows_out}
    if dom:
        out['dominant_kernel'] = {'kernel': 'conv_tc2_kernel<64>, 128 -> 128 @128^2, batch 12', 'us': dom['us'],
                                  'dram_bytes_per_launch': (dom['dram_read'] or 0) + (dom['dram_write'] or 0),
                                  'algorithmic_bytes': dom.get('algorithmic_bytes'), 'tensor_pipe_active': dom['tensor_active'],
                                  'metric': 'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg / 4 sub-partitions / '
                                            'sm__cycles_elapsed.max (counts UTCHMMA)'}
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'r02_ncu_summary.json'), 'w'), indent=1)
    open(os.path.join(ROOT, 'profiles', 'r02_ncu_kernels.md'), 'w').write(
        '# ncu --set full, one warm launch per kernel family (round 2)\n\n`bash tools/gpu_probe_r2c.sh` on a B200 (final round-2 code); batch 12 shapes of '
        'PHiSeg-7/5 (tools/ncu_probe.py).  Durations under ncu are cold-cache and serialised (clock control off): use them '
        'for DRAM bytes / counters, the in-step times are in `r02_layer_times_singlestream.txt` and the bench line.\n'
        'Tensor pipe active = `sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg` / 4 / `sm__cycles_elapsed.max` '
        '(this counter does count `UTCHMMA`).\n\n' + '\n'.join(md) + '\n')
    print('\n'.join(md))


if __name__ == '__main__':
    main(sys.argv[1:])
