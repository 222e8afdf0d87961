"""Summarise ncu outputs from gpurun_out/ into profiles/ (tracked).  Usage: python tools/summarise_profiles.py r01"""
import collections
import csv
import io
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
os.makedirs('profiles', exist_ok=True)

# ---- launch list: per-kernel totals and shares
rows = []
with open('gpurun_out/%s_launches.csv' % tag) as f:
    lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(io.StringIO(''.join(lines))):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        val = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ns = val * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        rows.append((r['Kernel Name'], ns))
tot = collections.defaultdict(lambda: [0, 0.0])
for k, ns in rows:
    name = k.split('(')[0]
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
with open('profiles/%s_launch_summary.md' % tag, 'w') as f:
    f.write('# %s: ncu launch list of `python bench.py --steps 1 --warmup 1 --no-cpu-baseline` (first 1200 launches)\n\n' % tag)
    f.write('`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200` -- cold-cache, serialised: compare shares, not absolutes.\n\n')
    f.write('| kernel | launches | total us | share |\n|---|---:|---:|---:|\n')
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write('| `%s` | %d | %.1f | %.1f%% |\n' % (name[:110], n, ns / 1e3, 100 * ns / total))
    f.write('\nTotal %.2f ms over %d launches.\n' % (total / 1e6, len(rows)))
print(open('profiles/%s_launch_summary.md' % tag).read())

# ---- full capture: key metrics per captured launch
rep = 'gpurun_out/%s_gemm.ncu-rep' % tag
if os.path.exists(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
            'launch__registers_per_thread', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
            'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'sm__cycles_elapsed.max',
            'sm__cycles_elapsed.max.per_second', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active']
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units = rd[0], rd[1]
    with open('profiles/%s_gemm_full.md' % tag, 'w') as f:
        f.write('# %s: `ncu --set full --clock-control none -k regex:gemm_tc -s 12 -c 5` (GEMM launches of consecutive decode steps: gemm_tc_pair_kernel<144,3> = att_lstm / lang_lstm / logit, gemm_tc_kernel<64,3,1,1> = h2att)\n\n' % tag)
        cols = [i for i, h in enumerate(hdr) if h in want or h in ('Kernel Name', 'Grid Size', 'Block Size')]
        for r in rd[2:]:
            f.write('## launch id %s\n\n' % r[0])
            for i in cols:
                f.write('- %s = %s %s\n' % (hdr[i], r[i], units[i]))
            f.write('\n')
    print(open('profiles/%s_gemm_full.md' % tag).read()[:6000])

# ---- non-GEMM kernels of the step (one launch each)
rep = 'gpurun_out/%s_small.ncu-rep' % tag
if os.path.exists(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units = rd[0], rd[1]
    with open('profiles/%s_small_kernels.md' % tag, 'w') as f:
        f.write('# %s: `ncu --set full --clock-control none` of the non-GEMM kernels of one decode step (B=256, beam 5)\n\n' % tag)
        cols = [i for i, h in enumerate(hdr) if h in want or h in ('Kernel Name', 'Grid Size', 'Block Size')]
        for r in rd[2:]:
            f.write('## %s\n\n' % r[hdr.index('Kernel Name')].split('(')[0])
            for i in cols:
                if hdr[i] != 'Kernel Name':
                    f.write('- %s = %s %s\n' % (hdr[i], r[i], units[i]))
            f.write('\n')
