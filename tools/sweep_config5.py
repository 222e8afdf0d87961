"""GPU: BASELINE.json configs[4] -- UpDown beam-width sweep {1,3,5,10} x batch {64..1024}, 1 GPU: captions/s and the fraction of the
tensor roofline the GEMM launches reach at each point.  Writes gpurun_out/<tag>_config5_sweep.{md,json} (copied to profiles/)."""
import json, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from imagecaptioning.pytorch_b200 import synthetic as syn       # seeded synthetic weights / inputs (profiling tools never touch oracle/)

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
mode = sys.argv[2] if len(sys.argv) > 2 else 'tc_f16x3'
peak = 1442.4
pp = os.path.join(REPO, 'MEASURED_PEAKS.json')
if os.path.exists(pp):
    peak = float(json.load(open(pp))['bf16_tflops_sustained'])
model = syn.build_model('updown', seed=1234, logit_scale=12.0, mode=mode, **bench.CFG)
rows = []
for beam in (1, 3, 5, 10):
    for B in (64, 128, 256, 512, 1024):
        ins = [syn.make_inputs(B, 36, 2048, 2048, seed=7 + i) for i in range(2)]
        ins = [(a.cuda(), b.cuda()) for a, b in ins]
        opt = {'beam_size': beam, 'sample_n': 1}
        with torch.no_grad():
            for i in range(2):
                model(*ins[i % 2], None, opt=opt, mode='sample')
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 4
            e0.record()
            for i in range(n):
                model(*ins[i % 2], None, opt=opt, mode='sample')
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            model.set_profiling(True)
            model(*ins[0], None, opt=opt, mode='sample')
            prof = model.read_profile()
            model.set_profiling(False)
        g_ms = sum(v[0] for v in prof.values())
        g_fl = sum(v[1] for v in prof.values())
        tf = g_fl / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
        rows.append({'beam': beam, 'batch': B, 'rows': B * beam, 'ms_per_batch': ms, 'captions_per_s': B / (ms / 1e3), 'gemm_ms': g_ms,
                     'gemm_tflops_algorithmic': tf, 'frac_of_bf16_peak': tf / peak, 'frac_of_3pass_ceiling': tf / (peak / 3), 'gemm_share': g_ms / ms})
        print(rows[-1], flush=True)
        del ins
        torch.cuda.empty_cache()
os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
json.dump({'mode': mode, 'peak_tflops': peak, 'rows': rows}, open(os.path.join(REPO, 'gpurun_out', '%s_config5_sweep.json' % tag), 'w'), indent=1)
with open(os.path.join(REPO, 'gpurun_out', '%s_config5_sweep.md' % tag), 'w') as f:
    f.write('# %s: UpDown beam x batch sweep (BASELINE configs[4]), 1 x B200, numeric mode %s, T=20, 36x2048 features, V=9487\n\n' % (tag, mode))
    f.write('CUDA-event time of `model(fc, att, None, opt={beam_size}, mode="sample")`, features resident, 2 warm-up + 4 timed batches. GEMM columns: all '
            '`gemm_tc_kernel` launches of one batch (algorithmic 2MNK FLOPs / summed CUDA-event time) against the measured bf16 peak %.0f TFLOP/s; the '
            '3-pass split-fp16 scheme caps the algorithmic rate at 1/3 of it.\n\n' % peak)
    f.write('| beam | batch | rows | ms/batch | captions/s | GEMM ms | GEMM TFLOP/s (alg.) | of bf16 peak | of 3-pass ceiling | GEMM share |\n|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n')
    for r in rows:
        f.write('| %d | %d | %d | %.2f | %.0f | %.2f | %.0f | %.3f | %.2f | %.0f%% |\n' % (r['beam'], r['batch'], r['rows'], r['ms_per_batch'], r['captions_per_s'],
                r['gemm_ms'], r['gemm_tflops_algorithmic'], r['frac_of_bf16_peak'], r['frac_of_3pass_ceiling'], 100 * r['gemm_share']))
