"""GPU: warm-cache per-kernel time table of one beam-5 decode (CUPTI via torch.profiler), B=256."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from imagecaptioning.pytorch_b200 import synthetic as syn       # seeded synthetic weights / inputs (profiling tools never touch oracle/)
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
beam = int(sys.argv[2]) if len(sys.argv) > 2 else 5
FAM = sys.argv[3] if len(sys.argv) > 3 else 'updown'
if FAM == 'transformer':
    model = syn.build_model('transformer', seed=1234, logit_scale=3.0, mode='tc_f16x3', heads=8, **dict(bench.CFG, E=512, H=2048, A=6))
elif FAM == 'aoa':
    model = syn.build_model('aoa', seed=1234, logit_scale=6.0, mode='tc_f16x3', heads=8, **dict(bench.CFG, E=1024, H=1024, A=0))
else:
    model = syn.build_model('updown', seed=1234, logit_scale=12.0, mode='tc_f16x3', **bench.CFG)
fc, att = syn.make_inputs(B, 36, 2048, 2048, seed=1)
fc, att = fc.cuda(), att.cuda()
opt = {'beam_size': beam, 'sample_n': 1}
with torch.no_grad():
    for _ in range(3):
        model(fc, att, None, opt=opt, mode='sample')
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            model(fc, att, None, opt=opt, mode='sample')
        torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[2] for r in rows)
print('%s: warm per-kernel table over 3 decodes (B=%d beam=%d): total kernel time %.2f ms per decode' % (FAM, B, beam, tot / 3e3))
for k, n, t in sorted(rows, key=lambda r: -r[2])[:25]:
    print('%-90s n=%5d  %8.1f us/decode  %6.1f us/launch  %5.1f%%' % (k[:90], n // 3, t / 3, t / n, 100 * t / tot))
