"""GPU: warm per-kernel time table of one UpDown SCST training step (batch 10 x 5 samples)."""
import argparse, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import imagecaptioning.pytorch_b200 as b200
from imagecaptioning.pytorch_b200 import synthetic as syn       # seeded synthetic weights / inputs (profiling tools never touch oracle/)
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10
FAM = sys.argv[2] if len(sys.argv) > 2 else 'updown'
if FAM == 'aoa':
    model = syn.build_model('aoa', seed=1234, logit_scale=6.0, mode='tc_f16x3', heads=8, **dict(bench.CFG, E=1024, H=1024, A=0))
else:
    model = syn.build_model('updown', seed=1234, logit_scale=12.0, mode='tc_f16x3', **bench.CFG)
model.train()
df, ref_len = syn.document_frequency(syn.make_refs(500, 9487, seed=4))
table = b200.rewards.CiderDTable(df, ref_len)
fc, att = syn.make_inputs(B, 36, 2048, 2048, seed=1)
fc, att = fc.cuda(), att.cuda()
gts = syn.make_refs(B, 9487, seed=5)
for _ in range(2):
    model.scst_step(fc, att, gts, table, 5)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        model.scst_step(fc, att, gts, table, 5)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[2] for r in rows)
print('%s SCST step (B=%d, n=5): total kernel time %.2f ms per step' % (FAM, B, tot / 2e3))
for k, n, t in sorted(rows, key=lambda r: -r[2])[:22]:
    print('%-100s n=%5d  %9.1f us/step  %7.1f us/launch  %5.1f%%' % (k[:100], n // 2, t / 2, t / n, 100 * t / tot))
