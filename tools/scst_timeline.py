"""GPU: timeline of one SCST training step (engine step + Adam, as bench.py runs it) from the torch profiler's device trace:
span, per-stream busy time, idle gaps on the main stream, and the phase boundaries (sampling forward / reward / BPTT / optimizer)
located by marker kernels.  usage: python tools/scst_timeline.py [aoa|updown] [out.json]"""
import argparse as ap, json, os, sys, tempfile
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import imagecaptioning.pytorch_b200 as b200
from imagecaptioning.pytorch_b200 import synthetic as syn
import bench

FAM = sys.argv[1] if len(sys.argv) > 1 else 'aoa'
B, n = 10, 5
CFG = bench.CFG
if FAM == 'aoa':
    model = syn.build_model('aoa', seed=1234, logit_scale=6.0, mode='tc_f16x3', device='cuda', heads=8, **dict(CFG, E=1024, H=1024, A=0))
else:
    model = syn.build_model('updown', seed=1234, logit_scale=12.0, mode='tc_f16x3', device='cuda', **CFG)
model.train()
df, ref_len = syn.document_frequency(syn.make_refs(1000, CFG['V'], seed=4))
b200.rewards.reset_scorer()
b200.rewards.init_scorer(b200.rewards.CiderDTable(df, ref_len))
opt = ap.Namespace(sc_sample_method='greedy', sc_beam_size=1, train_sample_method='sample', train_beam_size=1, train_sample_n=n,
                   cider_reward_weight=1, bleu_reward_weight=0)
lw = b200.B200LossWrapper(model, opt)
optim = b200.optim.FusedAdam(model.parameters(), lr=5e-5, clip_value=0.1)
fc, att = syn.make_inputs(B, 36, CFG['F_fc'], CFG['F_att'], seed=99)
fc, att = fc.pin_memory(), att.pin_memory()
gts = syn.make_refs(B, CFG['V'], seed=5)
idx = torch.arange(B)


def step():
    out = lw(fc.to('cuda', non_blocking=True), att.to('cuda', non_blocking=True), None, None, None, gts, idx, True, False, False)
    optim.zero_grad(set_to_none=True)
    out['loss'].backward()
    optim.step()
    return float(out['loss'].detach())


for _ in range(4):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), 'trace.json')
prof.export_chrome_trace(path)
ev = json.load(open(path))['traceEvents']
ker = [e for e in ev if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
ker.sort(key=lambda e: e['ts'])
# split into steps at the H2D copies of the features (two pinned copies open each step)
starts = [i for i, e in enumerate(ker) if e.get('cat') == 'gpu_memcpy' and 'HtoD' in e['name'] and e['dur'] > 20]
starts = starts[::2] if len(starts) >= 6 else starts
lo = starts[1]
hi = starts[2] if len(starts) > 2 else len(ker)
K = ker[lo:hi]
t0 = K[0]['ts']
span = K[-1]['ts'] + K[-1]['dur'] - t0
streams = {}
for e in K:
    streams.setdefault(e['args'].get('stream'), []).append(e)
main = max(streams, key=lambda s: len(streams[s]))
print('%s SCST step: %d device activities over %.2f ms; streams: %s' % (FAM, len(K), span / 1e3, {s: len(v) for s, v in streams.items()}))
for s, v in streams.items():
    print('  stream %s: busy %.2f ms in %d activities, first at %.2f ms, last ends %.2f ms' % (s, sum(e['dur'] for e in v) / 1e3, len(v), (v[0]['ts'] - t0) / 1e3, (v[-1]['ts'] + v[-1]['dur'] - t0) / 1e3))
M = streams[main]
gaps = [(M[i + 1]['ts'] - M[i]['ts'] - M[i]['dur'], M[i]['name'], M[i + 1]['name'], M[i]['ts'] - t0) for i in range(len(M) - 1)]
tot_gap = sum(g[0] for g in gaps)
print('main stream: idle between activities %.2f ms total; median gap %.2f us; gaps > 10 us: %d (%.2f ms)' % (
    tot_gap / 1e3, sorted(g[0] for g in gaps)[len(gaps) // 2], sum(g[0] > 10 for g in gaps), sum(g[0] for g in gaps if g[0] > 10) / 1e3))
for g in sorted(gaps, key=lambda g: -g[0])[:12]:
    print('   gap %7.1f us at %6.2f ms   after %-48s before %s' % (g[0], g[3] / 1e3, g[1][:48].replace('capb200::(anonymous namespace)::', ''), g[2][:60].replace('capb200::(anonymous namespace)::', '')))


def first(sub, after=0.0):
    for e in M:
        if sub in e['name'] and e['ts'] - t0 >= after:
            return (e['ts'] - t0) / 1e3
    return None


marks = [('first decoder step (embed)', 'embed_relu_dropout'), ('reward (CIDEr-D)', 'cider'), ('criterion / d logits', 'scst_dlogits'), ('first embed backward', 'embed_backward'),
         ('refiner backward (enc_attn_backward)', 'seq_attn_backward'), ('optimizer (adam_kernel)', 'adam_kernel')]
for label, sub in marks:
    print('  %-40s first at %s ms' % (label, first(sub)))
# busiest kernels of the main stream by phase would need names per phase; print 1-ms buckets of busy fraction instead
nb = int(span / 1e3) + 1
busy = [0.0] * nb
for e in M:
    b = int((e['ts'] - t0) / 1e3)
    busy[min(b, nb - 1)] += e['dur']
print('main-stream busy fraction per ms:', ' '.join('%.2f' % (b / 1e3) for b in busy))
if len(sys.argv) > 2:
    json.dump([{'name': e['name'][:80], 'ts': e['ts'] - t0, 'dur': e['dur'], 'stream': e['args'].get('stream')} for e in K], open(sys.argv[2], 'w'))
