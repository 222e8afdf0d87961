"""GPU: wall-clock phases of one SCST training step as bench.py runs it (weights change every step)."""
import argparse, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import imagecaptioning.pytorch_b200 as b200
from imagecaptioning.pytorch_b200 import synthetic as syn       # seeded synthetic weights / inputs (profiling tools never touch oracle/)
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10
FAM = sys.argv[2] if len(sys.argv) > 2 else 'updown'
import torch.distributed as dist
rank, world, lrank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lrank)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lrank))
if FAM == 'aoa':
    model = syn.build_model('aoa', seed=1234, logit_scale=6.0, mode='tc_f16x3', heads=8, device=torch.device('cuda', lrank), **dict(bench.CFG, E=1024, H=1024, A=0))
else:
    model = syn.build_model('updown', seed=1234, logit_scale=12.0, mode='tc_f16x3', device=torch.device('cuda', lrank), **bench.CFG)
model.train()
df, ref_len = syn.document_frequency(syn.make_refs(500, 9487, seed=4))
b200.rewards.reset_scorer(); b200.rewards.init_scorer(b200.rewards.CiderDTable(df, ref_len))
opt = argparse.Namespace(sc_sample_method='greedy', sc_beam_size=1, train_sample_method='sample', train_beam_size=1, train_sample_n=5,
                         cider_reward_weight=1, bleu_reward_weight=0)
lw = b200.B200LossWrapper(model, opt)
optim = torch.optim.Adam(model.parameters(), lr=5e-5)
fc, att = syn.make_inputs(B, 36, 2048, 2048, seed=1)
fc, att = fc.cuda(), att.cuda()
gts = syn.make_refs(B, 9487, seed=5)
idx = torch.arange(B)
acc = {}
def tick(name, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1
for it in range(6):
    if it == 2:
        acc.clear()
    torch.cuda.synchronize(); t = time.perf_counter()
    out = lw(fc, att, None, None, None, gts, idx, True, False, False); t = tick('loss_wrapper (greedy + sample + reward + BPTT)', t)
    optim.zero_grad(set_to_none=True); t = tick('zero_grad', t)
    out['loss'].backward(); t = tick('autograd bridge backward', t)
    b200.parallel.allreduce_gradients(model.parameters()); t = tick('allreduce_gradients', t)
    torch.nn.utils.clip_grad_value_(model.parameters(), 0.1); t = tick('clip_grad_value_', t)
    optim.step(); t = tick('optimizer step', t)
for k, v in acc.items():
    print('rank %d  %-55s %8.2f ms/step' % (rank, k, v / 4 * 1e3), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
# inside the wrapper: engine rebind vs the C call
import cProfile, pstats
pr = cProfile.Profile()
optim.step()
pr.enable()
out = lw(fc, att, None, None, None, gts, idx, True, False, False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
