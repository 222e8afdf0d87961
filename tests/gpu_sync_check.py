"""Two-rank check of the overlapped gradient all-reduce of the fused SCST step (run under torchrun by tests/test_gpu_multi.py).

Each rank has its own images and references.  The synchronised gradients (B200LossWrapper.enable_gradient_sync: chunks all-reduced on a
communication stream as the engine finishes each gradient group -- recorded by external event nodes when the step is replayed as a CUDA graph)
must equal the average over the ranks of what an unsynchronised engine computes for the same (weights, inputs, seed), on the eager first
step, the captured second step and the replayed later steps alike.  An all-reduce that started before a group was complete would average
stale data and fail the comparison."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from helpers import build_pair, co                                      # noqa: E402
from oracle import ciderd_oracle as cdo                                 # noqa: E402


def main():
    import imagecaptioning.pytorch_b200 as b200
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl')
    family = sys.argv[1] if len(sys.argv) > 1 else 'aoa'
    cfg = {'aoa': dict(V=40, E=32, H=64, A=0, F_fc=32, F_att=40, T=7), 'updown': dict(V=40, E=32, H=48, A=24, F_fc=32, F_att=40, T=9),
           'transformer': dict(V=40, E=32, H=64, A=2, F_fc=32, F_att=40, T=7)}[family]
    B, R, n = 3, 9, 3

    def fresh():
        m, _ = build_pair(family, seed=27, logit_scale=5.0, mode='tc_f16x3', heads=4, **cfg)
        m.train()
        return m
    model, ref = fresh(), fresh()
    fc, att = co.make_inputs(B, R, cfg['F_fc'], cfg['F_att'], seed=40 + rank)
    gts = cdo.make_refs(B, cfg['V'], seed=2 + rank)
    df, ref_len = cdo.build_document_frequency(cdo.make_refs(200, cfg['V'], seed=4))
    table = b200.rewards.CiderDTable(df, ref_len)
    b200.rewards.reset_scorer()
    b200.rewards.init_scorer(table)
    opt = argparse.Namespace(sc_sample_method='greedy', sc_beam_size=1, train_sample_method='sample', train_beam_size=1, train_sample_n=n, cider_reward_weight=1,
                             bleu_reward_weight=0)
    lw = b200.B200LossWrapper(model, opt).enable_gradient_sync()
    worst = 0.0
    for step, seed in enumerate([7, 8, 7, 9, 7]):
        torch.manual_seed(seed)                                        # the wrapper draws the step seed from torch's generator
        out = lw(fc.cuda(), att.cuda(), None, None, None, gts, torch.arange(B), True, False, False)
        used = lw.last_step['seed']
        model.zero_grad(set_to_none=True)
        out['loss'].backward()
        torch.cuda.synchronize()
        got = lw.last_step['flat'].flat.clone()
        local = ref.scst_step(fc.cuda(), att.cuda(), gts, table, n, seed=used)['flat'].flat.clone()      # unsynchronised engine, same seed
        dist.all_reduce(local)
        local /= world
        err = float((got - local).abs().max()) / float(local.abs().max())
        worst = max(worst, err)
        assert err < 1e-5, 'rank %d step %d: synchronised gradients differ from the average of the local ones (%.3g)' % (rank, step, err)
        assert torch.equal(model.logit.weight.grad if hasattr(model, 'logit') else model.model.generator.proj.weight.grad,
                           lw.last_step['grads'][model.logit.weight if hasattr(model, 'logit') else model.model.generator.proj.weight])
    dist.barrier()
    if rank == 0:
        print('SYNC-OK %s worst relative difference %.2e over 5 steps (eager, captured, replayed)' % (family, worst))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
