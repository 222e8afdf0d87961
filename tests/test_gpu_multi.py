"""Two-GPU, one-process test of the nn.DataParallel calling convention of tools/train.py:86-88,185 (one host thread per GPU calling the
replicas concurrently): each device gets its own engine, kept across forwards; decode parity and the SCST loss/backward through the
replicas.  Skipped on single-GPU boxes."""
import argparse

import numpy as np
import pytest
import torch

from helpers import build_pair, check_decode, co

pytestmark = pytest.mark.gpu

CFG = dict(V=40, E=32, H=48, A=24, F_fc=32, F_att=40, T=9)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_data_parallel_decode_and_scst():
    import imagecaptioning.pytorch_b200 as b200
    from oracle import ciderd_oracle as cdo
    model, fam = build_pair('updown', seed=31, logit_scale=8.0, mode='tc_f16x3', device='cuda:0', **CFG)
    dp = torch.nn.DataParallel(model, device_ids=[0, 1])
    B, R, b = 6, 7, 3
    fc, att = co.make_inputs(B, R, CFG['F_fc'], CFG['F_att'], seed=5)
    handles = None
    for call in range(3):                               # eager, graph capture, graph replay -- on both devices
        with torch.no_grad():
            seq, lp = dp(fc.cuda(0), att.cuda(0), None, opt={'beam_size': b, 'sample_n': 1}, mode='sample')
        assert seq.shape == (B, CFG['T']) and seq.device.index == 0
        margins = []
        oseq, olp, _ = co.sample_beam(fam, fc, att, beam_size=b, record_margin=margins)
        check_decode(fam, fc, att, seq, lp, oseq, olp, margins)
        now = {d: s['engine'] for d, s in model._store.slots.items() if s['engine'] is not None}
        assert set(now) == {0, 1}
        assert handles is None or handles == now        # the replicas reuse the per-device engines
        handles = now
    # SCST through DataParallel(LossWrapper), as tools/train.py does
    gts = cdo.make_refs(B, CFG['V'], seed=3)
    df, ref_len = cdo.build_document_frequency(cdo.make_refs(100, CFG['V'], seed=4))
    b200.rewards.reset_scorer()
    b200.rewards.init_scorer(b200.rewards.CiderDTable(df, ref_len, device='cuda:0'))
    opt = argparse.Namespace(sc_sample_method='greedy', sc_beam_size=1, train_sample_method='sample', train_beam_size=1, train_sample_n=3,
                             cider_reward_weight=1, bleu_reward_weight=0, label_smoothing=0.0)
    lw = b200.B200LossWrapper(model, opt)
    dp_lw = torch.nn.DataParallel(lw, device_ids=[0, 1])       # the scorer builds its table on each replica's device on first use
    model.train()
    out = dp_lw(fc.cuda(0), att.cuda(0), None, None, None, gts, torch.arange(B), True, False, False)
    loss = out['loss'].mean()
    model.zero_grad()
    loss.backward()
    grads = [p.grad for p in model.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert sum(float(g.abs().max()) > 0 for g in grads) >= 15
    b200.rewards.reset_scorer()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('family', ['aoa', 'transformer'])
def test_data_parallel_replicas_see_updated_weights(family):
    """nn.DataParallel replicas are fresh Broadcast outputs every forward (version 0) and the caching allocator hands the same addresses back
    after an optimizer step, so (data_ptr, _version) cannot tell a replica's weights changed: the engines on GPU >= 1 must re-bind on every
    call.  decode -> in-place weight update (what optimizer.step does) -> decode: both GPUs must follow the new weights (checked per shard
    against the oracle built from the updated state dict)."""
    heads = 4
    cfg = dict(V=40, E=32, H=64, A=0, F_fc=32, F_att=40, T=7) if family == 'aoa' else dict(V=40, E=32, H=64, A=2, F_fc=32, F_att=40, T=7)
    model, fam = build_pair(family, seed=21, logit_scale=8.0, mode='tc_f16x3', device='cuda:0', heads=heads, **cfg)
    dp = torch.nn.DataParallel(model, device_ids=[0, 1])
    B, R = 6, 7
    fc, att = co.make_inputs(B, R, cfg['F_fc'], cfg['F_att'], seed=5)
    opt = {'sample_method': 'greedy', 'beam_size': 1}
    with torch.no_grad():
        seq0, _ = dp(fc.cuda(0), att.cuda(0), None, opt=opt, mode='sample')
        g = torch.Generator(device='cuda:0').manual_seed(1)
        for p in model.parameters():                        # an "optimizer step": in-place update of the master parameters
            p.add_(torch.randn(p.shape, generator=g, device='cuda:0') * 0.05 * p.abs().mean())
        seq1, lp1 = dp(fc.cuda(0), att.cuda(0), None, opt=opt, mode='sample')
    W1 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    fam1 = co.Family(family, W1, cfg['T'], heads=heads)
    margins = []
    oseq, olp = co.sample(fam1, fc, att, record_margin=margins)
    check_decode(fam1, fc, att, seq1, lp1, oseq, olp, margins)          # covers the second half of the batch = the GPU-1 replica
    assert not torch.equal(seq0.cpu(), seq1.cpu())                       # the update did change the captions


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('family', ['aoa', 'updown', 'transformer'])
def test_overlapped_gradient_sync_two_ranks(family):
    """One process per GPU (torchrun, NCCL): see tests/gpu_sync_check.py."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    port = 29600 + {'aoa': 1, 'updown': 2, 'transformer': 3}[family]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(here, 'gpu_sync_check.py'), family]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=os.path.dirname(here))
    assert r.returncode == 0 and 'SYNC-OK' in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
