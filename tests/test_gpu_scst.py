"""GPU parity of the SCST training step (UpDown): loss, reward and every parameter gradient against torch autograd through the
oracle on the CPU, with the engine's own samples and (when dropout is on) its own dropout masks replayed in the oracle."""
import argparse

import numpy as np
import pytest
import torch

from helpers import LOGP_TOL, build_pair, co

pytestmark = pytest.mark.gpu

CFG = dict(V=40, E=32, H=48, A=24, F_fc=32, F_att=40, T=9)


def _oracle_grads(W, fc, att, gts, df, ref_len, sample_seq, greedy_seq, n, drop):
    from oracle import ciderd_oracle as cdo
    Wg = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    fam = co.Family('updown', Wg, CFG['T'])
    fam.drop = drop
    _, lp = co.sample(fam, fc, att, sample_method='sample', sample_n=n, forced_tokens=sample_seq)
    reward, _ = cdo.self_critical_reward(greedy_seq.numpy(), gts, sample_seq.numpy(), df, ref_len)
    reward = torch.from_numpy(reward).float()
    loss = co.reward_criterion(lp, sample_seq, reward)
    loss.backward()
    return float(loss), reward, {k: v.grad for k, v in Wg.items()}, lp.detach()


@pytest.mark.parametrize('drop_prob', [0.0, 0.5])
@pytest.mark.parametrize('mode', ['tc_f16x3', 'simt_fp32'])
def test_scst_step_gradients(mode, drop_prob):
    import imagecaptioning.pytorch_b200 as b200
    from oracle import ciderd_oracle as cdo
    model, fam = build_pair('updown', seed=31, logit_scale=5.0, mode=mode, **CFG)
    W = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    B, R, n, T = 5, 11, 4, CFG['T']
    fc, att = co.make_inputs(B, R, CFG['F_fc'], CFG['F_att'], seed=4)
    gts = cdo.make_refs(B, CFG['V'], seed=2)
    df, ref_len = cdo.build_document_frequency(cdo.make_refs(200, CFG['V'], seed=4))
    table = b200.rewards.CiderDTable(df, ref_len)
    model.train()
    res = model.scst_step(fc.cuda(), att.cuda(), gts, table, n, temperature=1.0, drop_prob=drop_prob, seed=1234)
    torch.cuda.synchronize()
    sample_seq, greedy_seq = res['sample_seq'].cpu(), res['greedy_seq'].cpu()
    # greedy baseline = eval-mode greedy decode of the oracle
    og, _ = co.sample(fam, fc, att)
    assert torch.equal(greedy_seq, og)
    drop = None
    if drop_prob > 0:
        L, lib = b200._lib, b200._lib.load()
        N, E, H = B * n, CFG['E'], CFG['H']

        def mask(site, step, rows, cols):
            m = torch.empty(rows * cols, device='cuda')
            L.check(lib.capb200_dropout_mask(L.ptr(m), rows * cols, 1234, site, step, drop_prob, L.current_stream()), 'dropout_mask')
            return m.cpu().reshape(rows, cols)
        drop = {'fc': mask(0, 0, B, H), 'att': mask(1, 0, B * R, H).reshape(B, R, H),
                'xt': torch.stack([mask(2, t, N, E) for t in range(T)]), 'out': torch.stack([mask(3, t, N, H) for t in range(T)])}
        keep = float((drop['att'] > 0).float().mean())
        assert abs(keep - (1 - drop_prob)) < 0.03 and abs(float(drop['att'].max()) - 1 / (1 - drop_prob)) < 1e-6
    oloss, oreward, ograds, olp = _oracle_grads(W, fc, att, gts, df, ref_len, sample_seq, greedy_seq, n, drop)
    assert float((res['sample_logprobs'].cpu() - olp).abs().max()) < LOGP_TOL
    assert float((res['reward'].cpu() - oreward).abs().max()) < LOGP_TOL
    assert abs(float(res['loss']) - oloss) < LOGP_TOL
    name_of = {id(p): k for k, p in model.state_dict(keep_vars=True).items()}
    worst = 0.0
    for p, g in res['grads'].items():
        key = name_of[id(p)]
        ref = ograds[key]
        scale = float(ref.abs().max())
        err = float((g.cpu() - ref).abs().max())
        if scale > 1e-7:
            worst = max(worst, err / scale)
        assert err <= 5e-4 * scale + 2e-9, (key, err, scale)      # 5e-4 of the tensor's largest gradient entry
    assert abs(oloss) > 1e-2 and sum(float(v.abs().max()) > 1e-5 for v in ograds.values()) >= 15      # the comparison is not vacuous
    print('max relative gradient error', worst)


def test_loss_wrapper_backward_sets_param_grads():
    import imagecaptioning.pytorch_b200 as b200
    from oracle import ciderd_oracle as cdo
    model, _ = build_pair('updown', seed=31, logit_scale=5.0, mode='tc_f16x3', **CFG)
    B, R, n = 4, 7, 3
    fc, att = co.make_inputs(B, R, CFG['F_fc'], CFG['F_att'], seed=5)
    gts = cdo.make_refs(B, CFG['V'], seed=3)
    df, ref_len = cdo.build_document_frequency(cdo.make_refs(100, CFG['V'], seed=4))
    b200.rewards.reset_scorer()
    b200.rewards.init_scorer(b200.rewards.CiderDTable(df, ref_len))
    opt = argparse.Namespace(sc_sample_method='greedy', sc_beam_size=1, train_sample_method='sample', train_beam_size=1, train_sample_n=n,
                             cider_reward_weight=1, bleu_reward_weight=0)
    lw = b200.B200LossWrapper(model, opt)
    out = lw(fc.cuda(), att.cuda(), None, None, None, gts, torch.arange(B), True, False, False)
    assert out['loss'].requires_grad
    (2.0 * out['loss']).backward()
    step = lw.last_step
    for p, g in step['grads'].items():
        assert p.grad is not None and torch.allclose(p.grad, 2.0 * g)
    # an optimizer step changes the weights, the next call re-binds them (version counters) and still works
    torch.optim.SGD(model.parameters(), lr=1e-3).step()
    out2 = lw(fc.cuda(), att.cuda(), None, None, None, gts, torch.arange(B), True, False, False)
    assert torch.isfinite(out2['loss'])
    b200.rewards.reset_scorer()
