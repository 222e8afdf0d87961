"""GPU parity of the Transformer captioner engine (K/V-cached decode) against the goldens of the live reference and the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import LOGP_TOL, PARITY_MODES, build_pair, check_decode, co, first_divergence

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', PARITY_MODES)
def test_transformer_small_golden(golden_dir, mode):
    g = np.load(os.path.join(golden_dir, 'transformer_small.npz'))
    cfg = dict(zip(('V', 'E', 'H', 'A', 'F_fc', 'F_att', 'T'), (int(x) for x in g['cfg'])))
    B, R, b, seed, heads = (int(x) for x in g['meta'])
    model, fam = build_pair('transformer', seed=seed, logit_scale=10.0, mode=mode, heads=heads, **cfg)
    fc, att = co.make_inputs(B, R, cfg['F_fc'], cfg['F_att'], seed=seed)
    fcd, attd = fc.cuda(), att.cuda()
    with torch.no_grad():
        seq, lp = model(fcd, attd, None, opt={'sample_method': 'greedy', 'beam_size': 1}, mode='sample')
        assert np.array_equal(seq.cpu().numpy(), g['greedy_seq'])
        assert np.abs(lp.cpu().numpy() - g['greedy_lp']).max() < LOGP_TOL
        seq, lp = model(fcd, attd, None, opt={'beam_size': b, 'sample_n': 1}, mode='sample')
        assert np.array_equal(seq.cpu().numpy(), g['beam_seq'])
        assert np.abs(lp.cpu().numpy() - g['beam_lp']).max() < LOGP_TOL
        ps = np.array([[model.done_beams[i][j]['p'] for j in range(b)] for i in range(B)])
        assert np.abs(ps - g['done_p']).max() < 1e-3
        masks = torch.from_numpy(g['masks']).cuda()
        seq, lp = model(fcd, attd, masks, opt={'sample_method': 'greedy', 'beam_size': 1}, mode='sample')
        assert np.array_equal(seq.cpu().numpy(), g['masked_greedy_seq'])
        assert np.abs(lp.cpu().numpy() - g['masked_greedy_lp']).max() < LOGP_TOL
        seq, _ = model(fcd, attd, masks, opt={'beam_size': b, 'sample_n': 1}, mode='sample')
        assert np.array_equal(seq.cpu().numpy(), g['masked_beam_seq'])
        out = model(fcd, attd, torch.from_numpy(g['teacher_in']).cuda(), None)
        assert np.abs(out.cpu().numpy() - g['teacher_lp']).max() < LOGP_TOL
        forced = torch.from_numpy(g['sample_seq']).cuda()
        seq, lp = model._sample(fcd, attd, None, opt={'sample_method': 'sample', 'sample_n': 3}, forced_tokens=forced)
        assert np.abs(lp.cpu().numpy() - g['sample_lp']).max() < LOGP_TOL


@pytest.mark.parametrize('mode', PARITY_MODES)
@pytest.mark.parametrize('B,R,beam', [(1, 3, 2), (7, 36, 5), (3, 50, 1)])
def test_transformer_shapes_vs_oracle(mode, B, R, beam):
    """configs/transformer/transformer.yml widths (d_model 512, d_ff 2048, 8 heads) with 2+2 layers so the CPU oracle stays fast."""
    cfg = dict(V=301, E=512, H=2048, A=2, F_fc=64, F_att=2048, T=10)
    model, fam = build_pair('transformer', seed=B + 7, logit_scale=4.0, mode=mode, heads=8, **cfg)
    fc, att = co.make_inputs(B, R, cfg['F_fc'], cfg['F_att'], seed=B + R)
    margins = []
    with torch.no_grad():
        if beam > 1:
            seq, lp = model(fc.cuda(), att.cuda(), None, opt={'beam_size': beam, 'sample_n': 1}, mode='sample')
            oseq, olp, _ = co.sample_beam(fam, fc, att, beam_size=beam, record_margin=margins)
        else:
            seq, lp = model(fc.cuda(), att.cuda(), None, opt={'sample_method': 'greedy', 'beam_size': 1}, mode='sample')
            oseq, olp = co.sample(fam, fc, att, record_margin=margins)
    # chosen-token log-probs within 1e-4; the far tail of the full rows (log-probs down to -40) additionally gets a 1e-5 relative
    # allowance: the tensor-core accumulator truncates once per MMA
    check_decode(fam, fc, att, seq, lp, oseq, olp, margins)


def test_transformer_full_depth_vs_oracle():
    """The reference configuration: 6 + 6 layers, d_model 512, d_ff 2048, 8 heads, V = 9487, 36 regions, T = 20, beam 5."""
    cfg = dict(V=9487, E=512, H=2048, A=6, F_fc=64, F_att=2048, T=20)
    model, fam = build_pair('transformer', seed=1234, logit_scale=3.0, mode='tc_f16x3', heads=8, **cfg)
    fc, att = co.make_inputs(3, 36, cfg['F_fc'], cfg['F_att'], seed=1234)
    margins = []
    with torch.no_grad():
        seq, lp = model(fc.cuda(), att.cuda(), None, opt={'beam_size': 5, 'sample_n': 1}, mode='sample')
        oseq, olp, _ = co.sample_beam(fam, fc, att, beam_size=5, record_margin=margins)
    check_decode(fam, fc, att, seq, lp, oseq, olp, margins)
