"""GPU parity of the AoANet engine against the goldens of the live reference and the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import LOGP_TOL, PARITY_MODES, build_pair, check_decode, co, first_divergence

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', PARITY_MODES)
def test_aoa_small_golden(golden_dir, mode):
    g = np.load(os.path.join(golden_dir, 'aoa_small.npz'))
    cfg = dict(zip(('V', 'E', 'H', 'A', 'F_fc', 'F_att', 'T'), (int(x) for x in g['cfg'])))
    B, R, b, seed, heads = (int(x) for x in g['meta'])
    model, fam = build_pair('aoa', seed=seed, logit_scale=20.0, mode=mode, heads=heads, **cfg)
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
@pytest.mark.parametrize('B,R,n,beam', [(10, 36, 5, 1), (2, 61, 1, 3)])
def test_aoa_config_dims_vs_oracle(mode, B, R, n, beam):
    """configs/aoa.yml widths (E = H = 1024, 8 heads, 6 refiner layers) at the SCST shape (10 images x 5 samples) and a beam case."""
    cfg = dict(V=501, E=1024, H=1024, A=0, F_fc=16, F_att=2048, T=8)
    model, fam = build_pair('aoa', seed=6, logit_scale=6.0, mode=mode, heads=8, **cfg)
    fc, att = co.make_inputs(B, R, cfg['F_fc'], cfg['F_att'], seed=B + R)
    margins = []
    with torch.no_grad():
        if beam > 1:
            seq, lp = model(fc.cuda(), att.cuda(), None, opt={'beam_size': beam, 'sample_n': 1}, mode='sample')
            oseq, olp, _ = co.sample_beam(fam, fc, att, beam_size=beam, record_margin=margins)
        else:
            seq, lp = model(fc.cuda(), att.cuda(), None, opt={'sample_method': 'greedy', 'beam_size': 1, 'sample_n': n}, mode='sample')
            oseq, olp = co.sample(fam, fc, att, sample_n=n, record_margin=margins)
    strict = check_decode(fam, fc, att, seq, lp, oseq, olp, margins, sample_n=n)
    if beam == 1:
        assert strict, 'the SCST-shape greedy case is seeded to have unambiguous decisions'
