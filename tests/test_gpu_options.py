"""GPU parity of the decode options (SURVEY.md 8f "later"): decoding_constraint, block_trigrams, remove_bad_endings, suppress_UNK / unk_idx and
temperature in beam search, against outputs of the live reference (tests/golden/updown_options.npz, oracle/make_golden.py: gen_updown_options);
top-k / nucleus / gumbel sampling against the distribution they must draw from."""
import os

import numpy as np
import pytest
import torch

from helpers import LOGP_TOL, build_pair, co

pytestmark = pytest.mark.gpu

BAD = ('the', 'a', 'with')


def _setup(golden_dir, mode='tc_f16x3'):
    g = np.load(os.path.join(golden_dir, 'updown_options.npz'))
    cfg = dict(zip(('V', 'E', 'H', 'A', 'F_fc', 'F_att', 'T'), (int(x) for x in g['cfg'])))
    B, R, b, seed = (int(x) for x in g['meta'])
    model, fam = build_pair('updown', seed=seed, logit_scale=8.0, mode=mode, **cfg)
    vocab = {str(i): 'w%d' % i for i in range(1, cfg['V'] + 1)}
    vocab[str(cfg['V'])] = 'UNK'
    for w, name in zip(g['bad_words'].tolist(), BAD):
        vocab[str(w)] = name
    model.vocab = vocab
    model.bad_endings_ix = [int(k) for k, v in vocab.items() if v in BAD]
    fc, att = co.make_inputs(B, R, cfg['F_fc'], cfg['F_att'], seed=seed)
    return g, cfg, model, fc.cuda(), att.cuda(), b


def _same_rows(mine, ref, tol=LOGP_TOL):
    """Log-prob rows with -inf entries (decoding_constraint / bad endings) and -1000 offsets: same -inf pattern, finite values within tol."""
    mine, ref = np.asarray(mine), np.asarray(ref)
    ok = ~np.isnan(ref)                       # the reference multiplies finished rows by 0: 0 * -inf = nan where we store 0
    assert np.array_equal(np.isinf(mine) & ok, np.isinf(ref) & ok)
    fin = np.isfinite(ref) & ok
    assert np.abs(mine[fin] - ref[fin]).max() < tol + 1e-4 * 0            # -1000-shifted entries carry an absolute error of ~1e-4 in fp32
    assert np.all(mine[~ok] == 0)


@pytest.mark.parametrize('mode', ['tc_f16x3', 'simt_fp32'])
def test_greedy_options_golden(golden_dir, mode):
    g, cfg, model, fc, att, b = _setup(golden_dir, mode)
    with torch.no_grad():
        for tag, opt in (('g_plain', {}), ('g_con', {'decoding_constraint': 1}), ('g_tri', {'block_trigrams': 1}),
                         ('g_all', {'decoding_constraint': 1, 'block_trigrams': 1})):
            seq, lp = model(fc, att, None, opt=dict({'sample_method': 'greedy', 'beam_size': 1}, **opt), mode='sample')
            assert np.array_equal(seq.cpu().numpy(), g[tag + '_seq']), tag
            _same_rows(lp.cpu().numpy(), g[tag + '_lp'])
        # forced replay of the reference's sampled run with sample_n = 2: trigram blocking touches the first batch_size rows only
        forced = torch.from_numpy(g['s_tri_seq']).cuda()
        seq, lp = model._sample(fc, att, None, opt={'sample_method': 'sample', 'sample_n': 2, 'block_trigrams': 1, 'decoding_constraint': 1}, forced_tokens=forced)
        _same_rows(lp.cpu().numpy(), g['s_tri_lp'])


def test_remove_bad_endings_in_sample(golden_dir):
    """The reference's own _sample path for this option does not run on current torch (uint8 mask, AttModel.py:303), so the intended
    semantics are checked directly: after a bad-ending word the end token has log-prob -inf and is never chosen."""
    g, cfg, model, fc, att, b = _setup(golden_dir)
    with torch.no_grad():
        seq, lp = model(fc, att, None, opt={'sample_method': 'greedy', 'beam_size': 1, 'remove_bad_endings': 1}, mode='sample')
        ref_seq, ref_lp = model(fc, att, None, opt={'sample_method': 'greedy', 'beam_size': 1}, mode='sample')
    seq, lp, ref_lp = seq.cpu().numpy(), lp.cpu().numpy(), ref_lp.cpu().numpy()
    bad = set(model.bad_endings_ix)
    hits = 0
    for i in range(seq.shape[0]):
        for t in range(1, seq.shape[1]):
            if seq[i, t - 1] in bad:
                hits += 1
                assert np.isneginf(lp[i, t, 0]) and seq[i, t] != 0
            elif seq[i, t - 1] != 0:
                assert np.isfinite(lp[i, t, 0])
    assert hits > 0


@pytest.mark.parametrize('tag,opt', [('b_unk', {'suppress_UNK': 1}), ('b_temp', {'temperature': 0.7}), ('b_con', {'decoding_constraint': 1}),
                                     ('b_bad', {'remove_bad_endings': 1}),
                                     ('b_all', {'suppress_UNK': 1, 'decoding_constraint': 1, 'remove_bad_endings': 1, 'temperature': 1.3})])
def test_beam_options_golden(golden_dir, tag, opt):
    g, cfg, model, fc, att, b = _setup(golden_dir)
    with torch.no_grad():
        for _ in range(3):                                  # eager, graph capture, graph replay
            seq, lp = model(fc, att, None, opt=dict({'beam_size': b, 'sample_n': 1}, **opt), mode='sample')
    assert np.array_equal(seq.cpu().numpy(), g[tag + '_seq']), tag
    _same_rows(lp.cpu().numpy(), g[tag + '_lp'], tol=2e-4)
    for i in range(seq.shape[0]):
        for j in range(b):
            rec = model.done_beams[i][j]
            Lr = int(g[tag + '_done_len'][i, j])
            assert rec['seq'].cpu().tolist() == g[tag + '_done_seq'][i, j, :Lr].tolist()
            assert abs(rec['p'] - g[tag + '_done_p'][i, j]) < 1e-3
    # the lazily materialised rows of a finished beam carry the same edits
    rec = model.done_beams[1][0]
    _same_rows(rec['logps'].cpu().numpy(), g[tag + '_lp'][1, :rec['seq'].shape[0]], tol=2e-4)


def test_unk_idx_is_lowered_regardless_of_the_flag(golden_dir):
    """CaptionModel.py:161-162: with unk_idx set the column is lowered by 1000 at every step whatever suppress_UNK says."""
    g, cfg, model, fc, att, b = _setup(golden_dir)
    model.vocab = {k: ('w' + k) for k in model.vocab}       # no 'UNK' word: the elif branch applies
    model.unk_idx = 7
    with torch.no_grad():
        seq, lp = model(fc, att, None, opt={'beam_size': b, 'sample_n': 1}, mode='sample')
        model.unk_idx = None
        seq0, lp0 = model(fc, att, None, opt={'beam_size': b, 'sample_n': 1}, mode='sample')
    lp, lp0 = lp.cpu().numpy(), lp0.cpu().numpy()
    same = (seq.cpu().numpy() == seq0.cpu().numpy()).all(1)
    assert same.any() and not (seq.cpu().numpy() == 7).any()
    live = (np.abs(lp0[same]).sum(2) > 0)
    assert np.abs((lp[same][..., 7] - lp0[same][..., 7])[live] + 1000).max() < 1e-3


@pytest.mark.parametrize('method', ['top5', 'top0.6', 'gumbel'])
def test_truncated_samplers_draw_from_the_right_distribution(method):
    """First-step tokens over many rows of one image against the distribution the reference's sample_next_word defines
    (CaptionModel.py:375-406): top-k / nucleus renormalise the kept words of softmax(logp / T); gumbel is a plain multinomial draw."""
    cfg = dict(V=30, E=16, H=24, A=8, F_fc=16, F_att=16, T=6)
    model, fam = build_pair('updown', seed=3, logit_scale=3.0, mode='simt_fp32', **cfg)
    fc, att = co.make_inputs(1, 4, 16, 16, seed=3)
    n, temperature = 4000, 1.3
    torch.manual_seed(0)
    with torch.no_grad():
        seq, lp = model(fc.cuda(), att.cuda(), None, opt={'sample_method': method, 'sample_n': n, 'temperature': temperature}, mode='sample')
    _, olp = co.sample(fam, fc, att)
    logits = olp[0, 0] / (1.0 if method == 'gumbel' else temperature)
    p = torch.softmax(logits, 0)
    if method == 'top5':
        keep = torch.zeros_like(p, dtype=torch.bool)
        keep[p.topk(5).indices] = True
    elif method == 'top0.6':
        sp, si = p.sort(descending=True)
        m = sp.cumsum(0) < 0.6
        m = torch.cat([torch.ones(1, dtype=torch.bool), m[:-1]])
        keep = torch.zeros_like(p, dtype=torch.bool)
        keep[si[m]] = True
    else:
        keep = torch.ones_like(p, dtype=torch.bool)
    q = (p * keep) / (p * keep).sum()
    counts = np.bincount(seq[:, 0].cpu().numpy(), minlength=31).astype(np.float64)
    assert counts[~keep.numpy()].sum() == 0                     # nothing outside the kept set is ever drawn
    q = q.numpy()
    big = q * n > 5
    chi2 = float((((counts - q * n) ** 2) / np.maximum(q * n, 1e-12))[big].sum())
    dof = max(int(big.sum()) - 1, 1)
    assert chi2 < dof + 6 * (2 * dof) ** 0.5, (method, chi2, dof)
    assert float((lp[:, 0].cpu() - olp[0, 0]).abs().max()) < LOGP_TOL        # stored rows stay the un-tempered, un-truncated log-probs
