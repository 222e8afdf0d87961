"""CPU-only checks of the boundary: the shared library loads, exports every symbol include/capb200.h declares, refuses
to run without a GPU (no CPU fallback), and the host-side mirrors expose the reference's parameter names / shapes."""
import json
import os
import re

import pytest
import torch

from helpers import REPO, family_opt, make_opt


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    import imagecaptioning.pytorch_b200 as b200
    return b200._lib.load()


def test_header_symbols_exported(lib):
    import imagecaptioning.pytorch_b200 as b200
    header = open(os.path.join(REPO, 'include', 'capb200.h')).read()
    declared = set(re.findall(r'\b(capb200_[a-z0-9_]+)\s*\(', header))
    declared -= {'capb200_engine', 'capb200_cider_table'}
    assert len(declared) >= 18
    assert declared == set(b200._lib.SIGNATURES), declared ^ set(b200._lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.capb200_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour of a box without a GPU')
def test_no_cpu_fallback(lib):
    import ctypes
    import imagecaptioning.pytorch_b200 as b200
    cfg = b200._lib.ModelCfg(0, 60, 32, 32, 16, 48, 48, 8, 1)
    assert not lib.capb200_engine_create(ctypes.byref(cfg))
    assert b'no CUDA device' in lib.capb200_last_error()
    model = b200.setup(make_opt('updown', 60, 32, 32, 16, 48, 48, 8))
    with pytest.raises(RuntimeError, match='CUDA'):
        model(torch.zeros(2, 48), torch.zeros(2, 3, 48), None, opt={'beam_size': 1}, mode='sample')


def test_state_dict_keys_match_reference(golden_dir):
    import imagecaptioning.pytorch_b200 as b200
    g = json.load(open(os.path.join(golden_dir, 'state_dict_keys.json')))
    c = g['cfg']
    dims = {'updown': (60, 32, 40, 16), 'newfc': (60, 32, 40, 16), 'transformer': (60, 32, 64, 2), 'aoa': (60, 32, 32, 16)}   # V, E, H, A
    for fam, ref in g['keys'].items():
        V, E, H, A = dims[fam]
        try:
            m = b200.setup(family_opt(fam, V, E, H, A, 48, 56, 8, heads=4))
        except NotImplementedError:
            pytest.skip('%s mirror not built yet' % fam) if fam == 'aoa' else pytest.fail(fam)
            continue
        mine = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert mine == ref, (fam, set(mine) ^ set(ref))


def test_unsupported_options_raise():
    import imagecaptioning.pytorch_b200 as b200
    m = b200.setup(make_opt('updown', 60, 32, 32, 16, 48, 48, 8))
    fc, att = torch.zeros(2, 48), torch.zeros(2, 3, 48)
    for bad in ({'group_size': 2, 'beam_size': 2}, {'output_logsoftmax': 0}, {'sample_method': 'dbs'}):
        with pytest.raises(NotImplementedError):
            m(fc, att, None, opt=bad, mode='sample')
    # options the engine implements get past the guards and stop at the no-CPU-fallback check
    for ok in ({'block_trigrams': 1}, {'sample_method': 'top5'}, {'sample_method': 'top0.9'}, {'sample_method': 'gumbel'}, {'decoding_constraint': 1},
               {'remove_bad_endings': 1, 'beam_size': 2, 'sample_n': 1}, {'suppress_UNK': 1, 'beam_size': 2, 'sample_n': 1, 'temperature': 0.7}):
        with pytest.raises(RuntimeError, match='CUDA'):
            m(fc, att, None, opt=ok, mode='sample')
    with pytest.raises(NotImplementedError):
        b200.setup(make_opt('adaatt', 60, 32, 32, 16, 48, 48, 8))


def test_pack_references_layout():
    import numpy as np
    from oracle import ciderd_oracle as cdo
    import imagecaptioning.pytorch_b200 as b200
    gts = cdo.make_refs(3, 20, n_refs=4, L=16, seed=1) + [np.zeros((2, 7), dtype=np.int64)]
    refs, offs, L = b200.rewards.pack_references(gts, 'cpu')
    assert L == 16 and refs.shape == (14, 16) and offs.tolist() == [0, 4, 8, 12, 14]
    assert refs[:4].tolist() == gts[0].tolist()
