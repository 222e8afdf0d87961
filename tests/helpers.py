"""Shared helpers for the parity tests."""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle import caption_oracle as co   # noqa: E402  (tests are allowed to use the oracle as the checker)

LOGP_TOL = 1e-4      # BASELINE.json north_star: log-probs within 1e-4 (fp32)
PARITY_MODES = ['simt_fp32', 'tc_f16x3']


def make_opt(family, V, E, H, A, F_fc, F_att, T):
    return argparse.Namespace(vocab_size=V, input_encoding_size=E, rnn_size=H, num_layers=1, drop_prob_lm=0.5, max_length=T, seq_length=T,
                              fc_feat_size=F_fc, att_feat_size=F_att, att_hid_size=A, vocab={str(i): 'w%d' % i for i in range(1, V + 1)},
                              caption_model=family, use_bn=0, logit_layers=1)


def build_pair(family, V, E, H, A, F_fc, F_att, T, seed, logit_scale, mode, device='cuda'):
    """Returns (B200 model on the GPU, oracle Family on the CPU) sharing the same synthetic weights."""
    import imagecaptioning.pytorch_b200 as b200
    W = co.make_weights(family, V, E, H, A, F_fc, F_att, seed=seed, logit_scale=logit_scale)
    model = b200.setup(make_opt(family, V, E, H, A, F_fc, F_att, T), numeric_mode=mode)
    model.load_state_dict(W, strict=True)
    model = model.to(device).eval()
    return model, co.Family(family, W, T)


def first_divergence(a, b):
    """Index of the first column where two id matrices differ, per row (-1 = identical)."""
    a, b = np.asarray(a), np.asarray(b)
    out = np.full(a.shape[0], -1)
    for i in range(a.shape[0]):
        d = np.nonzero(a[i] != b[i])[0]
        if d.size:
            out[i] = d[0]
    return out
