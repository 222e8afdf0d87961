"""Host-side post-processing next to the decode path.

    decode_sequence(ix_to_word, seq)        captioning/utils/misc.py:62-84

The reference walks the id tensor with one ``.item()`` (a device synchronisation when ``seq`` lives on the GPU) per token; here the
ids cross to the host once and the per-row work is plain Python on a list.  Same output strings, including the
``REMOVE_BAD_ENDINGS`` environment switch and the BPE ``'@@ '`` merge.
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch

BAD_ENDINGS = ['with', 'in', 'on', 'of', 'a', 'at', 'to', 'for', 'an', 'this', 'his', 'her', 'that', 'the']      # misc.py:17-18


def decode_sequence(ix_to_word: Dict[str, str], seq) -> List[str]:
    """seq: [N, D] integer tensor (or array) with 0 = end token -> one space-joined caption per row."""
    rows = seq.detach().cpu().tolist() if isinstance(seq, torch.Tensor) else [list(map(int, r)) for r in seq]
    strip = int(os.getenv('REMOVE_BAD_ENDINGS', '0'))
    out = []
    for row in rows:
        words = []
        for ix in row:
            if ix <= 0:
                break
            words.append(ix_to_word[str(int(ix))])
        txt = ' '.join(words)
        if strip:
            # misc.py:76-82: drop trailing bad endings; a caption made only of bad endings is kept whole
            parts = txt.split(' ')
            keep = len(parts)
            for j in range(len(parts)):
                if parts[-j - 1] not in BAD_ENDINGS:
                    keep = len(parts) - j
                    break
            txt = ' '.join(parts[:keep])
        out.append(txt.replace('@@ ', ''))
    return out
