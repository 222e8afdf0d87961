"""Recipe that makes the UNMODIFIED reference available as the CPU arm on the GPU box (test infrastructure, not product).

    python oracle/build_ref.py            # needs /root/reference (build container); writes oracle/_ref/

The reference is pure Python, so "building" it is a verbatim copy of the modules the hot path imports:

    /root/reference/captioning/                         -> oracle/_ref/captioning/
    /root/reference/cider/pyciderevalcap/{cider,ciderD} -> oracle/_ref/cider/pyciderevalcap/...
    /root/reference/coco-caption/pycocoevalcap/bleu     -> oracle/_ref/coco-caption/pycocoevalcap/bleu   (rewards.py:15-16 imports it)

``oracle/_ref/`` is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU box
with the snapshot exactly like the built ``libcapb200.so``.  ``oracle/ref_runtime.py`` imports the copy from a scratch working
directory (captioning/utils/rewards.py:12,15 and cider/pyciderevalcap/ciderD/ciderD_scorer.py:109 use cwd-relative paths).
Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may touch it.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
DST = os.path.join(HERE, '_ref')

COPIES = [
    ('captioning', 'captioning'),
    ('cider/pyciderevalcap/__init__.py', 'cider/pyciderevalcap/__init__.py'),
    ('cider/pyciderevalcap/cider', 'cider/pyciderevalcap/cider'),
    ('cider/pyciderevalcap/ciderD', 'cider/pyciderevalcap/ciderD'),
    ('coco-caption/pycocoevalcap/__init__.py', 'coco-caption/pycocoevalcap/__init__.py'),
    ('coco-caption/pycocoevalcap/bleu', 'coco-caption/pycocoevalcap/bleu'),
]


def build(force: bool = False, verbose: bool = True) -> str | None:
    """Copies the reference modules into oracle/_ref/ (no edits).  Returns the destination, or None when /root/reference is absent
    (the GPU box: the copy made in the build container is used as is)."""
    if not os.path.isdir(REF):
        return DST if os.path.isdir(os.path.join(DST, 'captioning')) else None
    stamp = os.path.join(DST, '.complete')
    if os.path.exists(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    ignore = shutil.ignore_patterns('__pycache__', '*.pyc')
    for src, dst in COPIES:
        s, d = os.path.join(REF, src), os.path.join(DST, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=ignore)
        else:
            shutil.copy2(s, d)
    with open(stamp, 'w') as f:
        f.write('verbatim copy of %s (captioning, ciderD, bleu); see oracle/build_ref.py\n' % REF)
    if verbose:
        print('copied the reference hot-path modules to', DST)
    return DST


if __name__ == '__main__':
    build(force='--force' in sys.argv)
