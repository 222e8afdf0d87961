"""CPU oracle for the SCST reward (CIDEr-D over token ids) -- TEST INFRASTRUCTURE ONLY.

Pure-Python / numpy float64 restatement of the arithmetic the reference performs on the self-critical
reward path; never imported by the product package.  Pinned by ``tests/golden/ciderd_*.npz`` which
``oracle/make_golden.py`` produced by running the reference's own scorer in the build container.

Reference lines followed (paths relative to /root/reference):
  tokens_through_eos        captioning/utils/rewards.py:33-39   (array_to_str keeps the first 0 as a token)
  ngram_counts              cider/pyciderevalcap/ciderD/ciderD_scorer.py:17-32   (precook, n = 1..4)
  tfidf_vector              cider/pyciderevalcap/ciderD/ciderD_scorer.py:156-180 (counts2vec; length = bigram count)
  clipped_similarity        cider/pyciderevalcap/ciderD/ciderD_scorer.py:53-79   (sim)
  ciderd_scores             cider/pyciderevalcap/ciderD/ciderD_scorer.py:182-208 (compute_cider, df from a pickle)
  build_document_frequency  scripts/prepro_ngrams.py:17-54 + ciderD_scorer.py:143-153
  self_critical_reward      captioning/utils/rewards.py:41-81
"""
from __future__ import annotations

import math
from collections import defaultdict
from typing import Dict, List, Sequence, Tuple

import numpy as np

NGRAM_MAX = 4
SIGMA = 6.0


def tokens_through_eos(row: Sequence[int]) -> List[int]:
    out = []
    for tok in row:
        out.append(int(tok))
        if int(tok) == 0:
            break
    return out


def ngram_counts(tokens: Sequence[int]) -> Dict[Tuple[int, ...], int]:
    counts: Dict[Tuple[int, ...], int] = defaultdict(int)
    for k in range(1, NGRAM_MAX + 1):
        for i in range(len(tokens) - k + 1):
            counts[tuple(tokens[i:i + k])] += 1
    return counts


def tfidf_vector(counts, df: Dict[Tuple[int, ...], float], log_ref_len: float):
    vec = [dict() for _ in range(NGRAM_MAX)]
    norm = [0.0] * NGRAM_MAX
    length = 0
    for gram, tf in counts.items():
        idf = log_ref_len - math.log(max(1.0, df.get(gram, 0.0)))
        n = len(gram) - 1
        vec[n][gram] = float(tf) * idf
        norm[n] += vec[n][gram] ** 2
        if n == 1:
            length += tf
    return vec, [math.sqrt(x) for x in norm], length


def clipped_similarity(vec_h, vec_r, norm_h, norm_r, len_h, len_r) -> np.ndarray:
    delta = float(len_h - len_r)
    val = np.zeros(NGRAM_MAX)
    for n in range(NGRAM_MAX):
        for gram, w in vec_h[n].items():
            r = vec_r[n].get(gram, 0.0)
            val[n] += min(w, r) * r
        if norm_h[n] != 0 and norm_r[n] != 0:
            val[n] /= norm_h[n] * norm_r[n]
        val[n] *= math.e ** (-(delta ** 2) / (2 * SIGMA ** 2))
    return val


def ciderd_scores(hyps: Sequence[Sequence[int]], refs: Sequence[Sequence[Sequence[int]]],
                  df: Dict[Tuple[int, ...], float], ref_len: float) -> np.ndarray:
    """hyps[i]: token list (already cut through the first 0); refs[i]: list of token lists. ref_len = #images."""
    log_ref_len = math.log(float(ref_len))
    out = []
    for hyp, rlist in zip(hyps, refs):
        vec, norm, length = tfidf_vector(ngram_counts(hyp), df, log_ref_len)
        score = np.zeros((len(rlist), NGRAM_MAX))
        for rid, ref in enumerate(rlist):
            vec_r, norm_r, len_r = tfidf_vector(ngram_counts(ref), df, log_ref_len)
            score[rid] += clipped_similarity(vec, vec_r, norm, norm_r, length, len_r)
        avg = np.mean(score, 1)
        out.append(np.sum(avg) / len(rlist) * 10.0)
    return np.array(out)


def build_document_frequency(ref_rows_per_image: Sequence[Sequence[Sequence[int]]]):
    """DF over images: an n-gram counts once per image whose references contain it. Rows are cut through the first 0."""
    df: Dict[Tuple[int, ...], float] = defaultdict(float)
    for rows in ref_rows_per_image:
        seen = set()
        for row in rows:
            seen.update(ngram_counts(tokens_through_eos(row)).keys())
        for gram in seen:
            df[gram] += 1.0
    return dict(df), len(ref_rows_per_image)


def self_critical_reward(greedy: np.ndarray, gts: Sequence[np.ndarray], sampled: np.ndarray,
                         df: Dict[Tuple[int, ...], float], ref_len: float, cider_weight: float = 1.0):
    """reward[i*n+j, :] = score(sample j of image i) - score(greedy of image i), repeated over the time axis."""
    B = len(gts)
    S = sampled.shape[0]
    n = S // B
    hyps = [tokens_through_eos(sampled[i]) for i in range(S)] + [tokens_through_eos(greedy[i]) for i in range(B)]
    ref_tok = [[tokens_through_eos(r) for r in gts[i]] for i in range(B)]
    refs = [ref_tok[i // n] for i in range(S)] + [ref_tok[i] for i in range(B)]
    scores = cider_weight * ciderd_scores(hyps, refs, df, ref_len)
    diff = scores[:S].reshape(B, n) - scores[-B:][:, None]
    return np.repeat(diff.reshape(S)[:, None], sampled.shape[1], 1), scores


def get_scores(gts: Sequence[np.ndarray], sampled: np.ndarray, df: Dict[Tuple[int, ...], float], ref_len: float,
               cider_weight: float = 1.0) -> np.ndarray:
    """captioning/utils/rewards.py:83-114 (CIDEr-D term): one score per sampled caption against its image's references."""
    B, S = len(gts), sampled.shape[0]
    n = S // B
    ref_tok = [[tokens_through_eos(r) for r in gts[i]] for i in range(B)]
    hyps = [tokens_through_eos(sampled[i]) for i in range(S)]
    return cider_weight * ciderd_scores(hyps, [ref_tok[i // n] for i in range(S)], df, ref_len)


from imagecaptioning.pytorch_b200.synthetic import make_refs      # noqa: E402,F401  (seeded synthetic references, shared with bench.py)
