"""Device-side mirror of captioning/utils/rewards.py for the SCST inner loop (CIDEr-D reward only).

    init_scorer(cached_tokens)                                   rewards.py:25-31
    get_self_critical_reward(greedy_res, data_gts, gen_result, opt)   rewards.py:41-81
    get_scores(data_gts, gen_result, opt)                             rewards.py:83-114

The reference moves both id tensors to the host, formats every id as a string and walks Python dicts; here the ids
never leave the GPU: n-gram extraction, the document-frequency lookup (open-addressing hash table built once from the
``scripts/prepro_ngrams.py`` pickle), the clipped tf-idf cosine and the self-critical difference run in csrc/reward.cu.
``bleu_reward_weight`` must be 0 (its default, opts.py:171).
"""
from __future__ import annotations

import os
import pickle
import threading
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


class CiderDTable:
    """Owns the device hash tables n-gram -> idf (capb200_cider_table), one per GPU, built lazily on the device that asks (under
    nn.DataParallel every replica scores its shard on its own GPU)."""

    def __init__(self, document_frequency: Dict[Tuple, float], ref_len: float, device=None):
        n = len(document_frequency)
        keys = np.full((max(n, 1), 4), -1, dtype=np.int32)
        vals = np.zeros((max(n, 1),), dtype=np.float64)
        for i, (k, v) in enumerate(document_frequency.items()):
            keys[i, :len(k)] = [int(t) for t in k]        # pickle keys are tuples of id strings (prepro_ngrams.py:42-45)
            vals[i] = float(v)
        self._keys, self._vals = keys, vals
        self.ref_len = float(ref_len)
        self.entries = n
        self._handles = {}
        self._lock = threading.Lock()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self._handle(self.device.index if self.device.index is not None else torch.cuda.current_device())

    def _handle(self, index: int):
        with self._lock:
            h = self._handles.get(index)
            if h is None:
                lib = _lib.load()
                with torch.cuda.device(index):
                    h = lib.capb200_cider_table_create(self._keys.ctypes.data, self._vals.ctypes.data, self.entries, self.ref_len, _lib.current_stream())
                if not h:
                    raise RuntimeError('capb200 cider_table_create failed: %s' % lib.capb200_last_error().decode())
                self._handles[index] = h
            return h

    @property
    def _h(self):
        """Handle of the table on the calling thread's current CUDA device."""
        return self._handle(torch.cuda.current_device())

    @classmethod
    def from_pickle(cls, path: str, device=None):
        with open(path, 'rb') as f:
            pk = pickle.load(f, encoding='latin1')
        return cls(pk['document_frequency'], pk['ref_len'], device)

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.load().capb200_cider_table_destroy(h)
            self._handles = {}
        except Exception:
            pass


CiderD_scorer: Optional[CiderDTable] = None


def init_scorer(cached_tokens, device=None):
    """Same contract as the reference: ``cached_tokens`` names ``data/<cached_tokens>.p`` relative to the cwd
    (ciderD_scorer.py:109); an existing path or an already built CiderDTable is accepted too.  Idempotent."""
    global CiderD_scorer
    if CiderD_scorer is not None:
        return CiderD_scorer
    if isinstance(cached_tokens, CiderDTable):
        CiderD_scorer = cached_tokens
    else:
        path = cached_tokens if os.path.exists(str(cached_tokens)) else os.path.join('data', str(cached_tokens) + '.p')
        CiderD_scorer = CiderDTable.from_pickle(path, device)
    return CiderD_scorer


def reset_scorer():
    global CiderD_scorer
    CiderD_scorer = None


class _Staging:
    """Per-step reference upload: round-robin PINNED host blocks (a block is reused only after the copy that read it has completed) feeding ONE
    device buffer per GPU -- the padded reference rows and the offsets travel in a single asynchronous H2D copy, and the device addresses stay
    the same from step to step (copies and the kernels that read them are ordered on the stream), which is what lets the engine replay the
    SCST step as a CUDA graph."""
    SLOTS = 4

    def __init__(self):
        self.host = {}
        self.dev = {}
        self.turn = 0

    def upload(self, host_rows: np.ndarray, offs: np.ndarray, device):
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        head = (offs.size + 1023) // 1024 * 1024          # offsets first, in a fixed-size head: both device addresses are independent of the row count
        n = head + host_rows.size
        cap = max(n, 1 << 14)
        key = (idx, self.turn % self.SLOTS)
        self.turn += 1
        slot = self.host.get(key)
        if slot is None or slot[0].numel() < n:
            slot = [torch.empty(cap, dtype=torch.int32).pin_memory(), None]
            self.host[key] = slot
        pinned, ev = slot
        if ev is not None:
            ev.synchronize()
        devbuf = self.dev.get(idx)
        if devbuf is None or devbuf.numel() < n:
            devbuf = torch.empty(cap, dtype=torch.int32, device=dev)
            self.dev[idx] = devbuf
        flat = pinned.numpy()
        flat[:offs.size] = offs
        flat[head:n] = host_rows.reshape(-1)
        devbuf[:n].copy_(pinned[:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        slot[1] = ev
        return devbuf[head:n].view(host_rows.shape), devbuf[:offs.size]


_staging = _Staging()


def pack_references(data_gts: Sequence, device) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """list[B] of int arrays [n_refs_i, L] (dataloader.py:213) -> (refs int32 [total, L], offsets int32 [B+1], L) on the device.
    The returned tensors are views of the device staging buffer, overwritten by the next upload (in stream order): consume them in the step
    they were packed for."""
    L = max(int(np.asarray(g).shape[1]) for g in data_gts)
    total = sum(int(np.asarray(g).shape[0]) for g in data_gts)
    rows = np.zeros((total, L), dtype=np.int32)
    offs = np.zeros(len(data_gts) + 1, dtype=np.int32)
    at = 0
    for i, g in enumerate(data_gts):
        g = np.asarray(g)
        rows[at:at + g.shape[0], :g.shape[1]] = g
        at += g.shape[0]
        offs[i + 1] = at
    if torch.device(device).type != 'cuda':
        return torch.from_numpy(rows), torch.from_numpy(offs), L
    refs, offsets = _staging.upload(rows, offs, device)
    return refs, offsets, L


def cider_scores_and_reward(greedy_res: torch.Tensor, data_gts: Sequence, gen_result: torch.Tensor, table: Optional[CiderDTable] = None):
    table = table or CiderD_scorer
    if table is None:
        raise RuntimeError('init_scorer(cached_tokens) must be called before the SCST reward (tools/train.py:150-152)')
    dev = gen_result.device
    if dev.type != 'cuda':
        raise RuntimeError('capb200: the reward kernel runs on CUDA tensors only')
    B = len(data_gts)
    S, T = gen_result.shape
    assert greedy_res.shape[0] == B and S % B == 0
    sampled = gen_result.detach().to(torch.long).contiguous()
    greedy = greedy_res.detach().to(torch.long).contiguous()
    refs, offsets, L = pack_references(data_gts, dev)
    scores = torch.empty(S + B, dtype=torch.float64, device=dev)
    reward = torch.empty(S, T, dtype=torch.float32, device=dev)
    lib = _lib.load()
    _lib.check(lib.capb200_self_critical_reward(table._h, _lib.ptr(sampled), S, _lib.ptr(greedy), B, T, _lib.ptr(refs), _lib.ptr(offsets), L,
                                                _lib.ptr(scores), _lib.ptr(reward), _lib.current_stream()), 'self_critical_reward')
    return scores, reward


def get_self_critical_reward(greedy_res, data_gts, gen_result, opt):
    """reward[i*n+j, :] = CIDEr-D(sample j of image i) - CIDEr-D(greedy of image i), as a device fp32 tensor [S, T]
    (the reference returns the same values as a host float64 array that LossWrapper immediately moves back to the GPU)."""
    if getattr(opt, 'bleu_reward_weight', 0) > 0:
        raise NotImplementedError('BLEU reward is out of scope of the B200 engine (bleu_reward_weight defaults to 0)')
    w = float(getattr(opt, 'cider_reward_weight', 1))
    _, reward = cider_scores_and_reward(greedy_res, data_gts, gen_result)
    return reward if w == 1.0 else reward * w


def cider_scores(data_gts: Sequence, gen_result: torch.Tensor, table: Optional[CiderDTable] = None, with_reward: bool = False):
    """CIDEr-D of every sampled caption (float64 [S]) and, optionally, the leave-one-out reward [S, T] (capb200_cider_scores)."""
    table = table or CiderD_scorer
    if table is None:
        raise RuntimeError('init_scorer(cached_tokens) must be called before the structure-loss scores (tools/train.py:150-152)')
    dev = gen_result.device
    if dev.type != 'cuda':
        raise RuntimeError('capb200: the reward kernel runs on CUDA tensors only')
    B = len(data_gts)
    S, T = gen_result.shape
    assert S % B == 0
    sampled = gen_result.detach().to(torch.long).contiguous()
    refs, offsets, L = pack_references(data_gts, dev)
    scores = torch.empty(S, dtype=torch.float64, device=dev)
    reward = torch.empty(S, T, dtype=torch.float32, device=dev) if with_reward else None
    lib = _lib.load()
    _lib.check(lib.capb200_cider_scores(table._h, _lib.ptr(sampled), S, B, T, _lib.ptr(refs), _lib.ptr(offsets), L, _lib.ptr(scores),
                                        _lib.ptr(reward) if with_reward else None, _lib.current_stream()), 'cider_scores')
    return (scores, reward) if with_reward else scores


def get_scores(data_gts, gen_result, opt):
    """rewards.py:83-114 with the CIDEr-D term only: ``cider_reward_weight * CIDEr-D`` per sampled caption, float64 [S] on the device
    (the reference returns the same values as a host numpy array)."""
    if getattr(opt, 'bleu_reward_weight', 0) > 0:
        raise NotImplementedError('BLEU reward is out of scope of the B200 engine (bleu_reward_weight defaults to 0)')
    w = float(getattr(opt, 'cider_reward_weight', 1))
    scores = cider_scores(data_gts, gen_result)
    return scores if w == 1.0 else scores * w
