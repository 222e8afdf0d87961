"""Host-side mirror of the reference's model surface for the decode hot path.

``B200UpDownModel`` / ``B200NewFCModel`` present exactly what ``captioning.models.setup(opt)`` returns for
``caption_model in ('updown', 'topdown', 'newfc')`` (captioning/models/__init__.py:20-73):

  * the same constructor argument (``opt`` namespace) and attributes (vocab_size, seq_length, bos/eos/pad/unk_idx,
    vocab, bad_endings_ix, ss_prob, done_beams)                                   AttModel.py:52-97
  * the same ``state_dict`` keys and tensor layouts, so reference checkpoints load unchanged (tools/train.py:79-80)
  * ``forward(*args, mode=...)`` dispatching to ``_forward`` / ``_sample`` (/ ``_sample_beam``)   CaptionModel.py:29-33
  * ``_sample(fc_feats, att_feats, att_masks=None, opt={})`` -> (seq int64 [B*n, T], seqLogprobs fp32 [B*n, T, V+1])
    and ``self.done_beams`` after beam search                                      AttModel.py:218-352

The parameters are ordinary ``nn.Parameter``s owned by PyTorch; every timestep of every decode runs in the hand-written
sm_100a kernels behind the C ABI (include/capb200.h).  Nothing here computes on the CPU and nothing falls back to
PyTorch ops: unsupported decode options raise.
"""
from __future__ import annotations

import ctypes
import threading
from collections.abc import Mapping
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

BAD_ENDINGS = ['a', 'an', 'the', 'in', 'for', 'at', 'of', 'with', 'before', 'after', 'on', 'upon', 'near', 'to', 'is', 'are', 'am', 'the']

_PENALTY = {'': 0, 'wu': 1, 'avg': 2}


class _DoneBeam(Mapping):
    """One finished hypothesis, dict-compatible with CaptionModel.py:190-195.  ``logps`` ([len, V+1]) is gathered from the
    engine's per-step slab on first access instead of being copied for every beam of every image."""

    def __init__(self, owner, image, rank, seq, p, raw):
        self._owner, self._image, self._rank = owner, image, rank
        self._seq, self._p, self._raw = seq, p, raw
        self._logps = None

    def _materialise(self):
        if self._logps is None:
            self._logps = self._owner._beam_logps(self._image, self._rank)[: self._seq.shape[0]]
        return self._logps

    def __getitem__(self, key):
        if key == 'seq':
            return self._seq
        if key == 'p':
            return self._p
        if key == 'logps':
            return self._materialise()
        if key == 'unaug_p':
            return float(self._materialise().sum().item())
        if key == 'sum_logp':          # extension: the raw (un-penalised) running sum
            return self._raw
        raise KeyError(key)

    def __iter__(self):
        return iter(('seq', 'logps', 'unaug_p', 'p'))

    def __len__(self):
        return 4


class _EngineStore:
    """Per-device engine handles of one model, shared BY REFERENCE between the module and the shallow replicas nn.DataParallel makes on
    every forward (replicate() copies __dict__), so each GPU keeps its engine, workspaces and CUDA graphs across steps.  Only the
    module that created the store frees the engines."""

    def __init__(self, owner):
        self.owner = id(owner)
        self.lock = threading.RLock()
        self.slots = {}

    def slot(self, dev):
        with self.lock:
            return self.slots.setdefault(dev, {'engine': None, 'key': None, 'versions': None, 'keepalive': None, 'flat': None, 'bufs': {}})

    def __deepcopy__(self, memo):
        fresh = _EngineStore(None)
        fresh.owner = None          # claimed by the copy at its first _ensure_engine
        return fresh

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.owner, self.lock, self.slots = None, threading.RLock(), {}


_tls = threading.local()           # device index the calling thread is decoding on (set by _ensure_engine)


def _on_device(fn):
    """Runs a call surface with the tensors' device current: the C ABI launches on the current device's stream and never calls
    cudaSetDevice itself, so a model living on cuda:1 while cuda:0 is current (model.to('cuda:1') without set_device) would otherwise launch
    on the wrong device with foreign pointers.  PyTorch modules handle that case transparently; so does this."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapped


def _slot_property(field):
    def get(self):
        return self._store.slot(getattr(_tls, 'dev', None))[field]

    def put(self, value):
        self._store.slot(getattr(_tls, 'dev', None))[field] = value
    return property(get, put)


class B200CaptionModel(nn.Module):
    """Common machinery: engine life-cycle, weight binding and the three call surfaces."""

    family = None           # _lib.FAMILY_*
    family_name = ''

    def __init__(self, opt, numeric_mode: Optional[str] = None):
        super().__init__()
        self.vocab_size = opt.vocab_size
        self.input_encoding_size = opt.input_encoding_size
        self.rnn_size = opt.rnn_size
        self.num_layers = getattr(opt, 'num_layers', 1)
        self.drop_prob_lm = getattr(opt, 'drop_prob_lm', 0.5)
        self.seq_length = getattr(opt, 'max_length', 20) or opt.seq_length
        self.fc_feat_size = opt.fc_feat_size
        self.att_feat_size = opt.att_feat_size
        self.att_hid_size = opt.att_hid_size
        self.bos_idx = getattr(opt, 'bos_idx', 0)
        self.eos_idx = getattr(opt, 'eos_idx', 0)
        self.pad_idx = getattr(opt, 'pad_idx', 0)
        self.unk_idx = getattr(opt, 'unk_idx', None)
        if (self.bos_idx, self.eos_idx, self.pad_idx) != (0, 0, 0):
            raise NotImplementedError('capb200 engine assumes bos = eos = pad = 0 (AttModel.py:65-67 defaults)')
        if getattr(opt, 'use_bn', 0):
            raise NotImplementedError('use_bn is not on the B200 decode path')
        if getattr(opt, 'logit_layers', 1) != 1:
            raise NotImplementedError('logit_layers > 1 is not on the B200 decode path')
        self.ss_prob = 0.0
        self.vocab = opt.vocab
        self.bad_endings_ix = [int(k) for k, v in self.vocab.items() if v in BAD_ENDINGS]
        self.numeric_mode = numeric_mode or getattr(opt, 'b200_numeric_mode', 'tc_f16x3')
        if self.numeric_mode not in _lib.MODES:
            raise ValueError('numeric_mode must be one of %s' % sorted(_lib.MODES))
        self.done_beams = []
        self._store = _EngineStore(self)

    # ---- engine plumbing --------------------------------------------------------------------------------------------
    _engine = _slot_property('engine')
    _engine_key = _slot_property('key')
    _bound_versions = _slot_property('versions')
    _keepalive = _slot_property('keepalive')
    _flat = _slot_property('flat')            # grad_sync.FlatGrads of this device (persistent flat gradient buffer of the fused training steps)
    _bufs = _slot_property('bufs')            # persistent per-shape output buffers of the fused training steps

    def _grad_groups(self):
        """[(name, parameter)] lists in the order the engine completes the gradients (include/capb200.h: *_set_grad_events)."""
        raise NotImplementedError

    def _flat_grads(self, device):
        """The persistent flat gradient buffer of this device, its {name: view} table and the engine-recorded group events.  Keyed by name:
        nn.DataParallel replicas carry different Parameter objects every forward but the same names and shapes."""
        from .grad_sync import FlatGrads
        groups = self._grad_groups()
        sig = tuple((n, tuple(p.shape)) for g in groups for n, p in g)
        fg = self._flat
        if fg is None or fg.sig != sig:
            fg = FlatGrads([[p for _, p in g] for g in groups], device)
            fg.sig = sig
            fg.by_name = {n: fg.view(p) for g in groups for n, p in g}
            self._flat = fg
        return fg

    def _step_buffers(self, key, make):      # (nn.Module owns the name _buffers)
        bufs = self._bufs
        if key not in bufs:
            bufs.clear()            # one live shape at a time: the buffers are large (the [N, T, V+1] log-prob block)
            bufs[key] = make()
        return bufs[key]

    def _enter_device(self, device):
        """Selects the per-device engine slot for this thread; returns the loaded library."""
        if device.type != 'cuda':
            raise RuntimeError('capb200: the decode engine runs on CUDA devices only (no CPU fallback); got %s' % device)
        _tls.dev = device.index if device.index is not None else torch.cuda.current_device()
        if self._store.owner is None:
            self._store.owner = id(self)
        return _lib.load()

    def _weight_table(self):
        raise NotImplementedError

    def _bind_key(self, tensors):
        """What the engine's derived weight copies (fp16 planes, fused QKV blocks, gate tables) were built from.  (data_ptr, _version)
        identifies the contents only for tensors this module owns: the parameters of an nn.DataParallel replica are fresh Broadcast outputs
        every forward (version 0, and the caching allocator hands the same addresses back after an optimizer step), so replicas re-bind on
        every call; the bound tensors are also kept alive so a freed-and-reused address can never look unchanged."""
        if getattr(self, '_is_replica', False) or any(not t.is_leaf for t in tensors):
            return None
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def _ensure_engine(self, device):
        lib = self._enter_device(device)
        key = (_tls.dev, self.numeric_mode)
        if self._engine is None or self._engine_key != key:
            self._destroy_engine()
            cfg = _lib.ModelCfg(self.family, self.vocab_size, self.input_encoding_size, self.rnn_size, self.att_hid_size, self.fc_feat_size,
                                self.att_feat_size, self.seq_length, _lib.MODES[self.numeric_mode])
            with torch.cuda.device(device):
                eng = lib.capb200_engine_create(ctypes.byref(cfg))
            if not eng:
                raise RuntimeError('capb200 engine_create failed: %s' % lib.capb200_last_error().decode())
            self._engine, self._engine_key, self._bound_versions = eng, key, None
        table = self._weight_table()
        versions = self._bind_key(list(table.values()))
        if versions is None or versions != self._bound_versions:
            w = _lib.Weights()
            keep = []
            for name, t in table.items():
                if t.device != device or t.dtype != torch.float32:
                    raise RuntimeError('capb200: parameter %s must be a float32 tensor on %s' % (name, device))
                tc = t.detach().contiguous()
                keep.append(tc)
                setattr(w, name, tc.data_ptr())
            _lib.check(lib.capb200_engine_bind_weights(self._engine, ctypes.byref(w), _lib.current_stream()), 'bind_weights')
            self._keepalive = keep
            self._bound_versions = versions
        return lib

    def _free_engine(self, handle):
        _lib.load().capb200_engine_destroy(handle)

    def _destroy_engine(self):
        if self._engine is not None:
            self._free_engine(self._engine)
            self._engine = None

    def __del__(self):
        try:
            store = self.__dict__.get('_store')
            if store is None or store.owner != id(self):
                return                                   # DataParallel replica: the engines belong to the original module
            for slot in store.slots.values():
                if slot['engine'] is not None:
                    self._free_engine(slot['engine'])
                    slot['engine'] = None
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return 0 if self._engine is None else int(_lib.load().capb200_engine_launch_count(self._engine))

    GEMM_IDS = ('fc_embed', 'att_embed', 'ctx2att', 'fc_gate_bias', 'att_lstm', 'h2att', 'lang_lstm', 'logit', 'newfc_core')

    def set_profiling(self, enable: bool):
        _lib.check(_lib.load().capb200_engine_set_profiling(self._engine, int(enable)), 'set_profiling')

    def read_profile(self, reset=True):
        """{gemm name: (milliseconds, algorithmic FLOPs, launches)} accumulated since the last reset (device-side cudaEvents)."""
        import numpy as np
        ms, fl, calls = np.zeros(9), np.zeros(9), np.zeros(9, dtype=np.int64)
        _lib.check(_lib.load().capb200_engine_read_profile(self._engine, int(reset), ms.ctypes.data, fl.ctypes.data, calls.ctypes.data, 9), 'read_profile')
        return {n: (float(ms[i]), float(fl[i]), int(calls[i])) for i, n in enumerate(self.GEMM_IDS)}

    # ---- reference surface ------------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        mode = kwargs.pop('mode', 'forward')
        return getattr(self, '_' + mode)(*args, **kwargs)

    @staticmethod
    def _f32(t):
        return None if t is None else t.detach().to(torch.float32).contiguous()

    def _clip(self, att_feats, att_masks):
        """clip_att (AttModel.py:106-112): cut the region axis to the longest valid length (one host sync, as the reference)."""
        if att_masks is not None:
            max_len = int(att_masks.detach().long().sum(1).max().item())
            att_feats = att_feats[:, :max_len]
            att_masks = att_masks[:, :max_len]
        return self._f32(att_feats), self._f32(att_masks)

    def _check_opts(self, opt):
        if opt.get('group_size', 1) != 1:
            raise NotImplementedError('diverse beam search (group_size > 1) is out of scope of the B200 engine (SURVEY.md section 8f)')
        if opt.get('output_logsoftmax', 1) != 1:
            raise NotImplementedError('output_logsoftmax=0 is out of scope of the B200 engine')

    def _decode_edits(self, opt, device, beam, batch_size=0):
        """The reference's per-step log-prob edits as a capb200_decode_edits (CaptionModel.py:118-120,154-162; AttModel.py:265-332)."""
        ed = _lib.DecodeEdits.none()
        keep = None
        ed.decoding_constraint = 1 if opt.get('decoding_constraint', 0) else 0
        if opt.get('remove_bad_endings', 0) and self.bad_endings_ix:
            keep = torch.tensor(sorted(set(self.bad_endings_ix)), dtype=torch.int32, device=device)
            ed.n_bad_endings, ed.bad_endings = keep.numel(), keep.data_ptr()
        if beam:
            # CaptionModel.py:120 reads opt.get('suppress_UNK', 0); the elif branch lowers unk_idx whatever the flag says (:161-162)
            if opt.get('suppress_UNK', 0) and self.vocab.get(str(self.vocab_size)) == 'UNK':
                ed.unk_col = self.vocab_size
            elif self.unk_idx is not None:
                ed.unk_col = int(self.unk_idx)
            if opt.get('block_trigrams', 0):
                pass                                 # beam search ignores it (the option is only read by _sample, AttModel.py:267)
        elif opt.get('block_trigrams', 0):
            ed.block_trigrams, ed.trigram_rows = 1, int(batch_size)
        return ed, keep

    @_on_device
    def _sample(self, fc_feats, att_feats, att_masks=None, opt={}, forced_tokens=None):
        sample_method = opt.get('sample_method', 'greedy')
        beam_size = opt.get('beam_size', 1)
        temperature = float(opt.get('temperature', 1.0))
        sample_n = int(opt.get('sample_n', 1))
        self._check_opts(opt)
        if beam_size > 1 and sample_method in ('greedy', 'beam_search'):
            return self._sample_beam(fc_feats, att_feats, att_masks, opt)
        top = 0.0
        if forced_tokens is not None:
            method = _lib.SAMPLE_FORCED
        elif sample_method == 'greedy':
            method = _lib.SAMPLE_GREEDY
        elif sample_method == 'sample':
            method = _lib.SAMPLE_MULTINOMIAL
        elif sample_method == 'gumbel':
            # argmax(logprobs + Gumbel noise) / temperature-free: a multinomial draw at temperature 1 (CaptionModel.py:375-385)
            method, temperature = _lib.SAMPLE_MULTINOMIAL, 1.0
        elif sample_method.startswith('top'):
            top = float(sample_method[3:])           # CaptionModel.py:387-402: 0 < x < 1 nucleus, else top-k
            if top <= 0:
                raise ValueError('sample_method %r: top-k needs k >= 1, nucleus sampling 0 < p < 1' % sample_method)
            method = _lib.SAMPLE_TOPP if top < 1 else _lib.SAMPLE_TOPK
        else:
            raise NotImplementedError("sample_method %r is out of scope of the B200 engine" % sample_method)
        lib = self._ensure_engine(fc_feats.device)
        fc = self._f32(fc_feats)
        att, masks = self._clip(att_feats, att_masks)
        B = fc.shape[0]
        R = att.shape[1] if att is not None and att.dim() == 3 else 1
        N, T, V1 = B * sample_n, self.seq_length, self.vocab_size + 1
        # every (row, step) of both outputs is written by the engine (finished rows get pad / zero rows): no memset of the [N, T, V+1] block
        seq = torch.empty(N, T, dtype=torch.long, device=fc.device)
        logprobs = torch.empty(N, T, V1, dtype=torch.float32, device=fc.device)
        draws = method in (_lib.SAMPLE_MULTINOMIAL, _lib.SAMPLE_TOPK, _lib.SAMPLE_TOPP)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if draws else 0   # follows torch.manual_seed
        edits, keep_bad = self._decode_edits(opt, fc.device, beam=False, batch_size=B)
        so = _lib.SampleOpts(sample_n, method, temperature, seed, T, top, edits)
        tok = None
        if forced_tokens is not None:
            tok = forced_tokens.detach().to(torch.long).contiguous()
            assert tok.shape == (N, T)
        _lib.check(self._call_sample(lib, fc, att, masks, B, R, so, tok, T, seq, logprobs), 'decode_sample')
        return seq, logprobs

    # family-specific C-ABI entry points (overridden by the Transformer / AoA mirrors)
    def _call_sample(self, lib, fc, att, masks, B, R, so, tok, ld_tok, seq, logprobs):
        return lib.capb200_decode_sample(self._engine, _lib.ptr(fc), _lib.ptr(att), _lib.ptr(masks), B, R, ctypes.byref(so), _lib.ptr(tok), ld_tok,
                                         _lib.ptr(seq), _lib.ptr(logprobs), None, _lib.current_stream())

    def _call_beam(self, lib, fc, att, masks, B, R, bo, seq, logprobs, d_seq, d_len, d_p, d_raw):
        return lib.capb200_decode_beam(self._engine, _lib.ptr(fc), _lib.ptr(att), _lib.ptr(masks), B, R, ctypes.byref(bo), _lib.ptr(seq),
                                       _lib.ptr(logprobs), _lib.ptr(d_seq), _lib.ptr(d_len), _lib.ptr(d_p), _lib.ptr(d_raw), _lib.current_stream())

    def _call_record(self, lib, image, rank, dst):
        return lib.capb200_beam_record_logprobs(self._engine, image, rank, _lib.ptr(dst), _lib.current_stream())

    def _teacher_steps(self, seq):
        # the reference stops at the first column i >= 1 whose labels are all pad (AttModel.py:158-159)
        col_empty = (seq[:, 1:].sum(0) == 0).nonzero()
        return int(col_empty[0].item()) + 1 if col_empty.numel() > 0 else seq.shape[1]

    @_on_device
    def _sample_beam(self, fc_feats, att_feats, att_masks=None, opt={}):
        beam_size = opt.get('beam_size', 10)
        sample_n = opt.get('sample_n', 10)
        self._check_opts(opt)
        assert sample_n == 1 or sample_n == beam_size, 'when beam search, sample_n == 1 or beam search'
        assert beam_size <= self.vocab_size + 1
        n_kinds = int(bool(opt.get('decoding_constraint', 0))) + int(bool(opt.get('remove_bad_endings', 0)) and bool(self.bad_endings_ix)) + \
            int((bool(opt.get('suppress_UNK', 0)) and self.vocab.get(str(self.vocab_size)) == 'UNK') or self.unk_idx is not None)
        if beam_size + n_kinds > 16:
            raise NotImplementedError('beam_size + number of active decode edits must be <= 16 on the B200 engine')
        cfg = opt.get('length_penalty', '')
        kind, alpha = (cfg.split('_') + ['0'])[:2] if cfg else ('', '0')
        lib = self._ensure_engine(fc_feats.device)
        fc = self._f32(fc_feats)
        att, masks = self._clip(att_feats, att_masks)
        B = fc.shape[0]
        R = att.shape[1] if att is not None and att.dim() == 3 else 1
        T, V1 = self.seq_length, self.vocab_size + 1
        dev = fc.device
        # all six outputs are written in full by the engine (zero rows / pad beyond each caption's length): no 194 MB memset per call
        seq = torch.empty(B * sample_n, T, dtype=torch.long, device=dev)
        logprobs = torch.empty(B * sample_n, T, V1, dtype=torch.float32, device=dev)
        d_seq = torch.empty(B, beam_size, T, dtype=torch.long, device=dev)
        d_len = torch.empty(B, beam_size, dtype=torch.int32, device=dev)
        d_p = torch.empty(B, beam_size, dtype=torch.float32, device=dev)
        d_raw = torch.empty(B, beam_size, dtype=torch.float32, device=dev)
        edits, keep_bad = self._decode_edits(opt, dev, beam=True)
        bo = _lib.BeamOpts(beam_size, sample_n, _PENALTY[kind], float(alpha), float(opt.get('temperature', 1.0)), edits)
        _lib.check(self._call_beam(lib, fc, att, masks, B, R, bo, seq, logprobs, d_seq, d_len, d_p, d_raw), 'decode_beam')
        self._last_beam = (d_seq, d_len, d_p, d_raw)
        self.done_beams = _LazyDoneBeams(self, B, beam_size)
        return seq, logprobs

    def _beam_logps(self, image, rank):
        dst = torch.zeros(self.seq_length, self.vocab_size + 1, dtype=torch.float32, device=self._last_beam[0].device)
        _lib.check(self._call_record(_lib.load(), image, rank, dst), 'beam_record_logprobs')
        return dst

    @_on_device
    def _forward(self, fc_feats, att_feats, seq, att_masks=None):
        """Teacher forcing (AttModel.py:126-164).  Scheduled sampling (training with ss_prob > 0) lives in the fused XE step (xe_step /
        B200LossWrapper), which is what trains; this inference-style call is plain teacher forcing."""
        if self.training and self.ss_prob > 0.0 and torch.is_grad_enabled():
            raise NotImplementedError('scheduled sampling runs inside the fused XE step (B200LossWrapper / model.xe_step), not in a bare _forward call')
        lib = self._ensure_engine(fc_feats.device)
        fc = self._f32(fc_feats)
        att, masks = self._clip(att_feats, att_masks)
        B = fc.shape[0]
        if seq.dim() == 3:
            seq = seq.reshape(-1, seq.shape[2])
        seq = seq.detach().to(torch.long).contiguous()
        spi = seq.shape[0] // B
        L = seq.shape[1]
        if L > self.seq_length + 2:
            raise ValueError('label width %d exceeds what the engine was built for' % L)
        steps = self._teacher_steps(seq)
        R = att.shape[1] if att is not None and att.dim() == 3 else 1
        out = torch.zeros(B * spi, L, self.vocab_size + 1, dtype=torch.float32, device=fc.device)
        so = _lib.SampleOpts(spi, _lib.SAMPLE_TEACHER, 1.0, 0, steps)
        _lib.check(self._call_sample(lib, fc, att, masks, B, R, so, seq, L, None, out), 'forward_teacher')
        return out


class _LazyDoneBeams(list):
    """list[B] of list[beam] of finished-beam records; host copies happen on first indexing (one D2H for the whole batch)."""

    def __init__(self, owner, B, beam):
        super().__init__()
        self._owner, self._B, self._beam, self._built = owner, B, beam, False

    def _build(self):
        if self._built:
            return
        d_seq, d_len, d_p, d_raw = self._owner._last_beam
        lens = d_len.cpu().tolist()
        ps = d_p.double().cpu().tolist()
        raws = d_raw.cpu().tolist()
        for i in range(self._B):
            super().append([_DoneBeam(self._owner, i, j, d_seq[i, j, :lens[i][j]], ps[i][j], raws[i][j]) for j in range(self._beam)])
        self._built = True

    def __getitem__(self, i):
        self._build()
        return super().__getitem__(i)

    def __iter__(self):
        self._build()
        return super().__iter__()

    def __len__(self):
        return self._B


class _UpDownCoreParams(nn.Module):
    """Parameter container with the key names of UpDownCore + Attention (AttModel.py:615-622,719-726)."""

    def __init__(self, opt):
        super().__init__()
        self.att_lstm = nn.LSTMCell(opt.input_encoding_size + opt.rnn_size * 2, opt.rnn_size)
        self.lang_lstm = nn.LSTMCell(opt.rnn_size * 2, opt.rnn_size)
        self.attention = nn.Module()
        self.attention.h2att = nn.Linear(opt.rnn_size, opt.att_hid_size)
        self.attention.alpha_net = nn.Linear(opt.att_hid_size, 1)


class B200UpDownModel(B200CaptionModel):
    """Drop-in for captioning.models.AttModel.UpDownModel (AttModel.py:868-872)."""

    family = _lib.FAMILY_UPDOWN
    family_name = 'updown'

    def __init__(self, opt, numeric_mode=None):
        super().__init__(opt, numeric_mode)
        self.num_layers = 2
        V1 = self.vocab_size + 1
        self.embed = nn.Sequential(nn.Embedding(V1, self.input_encoding_size), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.fc_embed = nn.Sequential(nn.Linear(self.fc_feat_size, self.rnn_size), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.att_embed = nn.Sequential(nn.Linear(self.att_feat_size, self.rnn_size), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.logit = nn.Linear(self.rnn_size, V1)
        self.ctx2att = nn.Linear(self.rnn_size, self.att_hid_size)
        self.core = _UpDownCoreParams(opt)

    def _weight_table(self):
        c = self.core
        return {
            'embed': self.embed[0].weight, 'fc_embed_w': self.fc_embed[0].weight, 'fc_embed_b': self.fc_embed[0].bias,
            'att_embed_w': self.att_embed[0].weight, 'att_embed_b': self.att_embed[0].bias,
            'ctx2att_w': self.ctx2att.weight, 'ctx2att_b': self.ctx2att.bias, 'logit_w': self.logit.weight, 'logit_b': self.logit.bias,
            'att_lstm_w_ih': c.att_lstm.weight_ih, 'att_lstm_w_hh': c.att_lstm.weight_hh, 'att_lstm_b_ih': c.att_lstm.bias_ih,
            'att_lstm_b_hh': c.att_lstm.bias_hh, 'lang_lstm_w_ih': c.lang_lstm.weight_ih, 'lang_lstm_w_hh': c.lang_lstm.weight_hh,
            'lang_lstm_b_ih': c.lang_lstm.bias_ih, 'lang_lstm_b_hh': c.lang_lstm.bias_hh,
            'h2att_w': c.attention.h2att.weight, 'h2att_b': c.attention.h2att.bias,
            'alpha_w': c.attention.alpha_net.weight, 'alpha_b': c.attention.alpha_net.bias,
        }


    def _grad_groups(self):
        t = self._weight_table()
        first = ('logit_w', 'logit_b')
        return [[(k, t[k]) for k in first], [(k, v) for k, v in t.items() if k not in first]]

    def _grad_table(self, lib, device):
        """capb200_updown_grads pointing into the persistent flat buffer, the group events registered with the engine."""
        fg = self._flat_grads(device)
        g = _lib.UpdownGrads()
        for name in _lib.GRAD_FIELDS:
            setattr(g, name, fg.by_name[name].data_ptr())
        # the engine records the group events only for a listener (B200LossWrapper.enable_gradient_sync); without one the whole step may run as a CUDA graph
        table, n = fg.event_table() if getattr(self, '_grad_sync_on', False) else (None, 0)
        _lib.check(lib.capb200_engine_set_grad_events(self._engine, table, n), 'set_grad_events')
        return fg, g

    # ---- SCST training step (UpDown): greedy baseline + sampling with dropout + CIDEr-D reward + RewardCriterion + BPTT -----
    @_on_device
    def scst_step(self, fc_feats, att_feats, gts, table, sample_n, temperature=1.0, drop_prob=None, seed=None, upstream=1.0, baseline='greedy',
                  forced_tokens=None, att_masks=None, keep_rows=0):
        """Runs one self-critical step entirely on the device (capb200_updown_scst_step).  Returns a dict with 'loss' (0-dim),
        'reward' [N, T], 'sample_seq', 'greedy_seq', 'sample_logprobs' and 'grads' {parameter: gradient tensor}.
        ``baseline='greedy'`` is the self-critical step (loss_wrapper.py:56-73); ``'leave_one_out'`` the 'new_self_critical' structure
        loss (losses.py:168-187): no greedy decode, each sample is scored against the mean of the image's other samples, and the
        result carries 'scores' [B, n] (the raw CIDEr-D values the reference reports as out['reward'])."""
        from .rewards import pack_references
        lib = self._ensure_engine(fc_feats.device)
        fc = self._f32(fc_feats)
        att, masks = self._clip(att_feats, att_masks)          # clip_att: cut the region axis to the longest valid length (AttModel.py:106-112)
        dev = fc.device
        B, R = att.shape[0], att.shape[1]
        N, T, V1 = B * sample_n, self.seq_length, self.vocab_size + 1
        refs, offsets, L = pack_references(gts, dev)
        table_params = self._weight_table()
        fg, g = self._grad_table(lib, dev)
        grads = fg.by_name
        # outputs live in persistent buffers (overwritten by the next step of the same shape): the step writes every row of every one
        sample_seq, greedy_seq, logprobs, reward, loss = self._step_buffers(('scst', B, sample_n), lambda: (
            torch.zeros(N, T, dtype=torch.long, device=dev), torch.zeros(B, T, dtype=torch.long, device=dev),
            torch.zeros(N, T, V1, dtype=torch.float32, device=dev), torch.empty(N, T, dtype=torch.float32, device=dev),
            torch.empty(1, dtype=torch.float32, device=dev)))
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        p = self.drop_prob_lm if drop_prob is None else drop_prob
        if baseline not in ('greedy', 'leave_one_out'):
            raise ValueError("baseline must be 'greedy' or 'leave_one_out'")
        loo = baseline == 'leave_one_out'
        forced = None
        if forced_tokens is not None:       # replay a given draw (parity tests against the reference's own samples)
            forced = forced_tokens.detach().to(device=dev, dtype=torch.long).contiguous()
            assert forced.shape == (N, T)
        row_loss = torch.empty(N, dtype=torch.float32, device=dev) if keep_rows else None      # drop_worst: per-row losses (reduction 'none')
        so = _lib.ScstOpts(sample_n, float(temperature), seed, float(p), float(upstream), _lib.BASELINE_LEAVE_ONE_OUT if loo else _lib.BASELINE_GREEDY,
                           _lib.ptr(forced), _lib.ptr(masks), int(keep_rows), _lib.ptr(row_loss))
        _lib.check(lib.capb200_updown_scst_step(self._engine, _lib.ptr(fc), _lib.ptr(att), B, R, ctypes.byref(so), table._h, _lib.ptr(refs),
                                                _lib.ptr(offsets), L, ctypes.byref(g), _lib.ptr(sample_seq), _lib.ptr(greedy_seq), _lib.ptr(logprobs),
                                                _lib.ptr(reward), _lib.ptr(loss), _lib.current_stream()), 'updown_scst_step')
        res = {'loss': loss[0], 'reward': reward, 'sample_seq': sample_seq, 'greedy_seq': None if loo else greedy_seq, 'sample_logprobs': logprobs,
               'grads': {table_params[k]: grads[k] for k in table_params}, 'seed': seed, 'flat': fg, 'row_loss': row_loss}
        return res

    @_on_device
    def xe_step(self, fc_feats, att_feats, labels, masks, label_smoothing=0.0, drop_prob=None, seed=None, upstream=1.0, att_masks=None, keep_rows=0):
        """One cross-entropy step on the device (capb200_updown_xe_step): teacher-forced forward over ``labels[..., :-1]`` in train mode,
        LanguageModelCriterion / LabelSmoothing against ``labels[..., 1:]``, ``masks[..., 1:]`` (reduction 'mean'), BPTT.
        Returns {'loss', 'logprobs' [N, L-1, V+1], 'grads' {parameter: gradient}, 'seed'}."""
        lib = self._ensure_engine(fc_feats.device)
        fc = self._f32(fc_feats)
        att, region_masks = self._clip(att_feats, att_masks)
        dev = fc.device
        B, R = att.shape[0], att.shape[1]
        if labels.dim() == 3:
            labels = labels.reshape(-1, labels.shape[2])
            masks = masks.reshape(-1, masks.shape[2])
        labels = labels.detach().to(torch.long).contiguous()
        masks = masks.detach().to(torch.float32).contiguous()
        N, Lc = labels.shape
        if N % B != 0 or Lc > self.seq_length + 2 or masks.shape != labels.shape:
            raise ValueError('labels/masks must be [B * seq_per_img, <= seq_length + 2]')
        steps = self._teacher_steps(labels[:, :-1])
        V1 = self.vocab_size + 1
        table_params = self._weight_table()
        fg, g = self._grad_table(lib, dev)
        grads = fg.by_name
        logprobs = torch.zeros(N, Lc - 1, V1, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        p = self.drop_prob_lm if drop_prob is None else drop_prob
        # scheduled sampling (self.ss_prob, set by the trainer: tools/train.py:147-148): the words actually fed are returned as 'tokens_used'
        tokens_used = torch.zeros(N, Lc - 1, dtype=torch.long, device=dev) if self.ss_prob > 0.0 else None
        row_loss = torch.empty(N, dtype=torch.float32, device=dev) if keep_rows else None
        xo = _lib.XeOpts(N // B, steps, seed, float(p), float(label_smoothing), float(upstream), _lib.ptr(region_masks), float(self.ss_prob),
                         _lib.ptr(tokens_used), int(keep_rows), _lib.ptr(row_loss))
        _lib.check(lib.capb200_updown_xe_step(self._engine, _lib.ptr(fc), _lib.ptr(att), B, R, ctypes.byref(xo), _lib.ptr(labels), _lib.ptr(masks), Lc,
                                              ctypes.byref(g), _lib.ptr(logprobs), _lib.ptr(loss), _lib.current_stream()), 'updown_xe_step')
        return {'loss': loss[0], 'logprobs': logprobs, 'grads': {table_params[k]: grads[k] for k in table_params}, 'seed': seed, 'flat': fg,
                'tokens_used': tokens_used, 'row_loss': row_loss}


class _MaxoutCoreParams(nn.Module):
    """Parameter container with the key names of FCModel.LSTMCore (FCModel.py:13-23)."""

    def __init__(self, opt):
        super().__init__()
        self.i2h = nn.Linear(opt.input_encoding_size, 5 * opt.rnn_size)
        self.h2h = nn.Linear(opt.rnn_size, 5 * opt.rnn_size)


class B200NewFCModel(B200CaptionModel):
    """Drop-in for captioning.models.AttModel.NewFCModel (AttModel.py:904-945)."""

    family = _lib.FAMILY_NEWFC
    family_name = 'newfc'

    def __init__(self, opt, numeric_mode=None):
        super().__init__(opt, numeric_mode)
        V1 = self.vocab_size + 1
        self.embed = nn.Embedding(V1, self.input_encoding_size)
        self.fc_embed = nn.Linear(self.fc_feat_size, self.input_encoding_size)
        self.logit = nn.Linear(self.rnn_size, V1)
        self._core = _MaxoutCoreParams(opt)

    def _weight_table(self):
        return {
            'embed': self.embed.weight, 'fc_embed_w': self.fc_embed.weight, 'fc_embed_b': self.fc_embed.bias,
            'logit_w': self.logit.weight, 'logit_b': self.logit.bias,
            'i2h_w': self._core.i2h.weight, 'i2h_b': self._core.i2h.bias, 'h2h_w': self._core.h2h.weight, 'h2h_b': self._core.h2h.bias,
        }


def _mha_params(d_model):
    m = nn.Module()
    m.linears = nn.ModuleList([nn.Linear(d_model, d_model) for _ in range(4)])
    return m


def _ln_params(d_model):
    m = nn.Module()
    m.a_2 = nn.Parameter(torch.ones(d_model))
    m.b_2 = nn.Parameter(torch.zeros(d_model))
    return m


def _tfm_layer(d_model, d_ff, n_sub, with_src):
    layer = nn.Module()
    layer.self_attn = _mha_params(d_model)
    if with_src:
        layer.src_attn = _mha_params(d_model)
    layer.feed_forward = nn.Module()
    layer.feed_forward.w_1 = nn.Linear(d_model, d_ff)
    layer.feed_forward.w_2 = nn.Linear(d_ff, d_model)
    layer.sublayer = nn.ModuleList()
    for _ in range(n_sub):
        sub = nn.Module()
        sub.norm = _ln_params(d_model)
        layer.sublayer.append(sub)
    return layer


class B200TransformerModel(B200CaptionModel):
    """Drop-in for captioning.models.TransformerModel.TransformerModel (TransformerModel.py:237-363): same opt fields (N_enc, N_dec,
    d_model, d_ff, num_att_heads), same state_dict keys (att_embed.0.*, model.encoder/decoder.layers.*, model.tgt_embed.0.lut.weight,
    model.tgt_embed.1.pe buffer, model.generator.proj.*).  Decoding keeps a per-layer K/V cache on the device."""

    family_name = 'transformer'

    def __init__(self, opt, numeric_mode=None):
        super().__init__(opt, numeric_mode)
        import math
        self.N_enc = getattr(opt, 'N_enc', opt.num_layers)
        self.N_dec = getattr(opt, 'N_dec', opt.num_layers)
        self.d_model = getattr(opt, 'd_model', opt.input_encoding_size)
        self.d_ff = getattr(opt, 'd_ff', opt.rnn_size)
        self.h = getattr(opt, 'num_att_heads', 8)
        self.dropout = getattr(opt, 'dropout', 0.1)             # every nn.Dropout inside make_model (TransformerModel.py:240-253)
        if self.N_enc > _lib.TFM_MAX_LAYERS or self.N_dec > _lib.TFM_MAX_LAYERS:
            raise NotImplementedError('at most %d layers per stack' % _lib.TFM_MAX_LAYERS)
        V1, D = self.vocab_size + 1, self.d_model
        self.att_embed = nn.Sequential(nn.Linear(self.att_feat_size, D), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.model = nn.Module()
        self.model.encoder = nn.Module()
        self.model.encoder.layers = nn.ModuleList([_tfm_layer(D, self.d_ff, 2, False) for _ in range(self.N_enc)])
        self.model.encoder.norm = _ln_params(D)
        self.model.decoder = nn.Module()
        self.model.decoder.layers = nn.ModuleList([_tfm_layer(D, self.d_ff, 3, True) for _ in range(self.N_dec)])
        self.model.decoder.norm = _ln_params(D)
        emb = nn.Module()
        emb.lut = nn.Embedding(V1, D)
        pos = nn.Module()
        pe = torch.zeros(5000, D)
        position = torch.arange(0, 5000).unsqueeze(1).float()
        div_term = torch.exp(torch.arange(0, D, 2).float() * -(math.log(10000.0) / D))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        pos.register_buffer('pe', pe.unsqueeze(0))
        self.model.tgt_embed = nn.Sequential(emb, pos)
        self.model.generator = nn.Module()
        self.model.generator.proj = nn.Linear(D, V1)
        for p_ in self.model.parameters():        # Glorot init like make_model (TransformerModel.py:255-258)
            if p_.dim() > 1:
                nn.init.xavier_uniform_(p_)

    # ---- engine plumbing (own C-ABI entry points: capb200_tfm_*) ------------------------------------------------------
    def _tensors(self):
        out = [self.att_embed[0].weight, self.att_embed[0].bias, self.model.tgt_embed[0].lut.weight, self.model.tgt_embed[1].pe,
               self.model.generator.proj.weight, self.model.generator.proj.bias]
        out += list(self.model.encoder.parameters()) + list(self.model.decoder.parameters())
        return out

    def _ensure_engine(self, device):
        lib = self._enter_device(device)
        key = (_tls.dev, self.numeric_mode)
        if self._engine is None or self._engine_key != key:
            self._destroy_engine()
            cfg = _lib.TfmCfg(self.vocab_size, self.d_model, self.d_ff, self.h, self.N_enc, self.N_dec, self.att_feat_size, self.seq_length,
                              _lib.MODES[self.numeric_mode])
            with torch.cuda.device(device):
                eng = lib.capb200_tfm_create(ctypes.byref(cfg))
            if not eng:
                raise RuntimeError('capb200 tfm_create failed: %s' % lib.capb200_last_error().decode())
            self._engine, self._engine_key, self._bound_versions = eng, key, None
        tensors = self._tensors()
        versions = self._bind_key(tensors)
        if versions is None or versions != self._bound_versions:
            for t in tensors:
                if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError('capb200: parameters must be contiguous float32 tensors on %s' % device)
            w = _lib.TfmWeights()
            P = lambda t: t.data_ptr()

            def mha(dst, src):
                for name, lin in zip(('q', 'k', 'v', 'o'), src.linears):
                    setattr(dst, name + '_w', P(lin.weight))
                    setattr(dst, name + '_b', P(lin.bias))

            def common(dst, layer, n_sub):
                dst.w1_w, dst.w1_b = P(layer.feed_forward.w_1.weight), P(layer.feed_forward.w_1.bias)
                dst.w2_w, dst.w2_b = P(layer.feed_forward.w_2.weight), P(layer.feed_forward.w_2.bias)
                for j in range(n_sub):
                    setattr(dst, 'ln%d_a' % j, P(layer.sublayer[j].norm.a_2))
                    setattr(dst, 'ln%d_b' % j, P(layer.sublayer[j].norm.b_2))

            w.att_embed_w, w.att_embed_b = P(self.att_embed[0].weight), P(self.att_embed[0].bias)
            for i, layer in enumerate(self.model.encoder.layers):
                mha(w.enc[i].self_attn, layer.self_attn)
                common(w.enc[i], layer, 2)
            for i, layer in enumerate(self.model.decoder.layers):
                mha(w.dec[i].self_attn, layer.self_attn)
                mha(w.dec[i].src_attn, layer.src_attn)
                common(w.dec[i], layer, 3)
            w.enc_norm_a, w.enc_norm_b = P(self.model.encoder.norm.a_2), P(self.model.encoder.norm.b_2)
            w.dec_norm_a, w.dec_norm_b = P(self.model.decoder.norm.a_2), P(self.model.decoder.norm.b_2)
            w.lut, w.pe = P(self.model.tgt_embed[0].lut.weight), P(self.model.tgt_embed[1].pe)
            w.gen_w, w.gen_b = P(self.model.generator.proj.weight), P(self.model.generator.proj.bias)
            _lib.check(lib.capb200_tfm_bind_weights(self._engine, ctypes.byref(w), _lib.current_stream()), 'tfm_bind_weights')
            self._keepalive = tensors
            self._bound_versions = versions
        return lib

    def _free_engine(self, handle):
        _lib.load().capb200_tfm_destroy(handle)

    @property
    def launch_count(self) -> int:
        return 0 if self._engine is None else int(_lib.load().capb200_tfm_launch_count(self._engine))

    # ---- calls: the transformer ignores fc_feats (TransformerModel.py:305-310) ---------------------------------------
    def _call_sample(self, lib, fc, att, masks, B, R, so, tok, ld_tok, seq, logprobs):
        return lib.capb200_tfm_decode_sample(self._engine, _lib.ptr(att), _lib.ptr(masks), B, R, ctypes.byref(so), _lib.ptr(tok), ld_tok,
                                             _lib.ptr(seq), _lib.ptr(logprobs), None, _lib.current_stream())

    def _call_beam(self, lib, fc, att, masks, B, R, bo, seq, logprobs, d_seq, d_len, d_p, d_raw):
        return lib.capb200_tfm_decode_beam(self._engine, _lib.ptr(att), _lib.ptr(masks), B, R, ctypes.byref(bo), _lib.ptr(seq), _lib.ptr(logprobs),
                                           _lib.ptr(d_seq), _lib.ptr(d_len), _lib.ptr(d_p), _lib.ptr(d_raw), _lib.current_stream())

    def _call_record(self, lib, image, rank, dst):
        return lib.capb200_tfm_beam_record_logprobs(self._engine, image, rank, _lib.ptr(dst), _lib.current_stream())

    def _teacher_steps(self, seq):
        return seq.shape[1]            # one parallel pass in the reference: every position is computed (TransformerModel.py:340-348)

    # ---- training steps (capb200_tfm_xe_step / capb200_tfm_scst_step) ---------------------------------------------------------------
    def _slots(self):
        """[(path into capb200_tfm_weights / capb200_tfm_grads, parameter)]: path = (field,) | ('enc'|'dec', layer, field) | ('enc'|'dec', layer, attn, field)."""
        out = [(('att_embed_w',), self.att_embed[0].weight), (('att_embed_b',), self.att_embed[0].bias)]

        def layer_slots(kind, i, layer, attns, n_sub):
            for an in attns:          # q | k | v weights (and biases) back to back: one GEMM / one column reduction per triple in the engine
                lins = getattr(layer, an).linears
                for suffix, attr in (('_w', 'weight'), ('_b', 'bias')):
                    for name, lin in zip(('q', 'k', 'v'), lins):
                        out.append(((kind, i, an, name + suffix), getattr(lin, attr)))
                out.append(((kind, i, an, 'o_w'), lins[3].weight))
                out.append(((kind, i, an, 'o_b'), lins[3].bias))
            out.append(((kind, i, 'w1_w'), layer.feed_forward.w_1.weight)); out.append(((kind, i, 'w1_b'), layer.feed_forward.w_1.bias))
            out.append(((kind, i, 'w2_w'), layer.feed_forward.w_2.weight)); out.append(((kind, i, 'w2_b'), layer.feed_forward.w_2.bias))
            for j in range(n_sub):
                out.append(((kind, i, 'ln%d_a' % j), layer.sublayer[j].norm.a_2)); out.append(((kind, i, 'ln%d_b' % j), layer.sublayer[j].norm.b_2))

        for i, layer in enumerate(self.model.encoder.layers):
            layer_slots('enc', i, layer, ('self_attn',), 2)
        out += [(('enc_norm_a',), self.model.encoder.norm.a_2), (('enc_norm_b',), self.model.encoder.norm.b_2)]
        for i, layer in enumerate(self.model.decoder.layers):
            layer_slots('dec', i, layer, ('self_attn', 'src_attn'), 3)
        out += [(('dec_norm_a',), self.model.decoder.norm.a_2), (('dec_norm_b',), self.model.decoder.norm.b_2),
                (('lut',), self.model.tgt_embed[0].lut.weight), (('gen_w',), self.model.generator.proj.weight), (('gen_b',), self.model.generator.proj.bias)]
        return out

    @staticmethod
    def _slot_name(path):
        return '/'.join(str(x) for x in path)

    def _grad_groups(self):
        slots = [(self._slot_name(path), prm) for path, prm in self._slots()]
        late = [s for s in slots if s[0].startswith(('enc', 'att_embed'))]              # encoder + att_embed finish last
        early = [s for s in slots if not s[0].startswith(('enc', 'att_embed'))]
        return [early, late]

    def _grad_table(self, lib, device):
        fg = self._flat_grads(device)
        g = _lib.TfmWeights()
        for path, _ in self._slots():
            ptr = fg.by_name[self._slot_name(path)].data_ptr()
            dst = g
            for key in path[:-1]:
                dst = getattr(dst, key) if isinstance(key, str) else dst[key]
            setattr(dst, path[-1], ptr)
        # the engine records the group events only for a listener (B200LossWrapper.enable_gradient_sync); without one the whole step may run as a CUDA graph
        table, n = fg.event_table() if getattr(self, '_grad_sync_on', False) else (None, 0)
        _lib.check(lib.capb200_tfm_set_grad_events(self._engine, table, n), 'tfm_set_grad_events')
        return fg, g

    def _result_grads(self, fg):
        return {prm: fg.by_name[self._slot_name(path)] for path, prm in self._slots()}

    @_on_device
    def xe_step(self, fc_feats, att_feats, labels, masks, label_smoothing=0.0, drop_prob=None, seed=None, upstream=1.0, dropout=None, att_masks=None,
                keep_rows=0):
        """One cross-entropy step of the Transformer on the device (capb200_tfm_xe_step): the teacher-forced pass over every position
        (TransformerModel.py:340-348), LanguageModelCriterion / LabelSmoothing, backward through decoder and encoder.  ``drop_prob`` is
        att_embed's dropout (drop_prob_lm), ``dropout`` the Transformer's own rate.  Result as B200UpDownModel.xe_step."""
        lib = self._ensure_engine(att_feats.device)
        att, region_masks = self._clip(att_feats, att_masks)
        dev = att.device
        B, R = att.shape[0], att.shape[1]
        if labels.dim() == 3:
            labels = labels.reshape(-1, labels.shape[2])
            masks = masks.reshape(-1, masks.shape[2])
        labels = labels.detach().to(torch.long).contiguous()
        masks = masks.detach().to(torch.float32).contiguous()
        N, Lc = labels.shape
        if N % B != 0 or Lc > self.seq_length + 2 or masks.shape != labels.shape:
            raise ValueError('labels/masks must be [B * seq_per_img, <= seq_length + 2]')
        # self.ss_prob is ignored, as in the reference: TransformerModel._forward is one parallel pass with no scheduled-sampling branch
        V1 = self.vocab_size + 1
        fg, g = self._grad_table(lib, dev)
        logprobs = torch.empty(N, Lc - 1, V1, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        p_lm = self.drop_prob_lm if drop_prob is None else drop_prob
        p = self.dropout if dropout is None else dropout
        row_loss = torch.empty(N, dtype=torch.float32, device=dev) if keep_rows else None
        xo = _lib.TfmXeOpts(N // B, seed, float(label_smoothing), float(upstream), float(p_lm), float(p), _lib.ptr(region_masks), int(keep_rows), _lib.ptr(row_loss))
        _lib.check(lib.capb200_tfm_xe_step(self._engine, _lib.ptr(att), B, R, ctypes.byref(xo), _lib.ptr(labels), _lib.ptr(masks), Lc, ctypes.byref(g),
                                           _lib.ptr(logprobs), _lib.ptr(loss), _lib.current_stream()), 'tfm_xe_step')
        return {'loss': loss[0], 'logprobs': logprobs, 'grads': self._result_grads(fg), 'seed': seed, 'flat': fg, 'tokens_used': None, 'row_loss': row_loss}

    @_on_device
    def scst_step(self, fc_feats, att_feats, gts, table, sample_n, temperature=1.0, drop_prob=None, seed=None, upstream=1.0, baseline='greedy', dropout=None,
                  forced_tokens=None, att_masks=None, keep_rows=0):
        """One self-critical step of the Transformer on the device (capb200_tfm_scst_step): eval-mode greedy baseline (or leave-one-out),
        train-mode samples drawn position by position on the K/V tape, CIDEr-D reward, RewardCriterion, batched backward.  Result as
        B200UpDownModel.scst_step."""
        from .rewards import pack_references
        lib = self._ensure_engine(att_feats.device)
        att, masks = self._clip(att_feats, att_masks)
        dev = att.device
        B, R = att.shape[0], att.shape[1]
        N, T, V1 = B * sample_n, self.seq_length, self.vocab_size + 1
        refs, offsets, L = pack_references(gts, dev)
        fg, g = self._grad_table(lib, dev)
        if baseline not in ('greedy', 'leave_one_out'):
            raise ValueError("baseline must be 'greedy' or 'leave_one_out'")
        loo = baseline == 'leave_one_out'
        sample_seq, greedy_seq, logprobs, reward, loss = self._step_buffers(('scst', B, sample_n), lambda: (
            torch.zeros(N, T, dtype=torch.long, device=dev), torch.zeros(B, T, dtype=torch.long, device=dev),
            torch.zeros(N, T, V1, dtype=torch.float32, device=dev), torch.empty(N, T, dtype=torch.float32, device=dev),
            torch.empty(1, dtype=torch.float32, device=dev)))
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        p_lm = self.drop_prob_lm if drop_prob is None else drop_prob
        p = self.dropout if dropout is None else dropout
        forced = None
        if forced_tokens is not None:
            forced = forced_tokens.detach().to(device=dev, dtype=torch.long).contiguous()
            assert forced.shape == (N, T)
        row_loss = torch.empty(N, dtype=torch.float32, device=dev) if keep_rows else None
        so = _lib.TfmScstOpts(sample_n, float(temperature), seed, float(upstream), _lib.BASELINE_LEAVE_ONE_OUT if loo else _lib.BASELINE_GREEDY, float(p_lm),
                              float(p), _lib.ptr(forced), _lib.ptr(masks), int(keep_rows), _lib.ptr(row_loss))
        _lib.check(lib.capb200_tfm_scst_step(self._engine, _lib.ptr(att), B, R, ctypes.byref(so), table._h, _lib.ptr(refs), _lib.ptr(offsets), L,
                                             ctypes.byref(g), _lib.ptr(sample_seq), None if loo else _lib.ptr(greedy_seq), _lib.ptr(logprobs),
                                             _lib.ptr(reward), _lib.ptr(loss), _lib.current_stream()), 'tfm_scst_step')
        return {'loss': loss[0], 'reward': reward, 'sample_seq': sample_seq, 'greedy_seq': None if loo else greedy_seq, 'sample_logprobs': logprobs,
                'grads': self._result_grads(fg), 'seed': seed, 'flat': fg, 'row_loss': row_loss}


class B200AoAModel(B200CaptionModel):
    """Drop-in for captioning.models.AoAModel.AoAModel in the configs/aoa.yml configuration (refine=1, refine_aoa=1, use_ff=0,
    decoder_type='AoA', use_multi_head=2, multi_head_scale=1, mean_feats=1): same state_dict keys (no fc_embed), same surfaces."""

    family_name = 'aoa'

    def __init__(self, opt, numeric_mode=None):
        super().__init__(opt, numeric_mode)
        need = dict(refine=1, refine_aoa=1, use_ff=0, decoder_type='AoA', use_multi_head=2, multi_head_scale=1)
        for k, v in need.items():
            if getattr(opt, k, v) != v:
                raise NotImplementedError('AoA option %s=%r is outside the configs/aoa.yml configuration the B200 engine implements' % (k, getattr(opt, k)))
        if not getattr(opt, 'mean_feats', 1):
            raise NotImplementedError('mean_feats=0 is outside the configs/aoa.yml configuration')
        if getattr(opt, 'out_res', 0):
            raise NotImplementedError('out_res is outside the configs/aoa.yml configuration')
        self.num_layers = 2
        self.num_heads = opt.num_heads
        self.dropout_aoa = getattr(opt, 'dropout_aoa', 0.3)          # AoAModel.py:117
        self.ctx_drop = getattr(opt, 'ctx_drop', 0)                  # AoAModel.py:134
        H, E, V1 = self.rnn_size, self.input_encoding_size, self.vocab_size + 1
        self.embed = nn.Sequential(nn.Embedding(V1, E), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.att_embed = nn.Sequential(nn.Linear(self.att_feat_size, H), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.logit = nn.Linear(H, V1)
        self.ctx2att = nn.Linear(H, 2 * H)
        self.refiner = nn.Module()
        self.refiner.layers = nn.ModuleList()
        for _ in range(_lib.AOA_REFINER_LAYERS):
            layer = nn.Module()
            layer.self_attn = nn.Module()
            layer.self_attn.linears = nn.ModuleList([nn.Linear(H, H) for _ in range(3)])
            layer.self_attn.aoa_layer = nn.Sequential(nn.Linear(2 * H, 2 * H), nn.GLU())
            sub = nn.Module()
            sub.norm = _ln_params(H)
            layer.sublayer = nn.ModuleList([sub])
            self.refiner.layers.append(layer)
        self.refiner.norm = _ln_params(H)
        self.core = nn.Module()
        self.core.att_lstm = nn.LSTMCell(E + H, H)
        self.core.att2ctx = nn.Sequential(nn.Linear(2 * H, 2 * H), nn.GLU())
        self.core.attention = nn.Module()
        self.core.attention.norm = _ln_params(H)
        self.core.attention.linears = nn.ModuleList([nn.Linear(H, H)])

    def _tensors(self):
        return [prm for _, prm in self._slots()]          # attribute access: works on nn.DataParallel replicas too (their parameters() is empty)

    def _slots(self):
        """(field path in capb200_aoa_weights / capb200_aoa_grads, parameter) pairs."""
        out = [(('embed',), self.embed[0].weight), (('att_embed_w',), self.att_embed[0].weight), (('att_embed_b',), self.att_embed[0].bias)]
        for i, layer in enumerate(self.refiner.layers):
            for name, lin in zip(('q', 'k', 'v'), layer.self_attn.linears):
                out += [(('refiner', i, name + '_w'), lin.weight), (('refiner', i, name + '_b'), lin.bias)]
            out += [(('refiner', i, 'aoa_w'), layer.self_attn.aoa_layer[0].weight), (('refiner', i, 'aoa_b'), layer.self_attn.aoa_layer[0].bias),
                    (('refiner', i, 'ln_a'), layer.sublayer[0].norm.a_2), (('refiner', i, 'ln_b'), layer.sublayer[0].norm.b_2)]
        c = self.core
        out += [(('refiner_norm_a',), self.refiner.norm.a_2), (('refiner_norm_b',), self.refiner.norm.b_2),
                (('ctx2att_w',), self.ctx2att.weight), (('ctx2att_b',), self.ctx2att.bias),
                (('att_lstm_w_ih',), c.att_lstm.weight_ih), (('att_lstm_w_hh',), c.att_lstm.weight_hh),
                (('att_lstm_b_ih',), c.att_lstm.bias_ih), (('att_lstm_b_hh',), c.att_lstm.bias_hh),
                (('attn_norm_a',), c.attention.norm.a_2), (('attn_norm_b',), c.attention.norm.b_2),
                (('attn_q_w',), c.attention.linears[0].weight), (('attn_q_b',), c.attention.linears[0].bias),
                (('att2ctx_w',), c.att2ctx[0].weight), (('att2ctx_b',), c.att2ctx[0].bias),
                (('logit_w',), self.logit.weight), (('logit_b',), self.logit.bias)]
        return out

    def _grad_groups(self):
        slots = {'/'.join(str(x) for x in path): prm for path, prm in self._slots()}
        pick = lambda names: [(n, slots[n]) for n in names]
        groups = [pick(['logit_w', 'logit_b']),
                  pick(['att2ctx_w', 'att2ctx_b', 'attn_q_w', 'attn_q_b', 'att_lstm_w_ih', 'att_lstm_w_hh', 'att_lstm_b_ih', 'att_lstm_b_hh', 'attn_norm_a',
                        'attn_norm_b', 'embed']),
                  pick(['ctx2att_w', 'ctx2att_b', 'refiner_norm_a', 'refiner_norm_b'])]
        for l in reversed(range(_lib.AOA_REFINER_LAYERS)):
            # q | k | v weights (and biases) back to back: the engine then writes each triple with one GEMM / one column reduction
            groups.append(pick(['refiner/%d/%s' % (l, f) for f in ('q_w', 'k_w', 'v_w', 'q_b', 'k_b', 'v_b', 'aoa_w', 'aoa_b', 'ln_a', 'ln_b')]))
        groups.append(pick(['att_embed_w', 'att_embed_b']))
        assert sum(len(g) for g in groups) == len(slots)
        return groups

    def _grad_table(self, lib, device):
        fg = self._flat_grads(device)
        g = _lib.AoaWeights()
        for path, _ in self._slots():
            ptr = fg.by_name['/'.join(str(x) for x in path)].data_ptr()
            if len(path) == 1:
                setattr(g, path[0], ptr)
            else:
                setattr(g.refiner[path[1]], path[2], ptr)
        # the engine records the group events only for a listener (B200LossWrapper.enable_gradient_sync); without one the whole step may run as a CUDA graph
        table, n = fg.event_table() if getattr(self, '_grad_sync_on', False) else (None, 0)
        _lib.check(lib.capb200_aoa_set_grad_events(self._engine, table, n), 'aoa_set_grad_events')
        return fg, g

    def _fill_table(self, table, tensor_of):
        """Writes data pointers into an AoaWeights-layout ctypes struct; ``tensor_of`` maps id(parameter) -> tensor to point at."""
        for path, prm in self._slots():
            ptr = tensor_of[id(prm)].data_ptr()
            if len(path) == 1:
                setattr(table, path[0], ptr)
            else:
                setattr(table.refiner[path[1]], path[2], ptr)

    @_on_device
    def scst_step(self, fc_feats, att_feats, gts, table, sample_n, temperature=1.0, drop_prob=None, seed=None, upstream=1.0, baseline='greedy',
                  drop_attn=0.1, drop_aoa=None, drop_sublayer=0.1, ctx_drop=None, forced_tokens=None, att_masks=None, keep_rows=0):
        """One self-critical step of AoANet on the device (capb200_aoa_scst_step): eval-mode greedy baseline (or the leave-one-out baseline of
        'new_self_critical'), train-mode samples with every dropout site of AoAModel.py active, CIDEr-D reward, RewardCriterion, BPTT through
        the decoder and the six refiner layers.  ``fc_feats`` is unused (mean_feats=1).  Returns the dict of B200UpDownModel.scst_step."""
        from .rewards import pack_references
        lib = self._ensure_engine(att_feats.device)
        att, masks = self._clip(att_feats, att_masks)
        dev = att.device
        B, R = att.shape[0], att.shape[1]
        N, T, V1 = B * sample_n, self.seq_length, self.vocab_size + 1
        refs, offsets, L = pack_references(gts, dev)
        slots = self._slots()
        fg, g = self._grad_table(lib, dev)
        if baseline not in ('greedy', 'leave_one_out'):
            raise ValueError("baseline must be 'greedy' or 'leave_one_out'")
        loo = baseline == 'leave_one_out'
        sample_seq, greedy_seq, logprobs, reward, loss = self._step_buffers(('scst', B, sample_n), lambda: (
            torch.zeros(N, T, dtype=torch.long, device=dev), torch.zeros(B, T, dtype=torch.long, device=dev),
            torch.zeros(N, T, V1, dtype=torch.float32, device=dev), torch.empty(N, T, dtype=torch.float32, device=dev),
            torch.empty(1, dtype=torch.float32, device=dev)))
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        p = self.drop_prob_lm if drop_prob is None else drop_prob
        forced = None
        if forced_tokens is not None:
            forced = forced_tokens.detach().to(device=dev, dtype=torch.long).contiguous()
            assert forced.shape == (N, T)
        row_loss = torch.empty(N, dtype=torch.float32, device=dev) if keep_rows else None
        so = _lib.AoaScstOpts(sample_n, float(temperature), seed, float(upstream), _lib.BASELINE_LEAVE_ONE_OUT if loo else _lib.BASELINE_GREEDY, float(p),
                              float(drop_attn), float(self.dropout_aoa if drop_aoa is None else drop_aoa), float(drop_sublayer),
                              int(self.ctx_drop if ctx_drop is None else ctx_drop), _lib.ptr(forced), _lib.ptr(masks), int(keep_rows), _lib.ptr(row_loss))
        _lib.check(lib.capb200_aoa_scst_step(self._engine, _lib.ptr(att), B, R, ctypes.byref(so), table._h, _lib.ptr(refs), _lib.ptr(offsets), L,
                                             ctypes.byref(g), _lib.ptr(sample_seq), None if loo else _lib.ptr(greedy_seq), _lib.ptr(logprobs),
                                             _lib.ptr(reward), _lib.ptr(loss), _lib.current_stream()), 'aoa_scst_step')
        return {'loss': loss[0], 'reward': reward, 'sample_seq': sample_seq, 'greedy_seq': None if loo else greedy_seq, 'sample_logprobs': logprobs,
                'grads': {prm: fg.by_name['/'.join(str(x) for x in path)] for path, prm in slots}, 'seed': seed, 'flat': fg, 'row_loss': row_loss}

    @_on_device
    def xe_step(self, fc_feats, att_feats, labels, masks, label_smoothing=0.0, drop_prob=None, seed=None, upstream=1.0, drop_attn=0.1, drop_aoa=None,
                drop_sublayer=0.1, ctx_drop=None, att_masks=None, keep_rows=0):
        """One cross-entropy step of AoANet on the device (capb200_aoa_xe_step); arguments and result as B200UpDownModel.xe_step."""
        lib = self._ensure_engine(att_feats.device)
        att, region_masks = self._clip(att_feats, att_masks)
        dev = att.device
        B, R = att.shape[0], att.shape[1]
        if labels.dim() == 3:
            labels = labels.reshape(-1, labels.shape[2])
            masks = masks.reshape(-1, masks.shape[2])
        labels = labels.detach().to(torch.long).contiguous()
        masks = masks.detach().to(torch.float32).contiguous()
        N, Lc = labels.shape
        if N % B != 0 or Lc > self.seq_length + 2 or masks.shape != labels.shape:
            raise ValueError('labels/masks must be [B * seq_per_img, <= seq_length + 2]')
        steps = self._teacher_steps(labels[:, :-1])
        V1 = self.vocab_size + 1
        slots = self._slots()
        fg, g = self._grad_table(lib, dev)
        logprobs = torch.zeros(N, Lc - 1, V1, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        p = self.drop_prob_lm if drop_prob is None else drop_prob
        tokens_used = torch.zeros(N, Lc - 1, dtype=torch.long, device=dev) if self.ss_prob > 0.0 else None
        row_loss = torch.empty(N, dtype=torch.float32, device=dev) if keep_rows else None
        xo = _lib.AoaXeOpts(N // B, steps, seed, float(label_smoothing), float(upstream), float(p), float(drop_attn),
                            float(self.dropout_aoa if drop_aoa is None else drop_aoa), float(drop_sublayer), int(self.ctx_drop if ctx_drop is None else ctx_drop),
                            _lib.ptr(region_masks), float(self.ss_prob), _lib.ptr(tokens_used), int(keep_rows), _lib.ptr(row_loss))
        _lib.check(lib.capb200_aoa_xe_step(self._engine, _lib.ptr(att), B, R, ctypes.byref(xo), _lib.ptr(labels), _lib.ptr(masks), Lc, ctypes.byref(g),
                                           _lib.ptr(logprobs), _lib.ptr(loss), _lib.current_stream()), 'aoa_xe_step')
        return {'loss': loss[0], 'logprobs': logprobs, 'grads': {prm: fg.by_name['/'.join(str(x) for x in path)] for path, prm in slots}, 'seed': seed,
                'flat': fg, 'tokens_used': tokens_used, 'row_loss': row_loss}

    def _ensure_engine(self, device):
        lib = self._enter_device(device)
        key = (_tls.dev, self.numeric_mode)
        if self._engine is None or self._engine_key != key:
            self._destroy_engine()
            cfg = _lib.AoaCfg(self.vocab_size, self.input_encoding_size, self.rnn_size, self.num_heads, self.att_feat_size, self.seq_length,
                              _lib.MODES[self.numeric_mode])
            with torch.cuda.device(device):
                eng = lib.capb200_aoa_create(ctypes.byref(cfg))
            if not eng:
                raise RuntimeError('capb200 aoa_create failed: %s' % lib.capb200_last_error().decode())
            self._engine, self._engine_key, self._bound_versions = eng, key, None
        tensors = self._tensors()
        versions = self._bind_key(tensors)
        if versions is None or versions != self._bound_versions:
            for t in tensors:
                if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError('capb200: parameters must be contiguous float32 tensors on %s' % device)
            w = _lib.AoaWeights()
            self._fill_table(w, {id(t): t for t in tensors})
            _lib.check(lib.capb200_aoa_bind_weights(self._engine, ctypes.byref(w), _lib.current_stream()), 'aoa_bind_weights')
            self._keepalive = tensors
            self._bound_versions = versions
        return lib

    def _free_engine(self, handle):
        _lib.load().capb200_aoa_destroy(handle)

    @property
    def launch_count(self) -> int:
        return 0 if self._engine is None else int(_lib.load().capb200_aoa_launch_count(self._engine))

    def _call_sample(self, lib, fc, att, masks, B, R, so, tok, ld_tok, seq, logprobs):
        return lib.capb200_aoa_decode_sample(self._engine, _lib.ptr(att), _lib.ptr(masks), B, R, ctypes.byref(so), _lib.ptr(tok), ld_tok,
                                             _lib.ptr(seq), _lib.ptr(logprobs), None, _lib.current_stream())

    def _call_beam(self, lib, fc, att, masks, B, R, bo, seq, logprobs, d_seq, d_len, d_p, d_raw):
        return lib.capb200_aoa_decode_beam(self._engine, _lib.ptr(att), _lib.ptr(masks), B, R, ctypes.byref(bo), _lib.ptr(seq), _lib.ptr(logprobs),
                                           _lib.ptr(d_seq), _lib.ptr(d_len), _lib.ptr(d_p), _lib.ptr(d_raw), _lib.current_stream())

    def _call_record(self, lib, image, rank, dst):
        return lib.capb200_aoa_beam_record_logprobs(self._engine, image, rank, _lib.ptr(dst), _lib.current_stream())


def setup(opt, numeric_mode=None):
    """Factory with the contract of captioning.models.setup (captioning/models/__init__.py:20-73) for the families on the
    B200 hot path."""
    name = opt.caption_model
    if name in ('topdown', 'updown'):
        return B200UpDownModel(opt, numeric_mode)
    if name == 'newfc':
        return B200NewFCModel(opt, numeric_mode)
    if name == 'aoa':
        return B200AoAModel(opt, numeric_mode)
    if name == 'transformer':
        if getattr(opt, 'cached_transformer', False):
            raise NotImplementedError('cachedTransformer is a reference-side variant; the B200 engine always caches K/V')
        return B200TransformerModel(opt, numeric_mode)
    raise NotImplementedError('caption_model %r is not on the B200 decode path yet (SURVEY.md section 8)' % name)
