"""Host-side mirror of the evaluation loop around the decode hot path: captioning/utils/eval_utils.py:129-213 (``eval_split``) and the
feature hand-off of captioning/data/dataloader.py:317-332 (``get_batch``).

SURVEY.md section 8(f) rank 4: once the decode itself is fast, the reference's evaluation loop is dominated by host work --
``tensor.to(device)`` from pageable memory on the critical path, one ``.item()`` (a device synchronisation) per caption for the
perplexity and the entropy, one per token in ``decode_sequence``.  This module keeps the reference's contract

    eval_split(model, crit, loader, eval_kwargs) -> (mean loss, predictions, lang_stats)

and its ``predictions`` entries (``image_id``, ``caption``, ``perplexity``, ``entropy``) but

  * ``PrefetchLoader`` pins every batch's tensors and copies batch i+1 host -> device on a side stream while batch i decodes,
  * perplexity / entropy are reduced on the device for the whole batch and fetched with ONE device -> host copy,
  * captions are detokenised by ``utils.decode_sequence`` (one copy per batch).

Language evaluation (``lang_eval=1``: the Java METEOR / SPICE tool chain of coco-caption) is outside the hot path;
pass ``eval_kwargs['language_eval']`` = a callable ``(dataset, predictions, n_predictions, eval_kwargs, split) -> stats`` to plug the
reference's ``eval_utils.language_eval`` in.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Iterator, Optional, Tuple

import torch

from .utils import decode_sequence

_TENSOR_KEYS = ('fc_feats', 'att_feats', 'labels', 'masks', 'att_masks')


class PrefetchLoader:
    """Iterates ``loader.get_batch(split)`` (the reference's loader API, dataloader.py:317-332) one batch ahead: the tensors of batch i+1
    are pinned and copied to ``device`` on a side stream while the caller still works on batch i.  Yields the loader's dict with the five
    tensor entries replaced by device tensors (``None`` stays ``None``).  On a CPU ``device`` it degrades to a plain pass-through."""

    def __init__(self, loader, split: str, device='cuda', max_batches: Optional[int] = None):
        self.loader, self.split, self.max_batches = loader, split, max_batches
        self.device = torch.device(device)
        self.cuda = self.device.type == 'cuda'
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None

    def _stage(self, data: Dict[str, Any]) -> Tuple[Dict[str, Any], Optional[torch.cuda.Event]]:
        if not self.cuda:
            return data, None
        out = dict(data)
        with torch.cuda.stream(self.stream):
            for k in _TENSOR_KEYS:
                t = data.get(k)
                if t is None:
                    continue
                t = torch.as_tensor(t)
                if not t.is_cuda:
                    t = t.pin_memory() if not t.is_pinned() else t
                out[k] = t.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self) -> Iterator[Dict[str, Any]]:
        n = 0
        nxt = self._stage(self.loader.get_batch(self.split))
        while nxt is not None:
            data, ev = nxt
            n += 1
            wrapped = bool(data.get('bounds', {}).get('wrapped', False))
            more = not wrapped and (self.max_batches is None or n < self.max_batches)
            nxt = self._stage(self.loader.get_batch(self.split)) if more else None      # the copy of batch i+1 overlaps the work on batch i
            if ev is not None:
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                for k in _TENSOR_KEYS:
                    if isinstance(data.get(k), torch.Tensor):
                        data[k].record_stream(cur)
            yield data


def caption_stats(seq: torch.Tensor, seq_logprobs: torch.Tensor):
    """(perplexity [N], entropy [N]) exactly as eval_utils.py:173-174, as device tensors (no host synchronisation)."""
    denom = (seq > 0).to(seq_logprobs).sum(1) + 1
    entropy = -(torch.softmax(seq_logprobs, dim=2) * seq_logprobs).sum(2).sum(1) / denom
    perplexity = -seq_logprobs.gather(2, seq.unsqueeze(2)).squeeze(2).sum(1) / denom
    return perplexity, entropy


def eval_split_n(model, n_predictions, input_data, eval_kwargs: Dict[str, Any] = {}):
    """Contract of captioning/utils/eval_utils.py:230-283: ``sample_n`` captions per image, appended to ``n_predictions``.
    ``sample_n_method`` 'bs' (the sample_n best beams), 'sample' / 'gumbel' / 'top<k>' / 'top<p>' (sample_n draws, with their perplexity:
    read back with ONE transfer for the batch instead of one .item() per caption).  'dbs' and the remaining branch are diverse beam
    search (group_size > 1), which the engine refuses; the model's own NotImplementedError surfaces."""
    verbose = eval_kwargs.get('verbose', True)
    beam_size = eval_kwargs.get('beam_size', 1)
    sample_n = eval_kwargs.get('sample_n', 1)
    sample_n_method = eval_kwargs.get('sample_n_method', 'sample')
    fc_feats, att_feats, att_masks, data = input_data
    kw = dict(eval_kwargs)
    n_images = fc_feats.shape[0]
    if sample_n_method == 'bs':
        kw.update({'sample_n': 1, 'beam_size': sample_n, 'group_size': 1})
        with torch.no_grad():
            model(fc_feats, att_feats, att_masks, opt=kw, mode='sample')
        for k in range(n_images):
            sents = decode_sequence(model.vocab, torch.stack([model.done_beams[k][_]['seq'] for _ in range(sample_n)]))
            for sent in sents:
                n_predictions.append({'image_id': data['infos'][k]['id'], 'caption': sent})
    elif sample_n_method in ('sample', 'gumbel') or sample_n_method.startswith('top'):
        kw.update({'sample_n': sample_n, 'sample_method': sample_n_method, 'beam_size': 1})
        with torch.no_grad():
            seq, logprobs = model(fc_feats, att_feats, att_masks, opt=kw, mode='sample')
        perplexity = (-logprobs.gather(2, seq.unsqueeze(2)).squeeze(2).sum(1) / ((seq > 0).to(logprobs).sum(1) + 1)).cpu().tolist()
        for k, sent in enumerate(decode_sequence(model.vocab, seq)):
            n_predictions.append({'image_id': data['infos'][k // sample_n]['id'], 'caption': sent, 'perplexity': perplexity[k]})
    elif sample_n_method == 'dbs':
        kw.update({'beam_size': sample_n * beam_size, 'group_size': sample_n})
        with torch.no_grad():
            model(fc_feats, att_feats, att_masks, opt=kw, mode='sample')
        for k in range(n_images):
            sents = decode_sequence(model.vocab, torch.stack([model.done_beams[k][_]['seq'] for _ in range(0, sample_n * beam_size, beam_size)]))
            for sent in sents:
                n_predictions.append({'image_id': data['infos'][k]['id'], 'caption': sent})
    else:
        kw.update({'sample_method': sample_n_method[1:], 'group_size': sample_n, 'beam_size': 1})
        with torch.no_grad():
            seq, _ = model(fc_feats, att_feats, att_masks, opt=kw, mode='sample')
        for k, sent in enumerate(decode_sequence(model.vocab, seq)):
            n_predictions.append({'image_id': data['infos'][k // sample_n]['id'], 'caption': sent})
    if verbose:
        for entry in sorted(n_predictions[-n_images * sample_n:], key=lambda x: x['image_id']):
            print('image %s: %s' % (entry['image_id'], entry['caption']))


def eval_split(model, crit, loader, eval_kwargs: Dict[str, Any] = {}):
    """Contract of captioning/utils/eval_utils.py:129-213.  ``crit`` is the XE criterion (LanguageModelCriterion / LabelSmoothing)."""
    verbose = eval_kwargs.get('verbose', True)
    verbose_beam = eval_kwargs.get('verbose_beam', 0)
    verbose_loss = eval_kwargs.get('verbose_loss', 1)
    num_images = eval_kwargs.get('num_images', eval_kwargs.get('val_images_use', -1))
    split = eval_kwargs.get('split', 'val')
    lang_eval = eval_kwargs.get('language_eval', 0)
    dataset = eval_kwargs.get('dataset', 'coco')
    beam_size = eval_kwargs.get('beam_size', 1)
    sample_n = eval_kwargs.get('sample_n', 1)
    remove_bad_endings = eval_kwargs.get('remove_bad_endings', 0)
    os.environ['REMOVE_BAD_ENDINGS'] = str(remove_bad_endings)      # same global configuration channel as the reference (eval_utils.py:139)
    device = eval_kwargs.get('device', 'cuda')
    model.eval()
    loader.reset_iterator(split)
    n, loss, loss_sum, loss_evals = 0, 0.0, 0.0, 1e-8
    predictions, n_predictions = [], []
    for data in PrefetchLoader(loader, split, device):
        n += len(data['infos'])
        fc_feats, att_feats, labels, masks, att_masks = (data[k] for k in _TENSOR_KEYS)
        loss_t = None
        if labels is not None and verbose_loss:
            with torch.no_grad():
                loss_t = crit(model(fc_feats, att_feats, labels[..., :-1], att_masks), labels[..., 1:], masks[..., 1:])
        with torch.no_grad():
            kw = dict(eval_kwargs)
            kw.update({'sample_n': 1})
            seq, seq_logprobs = model(fc_feats, att_feats, att_masks, opt=kw, mode='sample')
            seq = seq.data
            perplexity, entropy = caption_stats(seq, seq_logprobs)
        # one device -> host transfer for the batch's scalars (the reference: 2 .item() per caption + 1 for the loss)
        scal = torch.stack([perplexity, entropy]).double()
        if loss_t is not None:
            scal = torch.cat([scal.reshape(-1), loss_t.reshape(1).double()])
        scal = scal.reshape(-1).cpu().tolist()
        N = seq.shape[0]
        perp, ent = scal[:N], scal[N:2 * N]
        if loss_t is not None:
            loss = scal[2 * N]
            loss_sum += loss
            loss_evals += 1
        if beam_size > 1 and verbose_beam:
            for i in range(fc_feats.shape[0]):
                print('\n'.join([decode_sequence(model.vocab, _['seq'].unsqueeze(0))[0] for _ in model.done_beams[i]]))
                print('--' * 10)
        sents = decode_sequence(model.vocab, seq)
        for k, sent in enumerate(sents):
            entry = {'image_id': data['infos'][k]['id'], 'caption': sent, 'perplexity': perp[k], 'entropy': ent[k]}
            if eval_kwargs.get('dump_path', 0) == 1:
                entry['file_name'] = data['infos'][k]['file_path']
            predictions.append(entry)
            if verbose:
                print('image %s: %s' % (entry['image_id'], entry['caption']))
        if sample_n > 1:
            eval_split_n(model, n_predictions, [fc_feats, att_feats, att_masks, data], eval_kwargs)
        ix1 = data['bounds']['it_max']
        if num_images != -1:
            ix1 = min(ix1, num_images)
        else:
            num_images = ix1
        for _ in range(n - ix1):
            predictions.pop()
        if verbose:
            print('evaluating validation preformance... %d/%d (%f)' % (n, ix1, loss))
        if num_images >= 0 and n >= num_images:
            break

    lang_stats = None
    if len(n_predictions) > 0 and 'perplexity' in n_predictions[0]:
        n_predictions = sorted(n_predictions, key=lambda x: x['perplexity'])
    if 'id' in eval_kwargs:         # the reference's side effect (eval_utils.py:217-219): language_eval and tools/eval.py read this file back
        os.makedirs('eval_results', exist_ok=True)
        torch.save((predictions, n_predictions), os.path.join('eval_results/', '.saved_pred_' + eval_kwargs['id'] + '_' + split + '.pth'))
    if callable(lang_eval):
        lang_stats = lang_eval(dataset, predictions, n_predictions, eval_kwargs, split)
    elif lang_eval == 1:
        raise NotImplementedError("language evaluation runs the reference's Java tool chain: pass eval_kwargs['language_eval'] = eval_utils.language_eval")
    model.train()
    return loss_sum / loss_evals, predictions, lang_stats
