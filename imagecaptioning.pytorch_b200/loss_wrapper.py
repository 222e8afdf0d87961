"""Mirror of captioning/modules/loss_wrapper.py (LossWrapper) and losses.RewardCriterion for the SCST branch.

    B200LossWrapper(model, opt).forward(fc_feats, att_feats, labels, masks, att_masks, gts, gt_indices,
                                        sc_flag, struc_flag, drop_worst_flag) -> {'loss', 'reward'}

sc_flag=True follows loss_wrapper.py:56-73 exactly: eval-mode greedy baseline, train-mode multinomial samples, CIDEr-D
self-critical reward and RewardCriterion -- every stage on the device through the C ABI.  The XE branch (sc_flag=False)
and the structure-loss branch are the next rows of SURVEY.md section 8(f) and raise for now.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .rewards import get_self_critical_reward


class RewardCriterion(nn.Module):
    """losses.py:18-37 on the device: -logp[seq] * reward * mask, mask = tokens up to and including the first EOS."""

    def forward(self, input, seq, reward, reduction='mean'):
        N, L = seq.shape
        V1 = input.shape[2]
        lp = input.detach().to(torch.float32).contiguous()
        sq = seq.detach().to(torch.long).contiguous()
        rw = reward.detach().to(torch.float32).contiguous()
        dev = lp.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        rows = torch.empty(N, dtype=torch.float32, device=dev)
        msum = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.check(_lib.load().capb200_reward_criterion_forward(_lib.ptr(lp), _lib.ptr(sq), _lib.ptr(rw), N, L, V1, _lib.ptr(loss), _lib.ptr(rows),
                                                                _lib.ptr(msum), _lib.current_stream()), 'reward_criterion_forward')
        self.mask_sum = msum
        return rows if reduction == 'none' else loss[0]

    def backward_logprobs(self, seq, reward, V1, upstream=1.0):
        """d loss_mean / d logprobs as a dense [N, L, V1] tensor (what autograd hands to the reference's log-softmax)."""
        N, L = seq.shape
        grad = torch.zeros(N, L, V1, dtype=torch.float32, device=seq.device)
        sq = seq.detach().to(torch.long).contiguous()
        rw = reward.detach().to(torch.float32).contiguous()
        _lib.check(_lib.load().capb200_reward_criterion_backward(_lib.ptr(sq), _lib.ptr(rw), N, L, V1, _lib.ptr(self.mask_sum), float(upstream),
                                                                 _lib.ptr(grad), _lib.current_stream()), 'reward_criterion_backward')
        return grad


class _ScstLoss(torch.autograd.Function):
    """Connects the engine-computed loss to the parameters: the gradients were produced by the engine's own BPTT during the forward
    call; backward hands them (scaled by the upstream gradient) to autograd, so loss.backward(), DDP hooks, clip_grad_value_ and the
    optimizers of tools/train.py:189-196 work unchanged."""

    @staticmethod
    def forward(ctx, loss_value, grad_list, *params):
        ctx.grad_list = grad_list
        return loss_value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (None, None) + tuple(g * grad_out for g in ctx.grad_list)


class B200LossWrapper(nn.Module):
    def __init__(self, model, opt):
        super().__init__()
        self.opt = opt
        self.model = model
        self.rl_crit = RewardCriterion()

    def forward(self, fc_feats, att_feats, labels, masks, att_masks, gts, gt_indices, sc_flag, struc_flag, drop_worst_flag):
        opt = self.opt
        out = {}
        reduction = 'none' if drop_worst_flag else 'mean'
        if struc_flag:
            raise NotImplementedError('structure losses are the next row of SURVEY.md section 8(f)')
        if not sc_flag:
            raise NotImplementedError('the XE stage is the next row of SURVEY.md section 8(f)')
        fused = (hasattr(self.model, 'scst_step') and att_masks is None and not drop_worst_flag and torch.is_grad_enabled() and
                 opt.sc_sample_method == 'greedy' and opt.sc_beam_size == 1 and opt.train_sample_method == 'sample' and opt.train_beam_size == 1 and
                 getattr(opt, 'bleu_reward_weight', 0) == 0 and getattr(opt, 'cider_reward_weight', 1) == 1)
        if fused:
            # whole step on the device incl. back-propagation through time (UpDown); dropout as in model.train()
            from . import rewards as _rw
            if _rw.CiderD_scorer is None:
                raise RuntimeError('init_scorer(cached_tokens) must be called before the SCST reward (tools/train.py:150-152)')
            self.model.train()
            gts = [gts[_] for _ in gt_indices.tolist()]
            res = self.model.scst_step(fc_feats, att_feats, gts, _rw.CiderD_scorer, opt.train_sample_n, temperature=getattr(opt, 'temperature', 1.0))
            params = list(res['grads'].keys())
            out['loss'] = _ScstLoss.apply(res['loss'], [res['grads'][p_] for p_ in params], *params)
            out['reward'] = res['reward'][:, 0].mean()
            self.last_step = res
            return out
        self.model.eval()
        with torch.no_grad():
            greedy_res, _ = self.model(fc_feats, att_feats, att_masks, mode='sample',
                                       opt={'sample_method': opt.sc_sample_method, 'beam_size': opt.sc_beam_size})
        self.model.train()
        gen_result, sample_logprobs = self.model(fc_feats, att_feats, att_masks,
                                                 opt={'sample_method': opt.train_sample_method, 'beam_size': opt.train_beam_size,
                                                      'sample_n': opt.train_sample_n}, mode='sample')
        gts = [gts[_] for _ in gt_indices.tolist()]
        reward = get_self_critical_reward(greedy_res, gts, gen_result, opt)
        loss = self.rl_crit(sample_logprobs, gen_result, reward, reduction=reduction)
        out['reward'] = reward[:, 0].mean()
        out['loss'] = loss
        return out
