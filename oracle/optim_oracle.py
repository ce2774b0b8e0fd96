"""TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's checker legs; never by tubedetr_b200/).

CPU restatement of the reference's optimizer-side step, one tensor at a time, in plain fp32 torch arithmetic:
  * gradient clipping   -- engine.py:149-150 -> torch.nn.utils.clip_grad_norm_ (L2 norm of the per-tensor L2 norms,
                           coef = max_norm / (norm + 1e-6) clamped to <= 1, every gradient scaled by it)
  * AdamW               -- main.py:410-414 -> torch.optim.AdamW (decoupled weight decay, bias correction, eps added after
                           the bias-corrected sqrt; default betas (0.9, 0.999), eps 1e-8), three LR groups main.py:381-405
  * EMA                 -- util/optim.py:8-25: w_ema = w_ema * decay + (1 - decay) * w over the state_dict
Parity is PINNED: tests/golden/optim.pt was produced by tests/golden/make_optim_golden.py running the reference's own call
sequence (torch's clip_grad_norm_ and AdamW + the reference's util/optim.update_ema imported from /root/reference);
tests/test_optim_oracle.py checks this restatement against it.
"""
import math

import torch


def clip_coef(grads, max_norm):
    norms = [g.float().norm(2) for g in grads if g is not None]
    total = torch.stack(norms).norm(2)
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adamw_ema_step(params, grads, exp_avg, exp_avg_sq, emas, lrs, weight_decay, step, betas=(0.9, 0.999), eps=1e-8,
                   max_norm=0.1, ema_decay=0.9998):
    """in place on the lists of tensors; lrs: one learning rate per tensor; step counts from 1.  Returns the grad norm."""
    b1, b2 = betas
    total, coef = clip_coef(grads, max_norm) if max_norm and max_norm > 0 else (None, 1.0)
    bc1 = 1 - b1 ** step
    bc2_sqrt = math.sqrt(1 - b2 ** step)
    for i, (p, g) in enumerate(zip(params, grads)):
        if g is None:
            continue
        g = g * coef
        p.mul_(1 - lrs[i] * weight_decay)
        exp_avg[i].lerp_(g, 1 - b1)
        exp_avg_sq[i].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (exp_avg_sq[i].sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(exp_avg[i], denom, value=-(lrs[i] / bc1))
    if emas is not None:
        for p, e in zip(params, emas):
            e.copy_(e * ema_decay + (1.0 - ema_decay) * p)
    return total
