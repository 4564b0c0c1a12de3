"""Oracle: one ``train_batch`` step in fp32 on the CPU (TEST INFRASTRUCTURE ONLY).

Functional restatement of ``PyTorchModel.train_batch`` (network.py:199-235) over a plain
``state_dict``: training-mode forward (BatchNorm with batch statistics, eps 1e-5, running
statistics updated with momentum 0.1 and the unbiased variance as torch does), loss =
``KLDivLoss(batchmean)(log_softmax(logits), pi) + MSELoss(value, z)`` (network.py:143-144,
217-222), gradients (torch autograd on the functional graph), ``clip_grad_norm_(3.0)`` over all
parameters (network.py:224), and ``torch.optim.Adam(lr, weight_decay)`` written out (weight decay
added to the gradient, bias-corrected moments, eps 1e-8; network.py:141).

Pinned against the reference by ``oracle/make_golden.py`` (``tests/golden/train_steps.npz`` and the
checkpoint the reference saved after three steps) in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
CLIP = 3.0


def param_names(sd):
    """Parameter keys in ``net.parameters()`` order (registration order, network.py:47-73)."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


class _Bf16RoundTrip(torch.autograd.Function):
    """Round to bfloat16 and back in the forward AND in the backward pass: the storage format of the CUDA step's
    activations (z, a) and activation gradients (dz, g).  Used by the "emulated" oracle, which separates what
    bf16 storage does to a gradient from what a kernel bug would do."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def _bn_train(x, sd, name, stats):
    y = F.batch_norm(x, None, None, sd[name + ".weight"], sd[name + ".bias"], True, 0.0, BN_EPS)
    with torch.no_grad():
        dims = [0] + list(range(2, x.dim()))
        n = x.numel() // x.shape[1]
        stats[name] = (x.mean(dims), x.var(dims, unbiased=False), n)
    return y


def forward_train(sd, x, stats, record=None, emulate_bf16=False):
    """Training-mode forward (network.py:94-117 with module.training = True).  ``record`` (a list) receives the
    activation after every 3x3 layer's BatchNorm + ReLU (stem first), for layer-by-layer checks.
    ``emulate_bf16``: store what the CUDA step stores in bfloat16 (trunk conv weights, conv outputs, activations,
    and - in the backward pass - their gradients); everything else stays fp32 as in the CUDA step."""
    keep = (lambda t: record.append(t)) if record is not None else (lambda t: None)
    q = _Bf16RoundTrip.apply if emulate_bf16 else (lambda t: t)
    qw = (lambda w: w + (w.to(torch.bfloat16).to(torch.float32) - w).detach()) if emulate_bf16 else (lambda w: w)
    h = q(F.relu(_bn_train(q(F.conv2d(x, sd["conv.weight"], padding=1)), sd, "bn", stats)))
    keep(h)
    i = 0
    while f"res_blocks.{i}.conv1.weight" in sd:
        pre = f"res_blocks.{i}."
        t = q(F.relu(_bn_train(q(F.conv2d(h, qw(sd[pre + "conv1.weight"]), padding=1)), sd, pre + "bn1", stats)))
        keep(t)
        t = _bn_train(q(F.conv2d(t, qw(sd[pre + "conv2.weight"]), padding=1)), sd, pre + "bn2", stats)
        h = q(F.relu(t + h))
        keep(h)
        i += 1
    p = F.relu(_bn_train(F.conv2d(h, sd["policy_conv.weight"]), sd, "policy_bn", stats))
    logits = F.linear(p.reshape(p.shape[0], -1), sd["policy_fc.weight"], sd["policy_fc.bias"])
    v = F.relu(_bn_train(F.conv2d(h, sd["value_conv.weight"]), sd, "value_bn", stats))
    v = F.relu(F.linear(v.reshape(v.shape[0], -1), sd["value_fc1.weight"], sd["value_fc1.bias"]))
    return logits, torch.tanh(F.linear(v, sd["value_fc2.weight"], sd["value_fc2.bias"]))


def losses(logits, value, pi, z):
    """network.py:217-222.  KLDivLoss(batchmean): sum(pi * (log pi - log_softmax)) / B with 0 log 0 = 0."""
    logp = F.log_softmax(logits, dim=1)
    kl = torch.where(pi > 0, pi * (torch.log(pi.clamp_min(1e-45)) - logp), torch.zeros_like(pi)).sum() / pi.shape[0]
    mse = ((value - z) ** 2).mean()
    return kl, mse


def gradients(sd, x, pi, z, record=None, emulate_bf16=False):
    """-> (policy_loss, value_loss, {name: grad}, bn batch statistics) without touching ``sd``."""
    names = param_names(sd)
    work = {k: (v.detach().clone().to(torch.float32).requires_grad_(k in names) if v.dtype.is_floating_point else v) for k, v in sd.items()}
    stats = {}
    logits, value = forward_train(work, x, stats, record, emulate_bf16)
    kl, mse = losses(logits, value, pi, z)
    (kl + mse).backward()
    return float(kl.detach()), float(mse.detach()), {k: work[k].grad.detach() for k in names}, stats


class Adam:
    """torch.optim.Adam(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4), written out."""

    def __init__(self, sd, lr=1e-3, weight_decay=1e-4, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.wd, self.b1, self.b2, self.eps = lr, weight_decay, b1, b2, eps
        self.t = 0
        self.m = {k: torch.zeros_like(sd[k]) for k in param_names(sd)}
        self.v = {k: torch.zeros_like(sd[k]) for k in param_names(sd)}

    def step(self, sd, grads):
        self.t += 1
        c1, c2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
        for k, g in grads.items():
            g = g + self.wd * sd[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (self.v[k].sqrt() / math.sqrt(c2)).add_(self.eps)
            sd[k].addcdiv_(self.m[k], denom, value=-self.lr / c1)


def clip(grads, max_norm=CLIP):
    """torch.nn.utils.clip_grad_norm_: scale by max_norm / (total_norm + 1e-6), clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in grads.items()}, float(total)


def train_step(sd, opt: Adam, x, pi, z):
    """One epoch of network.py:210-226 on ``sd`` in place.  Returns the reference's loss dictionary."""
    x, pi, z = (torch.as_tensor(a, dtype=torch.float32) for a in (x, pi, z))
    kl, mse, grads, stats = gradients(sd, x, pi, z)
    grads, _ = clip(grads)
    opt.step(sd, grads)
    for name, (mean, var, n) in stats.items():           # running statistics as nn.BatchNorm2d updates them
        sd[name + ".running_mean"].mul_(1 - BN_MOMENTUM).add_(mean, alpha=BN_MOMENTUM)
        sd[name + ".running_var"].mul_(1 - BN_MOMENTUM).add_(var * (n / max(n - 1, 1)), alpha=BN_MOMENTUM)
        sd[name + ".num_batches_tracked"] += 1
    return {"policy_loss": kl, "value_loss": mse, "total_loss": kl + mse}
