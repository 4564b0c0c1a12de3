"""Oracle: fp32 policy/value ResNet forward on the CPU (TEST INFRASTRUCTURE ONLY).

Functional restatement of ``AlphaZeroNet.forward`` + ``PyTorchModel.predict``
(network.py:85-117, 168-183) over a plain ``state_dict`` - no nn.Module, so the
same function checks weights coming from the reference, from this repository's
parameter container, or from a checkpoint file.  Eval-mode BatchNorm (running
statistics, eps 1e-5) as during search.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def n_blocks(sd) -> int:
    i = 0
    while f"res_blocks.{i}.conv1.weight" in sd:
        i += 1
    return i


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, BN_EPS)


def forward(sd, x: torch.Tensor, return_trunk: bool = False):
    """x float32[B,3,15,15] -> (logits[B,225], value[B,1]).  network.py:94-117."""
    sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items() if v.dtype.is_floating_point}
    h = F.relu(_bn(F.conv2d(x, sd["conv.weight"], padding=1), sd, "bn"))
    i = 0
    while f"res_blocks.{i}.conv1.weight" in sd:
        pre = f"res_blocks.{i}."
        t = F.relu(_bn(F.conv2d(h, sd[pre + "conv1.weight"], padding=1), sd, pre + "bn1"))
        t = _bn(F.conv2d(t, sd[pre + "conv2.weight"], padding=1), sd, pre + "bn2")
        h = F.relu(t + h)
        i += 1
    p = F.relu(_bn(F.conv2d(h, sd["policy_conv.weight"]), sd, "policy_bn"))
    logits = F.linear(p.reshape(p.shape[0], -1), sd["policy_fc.weight"], sd["policy_fc.bias"])
    v = F.relu(_bn(F.conv2d(h, sd["value_conv.weight"]), sd, "value_bn"))
    v = F.relu(F.linear(v.reshape(v.shape[0], -1), sd["value_fc1.weight"], sd["value_fc1.bias"]))
    value = torch.tanh(F.linear(v, sd["value_fc2.weight"], sd["value_fc2.bias"]))
    if return_trunk:
        return logits, value, h
    return logits, value


class CpuModel:
    """``nn_model`` protocol of the reference search: predict(X) -> (probs, values)
    (network.py:168-183): softmax over all 225 logits, float32 numpy out."""

    def __init__(self, sd):
        self.sd = sd
        self.rows = 0

    def predict(self, X: np.ndarray):
        with torch.no_grad():
            logits, value = forward(self.sd, torch.from_numpy(np.asarray(X, dtype=np.float32)))
            probs = F.softmax(logits, dim=1).numpy()
        self.rows += len(X)
        return probs, value.numpy()


def policy_kl(p_ref: np.ndarray, p: np.ndarray) -> np.ndarray:
    """KL(p_ref || p) per row, in float64, with the usual 0*log0 = 0."""
    a = np.asarray(p_ref, dtype=np.float64)
    b = np.maximum(np.asarray(p, dtype=np.float64), 1e-30)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(a > 0, a * (np.log(a) - np.log(b)), 0.0)
    return t.sum(axis=1)
