"""Drop-in ``PyTorchModel`` / ``AlphaZeroNet`` whose inference runs in the sm_100a CUDA library.

Mirrors network.py of the reference (network.py:29-73 module layout and initialisation,
:132-264 wrapper): the ``state_dict`` keys, constructor arguments, ``predict`` /
``predict_batch`` / ``train_batch`` / ``save`` / ``load`` signatures and the checkpoint
dictionary (``net``, ``opt``, ``board_size``, ``action_size``) are the same, so reference
snapshots load here and vice versa.

``predict`` (the search's leaf evaluator, network.py:168-183) is the hot path: it goes
through ``azg_net_forward`` - folded BatchNorm, bf16 tensor-core trunk, fused heads - and
raises if the CUDA library or a B200 is missing (no eager PyTorch fallback).
``train_batch`` (network.py:199-235) runs in the CUDA library too: tensor-core
forward / input-gradient / weight-gradient convolutions, training-mode BatchNorm, the reference's
KLDiv + MSE loss, clip 3.0 and Adam on fp32 master weights (``trainer.TrainEngine``).  The torch
autograd formulation of the same step is kept as ``train_batch_autograd`` (other channel counts than
64 / 128 / 256, and the cross-check in the tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def _conv3x3(c_in: int, c_out: int) -> nn.Conv2d:
    return nn.Conv2d(c_in, c_out, kernel_size=3, padding=1, bias=False)


def _conv1x1(c_in: int, c_out: int) -> nn.Conv2d:
    return nn.Conv2d(c_in, c_out, kernel_size=1, bias=False)


class ResidualBlock(nn.Module):
    """conv3x3 - BN - ReLU - conv3x3 - BN, skip connection, ReLU (network.py:9-26).  Attribute names and
    creation order are part of the checkpoint format and of the seed -> weights mapping."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv1, self.bn1 = _conv3x3(channels, channels), nn.BatchNorm2d(channels)
        self.conv2, self.bn2 = _conv3x3(channels, channels), nn.BatchNorm2d(channels)

    def forward(self, x):
        inner = F.relu(self.bn1(self.conv1(x)))
        return F.relu(self.bn2(self.conv2(inner)) + x)


class AlphaZeroNet(nn.Module):
    """Parameter container of the policy/value ResNet with the reference's module names, creation order
    and initialisation (network.py:41-83): the same seed yields the same weights and checkpoints are
    interchangeable.  ``forward`` is the autograd path used for training; inference runs in the CUDA
    library (``PyTorchModel.predict``)."""

    def __init__(self, in_channels: int = 3, board_size: int = 15, action_size: int = 15 * 15,
                 n_res_blocks: int = 6, channels: int = 128):
        super().__init__()
        cells = board_size * board_size
        self.board_size, self.action_size, self.channels = board_size, action_size, channels
        # stem, tower
        self.conv, self.bn = _conv3x3(in_channels, channels), nn.BatchNorm2d(channels)
        self.res_blocks = nn.ModuleList(ResidualBlock(channels) for _ in range(n_res_blocks))
        # policy head: 1x1 conv to two planes, then a dense layer over ch*cells + pixel
        self.policy_conv, self.policy_bn = _conv1x1(channels, 2), nn.BatchNorm2d(2)
        self.policy_fc = nn.Linear(2 * cells, action_size)
        # value head: 1x1 conv to one plane, 64 hidden units, tanh
        self.value_conv, self.value_bn = _conv1x1(channels, 1), nn.BatchNorm2d(1)
        self.value_fc1, self.value_fc2 = nn.Linear(cells, 64), nn.Linear(64, 1)
        self._reset_parameters()

    def _reset_parameters(self):
        """Kaiming-normal convolutions, Kaiming-uniform dense layers with zero bias, in module
        registration order (network.py:75-83) - the order fixes the random stream."""
        for layer in self.modules():
            if isinstance(layer, nn.Conv2d):
                nn.init.kaiming_normal_(layer.weight, nonlinearity="relu")
            elif isinstance(layer, nn.Linear):
                nn.init.kaiming_uniform_(layer.weight, nonlinearity="relu")
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, 0)

    def trunk(self, x: torch.Tensor) -> torch.Tensor:
        h = F.relu(self.bn(self.conv(x)))
        for block in self.res_blocks:
            h = block(h)
        return h

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        h = self.trunk(x)
        pol = F.relu(self.policy_bn(self.policy_conv(h))).flatten(1)
        val = F.relu(self.value_bn(self.value_conv(h))).flatten(1)
        val = F.relu(self.value_fc1(val))
        return self.policy_fc(pol), torch.tanh(self.value_fc2(val))

    def predict(self, state):
        """(logits, value) of ONE encoded state through the autograd module, as network.py:119-129.  The search
        never calls this (it uses ``PyTorchModel.predict``, which runs in the CUDA library)."""
        x = torch.as_tensor(state, dtype=torch.float32, device=next(self.parameters()).device).unsqueeze(0)
        return self(x)


def infer_architecture(state_dict) -> Tuple[int, int]:
    """(n_res_blocks, channels) of a reference-layout state_dict."""
    channels = int(state_dict["conv.weight"].shape[0])
    blocks = 0
    while f"res_blocks.{blocks}.conv1.weight" in state_dict:
        blocks += 1
    return blocks, channels


class PyTorchModel:
    def __init__(self, board_size: int = 15, action_size: Optional[int] = None, device: Optional[str] = None,
                 n_res_blocks: int = 3, channels: int = 64, lr: float = 1e-3, weight_decay: float = 1e-4):
        if board_size != 15:
            raise ValueError("azgomoku_b200 kernels are specialised for the 15x15 board")
        self.board_size = board_size
        self.action_size = action_size if action_size is not None else board_size * board_size
        if self.action_size != board_size * board_size:
            raise ValueError("azgomoku_b200 kernels are specialised for action_size == board_size**2 == 225")
        self.device = device or "cuda"
        if not str(self.device).startswith("cuda") or not torch.cuda.is_available():
            raise _lib.AzgError("PyTorchModel needs a CUDA device: azgomoku_b200 has no CPU fallback")
        self.net = AlphaZeroNet(3, board_size, self.action_size, n_res_blocks, channels).to(self.device)
        self.optimizer = torch.optim.Adam(self.net.parameters(), lr=lr, weight_decay=weight_decay)
        self.value_loss_fn = nn.MSELoss()
        self.policy_loss_fn = nn.KLDivLoss(reduction="batchmean")
        self._engine = None
        self._packed_version = None
        self._trainer = None
        self.train_graphs = os.environ.get("AZG_TRAIN_GRAPHS", "1") != "0"    # replay the training step from CUDA graphs

    # ------------------------------------------------------------------ CUDA inference engine
    def _ensure_engine(self):
        from .nn_engine import NetEngine
        if self._engine is None:
            self._engine = NetEngine(len(self.net.res_blocks), self.net.channels, torch.device(self.device))
        # _version catches in-place tensor ops; data_ptr catches `p.data = ...`; writes through `.data` views that
        # bypass both must call invalidate() (broadcast_model and load do)
        version = tuple((t._version, t.data_ptr()) for t in list(self.net.parameters()) + list(self.net.buffers()))
        if version != self._packed_version:
            self._engine.load_state_dict(self.net.state_dict())
            self._packed_version = version
        return self._engine

    def invalidate(self) -> None:
        """Forget the packed bf16 weights: the next predict / search repacks from ``self.net``.  Needed after
        writes the version counters cannot see (``tensor.data`` views, raw device copies)."""
        self._packed_version = None
        if getattr(self, "_trainer", None) is not None:         # the training engine keeps bf16 copies of the convolution weights too
            self._trainer._param_version = None                 # (collectives such as dist.broadcast do not bump tensor versions)

    def predict_device(self, planes: torch.Tensor, validate: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """planes float32[B,3,15,15] on the device -> (probs f32[B,225], values f32[B,1]) on the device.
        The CUDA stem consumes stone bitboards, so the input must be what ``get_encoded_state`` produces
        (gomoku.py:130-150): planes 0/1 binary and disjoint, plane 2 all ones.  Anything else raises instead
        of silently evaluating a different input (one fused device reduction; ``validate=False`` skips it)."""
        if validate and planes.numel():
            a, b = planes[:, 0], planes[:, 1]
            bad = (((a != 0) & (a != 1)) | ((b != 0) & (b != 1)) | ((a != 0) & (b != 0))).any() | (planes[:, 2] != 1).any()
            if bool(bad):
                raise ValueError("predict: input is not an encoded board state (planes 0/1 must be binary and disjoint, "
                                 "plane 2 all ones); the CUDA evaluator has no generic float stem")
        return self._ensure_engine().forward(planes)

    def predict(self, encoded_states: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """network.py:168-183: eval-mode forward + softmax over all 225 logits, numpy in / numpy out."""
        x = torch.from_numpy(np.ascontiguousarray(encoded_states, dtype=np.float32)).to(self.device)
        probs, values = self.predict_device(x)
        return probs.cpu().numpy(), values.cpu().numpy()

    def predict_batch(self, states_list: list) -> Tuple[np.ndarray, np.ndarray]:
        return self.predict(self.make_batch_from_states(states_list))

    # ------------------------------------------------------------------ training step (network.py:199-235)
    def _to_device(self, a) -> torch.Tensor:
        t = a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a, dtype=np.float32))
        return t.to(self.device, torch.float32)

    def losses(self, states, target_pis, target_vs):
        """(policy KL-divergence with batchmean reduction on log-softmax, value MSE) as in the reference."""
        logits, values = self.net(states)
        return self.policy_loss_fn(F.log_softmax(logits, dim=1), target_pis), self.value_loss_fn(values, target_vs)

    def _ensure_trainer(self, batch: int):
        """The CUDA training engine for batches of up to ``batch`` positions (None for channel counts other than 64 / 128 / 256)."""
        if self.net.channels not in (64, 128, 256):
            return None
        if self._trainer is None or self._trainer.max_batch < batch:
            from .trainer import TrainEngine
            if self._trainer is not None:
                self._trainer.close()
            self._trainer = TrainEngine(self.net, self.optimizer, max(int(batch), 32), self.device)
        return self._trainer

    def train_batch_async(self, states, target_pis, target_vs, world: int = 1, reduce_grads=None) -> torch.Tensor:
        """One step of network.py:210-226 without a host synchronisation; returns the device tensor
        [policy_loss, value_loss].  ``reduce_grads(flat_grads)`` (data-parallel training) is called between
        the backward pass and the update and must leave the SUM over ``world`` ranks in the tensor."""
        x, pi, z = self._to_device(states), self._to_device(target_pis), self._to_device(target_vs)
        tr = self._ensure_trainer(x.shape[0])
        if tr is None:
            raise _lib.AzgError("the CUDA training step covers 64, 128 and 256 channels; use train_batch_autograd")
        self.net.train()
        if self.train_graphs:
            losses = tr.step_graph(x, pi, z, world, reduce_grads)
        else:
            losses = tr.forward_backward(x, pi, z)
            if reduce_grads is not None:
                reduce_grads(tr.flat_grads)
            tr.apply(world)
        self._packed_version = None     # the kernels wrote the parameters: the inference engine must repack (the trainer's own copies are current)
        return losses

    def train_batch(self, states, target_pis, target_vs, epochs: int = 1) -> dict:
        """``epochs`` Adam steps on one batch: loss = KL + MSE, gradient norm clipped at 3.0 (network.py:199-235).
        Accepts numpy arrays (as the reference) or device tensors (no host copy)."""
        if self.net.channels not in (64, 128, 256):
            return self.train_batch_autograd(states, target_pis, target_vs, epochs)
        x, pi, z = self._to_device(states), self._to_device(target_pis), self._to_device(target_vs)
        sums = torch.zeros(2, dtype=torch.float32, device=x.device)
        for _ in range(epochs):
            sums += self.train_batch_async(x, pi, z)
        p, v = (sums / float(epochs)).tolist()          # the only host synchronisation of the call
        return {"policy_loss": p, "value_loss": v, "total_loss": p + v}

    def train_batch_autograd(self, states, target_pis, target_vs, epochs: int = 1) -> dict:
        """The same step through torch autograd (library kernels)."""
        self.net.train()
        x, pi, z = self._to_device(states), self._to_device(target_pis), self._to_device(target_vs)
        sums = np.zeros(3)
        for _ in range(epochs):
            self.optimizer.zero_grad()
            policy_loss, value_loss = self.losses(x, pi, z)
            total = policy_loss + value_loss
            total.backward()
            torch.nn.utils.clip_grad_norm_(self.net.parameters(), 3.0)
            self.optimizer.step()
            sums += (float(policy_loss.item()), float(value_loss.item()), float(total.item()))
        self.invalidate()
        p, v, t = (sums / float(epochs)).tolist()
        return {"policy_loss": p, "value_loss": v, "total_loss": t}

    # ------------------------------------------------------------------ checkpoints (network.py:240-258)
    def checkpoint(self) -> dict:
        """The reference's checkpoint dictionary (depth and width are not stored there either)."""
        return {"net": self.net.state_dict(), "opt": self.optimizer.state_dict(),
                "board_size": self.board_size, "action_size": self.action_size}

    def save(self, path: str) -> None:
        folder = os.path.dirname(path)
        if folder:
            os.makedirs(folder, exist_ok=True)
        torch.save(self.checkpoint(), path)

    def load(self, path: str, map_location: Optional[str] = None) -> None:
        state = torch.load(path, map_location=map_location or self.device)
        self.net.load_state_dict(state["net"])
        self.invalidate()
        opt = state.get("opt")
        if opt is not None:
            try:                                   # optimiser state of another architecture is skipped, as in the reference
                self.optimizer.load_state_dict(opt)
            except Exception:
                pass

    @classmethod
    def from_checkpoint(cls, path: str, device: Optional[str] = None, **kw) -> "PyTorchModel":
        """Build a model whose depth and width are read off the checkpoint's state_dict (the reference's
        checkpoints do not store n_res_blocks / channels, network.py:240-248) and load it."""
        state = torch.load(path, map_location="cpu")
        blocks, channels = infer_architecture(state["net"])
        model = cls(board_size=int(state.get("board_size", 15)), action_size=state.get("action_size"), device=device,
                    n_res_blocks=blocks, channels=channels, **kw)
        model.load(path)
        return model

    @staticmethod
    def make_batch_from_states(list_of_encoded_states: list) -> np.ndarray:
        return np.stack(list_of_encoded_states, axis=0).astype(np.float32)
