"""Python face of the azg_net_* leaf evaluator (tcgen05 trunk + fused heads).

Holds the packed weights and activation buffers of one network on one B200.
``load_state_dict`` takes the reference's state_dict (network.py:54-71 key names).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import azg_net_weights, check, lib, ptr

_BN = ("weight", "bias", "running_mean", "running_var")


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class NetEngine:
    def __init__(self, n_blocks: int, channels: int, device, max_batch: int = 4096):
        if not torch.cuda.is_available():
            raise _lib.AzgError("no CUDA device: azgomoku_b200 has no CPU fallback")
        self.device = torch.device(device)
        self.n_blocks, self.channels, self.max_batch = n_blocks, channels, max_batch
        h = C.c_void_p()
        check(lib.azg_net_create(self.device.index or 0, n_blocks, channels, max_batch, C.byref(h)))
        self._h = h
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            lib.azg_net_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def memory_bytes(self) -> int:
        return int(lib.azg_net_memory_bytes(self._h))

    def load_state_dict(self, sd) -> None:
        dev = self.device
        keep = {k: v.detach().to(dev, torch.float32).contiguous() for k, v in sd.items() if v.dtype.is_floating_point}
        w = azg_net_weights()
        a = lambda k: C.c_void_p(keep[k].data_ptr())
        w.conv_w = a("conv.weight")
        for j, f in enumerate(_BN):
            w.bn[j] = a(f"bn.{f}")
            w.policy_bn[j] = a(f"policy_bn.{f}")
            w.value_bn[j] = a(f"value_bn.{f}")
        for i in range(self.n_blocks):
            for t in (0, 1):
                w.res_conv_w[2 * i + t] = a(f"res_blocks.{i}.conv{t + 1}.weight")
                for j, f in enumerate(_BN):
                    w.res_bn[2 * i + t][j] = a(f"res_blocks.{i}.bn{t + 1}.{f}")
        w.policy_conv_w, w.policy_fc_w, w.policy_fc_b = a("policy_conv.weight"), a("policy_fc.weight"), a("policy_fc.bias")
        w.value_conv_w, w.value_fc1_w, w.value_fc1_b = a("value_conv.weight"), a("value_fc1.weight"), a("value_fc1.bias")
        w.value_fc2_w, w.value_fc2_b = a("value_fc2.weight"), a("value_fc2.bias")
        with torch.cuda.device(dev):
            check(lib.azg_net_load(self._h, C.byref(w), _stream()))
            torch.cuda.current_stream().synchronize()      # the fp32 sources may be freed after this

    def forward(self, planes: torch.Tensor, want_logits: bool = False):
        """planes f32[B,3,15,15] (device) -> probs f32[B,225], values f32[B,1] (and logits)."""
        x = planes.to(self.device, torch.float32).contiguous()
        n = x.shape[0]
        probs = torch.empty((n, 225), dtype=torch.float32, device=self.device)
        values = torch.empty((n, 1), dtype=torch.float32, device=self.device)
        logits = torch.empty((n, 225), dtype=torch.float32, device=self.device) if want_logits else None
        with torch.cuda.device(self.device):
            check(lib.azg_net_forward_planes(self._h, ptr(x), n, ptr(probs), ptr(values), ptr(logits), _stream()))
        return (probs, values, logits) if want_logits else (probs, values)

    def forward_leaves(self, engine, probs: torch.Tensor, values: torch.Tensor | None = None):
        """Evaluate ``engine``'s current leaf batch on the device (no host round trip)."""
        check(lib.azg_net_forward_leaves(self._h, engine._h, ptr(probs), ptr(values)))

    def trunk(self, planes: torch.Tensor, n_layers: int) -> torch.Tensor:
        """Test hook: activations after the stem + n_layers 3x3 layers, f32[B,C,15,15]."""
        x = planes.to(self.device, torch.float32).contiguous()
        out = torch.empty((x.shape[0], self.channels, 15, 15), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.azg_net_trunk_debug(self._h, ptr(x), x.shape[0], n_layers, ptr(out), _stream()))
        return out

    def profile(self, enable: bool):
        check(lib.azg_net_profile(self._h, int(enable)))

    def profile_read(self):
        """-> (milliseconds spent in the 3x3 trunk, conv3x3 launches covered) since the last read."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        check(lib.azg_net_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_counters(self) -> dict:
        out = (C.c_uint64 * 32)()
        check(lib.azg_net_profile_counters(self._h, out))
        keys = ("mma_wait_full", "mma_wait_tmem_empty", "mma_total", "producer_wait_empty", "producer_total",
                "epilogue_wait_tmem_full", "epilogue_total", "boards", "epi_wait_store_drain", "epi_residual_transpose",
                "epi_tmem_ld_wait", "epi_compute_stage", "epi_fence_store")
        res = {k: int(out[i]) + int(out[16 + i]) for i, k in enumerate(keys)}            # all layers
        for tag, off in (("plain", 0), ("residual", 16)):                               # and split by layer type
            res[tag] = {k: int(out[off + i]) for i, k in enumerate(keys)}
        return res

    def check(self):
        check(lib.azg_net_check(self._h, _stream()))
