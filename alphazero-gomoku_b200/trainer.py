"""Python face of the azg_train_* training step (tensor-core ``train_batch``).

``TrainEngine`` re-homes the parameters of an ``AlphaZeroNet`` into ONE flat fp32 vector (every
``nn.Parameter`` becomes a view into it, in ``net.parameters()`` order), keeps the gradients and the
Adam moments in flat vectors of the same shape, and runs ``PyTorchModel.train_batch``
(network.py:199-235 of the reference) in the CUDA library: training-mode forward, KLDiv + MSE loss,
backward, ``clip_grad_norm_(3.0)`` and Adam.  Because parameters and optimiser state stay ordinary torch
tensors (views), ``state_dict()`` / ``optimizer.state_dict()`` and with them the reference's checkpoint
format are unchanged, and the data-parallel gradient exchange is one all-reduce over ``flat_grads``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import azg_net_weights, azg_train_config, check, lib, ptr


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class TrainEngine:
    def __init__(self, net, optimizer: torch.optim.Optimizer, max_batch: int, device):
        self.device = torch.device(device)
        if not torch.cuda.is_available():
            raise _lib.AzgError("no CUDA device: azgomoku_b200 has no CPU fallback")
        self.net, self.optimizer, self.max_batch = net, optimizer, int(max_batch)
        group = optimizer.param_groups[0]
        cfg = azg_train_config(device=self.device.index or 0, n_blocks=len(net.res_blocks), channels=net.channels,
                               max_batch=self.max_batch, lr=float(group["lr"]), weight_decay=float(group["weight_decay"]),
                               beta1=float(group["betas"][0]), beta2=float(group["betas"][1]), eps=float(group["eps"]), clip=3.0,
                               bn_momentum=float(net.bn.momentum), bn_eps=float(net.bn.eps))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.azg_train_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.params = list(net.parameters())
        self.n = sum(p.numel() for p in self.params)
        if self.n != int(lib.azg_train_param_count(self._h)):
            raise _lib.AzgError("parameter count of the module does not match the CUDA trainer's layout")
        dev = self.device
        self.flat_params = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.offsets = []
        o = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_params[o:o + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_params[o:o + k].view_as(p)          # the module now lives in the flat vector
                self.offsets.append(o)
                o += k
        self._tracked = [m.num_batches_tracked for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
        self.step_tensor = torch.tensor(0.0)
        self.steps = 0
        self._adopt_optimizer_state()
        self._loss_parts = torch.zeros((self.max_batch, 2), dtype=torch.float32, device=dev)
        self._graphs = {}               # (batch, world) -> captured step

    def close(self):
        if getattr(self, "_h", None):
            lib.azg_train_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def memory_bytes(self) -> int:
        return int(lib.azg_train_memory_bytes(self._h))

    # ------------------------------------------------------------------ optimiser state <-> flat vectors
    def _state_is_ours(self) -> bool:
        st = self.optimizer.state
        for p, o in ((self.params[0], self.offsets[0]), (self.params[-1], self.offsets[-1])):
            s = st.get(p)
            if not s or "exp_avg" not in s or s["exp_avg"].data_ptr() != self.flat_m[o:].data_ptr():
                return False
        return True

    def _adopt_optimizer_state(self):
        """Make ``optimizer.state`` (torch.optim.Adam's own format: step, exp_avg, exp_avg_sq per parameter) views
        into the flat moment vectors, taking over whatever state is there (e.g. loaded from a checkpoint)."""
        step = 0
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                k = p.numel()
                s = self.optimizer.state[p]
                if "exp_avg" in s:
                    self.flat_m[o:o + k].copy_(s["exp_avg"].reshape(-1))
                    self.flat_v[o:o + k].copy_(s["exp_avg_sq"].reshape(-1))
                    step = max(step, int(s["step"]))
                else:
                    self.flat_m[o:o + k].zero_()
                    self.flat_v[o:o + k].zero_()
            self.steps = step
            self.step_tensor = torch.tensor(float(step))
            for p, o in zip(self.params, self.offsets):
                k = p.numel()
                s = self.optimizer.state[p]
                s["step"] = self.step_tensor                       # one shared counter (every parameter steps together)
                s["exp_avg"] = self.flat_m[o:o + k].view_as(p)
                s["exp_avg_sq"] = self.flat_v[o:o + k].view_as(p)
        self._bind()

    def _bind(self):
        net = self.net
        w = azg_net_weights()
        a = lambda t: C.c_void_p(t.data_ptr())
        for j, f in ((2, "running_mean"), (3, "running_var")):
            w.bn[j] = a(getattr(net.bn, f))
            w.policy_bn[j] = a(getattr(net.policy_bn, f))
            w.value_bn[j] = a(getattr(net.value_bn, f))
            for i, blk in enumerate(net.res_blocks):
                w.res_bn[2 * i][j] = a(getattr(blk.bn1, f))
                w.res_bn[2 * i + 1][j] = a(getattr(blk.bn2, f))
        self._stat_ptrs = tuple(b.data_ptr() for b in net.buffers())
        with torch.cuda.device(self.device):
            check(lib.azg_train_bind(self._h, ptr(self.flat_params), ptr(self.flat_grads), ptr(self.flat_m), ptr(self.flat_v),
                                     C.byref(w), int(self.steps), _stream()))
        self._param_version = tuple(p._version for p in self.params)

    def _refresh(self):
        """Before a step: notice optimiser state replaced from outside (``optimizer.load_state_dict``), buffers
        re-allocated, or parameters written through torch (``load_state_dict``, broadcasts) since the last pack."""
        if tuple(p.data_ptr() for p in self.params) != tuple(self.flat_params[o:].data_ptr() for o in self.offsets):
            raise _lib.AzgError("module parameters were re-allocated; they must stay views into the flat training vector")
        if not self._state_is_ours():
            self._adopt_optimizer_state()
        elif tuple(b.data_ptr() for b in self.net.buffers()) != self._stat_ptrs:
            self._bind()
        elif tuple(p._version for p in self.params) != self._param_version:
            with torch.cuda.device(self.device):
                check(lib.azg_train_pack(self._h, _stream()))
            self._param_version = tuple(p._version for p in self.params)

    # ------------------------------------------------------------------ the step
    def forward_backward(self, states: torch.Tensor, pis: torch.Tensor, zs: torch.Tensor) -> torch.Tensor:
        """Gradients of loss(states, pis, zs) into ``flat_grads``; returns the device tensor [2] =
        (policy_loss, value_loss) as the reference defines them (network.py:217-218)."""
        self._refresh()
        dev = self.device
        x = states.to(dev, torch.float32).contiguous()
        pi = pis.to(dev, torch.float32).contiguous()
        z = zs.to(dev, torch.float32).reshape(-1).contiguous()
        n = x.shape[0]
        if n > self.max_batch:
            raise ValueError(f"batch of {n} exceeds the trainer's max_batch {self.max_batch}")
        with torch.cuda.device(dev):
            check(lib.azg_train_forward_backward(self._h, ptr(x), ptr(pi), ptr(z), n, ptr(self._loss_parts), _stream()))
        return self._loss_parts[:n].sum(dim=0) / n

    def apply(self, world: int = 1):
        """clip_grad_norm_(3.0) + Adam on the flat vectors (``flat_grads`` holds the sum over ``world`` ranks)."""
        with torch.cuda.device(self.device):
            check(lib.azg_train_apply(self._h, int(world), _stream()))
        self.steps += 1
        self.step_tensor += 1
        torch._foreach_add_(self._tracked, 1)
        self._param_version = tuple(p._version for p in self.params)     # written by the kernel, not through torch

    # ------------------------------------------------------------------ the step as CUDA graphs
    def step_graph(self, states: torch.Tensor, pis: torch.Tensor, zs: torch.Tensor, world: int = 1, reduce_grads=None) -> torch.Tensor:
        """One full step replayed from CUDA graphs (about 150 kernel launches otherwise issued one by one from the
        host): forward + backward is one graph, clip + Adam + repack another; ``reduce_grads`` (the NCCL all-reduce
        of data-parallel training) runs between them.  Inputs are copied into static buffers; returns
        [policy_loss, value_loss] on the device."""
        self._refresh()
        n = int(states.shape[0])
        key = (n, int(world))
        g = self._graphs.get(key)
        if g is None:
            dev = self.device
            g = {"x": torch.zeros((n, 3, 15, 15), dtype=torch.float32, device=dev), "pi": torch.zeros((n, 225), dtype=torch.float32, device=dev),
                 "z": torch.zeros(n, dtype=torch.float32, device=dev), "loss": torch.zeros(2, dtype=torch.float32, device=dev)}
            g["x"][:, 2] = 1.0
            g["pi"][:] = 1.0 / 225
            saved = [t.clone() for t in (self.flat_params, self.flat_m, self.flat_v)] + [b.clone() for b in self.net.buffers()]
            steps = self.steps
            with torch.cuda.device(dev):
                # warm-up outside capture (lazy module loading, cudaFuncSetAttribute), then capture; the warm-up and the
                # capture's dry arithmetic must leave no trace: parameters, moments, statistics and step are restored
                check(lib.azg_train_forward_backward(self._h, ptr(g["x"]), ptr(g["pi"]), ptr(g["z"]), n, ptr(self._loss_parts), _stream()))
                check(lib.azg_train_apply(self._h, int(world), _stream()))
                torch.cuda.synchronize(dev)
                fb, ap = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(fb):
                    check(lib.azg_train_forward_backward(self._h, ptr(g["x"]), ptr(g["pi"]), ptr(g["z"]), n, ptr(self._loss_parts), _stream()))
                    g["loss"].copy_(self._loss_parts[:n].sum(dim=0) / n)
                with torch.cuda.graph(ap):
                    check(lib.azg_train_apply(self._h, int(world), _stream()))
            g["fb"], g["ap"] = fb, ap
            with torch.no_grad():
                for dst, src in zip([self.flat_params, self.flat_m, self.flat_v] + list(self.net.buffers()), saved):
                    dst.copy_(src)
            self.steps = steps
            self._bind()                 # restores the device step counter and repacks the restored weights
            self._graphs[key] = g
        g["x"].copy_(states.reshape(n, 3, 15, 15), non_blocking=True)
        g["pi"].copy_(pis, non_blocking=True)
        g["z"].copy_(zs.reshape(-1), non_blocking=True)
        g["fb"].replay()
        if reduce_grads is not None:
            reduce_grads(self.flat_grads)
        g["ap"].replay()
        self.steps += 1
        self.step_tensor += 1
        torch._foreach_add_(self._tracked, 1)
        self._param_version = tuple(p._version for p in self.params)
        return g["loss"].clone()

    def check(self) -> dict:
        out = (C.c_double * 3)()
        with torch.cuda.device(self.device):
            check(lib.azg_train_check(self._h, out, _stream()))
        return {"grad_norm": out[0], "clip_coef": out[1], "steps": int(out[2])}

    # ------------------------------------------------------------------ test hooks
    def gradients(self) -> dict:
        """name -> gradient tensor in the parameter's own layout (what ``p.grad`` holds after ``backward()``)."""
        out = torch.empty(self.n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.azg_train_export_grads(self._h, ptr(out), _stream()))
        names = [k for k, _ in self.net.named_parameters()]
        return {k: out[o:o + p.numel()].view_as(p) for k, p, o in zip(names, self.params, self.offsets)}

    def conv_grads(self, dz: torch.Tensor, a: torch.Tensor, layer: int):
        """The weight-gradient and input-gradient tensor-core kernels alone on given float32[n,C,15,15] tensors
        (rounded to bf16) with the current weights of trunk layer ``layer`` -> (dW [C,C,3,3], da [n,C,15,15])."""
        self._refresh()
        C_, n = self.net.channels, dz.shape[0]
        dz = dz.to(self.device, torch.float32).contiguous()
        a = a.to(self.device, torch.float32).contiguous()
        dw = torch.empty((C_, C_, 3, 3), dtype=torch.float32, device=self.device)
        da = torch.empty((n, C_, 15, 15), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.azg_train_debug_conv_grads(self._h, ptr(dz), ptr(a), n, layer, ptr(dw), ptr(da), _stream()))
        return dw, da

    def activation(self, what: int, layer: int, count: int) -> torch.Tensor:
        out = torch.empty((count, self.net.channels, 15, 15), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.azg_train_read_activation(self._h, what, layer, ptr(out), _stream()))
        return out
