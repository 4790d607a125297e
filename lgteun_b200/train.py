"""Host side of the training step (BASELINE.json configs[4]).

Reference interface mirrored (paths relative to /root/reference):
  * ``UnlgFormer.train_iter``          models/unlg_former.py:87-113  (forward in train() mode, L1 loss, backward, Adam step)
  * ``ReconstructionLoss('l1')``       models/base/losses.py:19-40   (nn.L1Loss, weight loss_cfg.rec_loss.w = 1)
  * ``Base_model.set_optim/set_sched`` models/base/base_model.py:116-144 (Adam(betas=(0.9, 0.999), lr=1.5e-3), StepLR)
  * data parallelism                   SURVEY §8e: one NCCL all-reduce of the flat gradient per step

``FlatParameters`` re-points every ``nn.Parameter`` of a ``lgteun_b200.Pansharpening`` at a slice of ONE flat fp32 CUDA
buffer laid out as the C ABI's weight table (``lgteun_weight_offset``), so the kernels read the live parameters without
copies, the gradient comes back as one flat buffer (``param.grad`` are views of it) and data-parallel training needs a
single ``all_reduce``.  ``Trainer.step`` is the fused step: train-mode forward, L1 loss, backward, all-reduce, Adam —
all hand-written kernels behind ``include/lgteun.h``; torch provides memory, streams and the NCCL call only.
There is no CPU path: everything raises without the CUDA library / a CUDA device."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import _abi
from .module import Pansharpening, UP_FACTOR


class FlatParameters:
    """One flat parameter buffer + one flat gradient buffer for a ``Pansharpening`` module on its CUDA device."""

    def __init__(self, module: Pansharpening):
        params = dict(module.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("FlatParameters needs the module on a CUDA device (call .cuda() first); there is no CPU path")
        self.device = dev
        self.module = module
        with torch.cuda.device(dev):
            self.handle = module._runtime(dev)
            self.layout = self.handle.flat_layout()
            n = self.handle.flat_numel()
            self.param = torch.zeros(n, dtype=torch.float32, device=dev)
            self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
            missing = [k for k, _, _ in self.layout if k not in params]
            if missing or len(self.layout) != len(params):
                raise RuntimeError(f"module parameters do not match the weight ABI (missing {missing[:3]}...)")
            for key, off, numel in self.layout:
                p = params[key]
                if p.numel() != numel or p.dtype != torch.float32:
                    raise RuntimeError(f"parameter {key}: expected {numel} float32 elements")
                view = self.param[off:off + numel].view(p.shape)
                view.copy_(p.data)
                p.data = view                                   # the module now lives inside the flat buffer
        self.live_keys = None

    def grad_views(self):
        """{state_dict key: view of the flat gradient} (dead priors included; their gradient is zero)."""
        params = dict(self.module.named_parameters())
        return {key: self.grad[off:off + numel].view(params[key].shape) for key, off, numel in self.layout}

    def is_live(self, key: str) -> bool:
        """Does the parameter reach the output?  Only the last prior does (models/unlg_former.py:63-67)."""
        if not key.startswith("prior_module."):
            return True
        return key.startswith(f"prior_module.{self.module.stage - 1}.")

    def attach_grads(self):
        """Expose the flat gradient as ``param.grad`` the way loss.backward() leaves it in the reference: live parameters get
        a view, parameters of the dead priors get None."""
        views = self.grad_views()
        for key, p in self.module.named_parameters():
            p.grad = views[key] if self.is_live(key) else None


def allreduce_gradients(flat_grad: torch.Tensor, group=None) -> float:
    """The one collective of data-parallel training: SUM all-reduce of the flat gradient over the ranks, in place.
    Returns the factor the optimiser applies afterwards (1 / world size): every rank's loss is a mean over ITS batch
    (nn.L1Loss, models/base/losses.py:29), so the mean over the global batch is the average of the rank gradients."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class Trainer:
    """The fused training step of ``UnlgFormer.train_iter`` for one process / one GPU; with an initialised
    ``torch.distributed`` process group (NCCL) the flat gradient is averaged over the ranks before the Adam update, i.e.
    synchronous data parallelism with per-rank batches (weak scaling)."""

    def __init__(self, module: Pansharpening, lr: float = 1.5e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 loss_weight: float = 1.0, dropout_p: float = 0.1, seed: int = 19971118, step_size: int = 0,
                 gamma: float = 0.85, process_group=None, cuda_graph: bool = True):
        # ONE flat buffer per module: the drop-in module's own train-mode forward (module._flat_parameters) must see the
        # buffer this trainer updates, otherwise a later module(ms, pan) would re-flatten into a new buffer and the
        # trainer would keep stepping the orphaned one
        self.flat = module._flat_parameters()
        self.module = module
        self.handle = self.flat.handle
        self.lr0, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.loss_weight, self.dropout_p, self.seed = float(loss_weight), float(dropout_p), int(seed)
        self.step_size, self.gamma = int(step_size), float(gamma)     # StepLR (configs/unlg_former.py:86); 0 = constant lr
        self.steps = 0
        dev = self.flat.device
        self.exp_avg = torch.zeros_like(self.flat.param)
        self.exp_avg_sq = torch.zeros_like(self.flat.param)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.group = process_group
        ddp = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if ddp else 1
        self.rank = dist.get_rank(process_group) if ddp else 0
        self._buf = {}
        # forward + loss + backward (~330 launches) replayed as ONE CUDA graph per input shape: static input buffers, the
        # dropout seed read from device memory (lgteun_train_set_seed_ptr).  The all-reduce and Adam stay outside the graph.
        self.cuda_graph = bool(cuda_graph) and os.environ.get("LGTEUN_TRAIN_GRAPH", "1") != "0"
        self._graphs = {}
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.allreduce_events = None      # set to [] to collect a CUDA-event pair around every step's all-reduce
        if self.world > 1:
            self.broadcast_state()

    def broadcast_state(self, src: int = 0):
        """Make every rank start from rank ``src``'s parameters and Adam moments (what torch DDP does at construction):
        averaged gradients applied to different weights would keep the replicas different for ever.  Call it again after
        loading a checkpoint on one rank."""
        if self.world > 1:
            for t in (self.flat.param, self.exp_avg, self.exp_avg_sq):
                dist.broadcast(t, src=src, group=self.group)
            steps = torch.tensor([self.steps], dtype=torch.int64, device=self.flat.device)
            dist.broadcast(steps, src=src, group=self.group)
            self.steps = int(steps.item())
            self.module._invalidate_runtime()

    def _check_alias(self):
        """The module's parameters must still live inside ``flat.param`` (a .to() / .cuda() / load of new Parameter objects
        moves them out): re-flattening silently would orphan the optimiser state, so raise instead."""
        base = self.flat.param.data_ptr()
        params = dict(self.module.named_parameters())
        for key, off, _ in self.flat.layout:
            if params[key].data_ptr() != base + 4 * off:
                raise RuntimeError(f"parameter {key} no longer aliases the trainer's flat buffer (the module was moved or its "
                                   "parameters were replaced after Trainer was created); build a new Trainer")
        if self.module._flat is not self.flat:
            raise RuntimeError("the module was re-flattened after this Trainer was created; build a new Trainer")

    def dropout_seed(self) -> int:
        """Per-step, per-rank seed of the counter-based dropout masks: ranks must not drop the same positions."""
        return (self.seed + self.steps + 0x9E3779B1 * self.rank) & 0x7FFFFFFFFFFFFFFF

    def lr(self) -> float:
        return self.lr0 * (self.gamma ** (self.steps // self.step_size)) if self.step_size > 0 else self.lr0

    def _buffers(self, shape):
        b = self._buf.get(shape)
        if b is None:
            dev = self.flat.device
            b = (torch.empty(shape, dtype=torch.float32, device=dev), torch.empty(shape, dtype=torch.float32, device=dev))
            self._buf = {shape: b}
        return b

    def forward_backward(self, ms: torch.Tensor, pan: torch.Tensor, gt: torch.Tensor):
        """train-mode forward, weighted L1 loss and backward; fills ``flat.grad`` (local gradient) and returns (out, loss)."""
        n, b, h, w = ms.shape
        if tuple(pan.shape) != (n, 1, UP_FACTOR * h, UP_FACTOR * w) or tuple(gt.shape) != (n, b, UP_FACTOR * h, UP_FACTOR * w):
            raise ValueError(f"shape mismatch: ms {tuple(ms.shape)}, pan {tuple(pan.shape)}, gt {tuple(gt.shape)}")
        for t in (ms, pan, gt):
            if t.device != self.flat.device or t.dtype != torch.float32:
                raise RuntimeError("Trainer needs float32 tensors on the module's CUDA device")
        ms, pan, gt = ms.contiguous(), pan.contiguous(), gt.contiguous()
        self._check_alias()
        if self.cuda_graph and not getattr(self.handle, "ext_masks", False) and not torch.cuda.is_current_stream_capturing():
            return self._graphed_forward_backward(ms, pan, gt)
        with torch.cuda.device(self.flat.device):
            stream = torch.cuda.current_stream().cuda_stream
            out, dout = self._buffers(tuple(gt.shape))
            self.handle.train_forward(self.flat.param.data_ptr(), ms.data_ptr(), pan.data_ptr(), out.data_ptr(), n, h, w,
                                      self.dropout_p, self.dropout_seed(), stream)
            self.handle.l1_loss(out.data_ptr(), gt.data_ptr(), out.numel(), self.loss_weight, self.loss.data_ptr(),
                                dout.data_ptr(), stream)
            self.handle.train_backward(dout.data_ptr(), self.flat.grad.data_ptr(), stream)
        return out, self.loss

    def _graphed_forward_backward(self, ms, pan, gt):
        """The same three native calls, captured once per shape and replayed.  The first call of a shape runs eagerly (it sizes
        the tape), the second captures; inputs are copied into the graph's static buffers, the seed into device memory."""
        key = tuple(ms.shape)
        dev = self.flat.device
        with torch.cuda.device(dev):
            out, dout = self._buffers(tuple(gt.shape))
            entry = self._graphs.get(key)
            if entry is None:                               # eager warm-up: allocations, function attributes, the tape
                self._graphs = {key: "warm"}                # (one shape at a time, like _buffers)
                stream = torch.cuda.current_stream().cuda_stream
                n, b, h, w = ms.shape
                self.handle.train_forward(self.flat.param.data_ptr(), ms.data_ptr(), pan.data_ptr(), out.data_ptr(), n, h, w,
                                          self.dropout_p, self.dropout_seed(), stream)
                self.handle.l1_loss(out.data_ptr(), gt.data_ptr(), out.numel(), self.loss_weight, self.loss.data_ptr(),
                                    dout.data_ptr(), stream)
                self.handle.train_backward(dout.data_ptr(), self.flat.grad.data_ptr(), stream)
                return out, self.loss
            if entry == "warm":
                n, b, h, w = ms.shape
                sms, span, sgt = torch.empty_like(ms), torch.empty_like(pan), torch.empty_like(gt)
                graph = torch.cuda.CUDAGraph()
                self.handle.train_set_seed_ptr(self._seed_dev.data_ptr())
                try:
                    torch.cuda.synchronize(dev)
                    # thread_local: the NCCL watchdog thread of a data-parallel job keeps polling events during the capture
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        stream = torch.cuda.current_stream().cuda_stream
                        self.handle.train_forward(self.flat.param.data_ptr(), sms.data_ptr(), span.data_ptr(), out.data_ptr(),
                                                  n, h, w, self.dropout_p, 0, stream)
                        self.handle.l1_loss(out.data_ptr(), sgt.data_ptr(), out.numel(), self.loss_weight,
                                            self.loss.data_ptr(), dout.data_ptr(), stream)
                        self.handle.train_backward(dout.data_ptr(), self.flat.grad.data_ptr(), stream)
                finally:
                    self.handle.train_set_seed_ptr(0)
                entry = self._graphs[key] = (graph, sms, span, sgt, self.flat.param.data_ptr(), self.flat.grad.data_ptr())
            graph, sms, span, sgt, p_ptr, g_ptr = entry
            if p_ptr != self.flat.param.data_ptr() or g_ptr != self.flat.grad.data_ptr():
                raise RuntimeError("the flat parameter / gradient buffers moved after the step was captured; build a new Trainer")
            sms.copy_(ms)
            span.copy_(pan)
            sgt.copy_(gt)
            self._seed_dev.fill_(self.dropout_seed())
            graph.replay()
        return out, self.loss

    def step(self, ms: torch.Tensor, pan: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        """One train_iter: returns the (local) loss as a 1-element CUDA tensor (no host sync)."""
        _, loss = self.forward_backward(ms, pan, gt)
        if self.allreduce_events is not None:                # measurement aid (bench.py): device time of the collective
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self.flat.device))
        gscale = allreduce_gradients(self.flat.grad, self.group)
        if self.allreduce_events is not None:
            e1.record(torch.cuda.current_stream(self.flat.device))
            self.allreduce_events.append((e0, e1))
        lr = self.lr()
        self.steps += 1
        with torch.cuda.device(self.flat.device):
            self.handle.adam_step(self.flat.param.data_ptr(), self.flat.grad.data_ptr(), self.exp_avg.data_ptr(),
                                  self.exp_avg_sq.data_ptr(), self.flat.param.numel(), lr, self.betas[0], self.betas[1], self.eps,
                                  self.steps, gscale, torch.cuda.current_stream().cuda_stream)
        self.module._invalidate_runtime()      # the eval-mode handle keeps a packed snapshot of the weights: refresh on next use
        return loss


class _TrainForward(torch.autograd.Function):
    """``Pansharpening.forward`` under autograd (the reference's ``train_iter`` calls ``G(lr, pan)`` and ``loss.backward()``,
    models/unlg_former.py:94,108-110): forward = lgteun_train_forward, backward = lgteun_train_backward; the parameters are
    passed through ``apply`` only so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, flat, ms, pan, dropout_p, seed, *params):
        n, b, h, w = ms.shape
        out = torch.empty((n, b, UP_FACTOR * h, UP_FACTOR * w), dtype=torch.float32, device=ms.device)
        with torch.cuda.device(ms.device):
            flat.handle.train_forward(flat.param.data_ptr(), ms.data_ptr(), pan.data_ptr(), out.data_ptr(), n, h, w,
                                      dropout_p, seed, torch.cuda.current_stream().cuda_stream)
        ctx.flat = flat
        ctx.generation = flat.handle.train_generation()   # a handle keeps ONE tape: backward() names the forward it belongs to
        ctx.keep = (ms, pan)                    # the tape refers to the caller's input buffers
        return out

    @staticmethod
    def backward(ctx, dout):
        flat = ctx.flat
        dout = dout.contiguous()
        with torch.cuda.device(dout.device):
            # raises (LGTEUN_ESTATE) if a later train-mode forward overwrote this node's tape: forward A, forward B,
            # backward A must not combine A's dout with B's activations
            flat.handle.train_backward(dout.data_ptr(), flat.grad.data_ptr(), torch.cuda.current_stream().cuda_stream,
                                       generation=ctx.generation)
        g = flat.grad.clone()                   # param.grad must not alias the buffer the next backward overwrites
        shapes = {k: p.shape for k, p in flat.module.named_parameters()}
        grads = tuple(g[off:off + numel].view(shapes[key]) if flat.is_live(key) else None for key, off, numel in flat.layout)
        return (None, None, None, None, None) + grads


def autograd_forward(module: Pansharpening, ms: torch.Tensor, pan: torch.Tensor, dropout_p: float, seed: int) -> torch.Tensor:
    flat = module._flat_parameters()
    params = dict(module.named_parameters())
    ordered = [params[key] for key, _, _ in flat.layout]
    return _TrainForward.apply(flat, ms.contiguous(), pan.contiguous(), float(dropout_p), int(seed), *ordered)
