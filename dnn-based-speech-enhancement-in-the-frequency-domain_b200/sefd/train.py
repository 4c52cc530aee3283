"""Autograd-free train step over the C ABI: forward (+loss sums in the epilogue) -> loss -> backward ->
[one NCCL all-reduce of the flat gradient buffer] -> fused Adam.  Equivalent to the body of the reference's
trainer.model_train loop (trainer.py:27-37) with torch.optim.Adam(lr) (train_interface.py:59).

Data-parallel semantics: utterances are sharded by batch, weights replicated, BatchNorm uses
per-rank statistics (standard DDP), gradients are summed over ranks and scaled by 1/world inside Adam.
"""
import os

import torch

from . import _lib
from . import dccrn as _dccrn
from . import dist as _dist
from . import ops as _ops
from .ops import LOSSES, ptr, stream


class TrainStep:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, loss="SI-SNR", process_group=None, perceptual=False,
                 graph=False):
        """perceptual = 'PMSQE': the step of trainer.model_perceptual_train (trainer.py:44-70) with r1 = r2 = 1,
        loss = (main + PMSQE(out_wav, clean)) / 2 (BASELINE configs[3]).
        graph = True: step() replays ONE CUDA graph per (input buffers, shape): the ~200 kernel launches of a step (forward,
        loss, backward with its side-stream folds, all-reduce, Adam) become one host call.  The first call with a new pair of
        input buffers runs eagerly (it initialises plans / workspaces), the second captures, later ones replay; the Adam step
        count lives on the device.  Inputs must then be written INTO the same device buffers every step (e.g. WaveFeeder slots)."""
        self.graph = bool(graph)
        self._graphs, self._seen = {}, set()
        if perceptual not in (False, None, "PMSQE"):
            raise NotImplementedError(f"TrainStep: perceptual={perceptual!r} (built here: 'PMSQE'; LMS runs through the "
                                      "autograd drop-in, models.DCCRN.loss(..., perceptual=True))")
        self.perceptual = perceptual or False
        self.model = model
        self.engine = model._get_engine()
        self.engine.sync()
        self.lr, self.betas, self.eps = lr, betas, eps
        self.kind = LOSSES[loss]
        self.exp_avg = torch.zeros_like(self.engine.flat)
        self.exp_avg_sq = torch.zeros_like(self.engine.flat)
        self.steps = 0
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=self.engine.flat.device)
        self._bc_dev = torch.zeros(2, device=self.engine.flat.device)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self._bufs = {}
        _dist.broadcast_from_rank0_(self.engine, self.pg)                 # every replica starts from rank 0's weights
        # data-parallel overlap: the decoder / LSTM / projection gradients (the tail of the flat buffer, 76 %
        # of it) are final before the encoder backward starts; their all-reduce runs on a side stream beside it, the encoder
        # slice is reduced when the backward ends.  SEFD_DP_OVERLAP=0 falls back to one all-reduce after the backward.
        self.overlap = (self.world > 1 and self.engine.family == "dccrn" and os.environ.get("SEFD_DP_OVERLAP", "1") != "0")
        self._tail_pending = False
        if self.overlap:
            self._comm_stream = torch.cuda.Stream()
            self._tail_event = torch.cuda.Event()
            self._tail_event.record()                                      # materialises the cudaEvent_t handle
            self._split = int(_lib.load().sefd_dccrn_grad_split(self.engine._layout.handle))

    def _scratch(self, B, L, dev):
        key = (B, L)
        if key not in self._bufs:
            self._bufs[key] = dict(
                wav=torch.empty(B, L, device=dev), dwav=torch.empty(B, L, device=dev),
                loss=torch.empty(1, device=dev), coef=torch.empty(2 * B, device=dev))
        return self._bufs[key]

    def forward_backward(self, noisy, clean, reduce_tail=False):
        """Fills engine.flat_grad with this rank's gradient; returns the loss tensor (1 element, device).  reduce_tail (used by
        step()): with the data-parallel overlap on, the tail slice of the gradient is already being all-reduced on return."""
        lib = _lib.load()
        eng = self.engine
        eng.sync()
        if noisy.dim() != 2 or noisy.shape != clean.shape:
            raise ValueError(f"TrainStep: expected noisy, clean of one shape [B, L], got {tuple(noisy.shape)} / {tuple(clean.shape)}")
        _ops._req(noisy, clean)       # CUDA, float32, contiguous: the kernels index rows with stride L (a strided view of a
        B, L = noisy.shape            # [B, 2, L] batch would silently pair the wrong rows)
        plan = eng.plan(B, L)
        ws = plan.workspace(noisy.device)
        s = self._scratch(B, L, noisy.device)
        st = stream()
        plan.generation += 1
        fwd, bwd = (lib.sefd_crn_forward, lib.sefd_crn_backward) if eng.family == "crn" else \
                   (lib.sefd_dccrn_forward, lib.sefd_dccrn_backward)
        _lib.check(fwd(plan.handle, ptr(eng.flat), ptr(eng.flat_buf), ptr(noisy), ptr(clean), 1,
                       None, None, ptr(s["wav"]), ptr(ws), plan.ws_bytes, st), eng.family + "_forward")
        _dccrn.bump_batches_tracked(eng.module)       # train-mode forward: BatchNorm2d.num_batches_tracked += 1
        _lib.check(lib.sefd_dccrn_loss(plan.handle, ptr(s["wav"]), ptr(clean), self.kind, 1, ptr(s["loss"]),
                                       ptr(s["coef"]), ptr(ws), st), "dccrn_loss")
        if self.perceptual == "PMSQE":
            if "pws" not in s:
                nbytes = lib.sefd_pmsqe_workspace_bytes(B, L)
                if nbytes == 0:
                    raise ValueError(f"PMSQE: waveforms must be 1..4 whole seconds at 16 kHz, got {L} samples")
                s.update(pws=torch.empty(nbytes, device=noisy.device, dtype=torch.uint8), ploss=torch.empty(1, device=noisy.device),
                         dpw=torch.empty(B, L, device=noisy.device), half=torch.full((1,), 0.5, device=noisy.device),
                         tables=_ops._pmsqe_tables(noisy.device))
            nb = s["pws"].numel()
            _lib.check(lib.sefd_pmsqe_forward(ptr(s["wav"]), ptr(clean), B, L, ptr(s["tables"]), ptr(s["pws"]), nb,
                                              ptr(s["ploss"]), st), "pmsqe_forward")
            _lib.check(lib.sefd_loss_backward(ptr(s["wav"]), ptr(clean), ptr(s["coef"]), ptr(s["half"]), ptr(s["dwav"]), B, L,
                                              st), "loss_backward")
            _lib.check(lib.sefd_pmsqe_backward(ptr(s["half"]), B, L, ptr(s["tables"]), ptr(s["pws"]), nb, ptr(s["dpw"]), st),
                       "pmsqe_backward")
            _lib.check(lib.sefd_axpby(ptr(s["dwav"]), ptr(s["dpw"]), 1.0, 1.0, B * L, st), "axpby")
            _lib.check(lib.sefd_axpby(ptr(s["loss"]), ptr(s["ploss"]), 0.5, 0.5, 1, st), "axpby")
        else:
            _lib.check(lib.sefd_loss_backward(ptr(s["wav"]), ptr(clean), ptr(s["coef"]), None, ptr(s["dwav"]), B, L, st),
                       "loss_backward")
        if self.overlap and reduce_tail:
            _lib.check(lib.sefd_dccrn_backward_overlap(plan.handle, ptr(eng.flat), ptr(s["dwav"]), ptr(eng.flat_grad), ptr(ws),
                                                       plan.ws_bytes, st, self._tail_event.cuda_event), "dccrn_backward_overlap")
            self._comm_stream.wait_event(self._tail_event)
            with torch.cuda.stream(self._comm_stream):
                _dist.allreduce_sum_(eng.flat_grad[self._split:], self.pg)
            self._tail_pending = True
        else:
            _lib.check(bwd(plan.handle, ptr(eng.flat), ptr(s["dwav"]), ptr(eng.flat_grad), ptr(ws),
                           plan.ws_bytes, st), eng.family + "_backward")
        return s["loss"]

    def _step_eager(self, noisy, clean):
        loss = self.forward_backward(noisy, clean, reduce_tail=True)
        eng = self.engine
        if self._tail_pending:                                             # head slice now, tail slice already in flight
            gscale = _dist.allreduce_sum_(eng.flat_grad[:self._split], self.pg)
            torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._tail_pending = False
        else:
            gscale = _dist.allreduce_sum_(eng.flat_grad, self.pg)          # single flat 14.7 MB buffer
        # the step count lives on the device (same arithmetic as the host-side sefd_adam_step: corrections in double)
        _lib.check(_lib.load().sefd_adam_step_dev(ptr(eng.flat), ptr(eng.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                                  eng.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                                  ptr(self._step_dev), ptr(self._bc_dev), gscale, stream()), "adam_step_dev")
        return loss

    def step(self, noisy, clean):
        self.steps += 1
        if not self.graph:
            return self._step_eager(noisy, clean)
        key = (noisy.data_ptr(), clean.data_ptr(), tuple(noisy.shape))
        g = self._graphs.get(key)
        if g is None:
            if key not in self._seen:                  # first sight of these buffers: eager (initialises plans, workspaces, streams)
                self._seen.add(key)
                return self._step_eager(noisy, clean)
            self.engine.sync()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):                  # records the step (nothing executes yet)
                self._graph_loss = self._step_eager(noisy, clean)
            self._graphs[key] = g
            self.launches_per_graph = None
        g.replay()
        return self._graph_loss


class FlatAdam:
    """Optimizer-shaped wrapper (zero_grad / step) that can stand where train_interface.py:59 builds
    torch.optim.Adam(model.parameters(), lr): after loss.backward() the gradients already sit in the model's flat
    gradient buffer, so step() is [all-reduce over ranks] + one fused Adam kernel.  One backward per step, like
    the reference loop (trainer.py:35-37)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None):
        self.model, self.lr, self.betas, self.eps, self.pg = model, lr, betas, eps, process_group
        self.engine = model._get_engine()
        self.engine.sync()
        self.exp_avg = torch.zeros_like(self.engine.flat)
        self.exp_avg_sq = torch.zeros_like(self.engine.flat)
        self.steps = 0
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        _dist.broadcast_from_rank0_(self.engine, process_group)            # like DDP: replicas start from rank 0's weights
        self.engine.backwards_since_step = 0

    def zero_grad(self, set_to_none=True):
        if not set_to_none:
            raise RuntimeError("FlatAdam reads the model's flat gradient buffer, which holds the LAST backward only: "
                               "zero_grad(set_to_none=False) / gradient accumulation is not supported")
        for p in self.model.parameters():
            p.grad = None

    def step(self):
        eng = self.engine
        if eng.backwards_since_step != 1:
            raise RuntimeError(f"FlatAdam.step(): {eng.backwards_since_step} backward passes since the last step; the flat "
                               "gradient buffer holds exactly one (no accumulation, trainer.py:35-37 runs one per step)")
        eng.backwards_since_step = 0
        if eng.flat.data_ptr() != self.exp_avg.data_ptr() and eng.flat.numel() != self.exp_avg.numel():
            raise RuntimeError("FlatAdam: the model's parameter layout changed")
        gscale = _dist.allreduce_sum_(eng.flat_grad, self.pg)
        self.steps += 1
        _lib.check(_lib.load().sefd_adam_step(ptr(eng.flat), ptr(eng.flat_grad), ptr(self.exp_avg),
                                              ptr(self.exp_avg_sq), eng.flat.numel(), self.lr, self.betas[0],
                                              self.betas[1], self.eps, self.steps, gscale, stream()),
                   "adam_step")


class FsnTrainStep:
    """Autograd-free FullSubNet train step over the C ABI = the body of trainer.fullsubnet_train (trainer.py:97-112) with
    cfg.loss = 'MSE' and torch.optim.Adam(lr): feature / target kernel (tools.stft x2 + mag_phase +
    build_complex_ideal_ratio_mask fused) -> FullSubNet forward (train mode: Philox inter-layer dropout, p = 0.8 like
    tools_for_model.py:746) -> MSE(cIRM, cRM) -> backward -> [all-reduce] -> fused Adam."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None, dropout=None, seed=0):
        self.model = model
        self.engine = model._get_engine()
        self.engine.sync()
        self.lr, self.betas, self.eps = lr, betas, eps
        self.dropout = model.dropout if dropout is None else float(dropout)
        self.seed = int(seed)
        self.exp_avg = torch.zeros_like(self.engine.flat)
        self.exp_avg_sq = torch.zeros_like(self.engine.flat)
        self.steps = 0
        self.pg = process_group
        self._bufs = {}
        _dist.broadcast_from_rank0_(self.engine, self.pg)

    def _scratch(self, B, L, dev):
        key = (B, L)
        if key not in self._bufs:
            Tf = _lib.load().sefd_fsn_frames(L)
            if Tf <= 0:
                raise ValueError(f"FsnTrainStep: waveforms of {L} samples are too short for the centred STFT")
            self._bufs[key] = dict(
                Tf=Tf, mag=torch.empty(B, 257, Tf, device=dev), cirm=torch.empty(B, 257, Tf, 2, device=dev),
                crm=torch.empty(B, 257, Tf, 2, device=dev), dcrm=torch.empty(B, 257, Tf, 2, device=dev),
                loss=torch.empty(1, device=dev), coef=torch.empty(2 * B, device=dev),
                red=torch.empty(8 * B, device=dev, dtype=torch.float64))
        return self._bufs[key]

    def forward_backward(self, noisy, clean):
        lib = _lib.load()
        eng = self.engine
        eng.sync()
        if noisy.dim() != 2 or noisy.shape != clean.shape:
            raise ValueError(f"FsnTrainStep: expected noisy, clean of one shape [B, L], got {tuple(noisy.shape)} / {tuple(clean.shape)}")
        _ops._req(noisy, clean)
        B, L = noisy.shape
        s = self._scratch(B, L, noisy.device)
        plan = eng.plan(B, s["Tf"])
        ws = plan.workspace(noisy.device)
        st = stream()
        n = 257 * s["Tf"] * 2
        plan.generation += 1
        _lib.check(lib.sefd_fsn_features(ptr(noisy), ptr(clean), B, L, ptr(s["mag"]), ptr(s["cirm"]), st), "fsn_features")
        _lib.check(lib.sefd_fsn_forward(plan.handle, ptr(eng.flat), ptr(s["mag"]), 1, self.dropout, None, None,
                                        self.seed + self.steps, ptr(s["crm"]), ptr(ws), plan.ws_bytes, st), "fsn_forward")
        _lib.check(lib.sefd_loss_forward(ptr(s["crm"]), ptr(s["cirm"]), B, n, LOSSES["MSE"], ptr(s["red"]), ptr(s["loss"]),
                                         ptr(s["coef"]), st), "loss_forward")
        _lib.check(lib.sefd_loss_backward(ptr(s["crm"]), ptr(s["cirm"]), ptr(s["coef"]), None, ptr(s["dcrm"]), B, n, st),
                   "loss_backward")
        _lib.check(lib.sefd_fsn_backward(plan.handle, ptr(eng.flat), ptr(s["dcrm"]), ptr(eng.flat_grad), ptr(ws),
                                         plan.ws_bytes, st), "fsn_backward")
        return s["loss"]

    def step(self, noisy, clean):
        loss = self.forward_backward(noisy, clean)
        eng = self.engine
        gscale = _dist.allreduce_sum_(eng.flat_grad, self.pg)
        self.steps += 1
        _lib.check(_lib.load().sefd_adam_step(ptr(eng.flat), ptr(eng.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                              eng.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.steps,
                                              gscale, stream()), "adam_step")
        return loss
