"""Host side of the DCCRN path: parameter containers with the reference's state_dict keys, the flat
parameter / gradient buffers the C ABI works on, the plan + workspace cache and the autograd bridge.

Reference being mirrored: models.py:15-323 (class DCCRN), tools_for_model.py:36-112,141-338.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .ops import LOSSES, MODES, PLAN_CBN, PLAN_NO_SKIP, PLAN_REAL_LSTM, ptr, stream

KERNEL_NUM = [32, 64, 128, 256, 256, 256]


# --------------------------------------------------------------------------------------------------
# parameter containers (no compute here: they only own tensors under the reference's names and
# reproduce the reference's initialisation, including how it consumes the torch RNG stream)
# --------------------------------------------------------------------------------------------------
class ConvParams(nn.Module):
    """weight/bias of one nn.Conv2d / nn.ConvTranspose2d of the reference (tools_for_model.py:233-241, 299-311):
    torch's default init is drawn first (kaiming-uniform weight, uniform bias), then overwritten by
    N(0, 0.05) weights and zero bias - the same order of RNG draws as the reference."""

    def __init__(self, shape, n_bias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape))
        self.bias = nn.Parameter(torch.empty(n_bias))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(shape[1] * shape[2] * shape[3])   # torch's fan_in: size(1) * receptive field
        nn.init.uniform_(self.bias, -bound, bound)


class ComplexConvParams(nn.Module):
    """real_conv / imag_conv pair of ComplexConv2d (transposed=False) or ComplexConvTranspose2d."""

    def __init__(self, cin, cout, transposed):
        super().__init__()
        ci, co = cin // 2, cout // 2
        shape = (ci, co, 5, 2) if transposed else (co, ci, 5, 2)
        self.real_conv = ConvParams(shape, co)
        self.imag_conv = ConvParams(shape, co)
        nn.init.normal_(self.real_conv.weight.data, std=0.05)
        nn.init.normal_(self.imag_conv.weight.data, std=0.05)
        nn.init.constant_(self.real_conv.bias, 0.0)
        nn.init.constant_(self.imag_conv.bias, 0.0)


class BatchNormParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class ComplexBatchNormParams(BatchNormParams):
    """ComplexBatchNorm(c) (tools_for_model.py:430-491): c // 2 complex features; reset_parameters draws Wri ~ U(-0.9, 0.9) from
    the global RNG at construction (same stream position as the reference).  A BatchNormParams subclass only so that the
    bookkeeping that looks for normalisation layers (num_batches_tracked) finds it; it owns none of BatchNorm2d's tensors."""

    def __init__(self, c):
        nn.Module.__init__(self)
        h = c // 2
        self.Wrr = nn.Parameter(torch.ones(h))
        self.Wri = nn.Parameter(torch.empty(h).uniform_(-.9, +.9))
        self.Wii = nn.Parameter(torch.ones(h))
        self.Br = nn.Parameter(torch.zeros(h))
        self.Bi = nn.Parameter(torch.zeros(h))
        self.register_buffer("RMr", torch.zeros(h))
        self.register_buffer("RMi", torch.zeros(h))
        self.register_buffer("RVrr", torch.ones(h))
        self.register_buffer("RVri", torch.zeros(h))
        self.register_buffer("RVii", torch.ones(h))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class PReLUParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.full((1,), 0.25))


class LSTMParams(nn.Module):
    """Parameters of a single-layer nn.LSTM, initialised U(-1/sqrt(H), 1/sqrt(H)) in nn.LSTM's order."""

    def __init__(self, input_size, hidden):
        super().__init__()
        k = 1.0 / math.sqrt(hidden)
        self.weight_ih_l0 = nn.Parameter(torch.empty(4 * hidden, input_size))
        self.weight_hh_l0 = nn.Parameter(torch.empty(4 * hidden, hidden))
        self.bias_ih_l0 = nn.Parameter(torch.empty(4 * hidden))
        self.bias_hh_l0 = nn.Parameter(torch.empty(4 * hidden))
        for p in (self.weight_ih_l0, self.weight_hh_l0, self.bias_ih_l0, self.bias_hh_l0):
            nn.init.uniform_(p, -k, k)


class LinearParams(nn.Module):
    def __init__(self, fin, fout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin))
        self.bias = nn.Parameter(torch.empty(fout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(fin)
        nn.init.uniform_(self.bias, -bound, bound)


class ComplexLSTMParams(nn.Module):
    """NavieComplexLSTM parameters (tools_for_model.py:141-160)."""

    def __init__(self, input_size, hidden_size, projection_dim=None):
        super().__init__()
        self.real_lstm = LSTMParams(input_size // 2, hidden_size // 2)
        self.imag_lstm = LSTMParams(input_size // 2, hidden_size // 2)
        if projection_dim is not None:
            self.r_trans = LinearParams(hidden_size // 2, projection_dim // 2)
            self.i_trans = LinearParams(hidden_size // 2, projection_dim // 2)


class RealConvParams(nn.Module):
    """RealConv2d / RealConvTranspose2d (tools_for_model.py:341-425): one nn.Conv2d / nn.ConvTranspose2d named `conv`,
    default torch init drawn first, then N(0, 0.05) weights and zero bias (same RNG consumption as the reference)."""

    def __init__(self, cin, cout, transposed):
        super().__init__()
        shape = (cin, cout, 5, 2) if transposed else (cout, cin, 5, 2)
        self.conv = ConvParams(shape, cout)
        nn.init.normal_(self.conv.weight.data, std=0.05)
        nn.init.constant_(self.conv.bias, 0.0)


class STFTBuffers(nn.Module):
    """Buffers the reference keeps in ConvSTFT / ConviSTFT (tools_for_model.py:16-33,46,81,88-89).  The CUDA
    kernels do not read them (they use the FFT closed form); they exist so checkpoints round-trip."""

    def __init__(self, win_len, fft_len, inverse):
        super().__init__()
        n = np.arange(win_len, dtype=np.float64)
        w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_len)
        k = np.arange(fft_len // 2 + 1, dtype=np.float64)[:, None]
        ang = 2.0 * np.pi * k * n[None, :] / fft_len
        kern = np.concatenate([np.cos(ang), -np.sin(ang)], 0)
        if inverse:
            kern = np.linalg.pinv(kern).T
        self.register_buffer("weight", torch.from_numpy((kern * w).astype(np.float32))[:, None, :])
        if inverse:
            self.register_buffer("window", torch.from_numpy(w.astype(np.float32))[None, :, None])
            self.register_buffer("enframe", torch.eye(win_len)[:, None, :])


# --------------------------------------------------------------------------------------------------
# plan + workspace cache
# --------------------------------------------------------------------------------------------------
class Plan:
    def __init__(self, B, L, mode, family="dccrn", skip=True, real_lstm=False, cbn=False):
        lib = _lib.load()
        self.family = family
        self.skip = skip
        self.real_lstm = real_lstm
        self.cbn = cbn
        if family == "crn":
            self.handle = lib.sefd_crn_plan_create(B, L)
        elif family == "fsn":                      # L = number of STFT frames of noisy_mag
            self.handle = lib.sefd_fsn_plan_create(B, L)
        else:
            self.handle = lib.sefd_dccrn_plan_create_ex(B, L, MODES[mode], (0 if skip else PLAN_NO_SKIP) | (PLAN_REAL_LSTM if real_lstm else 0) |
                                                        (PLAN_CBN if cbn else 0))
        if not self.handle:
            raise RuntimeError("sefd plan: " + lib.sefd_last_error().decode())
        self.B, self.L, self.T, self.mode = B, L, (L + 2 if family == "fsn" else L // 100 + 3), mode
        self.ws_bytes = lib.sefd_dccrn_workspace_bytes(self.handle)
        self.n_param = lib.sefd_dccrn_param_floats(self.handle)
        self.n_buf = lib.sefd_dccrn_buffer_floats(self.handle)
        self.params = self._entries(0, lib.sefd_dccrn_num_params(self.handle))
        self.buffers = self._entries(1, lib.sefd_dccrn_num_buffers(self.handle))
        self.ws = None
        self.generation = 0

    def _entries(self, kind, n):
        lib = _lib.load()
        out = []
        name = C.create_string_buffer(128)
        off, numel, ndim = C.c_longlong(), C.c_longlong(), C.c_int()
        shape = (C.c_longlong * 4)()
        for i in range(n):
            _lib.check(lib.sefd_dccrn_entry_info(self.handle, kind, i, name, 128, C.byref(off), C.byref(numel),
                                                 C.byref(ndim), shape), "entry_info")
            out.append((name.value.decode(), off.value, numel.value, tuple(shape[: ndim.value])))
        return out

    def workspace(self, device):
        if self.ws is None or self.ws.device != device:
            self.ws = torch.empty(self.ws_bytes, device=device, dtype=torch.uint8)
        return self.ws

    def tensor(self, name):
        """View of a named intermediate inside the workspace (tests / debugging)."""
        lib = _lib.load()
        off, ndim = C.c_longlong(), C.c_int()
        shape = (C.c_longlong * 4)()
        _lib.check(lib.sefd_dccrn_tensor_info(self.handle, name.encode(), C.byref(off), C.byref(ndim), shape),
                   "tensor_info")
        shp = tuple(shape[: ndim.value])
        n = int(np.prod(shp))
        return self.ws.view(torch.float32)[off.value: off.value + n].view(*shp)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().sefd_dccrn_plan_destroy(self.handle)
        except Exception:
            pass


class _Forward(torch.autograd.Function):
    """Bridges torch autograd to sefd_dccrn_forward / sefd_dccrn_backward."""

    @staticmethod
    def forward(ctx, engine, noisy, target, train, *params):
        plan = engine.plan(noisy.shape[0], noisy.shape[1])
        dev = noisy.device
        ws = plan.workspace(dev)
        B, L, T = plan.B, plan.L, plan.T
        out_real = torch.empty(B, 257, T, device=dev)
        out_imag = torch.empty(B, 257, T, device=dev)
        out_wav = torch.empty(B, L, device=dev)
        plan.generation += 1
        _lib.check(_lib.load().sefd_dccrn_forward(
            plan.handle, ptr(engine.flat), ptr(engine.flat_buf), ptr(noisy), ptr(target), int(train),
            ptr(out_real), ptr(out_imag), ptr(out_wav), ptr(ws), plan.ws_bytes, stream()), "dccrn_forward")
        ctx.engine, ctx.plan, ctx.generation = engine, plan, plan.generation
        ctx.set_materialize_grads(False)       # unused outputs arrive as None, not as zero tensors
        return out_real, out_imag, out_wav

    @staticmethod
    def backward(ctx, g_real, g_imag, g_wav):
        engine, plan = ctx.engine, ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("sefd: the activation workspace of this forward was overwritten by a later forward "
                               "of the same batch shape; call backward() before the next forward")
        lib = _lib.load()
        g_wav = None if g_wav is None else g_wav.contiguous().float()
        if g_real is None and g_imag is None:
            if g_wav is None:
                raise RuntimeError("sefd: backward() reached the DCCRN forward without any gradient")
            _lib.check(lib.sefd_dccrn_backward(plan.handle, ptr(engine.flat), ptr(g_wav), ptr(engine.flat_grad),
                                               ptr(plan.ws), plan.ws_bytes, stream()), "dccrn_backward")
        else:
            # a loss on the masked spectrum (perceptual LMS branch, models.py:305-312) sends gradients to out_real / out_imag
            ref = g_real if g_real is not None else g_imag
            g_real = (torch.zeros_like(ref) if g_real is None else g_real).contiguous().float()
            g_imag = (torch.zeros_like(ref) if g_imag is None else g_imag).contiguous().float()
            _lib.check(lib.sefd_dccrn_backward_spec(plan.handle, ptr(engine.flat), ptr(g_wav), ptr(g_real), ptr(g_imag),
                                                    ptr(engine.flat_grad), ptr(plan.ws), plan.ws_bytes, stream()),
                       "dccrn_backward_spec")
        engine.backwards_since_step += 1
        flat = engine.flat_grad.clone()     # p.grad must never alias the buffer the next backward overwrites (AccumulateGrad
        grads = tuple(flat[o: o + n].view(shape) for (_, o, n, shape) in plan.params)   # may keep the tensor it is handed)
        return (None, None, None, None) + grads


class _ForwardCRN(torch.autograd.Function):
    """Bridges torch autograd to sefd_crn_forward / sefd_crn_backward (CRN.forward, models.py:460-532)."""

    @staticmethod
    def forward(ctx, engine, noisy, target, train, *params):
        plan = engine.plan(noisy.shape[0], noisy.shape[1])
        dev = noisy.device
        ws = plan.workspace(dev)
        B, L, T = plan.B, plan.L, plan.T
        est_mags = torch.empty(B, 257, T, device=dev)
        target_mags = torch.empty(B, 257, T, device=dev) if target is not None else None
        out_wav = torch.empty(B, L, device=dev)
        plan.generation += 1
        _lib.check(_lib.load().sefd_crn_forward(
            plan.handle, ptr(engine.flat), ptr(engine.flat_buf), ptr(noisy), ptr(target), int(train),
            ptr(est_mags), ptr(target_mags), ptr(out_wav), ptr(ws), plan.ws_bytes, stream()), "crn_forward")
        ctx.engine, ctx.plan, ctx.generation = engine, plan, plan.generation
        if target_mags is None:
            target_mags = torch.zeros(0, device=dev)
        ctx.mark_non_differentiable(target_mags)
        ctx.set_materialize_grads(False)
        return est_mags, target_mags, out_wav

    @staticmethod
    def backward(ctx, g_est, _g_tgt, g_wav):
        engine, plan = ctx.engine, ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("sefd: the activation workspace of this forward was overwritten by a later forward "
                               "of the same batch shape; call backward() before the next forward")
        if g_est is None and g_wav is None:
            raise RuntimeError("sefd: backward() reached the CRN forward without any gradient")
        g_wav = None if g_wav is None else g_wav.contiguous().float()
        g_est = None if g_est is None else g_est.contiguous().float()
        _lib.check(_lib.load().sefd_crn_backward_spec(plan.handle, ptr(engine.flat), ptr(g_wav), ptr(g_est),
                                                      ptr(engine.flat_grad), ptr(plan.ws), plan.ws_bytes, stream()),
                   "crn_backward_spec")
        engine.backwards_since_step += 1
        flat = engine.flat_grad.clone()     # p.grad must never alias the buffer the next backward overwrites (AccumulateGrad
        grads = tuple(flat[o: o + n].view(shape) for (_, o, n, shape) in plan.params)   # may keep the tensor it is handed)
        return (None, None, None, None) + grads


def bump_batches_tracked(model):
    """BatchNorm2d bookkeeping after a train-mode forward: every layer's num_batches_tracked += 1 (torch semantics), one
    launch of the library's counter kernel for all layers; the device array of addresses is cached until a buffer moves."""
    bns = [m for m in model.modules() if isinstance(m, BatchNormParams)]
    if not bns:
        return
    addrs = tuple(m.num_batches_tracked.data_ptr() for m in bns)
    dev = bns[0].num_batches_tracked.device
    if dev.type != "cuda":
        raise RuntimeError("sefd: the model's buffers must live on a CUDA device")
    cache = model.__dict__.get("_nbt_cache")
    if cache is None or cache[0] != addrs:
        cache = (addrs, torch.tensor(addrs, dtype=torch.int64, device=dev))
        model.__dict__["_nbt_cache"] = cache
    _lib.check(_lib.load().sefd_counters_inc(cache[1].data_ptr(), len(addrs), 1, stream()), "counters_inc")


class Engine:
    """Owns the flat parameter / gradient / BN-statistics buffers of one DCCRN / CRN module and keeps the module's
    nn.Parameters aliased onto them."""

    def __init__(self, module, mode, family="dccrn", skip=True, real_lstm=False, cbn=False):
        self.module = module
        self.mode = mode
        self.family = family
        self.skip = skip
        self.real_lstm = real_lstm
        self.cbn = cbn
        self.plans = {}
        self.flat = self.flat_grad = self.flat_buf = None
        self._layout = Plan(1, 1 if family == "fsn" else 100, mode, family, skip, real_lstm, cbn)  # layout is independent of (B, L)
        self.param_list = None
        self.backwards_since_step = 0       # autograd backwards that wrote flat_grad since the last FlatAdam.step()

    def plan(self, B, L):
        key = (B, L)
        if key not in self.plans:
            self.plans[key] = Plan(B, L, self.mode, self.family, self.skip, self.real_lstm, self.cbn)
        return self.plans[key]

    def _named(self):
        if self.param_list is None:
            params = dict(self.module.named_parameters())
            bufs = dict(self.module.named_buffers())
            self.param_list = [(params[name], off, n, shape) for (name, off, n, shape) in self._layout.params]
            self.buf_list = [(name, off, n) for (name, off, n, _) in self._layout.buffers]
            for p, _, n, shape in self.param_list:
                assert tuple(p.shape) == tuple(shape) and p.numel() == n, (p.shape, shape)
            assert len(self.param_list) == len(params)
            self._bufs = bufs
        return self.param_list

    def sync(self):
        """Make every parameter / BN buffer a view into the flat buffers (re-flattens after .to(), etc.)."""
        plist = self._named()
        dev = plist[0][0].device
        if dev.type != "cuda":
            raise RuntimeError("sefd DCCRN runs on CUDA only (no CPU fallback): move the model with .to('cuda')")
        ok = self.flat is not None and self.flat.device == dev
        if ok:
            base = self.flat.data_ptr()
            ok = all(p.data_ptr() == base + 4 * off for p, off, _, _ in plist)
        if not ok:
            flat = torch.zeros(self._layout.n_param, device=dev)
            for p, off, n, shape in plist:
                flat[off: off + n].copy_(p.data.reshape(-1).float())
                p.data = flat[off: off + n].view(shape)
            self.flat = flat
            self.flat_grad = torch.zeros_like(flat)
        mod_bufs = dict(self.module.named_buffers())
        okb = self.flat_buf is not None and self.flat_buf.device == dev
        if okb:
            base = self.flat_buf.data_ptr()
            okb = all(mod_bufs[name].data_ptr() == base + 4 * off for name, off, _ in self.buf_list)
        if not okb:
            fb = torch.zeros(self._layout.n_buf, device=dev)
            for name, off, n in self.buf_list:
                fb[off: off + n].copy_(mod_bufs[name].reshape(-1).float())
                self._set_buffer(name, fb[off: off + n])
            self.flat_buf = fb

    def _set_buffer(self, dotted, tensor):
        mod = self.module
        parts = dotted.split(".")
        for a in parts[:-1]:
            mod = getattr(mod, a)
        mod._buffers[parts[-1]] = tensor

    def forward(self, noisy, target, train):
        self.sync()
        noisy = noisy.contiguous().float()
        if target is not None:
            target = target.contiguous().float()
        params = [p for p, _, _, _ in self.param_list]
        fn = _ForwardCRN if self.family == "crn" else _Forward
        return fn.apply(self, noisy, target, train, *params)
