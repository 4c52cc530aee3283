"""Drop-in for the on-path part of the reference's tools_for_model.py (`import tools_for_model as tools`,
trainer.py:7): the same class names, constructor arguments and tensor conventions ([B, C, F, T], real half
of the channels first), computed by libsefd.so.  Layout changes between the reference's NCHW tensors and
the library's channels-last tensors are the only thing torch does here.

Built: ConvSTFT, ConviSTFT ('complex' feature type, tools_for_model.py:36-112), ComplexConv2d and
ComplexConvTranspose2d in the DCCRN configuration (tools_for_model.py:199-338), complex_cat, Bar.
Anything else raises NotImplementedError (there is no PyTorch fallback).
"""
import sys
import time

import numpy as np

import torch
import torch.nn as nn

from sefd import dccrn as _d
from sefd import ops as _ops

EPSILON = np.finfo(np.float32).eps                 # tools_for_model.py: module constant


def init_kernels(win_len, win_inc, fft_len, win_type=None, invers=False):
    """tools_for_model.py:16-33: the windowed rFFT basis [2 F, 1, win] (or, invers=True, the transposed pseudo-inverse of the
    un-windowed basis times the window) and the window [1, win, 1], built on the host exactly as the reference builds them.  The
    CUDA kernels never read these matrices (they evaluate the same sums as FFTs, csrc/stft.cu); the function exists for code
    that wants the reference's buffers (checkpoints, analysis)."""
    if win_type == 'None' or win_type is None:
        window = np.ones(win_len)
    else:
        from scipy.signal import get_window
        window = get_window(win_type, win_len, fftbins=True)
    basis = np.fft.rfft(np.eye(fft_len))[:win_len]
    kernel = np.concatenate([np.real(basis), np.imag(basis)], 1).T
    if invers:
        kernel = np.linalg.pinv(kernel).T
    kernel = (kernel * window)[:, None, :]
    return torch.from_numpy(kernel.astype(np.float32)), torch.from_numpy(window[None, :, None].astype(np.float32))



def _check_stft(win_len, win_inc, fft_len, win_type, feature_type):
    if (win_len, win_inc, fft_len) != (400, 100, 512) or win_type not in ("hanning", "hann") or feature_type != "complex":
        raise NotImplementedError("sefd STFT kernels are built for win 400 / hop 100 / fft 512 / Hann / 'complex'")


class ConvSTFT(nn.Module):
    """tools_for_model.py:36-68.  forward(inputs [B, L] or [B, 1, L]) -> [B, 514, T] (real rows then imag rows)."""

    def __init__(self, win_len, win_inc, fft_len=None, win_type="hamming", feature_type="real", fix=True):
        super().__init__()
        _check_stft(win_len, win_inc, fft_len, win_type, feature_type)
        self.register_buffer("weight", _d.STFTBuffers(win_len, fft_len, inverse=False).weight)
        self.feature_type, self.stride, self.win_len, self.dim = feature_type, win_inc, win_len, fft_len

    def forward(self, inputs):
        if inputs.dim() == 3:
            inputs = inputs.squeeze(1)
        spec = _ops.stft(inputs.float())                                  # [B, 257, T, 2]
        return torch.cat([spec[..., 0], spec[..., 1]], 1)


class _ISTFT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, L):
        return _ops.istft(spec, L)

    @staticmethod
    def backward(ctx, g):
        return _ops.istft_backward(g), None


class ConviSTFT(nn.Module):
    """tools_for_model.py:71-112.  forward(inputs [B, 514, T]) -> [B, 1, L]."""

    def __init__(self, win_len, win_inc, fft_len=None, win_type="hamming", feature_type="real", fix=True):
        super().__init__()
        _check_stft(win_len, win_inc, fft_len, win_type, feature_type)
        b = _d.STFTBuffers(win_len, fft_len, inverse=True)
        self.register_buffer("weight", b.weight)
        self.register_buffer("window", b.window)
        self.register_buffer("enframe", b.enframe)
        self.feature_type, self.win_type, self.win_len, self.stride, self.dim = feature_type, win_type, win_len, win_inc, fft_len

    def forward(self, inputs, phase=None):
        if phase is not None:
            raise NotImplementedError("sefd ConviSTFT: magnitude/phase input is not on the built path")
        B, _, T = inputs.shape
        spec = torch.stack([inputs[:, :257], inputs[:, 257:]], -1).contiguous()
        return _ISTFT.apply(spec, (T - 3) * 100).unsqueeze(1)


def complex_cat(inputs, axis):
    """tools_for_model.py:184-193 (pure data movement)."""
    real, imag = [], []
    for data in inputs:
        r, i = torch.chunk(data, 2, axis)
        real.append(r)
        imag.append(i)
    return torch.cat(real + imag, axis)


def _to_cl(x):      # [B, C, F, T] -> [B, F, T, C]
    return x.permute(0, 2, 3, 1).contiguous()


def _from_cl(x):    # [B, F, T, C] -> [B, C, F, T]
    return x.permute(0, 3, 1, 2)


class _CConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wr, br, wi, bi):
        xc = _to_cl(x.float())
        ctx.save_for_backward(xc, wr, wi)
        return _from_cl(_ops.cconv2d_forward(xc, wr.contiguous(), br.contiguous(), wi.contiguous(), bi.contiguous()))

    @staticmethod
    def backward(ctx, g):
        xc, wr, wi = ctx.saved_tensors
        dx, dwr, dbr, dwi, dbi = _ops.cconv2d_backward(xc, wr.contiguous(), wi.contiguous(), _to_cl(g))
        return _from_cl(dx), dwr, dbr, dwi, dbi


class _CConvT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wr, br, wi, bi):
        # split the already complex_cat'ed input into the library's two sources [r0 | i0], [r1 | i1]
        B, C, F, T = x.shape
        h = C // 4
        r, i = x[:, : C // 2], x[:, C // 2:]
        x0 = _to_cl(torch.cat([r[:, :h], i[:, :h]], 1).float())
        x1 = _to_cl(torch.cat([r[:, h:], i[:, h:]], 1).float())
        ctx.save_for_backward(x0, x1, wr, wi)
        return _from_cl(_ops.cconvT2d_forward(x0, x1, wr.contiguous(), br.contiguous(), wi.contiguous(), bi.contiguous()))

    @staticmethod
    def backward(ctx, g):
        x0, x1, wr, wi = ctx.saved_tensors
        dx0, dx1, dwr, dbr, dwi, dbi = _ops.cconvT2d_backward(x0, x1, wr.contiguous(), wi.contiguous(), _to_cl(g))
        d0, d1 = _from_cl(dx0), _from_cl(dx1)
        h = d0.shape[1] // 2
        dx = torch.cat([d0[:, :h], d1[:, :h], d0[:, h:], d1[:, h:]], 1)
        return dx, dwr, dbr, dwi, dbi


def _dccrn_geometry(kernel_size, stride, padding, extra_ok):
    if tuple(kernel_size) != (5, 2) or tuple(stride) != (2, 1) or not extra_ok:
        raise NotImplementedError("sefd complex conv kernels are built for kernel (5,2), stride (2,1), "
                                  "padding (2,1|0) [+ output_padding (1,0)] as used by DCCRN")


class ComplexConv2d(nn.Module):
    """tools_for_model.py:199-269 in the DCCRN geometry (causal: one zero frame on the left of T)."""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0), dilation=1,
                 groups=1, causal=True, complex_axis=1):
        super().__init__()
        _dccrn_geometry(kernel_size, stride, padding,
                        tuple(padding) == (2, 1) and causal and dilation == 1 and groups == 1 and complex_axis == 1)
        p = _d.ComplexConvParams(in_channels, out_channels, transposed=False)
        self.real_conv, self.imag_conv = p.real_conv, p.imag_conv

    def forward(self, inputs):
        return _CConv.apply(inputs, self.real_conv.weight, self.real_conv.bias, self.imag_conv.weight, self.imag_conv.bias)


class ComplexConvTranspose2d(nn.Module):
    """tools_for_model.py:272-338 in the DCCRN geometry; output has T+1 frames like the reference."""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 output_padding=(0, 0), causal=False, complex_axis=1, groups=1):
        super().__init__()
        _dccrn_geometry(kernel_size, stride, padding,
                        tuple(padding) == (2, 0) and tuple(output_padding) == (1, 0) and groups == 1
                        and complex_axis == 1 and in_channels % 4 == 0)
        p = _d.ComplexConvParams(in_channels, out_channels, transposed=True)
        self.real_conv, self.imag_conv = p.real_conv, p.imag_conv

    def forward(self, inputs):
        return _CConvT.apply(inputs, self.real_conv.weight, self.real_conv.bias, self.imag_conv.weight, self.imag_conv.bias)


class ComplexBatchNorm(nn.Module):
    """tools_for_model.py:430-603 as a stand-alone layer on [B, C, F, T] (C = num_features channels: the real parts of the C / 2
    complex features in the first half of the channel axis, the imaginary parts in the second): same parameter / buffer names,
    initial values and RNG use as the reference.
    Inside models.DCCRN(use_cbn=True) the layer is fused with the PReLU that follows it (csrc/cbn.cu)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, complex_axis=1):
        super().__init__()
        if (eps, momentum, affine, track_running_stats, complex_axis) != (1e-5, 0.1, True, True, 1) or num_features % 8:
            raise NotImplementedError("sefd ComplexBatchNorm: built for eps 1e-5, momentum 0.1, affine, running statistics, "
                                      "complex_axis 1 and a multiple of 8 features")
        p = _d.ComplexBatchNormParams(num_features)
        self.num_features = num_features // 2
        for k in ("Wrr", "Wri", "Wii", "Br", "Bi"):
            setattr(self, k, getattr(p, k))
        for k in ("RMr", "RMi", "RVrr", "RVri", "RVii", "num_batches_tracked"):
            self.register_buffer(k, getattr(p, k))

    def forward(self, inputs):
        x = _to_cl(inputs)
        B, F, T, C = x.shape
        w3h = torch.cat([self.Wrr, self.Wri, self.Wii])
        b2h = torch.cat([self.Br, self.Bi])
        running = torch.cat([self.RMr, self.RMi, self.RVrr, self.RVri, self.RVii])
        z = _ops.complex_batch_norm(x.reshape(-1, C), w3h, b2h, running, self.training)
        if self.training:
            with torch.no_grad():
                h = C // 2
                for i, k in enumerate(("RMr", "RMi", "RVrr", "RVri", "RVii")):
                    getattr(self, k).copy_(running[i * h:(i + 1) * h])
                self.num_batches_tracked += 1
        return _from_cl(z.reshape(B, F, T, C))


class Bar(object):
    """Progress iterator used by the reference trainer loops (`for inputs, targets in tools.Bar(loader)`,
    trainer.py:23).  Own implementation: prints count, rate and ETA on one line."""

    def __init__(self, dataloader, width=30, out=sys.stdout):
        self.it = iter(dataloader)
        try:
            self.total = len(dataloader)
        except TypeError:
            self.total = None
        self.n, self.t0, self.width, self.out = 0, time.time(), width, out

    def __len__(self):
        return self.total if self.total is not None else 0

    def __iter__(self):
        return self

    def __next__(self):
        try:
            item = next(self.it)
        except StopIteration:
            if self.n:
                self.out.write("\n")
            raise
        self.n += 1
        el = time.time() - self.t0
        if self.total:
            done = int(self.width * self.n / self.total)
            eta = el / self.n * (self.total - self.n)
            self.out.write(f"\r[{'=' * done}{' ' * (self.width - done)}] {self.n}/{self.total} "
                           f"{el:6.1f}s eta {eta:6.1f}s")
        else:
            self.out.write(f"\r{self.n} it {el:6.1f}s")
        self.out.flush()
        return item


def stft(y, n_fft=512, hop_length=300, win_length=400):
    """tools_for_model.py:628-648 (torch.stft, centred, reflect padding, periodic Hann): [B, L] -> complex [B, 257, T]."""
    if (n_fft, hop_length, win_length) != (512, 300, 400):
        raise NotImplementedError("sefd: tools.stft is built for the reference geometry n_fft 512 / hop 300 / win 400")
    return _ops.fsn_stft(y.float().contiguous())


def mag_phase(complex_tensor):
    """tools_for_model.py:682-683."""
    return _ops.fsn_mag_phase(complex_tensor)


def build_complex_ideal_ratio_mask(noisy, clean):
    """tools_for_model.py:686-705 (+ compress_cIRM :708-717): complex [B, F, T] x2 -> [B, F, T, 2]."""
    return _ops.fsn_cirm(noisy, clean)


def compress_cIRM(mask, K=10, C=0.1):
    """tools_for_model.py:707-717: (-inf, +inf) -> [-K, K].  Tensors run the library's kernel; numpy arrays take the
    reference's own numpy branch (host data stays on the host)."""
    if not torch.is_tensor(mask):
        mask = -100 * (mask <= -100) + mask * (mask > -100)
        return K * (1 - np.exp(-C * mask)) / (1 + np.exp(-C * mask))
    if (K, C) != (10, 0.1):
        raise NotImplementedError("sefd: compress_cIRM is built for K = 10, C = 0.1")
    return _ops.fsn_compress_cirm(mask)


def decompress_cIRM(mask, K=10, limit=9.9):
    """tools_for_model.py:720-723."""
    if (K, limit) != (10, 9.9):
        raise NotImplementedError("sefd: decompress_cIRM is built for K = 10, limit = 9.9")
    return _ops.fsn_decompress_cirm(mask)


def fullsubnet_features(noisy_wav, clean_wav):
    """The four feature / target calls of trainer.fullsubnet_train (trainer.py:97-104) fused into one kernel:
    returns (noisy_mag [B, 257, T], cIRM [B, 257, T, 2])."""
    return _ops.fsn_features(noisy_wav.float().contiguous(), clean_wav.float().contiguous())


def istft(features, n_fft=512, hop_length=300, win_length=400, length=None, use_mag_phase=False):
    """tools_for_model.py:651-679 (torch.istft, centred, periodic Hann): [B, 257, T] complex (or [B, 257, T, 2], or
    (mag, phase) with use_mag_phase) -> [B, length]."""
    if (n_fft, hop_length, win_length) != (512, 300, 400):
        raise NotImplementedError("sefd: tools.istft is built for the reference geometry n_fft 512 / hop 300 / win 400")
    return _ops.fsn_istft(features, length=length, use_mag_phase=use_mag_phase)
