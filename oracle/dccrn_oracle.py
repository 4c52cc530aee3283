"""TEST INFRASTRUCTURE ONLY — CPU oracle for the DCCRN train-step path.

A functional (state-dict in, tensors out) torch-CPU restatement of what the reference
computes on the path  wave -> STFT -> 6x(complex conv + BN + PReLU) -> 2x complex LSTM
-> 6x(complex convT + BN + PReLU) -> mask -> ISTFT -> clamp -> loss,  with gradients
from torch autograd and the reference's Adam settings.  Every function cites the
reference file:line (relative to /root/reference) it restates.

Parity status: PINNED.  tests/golden/make_golden.py imports the unmodified reference
in the build container, runs it on seeded inputs and stores outputs / loss / per-parameter
gradient norms under tests/golden/; tests/test_oracle_golden.py checks this module
against those fixtures (and against the si_sdr doctest values of tools_for_loss.py:57-74).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this module.  The product never does: it fails loudly without its CUDA library.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# config.py:50-61 (the values BASELINE.json's configs are quoted on)
KERNEL_NUM = [32, 64, 128, 256, 256, 256]
WIN_LEN, WIN_INC, FFT_LEN = 400, 100, 512
RNN_LAYERS, RNN_UNITS = 2, 256
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


# --------------------------------------------------------------------------------------
# STFT / ISTFT bases  (tools_for_model.py:16-33)
# --------------------------------------------------------------------------------------
def periodic_hann(win_len: int) -> np.ndarray:
    """scipy.signal.get_window('hann', win_len, fftbins=True)  (tools_for_model.py:20)."""
    n = np.arange(win_len, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_len)


def stft_bases(win_len: int = WIN_LEN, fft_len: int = FFT_LEN) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Analysis kernel K_a[2F, win] and synthesis kernel K_s[2F, win] in float64.

    tools_for_model.py:22-31: rows of rfft(eye(N)) truncated to the first `win_len`
    samples; synthesis = pinv of the un-windowed analysis kernel, transposed, then both
    are multiplied by the window.
    """
    w = periodic_hann(win_len)
    n = np.arange(win_len, dtype=np.float64)[None, :]
    k = np.arange(fft_len // 2 + 1, dtype=np.float64)[:, None]
    ang = 2.0 * np.pi * k * n / fft_len
    kern = np.concatenate([np.cos(ang), -np.sin(ang)], 0)          # [2F, win]
    synth = np.linalg.pinv(kern).T                                 # pinv([2F,win]) is [win,2F]; .T -> [2F,win]
    return kern * w, synth * w, w


def conv_stft(wav: torch.Tensor, k_a: torch.Tensor, win_len: int = WIN_LEN, hop: int = WIN_INC) -> torch.Tensor:
    """ConvSTFT.forward, 'complex' feature type (tools_for_model.py:54-61). wav [B,L] -> [B,2F,T]."""
    x = F.pad(wav[:, None, :], [win_len - hop, win_len - hop])
    return F.conv1d(x, k_a[:, None, :], stride=hop)


def conv_istft(spec: torch.Tensor, k_s: torch.Tensor, window: torch.Tensor,
               win_len: int = WIN_LEN, hop: int = WIN_INC) -> torch.Tensor:
    """ConviSTFT.forward (tools_for_model.py:90-112). spec [B,2F,T] -> [B,L]."""
    out = F.conv_transpose1d(spec, k_s[:, None, :], stride=hop)
    t = (window.reshape(1, -1, 1).repeat(1, 1, spec.shape[-1])) ** 2
    eye = torch.eye(win_len, dtype=spec.dtype)[:, None, :]
    coff = F.conv_transpose1d(t, eye, stride=hop)
    out = out / (coff + 1e-8)
    return out[..., win_len - hop: -(win_len - hop)].squeeze(1)


# --------------------------------------------------------------------------------------
# parameter construction in the reference's RNG order (models.py:17-170)
# --------------------------------------------------------------------------------------
def init_state(seed: int = 0, kernel_num: Optional[List[int]] = None, skip_type: bool = True,
               lstm: str = "complex", use_cbn: bool = False) -> Dict[str, torch.Tensor]:
    """State dict with the reference's keys/shapes and, for a given torch seed, its values.

    Built by instantiating torch.nn modules in the same order as DCCRN.__init__
    (encoder convs models.py:63-80, complex LSTMs :83-95, decoder convTs :107-137) so the
    CPU RNG stream is consumed identically (tools_for_model.py:233-241, 299-311, 147-158).
    """
    import torch.nn as nn
    kn = [2] + list(kernel_num or KERNEL_NUM)
    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def cplx(prefix, mod_cls, cin, cout, **kw):
        rc = mod_cls(cin // 2, cout // 2, (5, 2), (2, 1), **kw)
        ic = mod_cls(cin // 2, cout // 2, (5, 2), (2, 1), **kw)
        nn.init.normal_(rc.weight.data, std=0.05)
        nn.init.normal_(ic.weight.data, std=0.05)
        sd[prefix + "real_conv.weight"] = rc.weight.data.clone()
        sd[prefix + "real_conv.bias"] = torch.zeros_like(rc.bias.data)
        sd[prefix + "imag_conv.weight"] = ic.weight.data.clone()
        sd[prefix + "imag_conv.bias"] = torch.zeros_like(ic.bias.data)

    def bn_prelu(prefix_bn, prefix_act, c):
        if use_cbn:                                      # ComplexBatchNorm(c) (tools_for_model.py:430-491): c // 2 complex features;
            h = c // 2                                   # reset_parameters draws Wri ~ U(-0.9, 0.9) from the global RNG
            sd[prefix_bn + "Wrr"] = torch.ones(h)
            sd[prefix_bn + "Wri"] = torch.empty(h).uniform_(-.9, +.9)
            sd[prefix_bn + "Wii"] = torch.ones(h)
            sd[prefix_bn + "Br"] = torch.zeros(h)
            sd[prefix_bn + "Bi"] = torch.zeros(h)
            sd[prefix_bn + "RMr"] = torch.zeros(h)
            sd[prefix_bn + "RMi"] = torch.zeros(h)
            sd[prefix_bn + "RVrr"] = torch.ones(h)
            sd[prefix_bn + "RVri"] = torch.zeros(h)
            sd[prefix_bn + "RVii"] = torch.ones(h)
            sd[prefix_bn + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
            sd[prefix_act + "weight"] = torch.full((1,), 0.25)
            return
        sd[prefix_bn + "weight"] = torch.ones(c)
        sd[prefix_bn + "bias"] = torch.zeros(c)
        sd[prefix_bn + "running_mean"] = torch.zeros(c)
        sd[prefix_bn + "running_var"] = torch.ones(c)
        sd[prefix_bn + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        sd[prefix_act + "weight"] = torch.full((1,), 0.25)

    for i in range(len(kn) - 1):
        cplx(f"encoder.{i}.0.", nn.Conv2d, kn[i], kn[i + 1], padding=[2, 0])
        bn_prelu(f"encoder.{i}.1.", f"encoder.{i}.2.", kn[i + 1])

    hidden_dim = FFT_LEN // (2 ** len(kn))
    if lstm == "real":                                   # cfg.lstm == 'real' (models.py:96-105): one 2-layer nn.LSTM + Linear
        m = nn.LSTM(input_size=hidden_dim * kn[-1], hidden_size=RNN_UNITS, num_layers=2, dropout=0.0)
        for name, p in m.named_parameters():
            sd[f"enhance.{name}"] = p.data.clone()
        m = nn.Linear(RNN_UNITS, hidden_dim * kn[-1])
        sd["tranform.weight"] = m.weight.data.clone()
        sd["tranform.bias"] = m.bias.data.clone()
    for l in range(RNN_LAYERS if lstm != "real" else 0):
        in_sz = (hidden_dim * kn[-1] if l == 0 else RNN_UNITS) // 2
        for part in ("real", "imag"):
            m = nn.LSTM(in_sz, RNN_UNITS // 2, num_layers=1)
            for name, p in m.named_parameters():
                sd[f"enhance.{l}.{part}_lstm.{name}"] = p.data.clone()
        if l == RNN_LAYERS - 1:
            for part in ("r", "i"):
                m = nn.Linear(RNN_UNITS // 2, hidden_dim * kn[-1] // 2)
                sd[f"enhance.{l}.{part}_trans.weight"] = m.weight.data.clone()
                sd[f"enhance.{l}.{part}_trans.bias"] = m.bias.data.clone()

    j = 0
    for idx in range(len(kn) - 1, 0, -1):
        cplx(f"decoder.{j}.0.", nn.ConvTranspose2d, kn[idx] * (2 if skip_type else 1), kn[idx - 1],   # models.py:107-169
             padding=(2, 0), output_padding=(1, 0))
        if idx != 1:
            bn_prelu(f"decoder.{j}.1.", f"decoder.{j}.2.", kn[idx - 1])
        j += 1

    k_a, k_s, w = stft_bases()
    sd["stft.weight"] = torch.from_numpy(k_a.astype(np.float32))[:, None, :]
    sd["istft.weight"] = torch.from_numpy(k_s.astype(np.float32))[:, None, :]
    sd["istft.window"] = torch.from_numpy(w.astype(np.float32))[None, :, None]
    sd["istft.enframe"] = torch.eye(WIN_LEN)[:, None, :]
    return sd


def trainable_keys(sd: Dict[str, torch.Tensor]) -> List[str]:
    """Keys that are nn.Parameters in the reference (everything but BN/STFT buffers)."""
    skip = ("running_mean", "running_var", "num_batches_tracked", ".RMr", ".RMi", ".RVrr", ".RVri", ".RVii")
    return [k for k in sd if not k.endswith(skip) and not k.startswith(("stft.", "istft."))]


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------
def complex_conv2d(x, wr, br, wi, bi):
    """ComplexConv2d.forward (tools_for_model.py:243-269): causal pad 1 on T, 4 real convs."""
    x = F.pad(x, [1, 0, 0, 0])
    real, imag = torch.chunk(x, 2, 1)
    kw = dict(stride=(2, 1), padding=(2, 0))
    r2r = F.conv2d(real, wr, br, **kw)
    i2i = F.conv2d(imag, wi, bi, **kw)
    r2i = F.conv2d(real, wi, bi, **kw)
    i2r = F.conv2d(imag, wr, br, **kw)
    return torch.cat([r2r - i2i, r2i + i2r], 1)


def complex_conv_transpose2d(x, wr, br, wi, bi):
    """ComplexConvTranspose2d.forward (tools_for_model.py:313-338)."""
    real, imag = torch.chunk(x, 2, 1)
    kw = dict(stride=(2, 1), padding=(2, 0), output_padding=(1, 0))
    r2r = F.conv_transpose2d(real, wr, br, **kw)
    i2i = F.conv_transpose2d(imag, wi, bi, **kw)
    r2i = F.conv_transpose2d(real, wi, bi, **kw)
    i2r = F.conv_transpose2d(imag, wr, br, **kw)
    return torch.cat([r2r - i2i, r2i + i2r], 1)


def batch_norm_train(x, gamma, beta):
    """nn.BatchNorm2d in train mode (models.py:76): biased batch variance, eps 1e-5.
    Returns (y, mean, biased_var)."""
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    y = (x - mean[None, :, None, None]) * torch.rsqrt(var + BN_EPS)[None, :, None, None]
    return y * gamma[None, :, None, None] + beta[None, :, None, None], mean, var


def batch_norm_eval(x, gamma, beta, rmean, rvar):
    s = gamma * torch.rsqrt(rvar + BN_EPS)
    return (x - rmean[None, :, None, None]) * s[None, :, None, None] + beta[None, :, None, None]


def complex_batch_norm(x, Wrr, Wri, Wii, Br, Bi, stats=None):
    """ComplexBatchNorm.forward (tools_for_model.py:492-599) on x [B, C, F, T] whose first C/2 channels are the real and the
    last C/2 the imaginary parts: 2x2 whitening with the inverse square root of the covariance (eps 1e-5 on its diagonal),
    then the affine map [[Wrr, Wri], [Wri, Wii]] and bias.  stats = None: batch statistics (train mode), returned as
    (Mr, Mi, Vrr, Vri, Vii) WITHOUT eps - what the reference lerps into its running buffers; stats = the five running buffers:
    eval mode.  The reference line 567 calls torch.addcmul with the pre-1.5 positional `value` (addcmul(t, -1, a, b) =
    t - a * b), which current torch rejects; the arithmetic restated here is that expression."""
    xr, xi = torch.chunk(x, 2, dim=1)
    v = lambda t: t[None, :, None, None]
    if stats is None:
        Mr, Mi = xr.mean(dim=(0, 2, 3)), xi.mean(dim=(0, 2, 3))
    else:
        Mr, Mi = stats[0], stats[1]
    xr, xi = xr - v(Mr), xi - v(Mi)
    if stats is None:
        Vrr, Vri, Vii = (xr * xr).mean(dim=(0, 2, 3)), (xr * xi).mean(dim=(0, 2, 3)), (xi * xi).mean(dim=(0, 2, 3))
        out_stats = (Mr, Mi, Vrr, Vri, Vii)
    else:
        Vrr, Vri, Vii = stats[2], stats[3], stats[4]
        out_stats = None
    Vrr_e, Vii_e = Vrr + BN_EPS, Vii + BN_EPS
    tau = Vrr_e + Vii_e
    delta = Vrr_e * Vii_e - Vri * Vri
    s = delta.sqrt()
    t = (tau + 2 * s).sqrt()
    rst = (s * t).reciprocal()
    Urr, Uii, Uri = (s + Vii_e) * rst, (s + Vrr_e) * rst, -Vri * rst
    Zrr, Zri = Wrr * Urr + Wri * Uri, Wrr * Uri + Wri * Uii
    Zir, Zii = Wri * Urr + Wii * Uri, Wri * Uri + Wii * Uii
    yr = v(Zrr) * xr + v(Zri) * xi + v(Br)
    yi = v(Zir) * xr + v(Zii) * xi + v(Bi)
    return torch.cat([yr, yi], 1), out_stats


def prelu(x, alpha):
    """nn.PReLU() with one shared slope (models.py:78)."""
    return torch.where(x > 0, x, alpha * x)


def lstm_seq(x, w_ih, w_hh, b_ih, b_hh):
    """Single-layer unidirectional nn.LSTM, zero initial state, gate order i,f,g,o. x [T,N,I] -> [T,N,H]."""
    T, N, _ = x.shape
    H = w_hh.shape[1]
    pre = x @ w_ih.t() + (b_ih + b_hh)
    h = x.new_zeros(N, H)
    c = x.new_zeros(N, H)
    outs = []
    for t in range(T):
        g = pre[t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 0)


# The explicit per-step loop above is the restatement; the library routine below is the same arithmetic as
# torch.nn.LSTM.forward (what the reference calls, tools_for_model.py:147-150) and is what the timed CPU arm
# uses so that the baseline costs what the reference costs.  tests/test_oracle_golden.py checks they agree.
USE_FUSED_LSTM = True


def lstm_fused(x, w_ih, w_hh, b_ih, b_hh):
    N, H = x.shape[1], w_hh.shape[1]
    h0 = x.new_zeros(1, N, H)
    out, _, _ = torch._VF.lstm(x, (h0, h0.clone()), [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, False, False, False)
    return out


def complex_lstm(real, imag, sd, prefix, project):
    """NavieComplexLSTM.forward (tools_for_model.py:162-177)."""
    def run(part, x):
        f = lstm_fused if USE_FUSED_LSTM else lstm_seq
        return f(x, sd[f"{prefix}{part}_lstm.weight_ih_l0"], sd[f"{prefix}{part}_lstm.weight_hh_l0"],
                 sd[f"{prefix}{part}_lstm.bias_ih_l0"], sd[f"{prefix}{part}_lstm.bias_hh_l0"])
    r2r, r2i, i2r, i2i = run("real", real), run("imag", real), run("real", imag), run("imag", imag)
    ro, io = r2r - i2i, i2r + r2i
    if project:
        ro = F.linear(ro, sd[prefix + "r_trans.weight"], sd[prefix + "r_trans.bias"])
        io = F.linear(io, sd[prefix + "i_trans.weight"], sd[prefix + "i_trans.bias"])
    return ro, io


def complex_cat(a, b):
    """complex_cat([a, b], 1) (tools_for_model.py:184-193)."""
    ar, ai = torch.chunk(a, 2, 1)
    br, bi = torch.chunk(b, 2, 1)
    return torch.cat([ar, br, ai, bi], 1)


# --------------------------------------------------------------------------------------
# model forward (models.py:176-284)
# --------------------------------------------------------------------------------------
def dccrn_forward(sd: Dict[str, torch.Tensor], wav: torch.Tensor, masking_mode: str = "C",
                  train: bool = True, taps: Optional[dict] = None):
    """Returns (out_real [B,F,T], out_imag [B,F,T], out_wav [B,L]).

    `taps`, if a dict, is filled with intermediates: 'spec', 'enc{i}_conv', 'enc{i}',
    'lstm_in_r', 'lstm{l}_r/i', 'dec{j}_conv', 'dec{j}', 'mask', 'bn_stats'.
    In train mode new BN running stats are returned through taps['bn_stats'] (the state
    dict itself is not mutated).
    """
    nF = FFT_LEN // 2 + 1
    k_a = sd["stft.weight"][:, 0, :].to(wav.dtype)
    k_s = sd["istft.weight"][:, 0, :].to(wav.dtype)
    win = sd["istft.window"].reshape(-1).to(wav.dtype)
    n_layers = len([k for k in sd if k.startswith("encoder.") and k.endswith(".0.real_conv.weight")])

    specs = conv_stft(wav, k_a)                                        # models.py:177
    real, imag = specs[:, :nF], specs[:, nF:]
    out = torch.stack([real, imag], 1)[:, :, 1:]                       # models.py:183-184 (DC dropped)
    if taps is not None:
        taps["spec"] = specs
        taps["bn_stats"] = {}

    def norm_act(x, pbn, pact):
        if pbn + "Wrr" in sd:                                          # use_cbn = True (models.py:76, 120, 151)
            aff = [sd[pbn + k] for k in ("Wrr", "Wri", "Wii", "Br", "Bi")]
            if train:
                y, st = complex_batch_norm(x, *aff)
                if taps is not None:
                    taps["bn_stats"][pbn] = tuple(t.detach() for t in st)
            else:
                y, _ = complex_batch_norm(x, *aff, stats=[sd[pbn + k] for k in ("RMr", "RMi", "RVrr", "RVri", "RVii")])
            return prelu(y, sd[pact + "weight"])
        if train:
            y, m, v = batch_norm_train(x, sd[pbn + "weight"], sd[pbn + "bias"])
            if taps is not None:
                n = x.numel() // x.shape[1]
                taps["bn_stats"][pbn] = (m.detach(), v.detach() * n / max(n - 1, 1))
        else:
            y = batch_norm_eval(x, sd[pbn + "weight"], sd[pbn + "bias"],
                                sd[pbn + "running_mean"], sd[pbn + "running_var"])
        return prelu(y, sd[pact + "weight"])

    enc_out = []
    for i in range(n_layers):                                          # models.py:195-198
        p = f"encoder.{i}.0."
        out = complex_conv2d(out, sd[p + "real_conv.weight"], sd[p + "real_conv.bias"],
                             sd[p + "imag_conv.weight"], sd[p + "imag_conv.bias"])
        if taps is not None:
            taps[f"enc{i}_conv"] = out
        out = norm_act(out, f"encoder.{i}.1.", f"encoder.{i}.2.")
        if taps is not None:
            taps[f"enc{i}"] = out
        enc_out.append(out)

    B, C, D, T = out.shape                                             # models.py:200-220
    o = out.permute(3, 0, 1, 2)
    if "enhance.weight_ih_l0" in sd:                                   # cfg.lstm == 'real' (models.py:213-218)
        x = o.reshape(T, B, C * D)
        lstm = lstm_fused if USE_FUSED_LSTM else lstm_seq
        for l in range(2):
            x = lstm(x, sd[f"enhance.weight_ih_l{l}"], sd[f"enhance.weight_hh_l{l}"], sd[f"enhance.bias_ih_l{l}"],
                     sd[f"enhance.bias_hh_l{l}"])
        if taps is not None:
            taps["lstm_real"] = x
        x = F.linear(x, sd["tranform.weight"], sd["tranform.bias"])
        out = x.reshape(T, B, C, D).permute(1, 2, 3, 0)
    else:                                                              # cfg.lstm == 'complex' (models.py:201-212)
        r = o[:, :, :C // 2].reshape(T, B, C // 2 * D)
        i_ = o[:, :, C // 2:].reshape(T, B, C // 2 * D)
        if taps is not None:
            taps["lstm_in_r"], taps["lstm_in_i"] = r, i_
        for l in range(RNN_LAYERS):
            r, i_ = complex_lstm(r, i_, sd, f"enhance.{l}.", project=(l == RNN_LAYERS - 1))
            if taps is not None:
                taps[f"lstm{l}_r"], taps[f"lstm{l}_i"] = r, i_
        r = r.reshape(T, B, C // 2, D)
        i_ = i_.reshape(T, B, C // 2, D)
        out = torch.cat([r, i_], 2).permute(1, 2, 3, 0)

    for j in range(n_layers):                                          # models.py:222-226
        p = f"decoder.{j}.0."
        if sd[p + "real_conv.weight"].shape[0] == out.shape[1]:      # cfg.skip_type (models.py:222-230): the
            out = complex_cat(out, enc_out[-1 - j])                     # weights carry 2x the input channels
        out = complex_conv_transpose2d(out, sd[p + "real_conv.weight"], sd[p + "real_conv.bias"],
                                       sd[p + "imag_conv.weight"], sd[p + "imag_conv.bias"])
        if taps is not None:
            taps[f"dec{j}_conv"] = out                                  # T+1 frames
        if j != n_layers - 1:
            # NB: the Sequential(convT, BN, PReLU) runs on all T+1 frames; the look-ahead
            # frame 0 is dropped only afterwards (models.py:225-226), so it is part of the
            # BN batch statistics and receives gradient through them.
            out = norm_act(out, f"decoder.{j}.1.", f"decoder.{j}.2.")
        out = out[..., 1:]
        if taps is not None:
            taps[f"dec{j}"] = out

    mask_real = F.pad(out[:, 0], [0, 0, 1, 0])                         # models.py:253-256
    mask_imag = F.pad(out[:, 1], [0, 0, 1, 0])
    if taps is not None:
        taps["mask"] = torch.stack([mask_real, mask_imag], 1)

    if masking_mode == "E":                                            # models.py:258-272
        spec_mags = torch.sqrt(real ** 2 + imag ** 2 + 1e-8)
        spec_phase = torch.atan2(imag, real)
        mask_mags = (mask_real ** 2 + mask_imag ** 2) ** 0.5
        real_phase = mask_real / (mask_mags + 1e-8)
        imag_phase = mask_imag / (mask_mags + 1e-8)
        mask_phase = torch.atan2(imag_phase, real_phase)
        est_mags = torch.tanh(mask_mags) * spec_mags
        est_phase = spec_phase + mask_phase
        out_real = est_mags * torch.cos(est_phase)
        out_imag = est_mags * torch.sin(est_phase)
    elif masking_mode == "C":                                          # models.py:273-274
        out_real = real * mask_real - imag * mask_imag
        out_imag = real * mask_imag + imag * mask_real
    elif masking_mode == "R":                                          # models.py:275-276
        out_real, out_imag = real * mask_real, imag * mask_imag
    elif masking_mode == "Direct(None make)":                          # models.py:238-243: spectral mapping
        out_real, out_imag = mask_real, mask_imag
    else:
        raise ValueError(masking_mode)

    out_wav = conv_istft(torch.cat([out_real, out_imag], 1), k_s, win)  # models.py:278-282
    out_wav = torch.clamp(out_wav, -1, 1)
    return out_real, out_imag, out_wav


# --------------------------------------------------------------------------------------
# losses (tools_for_loss.py:17-94; selection models.py:315-323)
# --------------------------------------------------------------------------------------
def _dot(a, b):
    return torch.sum(a * b, -1, keepdim=True)


def si_snr(s1, s2, eps=1e-8):
    """tools_for_loss.py:36-44 (s1 = estimate, s2 = target; no mean removal)."""
    a = _dot(s1, s2) / (_dot(s2, s2) + eps)
    tgt = a * s2
    noise = s1 - tgt
    return torch.mean(10 * torch.log10(_dot(tgt, tgt) / (_dot(noise, noise) + eps) + eps))


def sdr(s1, s2, eps=1e-8):
    """tools_for_loss.py:29-33 (energies are squared)."""
    sn = _dot(s1, s1)
    d = _dot(s1 - s2, s1 - s2)
    return torch.mean(10 * torch.log10(sn ** 2 / (d ** 2 + eps)))


def si_sdr(reference, estimation, eps=1e-8):
    """tools_for_loss.py:80-94 (batch mean of the ratio before the log)."""
    e = torch.sum(reference ** 2, -1, keepdim=True)
    a = torch.sum(reference * estimation, -1, keepdim=True) / e + eps
    proj = a * reference
    noise = estimation - proj
    ratio = torch.sum(proj ** 2, -1) / torch.sum(noise ** 2, -1) + eps
    return 10 * torch.log10(torch.mean(ratio) + eps)


def dccrn_loss(estimated, target, loss: str = "SI-SNR"):
    """DCCRN.loss, non-perceptual branch (models.py:315-323)."""
    if loss == "MSE":
        return F.mse_loss(estimated, target)
    if loss == "SDR":
        return -sdr(target, estimated)
    if loss == "SI-SNR":
        return -si_snr(estimated, target)
    if loss == "SI-SDR":
        return -si_sdr(target, estimated)
    raise ValueError(loss)


# --------------------------------------------------------------------------------------
# LMS perceptual loss (tools_for_loss.py:111-249)
# --------------------------------------------------------------------------------------
def mel_filterbank(num_coeffs: int, fft_size: int = FFT_LEN, fs: int = 16000) -> np.ndarray:
    """melFilterBank (tools_for_loss.py:144-188): [num_coeffs, fft_size/2+1] triangular filters whose edges are
    floor(bins * f / (fs/2)) of mel-spaced centre frequencies held in a float32 array."""
    max_hz, bins = fs / 2, fft_size // 2 + 1
    to_mel = lambda f: 1127.01048 * math.log(1 + f / 700.0)
    to_hz = lambda m: 700 * (math.exp(m / 1127.01048) - 1)
    c = np.arange(num_coeffs + 2).astype(np.float32) * (to_mel(max_hz) - to_mel(0)) / (num_coeffs + 1) + to_mel(0)
    for i in range(num_coeffs + 2):
        c[i] = to_hz(c[i])
        c[i] = math.floor(bins * c[i] / max_hz)
    mat = np.zeros((num_coeffs, bins))
    for i in range(1, num_coeffs + 1):
        lo, mid, hi = int(c[i - 1]), int(c[i]), int(c[i + 1])
        mat[i - 1, lo:mid] = (np.arange(lo, mid, dtype=np.float64) - lo) / max(mid - lo, 1)
        mat[i - 1, mid:hi] = 1 - (np.arange(mid, hi, dtype=np.float64) - mid) / max(hi - mid, 1)
    return mat


def lms_loss(clean_mags: torch.Tensor, est_mags: torch.Tensor, scales=(16, 32, 64)) -> torch.Tensor:
    """get_array_lms_loss(clean_array, est_array) (tools_for_loss.py:241-249) with perceptual_distance (:219-237),
    perceptual_transform (:195-216; note x.view(-1, 257) on a [257, T] array) and rmse (:123-131)."""
    banks = [torch.from_numpy(mel_filterbank(m).T.astype(np.float32)) for m in scales]
    total = 0
    for i in range(clean_mags.shape[0]):
        dists = []
        for fb in banks:
            lt = torch.log(torch.mm(clean_mags[i].reshape(-1, FFT_LEN // 2 + 1) / FFT_LEN, fb) + 1e-7)
            lp = torch.log(torch.mm(est_mags[i].reshape(-1, FFT_LEN // 2 + 1) / FFT_LEN, fb) + 1e-7)
            dists.append(torch.sqrt(torch.mean((lp - lt) ** 2, dim=-1) + 1e-7).mean())
        total = total + torch.stack(dists).mean()
    return total / clean_mags.shape[0]


def dccrn_lms_loss(sd, out_real, out_imag, target):
    """DCCRN.loss(..., perceptual=True), cfg.perceptual == 'LMS' (models.py:305-312)."""
    spec = conv_stft(target, sd["stft.weight"][:, 0, :].to(target.dtype))
    f = FFT_LEN // 2 + 1
    clean_mags = torch.sqrt(spec[:, :f] ** 2 + spec[:, f:] ** 2 + 1e-7)
    est_mags = torch.sqrt(out_real ** 2 + out_imag ** 2 + 1e-7)
    return lms_loss(clean_mags, est_mags)


# --------------------------------------------------------------------------------------
# train step (trainer.py:27-37 + train_interface.py:59)
# --------------------------------------------------------------------------------------
class OracleTrainer:
    """Holds leaf parameters + torch.optim.Adam(lr=1e-3), runs reference-equivalent steps on CPU."""

    def __init__(self, sd: Dict[str, torch.Tensor], masking_mode="C", loss="SI-SNR", lr=1e-3,
                 dtype=torch.float32):
        self.sd = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        self.keys = trainable_keys(self.sd)
        for k in self.keys:
            self.sd[k].requires_grad_(True)
        self.opt = torch.optim.Adam([self.sd[k] for k in self.keys], lr=lr)
        self.masking_mode, self.loss_name = masking_mode, loss

    def forward_backward(self, noisy, clean, taps=None):
        for k in self.keys:
            self.sd[k].grad = None
        t = {} if taps is None else taps
        _, _, wav = dccrn_forward(self.sd, noisy, self.masking_mode, train=True, taps=t)
        loss = dccrn_loss(wav, clean, self.loss_name)
        loss.backward()
        with torch.no_grad():                          # BN running stats, momentum 0.1 (models.py:76)
            for pbn, st in t["bn_stats"].items():
                if len(st) == 5:                       # ComplexBatchNorm: lerp_ of the five buffers (tools_for_model.py:527-550)
                    for k, val in zip(("RMr", "RMi", "RVrr", "RVri", "RVii"), st):
                        self.sd[pbn + k].lerp_(val.to(self.sd[pbn + k].dtype), BN_MOMENTUM)
                    self.sd[pbn + "num_batches_tracked"] += 1
                    continue
                m, v = st
                self.sd[pbn + "running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * m)
                self.sd[pbn + "running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * v)
                self.sd[pbn + "num_batches_tracked"] += 1
        return loss.detach(), wav.detach()

    def step(self, noisy, clean):
        loss, wav = self.forward_backward(noisy, clean)
        self.opt.step()
        return loss, wav

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: self.sd[k].grad for k in self.keys}


def synthetic_batch(B: int, L: int = 48000, seed: int = 1234, amp: float = 0.1):
    """SURVEY §8(d) synthetic inputs: U(-amp, amp), generator seed 1234, noisy drawn first."""
    g = torch.Generator().manual_seed(seed)
    noisy = (torch.rand(B, L, generator=g) * 2 - 1) * amp
    clean = (torch.rand(B, L, generator=g) * 2 - 1) * amp
    return noisy, clean
