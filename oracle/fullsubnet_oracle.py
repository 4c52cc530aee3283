"""TEST INFRASTRUCTURE ONLY — CPU oracle for the FullSubNet train step (BASELINE.json configs[2]).

Functional (state-dict in, tensors out) torch-CPU restatement of
  FullSubNet.forward / loss            models.py:626-682
  SequenceModel                        tools_for_model.py:726-795   (nn.LSTM x2 + Linear + activation)
  BaseModel.unfold                     tools_for_model.py:806-837
  BaseModel.offline_laplace_norm       tools_for_model.py:997-1011
  stft / mag_phase / build_complex_ideal_ratio_mask / compress_cIRM / decompress_cIRM   tools_for_model.py:628-717
  trainer.fullsubnet_train loop body   trainer.py:85-118
Configuration (config.py:71-80): sb_num_neighbors 15, fb_num_neighbors 0, num_freqs 257, look_ahead 2, LSTM, fullband
hidden 512 + ReLU, sub-band hidden 384 + no activation, offline_laplace_norm; STFT n_fft 512, hop int(400 * 0.75) = 300,
win 400 (torch.stft: centered, reflect padding, periodic Hann zero-padded to n_fft).
The reference's nn.LSTM carries dropout = 0.8 between the two layers (tools_for_model.py:746): a train-mode step is
stochastic, so parity is defined with dropout inactive - `dropout_mask` lets a caller inject a
fixed inter-layer mask to restate the train-mode arithmetic (inverted dropout, scale 1 / (1 - p)).

Parity status: PINNED.  tests/golden/make_golden.py (mode `fullsubnet`) imports the unmodified reference FullSubNet in the
build container and stores inputs / cIRM / cRM / loss / gradients under tests/golden/fullsubnet_golden.npz;
tests/test_fullsubnet_oracle.py checks this module against them; tests/test_fullsubnet_gpu.py checks the CUDA path
(csrc/fsnet.cu, lstm_seq.cu, lstm_step_tc.cu) against this module and the same fixtures.

Only tests/, __graft_entry__.smoke() and bench.py may import this module; the product never does.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

N_FFT, HOP, WIN, NUM_FREQS = 512, 300, 400, 257
SB_NEIGHBORS, FB_NEIGHBORS, LOOK_AHEAD = 15, 0, 2
FB_HIDDEN, SB_HIDDEN = 512, 384
EPSILON = float(np.finfo(np.float32).eps)               # tools_for_model.py:798


def init_state(seed: int = 0) -> Dict[str, torch.Tensor]:
    """State dict with the reference's keys / shapes / values for a torch seed: fb_model then sb_model, each nn.LSTM
    (2 layers) then nn.Linear, in the order of FullSubNet.__init__ (models.py:600-618); weight_init is False."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, (i, o, h) in (("fb_model", (NUM_FREQS, NUM_FREQS, FB_HIDDEN)),
                            ("sb_model", ((SB_NEIGHBORS * 2 + 1) + (FB_NEIGHBORS * 2 + 1), 2, SB_HIDDEN))):
        lstm = nn.LSTM(input_size=i, hidden_size=h, num_layers=2, batch_first=True, dropout=0.8)
        for k, p in lstm.named_parameters():
            sd[f"{name}.sequence_model.{k}"] = p.data.clone()
        fc = nn.Linear(h, o)
        sd[f"{name}.fc_output_layer.weight"] = fc.weight.data.clone()
        sd[f"{name}.fc_output_layer.bias"] = fc.bias.data.clone()
    return sd


def stft(y: torch.Tensor) -> torch.Tensor:
    """tools.stft (tools_for_model.py:628-648) = torch.stft(y, 512, 300, 400, hann_window(400), center=True, reflect),
    restated as explicit framing + rFFT: [B, L] -> complex [B, 257, 1 + L // 300]."""
    w = torch.hann_window(WIN, dtype=y.dtype)
    lpad = (N_FFT - WIN) // 2
    w = F.pad(w, [lpad, N_FFT - WIN - lpad])                                   # window centred in the n_fft frame
    yp = F.pad(y[:, None, :], [N_FFT // 2, N_FFT // 2], mode="reflect")[:, 0]
    frames = yp.unfold(-1, N_FFT, HOP)                                         # [B, T, 512]
    return torch.fft.rfft(frames * w, dim=-1).transpose(1, 2)                  # [B, 257, T]


def istft(spec: torch.Tensor, length: Optional[int] = None) -> torch.Tensor:
    """tools.istft (tools_for_model.py:651-679) = torch.istft(spec, 512, 300, 400, hann_window(400), length=length), restated:
    irfft per frame, centred window, overlap-add at hop 300, division by the overlap-added squared window, n_fft / 2 trimmed
    at the front: complex [B, 257, T] -> [B, length or 300 (T - 1)]."""
    B, _, T = spec.shape
    w = torch.hann_window(WIN, dtype=torch.float32)
    lpad = (N_FFT - WIN) // 2
    w = F.pad(w, [lpad, N_FFT - WIN - lpad])
    frames = torch.fft.irfft(spec.transpose(1, 2), n=N_FFT, dim=-1) * w               # [B, T, 512]
    full = N_FFT + HOP * (T - 1)
    y = frames.new_zeros(B, full)
    env = frames.new_zeros(full)
    for t in range(T):
        y[:, t * HOP:t * HOP + N_FFT] += frames[:, t]
        env[t * HOP:t * HOP + N_FFT] += w * w
    n = HOP * (T - 1) if length is None else length
    y, env = y[:, N_FFT // 2:N_FFT // 2 + n], env[N_FFT // 2:N_FFT // 2 + n]
    out = torch.where(env > 1e-11, y / env.clamp_min(1e-11), torch.zeros_like(y))
    return F.pad(out, [0, n - out.shape[1]]) if out.shape[1] < n else out


def mag_phase(c: torch.Tensor):
    return torch.abs(c), torch.angle(c)                                        # tools_for_model.py:682-683


def compress_cirm(mask: torch.Tensor, K: float = 10.0, C: float = 0.1) -> torch.Tensor:
    mask = -100.0 * (mask <= -100).to(mask.dtype) + mask * (mask > -100).to(mask.dtype)    # tools_for_model.py:709-713
    return K * (1 - torch.exp(-C * mask)) / (1 + torch.exp(-C * mask))


def decompress_cirm(mask: torch.Tensor, K: float = 10.0, limit: float = 9.9) -> torch.Tensor:
    dt = mask.dtype
    mask = limit * (mask >= limit).to(dt) - limit * (mask <= -limit).to(dt) + mask * (torch.abs(mask) < limit).to(dt)
    return -K * torch.log((K - mask) / (K + mask))                             # tools_for_model.py:720-723


def build_complex_ideal_ratio_mask(noisy: torch.Tensor, clean: torch.Tensor) -> torch.Tensor:
    den = noisy.real ** 2 + noisy.imag ** 2 + EPSILON                          # tools_for_model.py:697-705
    mr = (noisy.real * clean.real + noisy.imag * clean.imag) / den
    mi = (noisy.real * clean.imag - noisy.imag * clean.real) / den
    return compress_cirm(torch.stack((mr, mi), dim=-1))


def offline_laplace_norm(x: torch.Tensor) -> torch.Tensor:
    return x / (x.mean(dim=(1, 2, 3), keepdim=True) + 1e-5)                   # tools_for_model.py:1006-1010


def unfold(x: torch.Tensor, num_neighbor: int) -> torch.Tensor:
    """BaseModel.unfold: [B, C, F, T] -> [B, F, C, 2n+1, T]; reflect padding along frequency, sub-band f holds the
    rows f-n .. f+n of the padded spectrogram."""
    B, C, Fq, T = x.shape
    if num_neighbor < 1:
        return x.permute(0, 2, 1, 3).reshape(B, Fq, C, 1, T)
    xp = F.pad(x.reshape(B * C, 1, Fq, T), [0, 0, num_neighbor, num_neighbor], mode="reflect")[:, 0]   # [BC, F+2n, T]
    idx = torch.arange(Fq)[:, None] + torch.arange(2 * num_neighbor + 1)[None, :]                       # [F, 2n+1]
    out = xp[:, idx]                                                                                    # [BC, F, 2n+1, T]
    return out.reshape(B, C, Fq, 2 * num_neighbor + 1, T).permute(0, 2, 1, 3, 4).contiguous()


def lstm_layer(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh) -> torch.Tensor:
    """One nn.LSTM layer, batch_first, zero initial state, gate order i, f, g, o: x [B, T, I] -> [B, T, H]."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    pre = x @ w_ih.t() + (b_ih + b_hh)
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out: List[torch.Tensor] = []
    for t in range(T):
        g = pre[:, t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out, 1)


def sequence_model(sd, prefix: str, x: torch.Tensor, activation: Optional[str], dropout_mask=None) -> torch.Tensor:
    """SequenceModel.forward: [B, F, T] -> [B, O, T]."""
    p = prefix + ".sequence_model."
    o = x.permute(0, 2, 1)
    o = lstm_layer(o, sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"], sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"])
    if dropout_mask is not None:
        o = o * dropout_mask                              # inverted dropout between the layers (train mode)
    o = lstm_layer(o, sd[p + "weight_ih_l1"], sd[p + "weight_hh_l1"], sd[p + "bias_ih_l1"], sd[p + "bias_hh_l1"])
    o = F.linear(o, sd[prefix + ".fc_output_layer.weight"], sd[prefix + ".fc_output_layer.bias"])
    if activation == "ReLU":
        o = torch.relu(o)
    elif activation is not None:
        raise NotImplementedError(activation)
    return o.permute(0, 2, 1)


def fullsubnet_forward(sd, noisy_mag: torch.Tensor, taps: Optional[dict] = None, dropout_masks=None) -> torch.Tensor:
    """FullSubNet.forward (models.py:626-672): noisy_mag [B, 257, T] -> cRM [B, 257, T, 2].  Dropout inactive unless
    `dropout_masks` = (fb [B, T+2, 512], sb [B*257, T+2, 384]) multipliers (0 or 1 / (1 - p)) are injected."""
    mfb, msb = dropout_masks if dropout_masks is not None else (None, None)
    x = noisy_mag[:, None] if noisy_mag.dim() == 3 else noisy_mag
    x = F.pad(x, [0, LOOK_AHEAD])
    B, C, Fq, T = x.shape
    fb_in = offline_laplace_norm(x).reshape(B, C * Fq, T)
    fb_out = sequence_model(sd, "fb_model", fb_in, "ReLU", mfb).reshape(B, 1, Fq, T)
    fb_unf = unfold(fb_out, FB_NEIGHBORS).reshape(B, Fq, FB_NEIGHBORS * 2 + 1, T)
    nm_unf = unfold(x, SB_NEIGHBORS).reshape(B, Fq, SB_NEIGHBORS * 2 + 1, T)
    sb_in = offline_laplace_norm(torch.cat([nm_unf, fb_unf], dim=2))
    if taps is not None:
        taps["fb_out"], taps["sb_in"] = fb_out.detach(), sb_in.detach()
    sb_in = sb_in.reshape(B * Fq, (SB_NEIGHBORS * 2 + 1) + (FB_NEIGHBORS * 2 + 1), T)
    sb_mask = sequence_model(sd, "sb_model", sb_in, None, msb)
    sb_mask = sb_mask.reshape(B, Fq, 2, T).permute(0, 2, 1, 3)
    return sb_mask[:, :, :, LOOK_AHEAD:].permute(0, 2, 3, 1)


def train_step_loss(sd, noisy: torch.Tensor, clean: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """Loop body of trainer.fullsubnet_train (trainer.py:97-107) with cfg.loss = 'MSE': F.mse_loss(cIRM, cRM)."""
    nc, cc = stft(noisy), stft(clean)
    noisy_mag, _ = mag_phase(nc)
    cirm = build_complex_ideal_ratio_mask(nc, cc)
    crm = fullsubnet_forward(sd, noisy_mag, taps)
    if taps is not None:
        taps["noisy_mag"], taps["cIRM"], taps["cRM"] = noisy_mag.detach(), cirm.detach(), crm.detach()
    return F.mse_loss(cirm, crm, reduction="mean")
