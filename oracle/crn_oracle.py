"""TEST INFRASTRUCTURE ONLY — CPU oracle for the CRN train-step path (BASELINE.json configs[0]).

Functional (state-dict in, tensors out) torch-CPU restatement of class CRN (models.py:329-565) with
RealConv2d / RealConvTranspose2d (tools_for_model.py:341-425) and ConvSTFT 'real' (tools_for_model.py:54-68):
  wave -> STFT -> |X| (DC dropped) -> 6 x [Conv2d + BN + PReLU] -> nn.LSTM(512 -> 128) -> Linear(128 -> 512)
       -> 6 x [ConvTranspose2d on cat(out, skip) (+ BN + PReLU)] -> tanh(out) * |X| with the noisy phase -> ISTFT -> clamp.
Every function cites the reference file:line (relative to /root/reference) it restates.

Parity status: PINNED.  tests/golden/make_golden.py imports the unmodified reference CRN in the build container
and stores outputs / loss / gradients under tests/golden/crn_golden.npz; tests/test_oracle_golden.py checks this
module against them.

Only tests/, __graft_entry__.smoke() and bench.py may import this module; the product never does.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import dccrn_oracle as D

KERNEL_NUM = [1, 16, 32, 64, 128, 128, 128]        # [2] + cfg.dccrn_kernel_num, halved (models.py:362-363, 378-380)
RNN_INPUT, RNN_HIDDEN = 512, 128                    # config.py:68; rnn_units // 2 (models.py:359)


def init_state(seed: int = 0) -> Dict[str, torch.Tensor]:
    """State dict with the reference's keys / shapes / values for a torch seed: modules are instantiated in the
    order of CRN.__init__ (encoder models.py:376-389, LSTM :391-397, Linear :398, decoder :400-430) so the RNG
    stream is consumed identically; keys are emitted in the reference's registration order (encoder, decoder,
    enhance, tranform)."""
    import torch.nn as nn
    kn = KERNEL_NUM
    torch.manual_seed(seed)
    enc: Dict[str, torch.Tensor] = {}
    dec: Dict[str, torch.Tensor] = {}
    rnn: Dict[str, torch.Tensor] = {}

    def bn_prelu(sd, prefix_bn, prefix_act, c):
        sd[prefix_bn + "weight"] = torch.ones(c)
        sd[prefix_bn + "bias"] = torch.zeros(c)
        sd[prefix_bn + "running_mean"] = torch.zeros(c)
        sd[prefix_bn + "running_var"] = torch.ones(c)
        sd[prefix_bn + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        sd[prefix_act + "weight"] = torch.full((1,), 0.25)

    for i in range(6):
        m = nn.Conv2d(kn[i], kn[i + 1], (5, 2), (2, 1), padding=[2, 0])          # tools_for_model.py:373-377
        nn.init.normal_(m.weight.data, std=0.05)
        enc[f"encoder.{i}.0.conv.weight"] = m.weight.data.clone()
        enc[f"encoder.{i}.0.conv.bias"] = torch.zeros_like(m.bias.data)
        bn_prelu(enc, f"encoder.{i}.1.", f"encoder.{i}.2.", kn[i + 1])
    m = nn.LSTM(RNN_INPUT, RNN_HIDDEN)                                            # models.py:391-397
    for name, p in m.named_parameters():
        rnn["enhance." + name] = p.data.clone()
    m = nn.Linear(RNN_HIDDEN, RNN_INPUT)                                          # models.py:398
    rnn["tranform.weight"] = m.weight.data.clone()
    rnn["tranform.bias"] = m.bias.data.clone()
    for j, idx in enumerate(range(6, 0, -1)):
        m = nn.ConvTranspose2d(2 * kn[idx], kn[idx - 1], (5, 2), (2, 1), padding=(2, 0), output_padding=(1, 0))
        nn.init.normal_(m.weight.data, std=0.05)                                  # tools_for_model.py:414-418
        dec[f"decoder.{j}.0.conv.weight"] = m.weight.data.clone()
        dec[f"decoder.{j}.0.conv.bias"] = torch.zeros_like(m.bias.data)
        if idx != 1:
            bn_prelu(dec, f"decoder.{j}.1.", f"decoder.{j}.2.", kn[idx - 1])
    sd: Dict[str, torch.Tensor] = {}
    k_a, k_s, w = D.stft_bases()
    sd["stft.weight"] = torch.from_numpy(k_a.astype(np.float32))[:, None, :]
    sd["istft.weight"] = torch.from_numpy(k_s.astype(np.float32))[:, None, :]
    sd["istft.window"] = torch.from_numpy(w.astype(np.float32))[None, :, None]
    sd["istft.enframe"] = torch.eye(D.WIN_LEN)[:, None, :]
    sd.update(enc)
    sd.update(dec)
    sd.update(rnn)
    return sd


def trainable_keys(sd: Dict[str, torch.Tensor]) -> List[str]:
    return D.trainable_keys(sd)


def real_conv2d(x, w, b):
    """RealConv2d.forward (tools_for_model.py:379-387): causal left pad of 1 frame, Conv2d pad (2, 0)."""
    return F.conv2d(F.pad(x, [1, 0, 0, 0]), w, b, stride=(2, 1), padding=(2, 0))


def real_conv_transpose2d(x, w, b):
    """RealConvTranspose2d.forward (tools_for_model.py:420-424)."""
    return F.conv_transpose2d(x, w, b, stride=(2, 1), padding=(2, 0), output_padding=(1, 0))


def crn_forward(sd, wav, targets=None, train=True, taps=None):
    """CRN.forward (models.py:460-532), T-F masking branch.  Returns (est_mags, target_mags, out_wav)."""
    t = {} if taps is None else taps
    t.setdefault("bn_stats", {})
    k_a = sd["stft.weight"][:, 0, :].to(wav.dtype)
    spec = D.conv_stft(wav, k_a)                                     # [B, 514, T]
    real, imag = spec[:, :257], spec[:, 257:]
    mags = torch.sqrt(real ** 2 + imag ** 2)                         # tools_for_model.py:65
    phase = torch.atan2(imag, real)                                  # tools_for_model.py:66
    t["spec"] = spec.detach()
    out = mags[:, None, 1:]                                          # models.py:464-466

    def norm_act(x, pbn, pact):
        if train:
            y, m, v = D.batch_norm_train(x, sd[pbn + "weight"], sd[pbn + "bias"])
            n = x.numel() / x.shape[1]
            t["bn_stats"][pbn] = (m.detach(), (v * n / max(n - 1, 1)).detach())
        else:
            y = D.batch_norm_eval(x, sd[pbn + "weight"], sd[pbn + "bias"], sd[pbn + "running_mean"], sd[pbn + "running_var"])
        return D.prelu(y, sd[pact + "weight"])

    enc_out = []
    for i in range(6):                                               # models.py:470-473
        out = real_conv2d(out, sd[f"encoder.{i}.0.conv.weight"], sd[f"encoder.{i}.0.conv.bias"])
        t[f"enc{i}_conv"] = out.detach()
        out = norm_act(out, f"encoder.{i}.1.", f"encoder.{i}.2.")
        t[f"enc{i}"] = out.detach()
        enc_out.append(out)
    B, C, Dm, T = out.shape
    rnn_in = out.permute(3, 0, 1, 2).reshape(T, B, C * Dm)            # models.py:475-478
    lstm = D.lstm_fused if D.USE_FUSED_LSTM else D.lstm_seq
    h = lstm(rnn_in, sd["enhance.weight_ih_l0"], sd["enhance.weight_hh_l0"], sd["enhance.bias_ih_l0"], sd["enhance.bias_hh_l0"])
    t["lstm"] = h.detach()
    out = F.linear(h, sd["tranform.weight"], sd["tranform.bias"])   # models.py:480
    t["proj"] = out.detach()
    out = out.reshape(T, B, C, Dm).permute(1, 2, 3, 0)                # models.py:481-483
    for j in range(6):                                               # models.py:485-489
        out = torch.cat([out, enc_out[-1 - j]], 1)
        out = real_conv_transpose2d(out, sd[f"decoder.{j}.0.conv.weight"], sd[f"decoder.{j}.0.conv.bias"])
        t[f"dec{j}_conv"] = out.detach()
        if j != 5:
            out = norm_act(out, f"decoder.{j}.1.", f"decoder.{j}.2.")
        out = out[..., 1:]
        if j != 5:
            t[f"dec{j}"] = out.detach()
    out = F.pad(out.squeeze(1), [0, 0, 1, 0])                        # models.py:500-502
    target_mags = None
    if targets is not None:                                          # models.py:505
        ts = D.conv_stft(targets, k_a)
        target_mags = torch.sqrt(ts[:, :257] ** 2 + ts[:, 257:] ** 2)
    est_mags = torch.tanh(out) * mags                                # models.py:521-522
    out_spec = torch.cat([est_mags * torch.cos(phase), est_mags * torch.sin(phase)], 1)
    k_s = sd["istft.weight"][:, 0, :].to(wav.dtype)
    out_wav = D.conv_istft(out_spec, k_s, sd["istft.window"].to(wav.dtype)).squeeze(1)
    out_wav = torch.clamp(out_wav, -1, 1)                            # models.py:530
    return est_mags, target_mags, out_wav


def crn_loss(estimated, target, loss="MSE"):
    """CRN.loss, non-perceptual branch (models.py:558-565)."""
    return D.dccrn_loss(estimated, target, loss)


class OracleTrainer:
    """Leaf parameters + torch.optim.Adam(lr=1e-3): reference-equivalent CRN steps on CPU (trainer.py:27-37)."""

    def __init__(self, sd, loss="MSE", lr=1e-3, dtype=torch.float32):
        self.sd = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        self.keys = trainable_keys(self.sd)
        for k in self.keys:
            self.sd[k].requires_grad_(True)
        self.opt = torch.optim.Adam([self.sd[k] for k in self.keys], lr=lr)
        self.loss_name = loss

    def forward_backward(self, noisy, clean, taps=None):
        for k in self.keys:
            self.sd[k].grad = None
        t = {} if taps is None else taps
        est, tgt, wav = crn_forward(self.sd, noisy, clean, train=True, taps=t)
        loss = crn_loss(wav, clean, self.loss_name)
        loss.backward()
        with torch.no_grad():
            for pbn, (m, v) in t["bn_stats"].items():
                self.sd[pbn + "running_mean"].mul_(1 - D.BN_MOMENTUM).add_(D.BN_MOMENTUM * m)
                self.sd[pbn + "running_var"].mul_(1 - D.BN_MOMENTUM).add_(D.BN_MOMENTUM * v)
                self.sd[pbn + "num_batches_tracked"] += 1
        t["est_mags"], t["target_mags"] = est.detach(), tgt.detach()
        return loss.detach(), wav.detach()

    def step(self, noisy, clean):
        loss, wav = self.forward_backward(noisy, clean)
        self.opt.step()
        return loss, wav

    def grads(self):
        return {k: self.sd[k].grad for k in self.keys}
