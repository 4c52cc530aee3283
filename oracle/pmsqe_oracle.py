"""TEST INFRASTRUCTURE ONLY — CPU oracle for the PMSQE perceptual loss.

PARITY UNPINNED.  The reference computes this loss entirely inside third-party code that is absent from /root/reference
and from this image (no network): `asteroid.losses.{SingleSrcPMSQE, PITLossWrapper}` and
`asteroid_filterbanks.{STFTFB, Encoder, transforms.mag}`; the reference pins no version (README.md:30 names only
"Pytorch 1.9.0") and holds no test or golden value for PMSQE.  This module therefore restates the PUBLISHED algorithm
(Martin-Donas, Gomez, Gonzalez, Peinado: "A deep learning loss function based on the perceptual evaluation of the speech
quality", IEEE SPL 25(11), 2018, as implemented in asteroid's losses/pmsqe.py) anchored on the reference's own call
sites:
    tools_for_loss.py:255   pmsqe_stft = Encoder(STFTFB(kernel_size=512, n_filters=512, stride=256))
    tools_for_loss.py:256   pmsqe_loss = PITLossWrapper(SingleSrcPMSQE(), pit_from='pw_pt')
    tools_for_loss.py:259-269 get_array_pmsqe_loss: wav.view(N, -1, 16000) -> mag(stft) -> pmsqe_loss(est_spec, clean_spec)
The Bark matrix of asteroid (bark_matrix_16k.mat) is not available; it is rebuilt here from the ITU-T P.862 tables
(band membership x pow_dens_correction_factor, the way P.862's freq_warping forms its pitch power densities).  Every such
table is an INPUT of the CUDA path (include/sefd.h: sefd_pmsqe_*), so a site with asteroid installed can pass asteroid's
own tensors and get its numbers.  What the tests pin is the CUDA path against THIS restatement (value and gradient).

Only tests/, __graft_entry__.smoke() and bench.py may import this module; the product never does.
"""
from __future__ import annotations

import itertools
import os
import sys
from typing import Dict

import numpy as np
import torch

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                    "dnn-based-speech-enhancement-in-the-frequency-domain_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from sefd import p862_16k as P862   # noqa: E402  (numeric tables only)

NFFT, HOP, NBIN, NBARK, FS = 512, 256, 257, 49, 16000
ALPHA, BETA = 0.1, 0.309 * 0.1                      # SingleSrcPMSQE.__init__: alpha, beta = 0.309 * alpha
EPS = 1e-8


def stft_filters() -> np.ndarray:
    """STFTFB(n_filters=512, kernel_size=512, stride=256) analysis filters [514, 512]: rows [Re; Im] of fft(eye(512))[:257],
    scaled by 1 / (0.5 sqrt(kernel * n_filters / stride)) = 1/16, DC and Nyquist real rows further / sqrt(2), window
    sqrt(hanning(513)[:-1])."""
    f = np.fft.fft(np.eye(NFFT))
    f = f / (0.5 * np.sqrt(NFFT * NFFT / HOP))
    filt = np.vstack([np.real(f[:NBIN]), np.imag(f[:NBIN])])
    filt[0] /= np.sqrt(2.0)
    filt[NFFT // 2] /= np.sqrt(2.0)
    win = np.hanning(NFFT + 1)[:-1] ** 0.5
    return (filt * win).astype(np.float32)


def tables() -> Dict[str, torch.Tensor]:
    """SingleSrcPMSQE.populate_constants / register_16k_constants restated from the P.862 tables."""
    thr = np.asarray(P862.ABS_THRESH_POWER, dtype=np.float64)
    centre = np.asarray(P862.CENTRE_OF_BAND_BARK, dtype=np.float64)
    zw = 0.23 * np.clip(6.0 / (centre + 2.0), 1.0, 2.0) ** 0.15          # P.862 modified Zwicker power
    width = np.asarray(P862.WIDTH_OF_BAND_BARK, dtype=np.float64)
    bark = np.zeros((NBIN, NBARK))
    f = 0
    for k, n in enumerate(P862.NR_OF_HZ_BANDS_PER_BARK_BAND):           # bins 0..255 in order; the Nyquist bin is unused
        bark[f:f + n, k] = P862.POW_DENS_CORRECTION_FACTOR[k]
        f += n
    mask = np.zeros(NBIN)
    mask[11] = 0.5 * 25.0 / 31.25
    mask[12:104] = 1.0
    mask[104] = 0.5
    mask *= 2.0 * (NFFT + 2.0) / NFFT ** 2                                # sqrt_hann power correction factor 2.0
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    return {"bark": t(bark), "thr": t(thr), "zw": t(zw), "width": t(width), "mask": t(mask)}


def encoder_mag(wav: torch.Tensor) -> torch.Tensor:
    """transforms.mag(Encoder(STFTFB)(wav)): wav [N, S, 16000] -> [N, S, 257, 61] (power + 1e-8 under the root)."""
    N, S, L = wav.shape
    filt = torch.from_numpy(stft_filters()).to(wav.dtype)[:, None, :]
    spec = torch.nn.functional.conv1d(wav.reshape(N * S, 1, L), filt, stride=HOP)       # [N*S, 514, 61]
    re, im = spec[:, :NBIN], spec[:, NBIN:]
    return torch.sqrt(re ** 2 + im ** 2 + EPS).reshape(N, S, NBIN, -1)


def single_src_pmsqe(est: torch.Tensor, tgt: torch.Tensor, tb=None) -> torch.Tensor:
    """SingleSrcPMSQE.forward with bark_eq = gain_eq = True, no padding mask: est / tgt [B, 257, T] -> [B]."""
    tb = tb or tables()
    tb = {k: v.to(est.dtype) for k, v in tb.items()}
    thr, zw, width = tb["thr"], tb["zw"], tb["width"]
    Sp, Sl = P862.SP, P862.SL
    est, tgt = est.transpose(1, 2), tgt.transpose(1, 2)                   # [B, T, F]

    def at_sll(x):                                                        # magnitude_at_sll
        m = (x * tb["mask"]).mean(-1, keepdim=True).sum(-2, keepdim=True) / x.shape[1]
        return x * 1e7 / m

    def audible(b, factor):                                               # compute_audible_power
        return torch.where(b > thr * factor, b, torch.zeros_like(b)).sum(-1, keepdim=True)

    ref = Sp * at_sll(tgt) @ tb["bark"]                                   # bark_computation
    deg = Sp * at_sll(est) @ tb["bark"]
    # bark_freq_equalization
    not_silent = audible(ref, 100.0) >= 1e7
    cond = ref >= thr * 100.0
    zero = torch.zeros_like(ref)
    ref_t, deg_t = torch.where(cond, ref, zero), torch.where(cond, deg, zero)
    ppb_ref = torch.where(not_silent, ref_t, zero).sum(-2, keepdim=True)
    ppb_deg = torch.where(not_silent, deg_t, zero).sum(-2, keepdim=True)
    deg = torch.clamp((ppb_ref + 1000.0) / (ppb_deg + 1000.0), 0.01, 100.0) * deg
    # bark_gain_equalization
    gain = (audible(ref, 1.0) + 5e3) / (audible(deg, 1.0) + 5e3)
    deg = torch.clamp(gain, 3e-4, 5.0) * deg

    def loudness(b):                                                      # loudness_computation (Zwicker)
        ld = Sl * torch.pow(thr / 0.5, zw) * (torch.pow(0.5 + 0.5 * b / thr, zw) - 1.0)
        return torch.where(b < thr, torch.zeros_like(ld), ld)

    lr, ld = loudness(ref), loudness(deg)                                 # compute_distortion_tensors
    r = torch.abs(ld - lr)
    m = 0.25 * torch.minimum(lr, ld)
    sym = torch.clamp(r - m, min=EPS)
    asym = torch.pow((deg + 50.0) / (ref + 50.0), 1.2)
    af = torch.where(asym < 3.0, torch.zeros_like(asym), torch.clamp(asym, max=12.0))
    asd = af * sym
    # per_frame_distortion
    d_frame = torch.sqrt(((sym * width) ** 2 + EPS).sum(-1, keepdim=True)) * torch.sqrt(width.sum())
    da_frame = (asd * width).sum(-1, keepdim=True)
    w = torch.pow((audible(ref, 1.0) + 1e5) / 1e7, 0.04)
    wd = torch.clamp(d_frame / w, max=45.0)
    wda = torch.clamp(da_frame / w, max=45.0)
    return (ALPHA * wd + BETA * wda).mean(dim=(-1, -2))


def pit_pw_pt(est_spec: torch.Tensor, clean_spec: torch.Tensor, tb=None):
    """PITLossWrapper(loss, pit_from='pw_pt'): pw[b, i, j] = loss(est[:, i], clean[:, j]); the permutation with the lowest
    mean over sources per batch item; mean over the batch.  Returns (loss, pw, best permutation per item)."""
    N, S = est_spec.shape[:2]
    pw = torch.stack([torch.stack([single_src_pmsqe(est_spec[:, i], clean_spec[:, j], tb) for j in range(S)], 1)
                      for i in range(S)], 1)                              # [N, S(est), S(target)]
    perms = list(itertools.permutations(range(S)))
    loss_set = torch.stack([sum(pw[:, i, p[i]] for i in range(S)) / S for p in perms], 1)
    best, idx = loss_set.min(1)
    return best.mean(), pw, [perms[int(i)] for i in idx]


def get_array_pmsqe_loss(clean: torch.Tensor, est: torch.Tensor, tb=None) -> torch.Tensor:
    """tools_for_loss.py:259-269."""
    N = clean.shape[0]
    c = encoder_mag(clean.reshape(N, -1, FS))
    e = encoder_mag(est.reshape(N, -1, FS))
    return pit_pw_pt(e, c, tb)[0]
