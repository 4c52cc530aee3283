"""Thin torch-tensor wrappers over the op-level C ABI (include/sefd.h).

PyTorch is used for device memory, the current stream and autograd plumbing only; every FLOP below
runs in libsefd.so.  Tensors must be CUDA float32; there is no CPU path.
"""
import torch

from . import _lib

MODES = {"E": 1, "C": 2, "R": 3, "Direct(None make)": 5}
PLAN_NO_SKIP = 1          # include/sefd.h: SEFD_PLAN_NO_SKIP
PLAN_REAL_LSTM = 2        # include/sefd.h: SEFD_PLAN_REAL_LSTM (cfg.lstm = 'real')
PLAN_CBN = 4              # include/sefd.h: SEFD_PLAN_CBN (DCCRN(use_cbn=True): ComplexBatchNorm)
LOSSES = {"MSE": 0, "SDR": 1, "SI-SNR": 2, "SI-SDR": 3}


def _req(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("sefd: tensors must live on a CUDA device (no CPU fallback on this path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"sefd: expected float32 tensor, got {t.dtype} (the kernels read float*; cast with .float())")
        if not t.is_contiguous():
            raise RuntimeError("sefd: tensors must be contiguous")


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream(t=None):
    """Current stream of the tensor's device (of the current device when no tensor is given)."""
    return torch.cuda.current_stream(t.device if t is not None else None).cuda_stream


def frames(L):
    if L % 100:
        raise ValueError(f"waveform length {L} must be a multiple of the hop (100)")
    return L // 100 + 3


# ---- STFT / ISTFT -------------------------------------------------------------------------------
def stft(wav):
    """ConvSTFT 'complex' (tools_for_model.py:54-61): wav [B,L] -> spec [B,257,T,2]."""
    wav = wav.contiguous()
    _req(wav)
    B, L = wav.shape
    spec = torch.empty(B, 257, frames(L), 2, device=wav.device, dtype=torch.float32)
    _lib.check(_lib.load().sefd_stft_forward(ptr(wav), ptr(spec), B, L, stream()), "stft_forward")
    return spec


def istft(spec, L):
    spec = spec.contiguous()
    _req(spec)
    B = spec.shape[0]
    assert spec.shape[1:] == (257, frames(L), 2)
    wav = torch.empty(B, L, device=spec.device, dtype=torch.float32)
    _lib.check(_lib.load().sefd_istft_forward(ptr(spec), ptr(wav), B, L, stream()), "istft_forward")
    return wav


def istft_backward(dwav):
    dwav = dwav.contiguous()
    _req(dwav)
    B, L = dwav.shape
    dspec = torch.empty(B, 257, frames(L), 2, device=dwav.device, dtype=torch.float32)
    _lib.check(_lib.load().sefd_istft_backward(ptr(dwav), ptr(dspec), B, L, stream()), "istft_backward")
    return dspec


def mask_istft(spec, mask, mode, L, want_spec=True):
    """models.py:253-282. spec [B,257,T,2], mask [B,256,T,2] -> (out_real, out_imag, out_wav, raw_wav)."""
    _req(spec, mask)
    B, T = spec.shape[0], frames(L)
    dev = spec.device
    o_r = torch.empty(B, 257, T, device=dev) if want_spec else None
    o_i = torch.empty(B, 257, T, device=dev) if want_spec else None
    wav = torch.empty(B, L, device=dev)
    raw = torch.empty(B, L, device=dev)
    _lib.check(_lib.load().sefd_mask_istft_forward(ptr(spec), ptr(mask), MODES[mode], B, L, ptr(o_r), ptr(o_i),
                                                   ptr(wav), ptr(raw), stream()), "mask_istft_forward")
    return o_r, o_i, wav, raw


# ---- either transform geometry of config.py:55-61 (nfft 512: 400/100, 257 bins; nfft 1024: 800/200, 513 bins) ----
def _geometry(nfft, L):
    if nfft not in (512, 1024):
        raise ValueError(f"fft length {nfft} is not built (512 or 1024)")
    hop = 100 if nfft == 512 else 200
    return nfft // 2 + 1, L // hop + 3


def stft_n(wav, nfft):
    """ConvSTFT 'complex' for win = 25/32 nfft, hop = win/4: wav [B,L] -> spec [B,nfft/2+1,T,2]."""
    wav = wav.contiguous()
    _req(wav)
    B, L = wav.shape
    F, T = _geometry(nfft, L)
    spec = torch.empty(B, F, T, 2, device=wav.device, dtype=torch.float32)
    _lib.check(_lib.load().sefd_stft_forward_n(ptr(wav), ptr(spec), B, L, nfft, stream()), "stft_forward_n")
    return spec


def mask_istft_n(spec, mask, mode, L, nfft):
    """mask apply (mode None: plain ISTFT) + ConviSTFT + clamp for either geometry: -> wav [B,L]."""
    spec = spec.contiguous()
    _req(spec)
    B = spec.shape[0]
    F, T = _geometry(nfft, L)
    assert spec.shape[1:] == (F, T, 2), spec.shape
    if mask is not None:
        mask = mask.contiguous()
        _req(mask)
        assert mask.shape == (B, F - 1, T, 2), mask.shape
    wav = torch.empty(B, L, device=spec.device, dtype=torch.float32)
    _lib.check(_lib.load().sefd_mask_istft_forward_n(ptr(spec), ptr(mask), MODES[mode] if mask is not None else 0, B, L,
                                                     nfft, ptr(wav), stream()), "mask_istft_forward_n")
    return wav


def stft_mask_istft(wav, mask, mode, nfft):
    """wave -> STFT -> mask -> ISTFT -> clamp in one kernel (the spectrum never reaches HBM)."""
    wav, mask = wav.contiguous(), mask.contiguous()
    _req(wav, mask)
    B, L = wav.shape
    F, T = _geometry(nfft, L)
    assert mask.shape == (B, F - 1, T, 2), mask.shape
    out = torch.empty_like(wav)
    _lib.check(_lib.load().sefd_stft_mask_istft_fused(ptr(wav), ptr(mask), MODES[mode], B, L, nfft, ptr(out), stream()),
               "stft_mask_istft_fused")
    return out


def mask_istft_backward(dwav, raw, spec, mask, mode):
    _req(dwav, raw, spec, mask)
    B, L = dwav.shape
    dmask = torch.empty_like(mask)
    _lib.check(_lib.load().sefd_mask_istft_backward(ptr(dwav), ptr(raw), ptr(spec), ptr(mask), MODES[mode], B, L,
                                                    ptr(dmask), stream()), "mask_istft_backward")
    return dmask


# ---- losses -------------------------------------------------------------------------------------
class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, est, tgt, kind):
        # rows = everything but the last axis (tools_for_loss.py:17-44 reduce over the last axis with keepdim and then
        # average over the rest; for [B, L] waveforms rows = B, for the [B, 257, T] spectra of the Direct mode rows = B*257)
        ctx.shape = est.shape
        est, tgt = est.contiguous().reshape(-1, est.shape[-1]), tgt.contiguous().reshape(-1, tgt.shape[-1])
        _req(est, tgt)
        B, L = est.shape
        scratch = torch.empty(8 * B, device=est.device, dtype=torch.float64)
        loss = torch.empty(1, device=est.device)
        coef = torch.empty(2 * B, device=est.device)
        _lib.check(_lib.load().sefd_loss_forward(ptr(est), ptr(tgt), B, L, kind, ptr(scratch), ptr(loss), ptr(coef),
                                                 stream()), "loss_forward")
        ctx.save_for_backward(est, tgt, coef)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        est, tgt, coef = ctx.saved_tensors
        B, L = est.shape
        gout = gout.contiguous().float()
        d = torch.empty_like(est)
        _lib.check(_lib.load().sefd_loss_backward(ptr(est), ptr(tgt), ptr(coef), ptr(gout), ptr(d), B, L, stream()),
                   "loss_backward")
        return d.reshape(ctx.shape), None, None


def loss(est, tgt, name):
    """DCCRN.loss non-perceptual branch (models.py:315-323): returns the value that is minimised."""
    return _Loss.apply(est, tgt, LOSSES[name])


# ---- complex conv / convT (channels-last) ---------------------------------------------------------
def _cconv_ws(Cin, Cout, dev):
    n = _lib.load().sefd_cconv_workspace_bytes(Cin, Cout)
    return torch.empty((n + 255) // 256 * 256, device=dev, dtype=torch.uint8)


def cconv2d_forward(x, wr, br, wi, bi):
    """x [B,F,T,Cin] -> y [B,F/2,T,Cout] (ComplexConv2d, tools_for_model.py:243-269)."""
    _req(x, wr, br, wi, bi)
    B, F, T, Cin = x.shape
    Cout = 2 * wr.shape[0]
    y = torch.empty(B, F // 2, T, Cout, device=x.device)
    ws = _cconv_ws(Cin, Cout, x.device)
    _lib.check(_lib.load().sefd_cconv2d_forward(ptr(x), ptr(wr), ptr(br), ptr(wi), ptr(bi), ptr(y), B, F, T, Cin, Cout,
                                                ptr(ws), stream()), "cconv2d_forward")
    return y


def cconv2d_backward(x, wr, wi, dy, need_dx=True):
    _req(x, wr, wi, dy)
    B, F, T, Cin = x.shape
    Cout = dy.shape[-1]
    dx = torch.empty_like(x) if need_dx else None
    if need_dx:
        dx.zero_()
    dwr, dwi = torch.empty_like(wr), torch.empty_like(wi)
    dbr = torch.empty(Cout // 2, device=x.device)
    dbi = torch.empty(Cout // 2, device=x.device)
    ws = _cconv_ws(Cin, Cout, x.device)
    _lib.check(_lib.load().sefd_cconv2d_backward(ptr(x), ptr(wr), ptr(wi), ptr(dy), ptr(dx), ptr(dwr), ptr(dbr),
                                                 ptr(dwi), ptr(dbi), B, F, T, Cin, Cout, ptr(ws), stream()),
               "cconv2d_backward")
    return dx, dwr, dbr, dwi, dbi


def cconvT2d_forward(x0, x1, wr, br, wi, bi):
    """x0, x1 [B,F,T,Cin/2] (x1 = skip) -> y [B,2F,T+1,Cout] (ComplexConvTranspose2d on complex_cat)."""
    _req(x0, x1, wr, br, wi, bi)
    B, F, T, Ch = x0.shape
    Cin, Cout = 2 * Ch, 2 * wr.shape[1]
    y = torch.empty(B, 2 * F, T + 1, Cout, device=x0.device)
    ws = _cconv_ws(Cin, Cout, x0.device)
    _lib.check(_lib.load().sefd_cconvT2d_forward(ptr(x0), ptr(x1), ptr(wr), ptr(br), ptr(wi), ptr(bi), ptr(y), B, F, T,
                                                 Cin, Cout, ptr(ws), stream()), "cconvT2d_forward")
    return y


def cconvT2d_backward(x0, x1, wr, wi, dy):
    _req(x0, x1, wr, wi, dy)
    B, F, T, Ch = x0.shape
    Cin, Cout = 2 * Ch, dy.shape[-1]
    dx0, dx1 = torch.empty_like(x0), torch.empty_like(x1)
    dwr, dwi = torch.empty_like(wr), torch.empty_like(wi)
    dbr = torch.empty(Cout // 2, device=x0.device)
    dbi = torch.empty(Cout // 2, device=x0.device)
    ws = _cconv_ws(Cin, Cout, x0.device)
    _lib.check(_lib.load().sefd_cconvT2d_backward(ptr(x0), ptr(x1), ptr(wr), ptr(wi), ptr(dy), ptr(dx0), ptr(dx1),
                                                  ptr(dwr), ptr(dbr), ptr(dwi), ptr(dbi), B, F, T, Cin, Cout, ptr(ws),
                                                  stream()), "cconvT2d_backward")
    return dx0, dx1, dwr, dbr, dwi, dbi


# ---- BatchNorm + PReLU ------------------------------------------------------------------------
def bn_prelu_forward(y, gamma, beta, alpha, running_mean=None, running_var=None):
    """y [rows, C] -> (z, save[2,C])  (nn.BatchNorm2d train mode + nn.PReLU, models.py:76-78)."""
    _req(y, gamma, beta, alpha, running_mean, running_var)
    rows, Cc = y.shape
    z = torch.empty_like(y)
    save = torch.empty(2, Cc, device=y.device)
    scratch = torch.empty(2 * Cc + 1, device=y.device, dtype=torch.float64)
    _lib.check(_lib.load().sefd_bn_prelu_forward(ptr(y), ptr(z), rows, Cc, ptr(gamma), ptr(beta), ptr(alpha), ptr(save),
                                                 ptr(running_mean), ptr(running_var), ptr(scratch), stream()),
               "bn_prelu_forward")
    return z, save


def bn_prelu_backward(y, dz, gamma, beta, alpha, save):
    _req(y, dz, gamma, beta, alpha, save)
    rows, Cc = y.shape
    dy = torch.empty_like(y)
    dg, db = torch.empty_like(gamma), torch.empty_like(beta)
    da = torch.empty_like(alpha)
    scratch = torch.empty(2 * Cc + 1, device=y.device, dtype=torch.float64)
    _lib.check(_lib.load().sefd_bn_prelu_backward(ptr(y), ptr(dz), ptr(dy), rows, Cc, ptr(gamma), ptr(beta), ptr(alpha),
                                                  ptr(save), ptr(dg), ptr(db), ptr(da), ptr(scratch), stream()),
               "bn_prelu_backward")
    return dy, dg, db, da


class _Cbn(torch.autograd.Function):
    """ComplexBatchNorm on channels-last rows [rows, C] (sefd_cbn_prelu_forward / backward with a PReLU slope of 1)."""

    @staticmethod
    def forward(ctx, y, w3h, b2h, running5h, train):
        _req(y, w3h, b2h, running5h)
        rows, Cc = y.shape
        h = Cc // 2
        z = torch.empty_like(y)
        save = torch.empty(9 * h, device=y.device)
        one = torch.ones(1, device=y.device)
        scratch = torch.empty(6 * h + 1, device=y.device, dtype=torch.float64)
        _lib.check(_lib.load().sefd_cbn_prelu_forward(ptr(y), ptr(z), rows, Cc, ptr(w3h), ptr(b2h), ptr(one), ptr(save),
                                                      ptr(running5h), 0 if train else 1, ptr(scratch), stream()), "cbn_prelu_forward")
        ctx.save_for_backward(y, w3h, b2h, save, one)
        ctx.train = train
        return z

    @staticmethod
    def backward(ctx, dz):
        if not ctx.train:
            raise RuntimeError("sefd: ComplexBatchNorm backward is built for train mode (batch statistics)")
        y, w3h, b2h, save, one = ctx.saved_tensors
        dz = dz.contiguous()
        rows, Cc = y.shape
        h = Cc // 2
        dy, dw, db, da = torch.empty_like(y), torch.empty_like(w3h), torch.empty_like(b2h), torch.empty_like(one)
        scratch = torch.empty(6 * h + 1, device=y.device, dtype=torch.float64)
        coef = torch.empty(9 * h, device=y.device)
        _lib.check(_lib.load().sefd_cbn_prelu_backward(ptr(y), ptr(dz), ptr(dy), rows, Cc, ptr(w3h), ptr(b2h), ptr(one), ptr(save),
                                                       ptr(dw), ptr(db), ptr(da), ptr(scratch), ptr(coef), stream()), "cbn_prelu_backward")
        return dy, dw, db, None, None


def complex_batch_norm(y, w3h, b2h, running5h, train):
    return _Cbn.apply(y, w3h, b2h, running5h, train)


# ---- LSTM recurrence ------------------------------------------------------------------------------
def lstm_forward(w_hh, pregates):
    """w_hh [2,512,128]; pregates [2,rows,T,512] (consumed: overwritten with the activated gates)."""
    _req(w_hh, pregates)
    _, rows, T, _ = pregates.shape
    h = torch.empty(2, rows, T, 128, device=w_hh.device)
    c = torch.empty(2, rows, T, 128, device=w_hh.device)
    _lib.check(_lib.load().sefd_lstm_forward(ptr(w_hh), ptr(pregates), ptr(h), ptr(c), rows, T, stream()),
               "lstm_forward")
    return h, c


def lstm_backward(w_hh, gates, c, dh):
    _req(w_hh, gates, c, dh)
    _, rows, T, _ = gates.shape
    dg = torch.empty_like(gates)
    _lib.check(_lib.load().sefd_lstm_backward(ptr(w_hh), ptr(gates), ptr(c), ptr(dh), ptr(dg), rows, T, stream()),
               "lstm_backward")
    return dg


# ---- Adam -------------------------------------------------------------------------------------------
def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, gscale=1.0):
    _req(params, grads, exp_avg, exp_avg_sq)
    _lib.check(_lib.load().sefd_adam_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), params.numel(), lr,
                                          betas[0], betas[1], eps, step, gscale, stream()), "adam_step")


# ---- LMS perceptual loss (tools_for_loss.py:111-249) -------------------------------------------------
_MEL = {}


def mel_filterbank(num_coeffs, fft_size=512, fs=16000):
    """melFilterBank(numCoeffs, fftSize) of the reference (tools_for_loss.py:144-188), restated with the same numpy
    operations so that the float32 rounding of the band edges follows the installed numpy exactly as the reference's
    would: [num_coeffs, fft_size // 2 + 1] triangular filters."""
    import math
    import numpy as np
    max_hz, bins = fs / 2, fft_size // 2 + 1
    f2m = lambda f: 1127.01048 * math.log(1 + f / 700.0)
    m2f = lambda m: 700 * (math.exp(m / 1127.01048) - 1)
    max_mel, min_mel = f2m(max_hz), f2m(0)
    centres = np.array(range(num_coeffs + 2)).astype(np.float32)
    centres = centres * (max_mel - min_mel) / (num_coeffs + 1) + min_mel
    for i in range(num_coeffs + 2):
        centres[i] = m2f(centres[i])
        centres[i] = math.floor(bins * centres[i] / max_hz)
    mat = np.zeros((num_coeffs, bins))
    for i in range(1, num_coeffs + 1):
        lo, mid, hi = int(centres[i - 1]), int(centres[i]), int(centres[i + 1])
        for j in range(lo, mid):
            mat[i - 1, j] = (float(j) - lo) / (mid - lo)
        for j in range(mid, hi):
            mat[i - 1, j] = 1 - ((float(j) - mid) / (hi - mid))
    return mat


def _mel_matrices(device):
    key = str(device)
    if key not in _MEL:
        import numpy as np
        f = np.concatenate([mel_filterbank(m) for m in (16, 32, 64)], 0).astype(np.float32)     # [112, 257]
        ft = torch.from_numpy(f).contiguous().to(device)
        _MEL[key] = (ft.t().contiguous(), ft)                                                     # F [257,112], Ft [112,257]
    return _MEL[key]


class _Lms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, clean, mags):
        _req(a, b, clean)
        B, _, T = a.shape
        F, Ft = _mel_matrices(a.device)
        scratch = torch.empty(1, device=a.device, dtype=torch.float64)
        loss = torch.empty(1, device=a.device)
        _lib.check(_lib.load().sefd_lms_forward(ptr(a), ptr(b), ptr(clean), ptr(F), B, T, int(mags), ptr(scratch), ptr(loss),
                                                stream()), "lms_forward")
        ctx.save_for_backward(a, b if b is not None else a, clean)
        ctx.mags = mags
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        a, b, clean = ctx.saved_tensors
        B, _, T = a.shape
        F, Ft = _mel_matrices(a.device)
        gout = gout.contiguous().float()
        da = torch.empty_like(a)
        db = None if ctx.mags else torch.empty_like(a)
        _lib.check(_lib.load().sefd_lms_backward(ptr(a), None if ctx.mags else ptr(b), ptr(clean), ptr(F), ptr(Ft), ptr(gout),
                                                 B, T, int(ctx.mags), ptr(da), ptr(db), stream()), "lms_backward")
        return da, db, None, None


def lms_loss_spec(est_real, est_imag, target_wav):
    """DCCRN.loss(..., perceptual=True) with cfg.perceptual == 'LMS' (models.py:305-312): est_real / est_imag [B,257,T]
    masked spectrum, target_wav [B,L].  Gradients flow to est_real and est_imag."""
    clean = stft(target_wav)                                                       # [B,257,T,2]
    return _Lms.apply(est_real.contiguous(), est_imag.contiguous(), clean, False)


def lms_loss_mags(clean_mags, est_mags):
    """get_array_lms_loss(clean_array, est_array) (tools_for_loss.py:241-249) on magnitude arrays [B,257,T]."""
    return _Lms.apply(est_mags.contiguous(), None, clean_mags.contiguous(), True)


# ---- PMSQE perceptual loss (tools_for_loss.py:255-269; arithmetic of asteroid's SingleSrcPMSQE: parity unpinned) -------
_PMSQE = {}


def pmsqe_tables(bark_matrix=None, abs_thresh_power=None, modified_zwicker_power=None, width_of_band_bark=None,
                 mask_sll=None):
    """Packed float32 table vector of sefd_pmsqe_*: [bark 257x49 | thresholds 49 | Zwicker powers 49 | widths 49 | SLL mask
    257].  Defaults are SingleSrcPMSQE's 16 kHz constants rebuilt from the ITU-T P.862 tables (sefd/p862_16k.py): Bark
    matrix = band membership of the 256 bins x pow_dens_correction_factor, modified Zwicker power
    0.23 * clip(6 / (centre + 2), 1, 2) ** 0.15, SLL mask over bins 11..104 times the sqrt-hann power correction
    2 * (512 + 2) / 512**2.  Any argument may be replaced by asteroid's own buffer of the same name."""
    import numpy as np
    from . import p862_16k as p
    if bark_matrix is None:
        bark_matrix = np.zeros((257, 49))
        f = 0
        for k, n in enumerate(p.NR_OF_HZ_BANDS_PER_BARK_BAND):
            bark_matrix[f:f + n, k] = p.POW_DENS_CORRECTION_FACTOR[k]
            f += n
    if abs_thresh_power is None:
        abs_thresh_power = p.ABS_THRESH_POWER
    if modified_zwicker_power is None:
        modified_zwicker_power = 0.23 * np.clip(6.0 / (np.asarray(p.CENTRE_OF_BAND_BARK) + 2.0), 1.0, 2.0) ** 0.15
    if width_of_band_bark is None:
        width_of_band_bark = p.WIDTH_OF_BAND_BARK
    if mask_sll is None:
        mask_sll = np.zeros(257)
        mask_sll[11], mask_sll[12:104], mask_sll[104] = 0.5 * 25.0 / 31.25, 1.0, 0.5
        mask_sll = mask_sll * 2.0 * (512 + 2.0) / 512 ** 2
    parts = [np.asarray(bark_matrix, dtype=np.float64).reshape(257 * 49), np.asarray(abs_thresh_power, dtype=np.float64),
             np.asarray(modified_zwicker_power, dtype=np.float64), np.asarray(width_of_band_bark, dtype=np.float64),
             np.asarray(mask_sll, dtype=np.float64)]
    assert [a.size for a in parts] == [257 * 49, 49, 49, 49, 257]
    return torch.from_numpy(np.concatenate(parts).astype(np.float32))


def _pmsqe_tables(device):
    key = str(device)
    if key not in _PMSQE:
        t = pmsqe_tables()
        assert t.numel() == _lib.load().sefd_pmsqe_table_floats()
        _PMSQE[key] = t.to(device)
    return _PMSQE[key]


class _Pmsqe(torch.autograd.Function):
    @staticmethod
    def forward(ctx, est, clean, tables):
        _req(est, clean, tables)
        N, L = est.shape
        lib = _lib.load()
        nbytes = lib.sefd_pmsqe_workspace_bytes(N, L)
        if nbytes == 0:
            raise ValueError(f"PMSQE: waveforms must be 1..4 whole seconds at 16 kHz (tools_for_loss.py:264), got {L} samples")
        ws = torch.empty(nbytes, device=est.device, dtype=torch.uint8)
        loss = torch.empty(1, device=est.device)
        _lib.check(lib.sefd_pmsqe_forward(ptr(est), ptr(clean), N, L, ptr(tables), ptr(ws), nbytes, ptr(loss), stream()),
                   "pmsqe_forward")
        ctx.save_for_backward(ws, tables)
        ctx.shape = (N, L)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        ws, tables = ctx.saved_tensors
        N, L = ctx.shape
        gout = gout.contiguous().float()
        d = torch.empty(N, L, device=ws.device)
        _lib.check(_lib.load().sefd_pmsqe_backward(ptr(gout), N, L, ptr(tables), ptr(ws), ws.numel(), ptr(d), stream()),
                   "pmsqe_backward")
        return d, None, None


def pmsqe_loss(clean_wav, est_wav, tables=None):
    """get_array_pmsqe_loss(clean_array, est_array) (tools_for_loss.py:259-269): [N, L] waveforms (or [N, 1, L]), L a whole
    number of seconds; differentiable with respect to est_wav."""
    if clean_wav.dim() == 3:
        clean_wav, est_wav = clean_wav.reshape(clean_wav.shape[0], -1), est_wav.reshape(est_wav.shape[0], -1)
    tables = _pmsqe_tables(est_wav.device) if tables is None else tables.to(est_wav.device).float().contiguous()
    return _Pmsqe.apply(est_wav.contiguous().float(), clean_wav.contiguous().float(), tables)


# ---- FullSubNet feature / target side (trainer.fullsubnet_train, trainer.py:97-104) ---------------------------------------
def fsn_stft(y):
    """tools.stft (tools_for_model.py:628-648): [B, L] -> complex64 [B, 257, L // 300 + 1]."""
    _req(y)
    B, L = y.shape
    T = _lib.load().sefd_fsn_frames(L)
    out = torch.empty(B, 257, T, 2, device=y.device)
    _lib.check(_lib.load().sefd_fsn_stft(ptr(y), B, L, ptr(out), stream()), "fsn_stft")
    return torch.view_as_complex(out)


def fsn_features(noisy, clean):
    """The feature / target computation of the fullsubnet_train loop in one kernel: (noisy_mag [B,257,T], cIRM [B,257,T,2])."""
    _req(noisy, clean)
    B, L = noisy.shape
    T = _lib.load().sefd_fsn_frames(L)
    mag = torch.empty(B, 257, T, device=noisy.device)
    cirm = torch.empty(B, 257, T, 2, device=noisy.device)
    _lib.check(_lib.load().sefd_fsn_features(ptr(noisy), ptr(clean), B, L, ptr(mag), ptr(cirm), stream()), "fsn_features")
    return mag, cirm


def fsn_mag_phase(c):
    r = torch.view_as_real(c).contiguous()
    _req(r)
    mag, phase = torch.empty(c.shape, device=c.device), torch.empty(c.shape, device=c.device)
    _lib.check(_lib.load().sefd_fsn_mag_phase(ptr(r), c.numel(), ptr(mag), ptr(phase), stream()), "fsn_mag_phase")
    return mag, phase


def fsn_cirm(noisy, clean):
    a, b = torch.view_as_real(noisy).contiguous(), torch.view_as_real(clean).contiguous()
    _req(a, b)
    out = torch.empty_like(a)
    _lib.check(_lib.load().sefd_fsn_cirm(ptr(a), ptr(b), noisy.numel(), ptr(out), stream()), "fsn_cirm")
    return out


def fsn_compress_cirm(mask):
    m = mask.contiguous().float()
    _req(m)
    out = torch.empty_like(m)
    _lib.check(_lib.load().sefd_fsn_compress_cirm(ptr(m), m.numel(), ptr(out), stream()), "fsn_compress_cirm")
    return out


def fsn_decompress_cirm(mask):
    m = mask.contiguous().float()
    _req(m)
    out = torch.empty_like(m)
    _lib.check(_lib.load().sefd_fsn_decompress_cirm(ptr(m), m.numel(), ptr(out), stream()), "fsn_decompress_cirm")
    return out


def fsn_istft(features, length=None, use_mag_phase=False):
    """tools.istft (tools_for_model.py:651-679): complex [B, 257, T] (or its real view [B, 257, T, 2], or (mag, phase) with
    use_mag_phase) -> [B, length]; length defaults to 300 (T - 1) like torch.istft."""
    if use_mag_phase:
        mag, phase = features
        a, b = mag.contiguous().float(), phase.contiguous().float()
        _req(a, b)
        B, _, T = a.shape
    else:
        a = (torch.view_as_real(features) if torch.is_complex(features) else features).contiguous().float()
        b = None
        _req(a)
        B, _, T, _ = a.shape
    n = 300 * (T - 1) if length is None else int(length)
    out = torch.empty(B, n, device=a.device)
    _lib.check(_lib.load().sefd_fsn_istft(ptr(a), ptr(b), B, T, n, ptr(out), stream()), "fsn_istft")
    return out
