"""Drop-in for the reference's models.py (same import surface: `from models import DCCRN, CRN, FullSubNet`,
train_interface.py:7), backed by the sefd CUDA library instead of torch ops.

DCCRN mirrors models.py:15-323: zero-argument constructor reading `config`, identical state_dict keys and
shapes, identical initial values for a given torch seed, forward(inputs, targets=0) ->
(out_real [B,257,T], out_imag [B,257,T], out_wav [B,L]) and loss(estimated, target, ...).
"""
import torch
import torch.nn as nn

try:
    import config as cfg
except ImportError:                      # no reference checkout on the path
    from sefd import default_config as cfg

from sefd import dccrn as _d
from sefd import ops as _ops
import tools_for_loss as _tfl


class DCCRN(nn.Module):
    def __init__(self, rnn_layers=cfg.rnn_layers, rnn_units=cfg.rnn_units, win_len=cfg.win_len,
                 win_inc=cfg.win_inc, fft_len=cfg.fft_len, win_type=cfg.window, masking_mode=cfg.masking_mode,
                 use_cbn=False, kernel_size=5):
        super().__init__()
        kernel_num = list(cfg.dccrn_kernel_num)
        unsupported = []
        if (win_len, win_inc, fft_len) != (400, 100, 512):
            unsupported.append(f"STFT geometry {(win_len, win_inc, fft_len)} (built: (400, 100, 512))")
        if win_type not in ("hanning", "hann"):
            unsupported.append(f"window {win_type!r} (built: periodic Hann)")
        if kernel_num != _d.KERNEL_NUM or kernel_size != 5:
            unsupported.append(f"kernel_num {kernel_num} / kernel_size {kernel_size}")
        if rnn_layers != 2 or rnn_units != 256 or cfg.lstm != "complex":
            unsupported.append(f"rnn_layers={rnn_layers}, rnn_units={rnn_units}, lstm={cfg.lstm!r} (built: 2 x complex LSTM 256)")
        if use_cbn:
            unsupported.append("use_cbn=True (ComplexBatchNorm, tools_for_model.py:430-603)")
        if masking_mode not in _ops.MODES:
            unsupported.append(f"masking_mode {masking_mode!r} (built: E, C, R, Direct(None make))")
        if unsupported:
            raise NotImplementedError("sefd DCCRN: configuration outside the built hot path: " + "; ".join(unsupported))

        self.win_len, self.win_inc, self.fft_len, self.win_type = win_len, win_inc, fft_len, win_type
        self.rnn_units, self.hidden_layers, self.kernel_size = rnn_units, rnn_layers, kernel_size
        self.kernel_num = [2] + kernel_num
        self.masking_mode = masking_mode
        self.skip_type = bool(cfg.skip_type)

        self.stft = _d.STFTBuffers(win_len, fft_len, inverse=False)
        self.istft = _d.STFTBuffers(win_len, fft_len, inverse=True)
        self.encoder = nn.ModuleList()
        self.decoder = nn.ModuleList()
        kn = self.kernel_num
        for i in range(len(kn) - 1):                                     # models.py:63-80
            self.encoder.append(nn.Sequential(_d.ComplexConvParams(kn[i], kn[i + 1], transposed=False),
                                              _d.BatchNormParams(kn[i + 1]), _d.PReLUParams()))
        hidden_dim = fft_len // (2 ** len(kn))
        rnns = []
        for i in range(rnn_layers):                                      # models.py:83-95
            rnns.append(_d.ComplexLSTMParams(
                input_size=hidden_dim * kn[-1] if i == 0 else rnn_units, hidden_size=rnn_units,
                projection_dim=hidden_dim * kn[-1] if i == rnn_layers - 1 else None))
        self.enhance = nn.Sequential(*rnns)
        for idx in range(len(kn) - 1, 0, -1):                            # models.py:107-137
            mods = [_d.ComplexConvParams(kn[idx] * (2 if self.skip_type else 1), kn[idx - 1], transposed=True)]   # :138-169 without skip
            if idx != 1:
                mods += [_d.BatchNormParams(kn[idx - 1]), _d.PReLUParams()]
            self.decoder.append(nn.Sequential(*mods))
        self._engine = None
        self._last = None

    # nn.Module would try to register the engine's tensors otherwise
    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            eng = _d.Engine(self, self.masking_mode, skip=self.skip_type)
            self.__dict__["_engine"] = eng
        return eng

    def flatten_parameters(self):
        self._get_engine().sync()

    def forward(self, inputs, targets=0):
        eng = self._get_engine()
        tgt = targets if torch.is_tensor(targets) and targets.shape == inputs.shape else None
        out_real, out_imag, out_wav = eng.forward(inputs, tgt, self.training)
        if self.training:
            for m in self.modules():
                if isinstance(m, _d.BatchNormParams):
                    m.num_batches_tracked += 1
        self.__dict__["_last"] = (out_wav, tgt)
        if self.masking_mode == "Direct(None make)":                     # spectral mapping, models.py:232-250
            if tgt is None:
                raise ValueError("Direct(None make) needs the target waveforms (models.py:234)")
            tspec = _ops.stft(tgt)                                       # [B,257,T,2]
            return out_real, tspec[..., 0], out_imag, tspec[..., 1], out_wav
        return out_real, out_imag, out_wav

    def get_params(self, weight_decay=0.0):
        weights, biases = [], []
        for name, param in self.named_parameters():
            (biases if "bias" in name else weights).append(param)
        return [{"params": weights, "weight_decay": weight_decay}, {"params": biases, "weight_decay": 0.0}]

    def loss(self, estimated, target, real_spec=0, img_spec=0, perceptual=False):
        if perceptual:                                                   # models.py:304-314
            if cfg.perceptual == "LMS":
                # clean_mags = sqrt(|STFT(target)|^2 + 1e-7), est_mags = sqrt(real_spec^2 + img_spec^2 + 1e-7) and the
                # log-mel distance are one fused kernel; gradients reach the mask through real_spec / img_spec
                return _ops.lms_loss_spec(real_spec, img_spec, target)
            return _tfl.get_array_pmsqe_loss(target, estimated)
        if cfg.loss not in _ops.LOSSES:
            raise NotImplementedError(f"loss {cfg.loss!r}")
        return _ops.loss(estimated, target, cfg.loss)


def _not_built(name, row):
    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            raise NotImplementedError(f"sefd: {name} is not built yet ({row}); only the DCCRN hot path is. "
                                      "There is deliberately no PyTorch fallback.")
    _Stub.__name__ = name
    return _Stub


class CRN(nn.Module):
    """Drop-in for models.py:329-565: real-valued conv recurrent network with a magnitude T-F mask.
    forward(inputs, targets=0) -> (est_mags [B,257,T], target_mags [B,257,T], out_wav [B,L])."""

    def __init__(self, rnn_layers=cfg.rnn_layers, rnn_input_size=getattr(cfg, "rnn_input_size", 512),
                 rnn_units=cfg.rnn_units, win_len=cfg.win_len, win_inc=cfg.win_inc, fft_len=cfg.fft_len,
                 win_type=cfg.window, masking_mode=cfg.masking_mode, kernel_size=5):
        super().__init__()
        kernel_num = list(cfg.dccrn_kernel_num)
        unsupported = []
        if (win_len, win_inc, fft_len) != (400, 100, 512):
            unsupported.append(f"STFT geometry {(win_len, win_inc, fft_len)} (built: (400, 100, 512))")
        if win_type not in ("hanning", "hann"):
            unsupported.append(f"window {win_type!r} (built: periodic Hann)")
        if kernel_num != _d.KERNEL_NUM or kernel_size != 5:
            unsupported.append(f"kernel_num {kernel_num} / kernel_size {kernel_size}")
        if rnn_units != 256 or rnn_input_size != 512:
            unsupported.append(f"rnn_units={rnn_units}, rnn_input_size={rnn_input_size} (built: 256 / 512)")
        if not cfg.skip_type:
            unsupported.append("skip_type=False")
        if masking_mode == "Direct(None make)":
            unsupported.append("Direct spectral mapping (models.py:507-516)")
        if unsupported:
            raise NotImplementedError("sefd CRN: configuration outside the built path: " + "; ".join(unsupported))
        self.win_len, self.win_inc, self.fft_len, self.win_type = win_len, win_inc, fft_len, win_type
        self.rnn_input_size, self.rnn_units = rnn_input_size, rnn_units // 2
        self.hidden_layers, self.kernel_size = rnn_layers, kernel_size      # rnn_layers is not passed to nn.LSTM (models.py:391-397)
        self.kernel_num = [2] + kernel_num
        self.masking_mode = masking_mode
        self.stft = _d.STFTBuffers(win_len, fft_len, inverse=False)
        self.istft = _d.STFTBuffers(win_len, fft_len, inverse=True)
        self.encoder = nn.ModuleList()
        self.decoder = nn.ModuleList()
        kn = self.kernel_num
        for i in range(len(kn) - 1):                                     # models.py:376-389
            self.encoder.append(nn.Sequential(_d.RealConvParams(kn[i] // 2, kn[i + 1] // 2, transposed=False),
                                              _d.BatchNormParams(kn[i + 1] // 2), _d.PReLUParams()))
        self.enhance = _d.LSTMParams(self.rnn_input_size, self.rnn_units)   # models.py:391-397
        self.tranform = _d.LinearParams(self.rnn_units, self.rnn_input_size)   # (sic) models.py:398
        for idx in range(len(kn) - 1, 0, -1):                            # models.py:400-430
            mods = [_d.RealConvParams(kn[idx], kn[idx - 1] // 2, transposed=True)]
            if idx != 1:
                mods += [_d.BatchNormParams(kn[idx - 1] // 2), _d.PReLUParams()]
            self.decoder.append(nn.Sequential(*mods))

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            eng = _d.Engine(self, "E", family="crn")
            self.__dict__["_engine"] = eng
        return eng

    def flatten_parameters(self):
        self._get_engine().sync()

    def forward(self, inputs, targets=0):
        eng = self._get_engine()
        tgt = targets if torch.is_tensor(targets) and targets.shape == inputs.shape else None
        est_mags, target_mags, out_wav = eng.forward(inputs, tgt, self.training)
        if self.training:
            for m in self.modules():
                if isinstance(m, _d.BatchNormParams):
                    m.num_batches_tracked += 1
        return est_mags, target_mags, out_wav

    get_params = DCCRN.get_params

    def loss(self, estimated, target, out_mags=0, target_mags=0, perceptual=False):
        if perceptual:                                                   # models.py:553-557
            if cfg.perceptual == "LMS":
                return _tfl.get_array_lms_loss(target_mags, out_mags)
            return _tfl.get_array_pmsqe_loss(target, estimated)
        if cfg.loss not in _ops.LOSSES:
            raise NotImplementedError(f"loss {cfg.loss!r}")
        return _ops.loss(estimated, target, cfg.loss)


FullSubNet = _not_built("FullSubNet", "SURVEY.md §8 a14, BASELINE config 3")
