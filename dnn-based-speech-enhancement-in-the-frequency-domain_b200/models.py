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
        if rnn_layers != 2 or rnn_units != 256 or cfg.lstm not in ("complex", "real"):
            unsupported.append(f"rnn_layers={rnn_layers}, rnn_units={rnn_units}, lstm={cfg.lstm!r} (built: 2 layers, 256 units, 'complex' / 'real')")
        if masking_mode not in _ops.MODES:
            unsupported.append(f"masking_mode {masking_mode!r} (built: E, C, R, Direct(None make))")
        if unsupported:
            raise NotImplementedError("sefd DCCRN: configuration outside the built hot path: " + "; ".join(unsupported))

        self.win_len, self.win_inc, self.fft_len, self.win_type = win_len, win_inc, fft_len, win_type
        self.rnn_units, self.hidden_layers, self.kernel_size = rnn_units, rnn_layers, kernel_size
        self.kernel_num = [2] + kernel_num
        self.masking_mode = masking_mode
        self.skip_type = bool(cfg.skip_type)
        self.use_cbn = bool(use_cbn)
        norm = _d.ComplexBatchNormParams if use_cbn else _d.BatchNormParams      # models.py:76, 120, 151

        self.stft = _d.STFTBuffers(win_len, fft_len, inverse=False)
        self.istft = _d.STFTBuffers(win_len, fft_len, inverse=True)
        self.encoder = nn.ModuleList()
        self.decoder = nn.ModuleList()
        kn = self.kernel_num
        for i in range(len(kn) - 1):                                     # models.py:63-80
            self.encoder.append(nn.Sequential(_d.ComplexConvParams(kn[i], kn[i + 1], transposed=False),
                                              norm(kn[i + 1]), _d.PReLUParams()))
        hidden_dim = fft_len // (2 ** len(kn))
        self.lstm_type = cfg.lstm
        if cfg.lstm == "complex":
            rnns = []
            for i in range(rnn_layers):                                  # models.py:83-95
                rnns.append(_d.ComplexLSTMParams(
                    input_size=hidden_dim * kn[-1] if i == 0 else rnn_units, hidden_size=rnn_units,
                    projection_dim=hidden_dim * kn[-1] if i == rnn_layers - 1 else None))
            self.enhance = nn.Sequential(*rnns)
        else:                                                            # models.py:96-105: nn.LSTM(1024 -> 256, 2 layers) + Linear
            from sefd.fullsubnet import StackedLSTMParams
            self.enhance = StackedLSTMParams(hidden_dim * kn[-1], rnn_units, num_layers=2, dropout=0.0)
            self.tranform = _d.LinearParams(rnn_units, hidden_dim * kn[-1])
        for idx in range(len(kn) - 1, 0, -1):                            # models.py:107-137
            mods = [_d.ComplexConvParams(kn[idx] * (2 if self.skip_type else 1), kn[idx - 1], transposed=True)]   # :138-169 without skip
            if idx != 1:
                mods += [norm(kn[idx - 1]), _d.PReLUParams()]
            self.decoder.append(nn.Sequential(*mods))
        self._engine = None
        self._last = None

    # nn.Module would try to register the engine's tensors otherwise
    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            eng = _d.Engine(self, self.masking_mode, skip=self.skip_type, real_lstm=self.lstm_type == "real", cbn=self.use_cbn)
            self.__dict__["_engine"] = eng
        return eng

    def flatten_parameters(self):
        self._get_engine().sync()

    def forward(self, inputs, targets=0):
        eng = self._get_engine()
        tgt = targets if torch.is_tensor(targets) and targets.shape == inputs.shape else None
        out_real, out_imag, out_wav = eng.forward(inputs, tgt, self.training)
        if self.training:
            _d.bump_batches_tracked(self)
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
            _d.bump_batches_tracked(self)
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


class FullSubNet(nn.Module):
    """Drop-in for models.py:568-682: forward(noisy_mag [B, 257, T] or [B, 1, 257, T]) -> cRM [B, 257, T, 2];
    loss(estimated, target) with cfg.loss = 'MSE' (trainer.fullsubnet_train, trainer.py:97-107).

    The reference's nn.LSTM(dropout=0.8) between the stacked layers (tools_for_model.py:746) is active in train() mode:
    the masks come from a Philox stream seeded from torch's generator (one draw per forward), or from `dropout_masks`
    = (mask_fb [T+2, B, 512], mask_sb [T+2, B*257, 384]) multipliers when a caller injects them (parity tests)."""

    def __init__(self, sb_num_neighbors=getattr(cfg, "sb_num_neighbors", 15), fb_num_neighbors=getattr(cfg, "fb_num_neighbors", 0),
                 num_freqs=getattr(cfg, "num_freqs", 257), look_ahead=getattr(cfg, "look_ahead", 2),
                 sequence_model=getattr(cfg, "sequence_model", "LSTM"),
                 fb_output_activate_function=getattr(cfg, "fb_output_activate_function", "ReLU"),
                 sb_output_activate_function=getattr(cfg, "sb_output_activate_function", None),
                 fb_model_hidden_size=getattr(cfg, "fb_model_hidden_size", 512),
                 sb_model_hidden_size=getattr(cfg, "sb_model_hidden_size", 384),
                 weight_init=getattr(cfg, "weight_init", False), norm_type=getattr(cfg, "norm_type", "offline_laplace_norm")):
        super().__init__()
        got = (sb_num_neighbors, fb_num_neighbors, num_freqs, look_ahead, sequence_model, fb_output_activate_function,
               sb_output_activate_function, fb_model_hidden_size, sb_model_hidden_size, bool(weight_init), norm_type)
        built = (15, 0, 257, 2, "LSTM", "ReLU", None, 512, 384, False, "offline_laplace_norm")
        if got != built:
            raise NotImplementedError(f"sefd FullSubNet: configuration {got} outside the built path {built} (config.py:71-80)")
        from sefd import fullsubnet as _f
        self.fb_model = _f.SequenceModelParams(num_freqs, num_freqs, fb_model_hidden_size)                   # models.py:600-608
        self.sb_model = _f.SequenceModelParams((sb_num_neighbors * 2 + 1) + (fb_num_neighbors * 2 + 1), 2,  # models.py:610-618
                                               sb_model_hidden_size)
        self.sb_num_neighbors, self.fb_num_neighbors, self.look_ahead = sb_num_neighbors, fb_num_neighbors, look_ahead
        self.dropout = 0.8                                                                                # tools_for_model.py:746
        self.dropout_masks = None

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            eng = _d.Engine(self, None, family="fsn")
            self.__dict__["_engine"] = eng
        return eng

    def flatten_parameters(self):
        self._get_engine().sync()

    def forward(self, noisy_mag):
        from sefd import fullsubnet as _f
        if noisy_mag.dim() == 4:
            if noisy_mag.shape[1] != 1:
                raise ValueError("FullSubNet takes the mag feature as inputs (one channel, models.py:642)")
            noisy_mag = noisy_mag[:, 0]
        noisy_mag = noisy_mag.contiguous().float()
        _ops._req(noisy_mag)
        if noisy_mag.shape[1] != 257:
            raise ValueError(f"FullSubNet: expected 257 frequency bins, got {noisy_mag.shape[1]}")
        eng = self._get_engine()
        eng.sync()
        mask_fb = mask_sb = None
        seed = 0
        p = self.dropout if self.training else 0.0
        if self.training and p > 0:
            if self.dropout_masks is not None:
                mask_fb, mask_sb = (m.contiguous().float() for m in self.dropout_masks)
                T, B = noisy_mag.shape[2] + self.look_ahead, noisy_mag.shape[0]
                if tuple(mask_fb.shape) != (T, B, 512) or tuple(mask_sb.shape) != (T, B * 257, 384):
                    raise ValueError("FullSubNet.dropout_masks: expected [T+2, B, 512] and [T+2, B*257, 384]")
                _ops._req(mask_fb, mask_sb)
            else:
                seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        params = [q for q, _, _, _ in eng.param_list]
        return _f._ForwardFSN.apply(eng, noisy_mag, self.training, p, mask_fb, mask_sb, seed, *params)

    def loss(self, estimated, target):
        # called as model.loss(cIRM, cRM) (trainer.py:107): the second argument carries the gradient.  MSE is symmetric.
        if cfg.loss != "MSE":
            raise NotImplementedError(f"sefd FullSubNet: loss {cfg.loss!r} (built: 'MSE', models.py:675-676)")
        a, b = (target, estimated) if target.requires_grad or not estimated.requires_grad else (estimated, target)
        return _ops.loss(a.reshape(a.shape[0], -1), b.detach().reshape(b.shape[0], -1), "MSE")
