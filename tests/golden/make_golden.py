"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Imports /root/reference/{config,models,tools_for_model,tools_for_loss}.py through the shim of
SURVEY.md appendix A (stub matplotlib / asteroid, cfg.DEVICE='cpu', cfg.window='hann'), runs
the reference DCCRN forward + loss + backward (and Adam steps) on seeded inputs and stores
the results.  /root/reference does not exist on the GPU box; the fixtures travel instead.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SEFD_REFERENCE", "/root/reference")


def import_reference(perceptual=False):
    sys.path.insert(0, REF)
    for n in ["matplotlib", "matplotlib.pylab", "asteroid", "asteroid.losses", "asteroid_filterbanks"]:
        sys.modules[n] = types.ModuleType(n)

    class _Stub:
        def __init__(self, *a, **k):
            pass

        def to(self, *a, **k):
            return self

    sys.modules["asteroid.losses"].SingleSrcPMSQE = _Stub
    sys.modules["asteroid.losses"].PITLossWrapper = _Stub
    sys.modules["asteroid_filterbanks"].STFTFB = _Stub
    sys.modules["asteroid_filterbanks"].Encoder = _Stub
    sys.modules["asteroid_filterbanks"].transforms = None
    with contextlib.redirect_stdout(io.StringIO()):
        import config as cfg
    cfg.DEVICE, cfg.window = "cpu", "hann"
    cfg.loss, cfg.lstm, cfg.skip_type, cfg.perceptual = "SI-SNR", "complex", True, perceptual
    import models
    import tools_for_loss
    return cfg, models, tools_for_loss


def batch(B, L, seed=1234, amp=0.1):
    g = torch.Generator().manual_seed(seed)
    noisy = (torch.rand(B, L, generator=g) * 2 - 1) * amp
    clean = (torch.rand(B, L, generator=g) * 2 - 1) * amp
    return noisy, clean


def speechlike(B, L, seed=7):
    """Target correlated with the input (realistic SI-SNR range, exercises alpha != 0)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(B)])
    noisy = clean + 0.05 * torch.randn(B, L, generator=g)
    return noisy, clean


def main():
    cfg, models, tfl = import_reference()
    torch.set_num_threads(8)
    out = {}

    # ---- 0. init-stream pin: per-key sum / abs-sum of the seed-0 state dict -------------
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C")
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    keys = list(sd0.keys())
    out["init_keys"] = np.array(keys)
    out["init_sum"] = np.array([float(sd0[k].double().sum()) for k in keys])
    out["init_abs"] = np.array([float(sd0[k].double().abs().sum()) for k in keys])
    out["n_params"] = np.array(sum(p.numel() for p in m.parameters()))

    # ---- 1. small case, all mask modes x all losses ------------------------------------
    B, L = 2, 4000
    for inputs_name, (noisy, clean) in {"rand": batch(B, L), "speech": speechlike(B, L)}.items():
        out[f"small_{inputs_name}_noisy"] = noisy.numpy()
        out[f"small_{inputs_name}_clean"] = clean.numpy()
        for mode in ["C", "E", "R"]:
            torch.manual_seed(0)
            m = models.DCCRN(masking_mode=mode).train()
            for loss_name in (["SI-SNR", "SDR", "SI-SDR", "MSE"] if mode == "C" else ["SI-SNR"]):
                cfg.loss = loss_name
                m.load_state_dict(sd0)
                m.zero_grad()
                o_r, o_i, wav = m(noisy, clean)
                loss = m.loss(wav, clean)
                loss.backward()
                tag = f"small_{inputs_name}_{mode}_{loss_name}"
                out[tag + "_loss"] = np.array(loss.item())
                if loss_name == "SI-SNR":
                    out[tag + "_wav"] = wav.detach().numpy()
                    out[tag + "_out_real"] = o_r.detach().numpy()
                    out[tag + "_out_imag"] = o_i.detach().numpy()
                names = [n for n, _ in m.named_parameters()]
                out[tag + "_gnorm"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
                out["param_names"] = np.array(names)
                # full gradients of the small tensors + a strided sample of the big ones
                for n, p in m.named_parameters():
                    g = p.grad.detach().reshape(-1)
                    if loss_name == "SI-SNR" and mode == "C" and inputs_name == "speech":
                        out[tag + "_grad::" + n] = (g if g.numel() <= 4096 else g[:: g.numel() // 2048][:2048]).numpy()
            cfg.loss = "SI-SNR"
        # BN running stats after one train forward (mode C, last loop left them updated once per call)
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C").train()
    m.load_state_dict(sd0)
    noisy, clean = speechlike(B, L)
    m(noisy, clean)
    for k, v in m.state_dict().items():
        if "running" in k:
            out["small_speech_bn::" + k] = v.numpy()

    # eval-mode forward (running stats) for the validation path
    m.eval()
    with torch.no_grad():
        _, _, wav = m(noisy, clean)
    out["small_speech_C_eval_wav"] = wav.numpy()

    # ---- 2. three Adam steps on the small speech case -----------------------------------
    cfg.loss = "SI-SNR"
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C").train()
    m.load_state_dict(sd0)
    opt = torch.optim.Adam(m.parameters(), lr=cfg.learning_rate)
    losses = []
    for _ in range(3):
        _, _, wav = m(noisy, clean)
        loss = m.loss(wav, clean)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    out["adam3_losses"] = np.array(losses)
    out["adam3_param_sum"] = np.array([float(p.double().sum()) for p in m.parameters()])
    out["adam3_param_abs"] = np.array([float(p.double().abs().sum()) for p in m.parameters()])

    # ---- 3. full-length case (SURVEY §4 known answer) -----------------------------------
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C").train()
    noisy, clean = batch(2, 48000)
    _, _, wav = m(noisy, clean)
    loss = m.loss(wav, clean)
    loss.backward()
    out["full_loss"] = np.array(loss.item())
    out["full_wav_head"] = wav.detach()[:, :2048].numpy()
    out["full_wav_rms"] = np.array(float(wav.double().pow(2).mean().sqrt()))
    out["full_wav_sum"] = np.array(float(wav.double().sum()))
    out["full_gnorm_total"] = np.array(float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in m.parameters()))))
    out["full_gnorm"] = np.array([float(p.grad.double().norm()) for p in m.parameters()])

    # ---- 4. loss known answers (tools_for_loss.py:57-74 doctest, numpy seed 0) ---------
    np.random.seed(0)
    ref = np.random.randn(100)
    tr = torch.from_numpy(ref)[None]
    out["si_sdr_doc_inputs"] = ref
    out["si_sdr_doc_expected"] = np.array([-25.127672346460717, 0.481070445785553, 6.3704606032577304, 6.3704606032577304])
    out["si_sdr_ref_values"] = np.array([
        float(tfl.si_sdr(tr, torch.from_numpy(np.flip(ref).copy())[None])),
        float(tfl.si_sdr(tr, tr + torch.from_numpy(np.flip(ref).copy())[None])),
        float(tfl.si_sdr(tr, tr + 0.5)),
        float(tfl.si_sdr(tr, tr * 2 + 1)),
    ])
    a, b = batch(4, 1000, seed=5)
    out["loss_pair_a"], out["loss_pair_b"] = a.numpy(), b.numpy()
    out["loss_ref_si_snr"] = np.array(float(tfl.si_snr(a, b)))
    out["loss_ref_sdr"] = np.array(float(tfl.sdr(a, b)))
    out["loss_ref_si_sdr"] = np.array(float(tfl.si_sdr(a, b)))

    np.savez_compressed(os.path.join(HERE, "dccrn_golden.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "dccrn_golden.npz"))
    print(f"wrote dccrn_golden.npz ({sz/1e3:.1f} kB), {len(out)} arrays; full_loss={out['full_loss']}, "
          f"n_params={out['n_params']}")


def main_crn():
    """BASELINE.json configs[0]: CRN, magnitude T-F mask, MSE loss, batch 2, on the reference's own CPU path."""
    cfg, models, tfl = import_reference()
    torch.set_num_threads(8)
    out = {}
    cfg.loss = "MSE"
    torch.manual_seed(0)
    m = models.CRN(masking_mode="E")
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    keys = list(sd0.keys())
    out["init_keys"] = np.array(keys)
    out["init_sum"] = np.array([float(sd0[k].double().sum()) for k in keys])
    out["init_abs"] = np.array([float(sd0[k].double().abs().sum()) for k in keys])
    out["n_params"] = np.array(sum(p.numel() for p in m.parameters()))
    out["param_names"] = np.array([n for n, _ in m.named_parameters()])
    B, L = 2, 4000
    for inputs_name, (noisy, clean) in {"rand": batch(B, L), "speech": speechlike(B, L)}.items():
        for loss_name in ["MSE", "SI-SNR"]:
            cfg.loss = loss_name
            m.load_state_dict(sd0)
            m.train()
            m.zero_grad()
            est, tgt, wav = m(noisy, clean)
            loss = m.loss(wav, clean)
            loss.backward()
            tag = f"small_{inputs_name}_{loss_name}"
            out[tag + "_loss"] = np.array(loss.item())
            out[tag + "_gnorm"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
            if loss_name == "MSE":
                out[tag + "_wav"] = wav.detach().numpy()
                out[tag + "_est_mags"] = est.detach().numpy()
                out[tag + "_target_mags"] = tgt.detach().numpy()
                for n, p in m.named_parameters():
                    g = p.grad.detach().reshape(-1)
                    out[tag + "_grad::" + n] = (g if g.numel() <= 4096 else g[:: g.numel() // 2048][:2048]).numpy()
    cfg.loss = "MSE"
    noisy, clean = speechlike(B, L)
    m.load_state_dict(sd0)
    m.train()
    m(noisy, clean)
    for k, v in m.state_dict().items():
        if "running" in k:
            out["small_speech_bn::" + k] = v.numpy().copy()   # copy: the buffers are updated in place below
    m.eval()
    with torch.no_grad():
        _, _, wav = m(noisy, clean)
    out["small_speech_eval_wav"] = wav.numpy()
    # three Adam steps
    m.load_state_dict(sd0)
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=cfg.learning_rate)
    losses = []
    for _ in range(3):
        _, _, wav = m(noisy, clean)
        loss = m.loss(wav, clean)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    out["adam3_losses"] = np.array(losses)
    out["adam3_param_sum"] = np.array([float(p.double().sum()) for p in m.parameters()])
    # full-length config-1 case: batch 2, 3 s @ 16 kHz
    m.load_state_dict(sd0)
    m.train()
    m.zero_grad()
    noisy, clean = batch(2, 48000)
    _, _, wav = m(noisy, clean)
    loss = m.loss(wav, clean)
    loss.backward()
    out["full_loss"] = np.array(loss.item())
    out["full_wav_head"] = wav.detach()[:, :2048].numpy()
    out["full_wav_rms"] = np.array(float(wav.double().pow(2).mean().sqrt()))
    out["full_gnorm"] = np.array([float(p.grad.double().norm()) for p in m.parameters()])
    np.savez_compressed(os.path.join(HERE, "crn_golden.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "crn_golden.npz"))
    print(f"wrote crn_golden.npz ({sz/1e3:.1f} kB), {len(out)} arrays; full_loss={out['full_loss']}, n_params={out['n_params']}")


def main_lms():
    """LMS perceptual loss (tools_for_loss.py:111-249) and the DCCRN perceptual train step of
    trainer.model_perceptual_train (trainer.py:44-70): loss = (main + perceptual) / 2."""
    cfg, models, tfl = import_reference(perceptual="LMS")     # MEL_SCALES is fixed at import (tools_for_loss.py:117-120)
    torch.set_num_threads(8)
    out = {}
    # op level: two magnitude arrays, gradient with respect to the estimate
    g = torch.Generator().manual_seed(11)
    clean = torch.rand(2, 257, 43, generator=g) * 3 + 0.01
    est = (clean * (0.5 + torch.rand(2, 257, 43, generator=g))).requires_grad_(True)
    loss = tfl.get_array_lms_loss(clean, est)
    loss.backward()
    out["op_clean"], out["op_est"] = clean.numpy(), est.detach().numpy()
    out["op_loss"] = np.array(loss.item())
    out["op_grad"] = est.grad.numpy().copy()
    # model level
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C").train()
    noisy, clean_w = speechlike(2, 4000)
    real, imag, wav = m(noisy, clean_w)
    main = m.loss(wav, clean_w)
    perc = m.loss(wav, clean_w, real, imag, perceptual=True)
    total = (main + perc) / 2
    total.backward()
    out["model_main"], out["model_perc"], out["model_total"] = np.array(main.item()), np.array(perc.item()), np.array(total.item())
    out["param_names"] = np.array([n for n, _ in m.named_parameters()])
    out["model_gnorm"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
    for n, p in m.named_parameters():
        gr = p.grad.detach().reshape(-1)
        out["model_grad::" + n] = (gr if gr.numel() <= 4096 else gr[:: gr.numel() // 2048][:2048]).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "lms_golden.npz"), **out)
    print(f"wrote lms_golden.npz, op_loss={out['op_loss']}, model main/perc/total = {out['model_main']}, {out['model_perc']}, {out['model_total']}")


def main_direct():
    """masking_mode 'Direct(None make)' (models.py:232-250) with the loop of trainer.dccrn_direct_train (trainer.py:122-150):
    loss = (loss(out_real, target_real) + loss(out_imag, target_imag)) / 2."""
    cfg, models, tfl = import_reference()
    torch.set_num_threads(8)
    out = {}
    noisy, clean = speechlike(2, 4000)
    for loss_name in ["MSE", "SI-SNR"]:
        cfg.loss = loss_name
        torch.manual_seed(0)
        m = models.DCCRN(masking_mode="Direct(None make)").train()
        o_r, t_r, o_i, t_i, wav = m(noisy, clean)
        loss = (m.loss(o_r, t_r) + m.loss(o_i, t_i)) / 2
        loss.backward()
        out[loss_name + "_loss"] = np.array(loss.item())
        out[loss_name + "_gnorm"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
        if loss_name == "MSE":
            out["wav"], out["out_real"], out["target_real"] = wav.detach().numpy(), o_r.detach().numpy(), t_r.detach().numpy()
            out["param_names"] = np.array([n for n, _ in m.named_parameters()])
    cfg.loss = "SI-SNR"
    np.savez_compressed(os.path.join(HERE, "direct_golden.npz"), **out)
    print("wrote direct_golden.npz", out["MSE_loss"], out["SI-SNR_loss"])


def main_noskip():
    """cfg.skip_type = False: decoder without skip connections (models.py:138-169, 227-230).  cfg.skip_type is read when the
    model is constructed and again in forward, so it is set before either."""
    cfg, models, tfl = import_reference()
    cfg.skip_type = False
    torch.set_num_threads(8)
    out = {}
    noisy, clean = speechlike(2, 4000)
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C").train()
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    o_r, o_i, wav = m(noisy, clean)
    loss = m.loss(wav, clean)
    loss.backward()
    out["loss"] = np.array(loss.item())
    out["wav"], out["out_real"] = wav.detach().numpy(), o_r.detach().numpy()
    names = [n for n, _ in m.named_parameters()]
    out["param_names"] = np.array(names)
    out["param_shapes"] = np.array([str(tuple(p.shape)) for _, p in m.named_parameters()])
    out["gnorm"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
    for n, p in m.named_parameters():                      # small tensors whole, large ones as a strided sample (stride 97)
        gflat = p.grad.reshape(-1)
        out["grad:" + n] = (gflat if gflat.numel() <= 4096 else gflat[::97]).numpy().copy()
    for k in ("decoder.0.0.real_conv.weight", "decoder.5.0.imag_conv.weight", "encoder.0.0.real_conv.weight"):
        out["init:" + k] = sd0[k].reshape(-1)[::97].numpy().copy()           # pins the RNG stream of the construction order
    m.eval()
    with torch.no_grad():
        out["eval_wav"] = m(noisy)[2].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "noskip_golden.npz"), **out)
    print("wrote noskip_golden.npz", out["loss"], len(names))


def main_fullsubnet():
    """BASELINE.json configs[2] at a small size: FullSubNet, loop body of trainer.fullsubnet_train (trainer.py:97-107), MSE
    between cIRM and cRM.  The model is put in eval() so that the inter-layer LSTM dropout (0.8) is inactive (SURVEY.md 8(d)
    config 3: eval-mode dropout for parity); autograd still runs."""
    cfg, models, tfl = import_reference()
    import tools_for_model as tools
    torch.set_num_threads(8)
    cfg.loss = "MSE"
    out = {}
    torch.manual_seed(0)
    m = models.FullSubNet()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    keys = list(sd0.keys())
    out["init_keys"] = np.array(keys)
    out["init_shapes"] = np.array([str(tuple(sd0[k].shape)) for k in keys])
    out["init_sum"] = np.array([float(sd0[k].double().sum()) for k in keys])
    out["init_abs"] = np.array([float(sd0[k].double().abs().sum()) for k in keys])
    out["n_params"] = np.array(sum(p.numel() for p in m.parameters()))
    m.eval()
    noisy, clean = speechlike(2, 4000)
    nc, cc = tools.stft(noisy), tools.stft(clean)
    noisy_mag, _ = tools.mag_phase(nc)
    cirm = tools.build_complex_ideal_ratio_mask(nc, cc)
    crm = m(noisy_mag)
    loss = m.loss(cirm, crm)
    loss.backward()
    out["noisy_mag"], out["cIRM"], out["cRM"] = noisy_mag.numpy(), cirm.numpy(), crm.detach().numpy()
    out["loss"] = np.array(loss.item())
    out["param_names"] = np.array([n for n, _ in m.named_parameters()])
    out["gnorm"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
    for n, p in m.named_parameters():
        g = p.grad.detach().reshape(-1)
        out["grad::" + n] = (g if g.numel() <= 4096 else g[::997]).numpy().copy()
    out["decompressed"] = tools.decompress_cIRM(crm.detach()).numpy()          # trainer.py:341: the validation path's inverse
    np.savez_compressed(os.path.join(HERE, "fullsubnet_golden.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "fullsubnet_golden.npz"))
    print(f"wrote fullsubnet_golden.npz ({sz/1e3:.1f} kB); loss={out['loss']}, n_params={out['n_params']}, T={crm.shape[2]}")


def main_reallstm():
    """cfg.lstm = 'real' (models.py:96-105, 213-218): one 2-layer nn.LSTM(1024 -> 256) + Linear(256 -> 1024) instead of
    the complex LSTM pair.  Oracle fixture only: the CUDA path of this variant is not built yet."""
    cfg, models, tfl = import_reference()
    cfg.lstm = "real"
    torch.set_num_threads(8)
    out = {}
    noisy, clean = speechlike(2, 4000)
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C").train()
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    o_r, o_i, wav = m(noisy, clean)
    loss = m.loss(wav, clean)
    loss.backward()
    out["loss"] = np.array(loss.item())
    out["wav"], out["out_real"] = wav.detach().numpy(), o_r.detach().numpy()
    out["param_names"] = np.array([n for n, _ in m.named_parameters()])
    out["param_shapes"] = np.array([str(tuple(p.shape)) for _, p in m.named_parameters()])
    out["gnorm"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
    for n, p in m.named_parameters():
        gflat = p.grad.reshape(-1)
        out["grad:" + n] = (gflat if gflat.numel() <= 4096 else gflat[::997]).numpy().copy()
    for k in ("enhance.weight_hh_l1", "tranform.weight", "decoder.0.0.real_conv.weight"):
        out["init:" + k] = sd0[k].reshape(-1)[::97].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "reallstm_golden.npz"), **out)
    print("wrote reallstm_golden.npz", out["loss"], len(out["param_names"]))


def main_cbn():
    """use_cbn = True (models.py:26, 76, 120, 151; ComplexBatchNorm tools_for_model.py:430-603).  The reference module calls
    torch.addcmul with the positional `value` argument of torch < 1.5 (line 567), which current torch rejects, so it cannot run
    as is; the ONLY change made here is a shim that accepts that legacy call form (addcmul(t, value, a, b) = t + value * a * b) -
    the reference files themselves are imported unmodified."""
    cfg, models, tfl = import_reference()
    torch.set_num_threads(8)
    _addcmul = torch.addcmul

    def addcmul_legacy(inp, *args, **kw):
        if len(args) == 3 and not torch.is_tensor(args[0]):
            return _addcmul(inp, args[1], args[2], value=args[0])
        return _addcmul(inp, *args, **kw)

    torch.addcmul = addcmul_legacy
    out = {}
    noisy, clean = speechlike(2, 4000)
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C", use_cbn=True).train()
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    out["state_keys"] = np.array(list(sd0.keys()))
    o_r, o_i, wav = m(noisy, clean)
    loss = m.loss(wav, clean)
    loss.backward()
    out["loss"] = np.array(loss.item())
    out["wav"], out["out_real"] = wav.detach().numpy(), o_r.detach().numpy()
    out["param_names"] = np.array([n for n, _ in m.named_parameters()])
    out["param_shapes"] = np.array([str(tuple(p.shape)) for _, p in m.named_parameters()])
    out["gnorm"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
    for n, p in m.named_parameters():
        gflat = p.grad.reshape(-1)
        out["grad:" + n] = (gflat if gflat.numel() <= 4096 else gflat[::997]).numpy().copy()
    for k in ("encoder.0.1.Wri", "encoder.3.1.Wri", "decoder.2.1.Wri", "decoder.0.0.real_conv.weight", "enhance.1.imag_lstm.weight_hh_l0"):
        out["init:" + k] = sd0[k].reshape(-1)[::7].numpy().copy()
    for k, v in m.state_dict().items():                      # running buffers after one train-mode forward
        if k.split(".")[-1] in ("RMr", "RMi", "RVrr", "RVri", "RVii", "num_batches_tracked"):
            out["buf:" + k] = v.detach().numpy().copy()
    m.eval()
    with torch.no_grad():
        _, _, wav_e = m(noisy, clean)
    out["wav_eval"] = wav_e.numpy()
    torch.addcmul = _addcmul
    np.savez_compressed(os.path.join(HERE, "cbn_golden.npz"), **out)
    print("wrote cbn_golden.npz", out["loss"], len(out["param_names"]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cbn":
        main_cbn()
    elif len(sys.argv) > 1 and sys.argv[1] == "reallstm":
        main_reallstm()
    elif len(sys.argv) > 1 and sys.argv[1] == "fullsubnet":
        main_fullsubnet()
    elif len(sys.argv) > 1 and sys.argv[1] == "noskip":
        main_noskip()
    elif len(sys.argv) > 1 and sys.argv[1] == "direct":
        main_direct()
    elif len(sys.argv) > 1 and sys.argv[1] == "lms":
        main_lms()
    elif len(sys.argv) > 1 and sys.argv[1] == "crn":
        main_crn()
    else:
        main()
