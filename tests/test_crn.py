"""CRN path (BASELINE.json configs[0]: magnitude T-F mask, MSE loss, batch 2).

CPU part: the oracle (oracle/crn_oracle.py) against fixtures produced by the UNMODIFIED reference CRN
(tests/golden/make_golden.py crn).  GPU part (-m gpu): the drop-in models.CRN (C ABI, CUDA kernels) against the oracle
and the same fixtures, both GEMM engines.
"""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import crn_oracle as O
from oracle import dccrn_oracle as D

DEV = "cuda"
REPORT = os.path.join(ROOT, "gpurun_out", "crn_report.txt")


@pytest.fixture(scope="module")
def crn_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "crn_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="module")
def sd0():
    return O.init_state(0)


def _inputs(name, B=2, L=4000):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    if name == "rand":
        return D.synthetic_batch(B, L)
    g = torch.Generator().manual_seed(7)                      # make_golden.speechlike
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(B)])
    return clean + 0.05 * torch.randn(B, L, generator=g), clean


# ------------------------------------------------------------------------------------------------
# CPU: oracle pinned to the reference
# ------------------------------------------------------------------------------------------------
def test_crn_init_stream_matches_reference(crn_golden, sd0):
    keys = [str(k) for k in crn_golden["init_keys"]]
    assert keys == list(sd0.keys())
    for k, s, a in zip(keys, crn_golden["init_sum"], crn_golden["init_abs"]):
        v = sd0[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), k
    assert sum(sd0[k].numel() for k in O.trainable_keys(sd0)) == int(crn_golden["n_params"]) == 1703436
    assert [str(n) for n in crn_golden["param_names"]] == O.trainable_keys(sd0)


@pytest.mark.parametrize("inputs", ["rand", "speech"])
@pytest.mark.parametrize("loss_name", ["MSE", "SI-SNR"])
def test_crn_oracle_forward_backward(crn_golden, sd0, inputs, loss_name):
    noisy, clean = _inputs(inputs)
    tr = O.OracleTrainer(sd0, loss=loss_name)
    taps = {}
    loss, wav = tr.forward_backward(noisy, clean, taps)
    tag = f"small_{inputs}_{loss_name}"
    assert float(loss) == pytest.approx(float(crn_golden[tag + "_loss"]), rel=2e-5, abs=1e-9)
    gn = np.array([float(tr.sd[k].grad.double().norm()) for k in tr.keys])
    ref = crn_golden[tag + "_gnorm"]
    np.testing.assert_allclose(gn, ref, rtol=2e-3, atol=2e-4 * ref.max())
    if loss_name == "MSE":
        np.testing.assert_allclose(wav.numpy(), crn_golden[tag + "_wav"], atol=2e-6)
        np.testing.assert_allclose(taps["est_mags"].numpy(), crn_golden[tag + "_est_mags"], atol=2e-4, rtol=1e-4)
        np.testing.assert_allclose(taps["target_mags"].numpy(), crn_golden[tag + "_target_mags"], atol=2e-4, rtol=1e-4)
        gmax = max(float(tr.sd[k].grad.abs().max()) for k in tr.keys)
        for k in tr.keys:
            g = tr.sd[k].grad.reshape(-1)
            g = g if g.numel() <= 4096 else g[:: g.numel() // 2048][:2048]
            r = crn_golden[tag + "_grad::" + k]
            # conv biases in front of a BatchNorm have an analytically zero gradient (pure rounding noise): floor
            assert np.abs(g.numpy() - r).max() <= 2e-3 * np.abs(r).max() + 1e-6 * gmax, k


def test_crn_oracle_bn_eval_and_adam(crn_golden, sd0):
    noisy, clean = _inputs("speech")
    tr = O.OracleTrainer(sd0, loss="MSE")
    tr.forward_backward(noisy, clean)
    for k in tr.sd:
        if "running" in k:
            np.testing.assert_allclose(tr.sd[k].numpy(), crn_golden["small_speech_bn::" + k], rtol=1e-4, atol=1e-6)
    with torch.no_grad():
        _, _, wav = O.crn_forward({k: v.detach() for k, v in tr.sd.items()}, noisy, clean, train=False)
    np.testing.assert_allclose(wav.numpy(), crn_golden["small_speech_eval_wav"], atol=5e-6)
    tr = O.OracleTrainer(sd0, loss="MSE")
    losses = [float(tr.step(noisy, clean)[0]) for _ in range(3)]
    np.testing.assert_allclose(losses, crn_golden["adam3_losses"], rtol=2e-3)


def test_crn_oracle_full_length(crn_golden, sd0):
    noisy, clean = D.synthetic_batch(2, 48000)
    tr = O.OracleTrainer(sd0, loss="MSE")
    loss, wav = tr.forward_backward(noisy, clean)
    assert float(loss) == pytest.approx(float(crn_golden["full_loss"]), rel=2e-5)
    np.testing.assert_allclose(wav[:, :2048].numpy(), crn_golden["full_wav_head"], atol=2e-6)


# ------------------------------------------------------------------------------------------------
# GPU: CUDA path vs oracle / golden
# ------------------------------------------------------------------------------------------------
def _report(line):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(line + "\n")


def _build(sd0):
    import models
    m = models.CRN(masking_mode="E")
    m.load_state_dict(sd0)
    return m.to(DEV).train()


def _cl(x):
    return x.permute(0, 2, 3, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("inputs", ["rand", "speech"])
def test_crn_gpu_forward_backward_every_tensor(crn_golden, sd0, engine, inputs):
    import models
    tf = engine == 1
    noisy, clean = _inputs(inputs)
    tr = O.OracleTrainer(sd0, loss="MSE")
    taps = {}
    loss_ref, wav_ref = tr.forward_backward(noisy, clean, taps)
    m = _build(sd0)
    models.cfg.loss = "MSE"
    est, tgt, wav = m(noisy.to(DEV), clean.to(DEV))
    loss = m.loss(wav, clean.to(DEV))
    loss.backward()
    plan = m._get_engine().plan(*noisy.shape)
    failures = []

    def chk(name, got, ref, atol_rel, rtol=1e-4):
        if tf:
            atol_rel, rtol = 2e-2, 1e-2
        g64, r64 = got.detach().double().cpu(), ref.detach().double().cpu()
        assert g64.shape == r64.shape, (name, g64.shape, r64.shape)
        s = float(r64.abs().max())
        e = float((g64 - r64).abs().max())
        ok = bool(((g64 - r64).abs() <= atol_rel * s + rtol * r64.abs()).all())
        _report(f"[crn {inputs}{'/tf32' if tf else ''}] {name:28s} max|err|={e:.3e} max|ref|={s:.3e} {'ok' if ok else 'FAIL'}")
        if not ok:
            failures.append(name)

    for i in range(6):
        chk(f"enc{i}.y", plan.tensor(f"enc{i}.y"), _cl(taps[f"enc{i}_conv"]), 4e-6)
        chk(f"enc{i}.z", plan.tensor(f"enc{i}.z"), _cl(taps[f"enc{i}"]), 4e-6)
    chk("lstm h", plan.tensor("H")[0], taps["lstm"].permute(1, 0, 2), 5e-6)
    T, B = taps["proj"].shape[:2]
    chk("tranform", plan.tensor("U"), taps["proj"].reshape(T, B, 128, 4).permute(1, 3, 0, 2), 5e-6)
    for j in range(6):
        chk(f"dec{j}.y", plan.tensor(f"dec{j}.y"), _cl(taps[f"dec{j}_conv"]), 5e-6)
        if j < 5:
            chk(f"dec{j}.z", plan.tensor(f"dec{j}.z"), _cl(taps[f"dec{j}"]), 5e-6)
    chk("out_wav", wav, wav_ref, 2e-5)
    chk("est_mags", est, taps["est_mags"], 2e-5)
    chk("target_mags", tgt, taps["target_mags"], 2e-5)
    chk("wav vs golden", wav, torch.from_numpy(crn_golden[f"small_{inputs}_MSE_wav"]), 2e-5)
    if abs(float(loss) - float(loss_ref)) > (5e-3 if tf else 2e-4) * abs(float(loss_ref)) + 1e-9:
        failures.append("loss")
    assert float(loss) == pytest.approx(float(crn_golden[f"small_{inputs}_MSE_loss"]), rel=5e-3 if tf else 2e-4)

    grads = tr.grads()
    gmax = max(float(g.abs().max()) for g in grads.values())
    for name, p in m.named_parameters():
        ref = grads[name]
        g64, r64 = p.grad.detach().double().cpu(), ref.detach().double()
        e, s = float((g64 - r64).abs().max()), float(r64.abs().max())
        if name.endswith("conv.bias") and not name.startswith("decoder.5."):
            ok = float(g64.abs().max()) <= 1e-4 * gmax          # exactly zero in front of a BatchNorm
        else:
            cosv = float((g64.reshape(-1) * r64.reshape(-1)).sum() / (g64.norm() * r64.norm() + 1e-30))
            nr = float(g64.norm() / (r64.norm() + 1e-30))
            if tf:
                ok = (cosv > 0.99 and abs(nr - 1) < 0.05) or e <= 2e-2 * gmax
            else:
                # element-wise bound; a tensor may instead differ by ONE PReLU branch decision (an activation within an
                # ulp of 0 whose sign differs between the CPU and the GPU arithmetic - DESIGN.md "conditioning"): that
                # moves the upstream gradients by <~1 % but leaves direction and norm intact
                ok = e <= (2e-2 if name.endswith(".2.weight") else 2e-3) * s + 1e-6 * gmax or \
                    (cosv > 0.9995 and abs(nr - 1) < 0.02)
        _report(f"[crn {inputs}{'/tf32' if tf else ''}] grad {name:32s} max|err|={e:.3e} max|ref|={s:.3e} {'ok' if ok else 'FAIL'}")
        if not ok:
            failures.append("grad " + name)
    assert not failures, failures


@pytest.mark.gpu
def test_crn_gpu_eval_adam_and_full_length(crn_golden, sd0, engine):
    import models
    from sefd.train import TrainStep
    tf = engine == 1
    models.cfg.loss = "MSE"
    noisy, clean = _inputs("speech")
    m = _build(sd0)
    m(noisy.to(DEV), clean.to(DEV))
    sd = m.state_dict()
    for k in sd:
        if "running" in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), crn_golden["small_speech_bn::" + k], rtol=2e-2 if tf else 1e-4,
                                       atol=1e-3 if tf else 1e-6)
    m.eval()
    with torch.no_grad():
        _, _, wav = m(noisy.to(DEV), clean.to(DEV))
    np.testing.assert_allclose(wav.cpu().numpy(), crn_golden["small_speech_eval_wav"], atol=3e-3 if tf else 2e-5)
    # three fused train steps (forward + MSE + backward + Adam) through the C ABI
    m = _build(sd0)
    ts = TrainStep(m, lr=1e-3, loss="MSE")
    losses = [float(ts.step(noisy.to(DEV), clean.to(DEV))) for _ in range(3)]
    np.testing.assert_allclose(losses, crn_golden["adam3_losses"], rtol=2e-2 if tf else 2e-3)
    # BASELINE configs[0]: batch 2, 3 s @ 16 kHz
    m = _build(sd0)
    noisy, clean = D.synthetic_batch(2, 48000)
    _, _, wav = m(noisy.to(DEV), clean.to(DEV))
    loss = m.loss(wav, clean.to(DEV))
    loss.backward()
    assert float(loss) == pytest.approx(float(crn_golden["full_loss"]), rel=5e-3 if tf else 2e-5)
    rmse = float((wav[:, :2048].cpu() - torch.from_numpy(crn_golden["full_wav_head"])).pow(2).mean().sqrt())
    _report(f"[crn full engine={engine}] loss {float(loss):.8f} wav RMSE vs reference {rmse:.3e}")
    assert rmse < 1e-4
    names = [n for n, _ in m.named_parameters()]
    gn = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
    ref = crn_golden["full_gnorm"]
    for i, n in enumerate(names):
        if n.endswith("conv.bias") and not n.startswith("decoder.5."):
            continue
        assert abs(gn[i] - ref[i]) <= (5e-2 if tf else 5e-3) * ref[i] + 1e-3 * ref.max(), (n, gn[i], ref[i])


@pytest.mark.gpu
def test_crn_gpu_gradient_through_est_mags(sd0):
    """A loss on the magnitude output (trainer.crn_direct_train: model.loss(output_mag, target_mag), trainer.py:168-169;
    CRN.loss perceptual branch models.py:553-555) sends a gradient to est_mags; checked against the oracle's autograd.
    (LMS itself is not used here: tanh(out) * |X| can be negative, so log-mel of est_mags is NaN in the reference too.)
    fp32 engine; (waveform MSE + magnitude MSE) / 2."""
    import models
    from sefd import _lib
    lib = _lib.load()
    lib.sefd_set_engine(0)
    models.cfg.loss = "MSE"
    try:
        noisy, clean = _inputs("speech")
        tr = O.OracleTrainer(sd0, loss="MSE")
        for k in tr.keys:
            tr.sd[k].grad = None
        est_r, tgt_r, wav_r = O.crn_forward(tr.sd, noisy, clean, train=True, taps={})
        ref_total = (O.crn_loss(wav_r, clean, "MSE") + O.crn_loss(est_r, tgt_r, "MSE")) / 2
        ref_total.backward()
        m = _build(sd0)
        est, tgt, wav = m(noisy.to(DEV), clean.to(DEV))
        total = (m.loss(wav, clean.to(DEV)) + m.loss(est, tgt)) / 2
        total.backward()
        assert float(total) == pytest.approx(float(ref_total), rel=2e-4)
        grads = tr.grads()
        gmax = max(float(g.abs().max()) for g in grads.values())
        for name, p in m.named_parameters():
            if name.endswith("conv.bias") and not name.startswith("decoder.5."):
                continue
            g64, r64 = p.grad.detach().double().cpu(), grads[name].detach().double()
            e, s = float((g64 - r64).abs().max()), float(r64.abs().max())
            cosv = float((g64.reshape(-1) * r64.reshape(-1)).sum() / (g64.norm() * r64.norm() + 1e-30))
            assert e <= 5e-3 * s + 1e-6 * gmax or cosv > 0.9995, (name, e, s, cosv)
    finally:
        models.cfg.loss = "SI-SNR"
        lib.sefd_set_engine(1)
