"""use_cbn = True (models.py:26, 76, 120, 151): ComplexBatchNorm (tools_for_model.py:430-603) in place of BatchNorm2d.
The reference module cannot run on current torch (legacy positional `value` of torch.addcmul, line 567); the fixture
tests/golden/cbn_golden.npz comes from the unmodified reference files behind a shim for exactly that call form
(tests/golden/make_golden.py cbn).  Oracle restatement against the fixture on CPU, the CUDA path against oracle and fixture."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import dccrn_oracle as O

BUFS = ("RMr", "RMi", "RVrr", "RVri", "RVii")


def _speech(B=2, L=4000):
    g = torch.Generator().manual_seed(7)
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(B)])
    return clean + 0.05 * torch.randn(B, L, generator=g), clean


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "cbn_golden.npz"), allow_pickle=False)


def test_cbn_oracle(gold):
    sd0 = O.init_state(0, use_cbn=True)
    assert sorted(str(k) for k in gold["state_keys"]) == sorted(sd0.keys())                     # the reference's state_dict keys
    for k in [str(n)[5:] for n in gold.files if n.startswith("init:")]:
        np.testing.assert_array_equal(sd0[k].reshape(-1)[::7].numpy(), gold["init:" + k])      # same RNG stream
    noisy, clean = _speech()
    tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
    loss, wav = tr.forward_backward(noisy, clean)
    assert float(loss) == pytest.approx(float(gold["loss"]), rel=2e-5)
    np.testing.assert_allclose(wav.numpy(), gold["wav"], atol=2e-6)
    names = [str(n) for n in gold["param_names"]]
    assert sorted(names) == sorted(tr.keys)
    assert [str(tuple(tr.sd[k].shape)) for k in names] == [str(s) for s in gold["param_shapes"]]
    ref = gold["gnorm"]
    gn = np.array([float(tr.sd[k].grad.double().norm()) for k in names])
    np.testing.assert_allclose(gn, ref, rtol=2e-3, atol=2e-4 * ref.max())
    for k in names:
        if k.endswith("_conv.bias") and not k.startswith("decoder.5"):
            continue                          # zero by the mean removal of the normalisation (rounding noise only)
        g = tr.sd[k].grad.reshape(-1)
        g = g if g.numel() <= 4096 else g[::997]
        r = gold["grad:" + k]
        np.testing.assert_allclose(g.numpy(), r, atol=2e-3 * max(float(np.abs(r).max()), 1e-3), err_msg=k)
    for n in gold.files:                                    # running buffers after the train-mode forward
        if n.startswith("buf:"):
            np.testing.assert_allclose(tr.sd[n[4:]].numpy(), gold[n], rtol=1e-5, atol=1e-7, err_msg=n)
    with torch.no_grad():                                   # eval mode reads the running buffers
        _, _, wav_e = O.dccrn_forward(tr.sd, noisy, "C", train=False)
    np.testing.assert_allclose(wav_e.numpy(), gold["wav_eval"], atol=5e-6)


def test_dropin_layout_for_cbn():
    import models
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C", use_cbn=True)
    ref = O.init_state(0, use_cbn=True)
    sd = m.state_dict()
    assert set(sd.keys()) == set(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape, k
        if sd[k].is_floating_point() and not k.startswith(("stft.", "istft.")):
            assert torch.equal(sd[k], ref[k]), k            # same RNG stream as the reference constructor (Wri ~ U(-0.9, 0.9))
    assert [n for n, _ in m.named_parameters() if n.startswith("encoder.0.")] == [
        "encoder.0.0.real_conv.weight", "encoder.0.0.real_conv.bias", "encoder.0.0.imag_conv.weight", "encoder.0.0.imag_conv.bias",
        "encoder.0.1.Wrr", "encoder.0.1.Wri", "encoder.0.1.Wii", "encoder.0.1.Br", "encoder.0.1.Bi", "encoder.0.2.weight"]


@pytest.mark.gpu
def test_cbn_gpu(engine, gold):
    """Drop-in DCCRN(use_cbn=True) through the C ABI (plan flag SEFD_PLAN_CBN, csrc/cbn.cu): waveform, loss, every gradient and the
    running buffers against the oracle and the reference's own fixture values; then the eval-mode forward."""
    import models
    models.cfg.loss = "SI-SNR"
    sd0 = O.init_state(0, use_cbn=True)
    noisy, clean = _speech()
    tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
    loss_ref, wav_ref = tr.forward_backward(noisy, clean)
    m = models.DCCRN(masking_mode="C", use_cbn=True)
    m.load_state_dict(sd0)
    m = m.cuda().train()
    _, _, wav = m(noisy.cuda(), clean.cuda())
    loss = m.loss(wav, clean.cuda())
    loss.backward()
    torch.cuda.synchronize()
    tf = engine == 1
    rmse = float((wav.detach().cpu() - wav_ref).pow(2).mean().sqrt())
    assert rmse < (1e-4 if tf else 2e-6), rmse
    np.testing.assert_allclose(wav.detach().cpu().numpy(), gold["wav"], atol=2e-3 if tf else 2e-5)
    assert float(loss.detach()) == pytest.approx(float(gold["loss"]), rel=5e-3 if tf else 2e-4)
    grads = tr.grads()
    amax = max(float(v.abs().max()) for n, v in grads.items() if n.endswith(".2.weight"))
    for k, p in m.named_parameters():
        if k.endswith("_conv.bias") and not k.startswith("decoder.5"):
            continue                      # zero by the mean removal
        g, r = p.grad.detach().cpu().double().reshape(-1), grads[k].double().reshape(-1)
        if k.endswith(".2.weight"):       # the single PReLU slope: one cancelling global sum
            assert abs(float(g[0] - r[0])) <= 2e-2 * amax, k
            continue
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
        nr = float(g.norm() / (r.norm() + 1e-30))
        assert cos > (0.99 if tf else 0.9995) and abs(nr - 1) < (0.05 if tf else 0.02), (k, cos, nr)
    sd = m.state_dict()
    for k in sd:                                            # running buffers after one train-mode forward
        if k.split(".")[-1] in BUFS:
            ref = tr.sd[k].numpy()                          # means / covariances of TF32 conv outputs: error relative to the buffer's scale
            tol = dict(rtol=5e-3, atol=2e-3 * float(np.abs(ref).max())) if tf else dict(rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(sd[k].cpu().numpy(), ref, err_msg=k, **tol)
            np.testing.assert_allclose(sd[k].cpu().numpy(), gold["buf:" + k], err_msg=k, **tol)
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == 1
    m.eval()
    with torch.no_grad():
        _, _, wav_e = m(noisy.cuda(), clean.cuda())
        _, _, wav_eo = O.dccrn_forward(tr.sd, noisy, "C", train=False)
    assert float((wav_e.cpu() - wav_eo).pow(2).mean().sqrt()) < (2e-4 if tf else 5e-6)


@pytest.mark.gpu
def test_standalone_complex_batch_norm_layer():
    """tools_for_model.ComplexBatchNorm (the bare layer, op-level sefd_cbn_prelu_forward / backward with slope 1) against the
    oracle's complex_batch_norm and its autograd: output, input / parameter gradients, running buffers, eval mode."""
    import tools_for_model as T
    torch.manual_seed(3)
    layer = T.ComplexBatchNorm(32).cuda().train()
    with torch.no_grad():
        layer.Wrr.uniform_(0.5, 1.5); layer.Wii.uniform_(0.5, 1.5); layer.Br.normal_(); layer.Bi.normal_()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 32, 6, 9, generator=g)
    x[:, 16:] = 0.5 * x[:, :16] + 0.7 * x[:, 16:]                 # correlated real / imaginary parts: Vri != 0
    dz = torch.randn(2, 32, 6, 9, generator=g)
    ref_p = [getattr(layer, k).detach().cpu().double().requires_grad_(True) for k in ("Wrr", "Wri", "Wii", "Br", "Bi")]
    xr = x.double().requires_grad_(True)
    y_ref, st = O.complex_batch_norm(xr, *ref_p)
    (y_ref * dz.double()).sum().backward()
    xg = x.cuda().requires_grad_(True)
    y = layer(xg)
    (y * dz.cuda()).sum().backward()
    assert float((y.detach().cpu().double() - y_ref.detach()).abs().max()) < 2e-5
    assert float((xg.grad.cpu().double() - xr.grad).abs().max()) < 2e-5 * float(xr.grad.abs().max()) + 1e-6
    for k, r in zip(("Wrr", "Wri", "Wii", "Br", "Bi"), ref_p):
        got = getattr(layer, k).grad.cpu().double()
        assert float((got - r.grad).abs().max()) < 1e-4 * float(r.grad.abs().max()) + 1e-5, k
    for k, v, init in zip(BUFS, st, (0.0, 0.0, 1.0, 0.0, 1.0)):
        want = init + 0.1 * (v.detach() - init)
        assert float((getattr(layer, k).cpu().double() - want).abs().max()) < 1e-5, k
    assert int(layer.num_batches_tracked) == 1
    layer.eval()
    with torch.no_grad():
        ye = layer(x.cuda())
        ye_ref, _ = O.complex_batch_norm(x.double(), *[p.detach() for p in ref_p],
                                         stats=[getattr(layer, k).cpu().double() for k in BUFS])
    assert float((ye.cpu().double() - ye_ref).abs().max()) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("skip,lstm", [(False, "complex"), (True, "real"), (False, "real")])
def test_plan_flag_combinations(skip, lstm):
    """SEFD_PLAN_CBN together with SEFD_PLAN_NO_SKIP / SEFD_PLAN_REAL_LSTM (the flags are independent in the reference too: use_cbn,
    cfg.skip_type, cfg.lstm): waveform, loss and the flat gradient of the drop-in against the oracle, default engine."""
    import models
    old = (models.cfg.skip_type, models.cfg.lstm, models.cfg.loss)
    models.cfg.skip_type, models.cfg.lstm, models.cfg.loss = skip, lstm, "SI-SNR"
    try:
        sd0 = O.init_state(0, skip_type=skip, lstm=lstm, use_cbn=True)
        noisy, clean = _speech()
        tr = O.OracleTrainer(sd0, masking_mode="E", loss="SI-SNR")
        loss_ref, wav_ref = tr.forward_backward(noisy, clean)
        m = models.DCCRN(masking_mode="E", use_cbn=True)
        m.load_state_dict(sd0)
        m = m.cuda().train()
        _, _, wav = m(noisy.cuda(), clean.cuda())
        loss = m.loss(wav, clean.cuda())
        loss.backward()
        torch.cuda.synchronize()
        assert float((wav.detach().cpu() - wav_ref).pow(2).mean().sqrt()) < 1e-4
        assert float(loss.detach()) == pytest.approx(float(loss_ref), rel=5e-3, abs=5e-3)      # SI-SNR in dB, may sit near 0
        grads = tr.grads()
        keys = [k for k, _ in m.named_parameters() if not k.endswith(".2.weight") and not (k.endswith("_conv.bias") and not k.startswith("decoder.5"))]
        got = torch.cat([dict(m.named_parameters())[k].grad.detach().cpu().double().reshape(-1) for k in keys])
        ref = torch.cat([grads[k].double().reshape(-1) for k in keys])
        cos = float((got * ref).sum() / (got.norm() * ref.norm()))
        assert cos > 0.999 and abs(float(got.norm() / ref.norm()) - 1) < 0.02, (cos, float(got.norm() / ref.norm()))
    finally:
        models.cfg.skip_type, models.cfg.lstm, models.cfg.loss = old
