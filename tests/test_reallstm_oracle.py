"""cfg.lstm = 'real' (models.py:96-105, 213-218): oracle restatement against fixtures of the unmodified reference
(tests/golden/make_golden.py reallstm), and the CUDA path (time-major LSTM layer engine of lstm_seq.cu / lstm_step_tc.cu behind
the DCCRN plan flag SEFD_PLAN_REAL_LSTM) against the oracle and the same fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import dccrn_oracle as O


def _speech(B=2, L=4000):
    g = torch.Generator().manual_seed(7)
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(B)])
    return clean + 0.05 * torch.randn(B, L, generator=g), clean


def test_real_lstm_oracle():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reallstm_golden.npz"), allow_pickle=False)
    sd0 = O.init_state(0, lstm="real")
    for k in ("enhance.weight_hh_l1", "tranform.weight", "decoder.0.0.real_conv.weight"):
        np.testing.assert_array_equal(sd0[k].reshape(-1)[::97].numpy(), gold["init:" + k])      # same RNG stream
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
            continue                          # zero by BatchNorm (rounding noise only)
        g = tr.sd[k].grad.reshape(-1)
        g = g if g.numel() <= 4096 else g[::997]
        r = gold["grad:" + k]
        np.testing.assert_allclose(g.numpy(), r, atol=2e-3 * max(float(np.abs(r).max()), 1e-3), err_msg=k)


def test_dropin_layout_for_real_lstm():
    import models
    models.cfg.lstm = "real"
    try:
        torch.manual_seed(0)
        m = models.DCCRN(masking_mode="C")
        ref = O.init_state(0, lstm="real")
        sd = m.state_dict()
        assert set(sd.keys()) == set(ref.keys())
        assert [k for k in sd if k.startswith(("enhance.", "tranform."))] == [k for k in ref if k.startswith(("enhance.", "tranform."))]
        for k in ref:
            assert sd[k].shape == ref[k].shape, k
            if sd[k].is_floating_point() and not k.startswith(("stft.", "istft.")):
                assert torch.equal(sd[k], ref[k]), k            # same RNG stream as the reference constructor
    finally:
        models.cfg.lstm = "complex"


@pytest.mark.gpu
def test_real_lstm_gpu(engine):
    """Drop-in DCCRN with cfg.lstm = 'real' through the C ABI: waveform, loss and every gradient against the oracle, and the
    reference's own fixture values."""
    import models
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reallstm_golden.npz"), allow_pickle=False)
    models.cfg.lstm, models.cfg.loss = "real", "SI-SNR"
    try:
        sd0 = O.init_state(0, lstm="real")
        noisy, clean = _speech()
        tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
        loss_ref, wav_ref = tr.forward_backward(noisy, clean)
        m = models.DCCRN(masking_mode="C")
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
        for k, p in m.named_parameters():
            if k.endswith("_conv.bias") and not k.startswith("decoder.5"):
                continue                      # zero by BatchNorm
            g, r = p.grad.detach().cpu().double().reshape(-1), grads[k].double().reshape(-1)
            cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
            nr = float(g.norm() / (r.norm() + 1e-30))
            if k.endswith(".2.weight"):       # the single PReLU slope: one cancelling global sum
                assert abs(float(g[0] - r[0])) <= 2e-2 * max(float(v.abs().max()) for n, v in grads.items() if n.endswith(".2.weight")), k
            else:
                assert cos > (0.99 if tf else 0.9995) and abs(nr - 1) < (0.05 if tf else 0.02), (k, cos, nr)
    finally:
        models.cfg.lstm = "complex"
