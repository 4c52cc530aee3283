"""LMS perceptual loss (tools_for_loss.py:111-249) and the DCCRN perceptual train step
(trainer.py:44-70: loss = (main + perceptual) / 2).  CPU: oracle vs fixtures from the unmodified reference
(tests/golden/make_golden.py lms).  GPU (-m gpu): CUDA kernels through the drop-in modules vs oracle and fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import dccrn_oracle as O

DEV = "cuda"


@pytest.fixture(scope="module")
def lms_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "lms_golden.npz"), allow_pickle=False)


def _speech(B=2, L=4000):
    g = torch.Generator().manual_seed(7)
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(B)])
    return clean + 0.05 * torch.randn(B, L, generator=g), clean


def test_lms_oracle_op(lms_golden):
    clean = torch.from_numpy(lms_golden["op_clean"])
    est = torch.from_numpy(lms_golden["op_est"]).requires_grad_(True)
    loss = O.lms_loss(clean, est)
    loss.backward()
    assert float(loss) == pytest.approx(float(lms_golden["op_loss"]), rel=1e-6)
    np.testing.assert_allclose(est.grad.numpy(), lms_golden["op_grad"], rtol=1e-4, atol=1e-9)


def test_lms_oracle_model(lms_golden):
    sd0 = O.init_state(0)
    noisy, clean = _speech()
    tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
    o_r, o_i, wav = O.dccrn_forward(tr.sd, noisy, "C", train=True, taps={})
    main = O.dccrn_loss(wav, clean, "SI-SNR")
    perc = O.dccrn_lms_loss(tr.sd, o_r, o_i, clean)
    total = (main + perc) / 2
    total.backward()
    assert float(main) == pytest.approx(float(lms_golden["model_main"]), rel=2e-5)
    assert float(perc) == pytest.approx(float(lms_golden["model_perc"]), rel=2e-5)
    names = [str(n) for n in lms_golden["param_names"]]
    gn = np.array([float(tr.sd[k].grad.double().norm()) for k in names])
    np.testing.assert_allclose(gn, lms_golden["model_gnorm"], rtol=2e-3, atol=2e-4 * lms_golden["model_gnorm"].max())


@pytest.mark.gpu
def test_lms_gpu_op(lms_golden):
    import tools_for_loss as tfl
    clean = torch.from_numpy(lms_golden["op_clean"]).to(DEV)
    est = torch.from_numpy(lms_golden["op_est"]).to(DEV).requires_grad_(True)
    loss = tfl.get_array_lms_loss(clean, est)
    loss.backward()
    assert float(loss) == pytest.approx(float(lms_golden["op_loss"]), rel=2e-5)
    g = est.grad.cpu().numpy()
    ref = lms_golden["op_grad"]
    assert np.abs(g - ref).max() <= 2e-4 * np.abs(ref).max()


@pytest.mark.gpu
def test_lms_gpu_model_perceptual_step(lms_golden, engine):
    import models
    tf = engine == 1
    models.cfg.loss, models.cfg.perceptual = "SI-SNR", "LMS"
    try:
        sd0 = O.init_state(0)
        noisy, clean = _speech()
        m = models.DCCRN(masking_mode="C")
        m.load_state_dict(sd0)
        m = m.to(DEV).train()
        real, imag, wav = m(noisy.to(DEV))                              # trainer.py:59-61
        main = m.loss(wav, clean.to(DEV))
        perc = m.loss(wav, clean.to(DEV), real, imag, perceptual=True)
        total = (main + perc) / 2
        total.backward()
        assert float(main) == pytest.approx(float(lms_golden["model_main"]), rel=5e-3 if tf else 2e-4)
        assert float(perc) == pytest.approx(float(lms_golden["model_perc"]), rel=5e-3 if tf else 2e-4)
        names = [str(n) for n in lms_golden["param_names"]]
        ref = lms_golden["model_gnorm"]
        gmax = ref.max()
        for i, (n, p) in enumerate(m.named_parameters()):
            assert n == names[i]
            if n.endswith("_conv.bias") and not n.startswith("decoder.5."):
                continue
            gn = float(p.grad.double().norm())
            assert abs(gn - ref[i]) <= (5e-2 if tf else 5e-3) * ref[i] + 1e-4 * gmax, (n, gn, ref[i])
            r = lms_golden["model_grad::" + n]
            g = p.grad.detach().reshape(-1)
            g = (g if g.numel() <= 4096 else g[:: g.numel() // 2048][:2048]).cpu().numpy()
            cosv = float((g.astype(np.float64) * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
            assert cosv > (0.99 if tf else 0.9995) or np.abs(r).max() < 1e-5 * gmax, (n, cosv)
    finally:
        models.cfg.perceptual = False
