"""Ragged / minimal shapes of the DCCRN and CRN paths against the oracle: batch 1, odd batch, waveform lengths that are
not multiples of the 128-frame GEMM tile, the 32-frame STFT tile or the 29-hop ISTFT chunk."""
import numpy as np
import pytest
import torch

from oracle import crn_oracle as CO
from oracle import dccrn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L", [(1, 1600), (3, 12300), (5, 100)])
def test_dccrn_odd_shapes(engine, B, L):
    import models
    tf = engine == 1
    models.cfg.loss = "SI-SNR"
    sd0 = O.init_state(0)
    noisy, clean = O.synthetic_batch(B, L, seed=3)
    tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
    loss_ref, wav_ref = tr.forward_backward(noisy, clean)
    m = models.DCCRN(masking_mode="C")
    m.load_state_dict(sd0)
    m = m.cuda().train()
    _, _, wav = m(noisy.cuda(), clean.cuda())
    loss = m.loss(wav, clean.cuda())
    loss.backward()
    rmse = float((wav.detach().cpu() - wav_ref).pow(2).mean().sqrt())
    assert rmse < (2e-3 if tf else 2e-6), rmse
    assert float(loss) == pytest.approx(float(loss_ref), rel=2e-2 if tf else 5e-4, abs=1e-3)
    grads = tr.grads()
    tot_ref = float(torch.sqrt(sum(g.double().pow(2).sum() for g in grads.values())))
    tot = float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in m.parameters())))
    assert abs(tot - tot_ref) <= (0.1 if tf else 0.02) * tot_ref, (tot, tot_ref)
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())


@pytest.mark.parametrize("B,L", [(1, 1600), (3, 12300)])
def test_crn_odd_shapes(engine, B, L):
    import models
    tf = engine == 1
    models.cfg.loss = "MSE"
    try:
        sd0 = CO.init_state(0)
        noisy, clean = O.synthetic_batch(B, L, seed=3)
        tr = CO.OracleTrainer(sd0, loss="MSE")
        loss_ref, wav_ref = tr.forward_backward(noisy, clean)
        m = models.CRN()
        m.load_state_dict(sd0)
        m = m.cuda().train()
        _, _, wav = m(noisy.cuda(), clean.cuda())
        loss = m.loss(wav, clean.cuda())
        loss.backward()
        rmse = float((wav.detach().cpu() - wav_ref).pow(2).mean().sqrt())
        assert rmse < (2e-3 if tf else 2e-6), rmse
        assert float(loss) == pytest.approx(float(loss_ref), rel=2e-2 if tf else 5e-4)
        assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    finally:
        models.cfg.loss = "SI-SNR"


def test_rejects_bad_inputs():
    import models
    m = models.DCCRN(masking_mode="C").cuda()
    with pytest.raises((RuntimeError, ValueError)):
        m(torch.zeros(2, 4050, device="cuda"))             # not a multiple of the hop
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 4000))                            # CPU tensor: no fallback
