"""Parity AT the bench configuration (BASELINE configs[1]: DCCRN mask C, SI-SNR, batch 32, 3 s @ 16 kHz, the product's tcgen05 TF32
engine): enhanced waveform, loss and the whole flat gradient of one train step against the CPU oracle on the same 32 pairs.
North-star bar: waveform RMSE < 1e-4.  (The oracle's forward + backward at this size takes ~10-20 s of host time.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bench_configuration_against_oracle():
    import models
    from oracle import dccrn_oracle as O
    from sefd import _lib
    _lib.load().sefd_set_engine(1)
    models.cfg.loss = "SI-SNR"
    B, L = 32, 48000
    sd0 = O.init_state(0)
    noisy, clean = O.synthetic_batch(B, L)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
    loss_ref, wav_ref = tr.forward_backward(noisy, clean)
    m = models.DCCRN(masking_mode="C")
    m.load_state_dict(sd0)
    m = m.cuda().train()
    _, _, wav = m(noisy.cuda(), clean.cuda())
    loss = m.loss(wav, clean.cuda())
    loss.backward()
    torch.cuda.synchronize()
    rmse = float((wav.detach().cpu() - wav_ref).pow(2).mean().sqrt())
    assert rmse < 1e-4, rmse                                                     # north-star tolerance
    assert float(loss.detach()) == pytest.approx(float(loss_ref), rel=1e-3)
    g = torch.cat([p.grad.detach().cpu().double().reshape(-1) for _, p in m.named_parameters()])
    r = torch.cat([tr.sd[n].grad.double().reshape(-1) for n, _ in m.named_parameters()])
    cos, nr = float((g * r).sum() / (g.norm() * r.norm())), float(g.norm() / r.norm())
    assert cos > 0.999 and abs(nr - 1.0) < 0.01, (cos, nr)
