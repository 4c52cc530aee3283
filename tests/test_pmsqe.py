"""PMSQE perceptual loss (tools_for_loss.py:255-269).  PARITY UNPINNED: asteroid is absent from the reference tree and from
this image, so there are no reference fixtures; the CUDA path is checked against oracle/pmsqe_oracle.py (value, pairwise
matrix, PIT choice and gradient through torch autograd of the oracle), and the oracle against the properties the published
algorithm guarantees."""
import itertools

import numpy as np
import pytest
import torch

from oracle import pmsqe_oracle as Q


def _pair(N=2, seconds=3, seed=3, noise=0.05):
    g = torch.Generator().manual_seed(seed)
    L = 16000 * seconds
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(N)])
    return clean, clean + noise * torch.randn(N, L, generator=g)


def test_oracle_stft_filters_and_tables():
    f = Q.stft_filters()
    assert f.shape == (514, 512)
    x = torch.randn(1, 1, 16000, generator=torch.Generator().manual_seed(0))
    mag = Q.encoder_mag(x)
    assert mag.shape == (1, 1, 257, 61)
    # the STFTFB scaling makes the frame transform energy-preserving under the sqrt-hann window at 50 % overlap:
    # sum over bins of |X|^2 (one-sided, DC/Nyquist halved) = sum_n (w[n] x[n])^2 / 128 * ... checked numerically
    fr = x[0, 0, :512] * torch.from_numpy(np.hanning(513)[:-1] ** 0.5).float()
    X = torch.fft.rfft(fr) / 16.0
    X[0] /= 2 ** 0.5
    X[256] /= 2 ** 0.5
    np.testing.assert_allclose(mag[0, 0, :, 0].numpy(), torch.sqrt(X.abs() ** 2 + 1e-8).numpy(), rtol=2e-4, atol=2e-5)
    tb = Q.tables()
    assert tb["bark"].shape == (257, 49) and int((tb["bark"] != 0).sum()) == 256      # every bin below Nyquist in one band
    assert float(tb["bark"][256].abs().sum()) == 0.0
    assert abs(float(tb["zw"][0]) - 0.25520097857560436) < 1e-7                        # asteroid's first Zwicker power
    assert abs(float(tb["zw"][4]) - 0.25168783742879913) < 1e-7
    assert abs(float(tb["mask"][11]) - 0.4 * 2.0 * 514 / 512 ** 2) < 1e-9


def test_oracle_properties():
    clean, noisy = _pair()
    same = float(Q.get_array_pmsqe_loss(clean, clean))
    l1 = float(Q.get_array_pmsqe_loss(clean, noisy))
    l2 = float(Q.get_array_pmsqe_loss(clean, clean + 4.0 * (noisy - clean)))
    assert 0.0 <= same < 1e-2 < l1 < l2 < 45.0 * (Q.ALPHA + Q.BETA)          # identical < noisy < noisier < the frame cap
    # SLL equalisation makes the loss invariant to the level of either signal
    assert float(Q.get_array_pmsqe_loss(clean, 0.5 * noisy)) == pytest.approx(l1, rel=2e-3)
    # PIT: permuting the 1-second chunks of the estimate alone leaves the loss unchanged
    perm = noisy.reshape(2, 3, 16000)[:, [2, 0, 1]].reshape(2, -1)
    assert float(Q.get_array_pmsqe_loss(clean, perm)) == pytest.approx(l1, rel=1e-5)
    c = Q.encoder_mag(clean.reshape(2, 3, 16000))
    e = Q.encoder_mag(perm.reshape(2, 3, 16000))
    assert Q.pit_pw_pt(e, c)[2] == [(2, 0, 1), (2, 0, 1)]                       # est chunk i holds clean chunk perm[i]


def test_host_tables_match_oracle():
    from sefd import ops
    t = ops.pmsqe_tables()
    tb = Q.tables()
    ref = torch.cat([tb["bark"].reshape(-1), tb["thr"], tb["zw"], tb["width"], tb["mask"]])
    assert t.shape == ref.shape
    np.testing.assert_allclose(t.numpy(), ref.numpy(), rtol=1e-6)
    from sefd import _lib
    assert _lib.load().sefd_pmsqe_table_floats() == t.numel()
    assert _lib.load().sefd_pmsqe_workspace_bytes(2, 48000) > 0
    assert _lib.load().sefd_pmsqe_workspace_bytes(2, 48001) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("seconds,N", [(3, 2), (1, 3), (2, 1)])
def test_pmsqe_gpu_value_and_gradient(seconds, N):
    import tools_for_loss as tfl
    from sefd import ops
    clean, noisy = _pair(N, seconds)
    if seconds == 3:                                     # make PIT pick a non-identity permutation for one utterance
        noisy = noisy.clone()
        noisy[0] = noisy[0].reshape(3, 16000)[[1, 2, 0]].reshape(-1)
    est = noisy.clone().requires_grad_(True)
    ref = Q.get_array_pmsqe_loss(clean, est)
    ref.backward()
    est_g = noisy.clone().cuda().requires_grad_(True)
    loss = tfl.get_array_pmsqe_loss(clean.cuda(), est_g)
    (2.0 * loss).backward()                              # upstream factor exercises gout
    assert float(loss.detach()) == pytest.approx(float(ref.detach()), rel=2e-4)
    g, r = est_g.grad.cpu().double() / 2.0, est.grad.double()
    assert bool(torch.isfinite(g).all())
    cos = float((g * r).sum() / (g.norm() * r.norm()))
    assert cos > 0.9999 and abs(float(g.norm() / r.norm()) - 1.0) < 2e-3, (cos, float(g.norm()), float(r.norm()))
    assert float((g - r).abs().max()) <= 5e-3 * float(r.abs().max())
    # the pairwise matrix the PIT search saw
    S = seconds
    c = Q.encoder_mag(clean.reshape(N, S, 16000))
    e = Q.encoder_mag(noisy.reshape(N, S, 16000))
    _, pw, perms = Q.pit_pw_pt(e, c)
    if seconds == 3:
        assert perms[0] != (0, 1, 2)


@pytest.mark.gpu
def test_pmsqe_perceptual_train_step_glue():
    """trainer.model_perceptual_train (trainer.py:44-70) with cfg.perceptual = 'PMSQE': (main + perceptual) / 2, backward
    through out_wav into every parameter."""
    import models
    from oracle import dccrn_oracle as O
    models.cfg.loss, models.cfg.perceptual = "SI-SNR", "PMSQE"
    try:
        clean, noisy = _pair(2, 1)
        m = models.DCCRN(masking_mode="C")
        m.load_state_dict(O.init_state(0))
        m = m.cuda().train()
        real_spec, img_spec, out = m(noisy.cuda())
        main = m.loss(out, clean.cuda())
        perc = m.loss(out, clean.cuda(), real_spec, img_spec, perceptual=True)
        ((main + perc) / 2).backward()
        assert 0.0 < float(perc.detach()) < 6.5
        ref = Q.get_array_pmsqe_loss(clean, out.detach().cpu())
        assert float(perc.detach()) == pytest.approx(float(ref), rel=5e-4)
        for n, p in m.named_parameters():
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
    finally:
        models.cfg.perceptual = False


@pytest.mark.gpu
def test_pmsqe_trainstep_matches_autograd_path():
    """sefd.train.TrainStep(perceptual='PMSQE') (the bench's device-resident step) fills the same flat gradient as the
    autograd drop-in running the loop body of trainer.model_perceptual_train."""
    import models
    from oracle import dccrn_oracle as O
    from sefd.train import TrainStep
    models.cfg.loss, models.cfg.perceptual = "SI-SNR", "PMSQE"
    try:
        clean, noisy = _pair(2, 1)
        clean, noisy = clean.cuda(), noisy.cuda()
        sd0 = O.init_state(0)
        m = models.DCCRN(masking_mode="C")
        m.load_state_dict(sd0)
        m = m.cuda().train()
        real_spec, img_spec, out = m(noisy, clean)
        loss = (m.loss(out, clean) + m.loss(out, clean, real_spec, img_spec, perceptual=True)) / 2
        loss.backward()
        ga = m._get_engine().flat_grad.clone()
        m2 = models.DCCRN(masking_mode="C")
        m2.load_state_dict(sd0)
        m2 = m2.cuda().train()
        ts = TrainStep(m2, loss="SI-SNR", perceptual="PMSQE")
        l2 = ts.forward_backward(noisy, clean)
        gb = m2._get_engine().flat_grad
        assert float(l2) == pytest.approx(float(loss.detach()), rel=1e-5)
        assert float((ga - gb).abs().max()) <= 1e-5 * float(ga.abs().max())
    finally:
        models.cfg.perceptual = False
