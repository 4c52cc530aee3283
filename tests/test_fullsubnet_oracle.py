"""FullSubNet: the oracle restatement against fixtures of the unmodified reference
(tests/golden/make_golden.py fullsubnet), and the feature / target kernels.  The model's CUDA path is checked in
tests/test_fullsubnet_gpu.py."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import fullsubnet_oracle as FS


@pytest.fixture(scope="module")
def fs_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "fullsubnet_golden.npz"), allow_pickle=False)


def _speech(B=2, L=4000):
    g = torch.Generator().manual_seed(7)
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(B)])
    return clean + 0.05 * torch.randn(B, L, generator=g), clean


def test_init_matches_reference_rng_stream(fs_golden):
    sd = FS.init_state(0)
    keys = [str(k) for k in fs_golden["init_keys"]]
    assert list(sd.keys()) == keys
    assert [str(tuple(sd[k].shape)) for k in keys] == [str(s) for s in fs_golden["init_shapes"]]
    np.testing.assert_allclose([float(sd[k].double().sum()) for k in keys], fs_golden["init_sum"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose([float(sd[k].double().abs().sum()) for k in keys], fs_golden["init_abs"], rtol=1e-9)
    assert sum(v.numel() for v in sd.values()) == int(fs_golden["n_params"]) == 5637635


def test_features_and_masks(fs_golden):
    noisy, clean = _speech()
    nc, cc = FS.stft(noisy), FS.stft(clean)
    ref = torch.stft(noisy, 512, 300, 400, window=torch.hann_window(400), return_complex=True)
    assert float((nc - ref).abs().max()) < 2e-5                                 # explicit framing + rFFT == torch.stft
    mag, _ = FS.mag_phase(nc)
    np.testing.assert_allclose(mag.numpy(), fs_golden["noisy_mag"], atol=2e-5)
    cirm = FS.build_complex_ideal_ratio_mask(nc, cc)
    np.testing.assert_allclose(cirm.numpy(), fs_golden["cIRM"], atol=2e-3, rtol=1e-3)   # 1 / (|noisy|^2 + eps) amplifies
    x = torch.randn(2, 1, 9, 5, generator=torch.Generator().manual_seed(1))
    u = FS.unfold(x, 2)
    assert u.shape == (2, 9, 1, 5, 5)
    xp = torch.nn.functional.pad(x.reshape(2, 1, 9, 5), [0, 0, 2, 2], mode="reflect")
    assert torch.equal(u[:, 4, 0], xp[:, 0, 4:9])
    assert torch.equal(u[:, 0, 0, 0], x[:, 0, 2])                               # reflect: row -2 is row 2
    m = torch.from_numpy(fs_golden["cRM"])
    np.testing.assert_allclose(FS.decompress_cirm(m).numpy(), fs_golden["decompressed"], rtol=1e-5, atol=1e-6)


def test_forward_loss_and_gradients(fs_golden):
    sd = {k: v.clone().requires_grad_(True) for k, v in FS.init_state(0).items()}
    noisy, clean = _speech()
    taps = {}
    loss = FS.train_step_loss(sd, noisy, clean, taps)
    loss.backward()
    np.testing.assert_allclose(taps["cRM"].numpy(), fs_golden["cRM"], atol=2e-5)
    assert float(loss.detach()) == pytest.approx(float(fs_golden["loss"]), rel=2e-4)
    names = [str(n) for n in fs_golden["param_names"]]
    gn = np.array([float(sd[k].grad.double().norm()) for k in names])
    np.testing.assert_allclose(gn, fs_golden["gnorm"], rtol=2e-3, atol=1e-7)
    for k in names:
        g = sd[k].grad.reshape(-1)
        g = g if g.numel() <= 4096 else g[::997]
        r = fs_golden["grad::" + k]
        np.testing.assert_allclose(g.numpy(), r, atol=2e-3 * max(float(np.abs(r).max()), 1e-8), err_msg=k)


def test_dropin_layout_matches_reference_and_refuses_cpu(fs_golden):
    """The drop-in constructs with the reference's state_dict keys / shapes / initial values for a torch seed (it consumes
    the RNG like nn.LSTM + nn.Linear do) and refuses to run without CUDA (no CPU fallback)."""
    import models
    torch.manual_seed(0)
    m = models.FullSubNet()
    sd = m.state_dict()
    ref = FS.init_state(0)
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]), k
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 257, 5))
    with pytest.raises(NotImplementedError):
        models.FullSubNet(sb_model_hidden_size=256)


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(2, 4000), (3, 48000), (1, 4801)])
def test_feature_kernels_gpu(fs_golden, B, L):
    """tools.stft / mag_phase / build_complex_ideal_ratio_mask / decompress_cIRM and the fused feature kernel (the feature /
    target side of trainer.fullsubnet_train) against the oracle and, at the fixture size, the reference's own values."""
    import tools_for_model as tools
    if (B, L) == (2, 4000):
        noisy, clean = _speech(B, L)
    else:
        g = torch.Generator().manual_seed(L)
        clean = 0.1 * torch.randn(B, L, generator=g)
        noisy = clean + 0.05 * torch.randn(B, L, generator=g)
    nc_ref, cc_ref = FS.stft(noisy), FS.stft(clean)
    nd, cd = noisy.cuda(), clean.cuda()
    nc, cc = tools.stft(nd), tools.stft(cd)
    assert nc.shape == nc_ref.shape and nc.dtype == torch.complex64
    scale = float(nc_ref.abs().max())
    assert float((nc.cpu() - nc_ref).abs().max()) < 2e-6 * scale + 1e-6
    mag, phase = tools.mag_phase(nc)
    mref, pref = FS.mag_phase(nc_ref)
    assert float((mag.cpu() - mref).abs().max()) < 2e-6 * scale + 1e-6
    big = mref > 1e-3 * scale                                   # the angle of a near-zero bin is ill-conditioned
    dphi = torch.remainder(phase.cpu() - pref + np.pi, 2 * np.pi) - np.pi
    assert float(dphi[big].abs().max()) < 1e-3
    # the mask divides by |noisy|^2: compare where the division is well conditioned, and bound everything by K = 10
    cirm = tools.build_complex_ideal_ratio_mask(nc, cc)
    cref = FS.build_complex_ideal_ratio_mask(nc_ref, cc_ref)
    assert cirm.shape == cref.shape and float(cirm.abs().max()) <= 10.0
    ok = big[..., None].expand_as(cref)
    assert float((cirm.cpu() - cref)[ok].abs().max()) < 5e-3
    fmag, fcirm = tools.fullsubnet_features(nd, cd)
    assert float((fmag - mag).abs().max()) < 2e-6 * scale + 1e-6
    assert float((fcirm.cpu() - cref)[ok].abs().max()) < 5e-3
    dec = tools.decompress_cIRM(cirm)
    np.testing.assert_allclose(dec.cpu().numpy(), FS.decompress_cirm(cirm.cpu()).numpy(), rtol=2e-4, atol=2e-4)
    if (B, L) == (2, 4000):
        np.testing.assert_allclose(fmag.cpu().numpy(), fs_golden["noisy_mag"], atol=2e-5)
        gref = torch.from_numpy(fs_golden["cIRM"])
        assert float((fcirm.cpu() - gref)[ok].abs().max()) < 5e-3


def test_istft_oracle_matches_torch():
    noisy, _ = _speech(2, 4000)
    spec = FS.stft(noisy)
    for length in (None, 4000, 3900):
        ref = torch.istft(spec, 512, 300, 400, window=torch.hann_window(400), length=length)
        got = FS.istft(spec, length)
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) < 2e-6
    assert float((FS.istft(spec, 4000) - noisy).abs().max()) < 1e-5          # analysis-synthesis round trip


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(2, 4000), (3, 48000), (1, 4801)])
def test_istft_gpu(B, L):
    import tools_for_model as tools
    g = torch.Generator().manual_seed(L + 1)
    x = 0.1 * torch.randn(B, L, generator=g)
    spec = FS.stft(x)
    # an arbitrary (not STFT-consistent) spectrum exercises the overlap-add, not just the round trip
    spec2 = spec * torch.exp(1j * 0.3 * torch.randn(spec.shape, generator=g)) * (1 + 0.2 * torch.randn(spec.shape, generator=g))
    for s, n in ((spec, L), (spec2, L), (spec2, None), (spec2, L - 77)):
        ref = FS.istft(s, n)
        got = tools.istft(s.cuda(), length=n)
        assert got.shape == ref.shape
        assert float((got.cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    got = tools.istft(torch.view_as_real(spec2).cuda(), length=L)                 # the reference passes the real view
    assert float((got.cpu() - FS.istft(spec2, L)).abs().max()) < 2e-5
    mag, phase = spec2.abs(), torch.angle(spec2)
    got = tools.istft((mag.cuda(), phase.cuda()), length=L, use_mag_phase=True)
    assert float((got.cpu() - FS.istft(spec2, L)).abs().max()) < 5e-5
    assert float((tools.istft(tools.stft(x.cuda()), length=L).cpu() - x).abs().max()) < 1e-5     # round trip on the device
