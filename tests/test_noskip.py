"""cfg.skip_type = False: DCCRN decoder without skip connections (models.py:138-169, 227-230).
CPU: oracle and plan layout vs fixtures of the unmodified reference (make_golden.py noskip); GPU: drop-in vs the fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import dccrn_oracle as O


@pytest.fixture(scope="module")
def noskip_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "noskip_golden.npz"), allow_pickle=False)


def _speech(B=2, L=4000):
    g = torch.Generator().manual_seed(7)
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(B)])
    return clean + 0.05 * torch.randn(B, L, generator=g), clean


def _sample(g):
    g = g.reshape(-1)
    return g if g.numel() <= 4096 else g[::97]


def test_noskip_oracle(noskip_golden):
    sd0 = O.init_state(0, skip_type=False)
    for k in ("decoder.0.0.real_conv.weight", "decoder.5.0.imag_conv.weight", "encoder.0.0.real_conv.weight"):
        np.testing.assert_array_equal(sd0[k].reshape(-1)[::97].numpy(), noskip_golden["init:" + k])   # same RNG stream
    noisy, clean = _speech()
    tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
    loss, wav = tr.forward_backward(noisy, clean)
    assert float(loss) == pytest.approx(float(noskip_golden["loss"]), rel=2e-5)
    np.testing.assert_allclose(wav.numpy(), noskip_golden["wav"], atol=2e-6)
    names = [str(n) for n in noskip_golden["param_names"]]
    assert sorted(names) == sorted(tr.keys)
    ref = noskip_golden["gnorm"]
    gn = np.array([float(tr.sd[k].grad.double().norm()) for k in names])
    np.testing.assert_allclose(gn, ref, rtol=2e-3, atol=2e-4 * ref.max())
    for k in names:
        if k.endswith("_conv.bias") and not k.startswith("decoder.5"):
            continue                          # zero by BatchNorm (rounding noise only)
        r = noskip_golden["grad:" + k]
        np.testing.assert_allclose(_sample(tr.sd[k].grad).numpy(), r, atol=2e-3 * max(np.abs(r).max(), 1e-3), err_msg=k)
    with torch.no_grad():
        ev = O.dccrn_forward(tr.sd, noisy, "C", train=False)[2]
    # the reference's eval pass ran after one train-mode forward: running stats moved once
    np.testing.assert_allclose(ev.numpy(), noskip_golden["eval_wav"], atol=5e-6)


def test_noskip_plan_layout(noskip_golden):
    from sefd import dccrn as d
    p = d.Plan(1, 100, "C", skip=False)
    names = [str(n) for n in noskip_golden["param_names"]]
    shapes = [str(s) for s in noskip_golden["param_shapes"]]
    assert [e[0] for e in p.params] == names
    assert [str(tuple(e[3])) for e in p.params] == shapes
    from sefd import _lib
    lib = _lib.load()
    assert not lib.sefd_dccrn_plan_create_ex(1, 100, 2, 8)              # unknown flag bits are an error, not ignored
    assert b"flag" in lib.sefd_last_error()


@pytest.mark.gpu
def test_noskip_gpu(noskip_golden, engine):
    import models
    tf = engine == 1
    models.cfg.skip_type, models.cfg.loss = False, "SI-SNR"
    try:
        sd0 = O.init_state(0, skip_type=False)
        noisy, clean = _speech()
        m = models.DCCRN(masking_mode="C")
        m.load_state_dict(sd0)
        m = m.cuda().train()
        o_r, o_i, wav = m(noisy.cuda(), clean.cuda())
        loss = m.loss(wav, clean.cuda())
        loss.backward()
        # every decoder tensor against the oracle's taps (channels-last workspace views)
        tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
        taps = {}
        tr.forward_backward(noisy, clean, taps)
        plan = m._get_engine().plan(*noisy.shape)
        for j in range(6):
            ref = taps[f"dec{j}_conv"].detach().permute(0, 2, 3, 1)
            got = plan.tensor(f"dec{j}.y").cpu()
            assert got.shape == ref.shape
            assert float((got - ref).abs().max()) <= (3e-2 if tf else 1e-4) * float(ref.abs().max()), f"dec{j}.y"
        assert float(loss.detach()) == pytest.approx(float(noskip_golden["loss"]), rel=5e-3 if tf else 1e-4)
        rmse = float((wav.detach().cpu() - torch.from_numpy(noskip_golden["wav"])).pow(2).mean().sqrt())
        assert rmse < (2e-3 if tf else 1e-5)
        np.testing.assert_allclose(o_r.detach().cpu().numpy(), noskip_golden["out_real"], atol=0.3 if tf else 2e-4)
        ref = noskip_golden["gnorm"]
        tol = 5e-2 if tf else 5e-3
        amax = max(ref[i] for i, (n, _) in enumerate(m.named_parameters()) if n.endswith(".2.weight"))
        for i, (n, p) in enumerate(m.named_parameters()):
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
            if n.endswith("_conv.bias") and not n.startswith("decoder.5"):
                continue                      # zero by BatchNorm (rounding noise only)
            gn = float(p.grad.double().norm())
            if tf and n.endswith(".2.weight"):      # PReLU slope: one cancelling global sum, TF32 noise does not cancel
                assert abs(gn - ref[i]) <= 2e-2 * amax, (n, gn, ref[i])
                continue
            assert abs(gn - ref[i]) <= tol * ref[i] + 1e-4 * ref.max(), (n, gn, ref[i])
            if not tf:
                r = torch.from_numpy(noskip_golden["grad:" + n]).double()
                g = _sample(p.grad.detach().cpu()).double()
                # kink tolerance: element-wise, or direction + norm
                ok = bool(((g - r).abs() <= 2e-3 * r.abs().max() + 1e-6).all())
                cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
                assert ok or cos > 0.9995, (n, cos)
        m.eval()
        with torch.no_grad():
            ev = m(noisy.cuda())[2]
        rmse = float((ev.cpu() - torch.from_numpy(noskip_golden["eval_wav"])).pow(2).mean().sqrt())
        assert rmse < (3e-3 if tf else 2e-5)
    finally:
        models.cfg.skip_type = True
