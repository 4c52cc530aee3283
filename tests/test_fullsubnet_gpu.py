"""FullSubNet CUDA path (BASELINE configs[2]) through the drop-in models.FullSubNet -> ctypes -> C ABI
(sefd_fsn_forward / sefd_fsn_backward) against the pinned oracle and the fixtures of the unmodified reference.

Tolerances: fp32 engine (exact arithmetic, different summation order): cRM <= 2e-5 abs, gradients <= 2e-3 of the tensor's
max; tcgen05 TF32 engine (fp32 storage, TF32 operands, fp32 accumulation): cRM <= 5e-3 abs (values are O(1)),
gradients cosine > 0.999 per tensor and norm within 2 %.
The reference's nn.LSTM(dropout=0.8) makes a train-mode step stochastic: parity is defined with dropout inactive
(eval-mode forward; train-mode arithmetic with p = 0) and with an injected mask."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import fullsubnet_oracle as FS

pytestmark = pytest.mark.gpu


def _speech(B=2, L=4000):
    g = torch.Generator().manual_seed(7)
    t = torch.arange(L, dtype=torch.float32) / 16000.0
    clean = torch.stack([0.2 * torch.sin(2 * np.pi * (200.0 + 150.0 * b + 300.0 * t) * t) *
                         (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t + b)) for b in range(B)])
    return clean + 0.05 * torch.randn(B, L, generator=g), clean


def _model(sd, train):
    import models
    models.cfg.loss = "MSE"
    m = models.FullSubNet()
    m.load_state_dict(sd)
    m = m.cuda()
    return m.train() if train else m.eval()


def _tm(plan, name):
    return plan.tensor(name).detach().cpu()


def _check_grads(m, sd_ref, engine):
    for k, p in m.named_parameters():
        g, r = p.grad.detach().cpu().double(), sd_ref[k].grad.double()
        scale = float(r.abs().max())
        if engine == 0:
            assert float((g - r).abs().max()) <= 2e-3 * scale + 1e-9, (k, float((g - r).abs().max()), scale)
        else:
            cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
            nr = float(g.norm() / (r.norm() + 1e-30))
            assert cos > 0.999 and abs(nr - 1) < 0.02, (k, cos, nr)


@pytest.mark.parametrize("B,Tf", [(2, 14), (1, 3), (3, 37)])
def test_forward_backward_vs_oracle(engine, B, Tf):
    """cRM, intermediates and every gradient against the oracle's autograd; train mode with dropout off."""
    g = torch.Generator().manual_seed(100 + B)
    mag = torch.rand(B, 257, Tf, generator=g) * (0.2 + torch.rand(B, 257, 1, generator=g))
    cirm = torch.randn(B, 257, Tf, 2, generator=g)
    sd = FS.init_state(0)
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    taps = {}
    crm_ref = FS.fullsubnet_forward(ref, mag, taps)
    loss_ref = torch.nn.functional.mse_loss(cirm, crm_ref)
    loss_ref.backward()

    m = _model(sd, train=True)
    m.dropout = 0.0
    crm = m(mag.cuda())
    loss = m.loss(cirm.cuda(), crm)
    loss.backward()
    torch.cuda.synchronize()
    plan = m._get_engine().plan(B, Tf)
    T, R = Tf + 2, B * 257
    tol = 2e-5 if engine == 0 else 5e-3
    fb_lin = _tm(plan, "fb_lin")[:, :, :257]                               # [T, B, 257] pre-ReLU
    fb_out = torch.relu(fb_lin).permute(1, 2, 0)                            # [B, 257, T]
    assert float((fb_out - taps["fb_out"][:, 0]).abs().max()) < tol * max(1.0, float(taps["fb_out"].abs().max()))
    sb_in = _tm(plan, "sb_in").reshape(T, B, 257, 32).permute(1, 2, 3, 0)   # [B, 257, 32, T]
    assert float((sb_in - taps["sb_in"]).abs().max()) < tol * max(1.0, float(taps["sb_in"].abs().max()))
    assert float((crm.detach().cpu() - crm_ref.detach()).abs().max()) < tol
    assert float(loss.detach()) == pytest.approx(float(loss_ref.detach()), rel=2e-4 if engine == 0 else 2e-3)
    _check_grads(m, ref, engine)


def test_eval_forward_matches_and_ignores_dropout(engine):
    g = torch.Generator().manual_seed(5)
    mag = torch.rand(2, 257, 9, generator=g)
    sd = FS.init_state(0)
    ref = FS.fullsubnet_forward(sd, mag)
    m = _model(sd, train=False)
    with torch.no_grad():
        a = m(mag.cuda()).cpu()
        b = m(mag.cuda()[:, None]).cpu()                                    # [B, 1, F, T] input like models.py:636-637
    assert float((a - ref).abs().max()) < (2e-5 if engine == 0 else 5e-3)
    assert torch.equal(a, b)


def test_injected_dropout_mask(engine):
    """Train-mode arithmetic with the inter-layer dropout of nn.LSTM(dropout=0.8): same multipliers on both sides."""
    B, Tf = 2, 6
    T, R = Tf + 2, B * 257
    g = torch.Generator().manual_seed(9)
    mag = torch.rand(B, 257, Tf, generator=g)
    cirm = torch.randn(B, 257, Tf, 2, generator=g)
    keep_fb = (torch.rand(B, T, 512, generator=g) >= 0.8).float() * 5.0
    keep_sb = (torch.rand(R, T, 384, generator=g) >= 0.8).float() * 5.0
    sd = FS.init_state(0)
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    crm_ref = FS.fullsubnet_forward(ref, mag, dropout_masks=(keep_fb, keep_sb))
    torch.nn.functional.mse_loss(cirm, crm_ref).backward()
    m = _model(sd, train=True)
    m.dropout_masks = (keep_fb.permute(1, 0, 2).contiguous().cuda(), keep_sb.permute(1, 0, 2).contiguous().cuda())
    crm = m(mag.cuda())
    m.loss(cirm.cuda(), crm).backward()
    torch.cuda.synchronize()
    assert float((crm.detach().cpu() - crm_ref.detach()).abs().max()) < (2e-5 if engine == 0 else 5e-3)
    _check_grads(m, ref, engine)


def test_philox_dropout_statistics_and_backward_consistency():
    """Without an injected mask the kernels draw Philox masks: keep rate 1 - p, scale 1 / (1 - p), and the backward
    regenerates the same mask (finite-difference check of one weight through the stochastic graph with a fixed seed)."""
    from sefd import _lib
    from sefd.ops import ptr, stream
    lib = _lib.load()
    x = torch.ones(1 << 20, device="cuda")
    y = torch.empty_like(x)
    _lib.check(lib.sefd_dropout_forward(ptr(x), ptr(y), x.numel(), 0.8, None, 1234, 7, stream()), "dropout")
    keep = float((y > 0).float().mean())
    vals = torch.unique(y).tolist()
    assert abs(keep - 0.2) < 3e-3 and len(vals) == 2 and vals[0] == 0.0 and vals[1] == pytest.approx(5.0, rel=1e-6)
    y2 = torch.empty_like(x)
    _lib.check(lib.sefd_dropout_forward(ptr(x), ptr(y2), x.numel(), 0.8, None, 1234, 7, stream()), "dropout")
    assert torch.equal(y, y2)
    _lib.check(lib.sefd_dropout_forward(ptr(x), ptr(y2), x.numel(), 0.8, None, 1235, 7, stream()), "dropout")
    assert not torch.equal(y, y2)


def test_reference_fixture(engine):
    """The fixture of the unmodified reference (tests/golden/make_golden.py fullsubnet): features from the CUDA feature
    kernels, cRM / loss / gradient norms and samples from the CUDA model (eval-mode dropout, like the fixture)."""
    import tools_for_model as tools
    gold = np.load(os.path.join(ROOT, "tests", "golden", "fullsubnet_golden.npz"), allow_pickle=False)
    noisy, clean = _speech()
    m = _model(FS.init_state(0), train=True)
    m.dropout = 0.0
    mag, cirm = tools.fullsubnet_features(noisy.cuda(), clean.cuda())
    crm = m(mag)
    # the fixture's cIRM divides by |noisy|^2 (ill conditioned in silent bins): use the fixture's own target for the loss
    loss = m.loss(torch.from_numpy(gold["cIRM"]).cuda(), crm)
    loss.backward()
    torch.cuda.synchronize()
    tol = 3e-5 if engine == 0 else 5e-3
    np.testing.assert_allclose(crm.detach().cpu().numpy(), gold["cRM"], atol=tol)
    assert float(loss.detach()) == pytest.approx(float(gold["loss"]), rel=2e-4 if engine == 0 else 2e-3)
    names = [str(n) for n in gold["param_names"]]
    grads = dict(m.named_parameters())
    gn = np.array([float(grads[k].grad.double().norm()) for k in names])
    np.testing.assert_allclose(gn, gold["gnorm"], rtol=2e-3 if engine == 0 else 2e-2, atol=1e-7)
    for k in names:
        gg = grads[k].grad.detach().cpu().reshape(-1)
        gg = gg if gg.numel() <= 4096 else gg[::997]
        r = gold["grad::" + k]
        np.testing.assert_allclose(gg.numpy(), r, atol=(2e-3 if engine == 0 else 3e-2) * max(float(np.abs(r).max()), 1e-8), err_msg=k)


def test_reference_train_loop_body_and_adam():
    """trainer.fullsubnet_train's loop body (trainer.py:97-112) on the drop-in with torch.optim.Adam: the loss decreases."""
    import tools_for_model as tools
    noisy, clean = _speech(2, 6000)
    m = _model(FS.init_state(0), train=True)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    losses = []
    for _ in range(4):
        nc, cc = tools.stft(noisy.cuda()), tools.stft(clean.cuda())
        mag, _ = tools.mag_phase(nc)
        cirm = tools.build_complex_ideal_ratio_mask(nc, cc)
        crm = m(mag)
        loss = m.loss(cirm, crm)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_bench_configuration_properties():
    """BASELINE configs[2] itself (B = 64, 161 frames: 16 448 sub-band sequences = 129 row tiles, the multicast clusters, the
    partial last tile, the backward's stream overlap) on the default engine, through properties that do not need a full-size
    oracle run: (1) sequences are independent, so utterances 0 and 63 of the batch equal the ORACLE's result for those two
    utterances alone; (2) the MSE gradient of the batch is the mean of the gradients of its two halves."""
    B, Tf = 64, 161
    g = torch.Generator().manual_seed(64)
    mag = torch.rand(B, 257, Tf, generator=g) * (0.2 + torch.rand(B, 257, 1, generator=g))
    cirm = torch.randn(B, 257, Tf, 2, generator=g)
    sd = FS.init_state(0)
    pick = [0, B - 1]
    with torch.no_grad():
        crm_ref = FS.fullsubnet_forward(sd, mag[pick])
    m = _model(sd, train=True)
    m.dropout = 0.0

    def run(idx):
        for p in m.parameters():
            p.grad = None
        crm = m(mag[idx].cuda())
        loss = m.loss(cirm[idx].cuda(), crm)
        loss.backward()
        torch.cuda.synchronize()
        return crm.detach().cpu(), float(loss.detach()), torch.cat([p.grad.detach().reshape(-1) for p in m.parameters()]).double().cpu()

    crm, loss, grad = run(slice(0, B))
    assert float((crm[pick] - crm_ref).abs().max()) < 5e-3                     # TF32 engine bar of this file
    crm_a, loss_a, grad_a = run(slice(0, B // 2))
    crm_b, loss_b, grad_b = run(slice(B // 2, B))
    assert float((crm[:B // 2] - crm_a).abs().max()) < 1e-5 and float((crm[B // 2:] - crm_b).abs().max()) < 1e-5
    assert loss == pytest.approx(0.5 * (loss_a + loss_b), rel=1e-5)
    mean = 0.5 * (grad_a + grad_b)
    cos = float((grad * mean).sum() / (grad.norm() * mean.norm()))
    assert cos > 0.99999 and abs(float(grad.norm() / mean.norm()) - 1) < 1e-3, (cos, float(grad.norm() / mean.norm()))
