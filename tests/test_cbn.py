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
