"""cfg.lstm = 'real' (models.py:96-105, 213-218): oracle restatement against fixtures of the unmodified reference
(tests/golden/make_golden.py reallstm).  The CUDA path of this variant is not built yet (H = 256 needs the cluster-split
recurrent engine); the drop-in raises."""
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


def test_dropin_raises_for_real_lstm():
    import models
    models.cfg.lstm = "real"
    try:
        with pytest.raises(NotImplementedError):
            models.DCCRN(masking_mode="C")
    finally:
        models.cfg.lstm = "complex"
