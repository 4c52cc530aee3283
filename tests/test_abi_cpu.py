"""CPU-only checks of the drop-in boundary: libsefd.so loads, exports every symbol include/sefd.h declares,
the parameter layout equals the reference's named_parameters(), and the drop-in module reproduces the
reference's initial values for a torch seed.  No compute entry point is called (there is no GPU here)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT
from sefd import _lib
from sefd import dccrn as D


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sefd.h")).read()
    declared = set(re.findall(r"\b(sefd_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sefd.h but not exported by libsefd.so"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.sefd_abi_version() == 1


def test_plan_layout_matches_reference_named_parameters(golden):
    plan = D.Plan(2, 4000, "C")
    names = [n for n, _, _, _ in plan.params]
    assert names == [str(n) for n in golden["param_names"]]
    assert sum(n for _, _, n, _ in plan.params) == int(golden["n_params"]) == 3671053
    offs = [o for _, o, _, _ in plan.params]
    assert all(o % 4 == 0 for o in offs) and offs == sorted(offs)
    assert plan.T == 43 and plan.ws_bytes > 0
    assert [n for n, _, _, _ in plan.buffers][:2] == ["encoder.0.1.running_mean", "encoder.0.1.running_var"]


def test_plan_rejects_bad_geometry():
    with pytest.raises(RuntimeError, match="multiple of 100"):
        D.Plan(2, 4050, "C")
    lib = _lib.load()
    assert lib.sefd_dccrn_plan_create(2, 4000, 7) is None
    assert b"masking mode" in lib.sefd_last_error()


def test_dropin_init_matches_reference_rng_stream(golden):
    import models
    torch.manual_seed(0)
    m = models.DCCRN(masking_mode="C")
    sd = m.state_dict()
    keys = [str(k) for k in golden["init_keys"]]
    assert list(sd.keys()) == keys
    for k, s, a in zip(keys, golden["init_sum"], golden["init_abs"]):
        v = sd[k].double()
        assert abs(float(v.sum()) - s) <= 1e-6 * max(1.0, abs(a)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-6 * max(1.0, abs(a)), k


def test_dropin_refuses_cpu_and_unbuilt_configs():
    import models
    m = models.DCCRN(masking_mode="C")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 4000))
    models.DCCRN(masking_mode="Direct(None make)")          # spectral mapping is built
    with pytest.raises(NotImplementedError):
        models.DCCRN(masking_mode="X")
    cbn = models.DCCRN(use_cbn=True)           # ComplexBatchNorm is built (plan flag SEFD_PLAN_CBN); CUDA only as well
    with pytest.raises(RuntimeError, match="CUDA"):
        cbn(torch.zeros(1, 4000))
    crn = models.CRN()                         # built; like DCCRN it has no CPU path
    with pytest.raises(RuntimeError, match="CUDA"):
        crn(torch.zeros(1, 4000))
    with pytest.raises(NotImplementedError):
        models.CRN(masking_mode="Direct(None make)")
    fsn = models.FullSubNet()                  # built: CUDA only as well
    with pytest.raises(RuntimeError, match="CUDA"):
        fsn(torch.zeros(1, 257, 5))
    plan = D.Plan(2, 11, None, family="fsn")
    assert [n for n, _, _, _ in plan.params] == [k for k, _ in fsn.named_parameters()]
    assert sum(n for _, _, n, _ in plan.params) == 5637635


def test_crn_dropin_layout_and_init_match_reference():
    import numpy as np
    import models
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crn_golden.npz"))
    torch.manual_seed(0)
    m = models.CRN()
    sd = m.state_dict()
    keys = [str(k) for k in g["init_keys"]]
    assert list(sd.keys()) == keys
    for k, s_, a in zip(keys, g["init_sum"], g["init_abs"]):
        v = sd[k].double()
        assert abs(float(v.sum()) - s_) <= 1e-6 * max(1.0, abs(a)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-6 * max(1.0, abs(a)), k
    plan = D.Plan(2, 4000, "E", family="crn")
    assert [n for n, _, _, _ in plan.params] == [str(n) for n in g["param_names"]]
    assert sum(n for _, _, n, _ in plan.params) == int(g["n_params"]) == 1703436


def test_fft_index_algebra_on_host(tmp_path):
    src = os.path.join(PKG, "csrc", "fft512_host_test.cpp")
    exe = str(tmp_path / "fft_test")
    subprocess.check_call(["g++", "-O2", "-o", exe, src])
    out = subprocess.check_output([exe]).decode()
    assert "max abs err" in out


def test_host_side_guards_of_the_later_additions():
    """Argument handling that needs no GPU: table packing overrides, unbuilt variants fail loudly, CPU tensors are refused."""
    import numpy as np
    import tools_for_model as tools
    import tools_for_loss as tfl
    from sefd import ops
    from sefd.train import TrainStep
    t0 = ops.pmsqe_tables()
    t1 = ops.pmsqe_tables(bark_matrix=np.ones((257, 49)), mask_sll=np.zeros(257))
    assert t0.shape == t1.shape == (257 * 49 + 3 * 49 + 257,)
    assert float(t1[: 257 * 49].sum()) == 257 * 49 and float(t1[-257:].abs().sum()) == 0.0
    assert torch.equal(t0[257 * 49: 257 * 49 + 147], t1[257 * 49: 257 * 49 + 147])       # untouched tables keep their defaults
    with pytest.raises((AssertionError, ValueError)):
        ops.pmsqe_tables(bark_matrix=np.ones((10, 49)))
    x = torch.zeros(2, 4000)
    for call in (lambda: tools.stft(x), lambda: tools.fullsubnet_features(x, x), lambda: tfl.get_array_pmsqe_loss(x, x),
                 lambda: tools.istft(torch.zeros(2, 257, 14, 2))):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()
    with pytest.raises(NotImplementedError):
        tools.stft(x, n_fft=1024)
    with pytest.raises(NotImplementedError):
        tools.istft(torch.zeros(2, 257, 14, 2), hop_length=128)
    with pytest.raises(NotImplementedError):
        tools.decompress_cIRM(torch.zeros(4), K=5)

    class _M:                                      # TrainStep validates `perceptual` before it touches the model
        pass
    with pytest.raises(NotImplementedError):
        TrainStep(_M(), perceptual="LMS")


def test_init_kernels_matches_the_oracle_bases():
    """tools_for_model.init_kernels (reference :16-33) on the host: same matrices as the oracle's float64 bases; numpy branch of
    compress_cIRM (:715-716)."""
    import tools_for_model as T
    from oracle import dccrn_oracle as O
    k_a, k_s, w = O.stft_bases()
    k, win = T.init_kernels(400, 100, 512, "hann")
    assert k.shape == (514, 1, 400) and win.shape == (1, 400, 1)
    assert float(np.abs(k[:, 0].numpy() - k_a).max()) < 1e-6 and float(np.abs(win.reshape(-1).numpy() - w).max()) < 1e-7
    ki, _ = T.init_kernels(400, 100, 512, "hann", invers=True)
    assert float(np.abs(ki[:, 0].numpy() - k_s).max()) < 1e-8
    np.testing.assert_allclose(T.compress_cIRM(np.array([-200.0, 0.0, 1.0])), [10 * (1 - np.exp(10)) / (1 + np.exp(10)), 0.0, 10 * np.tanh(0.05)], rtol=1e-12)
