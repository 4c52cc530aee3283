"""Pin the CPU oracle (oracle/dccrn_oracle.py) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import dccrn_oracle as O


@pytest.fixture(scope="module")
def sd0():
    return O.init_state(0)


def test_init_stream_matches_reference(golden, sd0):
    keys = [str(k) for k in golden["init_keys"]]
    assert keys == list(sd0.keys()) or set(keys) == set(sd0.keys())
    for k, s, a in zip(keys, golden["init_sum"], golden["init_abs"]):
        v = sd0[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), k
    n = sum(sd0[k].numel() for k in O.trainable_keys(sd0))
    assert n == int(golden["n_params"]) == 3671053


def test_loss_known_answers(golden):
    ref = torch.from_numpy(golden["si_sdr_doc_inputs"])[None]
    flip = torch.flip(ref, [-1])
    got = [float(O.si_sdr(ref, flip)), float(O.si_sdr(ref, ref + flip)),
           float(O.si_sdr(ref, ref + 0.5)), float(O.si_sdr(ref, ref * 2 + 1))]
    # tools_for_loss.py:57-74 doctest values (the torch code adds eps, so 1e-4 abs)
    np.testing.assert_allclose(got, golden["si_sdr_doc_expected"], atol=1e-4)
    np.testing.assert_allclose(got, golden["si_sdr_ref_values"], rtol=1e-12)
    a, b = torch.from_numpy(golden["loss_pair_a"]), torch.from_numpy(golden["loss_pair_b"])
    assert float(O.si_snr(a, b)) == pytest.approx(float(golden["loss_ref_si_snr"]), rel=1e-6)
    assert float(O.sdr(a, b)) == pytest.approx(float(golden["loss_ref_sdr"]), rel=1e-6)
    assert float(O.si_sdr(a, b)) == pytest.approx(float(golden["loss_ref_si_sdr"]), rel=1e-6)


@pytest.mark.parametrize("inputs", ["rand", "speech"])
@pytest.mark.parametrize("mode", ["C", "E", "R"])
def test_forward_backward_small(golden, sd0, inputs, mode):
    noisy = torch.from_numpy(golden[f"small_{inputs}_noisy"])
    clean = torch.from_numpy(golden[f"small_{inputs}_clean"])
    for loss_name in (["SI-SNR", "SDR", "SI-SDR", "MSE"] if mode == "C" else ["SI-SNR"]):
        tr = O.OracleTrainer(sd0, masking_mode=mode, loss=loss_name)
        taps = {}
        for k in tr.keys:
            tr.sd[k].grad = None
        o_r, o_i, wav = O.dccrn_forward(tr.sd, noisy, mode, train=True, taps=taps)
        loss = O.dccrn_loss(wav, clean, loss_name)
        loss.backward()
        tag = f"small_{inputs}_{mode}_{loss_name}"
        assert float(loss) == pytest.approx(float(golden[tag + "_loss"]), rel=2e-5, abs=2e-5)
        if loss_name == "SI-SNR":
            np.testing.assert_allclose(wav.detach().numpy(), golden[tag + "_wav"], atol=2e-6)
            np.testing.assert_allclose(o_r.detach().numpy(), golden[tag + "_out_real"], atol=2e-5, rtol=1e-5)
            np.testing.assert_allclose(o_i.detach().numpy(), golden[tag + "_out_imag"], atol=2e-5, rtol=1e-5)
        names = [str(n) for n in golden["param_names"]]
        gn = np.array([float(tr.sd[n].grad.double().norm()) for n in names])
        ref = golden[tag + "_gnorm"]
        np.testing.assert_allclose(gn, ref, rtol=2e-3, atol=1e-5 * float(ref.max()))


def test_sampled_gradients(golden, sd0):
    noisy = torch.from_numpy(golden["small_speech_noisy"])
    clean = torch.from_numpy(golden["small_speech_clean"])
    tr = O.OracleTrainer(sd0)
    tr.forward_backward(noisy, clean)
    pre = "small_speech_C_SI-SNR_grad::"
    n_checked = 0
    for k in golden.files:
        if not k.startswith(pre):
            continue
        g = tr.sd[k[len(pre):]].grad.reshape(-1)
        g = g if g.numel() <= 4096 else g[:: g.numel() // 2048][:2048]
        ref = golden[k]
        # conv biases that feed a BatchNorm have an analytically zero gradient: both sides hold only
        # fp32 rounding noise there, hence the absolute floor
        np.testing.assert_allclose(g.numpy(), ref, rtol=5e-3, atol=max(2e-4 * float(np.abs(ref).max()), 3e-5))
        n_checked += 1
    assert n_checked > 50
    # BN running statistics after one train-mode forward (momentum 0.1, unbiased var)
    for k in golden.files:
        if k.startswith("small_speech_bn::"):
            np.testing.assert_allclose(tr.sd[k.split("::")[1]].detach().numpy(), golden[k], rtol=1e-4, atol=1e-6)


def test_eval_forward(golden, sd0):
    noisy = torch.from_numpy(golden["small_speech_noisy"])
    clean = torch.from_numpy(golden["small_speech_clean"])
    tr = O.OracleTrainer(sd0)
    tr.forward_backward(noisy, clean)            # one train forward updates the running stats
    with torch.no_grad():
        _, _, wav = O.dccrn_forward(tr.sd, noisy, "C", train=False)
    np.testing.assert_allclose(wav.numpy(), golden["small_speech_C_eval_wav"], atol=2e-6)


def test_three_adam_steps(golden, sd0):
    noisy = torch.from_numpy(golden["small_speech_noisy"])
    clean = torch.from_numpy(golden["small_speech_clean"])
    tr = O.OracleTrainer(sd0)
    losses = [float(tr.step(noisy, clean)[0]) for _ in range(3)]
    np.testing.assert_allclose(losses, golden["adam3_losses"], rtol=1e-3, atol=1e-3)
    names = [str(n) for n in golden["param_names"]]
    # biases in front of a BatchNorm only ever see rounding-noise gradients, and Adam turns noise into
    # +-lr steps, so they are excluded from the comparison
    keep = [i for i, n in enumerate(names) if not (n.endswith("_conv.bias") and not n.startswith("decoder.5."))]
    s = np.array([float(tr.sd[n].detach().double().abs().sum()) for n in names])
    np.testing.assert_allclose(s[keep], golden["adam3_param_abs"][keep], rtol=1e-3)


def test_full_length_known_answer(golden, sd0):
    noisy, clean = O.synthetic_batch(2)
    tr = O.OracleTrainer(sd0)
    loss, wav = tr.forward_backward(noisy, clean)
    assert float(loss) == pytest.approx(43.71657562, rel=1e-5)
    assert float(loss) == pytest.approx(float(golden["full_loss"]), rel=1e-5)
    np.testing.assert_allclose(wav[:, :2048].numpy(), golden["full_wav_head"], atol=2e-6)
    assert float(wav.double().pow(2).mean().sqrt()) == pytest.approx(float(golden["full_wav_rms"]), rel=1e-5)
    names = [str(n) for n in golden["param_names"]]
    gn = np.array([float(tr.sd[n].grad.double().norm()) for n in names])
    np.testing.assert_allclose(gn, golden["full_gnorm"], rtol=5e-3, atol=1e-5 * float(golden["full_gnorm"].max()))


def test_fused_lstm_equals_explicit_loop():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(37, 5, 128, generator=g, requires_grad=True)
    w = [torch.randn(512, 128, generator=g) * 0.1, torch.randn(512, 128, generator=g) * 0.1,
         torch.randn(512, generator=g) * 0.1, torch.randn(512, generator=g) * 0.1]
    a = O.lstm_seq(x, *w)
    b = O.lstm_fused(x, *w)
    np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), atol=2e-6)
    ga, = torch.autograd.grad(a.sum(), x)
    gb, = torch.autograd.grad(b.sum(), x)
    np.testing.assert_allclose(ga.numpy(), gb.numpy(), atol=2e-5)
