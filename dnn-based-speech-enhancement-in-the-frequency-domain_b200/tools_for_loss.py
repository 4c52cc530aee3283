"""Drop-in for the loss functions of the reference's tools_for_loss.py that sit on the DCCRN path
(`from tools_for_loss import sdr, si_sdr, si_snr, get_array_lms_loss, get_array_pmsqe_loss`, models.py:9).
The arithmetic runs in libsefd.so (csrc/stft.cu loss kernels); argument order and sign conventions follow
tools_for_loss.py:29-94."""
from sefd import ops as _ops


def l2_norm(s1, s2):
    raise NotImplementedError("l2_norm is folded into the sefd loss kernels")


def si_snr(s1, s2, eps=1e-8):
    """tools_for_loss.py:36-44: s1 = estimate, s2 = target; returns the batch-mean SI-SNR in dB."""
    return -_ops.loss(s1, s2, "SI-SNR")


def sdr(s1, s2, eps=1e-8):
    """tools_for_loss.py:29-33: s1 = target, s2 = estimate."""
    return -_ops.loss(s2, s1, "SDR")


def si_sdr(reference, estimation, eps=1e-8):
    """tools_for_loss.py:47-94."""
    return -_ops.loss(estimation, reference, "SI-SDR")


def get_array_lms_loss(clean_array, est_array):
    """tools_for_loss.py:241-249: mean over the batch of the multi-scale log-mel distance (scales 16/32/64) between two
    magnitude arrays [B,257,T]; differentiable with respect to est_array."""
    return _ops.lms_loss_mags(clean_array, est_array)


def get_array_pmsqe_loss(clean_array, est_array):
    """tools_for_loss.py:259-269: PIT-wrapped PMSQE between the 1-second chunks of the two waveform batches [N, L].
    PARITY UNPINNED: the reference delegates the arithmetic to asteroid (SingleSrcPMSQE / PITLossWrapper / STFTFB), which is
    not part of the reference tree; csrc/pmsqe.cu restates the published algorithm with every perceptual table as an input
    (sefd.ops.pmsqe_tables accepts asteroid's own buffers)."""
    return _ops.pmsqe_loss(clean_array, est_array)
