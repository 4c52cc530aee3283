#!/usr/bin/env python
"""BASELINE.json configs[4]: STFT -> mask -> ISTFT kernel microbench through the op-level C ABI.

    python bench_micro.py [--iters 20]

One long fp32 signal of T frames (win 400 / hop 100 / fft 512 -> 257 bins; the 513-bin geometry is not built:
the kernels are specialised to the reference's default 512-point transform, config.py:55-61).  Reports, per T,
achieved HBM GB/s of the two kernels of the training-realistic split:
  stft        : reads 4*hop (wave), writes 8*F (spectrum)                           per frame
  mask+istft  : reads 8*F (spectrum) + 8*F (mask), writes 4*hop (wave)              per frame
against MEASURED_PEAKS.json's copy bandwidth.  Inputs are resident in HBM; the buffers of the larger sizes
exceed L2 (126 MB) only from T = 16 384 up, so an L2 flush (256 MB memset) runs between timed launches.
One JSON line per size.
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import torch
from sefd import _lib
from sefd.ops import ptr, stream

F, HOP = 257, 100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured"
    except OSError:
        hbm, src = 6650.0, "fallback"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for T in (1024, 4096, 16384, 65536):
        L = HOP * (T - 3)
        g = torch.Generator(device="cpu").manual_seed(1234)
        wav = ((torch.rand(1, L, generator=g) * 2 - 1) * 0.1).to(dev)
        mask = torch.randn(1, 256, T, 2, generator=g).to(dev)
        spec = torch.empty(1, F, T, 2, device=dev)
        out = torch.empty(1, L, device=dev)
        st = stream()

        def run_stft():
            _lib.check(lib.sefd_stft_forward(ptr(wav), ptr(spec), 1, L, st), "stft")

        def run_istft():
            _lib.check(lib.sefd_mask_istft_forward(ptr(spec), ptr(mask), 2, 1, L, None, None, ptr(out), None, st), "mask_istft")

        res = {}
        for name, fn, nbytes in (("stft", run_stft, T * (4 * HOP + 8 * F)), ("mask_istft", run_istft, T * (16 * F + 4 * HOP))):
            for _ in range(3):
                fn()
            ms = 0.0
            for _ in range(args.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
            ms /= args.iters
            gbs = nbytes / ms / 1e6
            res[name] = {"us": round(ms * 1e3, 2), "gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm, 4),
                         "bytes_per_frame": nbytes // T}
        print(json.dumps({"metric": "STFT->mask->ISTFT microbench (BASELINE configs[4]), training-realistic split",
                          "frames": T, "bins": F, "hop": HOP, "dtype": "f32", "hbm_peak_gbs": hbm, "peak_source": src,
                          "l2": "256 MB flush between timed launches", **res}), flush=True)


if __name__ == "__main__":
    main()
