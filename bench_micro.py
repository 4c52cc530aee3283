#!/usr/bin/env python
"""BASELINE.json configs[4]: STFT -> mask -> ISTFT kernel microbench through the op-level C ABI.

    python bench_micro.py [--iters 20]

One long fp32 signal of T frames, T = 1k .. 64k, for both transform geometries of the reference's config.py:55-61
(fft 512: win 400 / hop 100 / F = 257 bins; fft 1024: win 800 / hop 200 / F = 513 bins).  Reports achieved HBM GB/s of
  fused       : wave -> STFT -> complex mask -> ISTFT -> wave in ONE kernel: reads 4*hop (wave) + 8*(F-1) (mask), writes
                4*hop                                                                  per frame
  stft        : reads 4*hop (wave), writes 8*F (spectrum)                              per frame  } the training-realistic
  mask_istft  : reads 8*F (spectrum) + 8*(F-1) (mask), writes 4*hop (wave)             per frame  } split
against MEASURED_PEAKS.json's copy bandwidth.  Inputs are resident in HBM; a 256 MB memset flushes L2 (126 MB) between
timed launches.  One JSON line per (geometry, size).
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import torch
from sefd import _lib
from sefd.ops import ptr, stream

GEOMETRIES = {512: (257, 100), 1024: (513, 200)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nfft", type=int, default=0, help="512 or 1024 (default: both)")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured"
    except OSError:
        hbm, src = 6650.0, "fallback"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for nfft in ((args.nfft,) if args.nfft else (512, 1024)):
        F, HOP = GEOMETRIES[nfft]
        for T in (1024, 4096, 16384, 65536):
            L = HOP * (T - 3)
            g = torch.Generator(device="cpu").manual_seed(1234)
            wav = ((torch.rand(1, L, generator=g) * 2 - 1) * 0.1).to(dev)
            mask = torch.randn(1, F - 1, T, 2, generator=g).to(dev)
            spec = torch.empty(1, F, T, 2, device=dev)
            out = torch.empty(1, L, device=dev)
            st = stream()

            def run_fused():
                _lib.check(lib.sefd_stft_mask_istft_fused(ptr(wav), ptr(mask), 2, 1, L, nfft, ptr(out), st), "fused")

            def run_stft():
                _lib.check(lib.sefd_stft_forward_n(ptr(wav), ptr(spec), 1, L, nfft, st), "stft")

            def run_istft():
                _lib.check(lib.sefd_mask_istft_forward_n(ptr(spec), ptr(mask), 2, 1, L, nfft, ptr(out), st), "mask_istft")

            res = {}
            for name, fn, nbytes in (("fused", run_fused, T * (8 * (F - 1) + 8 * HOP)),
                                     ("stft", run_stft, T * (4 * HOP + 8 * F)),
                                     ("mask_istft", run_istft, T * (8 * F + 8 * (F - 1) + 4 * HOP))):
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
            print(json.dumps({"metric": "STFT->mask->ISTFT microbench (BASELINE configs[4]): fused kernel and the "
                                        "training-realistic split", "frames": T, "bins": F, "hop": HOP, "fft": nfft,
                              "dtype": "f32", "hbm_peak_gbs": hbm, "peak_source": src,
                              "l2": "256 MB flush between timed launches", **res}), flush=True)


if __name__ == "__main__":
    main()
