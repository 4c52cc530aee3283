"""Development aid: time the DCCRN LSTM recurrence kernels through the op-level C ABI."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import torch
from sefd import _lib
from sefd.ops import ptr, stream
lib = _lib.load()
rows, T = 64, 483
w = torch.randn(2, 512, 128, device="cuda") * 0.05
g0 = torch.randn(2, rows, T, 512, device="cuda")
h = torch.empty(2, rows, T, 128, device="cuda"); c = torch.empty_like(h)
dh = torch.randn(2, rows, T, 128, device="cuda") * 0.01
dg = torch.empty_like(g0)
def run():
    g = g0.clone()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    _lib.check(lib.sefd_lstm_forward(ptr(w), ptr(g), ptr(h), ptr(c), rows, T, stream()), "f")
    e[1].record()
    _lib.check(lib.sefd_lstm_backward(ptr(w), ptr(g), ptr(c), ptr(dh), ptr(dg), rows, T, stream()), "b")
    e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), float(h.abs().sum()), float(dg.abs().sum())
for _ in range(3): r = run()
print("fwd %.3f ms (%.2f us/step)  bwd %.3f ms (%.2f us/step)  checksums %.4f %.4f" % (r[0], r[0] * 1e3 / T, r[1], r[1] * 1e3 / T, r[2], r[3]))
