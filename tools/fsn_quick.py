"""Quick device timing of the FullSubNet train step (development aid): python tools/fsn_quick.py [B] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import ctypes as C
import torch
import models
from sefd import _lib
from sefd.train import FsnTrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
models.cfg.loss = "MSE"
torch.manual_seed(0)
m = models.FullSubNet().cuda().train()
g = torch.Generator().manual_seed(1)
noisy = ((torch.rand(B, 48000, generator=g) * 2 - 1) * 0.1).cuda()
clean = ((torch.rand(B, 48000, generator=g) * 2 - 1) * 0.1).cuda()
ts = FsnTrainStep(m)
for _ in range(2):
    l = ts.step(noisy, clean)
torch.cuda.synchronize()
print("ws GB", m._get_engine().plan(B, 161).ws_bytes / 1e9, "loss", float(l))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    l = ts.step(noisy, clean)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"B={B}: {ms:.2f} ms/step, {B / ms * 1e3:.1f} utt/s, loss {float(l):.5f}")
lib.sefd_prof_reset(); lib.sefd_prof_enable(1)
ts.step(noisy, clean)
torch.cuda.synchronize()
lib.sefd_prof_enable(0)
names = ["tapgemm", "wgrad", "bn", "lstm", "stft/feat", "misc", "skinny"]
for c in range(7):
    t, n, f, b = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
    lib.sefd_prof_get(c, C.byref(t), C.byref(n), C.byref(f), C.byref(b))
    if n.value:
        print(f"  {names[c]:10s} {t.value:8.3f} ms  {n.value:5d} launches  {f.value / 1e9 / max(t.value, 1e-9):8.1f} TFLOP/s  {b.value / 1e6 / max(t.value, 1e-9):8.1f} GB/s")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
lib.sefd_prof_dump(os.path.join(ROOT, "gpurun_out", "fsn_profile.csv").encode())
