"""Debug helper: poison the activation workspace with NaN before the first step; any NaN in outputs or gradients
means some kernel reads workspace memory before it was written in that step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import torch
import models
from sefd import _lib
from oracle import dccrn_oracle as O
engine = int(sys.argv[1])
lib = _lib.load(); lib.sefd_set_engine(engine)
models.cfg.loss = "SI-SNR"
sd0 = O.init_state(0)
B, L = int(sys.argv[2]), int(sys.argv[3])
noisy, clean = O.synthetic_batch(B, L)
m = models.DCCRN(masking_mode="C"); m.load_state_dict(sd0); m = m.cuda().train()
eng = m._get_engine(); eng.sync()
plan = eng.plan(B, L)
ws = plan.workspace(torch.device("cuda", 0)); ws.fill_(0xFF)
_, _, wav = m(noisy.cuda(), clean.cuda())
loss = m.loss(wav, clean.cuda()); loss.backward(); torch.cuda.synchronize()
print("engine", engine, "loss", float(loss.detach()), "wav nan", int(torch.isnan(wav).sum()))
for n, p in m.named_parameters():
    k = int(torch.isnan(p.grad).sum())
    if k: print("NaN grad", n, k, "of", p.grad.numel())
names = ["spec"] + [f"enc{i}.{s}" for i in range(6) for s in ("y", "z", "dz", "dz2")] + [f"dec{i}.{s}" for i in range(6) for s in ("y", "z", "dz")] + ["U", "dU", "dX", "dH", "dG", "X1", "X2"]
for n in names:
    try: t = plan.tensor(n)
    except Exception: continue
    k = int(torch.isnan(t).sum())
    if k: print("NaN tensor", n, k, "of", t.numel())
