"""Debug helper: repeat the same train step N times in one process and report every tensor that differs from
the first repetition by more than atomics-rounding noise (finds races / stray writes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import torch
import models
from sefd import _lib
from oracle import dccrn_oracle as O
engine, B, L, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lib = _lib.load(); lib.sefd_set_engine(engine)
models.cfg.loss = "SI-SNR"
sd0 = O.init_state(0)
noisy, clean = O.synthetic_batch(B, L)
m = models.DCCRN(masking_mode="C"); m.load_state_dict(sd0); m = m.cuda().eval()   # eval: BN buffers stay fixed
m.train()
names = ["spec"] + [f"enc{i}.{s}" for i in range(6) for s in ("y", "z")] + ["X1", "X2", "U"] + \
        [f"dec{i}.{s}" for i in range(6) for s in ("y", "z")] + \
        ["dec5.dy"] + [f"dec{i}.{s}" for i in (4, 3, 2, 1, 0) for s in ("dz", "dy")] + ["dU", "dX", "dH", "dG"] + \
        [f"enc{i}.{s}" for i in (5, 4, 3, 2, 1, 0) for s in ("dz", "dz2", "dy")]
ref = None
nbad = 0
for rep in range(reps):
    m.load_state_dict(sd0)          # resets BN running stats so every repetition is identical
    for p in m.parameters(): p.grad = None
    _, _, wav = m(noisy.cuda(), clean.cuda())
    loss = m.loss(wav, clean.cuda()); loss.backward(); torch.cuda.synchronize()
    plan = m._get_engine().plan(B, L)
    cur = {}
    for n in names:
        try: cur[n] = plan.tensor(n).detach().clone()
        except Exception: pass
    for n, p in m.named_parameters(): cur["g." + n] = p.grad.detach().clone()
    if ref is None:
        ref = cur; continue
    first = True
    for k in cur:
        d = (cur[k] - ref[k]).abs()
        s = float(ref[k].abs().max())
        if float(d.max()) > 2e-5 * s + 1e-12:
            idx = (d > 0.5 * d.max()).nonzero()
            if first: nbad += 1
            print(f"rep {rep} {'FIRST ' if first else '      '}{k}: max {float(d.max()):.3e} scale {s:.3e} bad {int((d > 2e-5 * s).sum())}/{d.numel()} worst {idx[:4].tolist()}")
            if first and d.dim() == 4 and nbad <= 2:
                pm = d.amax(dim=3)                       # [B, F, T]
                bad_pos = (pm > 1e-4 * s).nonzero().tolist()
                print("   positions with error > 1e-4*scale:", len(bad_pos), bad_pos[:60])
                b0, f0, t0 = bad_pos[len(bad_pos) // 2]
                print("   diff at", (b0, f0, t0), "first 8 ch:", (cur[k] - ref[k])[b0, f0, t0, :8].tolist())
            first = False
print(f"engine {engine} B {B} L {L}: {nbad} divergent repetitions of {reps - 1}")
