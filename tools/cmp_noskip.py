"""Debug aid: per-tensor comparison of the skip_type=False forward against the oracle (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import dccrn_oracle as O
from test_noskip import _speech
import models
from sefd import _lib
_lib.load().sefd_set_engine(int(os.environ.get("ENGINE", "0")))
models.cfg.skip_type, models.cfg.loss = False, "SI-SNR"
sd0 = O.init_state(0, skip_type=False)
noisy, clean = _speech()
tr = O.OracleTrainer(sd0, masking_mode="C", loss="SI-SNR")
taps = {}
loss_ref, wav_ref = tr.forward_backward(noisy, clean, taps)
m = models.DCCRN(masking_mode="C"); m.load_state_dict(sd0); m = m.cuda().train()
o_r, o_i, wav = m(noisy.cuda(), clean.cuda())
loss = m.loss(wav, clean.cuda()); loss.backward()
plan = m._get_engine().plan(*noisy.shape)
cl = lambda x: x.permute(0, 2, 3, 1)
def chk(name, got, ref):
    g, r = got.detach().double().cpu(), ref.detach().double()
    print(f"{name:12s} max|err|={float((g-r).abs().max()):.3e} max|ref|={float(r.abs().max()):.3e}")
for i in range(6):
    chk(f"enc{i}.z", plan.tensor(f"enc{i}.z"), cl(taps[f"enc{i}"]))
B, T = noisy.shape[0], plan.T
U = plan.tensor("U")
chk("U real", U[..., :128], taps["lstm1_r"].reshape(T, B, 128, 4).permute(1, 3, 0, 2))
for j in range(6):
    chk(f"dec{j}.y", plan.tensor(f"dec{j}.y"), cl(taps[f"dec{j}_conv"]))
    if j < 5:
        chk(f"dec{j}.z", plan.tensor(f"dec{j}.z"), cl(taps[f"dec{j}"]))
chk("wav", wav, wav_ref)
print("loss", float(loss), float(loss_ref))
g = tr.grads()
for n, p in m.named_parameters():
    r = g[n].double(); q = p.grad.double().cpu()
    print(f"grad {n:40s} |got|={float(q.norm()):.4e} |ref|={float(r.norm()):.4e} cos={float((q*r).sum()/(q.norm()*r.norm()+1e-30)):.5f}")
