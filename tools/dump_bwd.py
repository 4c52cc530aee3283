"""Debug helper: run the small fp32 test case and dump backward intermediates to gpurun_out/<tag>.pt
(SEFD_LIB selects the library build, so two builds can be diffed offline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import torch
import models
from sefd import _lib
from oracle import dccrn_oracle as O
tag, engine = sys.argv[1], int(sys.argv[2])
lib = _lib.load()
lib.sefd_set_engine(engine)
models.cfg.loss = "SI-SNR"
sd0 = O.init_state(0)
noisy, clean = O.synthetic_batch(2, 4000)
m = models.DCCRN(masking_mode="C"); m.load_state_dict(sd0); m = m.cuda().train()
out = {}
for rep in range(2):
    for p in m.parameters(): p.grad = None
    _, _, wav = m(noisy.cuda(), clean.cuda())
    loss = m.loss(wav, clean.cuda()); loss.backward(); torch.cuda.synchronize()
    plan = m._get_engine().plan(*noisy.shape)
    for n in ["dec4.dz", "dec3.dz", "dec2.dz", "dec1.dz", "dec0.dz", "dU", "dX", "dH", "dG", "enc5.dz", "enc5.dz2", "enc4.dz", "enc4.dz2", "enc3.dz", "enc3.dz2", "enc2.dz", "enc2.dz2", "enc1.dz", "enc1.dz2", "enc0.dz", "enc0.dz2"]:
        try: out[f"{rep}.{n}"] = plan.tensor(n).detach().cpu().clone()
        except Exception as e: print("skip", n, e)
    for n, p in list(m.named_parameters())[:0]: out[f"{rep}.g.{n}"] = p.grad.detach().cpu().clone()
torch.save(out, f"/tmp/{tag}.pt")
print("dumped", tag, float(loss.detach()))
for k in out:
    if k.startswith("0."):
        d = (out[k] - out["1." + k[2:]]).abs().max()
        if d > 0: print("run-to-run", k, float(d), float(out[k].abs().max()))
if len(sys.argv) > 3:
    ref = torch.load(f"/tmp/{sys.argv[3]}.pt")
    for k in out:
        d = (out[k] - ref[k]).abs()
        if float(d.max()) > 1e-5 * float(ref[k].abs().max()):
            idx = (d > 0.5 * d.max()).nonzero()
            print("DIFF", k, "max", float(d.max()), "scale", float(ref[k].abs().max()), "n_bad", int((d > 1e-5 * ref[k].abs().max()).sum()), "of", d.numel(), "shape", tuple(d.shape), "worst idx", idx[:6].tolist())
