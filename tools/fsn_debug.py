"""Development aid: per-parameter gradient agreement of the CUDA FullSubNet against the oracle (tf32 engine)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import torch
from oracle import fullsubnet_oracle as FS
import models
B, Tf = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 14
g = torch.Generator().manual_seed(100 + B)
mag = torch.rand(B, 257, Tf, generator=g) * (0.2 + torch.rand(B, 257, 1, generator=g))
cirm = torch.randn(B, 257, Tf, 2, generator=g)
sd = FS.init_state(0)
ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
crm_ref = FS.fullsubnet_forward(ref, mag)
torch.nn.functional.mse_loss(cirm, crm_ref).backward()
models.cfg.loss = "MSE"
m = models.FullSubNet(); m.load_state_dict(sd); m = m.cuda().train(); m.dropout = 0.0
crm = m(mag.cuda()); m.loss(cirm.cuda(), crm).backward(); torch.cuda.synchronize()
print("crm max err", float((crm.detach().cpu() - crm_ref.detach()).abs().max()))
for k, p in m.named_parameters():
    a, r = p.grad.detach().cpu().double(), ref[k].grad.double()
    print(f"{k:45s} cos {float((a * r).sum() / (a.norm() * r.norm() + 1e-30)):.6f} ratio {float(a.norm() / (r.norm() + 1e-30)):.4f}")
