"""GPU diagnostic: structure of the tensor-core weight-gradient error (per tap / per 32-channel block)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200"))
import torch
from sefd import ops, _lib
from sefd.ops import ptr, stream
lib = _lib.load()
torch.manual_seed(0)
def cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
def bwd(x, wr, wi, dy, sentinel):
    B, F, T, Cin = x.shape; Cout = dy.shape[-1]
    dx = torch.zeros_like(x); dwr, dwi = torch.empty_like(wr), torch.empty_like(wi)
    dbr = torch.empty(Cout // 2, device="cuda"); dbi = torch.empty(Cout // 2, device="cuda")
    n = lib.sefd_cconv_workspace_bytes(Cin, Cout)
    ws = torch.full(((n + 255) // 256 * 64,), sentinel, device="cuda", dtype=torch.float32)
    _lib.check(lib.sefd_cconv2d_backward(ptr(x), ptr(wr), ptr(wi), ptr(dy), ptr(dx), ptr(dwr), ptr(dbr), ptr(dwi), ptr(dbi),
                                         B, F, T, Cin, Cout, ptr(ws), stream()), "bwd")
    torch.cuda.synchronize()
    return dx, dwr, ws
for (B, F, T, Cin, Cout) in [(1, 2, 32, 64, 64), (1, 8, 130, 64, 128), (2, 4, 21, 256, 256)]:
    x = torch.randn(B, F, T, Cin, device="cuda")
    wr = torch.randn(Cout // 2, Cin // 2, 5, 2, device="cuda") * 0.05
    wi = torch.randn(Cout // 2, Cin // 2, 5, 2, device="cuda") * 0.05
    dy = torch.randn(B, F // 2, T, Cout, device="cuda")
    res = {}
    for eng in (0, 1):
        lib.sefd_set_engine(eng)
        res[eng] = bwd(x, wr, wi, dy, 7.0)
    lib.sefd_set_engine(1)
    r, g = res[0][1], res[1][1]
    ws = res[1][2]
    nW = 10 * Cin * Cout
    part = ws[2 * nW: 18 * nW]
    print(f"== B{B} F{F} T{T} Cin{Cin} Cout{Cout}: |ref|={float(r.norm()):.3e} |got|={float(g.norm()):.3e} cos={cos(r, g):.4f} "
          f"nan={int(torch.isnan(g).sum())} zeros={float((g == 0).float().mean()):.3f}")
    for s in range(16):
        blk = part[s * nW:(s + 1) * nW]
        print(f"   partial {s}: sentinel frac {float((blk == 7.0).float().mean()):.3f} zero frac {float((blk == 0).float().mean()):.3f} "
              f"nan {int(torch.isnan(blk).sum())} absmax {float(blk[~torch.isnan(blk)].abs().max()) if blk.numel() else 0:.3e}")
        if float((blk == 7.0).float().mean()) == 1.0:
            break
    for kf in range(5):
        print("   tap kf", kf, " cos kt0 %.3f kt1 %.3f" % (cos(r[..., kf, 0], g[..., kf, 0]), cos(r[..., kf, 1], g[..., kf, 1])),
              " |got| %.3e |ref| %.3e" % (float(g[..., kf, :].norm()), float(r[..., kf, :].norm())))
    print("   dx cos", "%.5f" % cos(res[0][0], res[1][0]))
