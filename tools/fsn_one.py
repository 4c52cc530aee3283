"""Development aid: ONE FullSubNet train step (for ncu): python tools/fsn_one.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]
import torch
import models
from sefd.train import FsnTrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
models.cfg.loss = "MSE"
torch.manual_seed(0)
m = models.FullSubNet().cuda().train()
g = torch.Generator().manual_seed(1)
noisy = ((torch.rand(B, 48000, generator=g) * 2 - 1) * 0.1).cuda()
clean = ((torch.rand(B, 48000, generator=g) * 2 - 1) * 0.1).cuda()
ts = FsnTrainStep(m)
print(float(ts.step(noisy, clean)))
torch.cuda.synchronize()
