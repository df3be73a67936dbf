"""Time the fused training kernel: us/step for bs=8192 on a resident scene. usage: time_train.py side [bs] [D] [bc]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene_torch
side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
D = int(sys.argv[3]) if len(sys.argv) > 3 else 2
bc = int(sys.argv[4]) if len(sys.argv) > 4 else 64
img = make_scene_torch(4, side, side, 12, device="cuda")
scene = F.DeviceScene.from_image(img, 5)
torch.manual_seed(19920517)
model = LBDRNModel(4 * (2 * D + 1) ** 2, bc, 4, 2)
tr = F.FusedTrainer(model, scene, D, 1e-3, bs, 10, flags=F.Flags(), sampler="device")
tr.begin()
perm = torch.randperm(side * side, device="cuda")
tr.train_epoch(perm, 1e-3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
losses = tr.train_epoch(perm, 1e-3)
e1.record()
torch.cuda.synchronize()
n = losses.numel()
print(f"side={side} bs={bs} D={D} bc={bc} steps={n} {e0.elapsed_time(e1) * 1e3 / n:.2f} us/step  loss {losses[0].item():.5f} -> {losses[-1].item():.5f}")
cur = tr.current_params()
for _ in range(3):
    e0.record()
    mse = tr.scene_mse(cur)
    e1.record()
    torch.cuda.synchronize()
    print(f"eval pass {e0.elapsed_time(e1):.2f} ms  mse {mse:.7f}")
tr.close()
