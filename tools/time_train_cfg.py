"""Time the fused training kernel on other feature sets: us/step for bs=8192 on a resident scene.
usage: time_train_cfg.py side [coords|coords_col|abs|d0|c8]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene_torch
side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
what = sys.argv[2] if len(sys.argv) > 2 else "coords"
C, D, bits = 4, 2, 12
fl = F.Flags()
if what == "coords":
    fl = F.Flags(use_coordinates=True, embedding=True, use_colors=False)
elif what == "coords_col":
    fl = F.Flags(use_coordinates=True, embedding=True)
elif what == "abs":
    fl = F.Flags(relative=False)
elif what == "d0":
    D = 0
elif what == "c8":
    C, bits = 8, 16
img = make_scene_torch(C, side, side, bits, device="cuda")
scene = F.DeviceScene.from_image(img, 5)
torch.manual_seed(19920517)
model = LBDRNModel(fl.dim_in(C, D), 64, C, 2)
tr = F.FusedTrainer(model, scene, D, 1e-3, 8192, 10, flags=fl, sampler="device")
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
print(f"{what}: side={side} dim_in={fl.dim_in(C, D)} C={C} D={D} steps={n} {e0.elapsed_time(e1) * 1e3 / n:.2f} us/step  loss {losses[0].item():.5f} -> {losses[-1].item():.5f}")
cur = tr.current_params()
for _ in range(2):
    e0.record()
    mse = tr.scene_mse(cur)
    e1.record()
    torch.cuda.synchronize()
    print(f"  eval pass {e0.elapsed_time(e1):.2f} ms  mse {mse:.7f}")
tr.close()
