"""Small config-3 (D=3 bc256) training workload for ncu captures (one GPU): two epochs of a 1024^2 scene."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene_torch

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = make_scene_torch(4, side, side, 12, device="cuda")
scene = F.DeviceScene.from_image(img, 5)
torch.manual_seed(19920517)
model = LBDRNModel(4 * 49, 256, 4, 2)
tr = F.FusedTrainer(model, scene, 3, 1e-3, 8192, 2, flags=F.Flags(), sampler="device")
res = tr.run()
tr.close()
print("trained", len(res["losses"]), res["losses"][-1], res["val_mse"])
