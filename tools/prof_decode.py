"""Small decode / train workload for ncu captures (one GPU).  usage: prof_decode.py [decode|train] [side] [path]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene_torch

what = sys.argv[1] if len(sys.argv) > 1 else "decode"
side = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
path = sys.argv[3] if len(sys.argv) > 3 else "auto"
img = make_scene_torch(4, side, side, 12, device="cuda")
scene = F.DeviceScene.from_image(img, 5)
torch.manual_seed(19920517)
model = LBDRNModel(100, 64, 4, 2)
flat = (model.flat_params().view(torch.int32) & -65536).view(torch.float32).cuda()
if what == "decode":
    for _ in range(3):
        out = F.decode_image(scene.msb, flat, 5, 2, 64, 2, flags=F.Flags(), path=path, return_tensor=True, base_max=scene.msb_max)
    torch.cuda.synchronize()
    print("decoded", tuple(out.shape))
else:
    tr = F.FusedTrainer(model, scene, 2, 1e-3, 8192, 2, flags=F.Flags(), sampler="device")
    res = tr.run()
    tr.close()
    print("trained", len(res["losses"]), res["losses"][-1])
