"""Time decode with coordinate / positional-encoding features (BASELINE config 4). usage: time_coords.py side use_colors path"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene_torch
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
use_colors = int(sys.argv[2]) if len(sys.argv) > 2 else 1
path = sys.argv[3] if len(sys.argv) > 3 else "auto"
fl = F.Flags(use_coordinates=True, embedding=True, use_colors=bool(use_colors))
dim_in = fl.dim_in(4, 2)
img = make_scene_torch(4, side, side, 12, device="cuda")
scene = F.DeviceScene.from_image(img, 5)
torch.manual_seed(19920517)
flat = (LBDRNModel(dim_in, 64, 4, 2).flat_params().view(torch.int32) & -65536).view(torch.float32).cuda()
for _ in range(3):
    F.decode_image(scene.msb, flat, 5, 2, 64, 2, flags=fl, path=path, return_tensor=True, base_max=scene.msb_max)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    F.decode_image(scene.msb, flat, 5, 2, 64, 2, flags=fl, path=path, return_tensor=True, base_max=scene.msb_max)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"coords+PE use_colors={use_colors} dim_in={dim_in} side={side} path={path} {ms:.3f} ms  {side*side/ms/1e3:.1f} Mpix/s")
