"""Time the decode kernel alone (CUDA events) for a few settings. usage: time_decode.py side path [reps] [D] [bc]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene_torch
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
path = sys.argv[2] if len(sys.argv) > 2 else "auto"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
D = int(sys.argv[4]) if len(sys.argv) > 4 else 2
bc = int(sys.argv[5]) if len(sys.argv) > 5 else 64
dim_in = 4 * (2 * D + 1) ** 2
img = make_scene_torch(4, side, side, 12, device="cuda")
scene = F.DeviceScene.from_image(img, 5)
torch.manual_seed(19920517)
flat = (LBDRNModel(dim_in, bc, 4, 2).flat_params().view(torch.int32) & -65536).view(torch.float32).cuda()
for _ in range(3):
    F.decode_image(scene.msb, flat, 5, D, bc, 2, flags=F.Flags(), path=path, return_tensor=True, base_max=scene.msb_max)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    F.decode_image(scene.msb, flat, 5, D, bc, 2, flags=F.Flags(), path=path, return_tensor=True, base_max=scene.msb_max)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flop = 2 * (dim_in * bc + bc * bc + bc * 4)
print(f"side={side} D={D} bc={bc} path={path} {side*side*flop/ms/1e9:.1f} TFLOP/s occ={os.environ.get('LBDRN_TC_OCC')} {ms:.3f} ms  {side*side/ms/1e3:.1f} Mpix/s")
