"""The bench's config-3 encode leg alone (two epochs of D=3 bc256 incl. evaluation), three times on one scene, next to the
per-step kernel timing of tools/time_train.py: separates first-run costs from the steady state.  usage: enc3_leg.py [side]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene_torch
side = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
img = make_scene_torch(4, side, side, 12, device="cuda")
scene = F.DeviceScene.from_image(img, 5)
for rep in range(3):
    torch.manual_seed(19920517)
    tr = F.FusedTrainer(LBDRNModel(4 * 49, 256, 4, 2), scene, 3, 1e-3, 8192, 2, flags=F.Flags(), sampler="device")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = tr.run()
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    tr.close()
    print(f"rep {rep}: {s:.3f} s per 2 epochs = {s / len(r['losses']) * 1e6:.1f} us/step incl. evaluation, val mse {r['val_mse'][-1]:.6f}")
