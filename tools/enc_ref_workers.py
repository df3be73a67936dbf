"""Reference-sampler encode wall time against the number of concurrent host shuffles. usage: enc_ref_workers.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene_torch
img = make_scene_torch(4, 8192, 8192, 12, device="cuda")
scene = F.DeviceScene.from_image(img, 5)
del img
for w in ("10", "3", "4", "5", "6", "10", "4"):
    os.environ["LBDRN_PERM_WORKERS"] = w
    torch.manual_seed(19920517)
    tr = F.FusedTrainer(LBDRNModel(100, 64, 4, 2), scene, 2, 1e-3, 8192, 10, flags=F.Flags(), sampler="reference")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = tr.run()
    torch.cuda.synchronize()
    print(f"workers {w}: {time.perf_counter() - t0:.3f} s  val {res['val_mse'][-1]:.6f}", flush=True)
    tr.close()
