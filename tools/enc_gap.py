"""Encode wall time of FusedTrainer.run() on the bench scene (device sampler), three times: run-to-run spread.
History: with torch.randperm as the device sampler 1.92-1.96 s, with a cached permutation 1.87 s (tools/enc_gap.py, r1c)."""
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
del img

def once(tag, sampler="device"):
    torch.manual_seed(19920517)
    model = LBDRNModel(100, 64, 4, 2)
    tr = F.FusedTrainer(model, scene, 2, 1e-3, 8192, 10, flags=F.Flags(), sampler=sampler)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = tr.run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tr.close()
    print(f"{tag}: {dt:.3f} s  ({dt / len(res['losses']) * 1e6:.2f} us/step incl eval)  val {res['val_mse'][-1]:.6f}")

for _ in range(3):
    once("device sampler (lbdrn_randperm)")
for _ in range(2):
    once("reference sampler (lbdrn_host_randperm orders = torch.randperm)", "reference")
