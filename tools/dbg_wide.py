"""Debug driver: one wide-kernel decode of a tiny scene against the oracle. usage: dbg_wide.py [H W] [bc D]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lbdrn-msic_b200", "oracle", "oracle/shims"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import lbdrn_fused as F
import lbdrn_oracle as O
from LBDRNmodel import LBDRNModel
from synth_scene import make_scene
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
bc = int(sys.argv[3]) if len(sys.argv) > 3 else 256
D = int(sys.argv[4]) if len(sys.argv) > 4 else 3
C, K, nl = 4, 5, 2
dim_in = C * (2 * D + 1) ** 2
img = make_scene(C, H, W, 12, seed=5)
msb, _ = O.split_msb_lsb(img, K)
torch.manual_seed(5)
flat = LBDRNModel(dim_in, bc, C, nl).flat_params()
flat = (flat.view(torch.int32) & -65536).view(torch.float32).numpy()
out = F.decode_image(msb, flat, K, D, bc, nl, flags=F.Flags(), path=os.environ.get("LBDRN_PATH", "tensor"))
torch.cuda.synchronize()
ref = O.decode_image(msb, O.unflatten_params(flat, dim_in, bc, C, nl), K, D)
diff = np.abs(out.astype(np.int64) - ref.astype(np.int64))
print(f"H={H} W={W} bc={bc} D={D}: max diff {diff.max()}, {(diff != 0).sum()}/{diff.size} differ")
if diff.max() > 1:
    bad = np.argwhere(diff > 1)
    print("first bad (c,y,x):", bad[:10].tolist())
    print("bad per band:", [(diff[c] > 1).sum() for c in range(C)])
    ys, xs = np.unique(bad[:, 1]), np.unique(bad[:, 2])
    print("bad rows:", ys[:40].tolist(), "bad cols:", xs[:40].tolist())
