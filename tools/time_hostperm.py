"""Time the native host permutation (lbdrn_host_randperm32) for an 8192^2 scene, one and four threads at once.
usage: time_hostperm.py [n]"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lbdrn-msic_b200"))
import torch
import lbdrn_fused as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192 * 8192
for rep in range(2):
    t0 = time.perf_counter()
    p = F.permutation_from_seed(n, 1234 + rep, dtype=torch.int32)
    print(f"one thread: {time.perf_counter() - t0:.3f} s for {n} entries", flush=True)
    del p
outs = [None] * 4
def work(i):
    outs[i] = F.permutation_from_seed(n, 99 + i, dtype=torch.int32)
t0 = time.perf_counter()
th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
[t.start() for t in th]; [t.join() for t in th]
print(f"four threads at once: {time.perf_counter() - t0:.3f} s")
