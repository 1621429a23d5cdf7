"""One shape/effort, N eager calls after a warm-up: target for ncu captures."""
import argparse, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from effort_b200 import ops  # noqa: E402
from tools.sweep import rand_weights, make_v_gpu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="4096x14336")
ap.add_argument("--effort", type=float, default=0.25)
ap.add_argument("--n", type=int, default=6)
ap.add_argument("--copies", type=int, default=3)
a = ap.parse_args()
in_dim, out_dim = (int(x) for x in a.shape.split("x"))
ws = [rand_weights(out_dim, in_dim, 100 + c)[0] for c in range(a.copies)]
v = make_v_gpu(in_dim)
out = torch.empty(out_dim, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for k in range(a.n):
    ops.bucketMul(v, ws[k % a.copies], None, out, a.effort)
torch.cuda.synchronize()
print("selected", ops.lastSelected(), "of", in_dim * 16)
