"""A/B of one bucketMul operator between library builds (EFFORT_LIB=<path> selects an older .so)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from effort_b200 import ops  # noqa: E402
from tools.sweep import rand_weights, make_v_gpu  # noqa: E402

CASES = [("4096x1024", 0.25), ("4096x4096", 0.25), ("4096x14336", 0.25), ("4096x14336", 1.0)]
if os.environ.get("AB_CASES"):   # e.g. AB_CASES=4096x14336:1.0,4096x14336:0.25
    CASES = [(c.split(":")[0], float(c.split(":")[1])) for c in os.environ["AB_CASES"].split(",")]
for sh, eff in CASES:
    in_dim, out_dim = (int(x) for x in sh.split("x"))
    ws = [rand_weights(out_dim, in_dim, 100 + c)[0] for c in range(max(3, 400_000_000 // (2 * in_dim * out_dim)))]
    v = make_v_gpu(in_dim)
    out = torch.empty(out_dim, dtype=torch.float32, device="cuda")
    for w in ws[:3]:
        ops.bucketMul(v, w, None, out, eff)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for w in ws:
                ops.bucketMul(v, w, None, out, eff)
        for _ in range(20):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(50):
            g.replay()
        e1.record(side)
        torch.cuda.synchronize()
    print(f"{os.environ.get('EFFORT_LIB', 'current')[-30:]:>30s} {sh} effort {eff}: {e0.elapsed_time(e1) * 1e3 / (50 * len(ws)):.2f} us/op", flush=True)
    del ws
    torch.cuda.empty_cache()
