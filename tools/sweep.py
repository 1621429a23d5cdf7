"""Raw bucketMul GEMV sweep (BASELINE.json configs[4]): shapes x efforts, CUDA-event timing, rotating
weight copies so that every timed call streams from HBM (working set > 2x L2)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from effort_b200 import ops  # noqa: E402
from tools.clocks import ClockSampler  # noqa: E402


def rand_weights(out_dim, in_dim, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = (torch.randn((out_dim, in_dim), generator=g, device="cuda", dtype=torch.float32) * 0.02).half()
    t = ops.bucketize(w)
    return ops.ExpertWeights(t["buckets"], t["bucket.stats"], t["probes"], inDim=in_dim, outDim=out_dim), t


def make_v_gpu(n, seed=42):
    g = torch.Generator(device="cuda").manual_seed(seed)
    v = torch.randn(n, generator=g, device="cuda", dtype=torch.float32)
    idx = torch.randperm(n, generator=g, device="cuda")[: max(1, n // 100)]
    v[idx] *= 10
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4096x4096,4096x1024,4096x11008,4096x14336,14336x4096")
    ap.add_argument("--efforts", default="1.0,0.5,0.25,0.1")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--l2_bytes", type=int, default=300 << 20)
    ap.add_argument("--out", default="")
    ap.add_argument("--warm_s", type=float, default=0.5)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--no-graph", action="store_true", help="time eager launches (host-bound through python)")
    args = ap.parse_args()
    peak = 6570.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    for sh in args.shapes.split(","):
        in_dim, out_dim = (int(x) for x in sh.split("x"))
        mat_bytes = 2 * in_dim * out_dim
        copies = max(2, -(-args.l2_bytes // mat_bytes))
        ws = []
        for c in range(copies):
            ew, t = rand_weights(out_dim, in_dim, 100 + c)
            del t  # the repacked copy is owned by the handle
            ws.append(ew)
        torch.cuda.empty_cache()
        v = make_v_gpu(in_dim)
        out = torch.empty(out_dim, dtype=torch.float32, device="cuda")
        for eff in (float(e) for e in args.efforts.split(",")):
            for k in range(3 * copies):
                ops.bucketMul(v, ws[k % copies], None, out, eff)
            torch.cuda.synchronize()
            nsel = ops.lastSelected()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if args.no_graph:
                s.record()
                for k in range(args.iters):
                    ops.bucketMul(v, ws[k % copies], None, out, eff)
                e.record()
                torch.cuda.synchronize()
                us = s.elapsed_time(e) * 1000 / args.iters
            else:
                # the python->ctypes call costs ~20 us of host time: capture the enqueues once, replay
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=side):
                        for k in range(args.iters):
                            ops.bucketMul(v, ws[k % copies], None, out, eff)
                    import time as _t
                    t0 = _t.time()
                    while _t.time() - t0 < args.warm_s:      # let the SM clock ramp up
                        gr.replay()
                        torch.cuda.synchronize()
                    reps = args.reps
                    with ClockSampler(period=0.01) as cs:
                        s.record(side)
                        for _ in range(reps):
                            gr.replay()
                        e.record(side)
                        torch.cuda.synchronize()
                    clk = cs.summary()
                us = s.elapsed_time(e) * 1000 / (args.iters * reps)
            alg = eff * mat_bytes
            act = nsel * (out_dim // 16) * 2
            row = dict(shape=sh, effort=eff, us=round(us, 2), sel_frac=round(nsel / (in_dim * 16), 4),
                       alg_GBs=round(alg / us / 1e3, 1), act_GBs=round(act / us / 1e3, 1),
                       frac_of_peak=round(alg / us / 1e3 / peak, 3))
            if not args.no_graph:
                row["sm_mhz"] = clk["sm_mhz"]
            rows.append(row)
            print(json.dumps(row), flush=True)
        del ws
        torch.cuda.empty_cache()
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
