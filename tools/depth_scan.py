"""How far do the decode logits of the GPU chain drift from the CPU restatement as the model gets deeper?
(effort 0.25 on iid-Gaussian weights is a chaotic regime: row-selection flips compound layer by layer.)
Prints, per depth, cos-sim(GPU, CPU) next to cos-sim(CPU with 8 threads, CPU with 3 threads) -- the same restatement
with a different fp32 summation order -- for the fused chain (2) and the per-op chain (1)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from effort_b200.model import DecodeModel, MistralConfig  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.ref_decode import RefModel  # noqa: E402

O.set_cutoff_mode("select")
effort = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
cfg = MistralConfig(n_layers=16, vocab=2048, max_seq=16)
m_full = DecodeModel.random_init(cfg, seed=21, keep_reference_layout=True)
cpu = lambda t: t.cpu().numpy()
names = ["wq", "wk", "wv", "wo", "w1", "w2", "w3"]
layers = []
for L in m_full.layers:
    d = {n: {"buckets": cpu(ew.buckets), "stats": cpu(ew.stats), "probes": cpu(ew.probes), "in": ew.inSize, "out": ew.outSize}
         for n, ew in zip(names, L[:7])}
    d["attn_norm"], d["ffn_norm"] = cpu(L[7]), cpu(L[8])
    layers.append(d)
head = [cpu(x) for x in m_full.head]
for depth in (1, 2, 4, 8, 16):
    m = DecodeModel(MistralConfig(n_layers=depth, vocab=2048, max_seq=16))
    for i in range(depth):
        m.set_layer(i, *m_full.layers[i])
    m.set_head(*m_full.head)
    row = {"depth": depth}
    for chain in (2, 1):
        m.set_chain(chain)
        m.set_graphs(False)
        m.reset()
        a = RefModel(layers[:depth], *head, fast=True)
        b = RefModel(layers[:depth], *head, fast=True)
        cs_g, cs_c = [], []
        for t in (1, 77):
            m.step(torch.tensor([t], dtype=torch.int32, device="cuda"), effort=effort)
            torch.cuda.synchronize()
            got = m.logits().cpu().numpy()
            O.set_threads(8); wa = a.step(t, effort)
            O.set_threads(3); wb = b.step(t, effort)
            cs_g.append(round(O.cossim(got, wa), 6)); cs_c.append(round(O.cossim(wa, wb), 6))
        row[f"gpu_chain{chain}_vs_cpu"] = cs_g
        row["cpu_vs_cpu"] = cs_c
    print(row, flush=True)
