"""GPU: the decode-loop mirror (effort_model_*, runNetwork.swift:113-209) against the CPU restatement built on
the oracle's bucketMul, on a 2-layer model of the Mistral-7B layer shape."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.ref_decode import RefModel

pytestmark = pytest.mark.gpu


def _cpu(t):
    return t.cpu().numpy()


@pytest.fixture(scope="module")
def small_model():
    import torch
    from effort_b200 import ops
    from effort_b200.model import DecodeModel, MistralConfig
    cfg = MistralConfig(n_layers=2, vocab=2048, max_seq=64)
    m = DecodeModel.random_init(cfg, seed=7, keep_reference_layout=True)
    names = ["wq", "wk", "wv", "wo", "w1", "w2", "w3"]
    layers = []
    for L in m.layers:
        d = {}
        for n, ew in zip(names, L[:7]):
            d[n] = {"buckets": _cpu(ew.buckets), "stats": _cpu(ew.stats), "probes": _cpu(ew.probes), "in": ew.inSize,
                    "out": ew.outSize}
        d["attn_norm"], d["ffn_norm"] = _cpu(L[7]), _cpu(L[8])
        layers.append(d)
    ref = RefModel(layers, _cpu(m.head[0]), _cpu(m.head[1]), _cpu(m.head[2]))
    return m, ref


@pytest.mark.parametrize("use_graph,fused_glue", [(False, False), (True, False), (False, True), (True, True)])
def test_decode_matches_cpu_restatement(small_model, use_graph, fused_glue):
    import torch
    m, ref = small_model
    m.set_graphs(use_graph)
    m.set_fused_glue(fused_glue)
    m.reset()
    ref.pos, ref.kc, ref.vc = 0, [[] for _ in ref.layers], [[] for _ in ref.layers]
    toks = [1, 17, 400, 999, 5]
    for t in toks:
        tok = torch.tensor([t], dtype=torch.int32, device="cuda")
        m.step(tok, effort=0.5)
        torch.cuda.synchronize()
        got = m.logits().cpu().numpy()
        want = ref.step(t, 0.5)
        cs = O.cossim(got, want)
        assert cs > 0.9995, cs     # tiny selection flips (fp32 reorder of v near the cutoff) allowed
        assert m.next_token() == int(np.argmax(got))


def test_step_host_and_self_feeding(small_model):
    import torch
    m, _ = small_model
    m.set_graphs(True)
    m.reset()
    logits = np.zeros(m.cfg.vocab, np.float32)
    t = m.step_host(3, effort=0.25, logits=logits)
    assert t == int(np.argmax(logits))
    seq_a = [t]
    for _ in range(5):
        t = m.step_host(None, effort=0.25)
        seq_a.append(t)
    m.reset()
    t = m.step_host(3, effort=0.25)
    seq_b = [t]
    for _ in range(5):
        t = m.step_host(None, effort=0.25)
        seq_b.append(t)
    assert seq_a == seq_b                                  # deterministic generation (the reference's is not)


def test_model_directory_roundtrip(tmp_path):
    """convert -> bucketed-safetensors directory -> load (C-ABI loader and Python reader) -> decode: the loaded models
    reproduce the in-memory model of the same seed bit for bit (same weights, deterministic kernels); a
    percentLoad-truncated load (loader.swift:113-166) runs and correlates."""
    import torch
    from effort_b200.model import DecodeModel, MistralConfig
    cfg = MistralConfig(n_layers=1, vocab=1024, max_seq=32)
    index = DecodeModel.convert_random_to_directory(str(tmp_path), cfg, seed=11)
    assert index.endswith("buckets-FP16.safetensors.index.json")
    mem = DecodeModel.random_init(cfg, seed=11)
    models = {"native": DecodeModel.from_directory(str(tmp_path), cfg, native=True),
              "python": DecodeModel.from_directory(str(tmp_path), cfg, native=False),
              "load8": DecodeModel.from_directory(str(tmp_path), cfg, percent_load=8)}
    outs = {}
    for name, m in [("mem", mem)] + list(models.items()):
        m.set_graphs(False)
        m.reset()
        seq = []
        for t in (3, 500, 77):
            m.step(torch.tensor([t], dtype=torch.int32, device="cuda"), effort=0.25)
            torch.cuda.synchronize()
            seq.append(m.logits().cpu().numpy())
        outs[name] = seq
    for name in ("native", "python"):
        for a, b in zip(outs["mem"], outs[name]):
            assert np.array_equal(a, b), name
    # percentLoad 8 drops ranks 8..15.  On iid-Gaussian weights those ranks ARE selected at effort 0.25 (the row means
    # fall slowly with rank), so the truncated model only correlates with the full one (measured 0.86 on the first
    # token); real Mistral weights are what the reference's percentLoad knob is for (loader.swift:113-166).
    assert all(np.isfinite(x).all() for x in outs["load8"])
    assert O.cossim(outs["mem"][0], outs["load8"][0]) > 0.7
