"""The bucketed-safetensors container (names, per-layer sharding, index json, percentLoad truncation) round-trips
through effort_b200.weights_io exactly as convert.swift:59-127 / safetensors.swift:38-85 / loader.swift:113-166 lay it out."""
import json
import os

import numpy as np
import pytest

from effort_b200 import weights_io as W
from oracle import oracle as O
from tests.util import make_w


def test_roundtrip_names_index_and_percent_load(tmp_path):
    conv = {k: O.bucketize(make_w(o, 4096, s)) for k, (o, s) in {"wq": (4096, 1), "wk": (1024, 2)}.items()}
    files = []
    for layer in range(2):
        names = W.layer_tensor_names(layer)
        t = {f"layers.{layer}.attention_norm": np.ones(4096, np.float16), f"layers.{layer}.ffn_norm": np.ones(4096, np.float16)}
        for k in ("wq", "wk"):
            for part in ("buckets", "bucket.stats", "probes"):
                t[names[k] + part] = np.ascontiguousarray(conv[k][part])
        if layer == 0:
            t["model.norm"] = np.ones(4096, np.float16)
        files.append(t)
    index = W.save_model(str(tmp_path), files)
    assert os.path.basename(index) == "buckets-FP16.safetensors.index.json"
    wm = json.load(open(index))["weight_map"]
    assert wm["layers.1.attention.wk.bucket.stats"] == "buckets-FP16-00002-of-00002.safetensors"
    assert wm["model.norm"] == "buckets-FP16-00001-of-00002.safetensors"
    L = W.TensorLoader(str(tmp_path))
    assert L.has_tensor("layers.0.attention.wq.probes") and not L.has_tensor("layers.0.attention.wq.core")
    full = L.expert_weights("layers.1.attention.wq.", 4096, 4096)
    assert np.array_equal(full["buckets"].view(np.uint16), np.ascontiguousarray(conv["wq"]["buckets"]).view(np.uint16))
    part = L.expert_weights("layers.0.attention.wk.", 4096, 1024, percent_load=6)
    assert part["buckets"].shape == (6 * 4096, 64) and part["bucket.stats"].shape == (6 * 4096, 4)
    assert np.array_equal(part["buckets"].view(np.uint16), np.ascontiguousarray(conv["wk"]["buckets"]).view(np.uint16)[: 6 * 4096])
    with pytest.raises(KeyError):
        L["layers.7.ffn_norm"]
    with pytest.raises(ValueError):
        L.expert_weights("layers.0.attention.wk.", 4096, 4096)


# ---------------------------------------------------------------------------------------------------------
# the library's own reader (csrc/safetensors_io.cpp, C ABI effort_loader_*): same files, same answers as the
# Python safetensors package; TensorLoader's behaviours (helpers/safetensors.swift:136-216) as return codes
# ---------------------------------------------------------------------------------------------------------
def _write_model(tmp_path):
    rng = np.random.default_rng(3)
    conv = O.bucketize(make_w(1024, 4096, 5))
    files = []
    for layer in range(3):
        t = {f"layers.{layer}.attention_norm": rng.standard_normal(4096).astype(np.float16),
             f"layers.{layer}.attention.wk.buckets": np.ascontiguousarray(conv["buckets"]),
             f"layers.{layer}.attention.wk.bucket.stats": np.ascontiguousarray(conv["bucket.stats"]),
             f"layers.{layer}.attention.wk.probes": np.ascontiguousarray(conv["probes"]),
             f"layers.{layer}.q4.stats": rng.standard_normal((64, 2)).astype(np.float32)}
        if layer == 0:
            t["tok_embeddings.core.weight"] = rng.standard_normal((32, 16)).astype(np.float16)  # reached as "...core"
        files.append(t)
    W.save_model(str(tmp_path), files, description='quotes " and \\ and unicode é中 survive')
    return files


def test_native_loader_matches_python_reader(tmp_path):
    files = _write_model(tmp_path)
    N, P = W.NativeTensorLoader(str(tmp_path)), W.TensorLoader(str(tmp_path))
    assert sorted(N.names()) == sorted(k for f in files for k in f)
    for f in files:
        for name, want in f.items():
            got = N[name]
            assert got.dtype == want.dtype and got.shape == want.shape
            assert np.array_equal(got.view(np.uint8), np.ascontiguousarray(want).view(np.uint8))
            assert np.array_equal(P[name].view(np.uint8), got.view(np.uint8))
    v = N.get("layers.2.attention_norm", copy=False)           # zero-copy view into the mapping
    assert not v.flags.writeable and np.array_equal(v, files[2]["layers.2.attention_norm"])
    # fetchTensor's ".weight" fallback (safetensors.swift:141-146)
    assert N.has_tensor("tok_embeddings.core") and N["tok_embeddings.core"].shape == (32, 16)
    assert not N.has_tensor("layers.9.attention_norm")
    with pytest.raises(KeyError):
        N["layers.9.attention_norm"]
    # ExpertWeights percentLoad prefix (loader.swift:113-166) == the Python reader's
    a = N.expert_weights("layers.1.attention.wk.", 4096, 1024, percent_load=5)
    b = P.expert_weights("layers.1.attention.wk.", 4096, 1024, percent_load=5)
    for k in ("buckets", "bucket.stats", "probes"):
        assert a[k].shape == b[k].shape and np.array_equal(a[k].view(np.uint16), b[k].view(np.uint16))
    with pytest.raises(ValueError):
        N.expert_weights("layers.1.attention.wk.", 4096, 4096)
    N.close()


def test_native_loader_rejects_bad_containers(tmp_path):
    from effort_b200._lib import EffortError
    with pytest.raises(EffortError):                          # no index json
        W.NativeTensorLoader(str(tmp_path))
    _write_model(tmp_path)
    idx = tmp_path / "buckets-FP16.safetensors.index.json"
    wm = json.load(open(idx))
    # a tensor of an unsupported dtype is refused when asked for (safetensors.swift:176), the rest still loads
    from safetensors.numpy import save_file
    save_file({"ints": np.arange(4, dtype=np.int32), "ok": np.ones(3, np.float32)}, str(tmp_path / "extra.safetensors"))
    wm["weight_map"].update({"ints": "extra.safetensors", "ok": "extra.safetensors", "ghost": "missing.safetensors",
                             "trunc": "trunc.safetensors"})
    json.dump(wm, open(idx, "w"))
    raw = open(tmp_path / "extra.safetensors", "rb").read()
    open(tmp_path / "trunc.safetensors", "wb").write(raw[: len(raw) - 7])   # data section shorter than data_offsets say
    N = W.NativeTensorLoader(str(tmp_path))
    assert np.array_equal(N["ok"], np.ones(3, np.float32))
    for bad in ("ints", "ghost", "trunc"):
        with pytest.raises((EffortError, KeyError)):
            N[bad]
    json.dump({"weight_map": ["not", "a", "map"]}, open(idx, "w"))
    with pytest.raises(EffortError):
        W.NativeTensorLoader(str(tmp_path))
    open(idx, "w").write('{"weight_map": {"a": "b"')            # truncated json
    with pytest.raises(EffortError):
        W.NativeTensorLoader(str(tmp_path))


def test_bf16_tensors_come_back_as_fp16(tmp_path):
    """convertBF16 (safetensors.swift:207-210): BF16 on disk -> fp16 in memory, value for value."""
    import struct
    vals = np.array([0.0, -0.0, 1.0, -2.5, 3.140625, 65280.0, 1e-5, 6.1e-5, 7e4, -1e38, np.inf, -np.inf, 2.0 ** -24, 2.0 ** -26],
                    np.float32)
    bf = (vals.view(np.uint32) >> 16).astype(np.uint16)         # exact for these (chosen representable or truncated)
    vals_bf = (bf.astype(np.uint32) << 16).view(np.float32)
    hdr = json.dumps({"x": {"dtype": "BF16", "shape": [len(bf)], "data_offsets": [0, 2 * len(bf)]}}).encode()
    open(tmp_path / "m.safetensors", "wb").write(struct.pack("<Q", len(hdr)) + hdr + bf.tobytes())
    json.dump({"weight_map": {"x": "m.safetensors"}}, open(tmp_path / "m.safetensors.index.json", "w"))
    got = W.NativeTensorLoader(str(tmp_path), model="m")["x"]
    with np.errstate(over="ignore"):
        want = vals_bf.astype(np.float16)                       # numpy: round-to-nearest-even, overflow -> inf
    assert got.dtype == np.float16 and np.array_equal(got.view(np.uint16), want.view(np.uint16))


def test_bf16_to_f16_all_codes():
    import ctypes as C
    from effort_b200 import _lib
    L = _lib.load()
    src = np.arange(65536, dtype=np.uint16)
    dst = np.empty(65536, np.uint16)
    assert L.effort_bf16_to_f16(src.ctypes.data, dst.ctypes.data, src.size) == 0
    f32 = (src.astype(np.uint32) << 16).view(np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        want = f32.astype(np.float16).view(np.uint16)
    nan = np.isnan(f32)
    assert np.array_equal(dst[~nan], want[~nan])
    assert np.all(np.isnan(dst[nan].view(np.float16)))


def test_native_loader_survives_corrupted_headers(tmp_path):
    """Model files come from outside: a damaged header must produce an error code (or a tensor whose bytes lie inside
    the mapping), never a crash.  Mutates the header of a valid file byte-wise and field-wise."""
    import struct
    from effort_b200._lib import EffortError
    rng = np.random.default_rng(7)
    x = rng.standard_normal((8, 4)).astype(np.float16)
    hdr = json.dumps({"x": {"dtype": "F16", "shape": [8, 4], "data_offsets": [0, 64]}, "__metadata__": {"description": "d"}})
    json.dump({"weight_map": {"x": "f.safetensors"}}, open(tmp_path / "m.safetensors.index.json", "w"))

    def write(h: bytes, data=x.tobytes()):
        open(tmp_path / "f.safetensors", "wb").write(struct.pack("<Q", len(h)) + h + data)

    def probe():
        try:
            a = W.NativeTensorLoader(str(tmp_path), model="m")["x"]
            assert a.nbytes <= 64
        except (EffortError, KeyError):
            pass

    write(hdr.encode())
    assert np.array_equal(W.NativeTensorLoader(str(tmp_path), model="m")["x"], x)
    for field in ('"shape": [8, 4]', '"data_offsets": [0, 64]', '"dtype": "F16"'):
        for repl in ('"shape": [8, -4]', '"shape": [1e300, 4]', '"shape": [9007199254740993, 9007199254740993]', '"shape": "x"',
                     '"data_offsets": [64, 0]', '"data_offsets": [0, 1e30]', '"data_offsets": [-8, 56]', '"data_offsets": [0]',
                     '"dtype": "I64"', '"dtype": 16'):
            write(hdr.replace(field, repl).encode())
            probe()
    raw = bytearray(hdr.encode())
    for _ in range(1500):
        m = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            m[int(rng.integers(0, len(m)))] = int(rng.integers(0, 256))
        write(bytes(m))
        probe()
    for cut in range(0, len(raw), 7):            # truncated headers / files
        write(bytes(raw[:cut]))
        probe()
        open(tmp_path / "f.safetensors", "wb").write((struct.pack("<Q", len(raw)) + bytes(raw))[: 8 + cut])
        probe()
    open(tmp_path / "f.safetensors", "wb").write(struct.pack("<Q", 2 ** 63) + bytes(raw))   # absurd header size
    probe()


# ---------------------------------------------------------------------------------------------------------
# writer side: the C-ABI saver (effort_saver_*), convertMistral's name mapping, and the reference loader's name set
# ---------------------------------------------------------------------------------------------------------
def _fake_hf(n_layers, dim=32, hidden=64, kvd=16, vocab=48, seed=9):
    rng = np.random.default_rng(seed)
    t = {"model.norm.weight": rng.standard_normal(dim).astype(np.float16),
         "lm_head.weight": rng.standard_normal((vocab, dim)).astype(np.float16),
         "model.embed_tokens.weight": rng.standard_normal((vocab, dim)).astype(np.float16)}
    for n in range(n_layers):
        p = f"model.layers.{n}."
        for name, shape in (("self_attn.q_proj.weight", (dim, dim)), ("self_attn.k_proj.weight", (kvd, dim)),
                            ("self_attn.v_proj.weight", (kvd, dim)), ("self_attn.o_proj.weight", (dim, dim)),
                            ("mlp.gate_proj.weight", (hidden, dim)), ("mlp.down_proj.weight", (dim, hidden)),
                            ("mlp.up_proj.weight", (hidden, dim))):
            t[p + name] = rng.standard_normal(shape).astype(np.float16)
        t[p + "input_layernorm.weight"] = rng.standard_normal(dim).astype(np.float16)
        t[p + "post_attention_layernorm.weight"] = rng.standard_normal(dim).astype(np.float16)
    return t


def _stub_bucketize(w):
    """shape-faithful stand-in for effort_bucketize (the real one needs in >= 4096 and a GPU)"""
    out, inn = w.shape
    return {"buckets": np.zeros((inn * 16, max(1, out // 16)), np.float16), "bucket.stats": np.zeros((inn * 16, 4), np.float16),
            "probes": np.zeros((4096,), np.float16)}


@pytest.mark.parametrize("native", [True, False])
def test_convert_mistral_writes_what_the_reference_loader_reads(tmp_path, native):
    """convertMistral (convert.swift:59-127) name mapping from the HF checkpoint names, and the name set of the written
    directory against what Model.init / Layer.init / ExpertWeights read (loader.swift:60-166, 201-272): every required
    name present -- including attention `.core`, which the reference's attention loader reads unconditionally -- and
    nothing outside required + optional."""
    hf = _fake_hf(3)
    index = W.convert_mistral(lambda n: hf[n], str(tmp_path), 3, _stub_bucketize, native=native)
    wm = json.load(open(index))["weight_map"]
    req, opt = W.reference_loader_names(3)
    written = set(wm)
    assert req <= written, sorted(req - written)[:5]
    assert written <= req | opt, sorted(written - req - opt)[:5]
    assert wm["output.core"] == wm["tok_embeddings.core"] == "buckets-FP16-00001-of-00003.safetensors"   # convert.swift:70-74
    assert wm["layers.2.feed_forward.experts.0.w2.buckets"] == "buckets-FP16-00003-of-00003.safetensors"
    L = W.TensorLoader(str(tmp_path))
    assert np.array_equal(L["layers.1.attention.wk.core"], hf["model.layers.1.self_attn.k_proj.weight"])
    assert np.array_equal(L["layers.2.ffn_norm"], hf["model.layers.2.post_attention_layernorm.weight"])
    assert np.array_equal(L["layers.0.attention_norm"], hf["model.layers.0.input_layernorm.weight"])
    assert np.array_equal(L["output.core"], hf["lm_head.weight"])
    assert np.array_equal(L["tok_embeddings.core"], hf["model.embed_tokens.weight"])
    # w1 = gate, w2 = down, w3 = up (convert.swift:98-104): the bucket row counts give the input dims away
    assert L["layers.0.feed_forward.experts.0.w2.buckets"].shape[0] == 16 * 64
    assert L["layers.0.feed_forward.experts.0.w1.buckets"].shape[0] == 16 * 32


def test_native_saver_files_match_the_python_package(tmp_path):
    """effort_saver_* (TensorSaver, safetensors.swift:38-85 / 222-280): files readable by the safetensors package and by
    the library's own loader, byte-identical tensors, the reference's header metadata, index json; error codes."""
    import ctypes as C
    from effort_b200 import _lib
    rng = np.random.default_rng(1)
    files = [{"a.x": rng.standard_normal((5, 7)).astype(np.float16), "a.y": rng.standard_normal(11).astype(np.float32)},
             {"b \"quoted\"": rng.standard_normal((2, 3, 4)).astype(np.float16), "b.empty": np.zeros((0, 4), np.float32)}]
    s = W.NativeTensorSaver(str(tmp_path / "m"), "model")
    for i, f in enumerate(files):
        s.add_file(i, f)
    with pytest.raises(TypeError):
        s.add_file(0, {"bad": np.zeros(3, np.int32)})
    Lc = _lib.load()
    shape = (C.c_int64 * 1)(3)
    z = np.zeros(3, np.float16)
    assert Lc.effort_saver_add(s._h, 0, b"a.x", 0, 1, shape, z.ctypes.data, z.nbytes) == -5          # duplicate name
    assert Lc.effort_saver_add(s._h, 0, b"short", 0, 1, shape, z.ctypes.data, 4) == -4               # bytes != shape
    assert Lc.effort_saver_add(s._h, -1, b"neg", 0, 1, shape, z.ctypes.data, z.nbytes) == -1
    index = s.save()
    wm = json.load(open(index))["weight_map"]
    assert wm == {"a.x": "model-00001-of-00002.safetensors", "a.y": "model-00001-of-00002.safetensors",
                  "b \"quoted\"": "model-00002-of-00002.safetensors", "b.empty": "model-00002-of-00002.safetensors"}
    from safetensors import safe_open
    for i, f in enumerate(files):
        path = str(tmp_path / "m" / W.file_name("model", i, 2))
        with safe_open(path, framework="np") as r:
            assert r.metadata() == {"description": "Bucket weights format, see mixtral-kolinko at github"}
            assert set(r.keys()) == set(f)
            for k, a in f.items():
                got = r.get_tensor(k)
                assert got.dtype == a.dtype and got.shape == a.shape and np.array_equal(got, a)
    nl = W.NativeTensorLoader(str(tmp_path / "m"), "model")
    for f in files:
        for k, a in f.items():
            assert np.array_equal(nl[k], a)
    nl.close()
