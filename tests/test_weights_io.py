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
