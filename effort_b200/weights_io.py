"""Bucketed-safetensors model files: the on-disk contract of the reference, unchanged.

    writer  TensorSaver.save            helpers/safetensors.swift:38-85   one file per layer,
            "<model>-%05d-of-%05d.safetensors" + "<model>.safetensors.index.json" {"weight_map": {name: file}}
    names   convertMistral              convert.swift:59-127
            layers.N.attention.w{q,k,v,o}.{buckets,bucket.stats,probes,core}
            layers.N.feed_forward.experts.0.w{1,2,3}.{buckets,bucket.stats,probes}
            layers.N.{attention_norm,ffn_norm};  layer 0 also model.norm, output.core, tok_embeddings.core
    reader  ExpertWeights(prefix, wId, inDim, outDim, numExperts, percentLoad)   loader.swift:113-166: only the first
            percentLoad*inDim bucket rows / stats rows are kept (model.swift:143-146 copies countBytes of the prefix)

Host-side I/O only (numpy / safetensors); tensors go to the device through ops.ExpertWeights."""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np
from safetensors import safe_open
from safetensors.numpy import save_file

MODEL_FP16 = "buckets-FP16"  # convert.swift:61


def layer_tensor_names(layer: int) -> Dict[str, str]:
    """projection key -> tensor-name prefix (convert.swift:86-106)."""
    names = {f"w{s}": f"layers.{layer}.attention.w{s}." for s in "qkvo"}
    names.update({f"w{i}": f"layers.{layer}.feed_forward.experts.0.w{i}." for i in (1, 2, 3)})
    return names


# HF checkpoint names -> the reference's names (convertMistral, convert.swift:59-127)
HF_HEAD = {"model.norm": "model.norm.weight", "output.core": "lm_head.weight", "tok_embeddings.core": "model.embed_tokens.weight"}
HF_ATTN = {"wq": "self_attn.q_proj.weight", "wk": "self_attn.k_proj.weight", "wv": "self_attn.v_proj.weight",
           "wo": "self_attn.o_proj.weight"}
HF_MLP = {"w1": "mlp.gate_proj.weight", "w2": "mlp.down_proj.weight", "w3": "mlp.up_proj.weight"}   # convert.swift:98-104
HF_NORMS = {"attention_norm": "input_layernorm.weight", "ffn_norm": "post_attention_layernorm.weight"}


def reference_loader_names(n_layers: int, n_experts: int = 1):
    """The tensor names the reference's loader reads for a non-quantised model: (required, optional).
    Model.init (loader.swift:254-272), Layer.init (:201-225), ExpertWeights(elName:) for attention (:60-110: `.core` is
    read unconditionally, buckets only if `.probes` exists) and ExpertWeights(prefix, wId, ...) for the MLP (:113-166)."""
    req = {"model.norm", "output.core", "tok_embeddings.core"}
    opt = set()
    for n in range(n_layers):
        req |= {f"layers.{n}.ffn_norm", f"layers.{n}.attention_norm"}
        if n_experts > 1:
            req.add(f"layers.{n}.feed_forward.gate")
        for s_ in "okqv":
            el = f"layers.{n}.attention.w{s_}"
            req.add(el + ".core")
            opt |= {el + ".outliers", el + ".probes", el + ".buckets", el + ".bucket.stats"}
        for e in range(n_experts):
            for wid in ("w1", "w3", "w2"):
                f = f"layers.{n}.feed_forward.experts.{e}.{wid}."
                req |= {f + "probes", f + "buckets", f + "bucket.stats"}
                opt |= {f + "core"} | ({f + "outliers"} if e == 0 else set())
    return req, opt


def convert_mistral(source, path: str, n_layers: int, bucketize, model: str = MODEL_FP16, native: bool = True) -> str:
    """convertMistral (convert.swift:59-127): HF-named tensors in, the reference's bucketed model directory out.
    source: callable HF tensor name -> numpy float16 array ([out, in] for matrices); bucketize: [out,in] fp16 array ->
    {"buckets", "bucket.stats", "probes"} numpy arrays (the library's effort_bucketize on the GPU).  One file per layer;
    layer 0 also carries model.norm / output.core / tok_embeddings.core (:70-74); every attention projection keeps its
    dense `.core` next to the buckets (:91 -- the reference's attention loader requires it, loader.swift:60-63)."""
    saver = (NativeTensorSaver if native else PythonTensorSaver)(path, model)
    for n in range(n_layers):
        t = {}
        if n == 0:
            for new, old in HF_HEAD.items():
                t[new] = source(old)
        for new, old in HF_NORMS.items():
            t[f"layers.{n}.{new}"] = source(f"model.layers.{n}.{old}")
        for s_ in ("k", "o", "q", "v"):                                        # the reference's order, :86
            w = source(f"model.layers.{n}." + HF_ATTN["w" + s_])
            pref = f"layers.{n}.attention.w{s_}."
            for k, a in bucketize(w).items():
                t[pref + k] = a
            t[pref + "core"] = w
        for wid in ("w1", "w2", "w3"):
            w = source(f"model.layers.{n}." + HF_MLP[wid])
            pref = f"layers.{n}.feed_forward.experts.0.{wid}."
            for k, a in bucketize(w).items():
                t[pref + k] = a
        saver.add_file(n, t)
    return saver.save()


class PythonTensorSaver:
    """TensorSaver through the Python safetensors package (kept as the independent implementation for tests)."""

    def __init__(self, path: str, model: str = MODEL_FP16):
        self.path, self.model, self.files = path, model, {}

    def add_file(self, idx: int, tensors: dict):
        self.files[idx] = tensors

    def save(self) -> str:
        return save_model(self.path, [self.files[i] for i in range(len(self.files))], self.model,
                          "Bucket weights format, see mixtral-kolinko at github")


class NativeTensorSaver:
    """TensorSaver over the library's C-ABI writer (effort_saver_*, csrc/safetensors_io.cpp).  The number of files is
    part of every file name ("-%05d-of-%05d", safetensors.swift:67), so -- like the reference -- the saver collects the
    tensors of all layers and writes the shards and the index in save()."""

    def __init__(self, path: str, model: str = MODEL_FP16):
        import ctypes as C
        from . import _lib
        self._C, self._lib, self._L = C, _lib, _lib.load()
        self.path, self.model = path, model
        h = C.c_void_p()
        _lib.check(self._L.effort_saver_open(path.encode(), model.encode(), None, C.byref(h)), "effort_saver_open")
        self._h, self._keep = h, []

    _CODE = {np.dtype(np.float16): 0, np.dtype(np.float32): 2}

    def add_file(self, idx: int, tensors: dict):
        for name, a in tensors.items():
            a = np.ascontiguousarray(a)
            if a.dtype not in self._CODE:
                raise TypeError(f"{name}: only float16 / float32 tensors can be saved (safetensors.swift:231-247), got {a.dtype}")
            self._keep.append(a)   # the saver does not copy
            shape = (self._C.c_int64 * max(1, a.ndim))(*a.shape)
            self._lib.check(self._L.effort_saver_add(self._h, idx, name.encode(), self._CODE[a.dtype], a.ndim, shape,
                                                     a.ctypes.data, a.nbytes), f"effort_saver_add({name})")

    def save(self) -> str:
        self._lib.check(self._L.effort_saver_save(self._h), "effort_saver_save")
        self._L.effort_saver_close(self._h)
        self._h, self._keep = None, []
        return os.path.join(self.path, f"{self.model}.safetensors.index.json")


def file_name(model: str, idx: int, count: int) -> str:
    return f"{model}-{idx + 1:05d}-of-{count:05d}.safetensors"  # safetensors.swift:67


def save_model(path: str, files: list, model: str = MODEL_FP16, description: str = "effort bucketed weights") -> str:
    """files: list (one entry per layer) of {tensor name: numpy array}.  Returns the index path."""
    os.makedirs(path, exist_ok=True)
    weight_map = {}
    for i, tensors in enumerate(files):
        fname = file_name(model, i, len(files))
        save_file({k: np.ascontiguousarray(v) for k, v in tensors.items()}, os.path.join(path, fname),
                  metadata={"description": description})
        for k in tensors:
            weight_map[k] = fname
    index = os.path.join(path, f"{model}.safetensors.index.json")
    with open(index, "w") as f:
        json.dump({"weight_map": weight_map}, f, indent=2)
    return index


class TensorLoader:
    """TensorLoader (helpers/safetensors.swift:87-216): name -> tensor through the index, no caching."""

    def __init__(self, path: str, model: str = MODEL_FP16):
        self.path = path
        with open(os.path.join(path, f"{model}.safetensors.index.json")) as f:
            self.weight_map = json.load(f)["weight_map"]

    def has_tensor(self, name: str) -> bool:
        return name in self.weight_map

    def __getitem__(self, name: str) -> np.ndarray:
        if name not in self.weight_map:
            raise KeyError(f"Tensor not found in library: {name}")  # safetensors.swift:32
        with safe_open(os.path.join(self.path, self.weight_map[name]), framework="np") as f:
            return f.get_tensor(name)

    def expert_weights(self, prefix: str, in_dim: int, out_dim: int, percent_load: int = 16) -> dict:
        """The tensors of ExpertWeights(prefix...) truncated to the first percent_load ranks (loader.swift:113-166)."""
        rows = percent_load * in_dim
        b, s = self[prefix + "buckets"], self[prefix + "bucket.stats"]
        if b.shape != (16 * in_dim, out_dim // 16) or s.shape[0] != 16 * in_dim:
            raise ValueError(f"{prefix}: shapes {b.shape} / {s.shape} do not match in={in_dim} out={out_dim}")
        return {"buckets": np.ascontiguousarray(b[:rows]), "bucket.stats": np.ascontiguousarray(s[:rows]),
                "probes": self[prefix + "probes"], "percentLoad": percent_load}


class NativeTensorLoader:
    """The same interface over the library's own reader (csrc/safetensors_io.cpp: `effort_loader_*`, plain C ABI, no
    Python safetensors package): index json + per-file headers parsed in C++, files mmapped read-only.  Arrays are
    copies by default; `copy=False` returns read-only views into the mapping that are valid until `close()`."""

    _DT = {0: np.float16, 1: np.uint16, 2: np.float32}  # EFFORT_ST_F16 / BF16 (raw codes) / F32

    def __init__(self, path: str, model: str = MODEL_FP16):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.effort_loader_open(path.encode(), model.encode(), C.byref(h)), f"effort_loader_open({path}, {model})")
        self._h = h

    def close(self):
        if self._h:
            self._L.effort_loader_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def names(self):
        return [self._L.effort_loader_name(self._h, i).decode() for i in range(self._L.effort_loader_count(self._h))]

    def has_tensor(self, name: str) -> bool:
        return bool(self._L.effort_loader_has(self._h, name.encode()))

    def info(self, name: str):
        ti = self._lib.TensorInfo()
        rc = self._L.effort_loader_tensor(self._h, name.encode(), self._C.byref(ti))
        if rc == -6:  # EFFORT_ENOTLOADED: "not found in the safetensors lib" (safetensors.swift:148)
            raise KeyError(f"Tensor not found in library: {name}")
        self._lib.check(rc, f"effort_loader_tensor({name})")
        return ti

    def get(self, name: str, copy: bool = True, rows: int = None) -> np.ndarray:
        """rows: keep only the first `rows` entries of the leading dimension (the percentLoad prefix) -- only those
        bytes are touched."""
        ti = self.info(name)
        shape = [int(ti.shape[d]) for d in range(ti.ndim)]
        nbytes = int(ti.nbytes)
        if rows is not None and shape:
            rows = min(rows, shape[0])
            nbytes = nbytes // max(1, shape[0]) * rows
            shape[0] = rows
        buf = (self._C.c_char * nbytes).from_address(ti.data) if nbytes else b""
        a = np.frombuffer(buf, dtype=self._DT[ti.dtype]).reshape(shape)
        if ti.dtype == 1:  # BF16 -> fp16 like the reference does after loading (safetensors.swift:207-210)
            out = np.empty(a.shape, np.float16)
            self._lib.check(self._L.effort_bf16_to_f16(a.ctypes.data, out.ctypes.data, a.size), "effort_bf16_to_f16")
            return out
        if copy:
            return a.copy()
        a.flags.writeable = False
        return a

    __getitem__ = get

    def expert_weights(self, prefix: str, in_dim: int, out_dim: int, percent_load: int = 16) -> dict:
        """ExpertWeights(prefix...) truncated to the first percent_load ranks (loader.swift:113-166); only the kept
        prefix of the mapping is read."""
        rows = percent_load * in_dim
        bi, si = self.info(prefix + "buckets"), self.info(prefix + "bucket.stats")
        if (bi.ndim != 2 or (bi.shape[0], bi.shape[1]) != (16 * in_dim, out_dim // 16)) or si.shape[0] != 16 * in_dim:
            raise ValueError(f"{prefix}: shapes do not match in={in_dim} out={out_dim}")
        return {"buckets": self.get(prefix + "buckets", rows=rows), "bucket.stats": self.get(prefix + "bucket.stats", rows=rows),
                "probes": self.get(prefix + "probes"), "percentLoad": percent_load}
