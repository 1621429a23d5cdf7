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
