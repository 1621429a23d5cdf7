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
