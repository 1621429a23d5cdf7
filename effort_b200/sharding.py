"""Tensor-parallel sharding of bucketed weights.  The reference is single-device
(helpers/gpu.swift:36-38); the sharding below is new work that keeps the selection of the unsharded operator:

* column-parallel (split `out`: wq, wk, wv, w1, w3): rank g keeps the bucket COLUMNS [C*g/G, C*(g+1)/G) of every
  row (output groups of 16 stay intact), the full stats and probes and the full `v`.  Cutoff and selection are
  computed redundantly and identically on every rank => the concatenation of the shards' outputs equals the
  unsharded output bit for bit in the selected set (fp32 order per output is even unchanged).
* row-parallel (split `in`: wo, w2): rank g keeps the bucket ROWS whose input index lies in its slice
  (rows rank*in + i of the rank-major matrix), the matching stats, all probes, and the local slice of `v`.
  The cutoff of the reference probes input dims 0..4095 only (bucketMul.metal:158-163), so every rank needs the
  first 4096 entries of the FULL v (`v_cutoff`): an all-gather of 16 KB; the partial outputs are summed with one
  all-reduce (the NCCL all-reduce BASELINE.json names on down_proj; wo needs the same).

Works on numpy arrays or torch tensors (anything with reshape/slicing); shapes are the reference layout
(buckets [in*16, out/16], stats [in*16, 4], probes [4096])."""
from __future__ import annotations

RANKS = 16  # bucket size of the FP16 format (convert.swift:233)


def _contig(x):
    return x.contiguous() if hasattr(x, "contiguous") else __import__("numpy").ascontiguousarray(x)


def shard_columns(t: dict, in_dim: int, out_dim: int, rank: int, world: int) -> dict:
    C = out_dim // 16
    if C % world:
        raise ValueError(f"out/16 = {C} not divisible by {world}")
    c0, c1 = C * rank // world, C * (rank + 1) // world
    return {"buckets": _contig(t["buckets"].reshape(RANKS * in_dim, C)[:, c0:c1]),
            "bucket.stats": t["bucket.stats"], "probes": t["probes"],
            "in": in_dim, "out": out_dim // world, "out_offset": 16 * c0}


def shard_rows(t: dict, in_dim: int, out_dim: int, rank: int, world: int) -> dict:
    if in_dim % world:
        raise ValueError(f"in = {in_dim} not divisible by {world}")
    C = out_dim // 16
    i0, i1 = in_dim * rank // world, in_dim * (rank + 1) // world
    b = t["buckets"].reshape(RANKS, in_dim, C)[:, i0:i1, :].reshape(RANKS * (i1 - i0), C)
    s = t["bucket.stats"].reshape(RANKS, in_dim, 4)[:, i0:i1, :].reshape(RANKS * (i1 - i0), 4)
    return {"buckets": _contig(b), "bucket.stats": _contig(s), "probes": t["probes"],
            "in": i1 - i0, "out": out_dim, "in_offset": i0}


def mistral_layer_plan():
    """Megatron-style plan for one Mistral layer (runNetwork.swift:132-183): which projections split which way and
    where the two exchanges sit.  Heads (32 q / 8 kv) are split with the columns of wq/wk/wv."""
    return {"wq": "column", "wk": "column", "wv": "column", "wo": "row", "w1": "column", "w3": "column", "w2": "row",
            "exchanges": ["all_gather(attn_out[:4096]) -> cutoff of wo; all_reduce(wo out)",
                          "all_gather(x2[:4096]) -> cutoff of w2; all_reduce(w2 out)"]}
