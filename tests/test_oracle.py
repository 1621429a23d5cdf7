"""CPU tests: pin the oracle against every known-answer the reference offers for this path.

 * docs/bucketmul.html:50-133   -- the 12x12 / bucket-size-4 layout example (only KAT for the layout)
 * docs/equations.html:262-358  -- the 3x3 effort example (selection semantics)
 * tests/golden/q4_golden_*.npz -- outputs of the reference's own q4_draft.convert (bit exact)
plus structural properties of bucketize / findCutoff32 / bucketMul.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import make_v, make_w, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------------------------------------------------
# scalar formats
# ---------------------------------------------------------------------------------------------------------
def test_half_float_roundtrip_all_codes():
    L = O.lib()
    codes = np.arange(65536, dtype=np.uint16)
    ref = codes.view(np.float16).astype(np.float32)
    got = np.array([L.oracle_half_to_float(int(c)) for c in codes[::7]], dtype=np.float32)
    ok = ~np.isnan(ref[::7])  # NaN payload/quiet-bit conventions differ between converters; not on the path
    np.testing.assert_array_equal(got.view(np.uint32)[ok], ref[::7].view(np.uint32)[ok])
    assert np.all(np.isnan(got[~ok]))


def test_float_to_half_matches_numpy_rne():
    L = O.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([
        rng.standard_normal(4000).astype(np.float32) * 0.02,
        rng.standard_normal(2000).astype(np.float32) * 1e-6,   # subnormal halves
        rng.standard_normal(1000).astype(np.float32) * 3e4,    # near overflow
        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, 6.1e-5, 5.96e-8, 2.98e-8, 2.99e-8], np.float32),
    ])
    with np.errstate(over="ignore"):
        ref = xs.astype(np.float16).view(np.uint16)
    got = np.array([L.oracle_float_to_half(float(x)) for x in xs], dtype=np.uint16)
    np.testing.assert_array_equal(got, ref)


def test_bf16_round_is_rne():
    L = O.lib()
    rng = np.random.default_rng(1)
    xs = (rng.standard_normal(5000) * 1000).astype(np.float32)
    got = np.array([L.oracle_bf16_round(float(x)) for x in xs], dtype=np.float32)
    u = xs.view(np.uint32).astype(np.uint64)
    ref = (((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    np.testing.assert_array_equal(got, ref)
    assert L.oracle_bf16_round(999.0) == 1000.0  # the sentinel quirk of findCutoff32 (bucketMul.metal:156,179)


# ---------------------------------------------------------------------------------------------------------
# KAT 1: docs/bucketmul.html layout example
# ---------------------------------------------------------------------------------------------------------
ROW1 = [.46, .87, -.19, .27, .18, -.39, -.29, -.62, -.81, -.34, -.84, .33]
ROW2 = [-.87, .11, .03, .5, .43, .87, -.49, .59, .5, -.42, -.23, .02]


def _doc_matrix():
    wt = np.zeros((12, 12), dtype=np.float16)  # W^T: row i multiplies v_i
    wt[0], wt[1] = ROW1, ROW2
    return np.ascontiguousarray(wt.T)          # HF layout [out, in]


def _decode(row):
    bits = row.view(np.uint16)
    vals = (bits & 0xFFFC).view(np.float16).astype(np.float64)
    return [(round(float(v), 2), int(b & 3)) for v, b in zip(vals, bits)]


def test_kat_docs_bucket_layout():
    r = O.bucketize(_doc_matrix(), bsize=4, n_probes=None)
    b, s = r["buckets"], r["bucket.stats"]
    assert b.shape == (12 * 4, 12 // 4)  # "[inDim * bSize, outDim / bSize]" docs/bucketmul.html:234
    # expected (value, position) per bucket row; row index = rank*inDim + inIdx.  First input row is given in
    # full by the docs; the second row's annotations are recomputed from the printed matrix (the page prints
    # "-0.42 \searrow 2" where the matrix has -.42 at position 1 of its bucket -- a typo in the page).
    exp = {
        (0, 0): [(.87, 1), (-.62, 3), (-.84, 2)], (0, 1): [(-.87, 0), (.87, 1), (.5, 0)],
        (1, 0): [(.46, 0), (-.39, 1), (-.81, 0)], (1, 1): [(.5, 3), (.59, 3), (-.42, 1)],
        (2, 0): [(.27, 3), (-.29, 2), (-.34, 1)], (2, 1): [(.11, 1), (-.49, 2), (-.23, 2)],
        (3, 0): [(-.19, 2), (.18, 0), (.33, 3)], (3, 1): [(.03, 2), (.43, 0), (.02, 3)],
    }
    for (rank, i), want in exp.items():
        got = _decode(b[rank * 12 + i])
        for (gv, gp), (wv, wp) in zip(got, want):
            assert gp == wp and abs(gv - wv) <= 0.011, (rank, i, got, want)
    # "avg. abs." column of the page (0.777, 0.747, 0.553, 0.503, 0.3, 0.28, 0.233, 0.16)
    want_avg = {(0, 0): .777, (0, 1): .747, (1, 0): .553, (1, 1): .503, (2, 0): .3, (2, 1): .28, (3, 0): .233,
                (3, 1): .16}
    for (rank, i), a in want_avg.items():
        row = s[rank * 12 + i].astype(np.float64)
        assert np.all(row == row[0])                      # half4 with 4 identical lanes, convert.metal:114-117
        assert abs(row[3] - a) < 0.004, (rank, i, row[3], a)


def test_kat_docs_effort_example():
    """docs/equations.html:262-358: v=[1,10,1000], cutoff 100 keeps 5 of 9 products -> [1130,100,1256]."""
    v = np.array([1.0, 10.0, 1000.0])
    # sorted rows (weight, output index) as printed on the page
    rows = [[(256, 2), (8, 1), (2, 0)], [(13, 0), (3, 2), (1, 1)], [(1, 0), (1, 2), (0.1, 1)]]
    out = np.zeros(3)
    kept = 0
    for vi, row in zip(v, rows):
        for wgt, idx in row:
            if vi * wgt >= 100:      # "if el_w > el_cutoff" with cutoff/el_v, inclusive at the 5th product
                out[idx] += vi * wgt
                kept += 1
    assert kept == 5
    np.testing.assert_allclose(out, [1130, 100, 1256])
    dense = np.array([1132.0, 118.0, 1286.0])
    assert abs(O.cossim(out, dense) - 0.99989) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# Q4 converter: bit-exact against the reference's own outputs
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_q4_convert_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"q4_golden_{name}.npz"))
    t = O.q4_convert(g["core2"].view(np.float16))
    np.testing.assert_array_equal(np.ascontiguousarray(t["probes"]).view(np.uint16), g["probes"])
    np.testing.assert_array_equal(t["bucket.stats"], g["bucket_stats"])
    np.testing.assert_array_equal(np.ascontiguousarray(t["buckets"]).view(np.uint16), g["buckets"])
    np.testing.assert_array_equal(t["outliers"], g["outliers"])


@pytest.mark.parametrize("name", ["a", "b"])
def test_q4_mul_full_effort_equals_sign_avg_sum(name):
    """q4_draft.py:203-228 (the live, no-cutoff numpy bucketMul): out[idx] += v * sign(w) * avg."""
    g = np.load(os.path.join(GOLD, f"q4_golden_{name}.npz"))
    core2 = g["core2"].view(np.float16)
    inn, out = core2.shape
    v = g["v"]
    stats, buckets, outl = g["bucket_stats"], g["buckets"], g["outliers"]
    # cutoff below everything -> every row selected
    disp = O.prepare_dispatch_q4(v, stats, -1.0, out // 32, inn * 8)
    assert disp.shape[0] == inn * 8
    o32 = np.zeros(out, np.float32)
    o64 = np.zeros(out, np.float64)
    import ctypes as C
    L = O.lib()
    L.oracle_bucket_mul_q4_dispatch(buckets.ctypes.data_as(C.POINTER(C.c_uint16)),
                                    disp.ctypes.data_as(C.POINTER(C.c_float)), disp.shape[0], out // 32,
                                    o32.ctypes.data_as(C.POINTER(C.c_float)),
                                    o64.ctypes.data_as(C.POINTER(C.c_double)))
    # independent restatement straight from the matrix
    core = core2.copy()
    core[outl[:, 1].astype(int), outl[:, 2].astype(int)] = 0
    want = np.zeros(out)
    gq = core.reshape(inn, out // 8, 8).astype(np.float64)
    order = np.argsort(-np.abs(core.reshape(inn, out // 8, 8)), axis=-1)
    for i in range(inn):
        for rank in range(8):
            avg = float(stats[i * 8 + rank, 0])
            for bkt in range(out // 8):
                pos = order[i, bkt, rank]
                sgn = -1.0 if gq[i, bkt, pos] < 0 else 1.0   # nibble sign bit: value < 0 (q4_draft.py:265)
                want[bkt * 8 + pos] += float(np.float32(v[i]) * np.float32(avg)) * sgn
    np.testing.assert_allclose(o64, want, rtol=1e-6, atol=1e-7)
    assert rel_err(o32, want) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# bucketize properties
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def conv4096():
    w = make_w(4096, 4096, seed=1234)
    return w, O.bucketize(w)


def test_bucketize_is_a_permutation_with_pos_bits(conv4096):
    w, r = conv4096
    b = r["buckets"].view(np.uint16).reshape(16, 4096, 256)       # [rank, in, bucket]
    pos = b & 15
    # every (in, bucket) group: the 16 ranks carry 16 distinct positions
    assert np.array_equal(np.sort(pos, axis=0), np.broadcast_to(np.arange(16, dtype=np.uint16)[:, None, None], pos.shape))
    # value bits (upper 12) equal the source weight at out index bucket*16+pos
    wt = w.view(np.uint16).T                                       # [in, out]
    out_idx = (np.arange(256, dtype=np.int64)[None, None, :] * 16 + pos)
    src = np.take_along_axis(np.broadcast_to(wt[None], (16, 4096, 4096)), out_idx, axis=2)
    assert np.array_equal(b & 0xFFF0, src & 0xFFF0)
    # ranks are |w|-descending inside a group
    a = (src & 0x7FFF).astype(np.int32)
    assert np.all(a[:-1] >= a[1:])


def test_bucketize_stats_and_probes(conv4096):
    w, r = conv4096
    b = r["buckets"]
    s = r["bucket.stats"]
    assert s.shape == (65536, 4) and np.array_equal(s[:, 0], s[:, 3])
    mean = np.abs(b.astype(np.float32)).mean(axis=1)
    assert np.max(np.abs(s[:, 3].astype(np.float32) - mean) / mean) < 2e-3   # fp16 rounding of an fp32 mean
    np.testing.assert_array_equal(r["probes"].view(np.uint16), np.diag(w).view(np.uint16))  # convert.metal:20
    # stats fall with rank for every input dim (what makes low-effort truncation work)
    st = s[:, 3].astype(np.float32).reshape(16, 4096)
    assert np.all(st[:-1] >= st[1:])


def test_bucketize_probe_repeat_small_out():
    """out < 4096: rep = 4096/out probes per row, probes[id*rep+j] = w[id, id+j]  (convert.metal:14-22)."""
    w = make_w(1024, 4096, seed=5)
    r = O.bucketize(w)
    p = r["probes"].view(np.uint16)
    wu = w.view(np.uint16)
    for idx in (0, 1, 17, 1023):
        for j in range(4):
            assert p[idx * 4 + j] == wu[idx, idx + j]
    assert r["buckets"].shape == (4096 * 16, 64)


def test_bucketize_preconditions():
    with pytest.raises(ValueError):
        O.bucketize(make_w(4096, 2048))        # in < 4096: "probes not implemented" convert.swift:212
    with pytest.raises(ValueError):
        O.bucketize(make_w(3000 - 3000 % 16, 4096))  # out < 4096 and 4096 % out != 0


def test_bucketize_tie_rule_lower_index_first():
    w = np.zeros((16, 4096), dtype=np.float16)
    w[:, 0] = [0.5, -0.5, 0.25, 0.5] + [0.1] * 12
    b = O.bucketize(w, n_probes=None)["buckets"].view(np.uint16)
    ranks = b.reshape(16, 4096, 1)[:, 0, 0]
    assert list(ranks[:4] & 15) == [0, 1, 3, 2]          # |0.5| ties: positions 0,1,3 in index order, then 0.25
    assert list(ranks[4:] & 15) == list(range(4, 16))


# ---------------------------------------------------------------------------------------------------------
# cutoff + selection + MAC
# ---------------------------------------------------------------------------------------------------------
def test_find_cutoff_tracks_order_statistic(conv4096):
    w, r = conv4096
    v = make_v(4096)
    vals = np.sort(O.probe_vals(v, r["probes"]))[::-1]
    for effort in (1.0, 0.9, 0.5, 0.25, 0.1, 0.02):
        c, loops = O.find_cutoff(v, r["probes"], effort, return_loops=True)
        k = 4096 - O.effort_to_q(effort)
        above = int(np.sum(vals > c))
        assert 1 <= loops <= 101
        assert abs(above - k) <= 3 + 4096 * 0.002, (effort, above, k)   # bisection stops within 3 counts or 1e-5


def test_effort_zero_and_one_edges(conv4096):
    w, r = conv4096
    v = make_v(4096)
    full = O.bucket_mul(v, r["buckets"], r["bucket.stats"], r["probes"], 4096, 4096, 1.0)
    none = O.bucket_mul(v, r["buckets"], r["bucket.stats"], r["probes"], 4096, 4096, 0.0)
    assert full["n_selected"] > 0.97 * 65536
    assert none["n_selected"] < 0.01 * 65536


def test_selected_fraction_monotone_in_effort(conv4096):
    w, r = conv4096
    v = make_v(4096)
    prev = -1
    for effort in (0.05, 0.1, 0.25, 0.5, 0.75, 1.0):
        n = O.bucket_mul(v, r["buckets"], r["bucket.stats"], r["probes"], 4096, 4096, effort)["n_selected"]
        assert n >= prev
        prev = n


def test_config0_full_effort_vs_dense_cpu_gemv(conv4096):
    """BASELINE.json configs[0]: single 4096x4096 FP16 bucketMul at effort 1.0 vs host-CPU dense GEMV."""
    w, r = conv4096
    v = make_v(4096)
    res = O.bucket_mul(v, r["buckets"], r["bucket.stats"], r["probes"], 4096, 4096, 1.0)
    dense32, dense64 = O.basic_mul(v, w, cast_v=True)
    assert O.cossim(res["out32"], dense64) >= 0.999
    assert rel_err(res["out32"], res["out64"]) < 1e-5


def test_cossim_vs_dense_at_reference_effort_levels(conv4096):
    """docs/ryc/ryc0.3.png (real Mistral weights): ~1.0 down to 40 %, ~0.99 at ~22-25 %.  Synthetic iid Gaussian
    weights are the worst case for the method (no heavy tail to exploit): 0.9946 / 0.9655 / 0.895 here."""
    w, r = conv4096
    v = make_v(4096)
    _, dense64 = O.basic_mul(v, w, cast_v=True)
    cs = {e: O.cossim(O.bucket_mul(v, r["buckets"], r["bucket.stats"], r["probes"], 4096, 4096, e)["out64"], dense64)
          for e in (0.5, 0.25, 0.1)}
    assert cs[0.5] > 0.99 and cs[0.25] > 0.96 and cs[0.5] >= cs[0.25] >= cs[0.1] > 0.85


def test_dispatch_matches_brute_force_selection(conv4096):
    w, r = conv4096
    v = make_v(4096, seed=7)
    c = O.find_cutoff(v, r["probes"], 0.3)
    d = O.prepare_dispatch(v, r["bucket.stats"], c, 4096, 256, 65536)
    st = r["bucket.stats"][:, 3].astype(np.float32)
    rows = np.arange(65536)
    lhs = (np.float32(100000.0) * st) * np.abs(v[rows % 4096])
    want = rows[np.float32(c) < lhs]
    np.testing.assert_array_equal((d[:, 1] / 256).astype(np.int64), want)
    np.testing.assert_array_equal(d[:, 0], v[want % 4096])


def test_basic_mul_casts_v_to_fp16():
    w = make_w(64, 4096, seed=3)
    v = make_v(4096, seed=4)
    o32, o64 = O.basic_mul(v, w, cast_v=True)
    want = w.astype(np.float64) @ v.astype(np.float16).astype(np.float64)
    np.testing.assert_allclose(o64, want, rtol=1e-12)
    assert rel_err(o32, want) < 1e-5
    np.testing.assert_allclose(O.basic_mul_fast(v.astype(np.float16).astype(np.float32), w), want, rtol=2e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------------------
# the derivation the CUDA cutoff relies on (csrc/cutoff.cuh, "why five order statistics are enough"): the literal bisection only ever
# needs the capped count (k-3) + sum_{j=k-2..k+2} [T_j > b] of five order statistics
# ---------------------------------------------------------------------------------------------------------
def _direct_cutoff(vals, k):
    f32 = np.float32
    bf = lambda x: f32(O.lib().oracle_bf16_round(float(x)))
    srt = np.sort(vals)[::-1]
    T = [f32(np.inf) if r < 1 else (f32(-1) if r > len(vals) else srt[r - 1]) for r in range(k - 2, k + 3)]
    mn, mx = bf(min(f32(999), vals.min())), bf(max(f32(-999), vals.max()))
    nb, loops, min_c, max_c = f32((mn + mx) / f32(2)), 0, min(4096, k + 2), max(0, k - 3)
    while True:
        loops += 1
        c = (k - 3) + sum(1 for t in T if t > nb)
        if c < k:
            mx, max_c = nb, c
        else:
            mn, min_c = nb, c
        prev, nb = nb, f32((mx + mn) / f32(2))
        if c == k or f32(mx - mn) < f32(0.00001) or abs(max_c - min_c) < 3 or loops > 100:
            return nb, loops
        if nb == prev:
            return nb, 101


def test_direct_order_statistic_replay_equals_literal_bisection(conv4096):
    w, r = conv4096
    rng = np.random.default_rng(0)
    cases = [make_v(4096, s) for s in range(6)] + [
        np.zeros(4096, np.float32), np.ones(4096, np.float32), make_v(4096, 3) * 1e4, make_v(4096, 4) * 1e-6,
        np.eye(1, 4096, 17, dtype=np.float32)[0] * 5, np.round(make_v(4096, 9)), rng.integers(0, 3, 4096).astype(np.float32)]
    with np.errstate(all="ignore"):
        for v in cases:
            vals = O.probe_vals(v, r["probes"])
            for e in (1.0, 0.999, 0.9, 0.5, 0.25, 0.1, 0.001, 0.0):
                c, loops = O.find_cutoff(v, r["probes"], e, return_loops=True)
                c2, loops2 = _direct_cutoff(vals, 4096 - O.effort_to_q(e))
                assert np.float32(c).view(np.uint32) == np.float32(c2).view(np.uint32) and loops == loops2, (e, c, c2)


# ---------------------------------------------------------------------------------------------------------
# the path the CUDA kernels actually run (csrc/cutoff.cuh, group_cutoff): literal iterations while more than 128
# products lie inside the bracket (counted against bf16-TRUNCATED thresholds, two products per register), then
# the five order statistics are read from the rank-sorted bracket and one thread replays the rest of the loop in
# the capped domain, continuing from the state phase A left
# ---------------------------------------------------------------------------------------------------------
def _trunc_bf16(t):
    return (np.float32(t).view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def _group_cutoff(vals, k, inside_max=128):
    f32 = np.float32
    bf = lambda x: f32(O.lib().oracle_bf16_round(float(x)))
    good = vals[vals >= 0]
    mn = bf(min(f32(999), good.min())) if len(good) else bf(f32(999))
    mx = bf(max(f32(-999), good.max())) if len(good) else bf(f32(-999))
    nb, loops, min_c, max_c = f32((mn + mx) / f32(2)), 0, 4096, 0

    def step(c):
        nonlocal mn, mx, nb, loops, min_c, max_c
        if c < k:
            mx, max_c = nb, c
        else:
            mn, min_c = nb, c
        prev, nb = nb, f32((mx + mn) / f32(2))
        if c == k or f32(mx - mn) < f32(0.00001) or abs(max_c - min_c) < 3 or loops > 100:
            return True
        if nb == prev:
            loops = 101
            return True
        return False

    while (min_c - max_c) > inside_max:          # phase A
        loops += 1
        c_exact = int(np.count_nonzero(vals > nb))
        c = int(np.count_nonzero(vals > _trunc_bf16(nb)))   # what HSET2.BF16 computes
        assert c == c_exact
        if step(c):
            return nb, loops
    inside = np.sort(vals[(vals > mn) & (vals <= mx)])[::-1]
    assert len(inside) <= inside_max
    above = max_c
    assert above == int(np.count_nonzero(vals > mx))
    T = []
    for r in range(k - 2, k + 3):
        idx = r - above - 1
        T.append(f32(np.inf) if idx < 0 else (f32(-1) if idx >= len(inside) else inside[idx]))
    min_c, max_c = min(min_c, k + 2), max(max_c, k - 3)
    while True:                                   # scalar replay
        loops += 1
        assert (not (T[2] > nb)) == ((k - 3) + sum(1 for t in T if t > nb) < k)   # the one compare on the critical path
        if step((k - 3) + sum(1 for t in T if t > nb)):
            return nb, loops


def test_group_cutoff_path_equals_literal_bisection(conv4096):
    w, r = conv4096
    rng = np.random.default_rng(1)
    cases = [make_v(4096, s) for s in range(6)] + [
        np.zeros(4096, np.float32), np.ones(4096, np.float32), make_v(4096, 3) * 1e4, make_v(4096, 4) * 1e-6,
        np.eye(1, 4096, 17, dtype=np.float32)[0] * 5, np.round(make_v(4096, 9)), rng.integers(0, 3, 4096).astype(np.float32),
        np.where(rng.random(4096) < 0.9, 0, make_v(4096, 11)).astype(np.float32)]
    with np.errstate(all="ignore"):
        for v in cases:
            vals = O.probe_vals(v, r["probes"])
            for e in (1.0, 0.999, 0.9, 0.5, 0.25, 0.1, 0.001, 0.0):
                c, loops = O.find_cutoff(v, r["probes"], e, return_loops=True)
                for inside_max in (128, 16):
                    c2, loops2 = _group_cutoff(vals, 4096 - O.effort_to_q(e), inside_max)
                    assert np.float32(c).view(np.uint32) == np.float32(c2).view(np.uint32) and loops == loops2, (e, c, c2, loops, loops2)


def test_multithreaded_cpu_arm_equals_the_single_thread_oracle():
    """bench.py's CPU arm (oracle_bucket_mul_mt: parallel dispatch, row-chunked MAC into private accumulators) selects
    the same rows and sums to the same vector as the literal single-thread restatement, for any thread count."""
    for (i, o) in [(4096, 1024), (4096, 4096)]:
        r = O.bucketize(make_w(o, i, 5))
        v = make_v(i, 42)
        for eff in (0.25, 1.0):
            ref = O.bucket_mul(v, r["buckets"], r["bucket.stats"], r["probes"], i, o, eff)
            before = O.num_threads()
            try:
                for nt in (1, 3, 8):
                    O.set_threads(nt)
                    out, n = O.bucket_mul_mt(v, r["buckets"], r["bucket.stats"], r["probes"], i, o, eff)
                    assert n == ref["n_selected"]
                    assert rel_err(out, ref["out64"]) < 1e-5
            finally:
                O.set_threads(before)
