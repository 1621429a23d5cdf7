"""A numpy model of the device's exact-select search (select_cutoff_group, csrc/bucket_mul_v4.cuh) against the oracle's
definition of the select cutoff (oracle_select_cutoff): the quaternary interval search over the 15-bit bf16 key space
returns the (k+1)-th largest probe product for every hint -- the hint only changes the number of passes.  The GPU
tests pin the kernel to the oracle; this pins the ALGORITHM on the CPU side of the suite."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import make_v


def _bf16_bits(x):
    """round-to-nearest-even float32 -> bfloat16 bits (cvt.rn.bf16.f32)"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint32) & 0xFFFF


def _keys(v, probes_f16):
    """bf16(|1e5 * v * bf16(probe)|) as 15-bit keys (sign cleared): findCutoff32's products, bucketMul.metal:166-171"""
    pb = (_bf16_bits(probes_f16.astype(np.float32)) << 16).astype(np.uint32).view(np.float32)
    x = np.abs((np.float32(1e5) * v.astype(np.float32)).astype(np.float32) * pb).astype(np.float32)
    return _bf16_bits(x) & 0x7FFF


def _search(keys, k, hint_key):
    """mirror of select_cutoff_group: interval (L, R) with Q(L) true, Q(R) false, Q(x) = [#{keys > x} >= k+1]"""
    need = k + 1
    L, R = -1, 0x7FFF
    first = 16 < hint_key < 0x7F00
    rounds = 0
    while R - L > 1:
        if first:
            p1, p2, p3 = hint_key - 16, hint_key, hint_key + 16
        else:
            w = R - L
            p1, p2, p3 = L + max(1, w >> 2), L + max(1, w >> 1), L + max(1, (3 * w) >> 2)
            p2, p3 = min(p2, R - 1), min(p3, R - 1)
        first = False
        q1, q2, q3 = [(int((keys > p).sum()) >= need) for p in (p1, p2, p3)]
        if not q1:
            R = p1
        elif not q2:
            L, R = p1, p2
        elif not q3:
            L, R = p2, p3
        else:
            L = p3
        rounds += 1
        assert rounds <= 12
    return (L + 1) << 16, rounds


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("effort", [0.05, 0.25, 0.5, 0.9, 1.0])
def test_search_returns_the_order_statistic_for_every_hint(seed, effort):
    rng = np.random.default_rng(seed)
    v = make_v(4096, seed)
    probes = (rng.standard_normal(4096) * 0.02).astype(np.float16)
    want = np.float32(O.select_cutoff(v, probes.view(np.uint16), effort))
    keys = _keys(v, probes)
    q = O.effort_to_q(effort, 4096)
    k = 4096 - q
    want_key = int(want.view(np.uint32)) >> 16
    for hint in (0, want_key, want_key + 3, max(17, want_key - 9), want_key + 16, want_key + 200, 20, 0x7E00):
        bits, rounds = _search(keys, k, hint)
        got = np.uint32(bits).view(np.float32) if k < 4096 else np.float32(0)
        if k >= 4096:
            continue
        assert got == want, (hint, got, want)
        if hint == 0:
            assert rounds <= 8
        if 16 < want_key and abs(hint - want_key) < 16 and 16 < hint < 0x7F00:
            assert rounds <= 3, (hint, want_key, rounds)      # bracketing pass + two refinements


def test_ties_and_degenerate_inputs():
    probes = np.full(4096, 0.02, np.float16)
    for v in (np.zeros(4096, np.float32), np.ones(4096, np.float32), np.linspace(-1, 1, 4096).astype(np.float32)):
        for effort in (0.1, 0.5):
            want = np.float32(O.select_cutoff(v, probes.view(np.uint16), effort))
            keys = _keys(v, probes)
            k = 4096 - O.effort_to_q(effort, 4096)
            for hint in (0, 300, 0x3F80):
                bits, _ = _search(keys, k, hint)
                assert np.uint32(bits).view(np.float32) == want, (effort, hint)
