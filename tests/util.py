"""Synthetic inputs shared by tests, smoke() and bench.py."""
import numpy as np


def make_w(out_dim: int, in_dim: int, seed: int = 1234, scale: float = 0.02) -> np.ndarray:
    """W [out,in] fp16 ~ N(0, scale^2) (Mistral-like scale)."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((out_dim, in_dim), dtype=np.float32) * scale).astype(np.float16)


def make_v(in_dim: int, seed: int = 42) -> np.ndarray:
    """v [in] fp32 ~ N(0,1) with 1% of entries x10 (post-RMSNorm-like heavy tail)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(in_dim).astype(np.float32)
    idx = rng.choice(in_dim, size=max(1, in_dim // 100), replace=False)
    v[idx] *= 10.0
    return v


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
