"""Order-independent digests of half pair lists (shared by the gpu tests and tests/ref_run0.py)."""
import numpy as np


def pair_digest(pl):
    """Order-independent digest of a half pair list of tags: (count, sum and xor of a 64-bit mix of
    (min, max)) and the per-tag degree. Two lists with equal digests hold the same pair set (up to a
    2^-64 collision)."""
    pl = np.asarray(pl, dtype=np.uint32).reshape(-1, 2)
    lo = np.minimum(pl[:, 0], pl[:, 1]).astype(np.uint64)
    hi = np.maximum(pl[:, 0], pl[:, 1]).astype(np.uint64)
    k = (lo << np.uint64(32)) | hi
    # splitmix64 finaliser
    k = (k ^ (k >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    k = (k ^ (k >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    k = k ^ (k >> np.uint64(31))
    return len(pl), int(np.add.reduce(k, dtype=np.uint64)), int(np.bitwise_xor.reduce(k)) if len(k) else 0


def degree(pl, N):
    pl = np.asarray(pl, dtype=np.uint32).reshape(-1, 2)
    return (np.bincount(pl[:, 0], minlength=N) + np.bincount(pl[:, 1], minlength=N)).astype(np.uint32)
