"""Seeded synthetic indices shared by the oracle and the CUDA path (test helpers; uses oracle/ builders)."""
import numpy as np

import oracle as O


def clustered(n, dim, n_blobs=32, sigma=0.15, seed=1234):
    """Gaussian blobs in [0,1]^dim (py/create_test_hdf5.py-style data; uniform 768-d makes IVF recall meaningless)."""
    rng = np.random.default_rng(seed)
    centers = rng.random((n_blobs, dim), dtype=np.float32)
    lab = rng.integers(0, n_blobs, n)
    return (centers[lab] + sigma * rng.standard_normal((n, dim)).astype(np.float32)).astype(np.float32)


def uniform(n, dim, seed=1234):
    """rs/utils/src/test_utils.rs:6-9 generate_random_vector: i.i.d. uniform[0,1) f32."""
    return np.random.default_rng(seed).random((n, dim), dtype=np.float32)


def build_ivf_arrays(X, nlist, seed=7, iters=8, max_clusters=1, threshold=0.1, sample=None):
    rng = np.random.default_rng(seed)
    S = X if sample is None or sample >= len(X) else X[rng.choice(len(X), sample, replace=False)]
    cents = O.kmeans(S, nlist, iters=iters, seed=seed)
    offsets, ids = O.build_posting_lists(X, cents, max_clusters, threshold)
    return cents, offsets, ids


def doc_ids_for(n, seed=3, wide=True):
    """u128 doc ids: distinct, not monotone in the point id, some above 2^64."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n).astype(np.uint64)
    pairs = np.zeros((n, 2), dtype=np.uint64)
    pairs[:, 0] = perm * np.uint64(7919) + np.uint64(11)
    if wide:
        pairs[:, 1] = (perm % np.uint64(3))
    return pairs


def result_lists(od, os_, cnt):
    out = []
    for b in range(len(cnt)):
        n = int(cnt[b])
        out.append(([int(lo) | (int(hi) << 64) for lo, hi in np.asarray(od[b, :n], dtype=np.uint64)],
                    np.asarray(os_[b, :n], dtype=np.float32)))
    return out
