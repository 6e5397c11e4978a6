#!/usr/bin/env python
"""ncu target for config 4: BQ brute force over a smaller base (same kernel, same per-row work), 3 calls."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jvector_b200 as jv  # noqa: E402

n, dim, nq, k = int(os.environ.get("N", 1_000_000)), 1536, 1000, 100
jv.init(0)
rng = np.random.default_rng(1)
words = rng.integers(0, 2**63, size=(n, dim // 64), dtype=np.int64).astype(np.uint64)  # random bit packs: same work as real ones
bqv = jv.BQVectors(words, dim)
q = rng.standard_normal((nq, dim)).astype(np.float32)
for _ in range(3):
    nodes, scores, keys = jv.topk_bruteforce(bqv, jv.VectorSimilarityFunction.COSINE, q, k)
print("ok", nodes[0, :5])
