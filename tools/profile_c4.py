#!/usr/bin/env python
"""ncu / timing target for config 4: BQ Hamming top-k of 1000 queries over N x 1536 bits (tensor-core contraction, csrc/bq_imma.cu).
   python tools/profile_c4.py                                            # times both paths (IMMA and the round-1 popcount kernels)
   ncu --profile-from-start off ... python tools/profile_c4.py --ncu     # profiles one IMMA call"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jvector_b200 as jv  # noqa: E402

n, dim, nq, k = int(os.environ.get("N", 1_000_000)), 1536, 1000, 100
ncu = "--ncu" in sys.argv
jv.init(0)
rng = np.random.default_rng(1)
words = rng.integers(0, 2**63, size=(n, dim // 64), dtype=np.int64).astype(np.uint64)  # random bit packs: same work as real ones
bqv = jv.BQVectors(words, dim)
q = rng.standard_normal((nq, dim)).astype(np.float32)
VSF = jv.VectorSimilarityFunction
jv.topk_bruteforce(bqv, VSF.COSINE, q, k)
if ncu:
    import torch
    torch.cuda.cudart().cudaProfilerStart()
    jv.topk_bruteforce(bqv, VSF.COSINE, q, k)
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
for mode in ("imma", "popc"):
    if mode == "popc":
        os.environ["JV_BQ_BRUTEFORCE"] = "popc"
    jv.topk_bruteforce(bqv, VSF.COSINE, q, k)
    t0 = time.perf_counter()
    for _ in range(5):
        nodes, scores, keys = jv.topk_bruteforce(bqv, VSF.COSINE, q, k)
    dt = (time.perf_counter() - t0) / 5
    print("%s: %.3f ms per 1000-query call (host pointers), %.1f G pairs/s" % (mode, dt * 1e3, nq * n / dt / 1e9), flush=True)
