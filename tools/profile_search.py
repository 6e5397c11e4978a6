#!/usr/bin/env python
"""Profiling helper (keeps ncu away from the graph build):
   python tools/profile_search.py --prepare   # build the c2 graph, save the adjacency under /tmp
   ncu ... python tools/profile_search.py --run [--workload c2|c3]   # load it, launch graph_search_kernel 3 times
Same data / parameters as bench.py (workload c2 / c3)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import jvector_b200 as jv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--prepare", action="store_true")
ap.add_argument("--run", action="store_true")
ap.add_argument("--workload", default="c2")
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--dir", default="/tmp/jv_profile")
a = ap.parse_args()
os.makedirs(a.dir, exist_ok=True)
VSF = jv.VectorSimilarityFunction
jv.init(0)
base = bench.gen_unit_rows(bench.SEED, a.n, 768)
vec = jv.F32Vectors(base)
if a.prepare:
    gi = jv.GraphIndexBuilder(VSF.DOT_PRODUCT, M=32, beamWidth=100, neighborOverflow=1.2, alpha=1.2, addHierarchy=True, seed=bench.SEED).build(vec)
    inf = gi.info()
    np.save(os.path.join(a.dir, "entry.npy"), np.array([inf["entry_node"], inf["levels"]]))
    for l in range(inf["levels"]):
        ids, adj = gi.level(l)
        np.save(os.path.join(a.dir, "ids%d.npy" % l), ids)
        np.save(os.path.join(a.dir, "adj%d.npy" % l), adj)
    if a.workload == "c3":
        import oracle_lib as o
        rs = np.random.default_rng(bench.SEED + 99)
        cb, _, _ = o.train_pq_numpy(rs, base[rs.choice(a.n, 20000, replace=False)], 96, 256, iters=6)
        np.save(os.path.join(a.dir, "cb.npy"), cb)
        np.save(os.path.join(a.dir, "codes.npy"), jv.pq_encode_all(base, cb, 96, 256))
    print("prepared", inf)
if a.run:
    entry, levels = np.load(os.path.join(a.dir, "entry.npy"))
    upper = [(np.load(os.path.join(a.dir, "ids%d.npy" % l)), np.load(os.path.join(a.dir, "adj%d.npy" % l))) for l in range(1, int(levels))]
    gi = jv.GraphIndex(np.load(os.path.join(a.dir, "adj0.npy")), int(entry), upper)
    queries = bench.gen_unit_rows(bench.SEED + 1, a.nq, 768)
    s = jv.GraphSearcher(gi)
    approx, rr = vec, None
    if a.workload == "c3":
        approx, rr = jv.PQVectors(np.load(os.path.join(a.dir, "codes.npy")), np.load(os.path.join(a.dir, "cb.npy")), 768, 256), vec
    for _ in range(a.reps):
        r = s.search(approx, queries, VSF.DOT_PRODUCT, 10, 100, reranker=rr)
        print("device_ms %.3f visited/q %.1f qps %.0f" % (r.device_ms, r.visitedCount / a.nq, a.nq / (r.device_ms / 1e3)))
