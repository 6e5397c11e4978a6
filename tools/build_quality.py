#!/usr/bin/env python
"""Graph quality of the batch-parallel device builder against the reference-order builder (the oracle's single-threaded restatement of
GraphIndexBuilder.addGraphNode, scoring through the reference's compiled kernels) on a 100 000-row subset of the c2 data: both graphs
are searched by the SAME device searcher at equal effort; recall@10 and visited nodes per query are printed for several rerankK.
    python tools/build_quality.py [--n 100000]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oracle_lib as o  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--nq", type=int, default=2000)
ap.add_argument("--no-oracle", action="store_true", help="skip the (slow) reference-order build")
ap.add_argument("--batch-div", default="", help="comma list of JV_BUILD_BATCH_DIV settings to compare (batch = inserted / div)")
a = ap.parse_args()
args = argparse.Namespace(impl="b200", n=a.n, dim=768, nq=a.nq, dist="latent", topk=10, gt_queries=a.nq)
cx = bench.Ctx(args)
jv, VSF, torch = cx.jv, cx.VSF, cx.torch
base_d = bench.gen_unit_rows_device(torch, bench.SEED, a.n, 768)
q = bench.gen_unit_rows_device(torch, bench.SEED + 1, a.nq, 768).cpu().numpy()
base = base_d.cpu().numpy()
vec = cx.adopt(base_d)
gt, _, _ = jv.topk_bruteforce(vec, VSF.DOT_PRODUCT, q, 10)
rows = []
if not a.no_oracle:
    L = o.load()
    kind = "reference kernels" if (os.path.exists(o.REF_SO) and L.jvo_use_ref(o.REF_SO.encode()) == 0) else "scalar port"
    t0 = time.time()
    adj = np.empty((a.n, 32), np.int32)
    entry = L.jvo_graph_build_f32(o.DOT_PRODUCT, o.fp(base), a.n, 768, 32, 100, 1.2, 1.2, o.ip(adj))
    L.jvo_use_ref(None)
    rows.append(("reference-order builder (oracle, %s)" % kind, jv.GraphIndex(adj, entry), time.time() - t0))
for div in [x for x in a.batch_div.split(",") if x]:
    os.environ["JV_BUILD_BATCH_DIV"] = div
    b = jv.GraphIndexBuilder(VSF.DOT_PRODUCT, M=32, beamWidth=100, neighborOverflow=1.2, alpha=1.2, addHierarchy=False, seed=bench.SEED)
    g = b.build(vec)
    rows.append(("device builder (batch = inserted / %s)" % div, g, b.device_ms / 1e3))
    del os.environ["JV_BUILD_BATCH_DIV"]
for window in (-1, 0):
    b = jv.GraphIndexBuilder(VSF.DOT_PRODUCT, M=32, beamWidth=100, neighborOverflow=1.2, alpha=1.2, addHierarchy=False, seed=bench.SEED, concurrent_window=window)
    g = b.build(vec)
    rows.append(("device builder (in-progress window %s)" % ("default" if window < 0 else "off"), g, b.device_ms / 1e3))
print("n = %d x 768 (c2 distribution), M = 32, efConstruction = 100, alpha 1.2, overflow 1.2, flat; %d queries, top-10" % (a.n, a.nq))
for name, g, secs in rows:
    _, ad = g.level(0)
    print("%-52s build %7.1f s  mean degree %.2f" % (name, secs, float((ad >= 0).sum(1).mean())))
    for oq in (1, 2, 5, 10):
        r = jv.GraphSearcher(g).search(vec, q, VSF.DOT_PRODUCT, 10, 10 * oq)
        print("    rerankK %3d: recall@10 %.4f  visited/query %.1f" % (10 * oq, bench.recall_at_k(r.nodes, gt, 10), r.visitedCount / a.nq))
