#!/usr/bin/env python
"""Search-kernel tuning / profiling helper: ONE process builds the c2 world (rows generated on the device, graph, PQ codes, FusedPQ
records) and then times graph_search_kernel under every requested setting of the runtime knobs, which the library re-reads on
each call:

    python tools/profile_search.py --sweep                      # c2 + c3 over JV_PQ_LUT_SMEM_M x JV_FUSED_PQ
    ncu --profile-from-start off ... python tools/profile_search.py --workload c3 --reps 1 --ncu   # only the timed launches are profiled

Same data and parameters as bench.py (workloads c2 / c3)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sweep", action="store_true")
ap.add_argument("--workload", default="c3", choices=["c2", "c3", "seam"])
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--smem-m", default="0,24,32,48,64,96")
ap.add_argument("--envs", default="", help="';'-separated settings, each 'A=1,B=2' (empty = defaults): every workload is timed under each")
ap.add_argument("--workloads", default="", help="comma list of c2,c3 for --envs")
ap.add_argument("--sustain", type=float, default=0.0, help="seconds of back-to-back launches before the timed reps (power-capped steady state, like bench.py); reports the MEAN of the reps")
ap.add_argument("--ncu", action="store_true", help="bracket the timed launches with cudaProfilerStart/Stop (ncu --profile-from-start off)")
a = ap.parse_args()
if a.ncu:
    # under a profiler the kernel must not wait for copies of another stream (the host-pointer search overlaps its H2D copy with the
    # kernel behind an arrival watermark; ncu synchronises before a profiled launch, so this is belt and braces)
    os.environ["JV_SEARCH_OVERLAP"] = "0"

args = argparse.Namespace(impl="b200", n=a.n, dim=768, nq=a.nq, dist="latent", topk=10, gt_queries=500)
cx = bench.Ctx(args)
jv, VSF = cx.jv, cx.VSF
w = bench.World2(cx)
s = jv.GraphSearcher(w.gi)


def run(approx, rr, label):
    best = None
    if a.ncu:
        s.search(approx, w.queries, VSF.DOT_PRODUCT, 10, 100, reranker=rr)  # warm-up outside the profiled range
        cx.torch.cuda.cudart().cudaProfilerStart()
    if a.sustain > 0:
        import subprocess
        import time
        t0 = time.time()
        while time.time() - t0 < a.sustain:
            s.search(approx, w.queries, VSF.DOT_PRODUCT, 10, 100, reranker=rr)
        ms = []
        for _ in range(a.reps):
            best = s.search(approx, w.queries, VSF.DOT_PRODUCT, 10, 100, reranker=rr)
            ms.append(best.device_ms)
        smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        print("%-44s sustained mean %8.3f ms (min %.3f max %.3f) qps %9.0f  [%s]" % (label, sum(ms) / len(ms), min(ms), max(ms), a.nq / (sum(ms) / len(ms) / 1e3), smi), flush=True)
        return best
    for _ in range(a.reps):
        r = s.search(approx, w.queries, VSF.DOT_PRODUCT, 10, 100, reranker=rr)
        best = r if best is None or r.device_ms < best.device_ms else best
    if a.ncu:
        cx.torch.cuda.cudart().cudaProfilerStop()
    rec = bench.recall_at_k(best.nodes[:w.ngt], w.gt_nodes, 10)
    print("%-44s device_ms %8.3f  qps %9.0f  visited/q %.1f  recall@10 %.4f" % (label, best.device_ms, a.nq / (best.device_ms / 1e3), best.visitedCount / a.nq, rec), flush=True)
    return best


def with_env(setting, fn):
    kv = [x.split("=", 1) for x in setting.split(",") if x]
    old = {k: os.environ.get(k) for k, _ in kv}
    for k, v in kv:
        os.environ[k] = v
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def make_pq():
    import oracle_lib as o
    rs = np.random.default_rng(bench.SEED + 99)
    sel = np.sort(rs.choice(a.n, min(a.n, 20000), replace=False))
    sample = w.base_dev[cx.torch.from_numpy(sel).cuda()].cpu().numpy()
    cb, _, _ = o.train_pq_numpy(rs, sample, 96, 256, iters=6)
    codes = jv.pq_encode_all(w.vec, cb, 96, 256)
    pqv = jv.PQVectors(codes, cb, 768, 256)
    w.gi.fuse_pq(pqv)
    return pqv


if a.envs:
    wl = [x for x in (a.workloads or "c2,c3").split(",") if x]
    pqv = make_pq() if "c3" in wl else None
    refs = {}
    for setting in a.envs.split(";"):
        for k in wl:
            r = with_env(setting, lambda: run(w.vec, None, "c2 [%s]" % setting) if k == "c2" else run(pqv, w.vec, "c3 [%s]" % setting))
            if k not in refs:
                refs[k] = r.nodes
            elif not np.array_equal(refs[k], r.nodes):
                print("   !! id lists differ from the first setting")
    sys.exit(0)
if a.workload == "seam":
    # the host-driven multi-query step (jv_query_batch_score -> score_ragged_kernel): 10 000 searches x 32 candidates
    rs = np.random.default_rng(bench.SEED + 3)
    mq = min(a.nq, 10000)
    off = np.arange(mq + 1, dtype=np.int32) * 32
    mids = rs.integers(0, a.n, mq * 32).astype(np.int32)
    qb = jv.QueryBatch(w.vec, w.queries[:mq], VSF.DOT_PRODUCT)
    qb.score_step(mids, off)
    if a.ncu:
        cx.torch.cuda.cudart().cudaProfilerStart()
    _, ms = qb.score_step(mids, off, return_ms=True)
    if a.ncu:
        cx.torch.cuda.cudart().cudaProfilerStop()
    print("score_ragged_kernel: %.3f ms for %d x 32 rows = %.1f GB/s" % (ms, mq, mq * 32 * 3080 / ms / 1e6))
    sys.exit(0)
if a.sweep or a.workload == "c2":
    run(w.vec, None, "c2 fp32 walk (48 registers, 5 CTAs/SM)")
    if a.sweep:
        os.environ["JV_SEARCH_WIDE"] = "1"
        run(w.vec, None, "c2 fp32 walk (64 registers, 4 CTAs/SM)")
        del os.environ["JV_SEARCH_WIDE"]
if a.sweep or a.workload == "c3":
    import oracle_lib as o
    rs = np.random.default_rng(bench.SEED + 99)
    sel = np.sort(rs.choice(a.n, min(a.n, 20000), replace=False))
    sample = w.base_dev[cx.torch.from_numpy(sel).cuda()].cpu().numpy()
    cb, _, _ = o.train_pq_numpy(rs, sample, 96, 256, iters=6)
    codes = jv.pq_encode_all(w.vec, cb, 96, 256)
    pqv = jv.PQVectors(codes, cb, 768, 256)
    w.gi.fuse_pq(pqv)
    if a.sweep:
        ref = None
        for fused in ("1", "0"):
            os.environ["JV_FUSED_PQ"] = fused
            for m in a.smem_m.split(","):
                os.environ["JV_PQ_LUT_SMEM_M"] = m
                r = run(pqv, w.vec, "c3 PQ walk fused=%s lut_smem_m=%s" % (fused, m))
                if ref is None:
                    ref = r.nodes
                elif not np.array_equal(ref, r.nodes):
                    print("   !! id lists differ from the first configuration")
        del os.environ["JV_FUSED_PQ"], os.environ["JV_PQ_LUT_SMEM_M"]
    else:
        run(pqv, w.vec, "c3 PQ walk (defaults)")
