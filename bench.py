#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

  python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU kernels on the host cores

Workload (configs[1], `c2`, the default): synthetic 1M x 768 float32 unit rows from a latent-factor model (see gen_unit_rows;
seeds fixed), DOT_PRODUCT, Vamana graph M=32 efConstruction=100 overflow 1.2 alpha 1.2 with hierarchy (built on the device,
untimed set-up), GraphSearcher top-10 with rerankK = 10 x overquery (default 10). A "step" = one batch of `nq` = 10 000 queries
searched to completion. Other workloads (parity / coverage cases, not the headline): `c1` siftsmall, `c3` the same data through
PQ (M=96, k=256) ADC + fp32 rerank, `c4` BQ Hamming brute force over a range-sharded base (NCCL all-gather + device merge),
`c5` device graph build + NVQ encode. One JSON line on stdout (rank 0).

value   : queries/s with the query batch already resident in HBM, device time from CUDA events on the launching stream
e2e     : queries/s through the host-pointer C-ABI call (H2D of the queries and D2H of the results inside the timed region)
roofline: graph_search_kernel, algorithmic bytes = scored vectors x (row bytes + 8) / device time, vs MEASURED_PEAKS.json
cpu_baseline: the same traversal (oracle/jv_oracle.c driver) calling the reference's own compiled kernels
              (oracle/_ref/libjvector.so) on the host cores, on a bounded sample of the same queries.
"""
import argparse
import ctypes as C
import json
import os

# several ranks share one host: keep each rank's BLAS / OpenMP pools to its share of the cores (data generation only)
_world = int(os.environ.get("WORLD_SIZE", "1"))
if _world > 1:
    _share = str(max(1, (os.cpu_count() or 1) // _world))
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(_v, _share)
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20260922


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE graph_search_kernel launch, from the committed ncu --set full captures
# (profiles/r1_ncu_search_c2.md, profiles/r1_ncu_search_c3.md); key = (workload, n, nq, rerankK, dist)
NCU_TRAFFIC = {("c2", 1_000_000, 10_000, 100, "latent"): 94.728e9, ("c3", 1_000_000, 10_000, 100, "latent"): None}  # c3 capture pre-dates the L2-resident LUT mode

LATENT = 32      # intrinsic dimensionality of the synthetic embedding model
NOISE = 0.25     # isotropic noise relative to the per-coordinate signal


def gen_unit_rows(seed, n, dim, dist="latent", chunk=65536):
    """Synthetic float32 unit rows.
    dist="latent" (default): x = normalise(z A + NOISE * e), z ~ N(0, I_64), A a fixed 64 x dim Gaussian map, e ~ N(0, I_dim):
        embedding-like data with neighbourhood structure (intrinsic dimension ~64), so recall@10 is a meaningful axis.
    dist="iid": i.i.d. N(0,1) rows normalised (SURVEY §8d's first suggestion). Measured here: graph search on 1M x 768 i.i.d.
        rows reaches recall@10 ~ 0.05 for the reference traversal and this one alike (distance concentration), so QPS "at
        recall" is meaningless on it; it is kept as an option, not as the headline workload."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), dtype=np.float32)
    A = None
    if dist == "latent":
        A = (np.random.default_rng(SEED + 7).standard_normal((LATENT, dim)) / np.sqrt(LATENT)).astype(np.float32)
    for i in range(0, n, chunk):
        j = min(n, i + chunk)
        blk = rng.standard_normal((j - i, dim), dtype=np.float32)
        if A is not None:
            z = rng.standard_normal((j - i, LATENT), dtype=np.float32)
            blk *= np.float32(NOISE)
            blk += z @ A
        blk /= np.linalg.norm(blk, axis=1, keepdims=True)
        out[i:j] = blk
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recall_at_k(found, truth, k):
    """jvector-examples/.../util/AccuracyMetrics.java:38-50 (recallFromSearchResults, k = topK)"""
    hits = 0
    for f, t in zip(found, truth):
        hits += len(set(int(x) for x in f[:k] if x >= 0) & set(int(x) for x in t[:k]))
    return hits / float(len(found) * k)


def dist_setup(gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    td = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, local, td


def host_graph(gi):
    """download the device graph into the oracle's host representation (CPU baseline / reference arm only)"""
    import oracle_lib as o
    inf = gi.info()
    _, adj0 = gi.level(0)
    upper = [gi.level(l) for l in range(1, inf["levels"])]
    return o.make_graph(adj0, inf["entry_node"], upper if upper else None)


def cpu_search(args, base, graph_host, queries, topK, rerankK, pq=None):
    """The reference arm / cpu_baseline leg: oracle traversal driver + the reference's own compiled kernels."""
    import oracle_lib as o
    L = o.load()
    kind = "port"
    if os.path.exists(o.REF_SO) and L.jvo_use_ref(o.REF_SO.encode()) == 0:
        kind = "reference"
    ds = o.Dataset()
    ds.kind = 1 if pq else 0
    ds.metric = o.DOT_PRODUCT
    ds.dim = base.shape[1]
    ds.base = o.fp(base)
    ds.n = base.shape[0]
    if pq:
        ds.codebooks, ds.M, ds.k, ds.centroid, ds.codes = o.fp(pq["codebooks"]), pq["M"], 256, None, o.bp(pq["codes"])
    nq = queries.shape[0]
    nodes = np.empty((nq, topK), np.int32)
    scores = np.empty((nq, topK), np.float32)
    scored = C.c_int64()
    threads = os.cpu_count() or 1
    secs = L.jvo_graph_search_batch(C.byref(graph_host), C.byref(ds), o.fp(queries), nq, topK, rerankK, threads, o.ip(nodes), o.fp(scores), C.byref(scored))
    isa = L.jvo_ref_isa().decode()
    L.jvo_use_ref(None)
    return {"seconds": secs, "qps": nq / secs, "scored": int(scored.value), "threads": threads, "kind": kind, "isa": isa, "nodes": nodes}


def run_c4(args, rank, world, local, td, jv, nat, lib):
    """configs[3]: 1M x 1536 BQ Hamming first pass, a batch of 1k queries against a base RANGE-SHARDED over the ranks; the only
    exchange is the all-gather of per-shard top-k keys + the device merge. A step = one 1000-query batch."""
    import torch

    from jvector_b200 import parallel as par
    VSF = jv.VectorSimilarityFunction
    dim = 1536 if args.dim == 768 else args.dim
    n, nq, k = args.n, (1000 if args.nq == 10_000 else args.nq), args.topk * args.overquery
    W = (dim + 63) // 64
    lo, hi = par.shard_range(n, rank, world)
    t0 = time.time()
    words = np.empty((hi - lo, W), dtype=np.uint64)
    chunk = 65536
    for c0 in range(0, n, chunk):  # chunk seeds are global, so every rank derives the same base and keeps its slice
        c1 = min(n, c0 + chunk)
        a, b_ = max(c0, lo), min(c1, hi)
        if a >= b_:
            continue
        rows = np.random.default_rng([SEED, c0]).standard_normal((c1 - c0, dim), dtype=np.float32)
        words[a - lo:b_ - lo] = jv.bq_encode_all(rows[a - c0:b_ - c0])
    queries = np.random.default_rng(SEED + 5).standard_normal((nq, dim), dtype=np.float32)
    log("[rank %d] BQ shard [%d, %d) encoded in %.1fs" % (rank, lo, hi, time.time() - t0))
    bqv = jv.BQVectors(words, dim)
    torch.cuda.set_device(local)
    sb = par.gpu_sharded_bruteforce(td if world > 1 else None, bqv, VSF.COSINE, lo)
    qd = torch.from_numpy(queries).cuda()

    def step(q):
        return sb.search(q, k)

    def sync():
        torch.cuda.synchronize()
        nat.check(lib.jv_device_synchronize())
        if td is not None:
            td.barrier()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup + 30):  # a FIXED count: every step holds a collective, all ranks must issue the same number
        step(qd)
    sync()
    l0 = lib.jv_kernel_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        keys = step(qd)
    sync()
    dev_s = time.perf_counter() - t0
    launches = lib.jv_kernel_launch_count() - l0
    # e2e: host queries in (pinned), host keys out
    hq = torch.from_numpy(queries).pin_memory()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step(hq.cuda(non_blocking=True)).cpu()
    sync()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    if td is not None:
        t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        dev_s, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
        td.all_reduce(c, op=td.ReduceOp.SUM)
        launches = int(c[0])
    # check against the CPU oracle on a few queries of the batch (rank 0, its own shard when world > 1 -> skip unless world == 1)
    peak, peak_src = measured_peaks()
    pairs = float(args.steps) * nq * n
    unique_bytes = float(args.steps) * (hi - lo) * W * 8
    out_d = {"metric": "queries_per_sec_bq_bruteforce", "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
             "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
             "config": {"workload": "c4: synthetic %dx%d BQ (sign bits of N(0,1) rows), Hamming first pass top-%d, %d queries/step, base range-sharded over %d GPU(s)"
                                    % (n, dim, k, nq, world),
                        "parallelism": "base sharded by node-id range; one all_gather of [nq][k] keys + device merge",
                        "l2": "flush not needed: every pass streams the whole shard (%.0f MB) once per step" % ((hi - lo) * W * 8 / 1e6)},
             "value": args.steps * nq / dev_s, "ms_per_step": 1e3 * dev_s / args.steps, "pairs_per_sec": pairs / dev_s,
             "e2e": {"value": args.steps * nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(queries.nbytes), "d2h_bytes_per_step": int(nq * k * 8)},
             "gpu_launches": int(launches), "clocks": clocks,
             "roofline": {"kernel": "topk_filter_bq_kernel", "bound": "hbm", "achieved": unique_bytes / dev_s / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": unique_bytes / dev_s / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                          "note": "popcount-issue bound, not HBM bound: %d x popc64 per pair; streamed-bytes form = %.0f GB/s per GPU"
                                  % (W, pairs / world * W * 8 / dev_s / 1e9)}}
    if rank == 0 and world == 1 and not args.no_cpu:
        import ctypes as C

        import oracle_lib as o
        L = o.load()
        nqs = min(nq, 256)
        qw = np.zeros((nqs, W), np.uint64)
        for i in range(nqs):
            L.jvo_bq_encode(o.fp(queries[i]), dim, o.wp(qw[i]))
        want = np.empty((nqs, k), np.int64)
        threads = os.cpu_count() or 1
        cpu_s = L.jvo_bq_bruteforce_batch(o.wp(words), n, dim, o.wp(qw), nqs, k, threads, o.lp(want))
        out_d["cpu_baseline"] = {"value": nqs / cpu_s, "unit": "queries/s", "cores": threads, "kind": "port",
                                 "pairs_per_sec": nqs * float(n) / cpu_s,
                                 "sample": "%d of the %d queries, scalar popcount loop (DefaultVectorUtilSupport.java:342-348; the reference has no native Hamming), %d threads, %.1f s"
                                           % (nqs, nq, threads, cpu_s)}
        out_d["parity"] = "keys bit-identical to the oracle for %d queries x %d rows: %s" % (nqs, n, bool(np.array_equal(keys[:nqs].cpu().numpy(), want)))
    if rank == 0:
        print(json.dumps(out_d), flush=True)
    if td is not None:
        td.destroy_process_group()
    return 0


def run_c5(args, rank, world, local, td, jv, nat, lib):
    """configs[4] shape: GraphIndexBuilder build (M=32, ef=100) + NVQ inline vectors (2 sub-vectors) on the device.
    A step = one full build of the n x 768 index followed by the NVQ encode of every row. Default n = 1M (10M x 768 = 30.7 GB
    fits one B200 but not this bench's few-minute budget for host data generation). Single GPU per rank (replicas)."""
    VSF = jv.VectorSimilarityFunction
    n, dim, nsub = args.n, args.dim, 2
    base = gen_unit_rows(SEED, n, dim, args.dist)
    queries = gen_unit_rows(SEED + 1, min(args.nq, 2000), dim, args.dist)
    vec = jv.F32Vectors(base)
    mean = base.mean(0).astype(np.float32)
    steps = max(1, min(args.steps, 3))
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.jv_kernel_launch_count()
    build_ms, enc_s, scored = [], [], 0
    gi = None
    for _ in range(steps):
        b = jv.GraphIndexBuilder(VSF.DOT_PRODUCT, M=32, beamWidth=100, neighborOverflow=1.2, alpha=1.2, addHierarchy=True, seed=SEED)
        gi = b.build(vec)
        build_ms.append(b.device_ms)
        scored = b.scored_vectors
        t0 = time.perf_counter()
        params, bys = jv.nvq_encode_all(vec, mean, nsub, True)  # rows already resident in HBM; params + bytes copied back to the host
        enc_s.append(time.perf_counter() - t0)
    launches = lib.jv_kernel_launch_count() - l0
    clocks = sampler.stop()
    nvq = jv.NVQVectors(bys, params, mean, nsub)
    gt, _, _ = jv.topk_bruteforce(vec, VSF.DOT_PRODUCT, queries, 10)
    res = jv.GraphSearcher(gi).search(vec, queries, VSF.DOT_PRODUCT, 10, 100, reranker=nvq)
    rec = recall_at_k(res.nodes, gt, 10)
    bm = float(np.median(build_ms))
    peak, peak_src = measured_peaks()
    out = {"metric": "build_inserts_per_sec", "unit": "vectors/s", "n_gpus": world, "steps": steps, "warmup": 0, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "c5: GraphIndexBuilder build of %dx%d float32 (%s) M=32 ef=100 overflow=1.2 alpha=1.2 hierarchy on the device, then NVQ (2 sub-vectors, learned) encode of every row"
                                  % (n, dim, args.dist), "parallelism": "one replica per GPU"},
           "value": n / (bm / 1e3), "ms_per_step": bm, "build_scored_vectors_per_sec": scored / (bm / 1e3),
           "nvq_encode_vectors_per_sec": n / float(np.median(enc_s)), "recall_at_10_fp32_walk_nvq_rerank": rec,
           "e2e": {"value": n / (bm / 1e3 + float(np.median(enc_s))), "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(bys.nbytes + params.nbytes)},
           "gpu_launches": int(launches), "clocks": clocks,
           "roofline": {"kernel": "graph_search_kernel (insert searches)", "bound": "hbm", "achieved": scored * (dim * 4 + 8) / (bm / 1e3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": scored * (dim * 4 + 8) / (bm / 1e3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                        "note": "whole-build time in the denominator (search + prune + back-links); the insert searches alone run near the c2 fraction"}}
    if rank == 0 and not args.no_cpu:
        import ctypes as C

        import oracle_lib as o
        L = o.load()
        kind = "reference" if (os.path.exists(o.REF_SO) and L.jvo_use_ref(o.REF_SO.encode()) == 0) else "port"
        ns = min(n, 200_000)
        p2 = np.empty((ns, nsub, 4), np.float32)
        b2 = np.empty((ns, dim), np.uint8)
        threads = os.cpu_count() or 1
        secs = L.jvo_nvq_encode_batch(o.fp(base), ns, dim, nsub, o.fp(mean), 1, threads, o.fp(p2), o.bp(b2))
        L.jvo_use_ref(None)
        same = (p2[:, :, 2] == params[:ns, :, 2])
        out["cpu_baseline"] = {"value": ns / secs, "unit": "vectors/s (NVQ encode)", "cores": threads, "kind": kind,
                               "sample": "NVQ encode of the first %d rows through the reference kernels (nvq_uniform_loss + 40 x nvq_loss + nvq_quantize_8bit), %.1f s; "
                                         "the reference's graph BUILD cannot run here (no JVM)" % (ns, secs)}
        out["parity"] = "growth-rate equal for %.4f of sub-vectors; bytes equal where equal: %s" % (
            float(same.mean()), bool(all(np.array_equal(bys[i, :dim // 2], b2[i, :dim // 2]) for i in np.flatnonzero(same[:, 0])[:2000])))
    if rank == 0:
        print(json.dumps(out), flush=True)
    if td is not None:
        td.destroy_process_group()
    return 0


def run_c1(args, rank, world, local, td, jv, nat, lib):
    """configs[0]: siftsmall 10k x 128, exact L2: brute-force top-100 against the shipped ground truth, and graph search
    (M=16, ef=100, overflow 1.2, alpha 1.2, no hierarchy: SiftSmall.java:86-93). A step = the 100 queries."""
    import oracle_lib as o
    VSF = jv.VectorSimilarityFunction
    base, queries, gt = o.load_siftsmall()
    vec = jv.F32Vectors(base)
    gi = jv.GraphIndexBuilder(VSF.EUCLIDEAN, M=16, beamWidth=100, neighborOverflow=1.2, alpha=1.2, addHierarchy=False, seed=SEED).build(vec)
    s = jv.GraphSearcher(gi)
    for _ in range(max(3, args.warmup)):
        res = s.search(vec, queries, VSF.EUCLIDEAN, 100, 100)
        nodes, _, _ = jv.topk_bruteforce(vec, VSF.EUCLIDEAN, queries, 100)
    l0 = lib.jv_kernel_launch_count()
    t_graph, t_bf, scored = 0.0, 0.0, 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        res = s.search(vec, queries, VSF.EUCLIDEAN, 100, 100)
        t_graph += time.perf_counter() - t0
        scored += res.visitedCount + 100
        t0 = time.perf_counter()
        nodes, _, _ = jv.topk_bruteforce(vec, VSF.EUCLIDEAN, queries, 100)
        t_bf += time.perf_counter() - t0
    launches = lib.jv_kernel_launch_count() - l0
    rec_graph = recall_at_k(res.nodes, gt, 100)
    rec_bf = recall_at_k(nodes, gt, 100)
    out = {"metric": "queries_per_sec_at_recall@100", "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "siftsmall (tests/golden/siftsmall)",
           "config": {"workload": "c1: siftsmall 10000x128 float32, exact L2, graph M=16 ef=100 topK=100 (e2e host-pointer calls; the data set fits in L2, launch-latency bound)"},
           "value": args.steps * 100 / t_graph, "ms_per_step": 1e3 * t_graph / args.steps, "recall_at_100": rec_graph,
           "bruteforce_queries_per_sec": args.steps * 100 / t_bf, "bruteforce_recall_at_100_vs_shipped_ground_truth": rec_bf,
           "scored_vectors_per_sec": scored / t_graph,
           "e2e": {"value": args.steps * 100 / t_graph, "unit": "queries/s", "h2d_bytes_per_step": int(queries.nbytes), "d2h_bytes_per_step": 100 * 100 * 8},
           "gpu_launches": int(launches), "roofline": {"bound": "hbm", "achieved": scored * 520 / t_graph / 1e9, "peak": measured_peaks()[0], "unit": "GB/s",
                                                       "frac": scored * 520 / t_graph / 1e9 / measured_peaks()[0], "traffic": None,
                                                       "note": "5 MB data set: L2 resident and launch bound, the fraction is not meaningful here"}}
    if not args.no_cpu:
        gh = host_graph(gi)
        r = cpu_search_generic(base, gh, queries, 100, 100, o.EUCLIDEAN)
        out["cpu_baseline"] = {"value": r["qps"], "unit": "queries/s", "cores": r["threads"], "kind": r["kind"], "isa": r["isa"],
                               "recall_at_100": recall_at_k(r["nodes"], gt, 100), "sample": "the 100 siftsmall queries, graph search, %d host threads" % r["threads"]}
    print(json.dumps(out), flush=True)
    return 0


def cpu_search_generic(base, graph_host, queries, topK, rerankK, metric):
    import oracle_lib as o
    L = o.load()
    kind = "reference" if (os.path.exists(o.REF_SO) and L.jvo_use_ref(o.REF_SO.encode()) == 0) else "port"
    ds = o.Dataset()
    ds.kind, ds.metric, ds.dim, ds.base, ds.n = 0, metric, base.shape[1], o.fp(base), base.shape[0]
    nq = queries.shape[0]
    nodes = np.empty((nq, topK), np.int32)
    scores = np.empty((nq, topK), np.float32)
    scored = C.c_int64()
    threads = min(os.cpu_count() or 1, nq)
    secs = L.jvo_graph_search_batch(C.byref(graph_host), C.byref(ds), o.fp(queries), nq, topK, rerankK, threads, o.ip(nodes), o.fp(scores), C.byref(scored))
    isa = L.jvo_ref_isa().decode()
    L.jvo_use_ref(None)
    return {"seconds": secs, "qps": nq / secs, "scored": int(scored.value), "threads": threads, "kind": kind, "isa": isa, "nodes": nodes}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--overquery", type=int, default=10)
    ap.add_argument("--gt-queries", type=int, default=1000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="also report overquery 1/2/5/10")
    ap.add_argument("--dist", default="latent", choices=["latent", "iid"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    if args.impl == "reference":
        # the CPU arm runs on rank 0 alone; the other ranks exit without joining any process group
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        rank, world, local, td = 0, 1, int(os.environ.get("LOCAL_RANK", "0")), None
    else:
        rank, world, local, td = dist_setup(args.gpus)

    import jvector_b200 as jv
    from jvector_b200 import _native as nat
    VSF = jv.VectorSimilarityFunction
    lib = nat.init(local)
    if args.workload == "c4":
        return run_c4(args, rank, world, local, td, jv, nat, lib)
    if args.workload == "c5":
        return run_c5(args, rank, world, local, td, jv, nat, lib)
    if args.workload == "c1":
        return run_c1(args, rank, world, local, td, jv, nat, lib)

    t0 = time.time()
    base = gen_unit_rows(SEED, args.n, args.dim, args.dist)
    queries = gen_unit_rows(SEED + 1 + rank, args.nq, args.dim, args.dist)
    log("[rank %d] data generated in %.1fs" % (rank, time.time() - t0))
    vec = jv.F32Vectors(base)
    t0 = time.time()
    builder = jv.GraphIndexBuilder(VSF.DOT_PRODUCT, M=32, beamWidth=100, neighborOverflow=1.2, alpha=1.2, addHierarchy=True, seed=SEED)
    gi = builder.build(vec)
    build_s = time.time() - t0
    log("[rank %d] graph built in %.1fs (device %.1fs) %s" % (rank, build_s, builder.device_ms / 1e3, gi.info()))

    topK, rerankK = args.topk, args.topk * args.overquery
    pq = None
    approx, reranker = vec, None
    row_bytes = args.dim * 4
    if args.workload == "c3":
        import oracle_lib as o
        M = args.dim // 8
        rs = np.random.default_rng(SEED + 99)
        sample = base[rs.choice(args.n, min(args.n, 20000), replace=False)]
        cb, _, _ = o.train_pq_numpy(rs, sample, M, 256, iters=6)
        codes = jv.pq_encode_all(vec, cb, M, 256)
        pq = {"codebooks": cb, "codes": codes, "M": M}
        approx, reranker = jv.PQVectors(codes, cb, args.dim, 256), vec

    # ground truth by exhaustive scoring with the reference's ordering key
    ngt = min(args.gt_queries, args.nq)
    gt_nodes, _, _ = jv.topk_bruteforce(vec, VSF.DOT_PRODUCT, queries[:ngt], topK)
    searcher = jv.GraphSearcher(gi)

    out = {"metric": "queries_per_sec_at_recall@10", "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s: synthetic %dx%d float32 unit rows (%s), DOT_PRODUCT, graph M=32 ef=100 overflow=1.2 alpha=1.2 hierarchy, "
                                  "GraphSearcher top-%d rerankK=%d, %d queries/step/GPU" % (args.workload, args.n, args.dim, args.dist, topK, rerankK, args.nq),
                      "parallelism": "replica per GPU, queries sharded (no data-path collective)",
                      "l2": "inputs %.2f GB >> 126 MB L2 (random row gathers)" % (base.nbytes / 1e9)}}

    if args.impl == "reference":
        gh = host_graph(gi)
        nqs = args.cpu_sample or args.nq
        for _ in range(args.warmup):
            cpu_search(args, base, gh, queries[: max(50, nqs // 10)], topK, rerankK, pq)
        secs, scored, last = 0.0, 0, None
        for _ in range(args.steps):
            last = cpu_search(args, base, gh, queries[:nqs], topK, rerankK, pq)
            secs += last["seconds"]
            scored += last["scored"]
        qps = args.steps * nqs / secs
        rec = recall_at_k(last["nodes"][:ngt], gt_nodes[: min(ngt, nqs)], topK)
        out.update({"impl": "reference", "value": qps, "ms_per_step": 1e3 * secs / args.steps, "recall_at_10": rec,
                    "scored_vectors_per_sec": scored / secs,
                    "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": last["threads"], "kind": last["kind"], "isa": last["isa"],
                                     "sample": "%d of the %d queries per step, all host threads" % (nqs, args.nq)},
                    "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0, "setup": "graph built with the device builder (untimed); timed path is CPU only"})
        print(json.dumps(out), flush=True)
        return 0

    # ---- device-resident leg: queries already in HBM ----
    nq = args.nq
    dq, dn, ds_ = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nat.check(lib.jv_device_malloc(C.byref(dq), queries.nbytes))
    nat.check(lib.jv_device_malloc(C.byref(dn), nq * topK * 4))
    nat.check(lib.jv_device_malloc(C.byref(ds_), nq * topK * 4))
    nat.check(lib.jv_memcpy_h2d(dq, queries.ctypes.data, queries.nbytes))
    st = nat.SearchStats()
    rr = reranker._h if reranker is not None else None

    def step_device():
        nat.check(lib.jv_graph_search_batch_device(gi._h, approx._h, rr, int(VSF.DOT_PRODUCT), dq, nq, topK, rerankK, dn, ds_, C.byref(st)))
        return st.device_ms, st.visited + nq + st.reranked

    def barrier():
        if td is not None:
            td.barrier()
        nat.check(lib.jv_device_synchronize())

    sampler = ClockSampler(local)
    sampler.start()
    t_w = time.time()
    while time.time() - t_w < 1.0:  # >= 1 s of load before timing so the clock samples are under load
        step_device()
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = lib.jv_kernel_launch_count()
    dev_ms, scored, t0 = 0.0, 0, time.time()
    for _ in range(args.steps):
        ms, sc = step_device()
        dev_ms += ms
        scored += sc
    barrier()
    wall_s = time.time() - t0
    launches = lib.jv_kernel_launch_count() - launches0
    nodes = np.empty((nq, topK), np.int32)
    nat.check(lib.jv_memcpy_d2h(nodes.ctypes.data, dn, nodes.nbytes))
    rec = recall_at_k(nodes[:ngt], gt_nodes, topK)

    # ---- end-to-end leg: pinned host buffers in, host results out, copies inside the timed region ----
    hq = np.ascontiguousarray(queries)
    hn = np.empty((nq, topK), np.int32)
    hs = np.empty((nq, topK), np.float32)
    for a in (hq, hn, hs):
        lib.jv_host_register(a.ctypes.data, a.nbytes)
    st2 = nat.SearchStats()

    def step_e2e():
        nat.check(lib.jv_graph_search_batch(gi._h, approx._h, rr, int(VSF.DOT_PRODUCT), nat.fp(hq), nq, topK, rerankK, nat.ip(hn), nat.fp(hs), C.byref(st2)))

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.time() - t0
    clocks = sampler.stop()
    for a in (hq, hn, hs):
        lib.jv_host_unregister(a.ctypes.data)

    # max over ranks
    if td is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_s, wall_s], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        dev_ms, e2e_s, wall_s = [float(x) for x in t.tolist()]
        c = torch.tensor([float(scored), rec, float(launches)], dtype=torch.float64, device="cuda")
        td.all_reduce(c, op=td.ReduceOp.SUM)
        scored, rec, launches = float(c[0]), float(c[1]) / world, int(c[2])
    total_q = args.steps * nq * world
    peak, peak_src = measured_peaks()
    per_unit = (row_bytes + 8) if args.workload == "c2" else None
    if args.workload == "c2":
        algo_bytes = scored * per_unit / world  # per GPU
    else:
        algo_bytes = ((st.visited + nq) * (approx.M + 8) + st.reranked * (row_bytes + 8)) * args.steps
    achieved = algo_bytes / (dev_ms / 1e3) / 1e9
    out.update({"value": total_q / (dev_ms / 1e3), "ms_per_step": dev_ms / args.steps, "recall_at_10": rec,
                "scored_vectors_per_sec": scored / (dev_ms / 1e3), "visited_per_query": st.visited / float(nq),
                "wall_ms_per_step": 1e3 * wall_s / args.steps,
                "e2e": {"value": total_q / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(hq.nbytes), "d2h_bytes_per_step": int(hn.nbytes + hs.nbytes)},
                "gpu_launches": int(launches), "clocks": clocks, "build_seconds": build_s,
                "roofline": {"kernel": "graph_search_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": NCU_TRAFFIC.get((args.workload, args.n, args.nq, rerankK, args.dist)), "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": algo_bytes / args.steps, "algorithmic_bytes_per_scored_vector": per_unit}})

    if rank == 0 and world == 1 and args.workload == "c2":
        # the host-expanded-frontier form of the same path (north_star's literal seam): one launch per hop / per multi-query step.
        # Reported beside the headline, not part of it: a single hop is launch-latency bound, a 10 000-query step is HBM bound.
        rs = np.random.default_rng(SEED + 3)
        sf = vec.score_function_for(queries[0], VSF.DOT_PRODUCT)
        ids32 = rs.integers(0, args.n, 32).astype(np.int32)
        for _ in range(50):
            sf.similarityToBatch(ids32)
        t0 = time.perf_counter()
        for _ in range(500):
            sf.similarityToBatch(ids32)
        hop_us = (time.perf_counter() - t0) / 500 * 1e6
        sf.close()
        mq = min(nq, 10000)
        off = (np.arange(mq + 1, dtype=np.int32) * 32)
        mids = rs.integers(0, args.n, mq * 32).astype(np.int32)
        jv.score_multi(vec, VSF.DOT_PRODUCT, queries[:mq], mids, off)
        t0 = time.perf_counter()
        for _ in range(5):
            jv.score_multi(vec, VSF.DOT_PRODUCT, queries[:mq], mids, off)
        step_s = (time.perf_counter() - t0) / 5
        out["host_driven"] = {"single_hop_32_candidates_us": hop_us, "multi_query_step": {"queries": mq, "candidates_per_query": 32,
                              "e2e_ms": 1e3 * step_s, "scored_vectors_per_sec_e2e": mq * 32 / step_s,
                              "note": "jv_score_batch / jv_score_multi through host pointers (H2D ids+queries, D2H scores inside the call)"}}

    if args.sweep and rank == 0:
        sweep = []
        for oq in (1, 2, 5, 10):
            rk = topK * oq
            nat.check(lib.jv_graph_search_batch_device(gi._h, approx._h, rr, int(VSF.DOT_PRODUCT), dq, nq, topK, rk, dn, ds_, C.byref(st)))
            nat.check(lib.jv_graph_search_batch_device(gi._h, approx._h, rr, int(VSF.DOT_PRODUCT), dq, nq, topK, rk, dn, ds_, C.byref(st)))
            nat.check(lib.jv_memcpy_d2h(nodes.ctypes.data, dn, nodes.nbytes))
            sweep.append({"overquery": oq, "qps": nq / (st.device_ms / 1e3), "recall_at_10": recall_at_k(nodes[:ngt], gt_nodes, topK),
                          "visited_per_query": st.visited / float(nq)})
        out["sweep"] = sweep

    if rank == 0 and world == 1 and not args.no_cpu:
        gh = host_graph(gi)
        nqs = args.cpu_sample or nq
        cpu_search(args, base, gh, queries[: max(20, nqs // 10)], topK, rerankK, pq)
        r = cpu_search(args, base, gh, queries[:nqs], topK, rerankK, pq)
        out["cpu_baseline"] = {"value": r["qps"], "unit": "queries/s", "cores": r["threads"], "kind": r["kind"], "isa": r["isa"],
                               "scored_vectors_per_sec": r["scored"] / r["seconds"],
                               "recall_at_10": recall_at_k(r["nodes"][:ngt], gt_nodes[: min(ngt, nqs)], topK),
                               "sample": "%d of the %d queries of one step, %d host threads, %.1f s" % (nqs, nq, r["threads"], r["seconds"])}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if td is not None:
        td.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
