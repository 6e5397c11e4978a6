"""N > 1 path on CPU: world_size-2 gloo processes run the sharded brute-force plumbing (range sharding, global node ids in
the keys, the single all_gather, the k-way merge) with the local top-k supplied by the CPU oracle. Checks that the merged
result equals the oracle's top-k over the unsharded base, bit for bit."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist

    import oracle_lib as o
    from jvector_b200 import parallel as par
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = o.load()
    rng = np.random.default_rng(77)
    n, dim, nq, k = 1001, 24, 7, 9
    base = o.random_unit_vectors(rng, n, dim)
    base[500] = base[3]  # exact duplicates in different shards: the tie must go to the smaller global id
    queries = o.random_unit_vectors(rng, nq, dim)
    lo, hi = par.shard_range(n, rank, world)
    shard = np.ascontiguousarray(base[lo:hi])

    def local_topk(qs, kk):
        out = np.empty((len(qs), kk), np.int64)
        for i in range(len(qs)):
            L.jvo_bruteforce_topk_f32(o.DOT_PRODUCT, o.fp(shard), hi - lo, dim, o.fp(np.ascontiguousarray(qs[i])), kk, o.lp(out[i]))
        return torch.from_numpy(par.rebase_keys_host(out, lo))

    res = par.ShardedBruteForce(dist, local_topk).search(queries, k).numpy()
    want = np.empty((nq, k), np.int64)
    for i in range(nq):
        L.jvo_bruteforce_topk_f32(o.DOT_PRODUCT, o.fp(base), n, dim, o.fp(queries[i]), k, o.lp(want[i]))
    ok = bool(np.array_equal(res, want))
    nodes, scores = par.keys_to_nodes_scores(res)
    ok = ok and nodes.min() >= 0 and nodes.max() < n and bool((np.diff(scores, axis=1) <= 0).all())
    with open(os.path.join(tmp, "rank%d.ok" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_topk_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / ("rank%d.ok" % r)).read() == "1"


def test_shard_ranges_and_rebase():
    from jvector_b200 import parallel as par
    for n in (1, 7, 8, 1000001):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    import oracle_lib as o
    L = o.load()
    k = np.array([L.jvo_topk_key(0.25, 5), par.KEY_MIN], np.int64)
    r = par.rebase_keys_host(k, 1000)
    assert r[0] == L.jvo_topk_key(0.25, 1005) and r[1] == par.KEY_MIN
    g = np.array([[[L.jvo_topk_key(0.5, 9), L.jvo_topk_key(0.1, 2)]], [[L.jvo_topk_key(0.5, 4), par.KEY_MIN]]], np.int64)
    m = par.merge_keys_host(g, 3)
    nodes, scores = par.keys_to_nodes_scores(m)
    assert nodes.tolist() == [[4, 9, 2]] and scores[0, 0] == np.float32(0.5)


def _exchange_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import ctypes as C

    import torch
    import torch.distributed as dist

    from jvector_b200 import parallel as par
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = par.SliceExchange(dist, "cpu")
    ok = True
    for count, width in ((1, 3), (2, 3), (5, 4), (8, 4), (7, 64), (16384, 32)):
        def produce(lo, hi, rows_ptr, deg_ptr):
            # what jv_builder_insert_slice does: write rows lo .. hi-1 of a virtual [count][width] array and of a [count] array
            rows = (C.c_int32 * (count * width)).from_address(rows_ptr.value) if lo == 0 else None
            for pos in range(lo, hi):
                row = (C.c_int32 * width).from_address(rows_ptr.value + pos * width * 4)
                for j in range(width):
                    row[j] = pos * 1000 + j
                C.c_int32.from_address(deg_ptr.value + pos * 4).value = pos + 7
            return 0
        rows, deg = ex.run(count, width, produce)
        want_rows = (torch.arange(count)[:, None] * 1000 + torch.arange(width)[None, :]).to(torch.int32)
        ok = ok and bool(torch.equal(rows[:count], want_rows)) and bool(torch.equal(deg[:count, 0], (torch.arange(count) + 7).to(torch.int32)))
    with open(os.path.join(tmp, "ex%d.ok" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_slice_exchange_gloo_world2(tmp_path):
    # the exchange step of the sharded build (jvector_b200/parallel.py SliceExchange): every rank produces its slice through the
    # pointer contract of jv_builder_insert_slice, after the all-gather every rank holds all positions in order — including
    # counts that do not divide by the world size and counts smaller than it
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / ("ex%d.ok" % r)).read() == "1"
