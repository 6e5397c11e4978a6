// Compiled by tests/test_cpp_mirror.py against include/jvector_b200.hpp + libjvector_b200.so (+ the oracle, as the CHECKER).
// Reads like the reference's TestVectorizationProvider / TestVectorGraph: random unit vectors, provider vs scalar reference.
//   no GPU : jv::init must throw jv::Error with JV_ERR_NO_DEVICE (no CPU fallback)            -> prints NO_DEVICE_OK
//   GPU    : scores within 1e-5 of the oracle, device-built graph recall vs brute force        -> prints PARITY_OK
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <set>
#include <vector>

#include "jvector_b200.hpp"
extern "C" {
#include "jv_oracle.h"
}

int main()
{
    try {
        jv::init(0);
    } catch (const jv::Error &e) {
        if (e.code == JV_ERR_NO_DEVICE) { std::puts("NO_DEVICE_OK"); return 0; }
        std::printf("unexpected error: %s\n", e.what());
        return 2;
    }
    const int n = 3000, dim = 96, nq = 32;
    std::mt19937 rng(5);
    std::normal_distribution<float> nd;
    std::vector<float> base((size_t)n * dim), queries((size_t)nq * dim);
    auto fill = [&](std::vector<float> &v, int rows) {
        for (int r = 0; r < rows; r++) {
            double s = 0;
            for (int j = 0; j < dim; j++) { v[(size_t)r * dim + j] = nd(rng); s += (double)v[(size_t)r * dim + j] * v[(size_t)r * dim + j]; }
            for (int j = 0; j < dim; j++) v[(size_t)r * dim + j] /= (float)std::sqrt(s);
        }
    };
    fill(base, n);
    fill(queries, nq);
    jv::F32Vectors vec(base.data(), n, dim);
    for (auto vsf : {jv::VectorSimilarityFunction::EUCLIDEAN, jv::VectorSimilarityFunction::DOT_PRODUCT, jv::VectorSimilarityFunction::COSINE}) {
        jv::ScoreFunction sf(vec, queries.data(), vsf);
        std::vector<int32_t> ids;
        for (int i = 0; i < 257; i++) ids.push_back((i * 37) % n);
        auto got = sf.similarityToBatch(ids);
        for (size_t i = 0; i < ids.size(); i++) {
            const float want = jvo_compare_f32((int)vsf, queries.data(), base.data() + (size_t)ids[i] * dim, dim);
            if (std::fabs(got[i] - want) > 1e-5f * std::fmax(std::fabs(want), 1e-2f)) { std::printf("score mismatch %g vs %g\n", got[i], want); return 3; }
        }
        if (std::fabs(sf.similarityTo(ids[3]) - got[3]) != 0.f) { std::puts("similarityTo != batch"); return 4; }
    }
    // device-built graph, device traversal vs the oracle's GraphSearcher restatement over the SAME adjacency
    jv::GraphIndexBuilder builder(jv::VectorSimilarityFunction::DOT_PRODUCT, 16, 60, 1.2f, 1.2f, false, 7);
    jv::GraphIndex g = builder.build(vec);
    jv::GraphSearcher searcher(g);
    auto res = searcher.search(vec, queries.data(), nq, jv::VectorSimilarityFunction::DOT_PRODUCT, 10, 40);
    std::vector<int32_t> adj = g.adjacency(0);
    jvo_graph og{};
    og.n = n; og.levels = 1; og.degree = g.maxDegree(); og.entry_node = 0; og.entry_level = 0; og.adj0 = adj.data();
    int same = 0, hits = 0;
    for (int q = 0; q < nq; q++) {
        jvo_scorer *sf = jvo_scorer_f32(JVO_DOT_PRODUCT, base.data(), n, dim, queries.data() + (size_t)q * dim);
        jvo_scorer_set_order(sf, 1);  // the kernels' summation order: the traversal must then agree id for id
        int32_t on[10];
        float os[10];
        jvo_graph_search(&og, sf, nullptr, 10, 40, on, os, nullptr);
        jvo_scorer_free(sf);
        bool eq = true;
        for (int i = 0; i < 10; i++) eq = eq && on[i] == res.nodes[(size_t)q * 10 + i];
        same += eq;
        std::vector<int64_t> keys(10);
        jvo_bruteforce_topk_f32(JVO_DOT_PRODUCT, base.data(), n, dim, queries.data() + (size_t)q * dim, 10, keys.data());
        std::set<int32_t> truth;
        for (auto k : keys) truth.insert(jvo_key_node(k));
        for (int i = 0; i < 10; i++) hits += truth.count(res.nodes[(size_t)q * 10 + i]);
    }
    if (same != nq) { std::printf("device traversal disagrees with the oracle: %d/%d\n", same, nq); return 5; }
    bool threw = false;
    try { jv::F32Vectors bad(nullptr, 0, 0); } catch (const jv::Error &e) { threw = e.code == JV_ERR_INVALID; }
    if (!threw) { std::puts("bad arguments did not throw"); return 6; }
    // ---- one process, every visible GPU (what a single JVM needs; with one device these are the 1-shard degenerates) ----
    const int ndev = std::min(jv_gpu_device_count(), 8);
    if (jv_gpu_init_mask((1u << ndev) - 1u) != JV_OK) { std::printf("init_mask: %s\n", jv_last_error()); return 7; }
    {
        // range-sharded BQ brute force: per-device streams, peer copies of the shard keys, merge on the first device
        const int bn = 40000, bdim = 512, bq = 19, k = 30, W = bdim / 64;
        std::vector<float> rows((size_t)bn * bdim), qs((size_t)bq * bdim);
        for (auto &x : rows) x = nd(rng);
        for (auto &x : qs) x = nd(rng);
        std::vector<uint64_t> words((size_t)bn * W), qw((size_t)bq * W);
        for (int i = 0; i < bn; i++) jvo_bq_encode(rows.data() + (size_t)i * bdim, bdim, words.data() + (size_t)i * W);
        for (int i = 0; i < bq; i++) jvo_bq_encode(qs.data() + (size_t)i * bdim, bdim, qw.data() + (size_t)i * W);
        jv_multi m = nullptr;
        if (jv_multi_register_bq(words.data(), bn, bdim, &m) != JV_OK) { std::printf("multi_register: %s\n", jv_last_error()); return 8; }
        if (jv_multi_shard_count(m) != ndev) { std::puts("shard count"); return 9; }
        std::vector<int64_t> got((size_t)bq * k), want((size_t)bq * k);
        if (jv_multi_topk_bruteforce(m, JV_COSINE, qs.data(), bq, k, got.data()) != JV_OK) { std::printf("multi_topk: %s\n", jv_last_error()); return 10; }
        jvo_bq_bruteforce_batch(words.data(), bn, bdim, qw.data(), bq, k, 8, want.data());
        if (got != want) { std::puts("multi-device brute force keys differ from the oracle"); return 11; }
        jv_multi_free(m);
    }
    {
        // replicas of graph + rows on every device, the query batch split across them
        std::vector<jv_dataset> dsets(ndev);
        std::vector<jv_graph> graphs(ndev);
        for (int d = 0; d < ndev; d++) {
            if (jv_gpu_set_device(d) != JV_OK || jv_dataset_register_f32(base.data(), n, dim, &dsets[d]) != JV_OK ||
                jv_graph_create(n, g.maxDegree(), adj.data(), 0, &graphs[d]) != JV_OK) { std::printf("replica %d: %s\n", d, jv_last_error()); return 12; }
        }
        jv_gpu_set_device(0);
        std::vector<int32_t> mn((size_t)nq * 10);
        std::vector<float> ms((size_t)nq * 10);
        jv_search_stats st;
        if (jv_multi_graph_search_batch(ndev, graphs.data(), dsets.data(), nullptr, JV_DOT_PRODUCT, queries.data(), nq, 10, 40, nullptr, mn.data(), ms.data(), &st) != JV_OK) {
            std::printf("multi_graph_search: %s\n", jv_last_error());
            return 13;
        }
        if (mn != res.nodes) { std::puts("replica search differs from the single-device search"); return 14; }
        for (int d = 0; d < ndev; d++) { jv_graph_free(graphs[d]); jv_dataset_free(dsets[d]); }
    }
    std::printf("PARITY_OK agree=%d/%d recall=%.3f visited=%lld devices=%d\n", same, nq, hits / (10.0 * nq), (long long)res.visitedCount, ndev);
    return 0;
}
