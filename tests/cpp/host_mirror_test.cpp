// Compiled by tests/test_cpp_mirror.py against include/jvector_b200.hpp + libjvector_b200.so (+ the oracle, as the CHECKER).
// Reads like the reference's TestVectorizationProvider / TestVectorGraph: random unit vectors, provider vs scalar reference.
//   no GPU : jv::init must throw jv::Error with JV_ERR_NO_DEVICE (no CPU fallback)            -> prints NO_DEVICE_OK
//   GPU    : scores within 1e-5 of the oracle, device-built graph recall vs brute force        -> prints PARITY_OK
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
    if (same < nq - 2) { std::printf("device traversal disagrees with the oracle: %d/%d\n", same, nq); return 5; }
    bool threw = false;
    try { jv::F32Vectors bad(nullptr, 0, 0); } catch (const jv::Error &e) { threw = e.code == JV_ERR_INVALID; }
    if (!threw) { std::puts("bad arguments did not throw"); return 6; }
    std::printf("PARITY_OK agree=%d/%d recall=%.3f visited=%lld\n", same, nq, hits / (10.0 * nq), (long long)res.visitedCount);
    return 0;
}
