// build.cu — GraphIndexBuilder.build on the device, exact fp32 scoring
// (base:graph/GraphIndexBuilder.java:436-448 build, :605-671 addGraphNode / updateNeighborsOneLayer;
//  base:graph/OnHeapGraphIndex.java:279-282 addEdges; base:graph/ConcurrentNeighborMap.java:104-110 insertDiverse,
//  :158-165 backlink, :286-321 insert with overflow, :214-222 enforceDegree;
//  base:graph/diversity/VamanaDiversityProvider.java:45-95 retainDiverse / isDiverse).
//
// The reference inserts nodes concurrently from a ForkJoinPool; every insert is (1) a beam search of the graph as it
// stands, (2) a Vamana robust prune of the beam, (3) back-links into the chosen neighbours, re-pruning a neighbour
// whose list exceeds overflow * M. Here a BATCH of nodes plays the role of the concurrently inserting threads:
//   search  : graph_search_kernel (search.cu) over the batch, one CTA per inserted node, beam = efConstruction
//   prune   : prune_kernel, one CTA per node: scores -> sort by the reference key -> retainDiverse restated over a running
//             max-similarity per candidate, updated in bulk whenever a neighbour is selected (SURVEY §3.2 seam v)
//   backlink: append u to adj[v] for every chosen v (atomic slot), then prune_kernel again over every v whose list
//             passed overflow * M
// Batches grow geometrically (x 33/32, JV_BUILD_BATCH_DIV) up to max_batch so early nodes see a connected graph. Neighbour lists in the
// reference are concurrency-order dependent, so parity is on scores and on the recall of the resulting graph
// (SURVEY §8d C5), not on identical adjacency.
#include <limits.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <cub/cub.cuh>

#include "kernels.h"

namespace jv {

constexpr int PRUNE_THREADS = 256;
constexpr int PRUNE_MAXC = 256;  // candidates per prune (beam or list capacity), power of two

struct PruneParams {
    DataDesc d;
    int metric;
    int degree;     // M: max selected
    int row_cap;    // slots per adjacency row
    float alpha;
    int32_t *adj;   // [n][row_cap]
    int *deg;       // [n]
    // mode 0: nodes = node_base + i, candidates = cand[i][0..cand_cnt) (search results, -1 padded)
    // mode 1: nodes = list[i], candidates = adj[v][0..min(deg, row_cap))
    int mode;
    int node_base;
    const int32_t *list;
    const int *count_ptr;  // number of nodes (device) or nullptr -> count
    int count;
    const int32_t *cand;
    int cand_stride;
    int *mark;  // cleared for pruned nodes in mode 1
    // mode 0: the W nodes around a batch node in insertion order are candidates too — the batch plays the reference's set of
    // concurrently inserting threads, whose in-progress nodes every insert scores and merges into its candidates
    // (GraphIndexBuilder.java:611-612,669,823-837 getConcurrentCandidates). Batch = nodes [batch_first, batch_first + batch_count).
    int window, batch_first, batch_count;
    // out-of-place results (sharded build: the slice's rows travel before they are applied): row i of this launch goes to
    // out_rows[(out_base + i) * out_stride ..] and out_deg[out_base + i]; nullptr = write adj / deg in place
    int32_t *out_rows;
    int *out_deg;
    int out_stride, out_base;
    int list_base;  // mode 1: nodes = list[list_base + i]
};

// candidate i of node v: the search result while i < cand_stride, then the in-progress window (skipping v itself)
__device__ __forceinline__ int32_t prune_candidate(const PruneParams &P, const int32_t *cand, int nc_search, int v, int i)
{
    if (i < nc_search) return cand[i];
    if (P.mode != 0 || P.window <= 0) return -1;
    const int j = i - nc_search;  // 0 .. window-1
    const int half = P.window >> 1;
    int w = v - half + j;
    if (w >= v) w += 1;  // skip v: the window holds `window` OTHER nodes
    if (w < P.batch_first || w >= P.batch_first + P.batch_count) return -1;
    return w;
}

__device__ __forceinline__ void bitonic_sort_desc_prune(long long *keys, int n_pow2)
{
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const long long a = keys[i], b = keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// similarity score of candidate row `c` (global) against a cached row `srow` (shared or global), one warp
template <int METRIC>
__device__ __forceinline__ float pair_rows(const float4 *__restrict__ c, const float4 *srow, int n4, int lane)
{
    float s0 = 0.f, s1 = 0.f, aa = 0.f, bb = 0.f;
#pragma unroll 2
    for (int i = lane; i < n4; i += 32) {
        const float4 x = __ldg(c + i);
        const float4 y = srow[i];
        if (METRIC == JV_METRIC_EUCLIDEAN) {
            float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
            s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s0 = fmaf(d2, d2, s0); s1 = fmaf(d3, d3, s1);
        } else {
            s0 = fmaf(x.x, y.x, s0); s1 = fmaf(x.y, y.y, s1); s0 = fmaf(x.z, y.z, s0); s1 = fmaf(x.w, y.w, s1);
            if (METRIC == JV_METRIC_COSINE) {
                aa = fmaf(x.x, x.x, aa); aa = fmaf(x.y, x.y, aa); aa = fmaf(x.z, x.z, aa); aa = fmaf(x.w, x.w, aa);
                bb = fmaf(y.x, y.x, bb); bb = fmaf(y.y, y.y, bb); bb = fmaf(y.z, y.z, bb); bb = fmaf(y.w, y.w, bb);
            }
        }
    }
    float s = group_sum<32>(s0 + s1);
    if (METRIC == JV_METRIC_COSINE) {
        aa = group_sum<32>(aa);
        bb = group_sum<32>(bb);
        s = __fdiv_rn(s, __fsqrt_rn(__fmul_rn(aa, bb)));
    }
    return score_map(METRIC, s);
}

// retainDiverse (VamanaDiversityProvider.java:45-95), restated so that the pair scores can be produced in bulk:
//   isDiverse(c) at alpha  <=>  max over the selected s of sim(c, s)  <=  score(c) * alpha.
// Keep M[c] = that running maximum. Selecting a candidate s updates M[c] for every still-eligible c in parallel (one warp
// per candidate, the new row cached in shared memory); the sequential part of the reference loop shrinks to "find the next
// c at or after the cursor with M[c] <= score(c) * alpha", a flag scan. A candidate whose M[c] already exceeds
// score(c) * alpha_max can never be selected in any pass (M only grows) and is dropped from further updates. The selection
// sequence — and therefore the neighbour list — is exactly the reference loop's, including the alpha = 1.0, 1.2 passes.
template <int METRIC>
__global__ void __launch_bounds__(PRUNE_THREADS) prune_kernel(PruneParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *blob = reinterpret_cast<float *>(smem_raw);                       // stride + 4
    float *newrow = blob + P.d.stride + 4;                                   // stride: the most recently selected row
    long long *keys = reinterpret_cast<long long *>(newrow + P.d.stride);    // PRUNE_MAXC
    float *maxsim = reinterpret_cast<float *>(keys + PRUNE_MAXC);            // PRUNE_MAXC
    int32_t *sel_ids = reinterpret_cast<int32_t *>(maxsim + PRUNE_MAXC);     // 128
    uint8_t *state = reinterpret_cast<uint8_t *>(sel_ids + 128);             // PRUNE_MAXC: 0 eligible, 1 selected, 2 dropped
    __shared__ float red[36];
    __shared__ int s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PRUNE_THREADS / 32;
    const int total = P.count_ptr ? *P.count_ptr : P.count;
    const int n4 = P.d.stride >> 2;
    // the last alpha the reference loop reaches: 1.0, 1.2, ... while <= alpha + 1e-6
    float alpha_max = 1.0f;
    for (float a = 1.0f; a <= P.alpha + 1e-6f; a += 0.2f) alpha_max = a;

    for (int it = blockIdx.x; it < total; it += gridDim.x) {
        const int v = P.mode == 0 ? P.node_base + it : P.list[P.list_base + it];
        const int32_t *cand;
        int nc, ncs;
        if (P.mode == 0) {
            cand = P.cand + (size_t)it * P.cand_stride;
            ncs = P.cand_stride;
            nc = ncs + P.window;
        } else {
            cand = P.adj + (size_t)v * P.row_cap;
            ncs = nc = min(P.deg[v], P.row_cap);
        }
        nc = min(nc, PRUNE_MAXC);
        prepare_blob(P.d, P.metric, P.d.rows + (size_t)v * P.d.stride, blob, red);
        // exact scores of the candidates against v, as sortable keys (two rows per warp at a time)
        for (int i = warp; i < PRUNE_MAXC; i += 2 * NW) {
            const int i2 = i + NW;
            const int32_t ca = i < nc ? prune_candidate(P, cand, ncs, v, i) : -1, cb = i2 < nc ? prune_candidate(P, cand, ncs, v, i2) : -1;
            const bool va = ca >= 0 && ca != v, vb = cb >= 0 && cb != v;
            long long ka = KEY_MIN, kb = KEY_MIN;
            if (va || vb) {
                float sa, sb;
                score_f32_pair<METRIC>(P.d, blob, va ? ca : cb, vb ? cb : ca, lane, sa, sb);
                if (va) ka = topk_key(sa, ca);
                if (vb) kb = topk_key(sb, cb);
            }
            if (lane == 0) {
                keys[i] = ka;
                if (i2 < PRUNE_MAXC) keys[i2] = kb;
            }
        }
        __syncthreads();
        bitonic_sort_desc_prune(keys, PRUNE_MAXC);
        // valid keys form a prefix (KEY_MIN sorts last); duplicates are adjacent after the sort and are dropped
        int nvalid;
        {
            int lo = 0, hi = PRUNE_MAXC;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (keys[mid] != KEY_MIN) lo = mid + 1;
                else hi = mid;
            }
            nvalid = lo;
        }
        for (int i = tid; i < PRUNE_MAXC; i += PRUNE_THREADS) {
            maxsim[i] = -3.0e38f;
            state[i] = (i < nvalid && !(i > 0 && keys[i - 1] == keys[i])) ? 0 : 2;
        }
        __syncthreads();

        int nsel = 0;
        float currentAlpha = 1.0f;
        while (currentAlpha <= P.alpha + 1e-6f && nsel < P.degree) {
            int cursor = 0;
            while (nsel < P.degree) {
                // next eligible candidate at or after the cursor that is diverse at this alpha
                if (tid == 0) s_next = INT_MAX;
                __syncthreads();
                for (int i = cursor + tid; i < nvalid; i += PRUNE_THREADS)
                    if (state[i] == 0 && !(maxsim[i] > __fmul_rn(key_score(keys[i]), currentAlpha))) { atomicMin(&s_next, i); break; }
                __syncthreads();
                const int pick = s_next;
                if (pick == INT_MAX) break;
                const int32_t c = key_node(keys[pick]);
                const float4 *crow = reinterpret_cast<const float4 *>(P.d.rows + (size_t)c * P.d.stride);
                if (tid == 0) { state[pick] = 1; sel_ids[nsel] = c; }
                for (int t = tid; t < n4; t += PRUNE_THREADS) reinterpret_cast<float4 *>(newrow)[t] = __ldg(crow + t);
                nsel++;
                cursor = pick + 1;
                __syncthreads();
                if (nsel >= P.degree) break;
                // fold the new neighbour into every still-eligible candidate's running maximum
                for (int i = warp; i < nvalid; i += NW) {
                    if (state[i] != 0) continue;
                    const float4 *row = reinterpret_cast<const float4 *>(P.d.rows + (size_t)key_node(keys[i]) * P.d.stride);
                    const float ps = pair_rows<METRIC>(row, reinterpret_cast<const float4 *>(newrow), n4, lane);
                    if (lane == 0) {
                        const float m = fmaxf(maxsim[i], ps);
                        maxsim[i] = m;
                        if (m > __fmul_rn(key_score(keys[i]), alpha_max)) state[i] = 2;  // can never become diverse
                    }
                }
                __syncthreads();
            }
            currentAlpha += 0.2f;
        }
        // the selected neighbours, in score order (NodeArray.retain keeps the sorted order)
        __syncthreads();
        if (tid == 0) {
            int w = 0;
            int32_t *row = P.out_rows ? P.out_rows + (size_t)(P.out_base + it) * P.out_stride : P.adj + (size_t)v * P.row_cap;
            const int width = P.out_rows ? P.out_stride : P.row_cap;
            for (int i = 0; i < nvalid; i++)
                if (state[i] == 1) row[w++] = key_node(keys[i]);
            for (int i = w; i < width; i++) row[i] = -1;
            if (P.out_rows) P.out_deg[P.out_base + it] = w;
            else P.deg[v] = w;
            if (P.mode == 1 && P.mark) P.mark[v] = 0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// prune_gram_kernel: the same retainDiverse, with the candidate x candidate similarity matrix produced up front by a
// shared-memory tiled Gram product on the CUDA cores (SURVEY §3.2 seam v: "compute the full candidate x candidate matrix
// and replay the sequential selection"). Rows are read ONCE (nc x 3 KB from L2) instead of once per selected neighbour;
// the selection loop then only touches shared memory. TILE = 128 candidates (beam) or 64 (re-prune of an adjacency row).
// Thread (ty, tx) of the 16 x 16 grid owns C[ty + 16 a][tx + 16 b]; K is walked in 32-float chunks staged in shared memory
// (row pitch 36 floats: 128-bit loads of a quarter-warp hit distinct banks).
// ------------------------------------------------------------------------------------------------
constexpr int GRAM_KC = 32;
constexpr int GRAM_PITCH = 36;

template <int METRIC, int TILE>
__global__ void __launch_bounds__(PRUNE_THREADS) prune_gram_kernel(PruneParams P)
{
    constexpr int NB = TILE / 16;       // rows / cols per thread
    constexpr int CP = TILE + 1;        // pitch of the similarity matrix
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *blob = reinterpret_cast<float *>(smem_raw);                         // stride + 4
    float *tileA = blob + P.d.stride + 4;                                      // TILE * GRAM_PITCH
    float *C = tileA + TILE * GRAM_PITCH;                                      // TILE * CP
    long long *keys = reinterpret_cast<long long *>(C + TILE * CP + ((TILE * CP) & 1));  // TILE (8-byte aligned)
    float *maxsim = reinterpret_cast<float *>(keys + TILE);                    // TILE
    int32_t *ids = reinterpret_cast<int32_t *>(maxsim + TILE);                 // TILE
    uint8_t *state = reinterpret_cast<uint8_t *>(ids + TILE);                  // TILE
    __shared__ float red[36];
    __shared__ int s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    constexpr int NW = PRUNE_THREADS / 32;
    const int total = P.count_ptr ? *P.count_ptr : P.count;
    float alpha_max = 1.0f;
    for (float a = 1.0f; a <= P.alpha + 1e-6f; a += 0.2f) alpha_max = a;

    for (int it = blockIdx.x; it < total; it += gridDim.x) {
        const int v = P.mode == 0 ? P.node_base + it : P.list[P.list_base + it];
        const int32_t *cand;
        int nc, ncs;
        if (P.mode == 0) {
            cand = P.cand + (size_t)it * P.cand_stride;
            ncs = P.cand_stride;
            nc = ncs + P.window;
        } else {
            cand = P.adj + (size_t)v * P.row_cap;
            ncs = nc = min(P.deg[v], P.row_cap);
        }
        nc = min(nc, TILE);
        prepare_blob(P.d, P.metric, P.d.rows + (size_t)v * P.d.stride, blob, red);
        for (int i = warp; i < TILE; i += 2 * NW) {
            const int i2 = i + NW;
            const int32_t ca = i < nc ? prune_candidate(P, cand, ncs, v, i) : -1, cb = (i2 < TILE && i2 < nc) ? prune_candidate(P, cand, ncs, v, i2) : -1;
            const bool va = ca >= 0 && ca != v, vb = cb >= 0 && cb != v;
            long long ka = KEY_MIN, kb = KEY_MIN;
            if (va || vb) {
                float sa, sb;
                score_f32_pair<METRIC>(P.d, blob, va ? ca : cb, vb ? cb : ca, lane, sa, sb);
                if (va) ka = topk_key(sa, ca);
                if (vb) kb = topk_key(sb, cb);
            }
            if (lane == 0) {
                keys[i] = ka;
                if (i2 < TILE) keys[i2] = kb;
            }
        }
        __syncthreads();
        bitonic_sort_desc_prune(keys, TILE);
        int nvalid;
        {
            int lo = 0, hi = TILE;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (keys[mid] != KEY_MIN) lo = mid + 1;
                else hi = mid;
            }
            nvalid = lo;
        }
        for (int i = tid; i < TILE; i += PRUNE_THREADS) {
            maxsim[i] = -3.0e38f;
            state[i] = (i < nvalid && !(i > 0 && keys[i - 1] == keys[i])) ? 0 : 2;
            ids[i] = i < nvalid ? key_node(keys[i]) : -1;
        }
        __syncthreads();

        // ---- Gram matrix of the candidate rows ----
        float acc[NB][NB];
#pragma unroll
        for (int a = 0; a < NB; a++)
#pragma unroll
            for (int b = 0; b < NB; b++) acc[a][b] = 0.f;
        for (int k0 = 0; k0 < P.d.stride; k0 += GRAM_KC) {
            // stage rows [0, TILE) x [k0, k0 + KC): TILE * 8 float4, zero beyond nvalid / beyond the row
            for (int t = tid; t < TILE * (GRAM_KC / 4); t += PRUNE_THREADS) {
                const int r = t / (GRAM_KC / 4), c4 = t - r * (GRAM_KC / 4);
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                const int32_t node = ids[r];
                if (node >= 0 && k0 + 4 * c4 < P.d.stride) val = __ldg(reinterpret_cast<const float4 *>(P.d.rows + (size_t)node * P.d.stride + k0) + c4);
                *reinterpret_cast<float4 *>(tileA + r * GRAM_PITCH + 4 * c4) = val;
            }
            __syncthreads();
#pragma unroll
            for (int k4 = 0; k4 < GRAM_KC / 4; k4++) {
                float4 av[NB], bv[NB];
#pragma unroll
                for (int a = 0; a < NB; a++) av[a] = *reinterpret_cast<const float4 *>(tileA + (ty + 16 * a) * GRAM_PITCH + 4 * k4);
#pragma unroll
                for (int b = 0; b < NB; b++) bv[b] = *reinterpret_cast<const float4 *>(tileA + (tx + 16 * b) * GRAM_PITCH + 4 * k4);
#pragma unroll
                for (int a = 0; a < NB; a++)
#pragma unroll
                    for (int b = 0; b < NB; b++) {
                        acc[a][b] = fmaf(av[a].x, bv[b].x, acc[a][b]);
                        acc[a][b] = fmaf(av[a].y, bv[b].y, acc[a][b]);
                        acc[a][b] = fmaf(av[a].z, bv[b].z, acc[a][b]);
                        acc[a][b] = fmaf(av[a].w, bv[b].w, acc[a][b]);
                    }
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < NB; a++)
#pragma unroll
            for (int b = 0; b < NB; b++) C[(ty + 16 * a) * CP + tx + 16 * b] = acc[a][b];
        __syncthreads();
        // raw dot products -> the reference's similarity scores (needs the diagonal for L2 / cosine)
        if (METRIC != JV_METRIC_DOT) {
            float diag_r[NB], diag_c[NB];
#pragma unroll
            for (int a = 0; a < NB; a++) diag_r[a] = C[(ty + 16 * a) * CP + ty + 16 * a];
#pragma unroll
            for (int b = 0; b < NB; b++) diag_c[b] = C[(tx + 16 * b) * CP + tx + 16 * b];
            __syncthreads();
#pragma unroll
            for (int a = 0; a < NB; a++)
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    float raw;
                    if (METRIC == JV_METRIC_EUCLIDEAN) raw = fmaxf(0.f, __fadd_rn(__fadd_rn(diag_r[a], diag_c[b]), -2.0f * acc[a][b]));
                    else raw = __fdiv_rn(acc[a][b], __fsqrt_rn(__fmul_rn(diag_r[a], diag_c[b])));
                    C[(ty + 16 * a) * CP + tx + 16 * b] = score_map(METRIC, raw);
                }
        } else {
#pragma unroll
            for (int a = 0; a < NB; a++)
#pragma unroll
                for (int b = 0; b < NB; b++) C[(ty + 16 * a) * CP + tx + 16 * b] = score_map(METRIC, acc[a][b]);
        }
        __syncthreads();

        // ---- the sequential selection, now over shared memory only ----
        int nsel = 0;
        float currentAlpha = 1.0f;
        while (currentAlpha <= P.alpha + 1e-6f && nsel < P.degree) {
            int cursor = 0;
            while (nsel < P.degree) {
                if (tid == 0) s_next = INT_MAX;
                __syncthreads();
                for (int i = cursor + tid; i < nvalid; i += PRUNE_THREADS)
                    if (state[i] == 0 && !(maxsim[i] > __fmul_rn(key_score(keys[i]), currentAlpha))) { atomicMin(&s_next, i); break; }
                __syncthreads();
                const int pick = s_next;
                if (pick == INT_MAX) break;
                nsel++;
                cursor = pick + 1;
                for (int i = tid; i < nvalid; i += PRUNE_THREADS) {
                    if (i == pick) state[i] = 1;
                    else if (state[i] == 0) {
                        const float m = fmaxf(maxsim[i], C[i * CP + pick]);
                        maxsim[i] = m;
                        if (m > __fmul_rn(key_score(keys[i]), alpha_max)) state[i] = 2;
                    }
                }
                __syncthreads();
            }
            currentAlpha += 0.2f;
        }
        __syncthreads();
        if (tid == 0) {
            int w = 0;
            int32_t *row = P.out_rows ? P.out_rows + (size_t)(P.out_base + it) * P.out_stride : P.adj + (size_t)v * P.row_cap;
            const int width = P.out_rows ? P.out_stride : P.row_cap;
            for (int i = 0; i < nvalid; i++)
                if (state[i] == 1) row[w++] = ids[i];
            for (int i = w; i < width; i++) row[i] = -1;
            if (P.out_rows) P.out_deg[P.out_base + it] = w;
            else P.deg[v] = w;
            if (P.mode == 1 && P.mark) P.mark[v] = 0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// prune_gram_tc_kernel: the same kernel with the Gram product on the 5th-generation tensor cores. The candidate x candidate
// matrix is GEMM-shaped work (X X^T, X = the candidate rows), 72 % of prune_gram_kernel<128>'s time on the CUDA cores
// (profiles/r2_build_prune.md). fp32 products are kept by the 3xTF32 split: x = hi + lo with hi = x truncated to tf32's 10
// mantissa bits and lo = x - hi (exact), D += hi hi^T + hi lo^T + lo hi^T (the dropped lo lo^T term is 2^-22 relative), fp32
// accumulation in TMEM. Per 32-float K chunk all 256 threads stage the rows into two K-major SWIZZLE_128B tiles (hi | lo, TILE
// rows x 128 B each; the same tile is the A and the B operand), double buffered, and one thread issues 12
// tcgen05.mma.kind::tf32 (M = 128, N = TILE, K = 8); the next chunk's global loads are in flight while it does. The similarity
// matrix C overlays the staging area once the last MMA has completed. The scores against v (the sort keys) stay exact fp32.
// For TILE = 64 the A operand still spans 128 rows (M = 128): rows 64..127 read whatever follows the tile and produce
// accumulator rows that are never read.
// ------------------------------------------------------------------------------------------------
namespace tc {
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bounded(uint64_t *bar, unsigned parity)
{
    for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) __trap();
}
template <int TILE>
struct Layout {
    static constexpr int TILE_BYTES = TILE * 128;                  // one operand tile: TILE rows x 32 floats
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES;             // hi | lo
    static constexpr int STAGES_END = 2 * STAGE_BYTES + (TILE < 128 ? (128 - TILE) * 128 : 0);  // + what an M = 128 A operand over-reads
    static constexpr int C_BYTES = (TILE * (TILE + 1) * 4 + 15) & ~15;
    static constexpr int AREA = ((STAGES_END > C_BYTES ? STAGES_END : C_BYTES) + 1023) & ~1023;
};
}  // namespace tc

template <int TILE>
static size_t gram_tc_smem_bytes(const DataDesc &d)
{
    size_t b = 1024 /* alignment slack */ + tc::Layout<TILE>::AREA + (size_t)(d.stride + 4) * 4 + (size_t)TILE * 8 + (size_t)TILE * 4 * 3 + TILE + 64;
    return (b + 15) & ~(size_t)15;
}

template <int METRIC, int TILE>
__global__ void __launch_bounds__(PRUNE_THREADS, TILE == 128 ? 3 : 4) prune_gram_tc_kernel(PruneParams P)
{
    using L = tc::Layout<TILE>;
    constexpr int CP = TILE + 1;
    constexpr int PASSES = TILE / 32;   // 256 threads stage 32 rows x 8 float4 per pass
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *area = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    float *C = reinterpret_cast<float *>(area);
    float *blob = reinterpret_cast<float *>(area + L::AREA);                   // stride + 4
    long long *keys = reinterpret_cast<long long *>(blob + ((P.d.stride + 4 + 1) & ~1));
    float *maxsim = reinterpret_cast<float *>(keys + TILE);
    float *diag = maxsim + TILE;
    int32_t *ids = reinterpret_cast<int32_t *>(diag + TILE);
    uint8_t *state = reinterpret_cast<uint8_t *>(ids + TILE);
    __shared__ float red[36];
    __shared__ int s_next;
    __shared__ __align__(8) uint64_t mma_bar[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PRUNE_THREADS / 32;
    const int total = P.count_ptr ? *P.count_ptr : P.count;
    float alpha_max = 1.0f;
    for (float a = 1.0f; a <= P.alpha + 1e-6f; a += 0.2f) alpha_max = a;

    if (tid == 0) {
        mbar_init(&mma_bar[0], 1);
        mbar_init(&mma_bar[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TILE) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // c = F32 (1) at [4,6), a = b = TF32 (2) at [7,10) / [10,13), K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    unsigned commits0 = 0u, commits1 = 0u;  // commits issued per staging buffer so far (every thread keeps the same count)
    const int nchunks = (P.d.stride + 31) / 32;
    const int c4 = tid & 7, r0 = tid >> 3;

    for (int it = blockIdx.x; it < total; it += gridDim.x) {
        const int v = P.mode == 0 ? P.node_base + it : P.list[P.list_base + it];
        const int32_t *cand;
        int nc, ncs;
        if (P.mode == 0) {
            cand = P.cand + (size_t)it * P.cand_stride;
            ncs = P.cand_stride;
            nc = ncs + P.window;
        } else {
            cand = P.adj + (size_t)v * P.row_cap;
            ncs = nc = min(P.deg[v], P.row_cap);
        }
        nc = min(nc, TILE);
        prepare_blob(P.d, P.metric, P.d.rows + (size_t)v * P.d.stride, blob, red);
        for (int i = warp; i < TILE; i += 2 * NW) {
            const int i2 = i + NW;
            const int32_t ca = i < nc ? prune_candidate(P, cand, ncs, v, i) : -1, cb = (i2 < TILE && i2 < nc) ? prune_candidate(P, cand, ncs, v, i2) : -1;
            const bool va = ca >= 0 && ca != v, vb = cb >= 0 && cb != v;
            long long ka = KEY_MIN, kb = KEY_MIN;
            if (va || vb) {
                float sa, sb;
                score_f32_pair<METRIC>(P.d, blob, va ? ca : cb, vb ? cb : ca, lane, sa, sb);
                if (va) ka = topk_key(sa, ca);
                if (vb) kb = topk_key(sb, cb);
            }
            if (lane == 0) {
                keys[i] = ka;
                if (i2 < TILE) keys[i2] = kb;
            }
        }
        __syncthreads();
        bitonic_sort_desc_prune(keys, TILE);
        int nvalid;
        {
            int lo = 0, hi = TILE;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (keys[mid] != KEY_MIN) lo = mid + 1;
                else hi = mid;
            }
            nvalid = lo;
        }
        for (int i = tid; i < TILE; i += PRUNE_THREADS) {
            maxsim[i] = -3.0e38f;
            state[i] = (i < nvalid && !(i > 0 && keys[i - 1] == keys[i])) ? 0 : 2;
            ids[i] = i < nvalid ? key_node(keys[i]) : -1;
        }
        __syncthreads();

        // ---- Gram matrix of the candidate rows on the tensor cores (3xTF32) ----
        const float *rowp[PASSES];
#pragma unroll
        for (int j = 0; j < PASSES; j++) {
            const int32_t node = ids[r0 + 32 * j];
            rowp[j] = node >= 0 ? P.d.rows + (size_t)node * P.d.stride + 4 * c4 : nullptr;
        }
        float4 pre[PASSES];
#pragma unroll
        for (int j = 0; j < PASSES; j++) pre[j] = (rowp[j] && 4 * c4 < P.d.stride) ? __ldg(reinterpret_cast<const float4 *>(rowp[j])) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < nchunks; c++) {
            const int b = c & 1;
            if (c >= 2) tc::wait_bounded(&mma_bar[b], ((b ? commits1 : commits0) - 1u) & 1u);  // the MMAs of chunk c - 2 have read this buffer
            unsigned char *hi_t = area + b * L::STAGE_BYTES, *lo_t = hi_t + L::TILE_BYTES;
#pragma unroll
            for (int j = 0; j < PASSES; j++) {
                const int r = r0 + 32 * j;
                const float4 x = pre[j];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = __fsub_rn(x.x, h.x);
                h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = __fsub_rn(x.y, h.y);
                h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = __fsub_rn(x.z, h.z);
                h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = __fsub_rn(x.w, h.w);
                const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c4 ^ (r & 7)) << 4);
                *reinterpret_cast<float4 *>(hi_t + off) = h;
                *reinterpret_cast<float4 *>(lo_t + off) = l;
            }
            if (c + 1 < nchunks) {
                const int k1 = (c + 1) * 32 + 4 * c4;
#pragma unroll
                for (int j = 0; j < PASSES; j++)
                    pre[j] = (rowp[j] && k1 < P.d.stride) ? __ldg(reinterpret_cast<const float4 *>(rowp[j] + (c + 1) * 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ah = smem_u32(hi_t), al = smem_u32(lo_t);
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    const uint64_t dh = tc::desc_sw128(ah + 32 * ks), dl = tc::desc_sw128(al + 32 * ks);
                    tc::mma_tf32(tmem, dh, dh, IDESC, (c | ks) != 0 ? 1u : 0u);
                    tc::mma_tf32(tmem, dh, dl, IDESC, 1u);
                    tc::mma_tf32(tmem, dl, dh, IDESC, 1u);
                }
                tc::commit(&mma_bar[b]);
            }
            if (b) commits1++;
            else commits0++;
        }
        if (commits0) tc::wait_bounded(&mma_bar[0], (commits0 - 1u) & 1u);
        if (commits1) tc::wait_bounded(&mma_bar[1], (commits1 - 1u) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();  // every thread is past its waits: the staging area may now be overwritten by C
        {
            // TMEM -> C: warp w reads lanes 32 (w % 4) .. (its quarter of the rows) and half of the columns
            const int quarter = warp & 3, half = warp >> 2;
            if (quarter * 32 < TILE) {
                const int i = quarter * 32 + lane;
#pragma unroll 1
                for (int cg = 0; cg < TILE / 64; cg++) {
                    const int col0 = half * (TILE / 2) + cg * 32;
                    uint32_t t[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]),
                          "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]), "=r"(t[16]), "=r"(t[17]), "=r"(t[18]), "=r"(t[19]), "=r"(t[20]),
                          "=r"(t[21]), "=r"(t[22]), "=r"(t[23]), "=r"(t[24]), "=r"(t[25]), "=r"(t[26]), "=r"(t[27]), "=r"(t[28]), "=r"(t[29]), "=r"(t[30]),
                          "=r"(t[31])
                        : "r"(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j++) C[i * CP + col0 + j] = __uint_as_float(t[j]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // raw dot products -> the reference's similarity scores (needs the diagonal for L2 / cosine)
        if (METRIC != JV_METRIC_DOT) {
            for (int i = tid; i < TILE; i += PRUNE_THREADS) diag[i] = C[i * CP + i];
            __syncthreads();
        }
        for (int idx = tid; idx < TILE * TILE; idx += PRUNE_THREADS) {
            const int i = idx / TILE, j = idx - i * TILE;
            const float dotv = C[i * CP + j];
            float raw;
            if (METRIC == JV_METRIC_EUCLIDEAN) raw = fmaxf(0.f, __fadd_rn(__fadd_rn(diag[i], diag[j]), -2.0f * dotv));
            else if (METRIC == JV_METRIC_COSINE) raw = __fdiv_rn(dotv, __fsqrt_rn(__fmul_rn(diag[i], diag[j])));
            else raw = dotv;
            C[i * CP + j] = score_map(METRIC, raw);
        }
        __syncthreads();

        // ---- the sequential selection, over shared memory only ----
        int nsel = 0;
        float currentAlpha = 1.0f;
        while (currentAlpha <= P.alpha + 1e-6f && nsel < P.degree) {
            int cursor = 0;
            while (nsel < P.degree) {
                if (tid == 0) s_next = INT_MAX;
                __syncthreads();
                for (int i = cursor + tid; i < nvalid; i += PRUNE_THREADS)
                    if (state[i] == 0 && !(maxsim[i] > __fmul_rn(key_score(keys[i]), currentAlpha))) { atomicMin(&s_next, i); break; }
                __syncthreads();
                const int pick = s_next;
                if (pick == INT_MAX) break;
                nsel++;
                cursor = pick + 1;
                for (int i = tid; i < nvalid; i += PRUNE_THREADS) {
                    if (i == pick) state[i] = 1;
                    else if (state[i] == 0) {
                        const float m = fmaxf(maxsim[i], C[i * CP + pick]);
                        maxsim[i] = m;
                        if (m > __fmul_rn(key_score(keys[i]), alpha_max)) state[i] = 2;
                    }
                }
                __syncthreads();
            }
            currentAlpha += 0.2f;
        }
        __syncthreads();
        if (tid == 0) {
            int w = 0;
            int32_t *row = P.out_rows ? P.out_rows + (size_t)(P.out_base + it) * P.out_stride : P.adj + (size_t)v * P.row_cap;
            const int width = P.out_rows ? P.out_stride : P.row_cap;
            for (int i = 0; i < nvalid; i++)
                if (state[i] == 1) row[w++] = ids[i];
            for (int i = w; i < width; i++) row[i] = -1;
            if (P.out_rows) P.out_deg[P.out_base + it] = w;
            else P.deg[v] = w;
            if (P.mode == 1 && P.mark) P.mark[v] = 0;
        }
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TILE) : "memory");
}

template <int TILE>
static size_t gram_smem_bytes(const DataDesc &d)
{
    size_t b = (size_t)(d.stride + 4) * 4 + (size_t)TILE * GRAM_PITCH * 4 + ((size_t)TILE * (TILE + 1) + 1) * 4 + (size_t)TILE * 8 + (size_t)TILE * 4 + (size_t)TILE * 4 + TILE;
    return (b + 15) & ~(size_t)15;
}

// ------------------------------------------------------------------------------------------------
// back-links (ConcurrentNeighborMap.backlink, ConcurrentNeighborMap.java:158-165): for every selected neighbour t of a new
// node u, append u to t's row. Deterministic: the (t, u) pairs of a batch are SORTED, pair i lands at deg_before[t] + (rank of i
// inside t's segment), so every replica of a sharded build applies a batch to bit-identical adjacency (atomics would order the
// appends by timing, and which back-link a full row drops would differ between replicas).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) apply_rows_kernel(int32_t *adj, int *deg, int row_cap, int degree, int first, int count, const int32_t *rows,
                                                         const int *rdeg, unsigned long long *pairs)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * row_cap) return;
    const int i = idx / row_cap, j = idx - i * row_cap;
    const int u = first + i;
    const int32_t t = j < degree ? rows[(size_t)i * degree + j] : -1;
    adj[(size_t)u * row_cap + j] = t;
    if (j == 0) deg[u] = rdeg[i];
    if (j < degree) {
        bool link = t >= 0;
        if (link && t >= first && t < first + count) {
            // t was inserted in this very batch and may have chosen u itself (both sit in each other's in-progress window): then u is
            // already in t's row and the back-link would duplicate it (NodeArray.insertSorted refuses duplicates)
            const int32_t *trow = rows + (size_t)(t - first) * degree;
            for (int x = 0; x < degree; x++)
                if (trow[x] == u) { link = false; break; }
        }
        pairs[(size_t)i * degree + j] = link ? (((unsigned long long)(unsigned)t << 32) | (unsigned)u) : ~0ull;  // ~0 sorts last
    }
}

__device__ __forceinline__ int lower_bound_u64(const unsigned long long *a, int n, unsigned long long key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// pairs sorted ascending. Thread i appends its u; the head of each segment publishes the new degree and flags the row when it
// passed overflow * M (ConcurrentNeighborMap.java:300: the row is then re-pruned).
__global__ void __launch_bounds__(256) backlink_sorted_kernel(int32_t *adj, int *deg, int row_cap, int hard_max, const unsigned long long *pairs, int npairs,
                                                              unsigned char *head_flag, int32_t *head_target, unsigned long long *dropped)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const unsigned long long p = pairs[i];
    head_flag[i] = 0;
    head_target[i] = -1;
    if (p == ~0ull) return;
    const int t = (int)(p >> 32), u = (int)(unsigned)p;
    const int seg0 = lower_bound_u64(pairs, npairs, (unsigned long long)(unsigned)t << 32);
    const int d0 = deg[t];  // no thread writes deg[t] before every pair of the segment has read it: see below
    const int pos = d0 + (i - seg0);
    if (pos < row_cap) adj[(size_t)t * row_cap + pos] = u;
    else atomicAdd(dropped, 1ull);
    if (i == seg0) {
        // segment head: remember the target and whether the row now passes overflow * M. The new degree itself is published by
        // publish_deg_kernel, a second launch, because the other pairs of the segment still read deg[t] in this one.
        const int seg1 = lower_bound_u64(pairs, npairs, (unsigned long long)((unsigned)t + 1u) << 32);
        const int nd = min(row_cap, d0 + (seg1 - seg0));
        head_target[i] = t;
        head_flag[i] = nd > hard_max ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) publish_deg_kernel(int *deg, int row_cap, const unsigned long long *pairs, int npairs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const unsigned long long p = pairs[i];
    if (p == ~0ull) return;
    const int t = (int)(p >> 32);
    if (i > 0 && (int)(pairs[i - 1] >> 32) == t) return;  // not a segment head
    const int seg1 = lower_bound_u64(pairs, npairs, (unsigned long long)((unsigned)t + 1u) << 32);
    deg[t] = min(row_cap, deg[t] + (seg1 - i));
}

__global__ void __launch_bounds__(256) flag_over_degree_kernel(const int *deg, int n, int degree, unsigned char *flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = deg[i] > degree ? 1 : 0;
}

__global__ void __launch_bounds__(256) iota_kernel(int32_t *a, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

// re-pruned rows coming back from the exchange: row i replaces adj[list[i]]
__global__ void __launch_bounds__(256) scatter_rows_kernel(int32_t *adj, int *deg, int row_cap, const int32_t *list, int count, const int32_t *rows, const int *rdeg)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * row_cap) return;
    const int i = idx / row_cap, j = idx - i * row_cap;
    const int v = list[i];
    adj[(size_t)v * row_cap + j] = rows[(size_t)i * row_cap + j];
    if (j == 0) deg[v] = rdeg[i];
}

__global__ void __launch_bounds__(256) compact_adj_kernel(const int32_t *adj, const int *deg, int n, int row_cap, int degree, int32_t *out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * degree) return;
    const int v = (int)(idx / degree), j = (int)(idx % degree);
    out[idx] = j < min(deg[v], degree) ? adj[(size_t)v * row_cap + j] : -1;
}

// rows [ids[i]] of a data set gathered into a dense one (upper levels of the hierarchy, repair pass): one warp per row
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ rows, int stride, const int32_t *__restrict__ ids, int count, float *__restrict__ out)
{
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= count) return;
    const float4 *src = reinterpret_cast<const float4 *>(rows + (size_t)ids[warp] * stride);
    float4 *dst = reinterpret_cast<float4 *>(out + (size_t)warp * stride);
    for (int i = lane; i < (stride >> 2); i += 32) dst[i] = ldg_stream(src + i);
}

cudaError_t launch_gather_rows(const DataDesc &f32, const int32_t *ids_dev, int count, float *out_dev, cudaStream_t s)
{
    if (count <= 0) return cudaSuccess;
    const long long blocks = ((long long)count * 32 + 255) / 256;
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, s>>>(f32.rows, f32.stride, ids_dev, count, out_dev);
    g_launches++;
    return cudaGetLastError();
}

static size_t prune_smem_bytes(const DataDesc &d)
{
    size_t b = (size_t)(d.stride + 4) * 4 + (size_t)d.stride * 4 + (size_t)PRUNE_MAXC * 8 + (size_t)PRUNE_MAXC * 4 + 128 * 4 + PRUNE_MAXC;
    return (b + 15) & ~(size_t)15;
}

template <int METRIC>
static cudaError_t launch_prune_t(const PruneParams &P, int grid, size_t smem, cudaStream_t s)
{
    // candidate sets that fit a tile go through the Gram-matrix kernel; anything larger keeps the incremental kernel
    const int ncmax = P.mode == 0 ? P.cand_stride + P.window : P.row_cap;
    cudaError_t e;
    // Gram products on the tensor cores (tcgen05, 3xTF32) unless JV_PRUNE_GRAM=ffma asks for the CUDA-core kernel
    static const bool use_tc = !(getenv("JV_PRUNE_GRAM") && getenv("JV_PRUNE_GRAM")[0] == 'f');
    if (ncmax <= 64 && use_tc) {
        const size_t gs = gram_tc_smem_bytes<64>(P.d);
        if ((e = cudaFuncSetAttribute(prune_gram_tc_kernel<METRIC, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs)) != cudaSuccess) return e;
        prune_gram_tc_kernel<METRIC, 64><<<grid, PRUNE_THREADS, gs, s>>>(P);
    } else if (ncmax <= 128 && use_tc) {
        const size_t gs = gram_tc_smem_bytes<128>(P.d);
        if ((e = cudaFuncSetAttribute(prune_gram_tc_kernel<METRIC, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs)) != cudaSuccess) return e;
        prune_gram_tc_kernel<METRIC, 128><<<grid, PRUNE_THREADS, gs, s>>>(P);
    } else if (ncmax <= 64) {
        const size_t gs = gram_smem_bytes<64>(P.d);
        if ((e = cudaFuncSetAttribute(prune_gram_kernel<METRIC, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs)) != cudaSuccess) return e;
        prune_gram_kernel<METRIC, 64><<<grid, PRUNE_THREADS, gs, s>>>(P);
    } else if (ncmax <= 128) {
        const size_t gs = gram_smem_bytes<128>(P.d);
        if ((e = cudaFuncSetAttribute(prune_gram_kernel<METRIC, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs)) != cudaSuccess) return e;
        prune_gram_kernel<METRIC, 128><<<grid, PRUNE_THREADS, gs, s>>>(P);
    } else {
        if ((e = cudaFuncSetAttribute(prune_kernel<METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        prune_kernel<METRIC><<<grid, PRUNE_THREADS, smem, s>>>(P);
    }
    g_launches++;
    return cudaGetLastError();
}

static cudaError_t launch_prune(const PruneParams &P, int grid, size_t smem, cudaStream_t s)
{
    if (P.metric == JV_METRIC_EUCLIDEAN) return launch_prune_t<JV_METRIC_EUCLIDEAN>(P, grid, smem, s);
    if (P.metric == JV_METRIC_DOT) return launch_prune_t<JV_METRIC_DOT>(P, grid, smem, s);
    return launch_prune_t<JV_METRIC_COSINE>(P, grid, smem, s);
}

#define JV_TRY(x)                      \
    do {                               \
        err = (x);                     \
        if (err != cudaSuccess) return err; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// GraphBuilder: the device-resident state of one flat (single level) build, advanced batch by batch. One rank of a sharded
// build calls insert_slice / reprune_slice for ITS part of a batch; the rows travel (NCCL all-gather in jvector_b200/parallel.py,
// nothing at all on one GPU) and every replica applies the whole batch with apply_new / apply_repruned, deterministically.
// ------------------------------------------------------------------------------------------------
struct GraphBuilder {
    DataDesc d;
    int metric = 0, sm_count = 0;
    BuildParams bp;
    int n = 0, degree = 0, beam = 0, hard_max = 0, row_cap = 0, max_batch = 0, window = 0, batch_div = 32;
    int inserted = 1;  // node 0 is the entry point with an empty list
    int32_t *adj = nullptr, *res_nodes = nullptr, *list = nullptr, *head_target = nullptr, *iota = nullptr;
    int *deg = nullptr, *mark = nullptr, *list_count = nullptr, *work_counter = nullptr;
    float *res_scores = nullptr;
    unsigned long long *dropped = nullptr, *pairs = nullptr, *pairs_sorted = nullptr;
    unsigned char *head_flag = nullptr, *node_flag = nullptr;
    SearchCounters *counters = nullptr;
    uint8_t *overflow = nullptr;
    void *scratch = nullptr, *cub_tmp = nullptr;
    size_t scratch_bytes = 0, cub_bytes = 0;
    int32_t *own_rows = nullptr;  // single-GPU path: the slice buffers are the builder's own
    int *own_deg = nullptr;
    int32_t *own_rp_rows = nullptr;
    int *own_rp_deg = nullptr;
    size_t own_rp_cap = 0;
    BuildStats st = {0, 0, 0, 0, 0};
    size_t psmem = 0;
    int prune_grid = 0;
};

static void builder_free(GraphBuilder *B)
{
    if (!B) return;
    cudaFree(B->adj); cudaFree(B->res_nodes); cudaFree(B->list); cudaFree(B->head_target); cudaFree(B->iota); cudaFree(B->deg); cudaFree(B->mark);
    cudaFree(B->list_count); cudaFree(B->work_counter); cudaFree(B->res_scores); cudaFree(B->dropped); cudaFree(B->pairs); cudaFree(B->pairs_sorted);
    cudaFree(B->head_flag); cudaFree(B->node_flag); cudaFree(B->counters); cudaFree(B->overflow); cudaFree(B->scratch); cudaFree(B->cub_tmp);
    cudaFree(B->own_rows); cudaFree(B->own_deg); cudaFree(B->own_rp_rows); cudaFree(B->own_rp_deg);
    delete B;
}

void builder_destroy(GraphBuilder *B) { builder_free(B); }

cudaError_t builder_create(const DataDesc &d, int metric, const BuildParams &bp, int sm_count, GraphBuilder **out, cudaStream_t s)
{
    const int n = (int)d.n;
    if (d.kind != KIND_F32 || bp.degree < 1 || bp.degree > 64 || bp.beam < 1 || bp.beam > PRUNE_MAXC) return cudaErrorInvalidValue;
    GraphBuilder *B = new GraphBuilder();
    B->d = d; B->metric = metric; B->bp = bp; B->sm_count = sm_count; B->n = n; B->degree = bp.degree; B->beam = bp.beam;
    B->hard_max = (int)(bp.overflow * bp.degree);  // ConcurrentNeighborMap.java:300
    B->row_cap = std::max(2 * bp.degree, B->hard_max + 1);
    if (B->row_cap > MAX_DEGREE) { delete B; return cudaErrorInvalidValue; }
    B->max_batch = bp.max_batch > 0 ? bp.max_batch : 16384;
    if (const char *e = getenv("JV_BUILD_BATCH_DIV")) B->batch_div = atoi(e) > 0 ? atoi(e) : 32;  // tuning knob (tools/build_quality.py)
    // in-progress window: as many concurrently inserting peers as still fit the 128-candidate Gram tile next to the beam
    B->window = bp.window >= 0 ? bp.window : std::max(0, std::min(32, 128 - bp.beam));
    B->psmem = prune_smem_bytes(d);
    B->prune_grid = sm_count * 5;  // ~10 KB of shared memory per CTA: register-limited residency
    cudaError_t err = cudaSuccess;
    const size_t mb = (size_t)B->max_batch;
#define ALLOC(ptr, bytes)                                                  \
    do {                                                                   \
        err = cudaMalloc((void **)&(ptr), (bytes));                        \
        if (err != cudaSuccess) { builder_free(B); return err; }           \
    } while (0)
    ALLOC(B->adj, (size_t)n * B->row_cap * 4);
    ALLOC(B->deg, (size_t)n * 4);
    ALLOC(B->mark, (size_t)n * 4);
    ALLOC(B->res_nodes, mb * B->beam * 4);
    ALLOC(B->res_scores, mb * B->beam * 4);
    ALLOC(B->list, (size_t)n * 4);
    ALLOC(B->iota, (size_t)n * 4);
    ALLOC(B->node_flag, (size_t)n);
    ALLOC(B->list_count, 16);
    ALLOC(B->work_counter, 16);
    ALLOC(B->dropped, 16);
    ALLOC(B->counters, sizeof(SearchCounters));
    ALLOC(B->overflow, mb);
    ALLOC(B->pairs, mb * B->degree * 8);
    ALLOC(B->pairs_sorted, mb * B->degree * 8);
    ALLOC(B->head_flag, mb * B->degree);
    ALLOC(B->head_target, mb * B->degree * 4);
    ALLOC(B->own_rows, mb * B->degree * 4);
    ALLOC(B->own_deg, mb * 4);
#undef ALLOC
    {
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, b1, B->pairs, B->pairs_sorted, (int)(mb * B->degree), 0, 64, s);
        cub::DeviceSelect::Flagged(nullptr, b2, B->iota, B->node_flag, B->list, B->list_count, n, s);
        B->cub_bytes = std::max(b1, b2) + 256;
        err = cudaMalloc(&B->cub_tmp, B->cub_bytes);
        if (err != cudaSuccess) { builder_free(B); return err; }
    }
    JV_TRY(cudaMemsetAsync(B->adj, 0xff, (size_t)n * B->row_cap * 4, s));
    JV_TRY(cudaMemsetAsync(B->deg, 0, (size_t)n * 4, s));
    JV_TRY(cudaMemsetAsync(B->mark, 0, (size_t)n * 4, s));
    JV_TRY(cudaMemsetAsync(B->dropped, 0, 16, s));
    JV_TRY(cudaMemsetAsync(B->list_count, 0, 16, s));
    JV_TRY(cudaMemsetAsync(B->counters, 0, sizeof(SearchCounters), s));
    iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(B->iota, n);
    g_launches++;
    *out = B;
    return cudaGetLastError();
}

// the next batch: 1/32 of what is already inserted, at most max_batch — every batch searches a graph that already holds 97 % of the
// nodes a sequential build would show it. Measured on 100k rows (profiles/r2_build_quality.md): recall@10 at rerankK 100 is 0.950
// with inserted / 2, 0.9585 with / 8, 0.9624 with / 32, against 0.9631 for the reference-order (sequential) builder.
bool builder_next_batch(GraphBuilder *B, int *first, int *count)
{
    if (B->inserted >= B->n) return false;
    int batch = B->inserted / B->batch_div;
    if (batch < 1) batch = 1;
    if (batch > B->max_batch) batch = B->max_batch;
    if (batch > B->n - B->inserted) batch = B->n - B->inserted;
    *first = B->inserted;
    *count = batch;
    return true;
}

static PruneParams prune_params(const GraphBuilder *B)
{
    PruneParams P;
    memset(&P, 0, sizeof(P));
    P.d = B->d; P.metric = B->metric; P.degree = B->degree; P.row_cap = B->row_cap; P.alpha = B->bp.alpha; P.adj = B->adj; P.deg = B->deg; P.mark = B->mark;
    return P;
}

// batch positions [lo, hi): beam search of the graph as it stands + robust prune of (beam U in-progress window) -> rows_out
// [batch][degree] at positions lo..hi-1, deg_out likewise. Reads adjacency, writes none of it.
cudaError_t builder_insert_slice(GraphBuilder *B, int first, int count, int lo, int hi, int32_t *rows_out, int *deg_out, cudaStream_t s)
{
    cudaError_t err = cudaSuccess;
    if (!rows_out) { rows_out = B->own_rows; deg_out = B->own_deg; }
    const int m = hi - lo;
    if (m <= 0) return cudaSuccess;
    GraphDesc g = {};
    g.n = B->n; g.degree = B->row_cap; g.levels = 1; g.entry_node = 0; g.entry_level = 0; g.adj0 = B->adj;
    SearchPlan plan;
    JV_TRY(plan_search(B->d, nullptr, g, B->beam, B->beam, m, 0, 0, B->sm_count, &plan));
    const size_t need = search_scratch_bytes(plan);
    if (need > B->scratch_bytes) {
        if (B->scratch) cudaFree(B->scratch);
        B->scratch = nullptr;
        B->scratch_bytes = 0;
        JV_TRY(cudaMalloc(&B->scratch, need));
        B->scratch_bytes = need;
    }
    SearchFilter lenient;
    memset(&lenient, 0, sizeof(lenient));
    lenient.lenient = 1;  // an insert search never fails the build: see SearchParams::lenient
    JV_TRY(launch_search(g, B->d, nullptr, B->metric, B->d.rows + (size_t)(first + lo) * B->d.stride, m, B->beam, B->beam, plan, B->scratch, B->work_counter,
                         B->res_nodes, B->res_scores, B->counters, B->overflow, nullptr, B->d.stride, &lenient, s));
    PruneParams P = prune_params(B);
    P.mode = 0; P.node_base = first + lo; P.count = m; P.cand = B->res_nodes; P.cand_stride = B->beam;
    P.window = B->window; P.batch_first = first; P.batch_count = count;
    P.out_rows = rows_out; P.out_deg = deg_out; P.out_stride = B->degree; P.out_base = lo;
    JV_TRY(launch_prune(P, m < B->prune_grid ? m : B->prune_grid, B->psmem, s));
    return cudaSuccess;
}

// the whole batch's rows (this rank's own slice + the gathered ones): write them, back-link them deterministically, and leave
// the SORTED list of rows that passed overflow * M in B->list (count in B->list_count[0])
cudaError_t builder_apply_new(GraphBuilder *B, int first, int count, const int32_t *rows, const int *rdeg, cudaStream_t s)
{
    cudaError_t err = cudaSuccess;
    if (!rows) { rows = B->own_rows; rdeg = B->own_deg; }
    const int npairs = count * B->degree;
    apply_rows_kernel<<<(count * B->row_cap + 255) / 256, 256, 0, s>>>(B->adj, B->deg, B->row_cap, B->degree, first, count, rows, rdeg, B->pairs);
    g_launches++;
    size_t tb = B->cub_bytes;
    JV_TRY(cub::DeviceRadixSort::SortKeys(B->cub_tmp, tb, B->pairs, B->pairs_sorted, npairs, 0, 64, s));
    g_launches++;
    backlink_sorted_kernel<<<(npairs + 255) / 256, 256, 0, s>>>(B->adj, B->deg, B->row_cap, B->hard_max, B->pairs_sorted, npairs, B->head_flag, B->head_target, B->dropped);
    g_launches++;
    publish_deg_kernel<<<(npairs + 255) / 256, 256, 0, s>>>(B->deg, B->row_cap, B->pairs_sorted, npairs);
    g_launches++;
    tb = B->cub_bytes;
    JV_TRY(cub::DeviceSelect::Flagged(B->cub_tmp, tb, B->head_target, B->head_flag, B->list, B->list_count, npairs, s));  // sorted by target: heads are
    g_launches++;
    B->inserted = first + count;
    B->st.batches++;
    return cudaGetLastError();
}

// rows list[lo .. hi) re-pruned (VamanaDiversityProvider.retainDiverse over the row itself). rows_out == nullptr: in place, with
// the count read on the device (one-GPU path, no host round trip); otherwise out-of-place into rows_out [.][row_cap] at lo..hi-1.
cudaError_t builder_reprune_slice(GraphBuilder *B, int lo, int hi, int32_t *rows_out, int *deg_out, cudaStream_t s)
{
    PruneParams P = prune_params(B);
    P.mode = 1; P.list = B->list;
    if (!rows_out) {
        P.count_ptr = B->list_count;
        return launch_prune(P, B->prune_grid, B->psmem, s);
    }
    if (hi <= lo) return cudaSuccess;
    P.count = hi - lo; P.list_base = lo; P.out_rows = rows_out; P.out_deg = deg_out; P.out_stride = B->row_cap; P.out_base = lo;
    return launch_prune(P, std::min(hi - lo, B->prune_grid), B->psmem, s);
}

cudaError_t builder_apply_repruned(GraphBuilder *B, int count, const int32_t *rows, const int *rdeg, cudaStream_t s)
{
    if (count <= 0) return cudaSuccess;
    scatter_rows_kernel<<<(count * B->row_cap + 255) / 256, 256, 0, s>>>(B->adj, B->deg, B->row_cap, B->list, count, rows, rdeg);
    g_launches++;
    return cudaGetLastError();
}

// cleanup() of the reference: enforceDegree on every row longer than M. Leaves the sorted list of those rows in B->list.
cudaError_t builder_collect_over_degree(GraphBuilder *B, cudaStream_t s)
{
    cudaError_t err = cudaSuccess;
    flag_over_degree_kernel<<<(B->n + 255) / 256, 256, 0, s>>>(B->deg, B->n, B->degree, B->node_flag);
    g_launches++;
    size_t tb = B->cub_bytes;
    JV_TRY(cub::DeviceSelect::Flagged(B->cub_tmp, tb, B->iota, B->node_flag, B->list, B->list_count, B->n, s));
    g_launches++;
    return cudaGetLastError();
}

cudaError_t builder_list_count(GraphBuilder *B, int *count_host, cudaStream_t s)
{
    cudaError_t err = cudaSuccess;
    JV_TRY(cudaMemcpyAsync(count_host, B->list_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    return cudaStreamSynchronize(s);
}

cudaError_t builder_finish(GraphBuilder *B, int32_t *adj_out_dev, BuildStats *stats, cudaStream_t s)
{
    cudaError_t err = cudaSuccess;
    const long long total = (long long)B->n * B->degree;
    compact_adj_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(B->adj, B->deg, B->n, B->row_cap, B->degree, adj_out_dev);
    g_launches++;
    JV_TRY(cudaGetLastError());
    SearchCounters hc;
    unsigned long long hd = 0;
    JV_TRY(cudaMemcpyAsync(&hc, B->counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
    JV_TRY(cudaMemcpyAsync(&hd, B->dropped, sizeof(hd), cudaMemcpyDeviceToHost, s));
    JV_TRY(cudaStreamSynchronize(s));
    B->st.searched = (long long)hc.visited;
    B->st.dropped_backlinks = (long long)hd;
    B->st.truncated_searches = (long long)hc.overflowed;
    if (stats) *stats = B->st;
    return cudaSuccess;
}

int builder_row_cap(const GraphBuilder *B) { return B->row_cap; }
int builder_degree(const GraphBuilder *B) { return B->degree; }
int builder_max_batch(const GraphBuilder *B) { return B->max_batch; }

// the whole build on one GPU: no exchange, no host synchronisation between batches
cudaError_t build_graph_flat(const DataDesc &d, int metric, const BuildParams &bp, int32_t *adj_out_dev, int sm_count,
                             BuildStats *stats, cudaStream_t s)
{
    GraphBuilder *B = nullptr;
    cudaError_t err = builder_create(d, metric, bp, sm_count, &B, s);
    if (err != cudaSuccess) return err;
    int first, count;
    while (err == cudaSuccess && builder_next_batch(B, &first, &count)) {
        err = builder_insert_slice(B, first, count, 0, count, nullptr, nullptr, s);
        if (err == cudaSuccess) err = builder_apply_new(B, first, count, nullptr, nullptr, s);
        if (err == cudaSuccess) err = builder_reprune_slice(B, 0, 0, nullptr, nullptr, s);
    }
    if (err == cudaSuccess) err = builder_collect_over_degree(B, s);
    if (err == cudaSuccess) err = builder_reprune_slice(B, 0, 0, nullptr, nullptr, s);
    if (err == cudaSuccess) err = builder_finish(B, adj_out_dev, stats, s);
    else if (stats) *stats = B->st;
    builder_free(B);
    return err;
}

}  // namespace jv
