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
// Batches grow geometrically (x1.5) up to max_batch so early nodes see a connected graph. Neighbour lists in the
// reference are concurrency-order dependent, so parity is on scores and on the recall of the resulting graph
// (SURVEY §8d C5), not on identical adjacency.
#include <limits.h>

#include <vector>

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
};

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
        const int v = P.mode == 0 ? P.node_base + it : P.list[it];
        const int32_t *cand;
        int nc;
        if (P.mode == 0) {
            cand = P.cand + (size_t)it * P.cand_stride;
            nc = P.cand_stride;
        } else {
            cand = P.adj + (size_t)v * P.row_cap;
            nc = min(P.deg[v], P.row_cap);
        }
        nc = min(nc, PRUNE_MAXC);
        prepare_blob(P.d, P.metric, P.d.rows + (size_t)v * P.d.stride, blob, red);
        // exact scores of the candidates against v, as sortable keys (two rows per warp at a time)
        for (int i = warp; i < PRUNE_MAXC; i += 2 * NW) {
            const int i2 = i + NW;
            const int32_t ca = i < nc ? cand[i] : -1, cb = i2 < nc ? cand[i2] : -1;
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
            int32_t *row = P.adj + (size_t)v * P.row_cap;
            for (int i = 0; i < nvalid; i++)
                if (state[i] == 1) row[w++] = key_node(keys[i]);
            for (int i = w; i < P.row_cap; i++) row[i] = -1;
            P.deg[v] = w;
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
        const int v = P.mode == 0 ? P.node_base + it : P.list[it];
        const int32_t *cand;
        int nc;
        if (P.mode == 0) {
            cand = P.cand + (size_t)it * P.cand_stride;
            nc = P.cand_stride;
        } else {
            cand = P.adj + (size_t)v * P.row_cap;
            nc = min(P.deg[v], P.row_cap);
        }
        nc = min(nc, TILE);
        prepare_blob(P.d, P.metric, P.d.rows + (size_t)v * P.d.stride, blob, red);
        for (int i = warp; i < TILE; i += 2 * NW) {
            const int i2 = i + NW;
            const int32_t ca = i < nc ? cand[i] : -1, cb = (i2 < TILE && i2 < nc) ? cand[i2] : -1;
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
            int32_t *row = P.adj + (size_t)v * P.row_cap;
            for (int i = 0; i < nvalid; i++)
                if (state[i] == 1) row[w++] = ids[i];
            for (int i = w; i < P.row_cap; i++) row[i] = -1;
            P.deg[v] = w;
            if (P.mode == 1 && P.mark) P.mark[v] = 0;
        }
        __syncthreads();
    }
}

template <int TILE>
static size_t gram_smem_bytes(const DataDesc &d)
{
    size_t b = (size_t)(d.stride + 4) * 4 + (size_t)TILE * GRAM_PITCH * 4 + ((size_t)TILE * (TILE + 1) + 1) * 4 + (size_t)TILE * 8 + (size_t)TILE * 4 + (size_t)TILE * 4 + TILE;
    return (b + 15) & ~(size_t)15;
}

// back-links: for every selected neighbour v of a new node u append u to v's row (ConcurrentNeighborMap.backlink)
__global__ void __launch_bounds__(256) backlink_kernel(int32_t *adj, int *deg, int row_cap, int degree, int node_base, int count, int hard_max,
                                                       int *mark, int32_t *prune_list, int *prune_count, unsigned long long *dropped)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * degree) return;
    const int u = node_base + idx / degree, j = idx % degree;
    const int32_t v = adj[(size_t)u * row_cap + j];
    if (v < 0) return;
    const int pos = atomicAdd(&deg[v], 1);
    if (pos < row_cap) adj[(size_t)v * row_cap + pos] = u;
    else atomicAdd(dropped, 1ull);
    if (pos + 1 > hard_max && atomicExch(&mark[v], 1) == 0) prune_list[atomicAdd(prune_count, 1)] = v;
}

__global__ void __launch_bounds__(256) clamp_deg_kernel(int *deg, int n, int row_cap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && deg[i] > row_cap) deg[i] = row_cap;
}

__global__ void __launch_bounds__(256) collect_over_degree_kernel(const int *deg, int n, int degree, int32_t *list, int *count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && deg[i] > degree) list[atomicAdd(count, 1)] = i;
}

__global__ void __launch_bounds__(256) compact_adj_kernel(const int32_t *adj, const int *deg, int n, int row_cap, int degree, int32_t *out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * degree) return;
    const int v = (int)(idx / degree), j = (int)(idx % degree);
    out[idx] = j < min(deg[v], degree) ? adj[(size_t)v * row_cap + j] : -1;
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
    const int ncmax = P.mode == 0 ? P.cand_stride : P.row_cap;
    cudaError_t e;
    if (ncmax <= 64) {
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
        if (err != cudaSuccess) goto done; \
    } while (0)

cudaError_t build_graph_flat(const DataDesc &d, int metric, const BuildParams &bp, int32_t *adj_out_dev, int sm_count,
                             BuildStats *stats, cudaStream_t s)
{
    const int n = (int)d.n;
    const int degree = bp.degree, beam = bp.beam;
    if (d.kind != KIND_F32 || degree < 1 || degree > 64 || beam < 1 || beam > PRUNE_MAXC) return cudaErrorInvalidValue;
    const int hard_max = (int)(bp.overflow * degree);  // ConcurrentNeighborMap.java:300
    int row_cap = 2 * degree;
    if (row_cap < hard_max + 1) row_cap = hard_max + 1;
    if (row_cap > MAX_DEGREE) return cudaErrorInvalidValue;
    const int max_batch = bp.max_batch > 0 ? bp.max_batch : 16384;

    cudaError_t err = cudaSuccess;
    int32_t *adj = nullptr, *res_nodes = nullptr, *prune_list = nullptr;
    int *deg = nullptr, *mark = nullptr, *prune_count = nullptr, *work_counter = nullptr;
    float *res_scores = nullptr;
    unsigned long long *dropped = nullptr;
    SearchCounters *counters = nullptr;
    uint8_t *overflow = nullptr;
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    BuildStats st = {0, 0, 0, 0};
    const size_t psmem = prune_smem_bytes(d);
    const int prune_grid = sm_count * 5;  // ~10 KB of shared memory per CTA: register-limited residency

    JV_TRY(cudaMalloc(&adj, (size_t)n * row_cap * sizeof(int32_t)));
    JV_TRY(cudaMemsetAsync(adj, 0xff, (size_t)n * row_cap * sizeof(int32_t), s));
    JV_TRY(cudaMalloc(&deg, (size_t)n * sizeof(int)));
    JV_TRY(cudaMemsetAsync(deg, 0, (size_t)n * sizeof(int), s));
    JV_TRY(cudaMalloc(&mark, (size_t)n * sizeof(int)));
    JV_TRY(cudaMemsetAsync(mark, 0, (size_t)n * sizeof(int), s));
    JV_TRY(cudaMalloc(&res_nodes, (size_t)max_batch * beam * sizeof(int32_t)));
    JV_TRY(cudaMalloc(&res_scores, (size_t)max_batch * beam * sizeof(float)));
    JV_TRY(cudaMalloc(&prune_list, (size_t)n * sizeof(int32_t)));
    JV_TRY(cudaMalloc(&prune_count, sizeof(int)));
    JV_TRY(cudaMalloc(&work_counter, sizeof(int)));
    JV_TRY(cudaMalloc(&dropped, sizeof(unsigned long long)));
    JV_TRY(cudaMemsetAsync(dropped, 0, sizeof(unsigned long long), s));
    JV_TRY(cudaMalloc(&counters, sizeof(SearchCounters)));
    JV_TRY(cudaMemsetAsync(counters, 0, sizeof(SearchCounters), s));
    JV_TRY(cudaMalloc(&overflow, (size_t)max_batch));

    {
        GraphDesc g = {};
        g.n = n; g.degree = row_cap; g.levels = 1; g.entry_node = 0; g.entry_level = 0;
        g.adj0 = adj; g.upper_row = nullptr; g.upper_adj = nullptr; g.upper_off = nullptr;
        int inserted = 1;  // node 0 is the entry point with an empty list
        while (inserted < n) {
            int batch = inserted / 2;
            if (batch < 1) batch = 1;
            if (batch > max_batch) batch = max_batch;
            if (batch > n - inserted) batch = n - inserted;
            // (1) beam search of the current graph for every node of the batch
            SearchPlan plan;
            JV_TRY(plan_search(d, nullptr, g, beam, beam, batch, 0, 0, sm_count, &plan));
            const size_t need = search_scratch_bytes(plan);
            if (need > scratch_bytes) {
                if (scratch) cudaFree(scratch);
                scratch = nullptr;
                JV_TRY(cudaMalloc(&scratch, need));
                scratch_bytes = need;
            }
            JV_TRY(launch_search(g, d, nullptr, metric, d.rows + (size_t)inserted * d.stride, batch, beam, beam, plan, scratch, work_counter,
                                 res_nodes, res_scores, counters, overflow, nullptr, d.stride, nullptr, s));
            // (2) robust prune of each beam -> the new node's list
            PruneParams P;
            P.d = d; P.metric = metric; P.degree = degree; P.row_cap = row_cap; P.alpha = bp.alpha; P.adj = adj; P.deg = deg;
            P.mode = 0; P.node_base = inserted; P.list = nullptr; P.count_ptr = nullptr; P.count = batch; P.cand = res_nodes;
            P.cand_stride = beam; P.mark = mark;
            JV_TRY(launch_prune(P, batch < prune_grid ? batch : prune_grid, psmem, s));
            // (3) back-links, then re-prune every neighbour that passed overflow * M
            JV_TRY(cudaMemsetAsync(prune_count, 0, sizeof(int), s));
            {
                const int threads = batch * degree;
                backlink_kernel<<<(threads + 255) / 256, 256, 0, s>>>(adj, deg, row_cap, degree, inserted, batch, hard_max, mark, prune_list, prune_count, dropped);
                g_launches++;
                clamp_deg_kernel<<<(n + 255) / 256, 256, 0, s>>>(deg, n, row_cap);
                g_launches++;
            }
            P.mode = 1; P.list = prune_list; P.count_ptr = prune_count; P.count = 0;
            JV_TRY(launch_prune(P, prune_grid, psmem, s));
            inserted += batch;
            st.batches++;
        }
        // cleanup(): enforceDegree on every list longer than M, then emit [n][degree]
        JV_TRY(cudaMemsetAsync(prune_count, 0, sizeof(int), s));
        collect_over_degree_kernel<<<(n + 255) / 256, 256, 0, s>>>(deg, n, degree, prune_list, prune_count);
        g_launches++;
        PruneParams P;
        P.d = d; P.metric = metric; P.degree = degree; P.row_cap = row_cap; P.alpha = bp.alpha; P.adj = adj; P.deg = deg;
        P.mode = 1; P.node_base = 0; P.list = prune_list; P.count_ptr = prune_count; P.count = 0; P.cand = nullptr; P.cand_stride = 0;
        P.mark = mark;
        JV_TRY(launch_prune(P, prune_grid, psmem, s));
        const long long total = (long long)n * degree;
        compact_adj_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(adj, deg, n, row_cap, degree, adj_out_dev);
        g_launches++;
        JV_TRY(cudaGetLastError());
        SearchCounters hc;
        unsigned long long hd = 0;
        JV_TRY(cudaMemcpyAsync(&hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
        JV_TRY(cudaMemcpyAsync(&hd, dropped, sizeof(hd), cudaMemcpyDeviceToHost, s));
        JV_TRY(cudaStreamSynchronize(s));
        st.searched = (long long)hc.visited;
        st.dropped_backlinks = (long long)hd;
        if (hc.overflowed) err = cudaErrorLaunchOutOfResources;
    }
done:
    if (stats) *stats = st;
    cudaFree(adj); cudaFree(deg); cudaFree(mark); cudaFree(res_nodes); cudaFree(res_scores); cudaFree(prune_list);
    cudaFree(prune_count); cudaFree(work_counter); cudaFree(dropped); cudaFree(counters); cudaFree(overflow); cudaFree(scratch);
    return err;
}

}  // namespace jv
