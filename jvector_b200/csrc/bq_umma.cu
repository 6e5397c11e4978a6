// bq_umma.cu — the FILTER pass of the BQ Hamming top-k (see bq_imma.cu for the whole pipeline) on the 5th-generation tensor
// cores: tcgen05.mma kind::i8 (u8 x u8 -> s32, exact), accumulators in TMEM, operands in shared memory in the canonical K-major
// SWIZZLE_128B layout, warp-specialised roles connected by mbarriers.
//
//   D[128 rows][256 queries] += A[128][K] * B[256][K]^T,  A bytes in {0, 2^s}, B bytes in {0, 2^(7-s)} expanded from bit words
//   (D = 128 * popcount(row & query), exact in s32), K = dim rounded to 128
//
// Why a second kernel next to the IMMA one: the legacy IMMA path tops out at 917 TOP/s and re-expands both operands in every
// warp tile; here the row tile is expanded ONCE per CTA tile into shared memory (producer warps), the query tile arrives as a
// pre-expanded, pre-swizzled 32 KB image by one bulk copy, and one elected thread issues 128 x 256 x 32 MMAs.
//
// Roles (640 threads): warp 0 = B producer (cp.async.bulk of the query image chunk), warp 1 = MMA issuer + TMEM owner,
// warps 4-7 and 16-19 = two epilogue groups, one per accumulator buffer, so each has two tiles' worth of MMA time to drain its
// tile (tcgen05.ld: warp w may touch TMEM lanes 32 (w % 4) ..), warps 8-15 = A producers (two threads per row:
// 64 bits -> 64 bytes each, swizzled 128-bit stores, fence.proxy.async, arrive). Pipelines: full[s] / empty[s] over 4 operand
// stages of 48 KB, tmem_full[a] / tmem_empty[a] over two 256-column accumulators, so the epilogue of one tile overlaps the MMAs
// of the next. The epilogue is branch-free per column: thresholds of the query tile sit in shared memory, a column costs
// IADD3 + ISETP + SEL into a survivor bit mask; the rare survivors (about 0.3 %) are re-read from TMEM in a warp-uniform loop and go
// to a per-warp shared-memory queue (ballot slots, no atomics) that the warp flushes after handing the accumulator back. Row
// popcounts come from the producers. Round-2 profiles (profiles/r2_ncu_bq_umma.md): the first version's per-element LDG + branch
// epilogue kept the MMA thread waiting half its time (6.8 ms); this one runs the filter pass in 2 ms.
// Every spin-wait is bounded and traps: a wrong barrier count aborts the launch instead of hanging the device.
#include <limits.h>

#include "kernels.h"

namespace jv {

namespace {

constexpr int UM = 128;           // rows per tile (MMA M)
constexpr int UN = 256;           // queries per tile (MMA N)
constexpr int UKC = 128;          // K bytes per stage (one 128-byte swizzle atom wide)
constexpr int USTAGES = 4;
constexpr int UTHREADS = 640;
constexpr int A_STAGE_BYTES = UM * UKC;  // 16 KB
constexpr int B_STAGE_BYTES = UN * UKC;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int N_A_PRODUCERS = 256;  // two threads per row
constexpr int NPAR = 8;               // ring of row-popcount buffers: the producers run up to USTAGES stages (= tiles when K is one
                                      // chunk) ahead of the MMA, which runs up to 2 tiles ahead of the epilogue
constexpr int WQ_CAP = 256;           // survivor queue of one epilogue warp and tile (expected ~25 entries); overflow appends directly

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, unsigned phase)
{
    for (unsigned spins = 0; !mbar_try_wait(bar, phase); ++spins)
        if (spins > (1u << 28)) __trap();  // a pipeline bug must abort the launch, not hang the device
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor; canonical layout
// ((8,n),2):((8,SBO),1) in 16-byte units): rows 128 B apart inside an 8-row atom, atoms SBO = 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused by swizzled K-major layouts), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                       // layout type SWIZZLE_128B, bits [61,64)
    return d;
}

// instruction descriptor (InstrDescriptor): c = S32 (2) at [4,6), a = b = U8 (0), K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t UMMA_IDESC = (2u << 4) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// the query tile as shared-memory images: [qtile][kchunk][UN rows x 128 B], swizzled exactly as the MMA reads it
__global__ void __launch_bounds__(256) bq_query_image_kernel(const uint32_t *__restrict__ qbits, int nq_pad, int W32, int kchunks, uint8_t *__restrict__ images)
{
    // one thread per (query, 16-byte chunk): 16 bits -> 16 bytes
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)nq_pad * kchunks * 8;
    if (idx >= total) return;
    const int j = (int)(idx & 7);
    const long long t = idx >> 3;
    const int kc = (int)(t % kchunks), q = (int)(t / kchunks);
    const uint32_t w = qbits[(size_t)q * W32 + kc * 4 + (j >> 1)];
    // K order inside a 32-bit word: output word s holds bits s, s + 8, s + 16, s + 24 (one per byte). The row side stores such a
    // bit as the byte value 2^s (a single AND with a shifted mask), the query side as 2^(7 - s): every matching bit pair
    // contributes exactly 128 to the s32 accumulator, so D = 128 * popcount(row & query).
    const int sb = (j & 1) * 4;
    uint4 v;
    v.x = ((w >> (sb + 0)) & 0x01010101u) << (7 - (sb + 0));
    v.y = ((w >> (sb + 1)) & 0x01010101u) << (7 - (sb + 1));
    v.z = ((w >> (sb + 2)) & 0x01010101u) << (7 - (sb + 2));
    v.w = ((w >> (sb + 3)) & 0x01010101u) << (7 - (sb + 3));
    const int qt = q / UN, r = q % UN;
    uint8_t *img = images + ((size_t)qt * kchunks + kc) * B_STAGE_BYTES;
    *reinterpret_cast<uint4 *>(img + (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)) = v;
}

struct UmmaParams {
    const uint32_t *rows;  // [n][W32]
    long long n;
    int W32, dim, kchunks;
    const uint8_t *images;  // [qtiles][kchunks][32 KB]
    int qtiles, nq;
    const int *t2, *pb;
    long long *buf;
    int *cnt;
    int cap;
    long long id_base;
};

__global__ void __launch_bounds__(UTHREADS, 1) bq_umma_filter_kernel(UmmaParams P)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *stage0 = smem;  // USTAGES x (A 16 KB | B 32 KB), 1024-byte aligned
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + USTAGES * STAGE_BYTES);
    uint64_t *empty = full + USTAGES;
    uint64_t *tmem_full = empty + USTAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    int *s_t2 = reinterpret_cast<int *>(smem + USTAGES * STAGE_BYTES + 256);  // [2 groups][UN] thresholds of the group's query tile
    int *s_parh = s_t2 + 2 * UN;                                                  // [NPAR tiles][2 halves][UM] row popcounts from the producers
    int2 *s_hits = reinterpret_cast<int2 *>(s_parh + NPAR * 2 * UM);                 // [8 warps][WQ_CAP] (row << 16 | column, par - 2 dot)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row_tiles = (P.n + UM - 1) / UM;
    const long long tiles = row_tiles * P.qtiles;

    if (tid == 0) {
        for (int s = 0; s < USTAGES; s++) {
            mbar_init(&full[s], N_A_PRODUCERS + 1);  // 256 row producers + the bulk copy's arrive.expect_tx
            mbar_init(&empty[s], 1);                 // one tcgen05.commit
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== B producer: one bulk copy of the query image chunk per stage =====
        if (lane == 0) {
            int stage = 0;
            unsigned phase = 0;
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int qt = (int)(tile / row_tiles);
                for (int kc = 0; kc < P.kchunks; kc++) {
                    mbar_wait_bounded(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], B_STAGE_BYTES);
                    bulk_g2s(stage0 + stage * STAGE_BYTES + A_STAGE_BYTES, P.images + ((size_t)qt * P.kchunks + kc) * B_STAGE_BYTES, B_STAGE_BYTES, &full[stage]);
                    if (++stage == USTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0, acc = 0;
            unsigned phase = 0, acc_phase = 0;
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait_bounded(&tmem_empty[acc], acc_phase ^ 1);  // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * UN);
                for (int kc = 0; kc < P.kchunks; kc++) {
                    mbar_wait_bounded(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stage0 + stage * STAGE_BYTES), b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < UKC / 32; k++)
                        umma_i8(d, umma_desc(a_addr + 32 * k), umma_desc(b_addr + 32 * k), UMMA_IDESC, (kc | k) != 0 ? 1u : 0u);
                    umma_commit(&empty[stage]);  // frees the stage when these MMAs have read it
                    if (++stage == USTAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);    // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if ((warp >= 4 && warp < 8) || warp >= 16) {
        // ===== epilogue: thread = one row of the tile = one TMEM lane; group g drains accumulator buffer g (every other tile) =====
        const int grp = warp >= 16 ? 1 : 0, wq4 = warp & 3;
        const int ltid = wq4 * 32 + lane;
        const int acc = grp;
        int cur_qt = -1;
        unsigned acc_phase = 0;
        int *my_t2 = s_t2 + grp * UN;
        int2 *wq = s_hits + (grp * 4 + wq4) * WQ_CAP;  // this warp's private survivor queue: no atomics, no cross-warp barrier
        const uint32_t lt_mask = (1u << lane) - 1u;
        auto append = [&](int q, long long rr, int x) {  // x = par - 2 dot
            const int hd = x + __ldg(P.pb + q);
            const long long key = topk_key(bq_score_from_hd(hd, P.dim), (int32_t)(rr + P.id_base));
            const int pos = atomicAdd(&P.cnt[q], 1);
            if (pos < P.cap) P.buf[(size_t)q * P.cap + pos] = key;
        };
        int it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
            if ((it & 1) != grp) continue;
            const int qt = (int)(tile / row_tiles), tbuf = it & (NPAR - 1);
            const long long row0 = (tile % row_tiles) * UM, rr = row0 + ltid;
            if (qt != cur_qt) {  // the thresholds of this query tile (padding queries never pass)
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                else asm volatile("bar.sync 2, 128;" ::: "memory");
                for (int c = ltid; c < UN; c += 128) {
                    const int q = qt * UN + c;
                    my_t2[c] = q < P.nq ? 128 * __ldg(P.t2 + q) : INT_MIN;  // the accumulators hold 128 * dot
                }
                cur_qt = qt;
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                else asm volatile("bar.sync 2, 128;" ::: "memory");
            }
            mbar_wait_bounded(&tmem_full[acc], acc_phase);
            tc_fence_after();
            // the row's popcount, summed by the two producer threads of the row while they expanded it (ordered before the
            // accumulator's completion through full[] -> MMA -> tmem_full[])
            const int par = s_parh[(tbuf * 2 + 0) * UM + ltid] + s_parh[(tbuf * 2 + 1) * UM + ltid];
            const int par128 = rr < P.n ? par << 7 : INT_MAX;  // rows past the end never pass
            const uint32_t taddr = tmem_base + ((uint32_t)(wq4 * 32) << 16) + (uint32_t)(acc * UN);
            int wcount = 0;  // warp-uniform
#pragma unroll 1
            for (int g = 0; g < UN / 32; g++) {
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                      "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                      "=r"(v[31])
                    : "r"(taddr + (uint32_t)(g * 32)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int4 *tp = reinterpret_cast<const int4 *>(my_t2 + g * 32);
                // pass 1, branch-free: one bit per column that survives (independent LEA / ISETP / SEL chains; the round-2 profile
                // showed a per-column branch costing ~90 cycles of dependent issue). hd - pb = par - 2 dot <= t2  <=>
                // 2 (128 dot) + 128 t2 >= 128 par
                uint32_t mask = 0;
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++) {
                    const int4 t = tp[c4];  // one broadcast LDS.128 per four columns
                    const int tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int c = c4 * 4 + e;
                        mask |= (2 * (int)v[c] + tt[e] >= par128) ? (1u << c) : 0u;
                    }
                }
                // pass 2: only the columns in which some lane of the warp has a survivor (about 3 of 32). The loop is warp-uniform;
                // the column's accumulator is read back from TMEM with a one-column load (the column index is warp-uniform, a
                // register array could only be indexed statically); slots in the warp's queue come from a ballot.
                uint32_t um = __reduce_or_sync(0xffffffffu, mask);
                while (um) {
                    const int c = __ffs(um) - 1;
                    um &= um - 1;
                    uint32_t dcol;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(dcol) : "r"(taddr + (uint32_t)(g * 32 + c)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const bool hit = (mask >> c) & 1u;
                    const uint32_t b = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int x = par - ((int)dcol >> 6);  // par - 2 dot
                        const int slot = wcount + __popc(b & lt_mask);
                        if (slot < WQ_CAP) wq[slot] = make_int2((ltid << 16) | (g * 32 + c), x);
                        else append(qt * UN + g * 32 + c, rr, x);
                    }
                    wcount += __popc(b);
                }
            }
            // the accumulator is drained: hand it back before the (slower) global flush of the queue
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            acc_phase ^= 1;
            __syncwarp();
            const int nh = min(wcount, WQ_CAP);
            for (int i = lane; i < nh; i += 32) {
                const int2 h = wq[i];
                append(qt * UN + (h.x & 0xffff), row0 + (h.x >> 16), h.y);
            }
            __syncwarp();
        }
    } else if (warp >= 8 && warp < 16) {
        // ===== A producers: two threads per row; 64 bits -> 64 bytes per stage and thread, swizzled 16-byte chunks =====
        const int t = (warp - 8) * 32 + lane, r = t & 127, half = t >> 7;
        int stage = 0, tbuf = 0;
        unsigned phase = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const long long rr = (tile % row_tiles) * UM + r;
            const bool live = rr < P.n;
            const uint2 *rp = reinterpret_cast<const uint2 *>(P.rows + (size_t)(live ? rr : 0) * P.W32) + half;
            uint2 w = make_uint2(0u, 0u);
            if (live) w = __ldg(rp);
            int pc = 0;
            for (int kc = 0; kc < P.kchunks; kc++) {
                uint2 wn = make_uint2(0u, 0u);
                if (live && kc + 1 < P.kchunks) wn = __ldg(rp + 2 * (kc + 1));  // the next chunk's words are in flight during this one
                mbar_wait_bounded(&empty[stage], phase ^ 1);
                unsigned char *dst = stage0 + stage * STAGE_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
                const uint32_t ws[2] = {w.x, w.y};
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    const int j = half * 4 + jj, sb = (jj & 1) * 4;
                    const uint32_t word = ws[jj >> 1];
                    uint4 v;  // bit (8 b + s) of the word -> byte b of output word s, value 2^s: one AND per output word
                    v.x = word & (0x01010101u << (sb + 0));
                    v.y = word & (0x01010101u << (sb + 1));
                    v.z = word & (0x01010101u << (sb + 2));
                    v.w = word & (0x01010101u << (sb + 3));
                    *reinterpret_cast<uint4 *>(dst + ((j ^ (r & 7)) << 4)) = v;
                }
                pc += __popc(w.x) + __popc(w.y);
                if (kc + 1 == P.kchunks) s_parh[(tbuf * 2 + half) * UM + r] = pc;  // this half row's popcount, for the epilogue
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                mbar_arrive(&full[stage]);
                if (++stage == USTAGES) { stage = 0; phase ^= 1; }
                w = wn;
            }
            tbuf = (tbuf + 1) & (NPAR - 1);
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace

size_t bq_umma_image_bytes(int nq, int W)
{
    const int qtiles = (nq + UN - 1) / UN, kchunks = (2 * W) / 4;
    return (size_t)qtiles * kchunks * B_STAGE_BYTES + 1024;
}

bool bq_umma_supported(const DataDesc &d)
{
    return d.kind == KIND_BQ && (d.W % 2) == 0 && d.W <= 32;
}

// the filter pass over queries [0, nq) (identity list): pairs with hd <= thr[q] append their keys to buf / cnt
cudaError_t launch_bq_umma_filter(const DataDesc &d, const uint32_t *qbits_dev, int nq, int nq_pad, const int *t2_dev, const int *pb_dev, long long *buf_dev,
                                  int *cnt_dev, int cap, long long id_base, uint8_t *images_dev, int sm_count, cudaStream_t s)
{
    const int W32 = 2 * d.W, kchunks = W32 / 4, qtiles = (nq + UN - 1) / UN;
    cudaError_t e;
    // padding queries of the last tile need defined (zero) bits: qbits rows [nq, nq_pad) are zero; rows past nq_pad are not read
    const int q_img = qtiles * UN;
    if (q_img > nq_pad) return cudaErrorInvalidValue;  // callers pad nq_pad to a multiple of 256
    {
        const long long total = (long long)q_img * kchunks * 8;
        bq_query_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(qbits_dev, q_img, W32, kchunks, images_dev);
        g_launches++;
    }
    UmmaParams P;
    P.rows = reinterpret_cast<const uint32_t *>(d.words);
    P.n = d.n; P.W32 = W32; P.dim = d.dim; P.kchunks = kchunks; P.images = images_dev; P.qtiles = qtiles; P.nq = nq; P.t2 = t2_dev; P.pb = pb_dev;
    P.buf = buf_dev; P.cnt = cnt_dev; P.cap = cap; P.id_base = id_base;
    const size_t smem = (size_t)USTAGES * STAGE_BYTES + 256 + 2 * UN * sizeof(int) + NPAR * 2 * UM * sizeof(int) + 8 * WQ_CAP * sizeof(int2);
    if ((e = cudaFuncSetAttribute(bq_umma_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    const long long tiles = ((d.n + UM - 1) / UM) * qtiles;
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    bq_umma_filter_kernel<<<grid, UTHREADS, smem, s>>>(P);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace jv
