// legacy_avx512.cpp — AVX-512 bodies of the legacy libjvector.so symbols (the host-synchronous n = 1 group of include/jvector_b200.h).
//
// Installed as `libjvector.so` this library serves every NON-batched similarity / PQ / NVQ call of jvector-native, so these
// symbols must not be slower than the reference's Highway kernels (native-c:src/jvector_simd_kernels.cpp:208-287 similarities,
// :543-724 partial sums / assemble-and-sum, :729-879 PQ pair table / decoded cosine, :1047-1641 NVQ). Each function here follows
// the scalar definition in legacy_host.cpp operation for operation (explicit FMAs, same bit tricks), 16 elements at a time; sums
// are folded from 16-lane accumulators, which stays inside the reference's own provider-to-provider tolerance. legacy_host.cpp
// dispatches here when the CPU has AVX-512 F/BW/VL/DQ (checked once), otherwise it runs its scalar bodies.
// tools/legacy_bench.c times every symbol beside oracle/_ref/libjvector.so.
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "legacy_avx512.h"

#define JV_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,fma")))

namespace jvl {

namespace {

JV_AVX512 inline __mmask16 tail_mask(size_t n) { return (__mmask16)((1u << n) - 1u); }

JV_AVX512 inline float hsum(__m512 v) { return _mm512_reduce_add_ps(v); }

// logistic_nqt of legacy_host.cpp, 16 lanes
JV_AVX512 inline __m512 logistic16(__m512 v, __m512 alpha, __m512 c0)
{
    const __m512 t = _mm512_fmadd_ps(v, alpha, c0);  // fma(v, alpha, -alpha * x0)
    const __m512i ti = _mm512_castps_si512(t);
    const __mmask16 neg = _mm512_movepi32_mask(ti);  // sign bit: f2i(t) < 0
    const __m512i p = _mm512_mask_blend_epi32(neg, _mm512_cvttps_epi32(_mm512_add_ps(t, _mm512_set1_ps(1.0f))), _mm512_cvttps_epi32(t));
    const __m512 e = _mm512_cvtepi32_ps(p);
    const __m512i m = _mm512_castps_si512(_mm512_fmadd_ps(_mm512_sub_ps(t, e), _mm512_set1_ps(0.5f), _mm512_set1_ps(1.0f)));
    const __m512 r = _mm512_castsi512_ps(_mm512_add_epi32(m, _mm512_slli_epi32(p, 23)));
    return _mm512_div_ps(r, _mm512_add_ps(r, _mm512_set1_ps(1.0f)));
}

// logit_nqt of legacy_host.cpp, 16 lanes
JV_AVX512 inline __m512 logit16(__m512 v, __m512 inv_alpha, __m512 x0)
{
    const __m512 z = _mm512_div_ps(v, _mm512_sub_ps(_mm512_set1_ps(1.0f), v));
    const __m512i t = _mm512_castps_si512(z);
    const __m512i p = _mm512_sub_epi32(_mm512_srli_epi32(_mm512_and_si512(t, _mm512_set1_epi32(0x7f800000)), 23), _mm512_set1_epi32(128));
    const __m512 m = _mm512_castsi512_ps(_mm512_add_epi32(_mm512_and_si512(t, _mm512_set1_epi32(0x007fffff)), _mm512_set1_epi32(0x3f800000)));
    return _mm512_fmadd_ps(_mm512_add_ps(m, _mm512_cvtepi32_ps(p)), inv_alpha, x0);
}

struct Nvq16 {
    __m512 scale, bias, isa, sx0;
    JV_AVX512 inline __m512 dq(__m512 b) const { return logit16(_mm512_fmadd_ps(b, scale, bias), isa, sx0); }
};

JV_AVX512 inline Nvq16 nvq16(const NvqScalars &c)
{
    Nvq16 r;
    r.scale = _mm512_set1_ps(c.scale); r.bias = _mm512_set1_ps(c.bias); r.isa = _mm512_set1_ps(c.isa); r.sx0 = _mm512_set1_ps(c.sx0);
    return r;
}

JV_AVX512 inline __m512 bytes16(const unsigned char *p, __mmask16 k)
{
    return _mm512_cvtepi32_ps(_mm512_cvtepu8_epi32(_mm_maskz_loadu_epi8(k, p)));
}

}  // namespace

// ---- float32 similarities: 4 independent 16-lane accumulators (the reference keeps 4 Highway vectors, :212-229) ----
JV_AVX512 float dot_avx512(const float *a, const float *b, size_t n)
{
    __m512 s0 = _mm512_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        s0 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i), s0);
        s1 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i + 16), _mm512_loadu_ps(b + i + 16), s1);
        s2 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i + 32), _mm512_loadu_ps(b + i + 32), s2);
        s3 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i + 48), _mm512_loadu_ps(b + i + 48), s3);
    }
    for (; i + 16 <= n; i += 16) s0 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i), s0);
    if (i < n) {
        const __mmask16 k = tail_mask(n - i);
        s1 = _mm512_fmadd_ps(_mm512_maskz_loadu_ps(k, a + i), _mm512_maskz_loadu_ps(k, b + i), s1);
    }
    return hsum(_mm512_add_ps(_mm512_add_ps(s0, s1), _mm512_add_ps(s2, s3)));
}

JV_AVX512 float l2_avx512(const float *a, const float *b, size_t n)
{
    __m512 s0 = _mm512_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m512 d0 = _mm512_sub_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i));
        const __m512 d1 = _mm512_sub_ps(_mm512_loadu_ps(a + i + 16), _mm512_loadu_ps(b + i + 16));
        const __m512 d2 = _mm512_sub_ps(_mm512_loadu_ps(a + i + 32), _mm512_loadu_ps(b + i + 32));
        const __m512 d3 = _mm512_sub_ps(_mm512_loadu_ps(a + i + 48), _mm512_loadu_ps(b + i + 48));
        s0 = _mm512_fmadd_ps(d0, d0, s0); s1 = _mm512_fmadd_ps(d1, d1, s1); s2 = _mm512_fmadd_ps(d2, d2, s2); s3 = _mm512_fmadd_ps(d3, d3, s3);
    }
    for (; i + 16 <= n; i += 16) {
        const __m512 d = _mm512_sub_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i));
        s0 = _mm512_fmadd_ps(d, d, s0);
    }
    if (i < n) {
        const __mmask16 k = tail_mask(n - i);
        const __m512 d = _mm512_sub_ps(_mm512_maskz_loadu_ps(k, a + i), _mm512_maskz_loadu_ps(k, b + i));
        s1 = _mm512_fmadd_ps(d, d, s1);
    }
    return hsum(_mm512_add_ps(_mm512_add_ps(s0, s1), _mm512_add_ps(s2, s3)));
}

JV_AVX512 float cosine_avx512(const float *a, const float *b, size_t n)
{
    __m512 s0 = _mm512_setzero_ps(), s1 = s0, aa0 = s0, aa1 = s0, bb0 = s0, bb1 = s0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m512 x0 = _mm512_loadu_ps(a + i), y0 = _mm512_loadu_ps(b + i), x1 = _mm512_loadu_ps(a + i + 16), y1 = _mm512_loadu_ps(b + i + 16);
        s0 = _mm512_fmadd_ps(x0, y0, s0); aa0 = _mm512_fmadd_ps(x0, x0, aa0); bb0 = _mm512_fmadd_ps(y0, y0, bb0);
        s1 = _mm512_fmadd_ps(x1, y1, s1); aa1 = _mm512_fmadd_ps(x1, x1, aa1); bb1 = _mm512_fmadd_ps(y1, y1, bb1);
    }
    for (; i < n; i += 16) {
        const __mmask16 k = n - i >= 16 ? (__mmask16)0xffff : tail_mask(n - i);
        const __m512 x = _mm512_maskz_loadu_ps(k, a + i), y = _mm512_maskz_loadu_ps(k, b + i);
        s0 = _mm512_fmadd_ps(x, y, s0); aa0 = _mm512_fmadd_ps(x, x, aa0); bb0 = _mm512_fmadd_ps(y, y, bb0);
    }
    const float s = hsum(_mm512_add_ps(s0, s1)), am = hsum(_mm512_add_ps(aa0, aa1)), bm = hsum(_mm512_add_ps(bb0, bb1));
    return s / sqrtf(am * bm);  // native-c:src/jvector_simd_kernels.cpp:285-286
}

// ---- PQ: assemble-and-sum = 32-bit gathers from the per-query table ----
JV_AVX512 float assemble_and_sum_avx512(const float *data, int dataBase, const unsigned char *c, size_t len)
{
    __m512 acc = _mm512_setzero_ps();
    const __m512i step = _mm512_set1_epi32(16 * dataBase);
    __m512i base = _mm512_mullo_epi32(_mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15), _mm512_set1_epi32(dataBase));
    size_t i = 0;
    for (; i + 16 <= len; i += 16) {
        const __m512i idx = _mm512_add_epi32(base, _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(c + i))));
        acc = _mm512_add_ps(acc, _mm512_i32gather_ps(idx, data, 4));
        base = _mm512_add_epi32(base, step);
    }
    if (i < len) {
        const __mmask16 k = tail_mask(len - i);
        const __m512i idx = _mm512_add_epi32(base, _mm512_cvtepu8_epi32(_mm_maskz_loadu_epi8(k, c + i)));
        acc = _mm512_add_ps(acc, _mm512_mask_i32gather_ps(_mm512_setzero_ps(), k, idx, data, 4));
    }
    return hsum(acc);
}

JV_AVX512 float pq_decoded_cosine_avx512(const unsigned char *c, size_t len, int clusterCount, const float *partialSums, const float *aMagnitude, float bMagnitude)
{
    __m512 s = _mm512_setzero_ps(), a = _mm512_setzero_ps();
    const __m512i step = _mm512_set1_epi32(16 * clusterCount);
    __m512i base = _mm512_mullo_epi32(_mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15), _mm512_set1_epi32(clusterCount));
    for (size_t i = 0; i < len; i += 16) {
        const __mmask16 k = len - i >= 16 ? (__mmask16)0xffff : tail_mask(len - i);
        const __m512i idx = _mm512_add_epi32(base, _mm512_cvtepu8_epi32(_mm_maskz_loadu_epi8(k, c + i)));
        s = _mm512_add_ps(s, _mm512_mask_i32gather_ps(_mm512_setzero_ps(), k, idx, partialSums, 4));
        a = _mm512_add_ps(a, _mm512_mask_i32gather_ps(_mm512_setzero_ps(), k, idx, aMagnitude, 4));
        base = _mm512_add_epi32(base, step);
    }
    return hsum(s) / sqrtf(hsum(a) * bMagnitude);
}

JV_AVX512 float assemble_and_sum_pq_avx512(const float *data, size_t subspaceCount, const unsigned char *c1, const unsigned char *c2, int clusterCount)
{
    const int k = clusterCount, block = k * (k + 1) / 2;
    __m512 acc = _mm512_setzero_ps();
    const __m512i vk = _mm512_set1_epi32(k), one = _mm512_set1_epi32(1), step = _mm512_set1_epi32(16 * block);
    __m512i base = _mm512_mullo_epi32(_mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15), _mm512_set1_epi32(block));
    for (size_t i = 0; i < subspaceCount; i += 16) {
        const __mmask16 m = subspaceCount - i >= 16 ? (__mmask16)0xffff : tail_mask(subspaceCount - i);
        const __m512i a = _mm512_cvtepu8_epi32(_mm_maskz_loadu_epi8(m, c1 + i)), b = _mm512_cvtepu8_epi32(_mm_maskz_loadu_epi8(m, c2 + i));
        const __m512i r = _mm512_min_epi32(a, b), c = _mm512_max_epi32(a, b);
        // r k - r (r - 1) / 2 + (c - r)
        const __m512i tri = _mm512_srai_epi32(_mm512_mullo_epi32(r, _mm512_sub_epi32(r, one)), 1);
        const __m512i idx = _mm512_add_epi32(base, _mm512_add_epi32(_mm512_sub_epi32(_mm512_mullo_epi32(r, vk), tri), _mm512_sub_epi32(c, r)));
        acc = _mm512_add_ps(acc, _mm512_mask_i32gather_ps(_mm512_setzero_ps(), m, idx, data, 4));
        base = _mm512_add_epi32(base, step);
    }
    return hsum(acc);
}

// ---- PQ: partial sums ----
// sub-vector size 8 (d / M = 8, the reference's default mFactor): one 512-bit load = two centroids; four loads are folded by
// in-lane horizontal adds into [A0 A1 A2 A3 | . | B0 B1 B2 B3 | .] (A/B = first / second centroid of each load), eight sums a step
JV_AVX512 static inline __m512 hadd16(__m512 a, __m512 b) { return _mm512_add_ps(_mm512_shuffle_ps(a, b, 0x88), _mm512_shuffle_ps(a, b, 0xDD)); }

JV_AVX512 static void partial_sums8_avx512(const float *codebook, int clusterCount, const float *q, float *out, int mode)
{
    __m512 qq = _mm512_setzero_ps();
    if (mode != 2) qq = _mm512_broadcast_f32x8(_mm256_loadu_ps(q));
    int c = 0;
    for (; c + 8 <= clusterCount; c += 8) {
        __m512 p[4];
        for (int i = 0; i < 4; i++) {
            const __m512 x = _mm512_loadu_ps(codebook + (size_t)(c + 2 * i) * 8);
            if (mode == 0) p[i] = _mm512_mul_ps(x, qq);
            else if (mode == 1) {
                const __m512 d = _mm512_sub_ps(x, qq);
                p[i] = _mm512_mul_ps(d, d);
            } else p[i] = _mm512_mul_ps(x, x);
        }
        const __m512 h = hadd16(hadd16(p[0], p[1]), hadd16(p[2], p[3]));       // per 128-bit lane: [sum4(p0) sum4(p1) sum4(p2) sum4(p3)]
        const __m512 r = _mm512_add_ps(h, _mm512_shuffle_f32x4(h, h, 0xB1));    // lane 0: first centroids, lane 2: second centroids
        const __m128 a = _mm512_castps512_ps128(r), b = _mm512_extractf32x4_ps(r, 2);
        _mm_storeu_ps(out + c, _mm_unpacklo_ps(a, b));
        _mm_storeu_ps(out + c + 4, _mm_unpackhi_ps(a, b));
    }
    for (; c < clusterCount; c++) {
        const float *cen = codebook + (size_t)c * 8;
        float sacc = 0.f;
        for (int j = 0; j < 8; j++) {
            if (mode == 0) sacc = fmaf(cen[j], q[j], sacc);
            else if (mode == 1) { const float d = cen[j] - q[j]; sacc = fmaf(d, d, sacc); }
            else sacc = fmaf(cen[j], cen[j], sacc);
        }
        out[c] = sacc;
    }
}

// any other sub-vector size: 16 centroids at a time, dimension j of all 16 fetched by one strided gather
JV_AVX512 void partial_sums_avx512(const float *codebook, size_t size, int clusterCount, const float *q, float *out, int mode /* 0 dot, 1 l2, 2 self */)
{
    if (size == 8) { partial_sums8_avx512(codebook, clusterCount, q, out, mode); return; }
    const __m512i lanes = _mm512_mullo_epi32(_mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15), _mm512_set1_epi32((int)size));
    for (int c0 = 0; c0 < clusterCount; c0 += 16) {
        const __mmask16 k = clusterCount - c0 >= 16 ? (__mmask16)0xffff : tail_mask((size_t)(clusterCount - c0));
        const float *cb = codebook + (size_t)c0 * size;
        __m512 acc = _mm512_setzero_ps();
        for (size_t j = 0; j < size; j++) {
            const __m512 col = _mm512_mask_i32gather_ps(_mm512_setzero_ps(), k, lanes, cb + j, 4);
            if (mode == 0) acc = _mm512_fmadd_ps(col, _mm512_set1_ps(q[j]), acc);
            else if (mode == 1) {
                const __m512 d = _mm512_sub_ps(col, _mm512_set1_ps(q[j]));
                acc = _mm512_fmadd_ps(d, d, acc);
            } else acc = _mm512_fmadd_ps(col, col, acc);
        }
        _mm512_mask_storeu_ps(out + c0, k, acc);
    }
}

// ---- NVQ ----
JV_AVX512 void nvq_quantize_avx512(const float *v, size_t n, float sa, float c0, float bias, float inv, unsigned char *dst)
{
    const __m512 vsa = _mm512_set1_ps(sa), vc0 = _mm512_set1_ps(c0), vbias = _mm512_set1_ps(bias), vinv = _mm512_set1_ps(inv), half = _mm512_set1_ps(0.5f);
    for (size_t i = 0; i < n; i += 16) {
        const __mmask16 k = n - i >= 16 ? (__mmask16)0xffff : tail_mask(n - i);
        const __m512 x = _mm512_maskz_loadu_ps(k, v + i);
        const __m512 a = _mm512_fmadd_ps(_mm512_sub_ps(logistic16(x, vsa, vc0), vbias), vinv, half);
        __m512i q = _mm512_cvttps_epi32(a);
        q = _mm512_min_epi32(_mm512_max_epi32(q, _mm512_setzero_si512()), _mm512_set1_epi32(255));
        _mm_mask_storeu_epi8(dst + i, k, _mm512_cvtepi32_epi8(q));
    }
}

JV_AVX512 float nvq_loss_avx512(const float *v, size_t n, const NvqScalars &c, float c0)
{
    const Nvq16 d = nvq16(c);
    const __m512 vsa = _mm512_set1_ps(c.sa), vc0 = _mm512_set1_ps(c0), vbias = _mm512_set1_ps(c.bias), vinv = _mm512_set1_ps(1.0f / c.scale), half = _mm512_set1_ps(0.5f);
    __m512 acc = _mm512_setzero_ps();
    for (size_t i = 0; i < n; i += 16) {
        const __mmask16 k = n - i >= 16 ? (__mmask16)0xffff : tail_mask(n - i);
        const __m512 x = _mm512_maskz_loadu_ps(k, v + i);
        const __m512 r = _mm512_mul_ps(_mm512_sub_ps(logistic16(x, vsa, vc0), vbias), vinv);
        const __m512 rq = _mm512_cvtepi32_ps(_mm512_cvttps_epi32(_mm512_add_ps(r, half)));
        const __m512 df = _mm512_sub_ps(x, d.dq(rq));
        acc = _mm512_mask3_fmadd_ps(df, df, acc, k);
    }
    return hsum(acc);
}

JV_AVX512 float nvq_uniform_loss_avx512(const float *v, size_t n, float minv, float maxv, float constant)
{
    const float delta = maxv - minv;
    const __m512 vmin = _mm512_set1_ps(minv), s1 = _mm512_set1_ps(constant / delta), s2 = _mm512_set1_ps(delta / constant), half = _mm512_set1_ps(0.5f);
    __m512 acc = _mm512_setzero_ps();
    for (size_t i = 0; i < n; i += 16) {
        const __mmask16 k = n - i >= 16 ? (__mmask16)0xffff : tail_mask(n - i);
        const __m512 x = _mm512_maskz_loadu_ps(k, v + i);
        const __m512 r = _mm512_mul_ps(_mm512_sub_ps(x, vmin), s1);
        const __m512 rec = _mm512_fmadd_ps(_mm512_cvtepi32_ps(_mm512_cvttps_epi32(_mm512_add_ps(r, half))), s2, vmin);
        const __m512 df = _mm512_sub_ps(x, rec);
        acc = _mm512_mask3_fmadd_ps(df, df, acc, k);
    }
    return hsum(acc);
}

JV_AVX512 float nvq_dot_avx512(const float *q, const unsigned char *b, size_t n, const NvqScalars &c)
{
    const Nvq16 d = nvq16(c);
    __m512 acc = _mm512_setzero_ps();
    for (size_t i = 0; i < n; i += 16) {
        const __mmask16 k = n - i >= 16 ? (__mmask16)0xffff : tail_mask(n - i);
        acc = _mm512_mask3_fmadd_ps(_mm512_maskz_loadu_ps(k, q + i), d.dq(bytes16(b + i, k)), acc, k);
    }
    return hsum(acc);
}

JV_AVX512 float nvq_l2_avx512(const float *q, const unsigned char *b, size_t n, const NvqScalars &c)
{
    const Nvq16 d = nvq16(c);
    __m512 acc = _mm512_setzero_ps();
    for (size_t i = 0; i < n; i += 16) {
        const __mmask16 k = n - i >= 16 ? (__mmask16)0xffff : tail_mask(n - i);
        const __m512 df = _mm512_sub_ps(_mm512_maskz_loadu_ps(k, q + i), d.dq(bytes16(b + i, k)));
        acc = _mm512_mask3_fmadd_ps(df, df, acc, k);
    }
    return hsum(acc);
}

JV_AVX512 void nvq_cosine_avx512(const float *q, const unsigned char *b, size_t n, const NvqScalars &c, const float *centroid, float *sum_out, float *mag_out)
{
    const Nvq16 d = nvq16(c);
    __m512 s = _mm512_setzero_ps(), bm = _mm512_setzero_ps();
    for (size_t i = 0; i < n; i += 16) {
        const __mmask16 k = n - i >= 16 ? (__mmask16)0xffff : tail_mask(n - i);
        const __m512 e = _mm512_add_ps(d.dq(bytes16(b + i, k)), _mm512_maskz_loadu_ps(k, centroid + i));
        s = _mm512_mask3_fmadd_ps(_mm512_maskz_loadu_ps(k, q + i), e, s, k);
        bm = _mm512_mask3_fmadd_ps(e, e, bm, k);
    }
    *sum_out = hsum(s);
    *mag_out = hsum(bm);
}

bool cpu_has_avx512()
{
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                           __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("fma");
    return ok;
}

}  // namespace jvl
