// legacy_simd.cpp — the three float32 similarity symbols of the reference's libjvector.so ABI
// (/root/reference/jvector-native/src/main/native/src/jvector_simd_kernel_list.h:37-39), host-synchronous by contract (see
// legacy_host.cpp). Written as fixed-width lane arrays so the host compiler vectorises them; target_clones gives AVX-512 /
// AVX2+FMA / baseline versions behind one symbol (resolved once at load, like the reference's CPUID dispatcher,
// jvector_simd.cpp:124-167). Summation order: 64 independent lanes, then a pairwise tree — within the reference's own
// tolerance between providers (rel 1e-4, native-c:tests/test_similarity.cpp).
#include <math.h>
#include <stddef.h>

#include "../../include/jvector_b200.h"
#include "legacy_avx512.h"

#define JV_CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))

namespace {
constexpr int W = 64;
inline float reduce(float *acc)
{
    for (int w = W / 2; w > 0; w >>= 1)
        for (int l = 0; l < w; l++) acc[l] += acc[l + w];
    return acc[0];
}
}  // namespace

extern "C" {

JV_CLONES float dot_product_f32(const float *a, size_t aoffset, const float *b, size_t boffset, size_t length)
{
    a += aoffset;
    b += boffset;
    if (jvl::cpu_has_avx512()) return jvl::dot_avx512(a, b, length);
    float acc[W] = {0.f};
    size_t i = 0;
    for (; i + W <= length; i += W)
        for (int l = 0; l < W; l++) acc[l] += a[i + l] * b[i + l];
    for (int l = 0; i < length; i++, l++) acc[l] += a[i] * b[i];
    return reduce(acc);
}

JV_CLONES float euclidean_f32(const float *a, size_t aoffset, const float *b, size_t boffset, size_t length)
{
    a += aoffset;
    b += boffset;
    if (jvl::cpu_has_avx512()) return jvl::l2_avx512(a, b, length);
    float acc[W] = {0.f};
    size_t i = 0;
    for (; i + W <= length; i += W)
        for (int l = 0; l < W; l++) {
            const float d = a[i + l] - b[i + l];
            acc[l] += d * d;
        }
    for (int l = 0; i < length; i++, l++) {
        const float d = a[i] - b[i];
        acc[l] += d * d;
    }
    return reduce(acc);
}

JV_CLONES float cosine_f32(const float *a, size_t aoffset, const float *b, size_t boffset, size_t length)
{
    a += aoffset;
    b += boffset;
    if (jvl::cpu_has_avx512()) return jvl::cosine_avx512(a, b, length);
    float s[W] = {0.f}, aa[W] = {0.f}, bb[W] = {0.f};
    size_t i = 0;
    for (; i + W <= length; i += W)
        for (int l = 0; l < W; l++) {
            const float x = a[i + l], y = b[i + l];
            s[l] += x * y;
            aa[l] += x * x;
            bb[l] += y * y;
        }
    for (int l = 0; i < length; i++, l++) {
        const float x = a[i], y = b[i];
        s[l] += x * y;
        aa[l] += x * x;
        bb[l] += y * y;
    }
    return reduce(s) / sqrtf(reduce(aa) * reduce(bb));  // native-c:src/jvector_simd_kernels.cpp:285-286
}

}  // extern "C"
