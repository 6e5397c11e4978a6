// legacy_avx512.h — AVX-512 bodies of the legacy libjvector.so symbols (see legacy_avx512.cpp); internal.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace jvl {

struct NvqScalars {  // the derived scalars of one NVQ sub-vector (legacy_host.cpp struct Nvq)
    float sa, isa, sx0, bias, scale;
};

bool cpu_has_avx512();
float dot_avx512(const float *a, const float *b, size_t n);
float l2_avx512(const float *a, const float *b, size_t n);
float cosine_avx512(const float *a, const float *b, size_t n);
float assemble_and_sum_avx512(const float *data, int dataBase, const unsigned char *codes, size_t len);
float pq_decoded_cosine_avx512(const unsigned char *codes, size_t len, int clusterCount, const float *partialSums, const float *aMagnitude, float bMagnitude);
float assemble_and_sum_pq_avx512(const float *data, size_t subspaceCount, const unsigned char *c1, const unsigned char *c2, int clusterCount);
void partial_sums_avx512(const float *codebook, size_t size, int clusterCount, const float *q, float *out, int mode);
void nvq_quantize_avx512(const float *v, size_t n, float sa, float c0, float bias, float inv, unsigned char *dst);
float nvq_loss_avx512(const float *v, size_t n, const NvqScalars &c, float c0);
float nvq_uniform_loss_avx512(const float *v, size_t n, float minv, float maxv, float constant);
float nvq_dot_avx512(const float *q, const unsigned char *b, size_t n, const NvqScalars &c);
float nvq_l2_avx512(const float *q, const unsigned char *b, size_t n, const NvqScalars &c);
void nvq_cosine_avx512(const float *q, const unsigned char *b, size_t n, const NvqScalars &c, const float *centroid, float *sum_out, float *mag_out);

}  // namespace jvl
