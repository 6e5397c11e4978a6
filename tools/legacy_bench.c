// Times the 24-symbol legacy ABI of libjvector_b200.so beside the reference's own libjvector.so (oracle/_ref) on one host core:
//   gcc -O2 -o /tmp/legacy_bench tools/legacy_bench.c -ldl -lm && /tmp/legacy_bench jvector_b200/lib/libjvector_b200.so oracle/_ref/libjvector.so
// Shapes: d = 768, PQ M = 96 / k = 256 / sub-vector size 8 (BASELINE config 3), NVQ sub-vector of 384.
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef float (*f_adc)(const float *, int, const unsigned char *, int, size_t);
typedef float (*f_adcpq)(const float *, size_t, const unsigned char *, int, const unsigned char *, int, int);
typedef float (*f_pqcos)(const unsigned char *, int, size_t, int, const float *, const float *, float);
typedef void (*f_ps)(const float *, int, size_t, int, const float *, int, float *);
typedef void (*f_psm)(const float *, int, size_t, int, float *);
typedef void (*f_nvqq)(const float *, size_t, float, float, float, float, unsigned char *);
typedef float (*f_nvql)(const float *, size_t, float, float, float, float, int);
typedef float (*f_nvqu)(const float *, size_t, float, float, int);
typedef float (*f_nvqd)(const float *, const unsigned char *, size_t, float, float, float, float);
typedef int64_t (*f_nvqc)(const float *, const unsigned char *, size_t, float, float, float, float, const float *);
typedef void (*f_shuf)(float *, size_t);
typedef float (*f_sim)(const float *, size_t, const float *, size_t, size_t);

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static float frand(void) { return (float)rand() / RAND_MAX - 0.5f; }

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s ours.so ref.so\n", argv[0]); return 2; }
    void *libs[2] = {dlopen(argv[1], RTLD_NOW | RTLD_LOCAL), dlopen(argv[2], RTLD_NOW | RTLD_LOCAL)};
    if (!libs[0] || !libs[1]) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    const int M = 96, k = 256, sz = 8, dim = 768, nv = 384, ncodes = 4096;
    float *lut = malloc(sizeof(float) * M * k), *mag = malloc(sizeof(float) * M * k), *cb = malloc(sizeof(float) * k * sz), *q = malloc(sizeof(float) * dim);
    float *table = malloc(sizeof(float) * (size_t)M * (k * (k + 1) / 2)), *v = malloc(sizeof(float) * nv), *cen = malloc(sizeof(float) * nv), *a = malloc(sizeof(float) * dim);
    unsigned char *codes = malloc((size_t)ncodes * M), *qb = malloc(nv);
    for (int i = 0; i < M * k; i++) { lut[i] = frand(); mag[i] = 0.1f + frand() * frand(); }
    for (int i = 0; i < k * sz; i++) cb[i] = frand();
    for (int i = 0; i < dim; i++) { q[i] = frand(); a[i] = frand(); }
    for (size_t i = 0; i < (size_t)M * (k * (k + 1) / 2); i++) table[i] = frand();
    for (int i = 0; i < nv; i++) { v[i] = 0.1f * frand(); cen[i] = 0.01f * frand(); }
    for (size_t i = 0; i < (size_t)ncodes * M; i++) codes[i] = (unsigned char)(rand() & 255);
    float minv = 1e9f, maxv = -1e9f;
    for (int i = 0; i < nv; i++) { if (v[i] < minv) minv = v[i]; if (v[i] > maxv) maxv = v[i]; }
    printf("%-44s %12s %12s %8s\n", "symbol (shape)", "ours ns", "ref ns", "ours/ref");
    double t[2];
    volatile float sink = 0.f;
#define TIME(label, reps, body)                                            \
    for (int L = 0; L < 2; L++) {                                          \
        void *h = libs[L];                                                 \
        (void)h;                                                           \
        for (int w = 0; w < 2; w++) {                                      \
            const double t0 = now();                                       \
            for (int r = 0; r < (reps); r++) { body; }                     \
            t[L] = (now() - t0) / (reps) * 1e9;                            \
        }                                                                  \
    }                                                                      \
    printf("%-44s %12.1f %12.1f %8.2f\n", label, t[0], t[1], t[0] / t[1]);
    TIME("dot_product_f32 (768)", 200000, sink += ((f_sim)dlsym(h, "dot_product_f32"))(a, 0, q, 0, dim))
    { f_sim fn[2] = {(f_sim)dlsym(libs[0], "dot_product_f32"), (f_sim)dlsym(libs[1], "dot_product_f32")};
      TIME("dot_product_f32 (768), symbol resolved once", 2000000, sink += fn[L](a, 0, q, 0, dim)) }
    { f_adc fn[2] = {(f_adc)dlsym(libs[0], "assemble_and_sum_f32"), (f_adc)dlsym(libs[1], "assemble_and_sum_f32")};
      TIME("assemble_and_sum_f32 (M=96)", 2000000, sink += fn[L](lut, k, codes, (r % ncodes) * M, M)) }
    { f_pqcos fn[2] = {(f_pqcos)dlsym(libs[0], "pq_decoded_cosine_similarity_f32"), (f_pqcos)dlsym(libs[1], "pq_decoded_cosine_similarity_f32")};
      TIME("pq_decoded_cosine_similarity_f32 (M=96)", 2000000, sink += fn[L](codes, (r % ncodes) * M, M, k, lut, mag, 1.0f)) }
    { f_adcpq fn[2] = {(f_adcpq)dlsym(libs[0], "assemble_and_sum_pq_f32"), (f_adcpq)dlsym(libs[1], "assemble_and_sum_pq_f32")};
      TIME("assemble_and_sum_pq_f32 (M=96)", 1000000, sink += fn[L](table, M, codes, (r % ncodes) * M, codes, ((r * 7 + 1) % ncodes) * M, k)) }
    { f_ps fn[2] = {(f_ps)dlsym(libs[0], "calculate_partial_sums_dot_f32"), (f_ps)dlsym(libs[1], "calculate_partial_sums_dot_f32")};
      TIME("calculate_partial_sums_dot_f32 (k=256, size 8)", 200000, fn[L](cb, r % M, sz, k, q, (r % M) * sz, lut)) }
    { f_ps fn[2] = {(f_ps)dlsym(libs[0], "calculate_partial_sums_euclidean_f32"), (f_ps)dlsym(libs[1], "calculate_partial_sums_euclidean_f32")};
      TIME("calculate_partial_sums_euclidean_f32 (k=256,8)", 200000, fn[L](cb, r % M, sz, k, q, (r % M) * sz, lut)) }
    { f_psm fn[2] = {(f_psm)dlsym(libs[0], "calculate_partial_sums_self_magnitude_f32"), (f_psm)dlsym(libs[1], "calculate_partial_sums_self_magnitude_f32")};
      TIME("calculate_partial_sums_self_magnitude (256,8)", 200000, fn[L](cb, r % M, sz, k, mag)) }
    for (int i = 0; i < M * k; i++) { lut[i] = frand(); mag[i] = 0.1f + frand() * frand(); }
    { f_nvqq fn[2] = {(f_nvqq)dlsym(libs[0], "nvq_quantize_8bit"), (f_nvqq)dlsym(libs[1], "nvq_quantize_8bit")};
      TIME("nvq_quantize_8bit (384)", 200000, fn[L](v, nv, 2.5f, 0.f, minv, maxv, qb)) }
    { f_nvql fn[2] = {(f_nvql)dlsym(libs[0], "nvq_loss"), (f_nvql)dlsym(libs[1], "nvq_loss")};
      TIME("nvq_loss (384)", 200000, sink += fn[L](v, nv, 2.5f, 0.f, minv, maxv, 8)) }
    { f_nvqu fn[2] = {(f_nvqu)dlsym(libs[0], "nvq_uniform_loss"), (f_nvqu)dlsym(libs[1], "nvq_uniform_loss")};
      TIME("nvq_uniform_loss (384)", 500000, sink += fn[L](v, nv, minv, maxv, 8)) }
    { f_nvqd fn[2] = {(f_nvqd)dlsym(libs[0], "nvq_dot_product_8bit"), (f_nvqd)dlsym(libs[1], "nvq_dot_product_8bit")};
      TIME("nvq_dot_product_8bit (384)", 500000, sink += fn[L](q, qb, nv, 2.5f, 0.f, minv, maxv)) }
    { f_nvqd fn[2] = {(f_nvqd)dlsym(libs[0], "nvq_square_l2_distance_8bit"), (f_nvqd)dlsym(libs[1], "nvq_square_l2_distance_8bit")};
      TIME("nvq_square_l2_distance_8bit (384)", 500000, sink += fn[L](q, qb, nv, 2.5f, 0.f, minv, maxv)) }
    { f_nvqc fn[2] = {(f_nvqc)dlsym(libs[0], "nvq_cosine_8bit_packed"), (f_nvqc)dlsym(libs[1], "nvq_cosine_8bit_packed")};
      TIME("nvq_cosine_8bit_packed (384)", 500000, sink += (float)fn[L](q, qb, nv, 2.5f, 0.f, minv, maxv, cen)) }
    return 0;
}
