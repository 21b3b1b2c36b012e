/* ORACLE — TEST INFRASTRUCTURE ONLY (not product code; see oracle/README.md).
 *
 * Counter-based pseudo-normal generator shared by the oracle and (as an independent CUDA
 * re-implementation, teochat_b200/csrc/init_kernels.cu) the product, so that "random-init
 * CLIP-L / LLaMA-2-7B weights" (BASELINE.json configs) are bit-identical on the CPU box and
 * on the GPU without shipping 13 GB fixtures.
 *
 *   u64 x   = splitmix64(seed + (i+1) * 0x9E3779B97F4A7C15)
 *   s       = sum of the four 16-bit fields of x          (Irwin-Hall, n = 4)
 *   value_i = (float)(s - 131070) * scale                 (one exactly-rounded fp32 multiply)
 *
 * with scale = std / sqrt(4 * (65536^2 - 1) / 12) computed by the caller in double and
 * rounded to float.  Integer arithmetic + one IEEE multiply ⇒ identical bits everywhere.
 * Replaces the reference's `normal_(std=...)` init (modeling_image.py:179-230; HF Llama
 * `_init_weights`), whose torch RNG stream is not reproducible across devices.
 */
#include <stdint.h>
#include <stddef.h>
#include <pthread.h>
#include <unistd.h>

static inline uint64_t splitmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

typedef struct { void *out; size_t lo, hi; uint64_t seed; float scale, mean; int kind; } job_t;

static void *run_job(void *arg) {
    job_t *j = (job_t *)arg;
    for (size_t i = j->lo; i < j->hi; ++i) {
        uint64_t x = splitmix64(j->seed + ((uint64_t)i + 1ULL) * 0x9E3779B97F4A7C15ULL);
        if (j->kind == 0) {
            int32_t s = (int32_t)(x & 0xFFFF) + (int32_t)((x >> 16) & 0xFFFF) +
                        (int32_t)((x >> 32) & 0xFFFF) + (int32_t)((x >> 48) & 0xFFFF);
            ((float *)j->out)[i] = j->mean + (float)(s - 131070) * j->scale;
        } else {
            ((uint8_t *)j->out)[i] = (uint8_t)(x >> 56);   /* top byte */
        }
    }
    return 0;
}

static void run_split(void *out, size_t n, uint64_t seed, float scale, float mean, int kind) {
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if (n < (size_t)1 << 16) nt = 1;
    pthread_t th[64];
    job_t jobs[64];
    size_t per = (n + (size_t)nt - 1) / (size_t)nt;
    for (long t = 0; t < nt; ++t) {
        size_t lo = per * (size_t)t, hi = lo + per;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        jobs[t] = (job_t){out, lo, hi, seed, scale, mean, kind};
        pthread_create(&th[t], 0, run_job, &jobs[t]);
    }
    for (long t = 0; t < nt; ++t) pthread_join(th[t], 0);
}

void teo_oracle_hash_normal_f32(float *out, size_t n, uint64_t seed, float scale, float mean) {
    run_split(out, n, seed, scale, mean, 0);
}

/* uniform bytes for synthetic frames: byte i = top byte of splitmix64(seed + (i+1)*golden) */
void teo_oracle_hash_u8(uint8_t *out, size_t n, uint64_t seed) {
    run_split(out, n, seed, 0.f, 0.f, 1);
}
