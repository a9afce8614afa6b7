// Stress check of the division-free overlap test of nms.cu (`overlaps`):
//     t = fmaf(-thr, denom, inter);  if (denom > 0 && |t| > 1e-5f*denom) decide by (t > 0)  else  inter/denom > thr
// must give the SAME decision as the IEEE division `inter/denom > thr` for every (inter, denom, thr) with |thr| <= 4.
// The risky region is a quotient within a few ulps of thr, so besides uniformly random operands the generator builds
// adversarial cases: denom random, inter = the float nearest thr*denom nudged by -40..+40 ulps.
// usage: verify_overlap_test [cases_per_thread]      (prints mismatches; exit code 1 if any)
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t rng_next(uint64_t* s) {  // xorshift64*
    uint64_t x = *s;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    *s = x;
    return x * 2685821657736338717ULL;
}
static inline float urand(uint64_t* s) { return (float)(rng_next(s) >> 40) * (1.0f / 16777216.0f); }
static inline float nudge(float x, int ulps) {
    int32_t b;
    memcpy(&b, &x, 4);
    b += ulps;  // x > 0: monotone in the integer representation
    memcpy(&x, &b, 4);
    return x;
}
static inline int fast_or_exact(float inter, float denom, float thr, int* used_fast) {
    const float t = fmaf(-thr, denom, inter);
    if (denom > 0.0f && fabsf(t) > 1e-5f * denom) {
        *used_fast = 1;
        return t > 0.0f;
    }
    *used_fast = 0;
    return inter / denom > thr;
}

int main(int argc, char** argv) {
    const long per_thread = argc > 1 ? atol(argv[1]) : 20000000L;
    long bad = 0, fast = 0, total = 0;
    const float thrs[] = {0.45f, 0.5f, 0.4f, 0.7f, 0.3f, 0.05f, 0.95f, 1.0f, 0.0f, 1e-3f, 2.5f, 4.0f, -0.5f, 0.6499999f};
    const int nthr = (int)(sizeof(thrs) / sizeof(thrs[0]));
#pragma omp parallel reduction(+ : bad, fast, total)
    {
        uint64_t s = 0x9E3779B97F4A7C15ULL ^ ((uint64_t)(omp_get_thread_num() + 1) * 0xD1B54A32D192ED03ULL);
        for (long i = 0; i < per_thread; ++i) {
            const float thr = thrs[rng_next(&s) % nthr];
            float inter, denom;
            const unsigned mode = (unsigned)(rng_next(&s) & 3);
            const float scale = exp2f((float)((int)(rng_next(&s) % 40) - 8));  // box areas from 2^-8 to 2^31
            denom = (0.001f + urand(&s)) * scale;
            if (mode == 0) {  // unrelated operands (inter <= denom as for real boxes, sometimes larger)
                inter = urand(&s) * denom * (1.0f + 0.1f * (float)(rng_next(&s) & 1));
            } else {  // adversarial: quotient within a few dozen ulps of the threshold
                inter = thr * denom;
                if (inter > 0.0f) inter = nudge(inter, (int)(rng_next(&s) % 81) - 40);
                if (mode == 3) denom = nudge(denom, (int)(rng_next(&s) % 9) - 4);
            }
            int used_fast;
            const int got = fast_or_exact(inter, denom, thr, &used_fast);
            const int ref = inter / denom > thr;
            fast += used_fast;
            ++total;
            if (got != ref) {
                ++bad;
                if (bad < 5) printf("MISMATCH inter=%a denom=%a thr=%a got=%d ref=%d\n", inter, denom, thr, got, ref);
            }
        }
    }
    printf("cases=%ld fast_path=%ld mismatches=%ld\n", total, fast, bad);
    return bad != 0;
}
