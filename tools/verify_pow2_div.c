/* RoIAlign window kernel (tensorrtx_b200/csrc/roi_align.cu, roi_window_taps<.., POW2>): when the sample count of a bin is a power
 * of two the kernel stores `output_val * (1.0f / count)` instead of the reference's `output_val / count` (rcnn/RoiAlign.cu:147).
 * Claim: identical bits for EVERY float (normals, denormal results, infinities, NaN payloads) and every count 2^k, k = 0..8
 * (a sampling grid is at most 18 x 18 per bin).  This program checks all 2^32 bit patterns (or every `stride`-th one).
 *   gcc -O2 -o verify_pow2_div verify_pow2_div.c && ./verify_pow2_div [stride]      exit code 0 = no mismatch */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    uint64_t checked = 0, bad = 0;
    for (int k = 0; k <= 8; ++k) {
        const float count = (float)(1 << k);
        const volatile float inv_v = 1.0f / count;  /* what the kernel multiplies with */
        const float inv = inv_v;
        for (uint64_t u = 0; u < (1ull << 32); u += stride) {
            const uint32_t b = (uint32_t)u;
            float x; memcpy(&x, &b, 4);
            const float q = x / count, m = x * inv;
            const uint32_t bq = bits(q), bm = bits(m);
            ++checked;
            if (bq != bm && !(q != q && m != m && ((bq ^ bm) & 0x7fffffffu) == 0)) {  /* NaN: same payload required */
                if (bad < 10) fprintf(stderr, "mismatch: x=%08x count=%g  div=%08x mul=%08x\n", b, count, bq, bm);
                ++bad;
            }
        }
    }
    printf("checked %llu (x, count) pairs, %llu mismatches\n", (unsigned long long)checked, (unsigned long long)bad);
    return bad ? 1 : 0;
}
