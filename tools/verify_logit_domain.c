// Stress check of the logit-domain class loop of the YoloLayer scan kernels (yolo_layout.cuh: Best / update_one /
// merge_one / finish_best, yolo_decode.cu: scan_classes with its group-max gate) against the reference loop
//     max = 0; cls = 0; for i: p = Logist(x_i); if (p > max) { max = p; cls = i; }   if (max < gate) drop
// (yolov8/plugin/yololayer.cu:193-203).  Same Logist (1/(1+expf(-x))) on both sides; what is verified is the
// ALGORITHM: running max logit + first argmax + max-before-it, class slices merged in ascending order, group replay
// only when a group raises the running maximum, one sigmoid at the end, collision replay.  Inputs: background noise
// with planted logits a few ulps apart, exact ties, saturating values, -inf / NaN / < -88 (probability underflow).
// usage: verify_logit_domain [cases_per_thread]
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NC 80
static inline float logist(float x) { return 1.0f / (1.0f + expf(-x)); }
static inline uint64_t rng_next(uint64_t* s) {
    uint64_t x = *s;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    *s = x;
    return x * 2685821657736338717ULL;
}
static inline float urand(uint64_t* s) { return (float)(rng_next(s) >> 40) * (1.0f / 16777216.0f); }
static inline float nrand(uint64_t* s) {  // Box-Muller
    float u = urand(s) + 1e-7f, v = urand(s);
    return sqrtf(-2.0f * logf(u)) * cosf(6.2831853f * v);
}
static inline float nudge(float x, int ulps) {
    int32_t b;
    memcpy(&b, &x, 4);
    b += (x >= 0) ? ulps : -ulps;
    memcpy(&x, &b, 4);
    return x;
}

typedef struct {
    float bx, b2;
    int bc;
} best_t;
static inline void update_one(best_t* s, float x, int cls) {
    const int up = x > s->bx;
    s->b2 = up ? s->bx : s->b2;
    s->bc = up ? cls : s->bc;
    s->bx = up ? x : s->bx;
}
static inline void merge_one(best_t* s, float m, float m2, int c) {
    if (m > s->bx) {
        s->b2 = fmaxf(s->bx, m2);
        s->bx = m;
        s->bc = c;
    }
}

// ours: returns 1 if the anchor passes the gate; *P, *cls as the kernel would store them
static int ours(const float* x, int slices, int U, float gate, float x_lo, float* P, int* cls, long* replays) {
    best_t tot = {x_lo, x_lo, 0};
    const int per = (NC + slices - 1) / slices;
    for (int w = 0; w < slices; ++w) {
        best_t s = {x_lo, x_lo, 0};
        const int c0 = w * per, c1 = c0 + per < NC ? c0 + per : NC;
        for (int r = c0; r < c1; r += U) {  // scan_classes: group max, replay only if it raises the running maximum
            float m = -INFINITY;
            for (int u = 0; u < U && r + u < c1; ++u) m = fmaxf(m, x[r + u]);
            if (m > s.bx)
                for (int u = 0; u < U && r + u < c1; ++u) update_one(&s, x[r + u], r + u);
        }
        if (w == 0) tot = s; else merge_one(&tot, s.bx, s.b2, s.bc);
    }
    float bp = 0.0f;
    int c = tot.bc;
    if (tot.bx > x_lo) {  // finish_best
        const float p = logist(tot.bx), p2 = logist(tot.b2);
        bp = p;
        if (!(p < gate)) {
            if (p == 0.0f) {
                c = 0;
            } else if (p2 == p) {
                ++*replays;
                for (int i = 0; i < c; ++i)
                    if (logist(x[i]) == p) {
                        c = i;
                        break;
                    }
            }
        }
    }
    *P = bp;
    *cls = c;
    return !(bp < gate);
}

int main(int argc, char** argv) {
    const long per_thread = argc > 1 ? atol(argv[1]) : 2000000L;
    long bad = 0, total = 0, passed = 0, replays = 0;
    const float gates[] = {0.1f, 0.25f, 0.5f, 0.0f, 1.0f, -1.0f, 0.999f};
    const int ng = (int)(sizeof(gates) / sizeof(gates[0]));
    const int slice_opts[] = {1, 2, 4, 8}, u_opts[] = {4, 5, 8, 10, 16, 20};
#pragma omp parallel reduction(+ : bad, total, passed, replays)
    {
        uint64_t s = 0xA0761D6478BD642FULL ^ ((uint64_t)(omp_get_thread_num() + 1) * 0xE7037ED1A0B428DBULL);
        float x[NC];
        for (long it = 0; it < per_thread; ++it) {
            const float gate = gates[rng_next(&s) % ng];
            float x_lo;  // yolo_fill_args
            if (gate <= 0.0f) x_lo = -INFINITY;
            else if (gate >= 1.0f) x_lo = 10.0f;
            else x_lo = logf(gate / (1.0f - gate)) - 0.05f;
            for (int i = 0; i < NC; ++i) x[i] = -7.0f + nrand(&s);
            const unsigned kind = (unsigned)(rng_next(&s) % 8);
            const int a = (int)(rng_next(&s) % NC), b = (int)(rng_next(&s) % NC), c = (int)(rng_next(&s) % NC);
            const float v = kind == 5 ? 12.0f + 12.0f * urand(&s) : -3.0f + 9.0f * urand(&s);
            if (kind >= 1) x[a] = v;
            if (kind == 2 || kind == 5) {  // near-ulp pair / triple
                x[b] = nudge(v, -(int)(rng_next(&s) % 12));
                x[c] = nudge(v, -(int)(rng_next(&s) % 12));
            }
            if (kind == 3) x[b] = v;  // exact tie
            if (kind == 4) {          // saturation: several classes with probability 1.0f
                x[a] = 17.0f + 20.0f * urand(&s);
                x[b] = 17.0f + 20.0f * urand(&s);
                x[c] = 17.0f + 20.0f * urand(&s);
            }
            if (kind == 6) {  // non-finite and underflowing logits
                x[b] = (rng_next(&s) & 1) ? -INFINITY : NAN;
                x[c] = -90.0f - 20.0f * urand(&s);
                if (rng_next(&s) & 1)
                    for (int i = 0; i < NC; ++i) x[i] = -89.0f - 30.0f * urand(&s);
            }
            if (kind == 7) x[c] = INFINITY;
            // reference
            float max = 0.0f;
            int rc = 0;
            for (int i = 0; i < NC; ++i) {
                const float p = logist(x[i]);
                if (p > max) {
                    max = p;
                    rc = i;
                }
            }
            const int ref_pass = !(max < gate);
            const int slices = slice_opts[rng_next(&s) % 4], U = u_opts[rng_next(&s) % 6];
            float P;
            int cls;
            const int pass = ours(x, slices, U, gate, x_lo, &P, &cls, &replays);
            ++total;
            passed += ref_pass;
            if (pass != ref_pass || (ref_pass && (memcmp(&P, &max, 4) != 0 || cls != rc))) {
                ++bad;
                if (bad < 5) printf("MISMATCH kind=%u gate=%g slices=%d U=%d: ref (%a, %d, pass %d) ours (%a, %d, pass %d)\n", kind, gate, slices, U, max, rc, ref_pass, P, cls, pass);
            }
        }
    }
    printf("cases=%ld passing_gate=%ld collision_replays=%ld mismatches=%ld\n", total, passed, replays, bad);
    return bad != 0;
}
