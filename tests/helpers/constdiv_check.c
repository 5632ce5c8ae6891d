/* Checks the constant-divisor sequence used by the tile kernel's fast t-statistic (dnascent_b200/csrc/seg.cu,
 * div_const): q0 = x*r, rem = fma(-q0, w, x), q = fma(rem, r, q0) with r = RN(1/w) must equal the IEEE quotient x/w.
 * usage: constdiv_check <float_stride> <n_double>   -> prints "<float mismatches> <double mismatches>"
 * float: every finite float with |x| >= 2^-124 whose bit pattern is a multiple of <float_stride> (1 = exhaustive),
 * double: n random doubles with exponents in [-60, 60] plus structured cases next to multiples of w. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static uint64_t rng(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
int main(int argc, char **argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 64;
    const long long nd = argc > 2 ? atoll(argv[2]) : 1000000;
    unsigned long long badf = 0, badd = 0;
    for (int w = 2; w <= 7; w++) {
        const float wf = (float)w, rf = 1.0f / wf;
        for (uint64_t b = 0; b < (1ull << 32); b += stride) {
            uint32_t u = (uint32_t)b; float x; memcpy(&x, &u, 4);
            if (!isfinite(x) || fabsf(x) < 0x1p-124f) continue;
            const float q0 = x * rf, rem = fmaf(-q0, wf, x), q = fmaf(rem, rf, q0), a = x / wf;
            if (memcmp(&a, &q, 4)) badf++;
        }
        const double wd = w, rd = 1.0 / wd;
        if (fabs(fma(-wd, rd, 1.0)) > 0x1p-54) badd += 1000000;   /* the host-side admission test must pass for 2..7 */
        uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)w;
        for (long long i = 0; i < nd; i++) {
            uint64_t m = rng(&s) & 0xFFFFFFFFFFFFFull; int ex = 1023 + (int)(rng(&s) % 121) - 60;
            uint64_t ub = ((uint64_t)ex << 52) | m; if (rng(&s) & 1) ub |= 1ull << 63;
            double x; memcpy(&x, &ub, 8);
            if ((i & 3) == 0) { uint64_t km = (rng(&s) & 0xFFFFFFFFFFFFFull) | (1ull << 52); x = (double)km * wd;
                                uint64_t ux; memcpy(&ux, &x, 8); ux += (rng(&s) % 5) - 2; memcpy(&x, &ux, 8); }
            const double q0 = x * rd, rem = fma(-q0, wd, x), q = fma(rem, rd, q0), a = x / wd;
            if (memcmp(&a, &q, 8)) badd++;
        }
    }
    printf("%llu %llu\n", badf, badd);
    return 0;
}
