/* Exhaustive check of oracle/exp_fast.h: every float x in (-80, 0] against the long-double exp.
 *   gcc -O2 -ffp-contract=off -fopenmp -o tools/_build/verify_exp_fast tools/verify_exp_fast.c -lm && tools/_build/verify_exp_fast
 * Reports: accepted values that differ from the correctly rounded result (must be 0), the fallback rate, the largest
 * observed error of the (yh, yl) pair in units of 2^-40 relative, and the inputs whose long-double reference itself
 * lies within 2^-58 of a rounding boundary (reference not decisive there). */
#include <stdio.h>
#include <stdlib.h>
#include "../oracle/exp_fast.h"

int main(int argc, char** argv) {
    float lim = -80.0f;
    uint32_t hi_bits; memcpy(&hi_bits, &lim, 4);           /* bits of -80 */
    const uint32_t first = 0x80000001u;                    /* smallest negative denormal */
    uint64_t wrong = 0, fallbacks = 0, total = 0, undecided = 0, dbl_differs = 0;
    double max_err = 0.0;
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
#pragma omp parallel for reduction(+:wrong,fallbacks,total,undecided,dbl_differs) reduction(max:max_err) schedule(static, 1 << 16)
    for (uint64_t ub = first; ub < hi_bits; ub += stride) {
        uint32_t u = (uint32_t)ub;
        float x; memcpy(&x, &u, 4);
        int fb, k; float yh, yl;
        const float got = dcmoe_exp_fast(x, &fb, &yh, &yl, &k);
        ++total;
        const long double t = expl((long double)x);
        const float want = (float)t;                       /* RN to binary32 (single rounding from 64-bit significand) */
        if ((float)exp((double)x) != want) ++dbl_differs;  /* the fallback / previous definition: double exp, then one rounding */
        if (fb) { ++fallbacks; continue; }
        /* is the reference decisive?  distance of t from the midpoint between want and its neighbour towards t */
        const float nb = t > (long double)want ? nextafterf(want, INFINITY) : nextafterf(want, -INFINITY);
        const long double mid = ((long double)want + (long double)nb) / 2;
        const long double rel = fabsl(t - mid) / t;
        if (rel < 0x1p-58L) ++undecided;
        if (got != want) { ++wrong; if (wrong < 10) printf("WRONG x=%a got=%a want=%a\n", x, got, want); }
        const long double approx = ldexpl((long double)yh + (long double)yl, k);
        const double err = (double)(fabsl(approx - t) / t * 0x1p40L);
        if (err > max_err) max_err = err;
    }
    printf("inputs %llu  accepted-but-wrong %llu  fallbacks %llu (%.3e)  reference-undecided %llu  max pair error %.4f x 2^-40  "
           "(float)exp((double)x) != correctly rounded: %llu\n",
           (unsigned long long)total, (unsigned long long)wrong, (unsigned long long)fallbacks, (double)fallbacks / (double)total,
           (unsigned long long)undecided, max_err, (unsigned long long)dbl_differs);
    return wrong ? 1 : 0;
}
