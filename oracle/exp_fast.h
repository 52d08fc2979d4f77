/* exp_fast.h -- correctly rounded expf for x in (-80, 0] WITHOUT double precision on the accepted path
 * (test infrastructure: the CPU statement of the arithmetic in unimoe_audio_b200/csrc/exp_fast.cuh; see that file).
 *
 * exp(x) = 2^k * 2^(i/32) * exp(r),  n = 32 k + i = rint(x * 32/ln2),  r = x - n ln2/32  (|r| <= ln2/64):
 *   - r as a float pair (three-part ln2: the first product is exact, the second goes through an exact two-product);
 *   - exp(r) = 1 + r + r^2 (1/2 + g), g = r (1/6 + r (1/24 + r (1/120 + r/720))), the leading terms carried as pairs;
 *   - times the pair table of 2^(i/32);  the result pair (yh, yl) approximates exp(x) 2^-k to ~2^-43;
 *   - Ziv's test: yh is the correctly rounded value when the tail yl stays 2^-14 ulp clear of a rounding boundary;
 *     otherwise *fallback = 1 and the caller evaluates (float)exp((double)x).
 * tools/verify_exp_fast.c runs ALL ~1.1e9 floats of the domain against the long-double exp: no accepted value differs.
 * Every operation is one IEEE binary32 operation (fmaf where written): CPU and CUDA agree bit for bit. */
#ifndef DCMOE_EXP_FAST_H_
#define DCMOE_EXP_FAST_H_
#include <math.h>
#include <stdint.h>
#include <string.h>

static const float kExpTabHi[32] = {0x1.0000000000000p+0f, 0x1.059b0e0000000p+0f, 0x1.0b55860000000p+0f, 0x1.11301e0000000p+0f, 0x1.172b840000000p+0f, 0x1.1d48740000000p+0f, 0x1.2387a60000000p+0f, 0x1.29e9e00000000p+0f, 0x1.306fe00000000p+0f, 0x1.371a740000000p+0f, 0x1.3dea640000000p+0f, 0x1.44e0860000000p+0f, 0x1.4bfdae0000000p+0f, 0x1.5342b60000000p+0f, 0x1.5ab07e0000000p+0f, 0x1.6247ec0000000p+0f, 0x1.6a09e60000000p+0f, 0x1.71f75e0000000p+0f, 0x1.7a11480000000p+0f, 0x1.82589a0000000p+0f, 0x1.8ace540000000p+0f, 0x1.93737c0000000p+0f, 0x1.9c49180000000p+0f, 0x1.a5503c0000000p+0f, 0x1.ae89fa0000000p+0f, 0x1.b7f7700000000p+0f, 0x1.c199be0000000p+0f, 0x1.cb720e0000000p+0f, 0x1.d5818e0000000p+0f, 0x1.dfc9740000000p+0f, 0x1.ea4afa0000000p+0f, 0x1.f507660000000p+0f};
static const float kExpTabLo[32] = {0x0.0p+0f, -0x1.9d4f520000000p-25f, 0x1.9f31220000000p-25f, -0x1.fdb4960000000p-25f, -0x1.c157420000000p-27f, -0x1.d2e8ca0000000p-25f, 0x1.ceac480000000p-25f, -0x1.5c04240000000p-25f, 0x1.4636e20000000p-25f, -0x1.18aac60000000p-25f, 0x1.8246840000000p-25f, 0x1.8624b40000000p-30f, -0x1.593abc0000000p-25f, -0x1.2c56100000000p-25f, -0x1.5bd5ec0000000p-27f, -0x1.f8b5500000000p-25f, 0x1.9fcef40000000p-26f, 0x1.1d8bee0000000p-25f, -0x1.829fd00000000p-25f, -0x1.accc7c0000000p-26f, 0x1.15506e0000000p-27f, -0x1.e647440000000p-25f, 0x1.51f8480000000p-27f, -0x1.b83b540000000p-25f, -0x1.a94b140000000p-26f, -0x1.a094380000000p-25f, -0x1.3d56b20000000p-27f, -0x1.8837cc0000000p-27f, -0x1.822dbc0000000p-27f, -0x1.908c940000000p-25f, 0x1.52486c0000000p-27f, -0x1.246eb00000000p-26f};

/* returns the candidate; yl_out (optional) receives the tail for accuracy statistics */
static inline float dcmoe_exp_fast(float x, int* fallback, float* yh_out, float* yl_out, int* k_out) {
    *fallback = 0;
    if (x == 0.0f) { if (yh_out) { *yh_out = 1.0f; *yl_out = 0.0f; *k_out = 0; } return 1.0f; }
    if (!(x > -80.0f) || x > 0.0f) { *fallback = 1; return 0.0f; }       /* also NaN */
    const float nf = rintf(x * 0x1.715476p+5f);                           /* 32 / ln 2 */
    const int n = (int)nf;
    const float r0 = fmaf(nf, -0x1.63p-6f, x);                            /* exact: 355/16384 has 9 significant bits */
    const float A2 = 0x1.bd0106p-18f, A3 = -0x1.cf79acp-45f;             /* ln2/32 = 355/16384 - A2 - A3 */
    const float p = nf * A2;
    const float pe = fmaf(nf, A2, -p);
    const float rh = r0 + p;
    const float bb = rh - r0;
    const float se = (r0 - (rh - bb)) + (p - bb);
    const float rl = (se + pe) + nf * A3;
    /* exp(r) */
    float g = fmaf(rh, 0x1.6c16c2p-10f, 0x1.111112p-7f);
    g = fmaf(rh, g, 0x1.555556p-5f);
    g = fmaf(rh, g, 0x1.555556p-3f);
    g = rh * g;
    const float p2 = rh * rh;
    const float p2e = fmaf(rh, rh, -p2) + 2.0f * rh * rl;
    const float a = 1.0f + rh;
    const float ae = rh - (a - 1.0f);
    const float h2 = 0.5f * p2;
    const float b = a + h2;
    const float be = h2 - (b - a);
    const float lo = ((ae + be) + rl) + fmaf(p2, g, 0.5f * p2e);
    /* times 2^(i/32) */
    const int i = n & 31, k = n >> 5;
    const float th = kExpTabHi[i], tl = kExpTabLo[i];
    const float m = b * th;
    const float me = fmaf(b, th, -m);
    const float ylo = me + fmaf(b, tl, lo * th);
    const float yh = m + ylo;
    const float yl = ylo - (yh - m);
    uint32_t u;
    memcpy(&u, &yh, 4);
    uint32_t ue = (u & 0x7f800000u) - (23u << 23);
    if ((u & 0x007fffffu) == 0u && yl < 0.0f) ue -= 1u << 23;             /* below a power of two the spacing halves */
    float ulp;
    memcpy(&ulp, &ue, 4);
    if (!(fabsf(yl) < ulp * 0x1.fff8p-2f)) { *fallback = 1; return 0.0f; }  /* (1/2 - 2^-14) ulp */
    if (yh_out) { *yh_out = yh; *yl_out = yl; *k_out = k; }
    return ldexpf(yh, k);                                                  /* exact: the result is a normal number */
}
#endif
