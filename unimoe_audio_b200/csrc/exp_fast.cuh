// exp_fast.cuh -- correctly rounded expf for the router's softmaxes without double precision on the hot path.
//
// ATen's bf16 softmax evaluates std::exp in fp32-from-double quality, i.e. the CORRECTLY ROUNDED binary32 exponential
// (reference utils/UniMoE_Audio_core.py:162, :373, :188 through torch.softmax).  The router used
// (float)exp((double)x) for that; B200's vector FP64 pipe issues at a small fraction of the FP32 rate and those ~30 DFMA per
// exponential were the router's critical path (profiles/r02_router_history.md).  This version keeps the same VALUES:
//
//   exp(x) = 2^k * 2^(i/32) * exp(r),  n = 32 k + i = rint(x * 32/ln2),  r = x - n ln2/32,  |r| <= ln2/64
//     r      as a float pair: three-part ln2/32 (355/16384 exactly representable in 9 bits -> the first product is exact;
//            the second goes through an exact two-product)
//     exp(r) = 1 + r + r^2 (1/2 + g),  g = r (1/6 + r (1/24 + r (1/120 + r/720))): leading terms carried as float pairs
//     2^(i/32) from a 32-entry pair table
//   The pair (yh, yl) approximates exp(x) 2^-k to 2^-43 (largest error over ALL 1,117,782,015 floats of (-80, 0):
//   0.13 x 2^-40).  Ziv's rounding test: yh is the correctly rounded result when yl stays 2^-14 ulp clear of a rounding
//   boundary; otherwise (1.5e-5 of all inputs) and outside (-80, 0] the double-precision evaluation runs.
// tools/verify_exp_fast.c proves on the CPU statement of the same arithmetic (oracle/exp_fast.h; every step one IEEE
// binary32 operation, so host and device agree bit for bit) that no accepted value differs from the correctly rounded
// exponential, and that (float)exp((double)x) equals it on the whole domain: the router's outputs do not change.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dcmoe {

static __device__ const float2 kExpTab32[32] = {   // 2^(i/32) = .x + .y
    {0x1.000000p+0f, 0x0.0p+0f}, {0x1.059b0ep+0f, -0x1.9d4f52p-25f}, {0x1.0b5586p+0f, 0x1.9f3122p-25f}, {0x1.11301ep+0f, -0x1.fdb496p-25f},
    {0x1.172b84p+0f, -0x1.c15742p-27f}, {0x1.1d4874p+0f, -0x1.d2e8cap-25f}, {0x1.2387a6p+0f, 0x1.ceac48p-25f}, {0x1.29e9e0p+0f, -0x1.5c0424p-25f},
    {0x1.306fe0p+0f, 0x1.4636e2p-25f}, {0x1.371a74p+0f, -0x1.18aac6p-25f}, {0x1.3dea64p+0f, 0x1.824684p-25f}, {0x1.44e086p+0f, 0x1.8624b4p-30f},
    {0x1.4bfdaep+0f, -0x1.593abcp-25f}, {0x1.5342b6p+0f, -0x1.2c5610p-25f}, {0x1.5ab07ep+0f, -0x1.5bd5ecp-27f}, {0x1.6247ecp+0f, -0x1.f8b550p-25f},
    {0x1.6a09e6p+0f, 0x1.9fcef4p-26f}, {0x1.71f75ep+0f, 0x1.1d8beep-25f}, {0x1.7a1148p+0f, -0x1.829fd0p-25f}, {0x1.82589ap+0f, -0x1.accc7cp-26f},
    {0x1.8ace54p+0f, 0x1.15506ep-27f}, {0x1.93737cp+0f, -0x1.e64744p-25f}, {0x1.9c4918p+0f, 0x1.51f848p-27f}, {0x1.a5503cp+0f, -0x1.b83b54p-25f},
    {0x1.ae89fap+0f, -0x1.a94b14p-26f}, {0x1.b7f770p+0f, -0x1.a09438p-25f}, {0x1.c199bep+0f, -0x1.3d56b2p-27f}, {0x1.cb720ep+0f, -0x1.8837ccp-27f},
    {0x1.d5818ep+0f, -0x1.822dbcp-27f}, {0x1.dfc974p+0f, -0x1.908c94p-25f}, {0x1.ea4afap+0f, 0x1.52486cp-27f}, {0x1.f50766p+0f, -0x1.246eb0p-26f}};

// the previous definition, kept as the rare path (and for |x| >= 80, NaN, x > 0)
static __device__ __noinline__ float exp_cr_double(float x) { return (float)exp((double)x); }

static __device__ __forceinline__ float exp_cr(float x) {
    if (x == 0.0f) return 1.0f;
    if (x > -80.0f && x < 0.0f) {
        const float nf = rintf(__fmul_rn(x, 0x1.715476p+5f));                 // 32 / ln 2
        const int n = (int)nf;
        const float r0 = __fmaf_rn(nf, -0x1.63p-6f, x);                       // exact
        const float A2 = 0x1.bd0106p-18f, A3 = -0x1.cf79acp-45f;              // ln2/32 = 355/16384 - A2 - A3
        const float p = __fmul_rn(nf, A2);
        const float pe = __fmaf_rn(nf, A2, -p);
        const float rh = __fadd_rn(r0, p);
        const float bb = __fsub_rn(rh, r0);
        const float se = __fadd_rn(__fsub_rn(r0, __fsub_rn(rh, bb)), __fsub_rn(p, bb));
        const float rl = __fadd_rn(__fadd_rn(se, pe), __fmul_rn(nf, A3));
        float g = __fmaf_rn(rh, 0x1.6c16c2p-10f, 0x1.111112p-7f);
        g = __fmaf_rn(rh, g, 0x1.555556p-5f);
        g = __fmaf_rn(rh, g, 0x1.555556p-3f);
        g = __fmul_rn(rh, g);
        const float p2 = __fmul_rn(rh, rh);
        const float p2e = __fadd_rn(__fmaf_rn(rh, rh, -p2), __fmul_rn(__fmul_rn(2.0f, rh), rl));
        const float a = __fadd_rn(1.0f, rh);
        const float ae = __fsub_rn(rh, __fsub_rn(a, 1.0f));
        const float h2 = __fmul_rn(0.5f, p2);
        const float b = __fadd_rn(a, h2);
        const float be = __fsub_rn(h2, __fsub_rn(b, a));
        const float lo = __fadd_rn(__fadd_rn(__fadd_rn(ae, be), rl), __fmaf_rn(p2, g, __fmul_rn(0.5f, p2e)));
        const int i = n & 31, k = n >> 5;
        const float2 t = __ldg(&kExpTab32[i]);
        const float m = __fmul_rn(b, t.x);
        const float me = __fmaf_rn(b, t.x, -m);
        const float ylo = __fadd_rn(me, __fmaf_rn(b, t.y, __fmul_rn(lo, t.x)));
        const float yh = __fadd_rn(m, ylo);
        const float yl = __fsub_rn(ylo, __fsub_rn(yh, m));
        const uint32_t u = __float_as_uint(yh);
        uint32_t ue = (u & 0x7f800000u) - (23u << 23);
        if ((u & 0x007fffffu) == 0u && yl < 0.0f) ue -= 1u << 23;             // below a power of two the spacing halves
        if (fabsf(yl) < __fmul_rn(__uint_as_float(ue), 0x1.fff8p-2f))         // (1/2 - 2^-14) ulp
            return __fmul_rn(yh, __uint_as_float((uint32_t)(k + 127) << 23)); // exact: k >= -116
    }
    return exp_cr_double(x);
}

}  // namespace dcmoe
