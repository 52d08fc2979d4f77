// route_token.cuh -- the per-token routing arithmetic of the DCMoE router (sm_100a), shared by the router kernels
// (router.cu: the persistent router, the one-CTA-per-block router, the decode front end).
// Everything follows the canonical arithmetic of oracle/route_oracle.c: explicit __f*_rn intrinsics, never contracted.
#pragma once

#include "common.cuh"
#include "exp_fast.cuh"

namespace dcmoe {
namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float pow2if(int q) { return __int_as_float((q + 127) << 23); }

// Sleef_expf_u10 restated with single-rounding intrinsics (see oracle/route_oracle.c)
__device__ __forceinline__ float exp_sleef_u10(float d) {
    float qf = rintf(__fmul_rn(d, 1.442695040888963407359924681001892137426645954152985934135449406931f));
    int q = (int)qf;
    float s = __fmaf_rn(qf, -0.693145751953125f, d);
    s = __fmaf_rn(qf, -1.428606765330187045e-06f, s);
    float u = 0.000198527617612853646278381f;
    u = __fmaf_rn(u, s, 0.00139304355252534151077271f);
    u = __fmaf_rn(u, s, 0.00833336077630519866943359f);
    u = __fmaf_rn(u, s, 0.0416664853692054748535156f);
    u = __fmaf_rn(u, s, 0.166666671633720397949219f);
    u = __fmaf_rn(u, s, 0.5f);
    u = __fadd_rn(1.0f, __fmaf_rn(__fmul_rn(s, s), u, s));
    u = __fmul_rn(__fmul_rn(u, pow2if(q >> 1)), pow2if(q - (q >> 1)));
    if (d < -104.0f) u = 0.0f;
    if (d > 100.0f) u = __int_as_float(0x7f800000);
    return u;
}

// exp_cr(x): correctly rounded expf (the bf16 path of ATen's softmax uses std::exp) -- exp_fast.cuh

template <bool BF16>
__device__ __forceinline__ float rnd(float v) {
    return BF16 ? bf16_round(v) : v;
}

// softmax over lanes [0, n) of a 16-lane group; lanes >= n must hold -inf.  Sequential sum in
// lane order, multiply by the reciprocal, round to D -- the ATen CPU order.
template <bool BF16>
__device__ __forceinline__ float softmax_lanes(float v, int j, const int n) {
    float m = v;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, off, 16));
    float e = 0.0f;
    if (j < n) e = BF16 ? exp_cr(__fsub_rn(v, m)) : exp_sleef_u10(__fsub_rn(v, m));
    float s = __shfl_sync(kFull, e, 0, 16);
#pragma unroll
    for (int i = 1; i < kMaxDyn; ++i) {
        if (i < n) s = __fadd_rn(s, __shfl_sync(kFull, e, i, 16));   // n is warp-uniform
    }
    float inv = __fdiv_rn(1.0f, s);
    return rnd<BF16>(__fmul_rn(e, inv));
}

// torch.sum over an inner dim of length n <= 16: ATen row_sum with 8 interleaved partial sums
__device__ __forceinline__ float row_sum8_lanes(float d, int n) {
    const int n8 = n >> 3;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc[k] = 0.0f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i < n8) acc[k] = __fadd_rn(acc[k], __shfl_sync(kFull, d, 8 * i + k, 16));
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxDyn; ++i) {
        if (i >= 8 * n8 && i < n) acc[0] = __fadd_rn(acc[0], __shfl_sync(kFull, d, i, 16));
    }
#pragma unroll
    for (int k = 1; k < 8; ++k) acc[0] = __fadd_rn(acc[0], acc[k]);
    return acc[0];
}

struct RouteConsts {
    float thr_p, thr_eps, plus_eps, finfo_min;
    int n_dyn, E;
    int fixed_k;          // > 0: mlp_dynamic_top_p == 0, every token selects fixed_k dynamic experts (core.py:256-257)
    int always_softmax;   // debug: evaluate the mixer softmax even when it is provably 1 (DCMOE_ROUTER_ALWAYS_SOFTMAX=1)
    unsigned long long* dbg;   // tuning (DCMOE_ROUTER_DEBUG=1): per-CTA cycle counters of router_tma_kernel, else nullptr
};

// exp of the canonical arithmetic for dtype D
template <bool BF16>
__device__ __forceinline__ float exp_D(float x) {
    return BF16 ? exp_cr(x) : exp_sleef_u10(x);
}

// sequential sum of lanes [0, n) of a 16-lane group, in lane order (ATen softmax order)
__device__ __forceinline__ float seq_sum_lanes(float e, const int n) {
    float s = __shfl_sync(kFull, e, 0, 16);
#pragma unroll
    for (int i = 1; i < kMaxDyn; ++i) {
        if (i < n) s = __fadd_rn(s, __shfl_sync(kFull, e, i, 16));
    }
    return s;
}

// descending rank of v among lanes [0, n) of the 16-lane group, ties broken by lower lane first
__device__ __forceinline__ int rank_desc_lanes(float v, int j, const int n) {
    int rank = 0;
#pragma unroll
    for (int i = 0; i < kMaxDyn; ++i) {
        if (i < n) {
            const float vi = __shfl_sync(kFull, v, i, 16);
            rank += (vi > v) || (vi == v && i < j);
        }
    }
    return rank;
}

// inverse of a permutation held one entry per lane: lane r gets the lane whose rank is r
__device__ __forceinline__ int inverse_perm_lanes(int rank, int j, const int n) {
    int src = 0;
#pragma unroll
    for (int i = 0; i < kMaxDyn; ++i) {
        if (i < n) {
            const int ri = __shfl_sync(kFull, rank, i, 16);
            if (ri == j) src = i;
        }
    }
    return src;
}

// Route one token per 16-lane group.  l = logit of lane j (D-representable fp32), am = padding mask.
// NDYN / NE > 0 fix the expert counts at compile time (the reference config: 9 dynamic + 2 shared), which
// trims every shuffle loop to its real length; 0 = read them from rc.
//
// The reference's iterative arg-max loop (core.py:103-147) selects experts in the order of the logits sorted
// descending with ties to the lower index, so one rank computation replaces the k arg-max reductions; the
// softmax inside iteration `it` has max == the it-th largest logit, e = 1 for it and exactly 0 for every
// dropped / already selected entry, i.e. it is exactly 1 unless another remaining logit lies within 2 % of
// the current maximum ("near tie").  Only near ties (about 3 % of iterations) evaluate it.
//
// DROP: the token-drop branch (core.py:314-316, :328-329) is compiled in; when `do_drop` (warp-uniform) is set, a dynamic
// column survives only where keep_j != 0, the dropped weights are zeroed and the weights are normalised a second time.
// The aux softmax (ga_out) is always that of the mask BEFORE the drop (core.py:293 precedes :302).
template <bool BF16, int NDYN, int NE, bool DROP = false>
__device__ __forceinline__ void route_token(float l, int j, int half, int am, const RouteConsts& rc, int& raw_out,
                                            int& mask_out, float& gw_out, float& ga_out, int keep_j = 1,
                                            bool do_drop = false) {
    const int n_dyn = NDYN ? NDYN : rc.n_dyn, E = NE ? NE : rc.E;
    const float ninf = __int_as_float(0xff800000);
    const bool dyn = j < n_dyn;
    const unsigned half_mask = 0xffffu << (half * 16);
    // ---- selection order: logits descending, ties -> lower index (torch.max first occurrence) ----
    const int rank_l = rank_desc_lanes(dyn ? l : ninf, j, n_dyn);
    const int src_l = inverse_perm_lanes(rank_l, j, n_dyn);          // lane r: index of the r-th largest logit
    const float sorted_l = __shfl_sync(kFull, l, src_l, 16);         // lane r: r-th largest logit
    const float top1 = __shfl_sync(kFull, sorted_l, 0, 16);
    // ---- Top-P count (core.py:162-166) ----
    // e_j = exp(l_j - top1) is evaluated once for every lane j < E: the aux softmax (max = top1 whenever an
    // expert is selected) and, when no shared logit exceeds top1, the 11-way softmax reuse the same values.
    const float e_all = (j < E) ? exp_D<BF16>(__fsub_rn(l, top1)) : 0.0f;
    const float e = dyn ? e_all : 0.0f;
    float inv = __fdiv_rn(1.0f, seq_sum_lanes(e, n_dyn));
    const float p = rnd<BF16>(__fmul_rn(e, inv));
    // sorted probabilities: a correctly rounded exp is monotone, so in bf16 the order of p is the order of l
    // (ties give equal values, and only the sorted VALUES matter); fp32 ranks p itself.
    int src_p = src_l;
    if (!BF16) src_p = inverse_perm_lanes(rank_desc_lanes(dyn ? p : ninf, j, n_dyn), j, n_dyn);
    const float sorted_p = __shfl_sync(kFull, p, src_p, 16);         // lane r: r-th largest probability
    float run = 0.0f;                                                // lane r ends with prefix c_r
#pragma unroll
    for (int i = 0; i < kMaxDyn; ++i) {
        if (i < n_dyn) {
            const float v = __shfl_sync(kFull, sorted_p, i, 16);
            if (i <= j) run = __fadd_rn(run, v);
        }
    }
    const bool below = dyn && !(rnd<BF16>(run) >= rc.thr_p);
    const int raw = rc.fixed_k > 0 ? rc.fixed_k : 1 + __popc(__ballot_sync(kFull, below) & half_mask);
    raw_out = raw;
    const int k = raw <= n_dyn ? raw : 0;
    const int kmax = max(k, __shfl_xor_sync(kFull, k, 16));
    // ---- mixer (core.py:103-147, eval branch) ----
    float rw = (dyn && rank_l < k) ? 1.0f : 0.0f;
    bool had_tie = false;
    for (int it = 0; it < kmax; ++it) {
        const float thr = __shfl_sync(kFull, sorted_l, it, 16);
        const float fac = fmaxf(fabsf(l), fabsf(thr));
        const float diff = rnd<BF16>(__fsub_rn(thr, l));
        // drop <=> rnd(diff / fac) > t, t = 2*eps.  Outside [0.75 t, 1.5 t] * fac the outcome is certain (the
        // rounding of the quotient moves it by < 0.4 %), so the IEEE division only runs for borderline lanes.
        bool drop = diff > rc.thr_eps * fac;
        const bool borderline = !(diff < 0.75f * rc.thr_eps * fac) && !(diff > 1.5f * rc.thr_eps * fac);  // also NaN, fac == 0
        if (__any_sync(kFull, borderline && dyn)) {
            const float ratio = rnd<BF16>(__fdiv_rn(diff, fac));
            if (borderline) drop = ratio > rc.thr_eps;
        }
        const bool remaining = dyn && rank_l >= it;
        const bool near_tie = remaining && rank_l != it && !drop && it < k;
        const unsigned ties = __ballot_sync(kFull, near_tie);
        if (ties != 0u || rc.always_softmax) {                        // warp-uniform
            const float sm = softmax_lanes<BF16>((remaining && !drop) ? l : ninf, j, n_dyn);
            if (it < k && rank_l == it) rw = sm;
            had_tie |= (ties & half_mask) != 0u;
        }
    }
    const int sel = (dyn && rank_l < k) ? 1 : 0;
    // ---- normalise (core.py:284) ----
    float rsum = (float)k;                                            // k exact ones when no near tie occurred
    if (__any_sync(kFull, had_tie) || rc.always_softmax) {
        const float full = row_sum8_lanes(dyn ? rw : 0.0f, n_dyn);
        if (had_tie || rc.always_softmax) rsum = full;
    }
    const float den = rnd<BF16>(__fadd_rn(rnd<BF16>(rsum), rc.plus_eps));
    rw = rnd<BF16>(__fdiv_rn(rw, den));
    // ---- padding mask, shared experts always on (core.py:286-291) ----
    int mk = dyn ? sel * am : (j < E ? 1 : 0);
    bool any_sel = (k > 0) && (am != 0);
    // ---- aux-loss softmax (core.py:370-373): selected logits, finfo.min elsewhere ----
    {
        // max = top1 if anything is selected (exp(finfo.min - top1) is exactly 0), else every entry is
        // finfo.min and exp(0) = 1
        const float ea = dyn ? (any_sel ? (mk ? e_all : 0.0f) : 1.0f) : 0.0f;
        const float ia = __fdiv_rn(1.0f, seq_sum_lanes(ea, n_dyn));
        ga_out = rnd<BF16>(__fmul_rn(ea, ia));
    }
    if (DROP && do_drop) {
        // ---- token drop (core.py:314-316, :326-329): AND with the capacity mask, zero the dropped weights,
        // normalise again (torch.sum -> ATen row_sum; the same rounding points as the first normalisation) ----
        if (dyn && !keep_j) mk = 0;
        if (dyn && !mk) rw = 0.0f;                                    // also where the padding mask cleared the column
        const float rs2 = row_sum8_lanes(dyn ? rw : 0.0f, n_dyn);
        const float den2 = rnd<BF16>(__fadd_rn(rnd<BF16>(rs2), rc.plus_eps));
        rw = rnd<BF16>(__fdiv_rn(rw, den2));
        any_sel = (__ballot_sync(kFull, dyn && mk) & half_mask) != 0u;
    }
    // ---- global weights (core.py:188-192): 11-way softmax over selected + shared ----
    {
        float ms = ninf;                                              // max of the shared logits
#pragma unroll
        for (int i = 0; i < kMaxDyn; ++i) {
            if (i >= n_dyn && i < E) ms = fmaxf(ms, __shfl_sync(kFull, l, i, 16));
        }
        // softmax max == top1: same exp arguments (after a token drop the largest surviving logit can lie below top1)
        const bool reuse = any_sel && top1 >= ms && !(DROP && do_drop);
        float eg = (j < E && mk) ? e_all : 0.0f;
        if (!__all_sync(kFull, reuse)) {                              // warp-uniform
            float m = any_sel ? fmaxf(top1, ms) : ms;
            if (DROP && do_drop) {                                     // max over the surviving columns
                float mm = (j < E && mk) ? l : ninf;
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) mm = fmaxf(mm, __shfl_xor_sync(kFull, mm, off, 16));
                m = mm;
            }
            const float eg2 = (j < E && mk) ? exp_D<BF16>(__fsub_rn(l, m)) : 0.0f;
            if (!reuse) eg = eg2;
        }
        const float ig = __fdiv_rn(1.0f, seq_sum_lanes(eg, E));
        const float G = rnd<BF16>(__fmul_rn(eg, ig));
        const float dsum = rnd<BF16>(row_sum8_lanes(dyn ? G : 0.0f, n_dyn));
        gw_out = dyn ? rnd<BF16>(__fmul_rn(rw, dsum)) : G;
    }
    mask_out = mk;
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Partial gate logits of ONE K slice of a 16-token block on mma.sync: rows = tokens [blk*16, blk*16+16), K range
// [ks*H/ksplit, (ks+1)*H/ksplit), columns = the E <= 16 router outputs; r[token][expert] receives the fp32 partial sums.
// The decode front end (front_small_kernel) sums the ksplit partials of a logit in slice order.
__device__ __forceinline__ void gate_slice16(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wg, int T, int H,
                                             int E, int blk, int ks, int ksplit, int lane, float (*r)[16]) {
    const int Kq = H / ksplit, k0 = ks * Kq;
    const int g = lane >> 2, tq = lane & 3;
    const int r0 = blk * kRouterBlock + g, r1 = r0 + 8;
    const bool v0 = r0 < T, v1 = r1 < T;
    const __nv_bfloat16* xr0 = x + (int64_t)(v0 ? r0 : 0) * H + k0 + tq * 8;
    const __nv_bfloat16* xr1 = x + (int64_t)(v1 ? r1 : 0) * H + k0 + tq * 8;
    const bool wv0 = g < E, wv1 = g + 8 < E;
    const __nv_bfloat16* w0 = wg + (int64_t)(wv0 ? g : 0) * H + k0 + tq * 8;
    const __nv_bfloat16* w1 = wg + (int64_t)(wv1 ? g + 8 : 0) * H + k0 + tq * 8;
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const int steps = Kq >> 5;
    for (int s0 = 0; s0 < steps; s0 += 4) {
        uint4 a[4], b[4], q0[4], q1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] = (v0 && s0 + u < steps) ? ld_nc_v4(xr0 + (s0 + u) * 32) : zero;   // zero past this slice's K range
            b[u] = (v1 && s0 + u < steps) ? ld_nc_v4(xr1 + (s0 + u) * 32) : zero;
            q0[u] = (wv0 && s0 + u < steps) ? ld_ca_v4(w0 + (s0 + u) * 32) : zero;
            q1[u] = (wv1 && s0 + u < steps) ? ld_ca_v4(w1 + (s0 + u) * 32) : zero;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            mma_bf16_16816(c0, a[u].x, b[u].x, a[u].y, b[u].y, q0[u].x, q0[u].y);
            mma_bf16_16816(c0, a[u].z, b[u].z, a[u].w, b[u].w, q0[u].z, q0[u].w);
            mma_bf16_16816(c1, a[u].x, b[u].x, a[u].y, b[u].y, q1[u].x, q1[u].y);
            mma_bf16_16816(c1, a[u].z, b[u].z, a[u].w, b[u].w, q1[u].z, q1[u].w);
        }
    }
    r[g][2 * tq] = c0[0];
    r[g][2 * tq + 1] = c0[1];
    r[g + 8][2 * tq] = c0[2];
    r[g + 8][2 * tq + 1] = c0[3];
    r[g][8 + 2 * tq] = c1[0];
    r[g][8 + 2 * tq + 1] = c1[1];
    r[g + 8][8 + 2 * tq] = c1[2];
    r[g + 8][8 + 2 * tq + 1] = c1[3];
}

// scalar operands rounded to D exactly as torch does for a wrapped python scalar (host)
inline RouteConsts make_route_consts(const dcmoe_config* cfg, bool bf16) {
    RouteConsts rc;
    rc.n_dyn = cfg->n_real + cfg->n_null;
    rc.E = rc.n_dyn + cfg->n_fix;
    auto r = [&](float v) { return bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v; };
    rc.thr_p = r((float)cfg->top_p);
    rc.thr_eps = r((float)(2.0 * cfg->jitter_eps));
    rc.plus_eps = r(1e-6f);
    rc.finfo_min = bf16 ? -3.3895313892515355e38f : -3.4028234663852886e38f;
    rc.always_softmax = 0;
    rc.dbg = nullptr;
    rc.fixed_k = cfg->top_p == 0.0 ? cfg->fixed_top_k : 0;
    return rc;
}

}  // namespace
}  // namespace dcmoe
