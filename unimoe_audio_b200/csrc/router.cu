// router.cu -- fused gate projection + Top-P router for the DCMoE layer (sm_100a).
//
// Replaces (reference utils/UniMoE_Audio_core.py):
//   :251        gate Linear            x[T,H] @ W_g^T -> logits[T,E]
//   :157-167    audio_dynamic_expert_selection   (softmax -> sort -> cumsum -> >= p -> count)
//   :262-282    the Python loop over top_k groups + audio_sparse_expert_mixer (:94-154, eval branch)
//   :284        normalisation, :286-291 padding mask / shared columns
//   :361-389    aux-loss per-token terms (block partial sums; finished in plan.cu)
//   :178-193    calculate_audio_global_routing_weight
//
// One CTA (4 warps) owns DCMOE_ROUTER_BLOCK = 16 tokens.
//   phase 1 (bf16): the skinny GEMM [16, H] x [H, 16] runs on mma.sync.m16n8k16 with the K
//     dimension split across the 4 warps.  x is streamed straight from HBM with 128-bit loads
//     (each row is read exactly once, full 32 B sectors); W_g (45 KB) stays L1/L2 resident.
//     The K order inside a 16-element MMA step is permuted identically for A and B so that each
//     lane's 16 B load feeds two MMA steps without any shuffle.  fp32 accumulators.
//     A CUDA-core FFMA version of this product needs 22.5 kFMA/token, i.e. about as long as the
//     HBM read of x itself; the tensor-core form makes the kernel purely HBM bound.
//   phase 1 (fp32): FFMA dot products (parity path, not a performance target).
//   phase 2: half-warp per token.  Lane j holds logit j; softmax / rank-sort / running sum /
//     arg-max with lowest-index tie-break are built from __shfl_sync / __ballot_sync.  All
//     floating point follows the canonical arithmetic of oracle/route_oracle.c (explicit
//     __f*_rn intrinsics, never contracted), so dynamic_top_k, expert_mask AND global_weight
//     are bit-identical to the oracle for identical logits.
#include <cstdlib>

#include "common.cuh"
#include "exp_fast.cuh"
#include "ptx.cuh"
#include "route_token.cuh"

namespace dcmoe {

namespace {

__device__ __forceinline__ float4 ld_bf16x4_as_float(const __nv_bfloat16* p) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
}

// MIXED (BF16 = false only): the fp32 gate of the training-mode forward (core.py:240-249) on a bf16 layer -- x and W_g
// are bf16 and are widened on load, logits and all routing arithmetic are fp32, global_weight is written in bf16
// (core.py:339).  keep [T, E] uint8 or nullptr: token_drop's capacity mask (dcmoe_drop_select).
template <bool BF16, int NDYN, int NE, bool MIXED = false>
__global__ void __launch_bounds__(128) router_kernel(const void* __restrict__ x_, const void* __restrict__ wg_,
                                                     const void* __restrict__ logits_in_,
                                                     const int32_t* __restrict__ attn_mask,
                                                     const uint8_t* __restrict__ keep, int64_t T, int H,
                                                     RouteConsts rc, void* __restrict__ logits_out_,
                                                     int64_t* __restrict__ top_k, int32_t* __restrict__ expert_mask,
                                                     void* __restrict__ gw_out_, int32_t* __restrict__ block_counts,
                                                     float* __restrict__ block_probs) {
    using elem_t = typename std::conditional<BF16, __nv_bfloat16, float>::type;
    __shared__ float red[4][kRouterBlock][16];
    __shared__ int s_cnt[kRouterBlock][kMaxDyn];
    __shared__ float s_prob[kRouterBlock][kMaxDyn];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t tok0 = (int64_t)blockIdx.x * kRouterBlock;
    const int E = NE ? NE : rc.E;
    const int n_dyn_k = NDYN ? NDYN : rc.n_dyn;

    if (logits_in_ == nullptr) {
        const int Kq = H >> 2;  // columns per warp
        const int k0 = warp * Kq;
        if constexpr (BF16) {
            const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_);
            const __nv_bfloat16* wg = static_cast<const __nv_bfloat16*>(wg_);
            const int g = lane >> 2, tq = lane & 3;
            const int64_t r0 = tok0 + g, r1 = r0 + 8;
            const bool v0 = r0 < T, v1 = r1 < T;
            const __nv_bfloat16* xr0 = x + (v0 ? r0 : 0) * (int64_t)H + k0 + tq * 8;
            const __nv_bfloat16* xr1 = x + (v1 ? r1 : 0) * (int64_t)H + k0 + tq * 8;
            const bool wv0 = g < E, wv1 = g + 8 < E;
            const __nv_bfloat16* w0 = wg + (int64_t)(wv0 ? g : 0) * H + k0 + tq * 8;
            const __nv_bfloat16* w1 = wg + (int64_t)(wv1 ? g + 8 : 0) * H + k0 + tq * 8;
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
            const int steps = Kq >> 5;  // 32-column steps
            for (int s0 = 0; s0 < steps; s0 += 4) {
                uint4 a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {  // batch the HBM loads: 8 x 16 B in flight per lane
                    a[u] = (v0 && s0 + u < steps) ? ld_nc_v4(xr0 + (s0 + u) * 32) : zero;   // zero past this warp's K range
                    b[u] = (v1 && s0 + u < steps) ? ld_nc_v4(xr1 + (s0 + u) * 32) : zero;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint4 q0 = (wv0 && s0 + u < steps) ? ld_ca_v4(w0 + (s0 + u) * 32) : zero;
                    uint4 q1 = (wv1 && s0 + u < steps) ? ld_ca_v4(w1 + (s0 + u) * 32) : zero;
                    mma_bf16_16816(c0, a[u].x, b[u].x, a[u].y, b[u].y, q0.x, q0.y);
                    mma_bf16_16816(c0, a[u].z, b[u].z, a[u].w, b[u].w, q0.z, q0.w);
                    mma_bf16_16816(c1, a[u].x, b[u].x, a[u].y, b[u].y, q1.x, q1.y);
                    mma_bf16_16816(c1, a[u].z, b[u].z, a[u].w, b[u].w, q1.z, q1.w);
                }
            }
            red[warp][g][2 * tq] = c0[0];
            red[warp][g][2 * tq + 1] = c0[1];
            red[warp][g + 8][2 * tq] = c0[2];
            red[warp][g + 8][2 * tq + 1] = c0[3];
            red[warp][g][8 + 2 * tq] = c1[0];
            red[warp][g][8 + 2 * tq + 1] = c1[1];
            red[warp][g + 8][8 + 2 * tq] = c1[2];
            red[warp][g + 8][8 + 2 * tq + 1] = c1[3];
        } else {
            // fp32 parity path: FFMA dot products, 2 tokens at a time, warp-reduced
            using in_t = typename std::conditional<MIXED, __nv_bfloat16, float>::type;
            const in_t* x = static_cast<const in_t*>(x_);
            const in_t* wg = static_cast<const in_t*>(wg_);
            auto ld4 = [](const in_t* p) -> float4 {
                if constexpr (MIXED) return ld_bf16x4_as_float(p);
                else return *reinterpret_cast<const float4*>(p);
            };
            for (int r = 0; r < kRouterBlock; r += 2) {
                const int64_t ra = tok0 + r, rb = ra + 1;
                const bool va = ra < T, vb = rb < T;
                float acc_a[kMaxDyn], acc_b[kMaxDyn];
#pragma unroll
                for (int e = 0; e < kMaxDyn; ++e) acc_a[e] = acc_b[e] = 0.f;
                for (int c = lane * 4; c < Kq; c += 128) {
                    float4 xa = va ? ld4(x + ra * (int64_t)H + k0 + c) : make_float4(0, 0, 0, 0);
                    float4 xb = vb ? ld4(x + rb * (int64_t)H + k0 + c) : make_float4(0, 0, 0, 0);
#pragma unroll
                    for (int e = 0; e < kMaxDyn; ++e) {
                        if (e < E) {
                            float4 w = ld4(wg + (int64_t)e * H + k0 + c);
                            acc_a[e] += xa.x * w.x + xa.y * w.y + xa.z * w.z + xa.w * w.w;
                            acc_b[e] += xb.x * w.x + xb.y * w.y + xb.z * w.z + xb.w * w.w;
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < kMaxDyn; ++e) {
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        acc_a[e] += __shfl_xor_sync(kFull, acc_a[e], off);
                        acc_b[e] += __shfl_xor_sync(kFull, acc_b[e], off);
                    }
                    if (lane == 0) {
                        red[warp][r][e] = acc_a[e];
                        red[warp][r + 1][e] = acc_b[e];
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- phase 2: half-warp per token ----
    const int half = lane >> 4, j = lane & 15;
    const elem_t* logits_in = static_cast<const elem_t*>(logits_in_);
    elem_t* logits_out = static_cast<elem_t*>(logits_out_);
    elem_t* gw_out = static_cast<elem_t*>(gw_out_);
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int tl = warp * 4 + round * 2 + half;  // token within block
        const int64_t t = tok0 + tl;
        const bool valid = t < T;
        float l = 0.0f;
        if (j < E) {
            if (logits_in != nullptr) {
                if (valid) l = BF16 ? __bfloat162float(((const __nv_bfloat16*)logits_in)[t * E + j])
                                    : ((const float*)logits_in)[t * E + j];
            } else {
                l = __fadd_rn(__fadd_rn(__fadd_rn(red[0][tl][j], red[1][tl][j]), red[2][tl][j]), red[3][tl][j]);
                l = rnd<BF16>(l);
            }
        }
        const int am = (attn_mask != nullptr && valid) ? (attn_mask[t] != 0) : 1;
        int raw, mk;
        float gw, ga;
        const int keep_j = (keep != nullptr && valid && j < E) ? (int)keep[t * E + j] : 1;
        // (a token past T routes a one-hot row: all-zero logits are route_token's slowest input, a nine-way tie)
        route_token<BF16, NDYN, NE, true>(valid ? l : (j == 0 ? 8.0f : 0.0f), j, half, am, rc, raw, mk, gw, ga, keep_j,
                                          keep != nullptr);
        if (valid && j < E) {
            if constexpr (BF16) {
                ((__nv_bfloat16*)logits_out)[t * E + j] = __float2bfloat16_rn(l);
                ((__nv_bfloat16*)gw_out)[t * E + j] = __float2bfloat16_rn(gw);
            } else {
                ((float*)logits_out)[t * E + j] = l;
                if constexpr (MIXED) ((__nv_bfloat16*)gw_out_)[t * E + j] = __float2bfloat16_rn(gw);   // core.py:339
                else ((float*)gw_out)[t * E + j] = gw;
            }
            expert_mask[t * E + j] = mk;
            if (j == 0) top_k[t] = raw;
        }
        s_cnt[tl][j] = (valid && j < n_dyn_k) ? mk : 0;
        s_prob[tl][j] = (valid && j < n_dyn_k) ? ga : 0.0f;
    }
    __syncthreads();
    if (tid < n_dyn_k) {  // fixed-order block partials -> deterministic aux loss and exact counts
        int cnt = 0;
        float pr = 0.0f;
#pragma unroll
        for (int r = 0; r < kRouterBlock; ++r) {
            cnt += s_cnt[r][tid];
            pr = __fadd_rn(pr, s_prob[r][tid]);
        }
        block_counts[(int64_t)blockIdx.x * n_dyn_k + tid] = cnt;
        block_probs[(int64_t)blockIdx.x * n_dyn_k + tid] = pr;
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA-fed persistent router (bf16, the production path).
// One CTA per SM, 29 warps:
//   warp 0       TMA producer: the 16-token x block (16 x H bf16 = 64 KB) is fetched as 64-column boxes
//                (cp.async.bulk.tensor.2d, 128B swizzle) into a 2-stage smem ring -- up to 128 KB in flight per SM
//                with no register cost, which is what it takes to stream HBM at full rate from ~7 blocks per SM;
//   warps 1-4    gate MMA: ldmatrix.x4 (swizzled) + mma.sync.m16n8k16 against W_g held in smem in fragment order,
//                K split four ways, partial logits into a 4-stage smem ring;
//   warps 5-28   routing: three groups of 8 warps (one 16-token block per group per round, half-warp per token)
//                running the shuffle/exp chains of route_token<> while the next blocks stream in.
// mbarriers pace the x ring; named barriers (bar.arrive / bar.sync) pace the logits ring.
constexpr int kTmaThreads = 29 * 32;
constexpr int kRouteGroups = 3;
constexpr int kXStageBytes = kRouterBlock * 2048 * 2;   // sized for H = 2048 (checked at launch: H <= 2048)
constexpr int kXStages = 2;
constexpr int kRedStages = 6;                           // stage s is always consumed by routing group s % 3
static_assert((kXStages & (kXStages - 1)) == 0, "x ring is indexed with a mask");
static_assert(kRedStages % kRouteGroups == 0 && 2 * kRedStages + kRouteGroups <= 15, "named barrier ids 1..15");
// W_g fragments in smem: n-tile 0 (experts 0-7) [k16][32 lanes] x 8 B, n-tile 1 (experts 8..E-1) [k16][4 (E-8) lanes] x 8 B
__host__ __device__ constexpr int router_wf_bytes(int H, int E) { return (H / 16) * (32 + 4 * (E > 8 ? E - 8 : 0)) * 8; }
__host__ __device__ constexpr int router_smem_bytes(int H, int E) {
    return kXStages * kXStageBytes + router_wf_bytes(H, E) + kRedStages * 4 * kRouterBlock * 16 * 4 +
           kRouteGroups * 2 * 2 * kRouterBlock * kMaxDyn * 4 + 64 + 1024;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}

// DBG: per-role cycle counters (DCMOE_ROUTER_DEBUG=1) -- a separate instantiation, so that the production kernel keeps
// its 56 registers (with the counters compiled in it needed 64 and spilled)
template <int NDYN, int NE, bool DBG>
__global__ void __launch_bounds__(kTmaThreads, 1)
router_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wg,
                  const int32_t* __restrict__ attn_mask, int64_t T, int H, int n_blocks, RouteConsts rc,
                  __nv_bfloat16* __restrict__ logits_out, int64_t* __restrict__ top_k,
                  int32_t* __restrict__ expert_mask, __nv_bfloat16* __restrict__ gw_out,
                  int32_t* __restrict__ block_counts, float* __restrict__ block_probs) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t xs = base;                                                 // [kXStages][H/64][16 rows][128 B]
    const int E_ = NE ? NE : rc.E;
    const int L1 = (E_ > 8 ? E_ - 8 : 0) * 4;                                 // lanes of n-tile 1 that hold real experts
    uint2* wf0 = reinterpret_cast<uint2*>(gbase + kXStages * kXStageBytes);   // [H/16][32]
    uint2* wf1 = wf0 + (H >> 4) * 32;                                         // [H/16][L1]
    float* red = reinterpret_cast<float*>(gbase + kXStages * kXStageBytes + router_wf_bytes(H, E_));   // [6][4][16][16]
    int* s_cnt = reinterpret_cast<int*>(red + kRedStages * 4 * kRouterBlock * 16);       // [3 groups][2 buffers][16][16]
    float* s_prob = reinterpret_cast<float*>(s_cnt + kRouteGroups * 2 * kRouterBlock * kMaxDyn);
    const uint32_t bars = smem_u32(s_prob + kRouteGroups * 2 * kRouterBlock * kMaxDyn);
    auto x_full = [&](int s) { return bars + 8u * s; };
    auto x_empty = [&](int s) { return bars + 8u * (kXStages + s); };

    unsigned long long* const dbg = DBG ? rc.dbg : nullptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = NE ? NE : rc.E;
    const int n_dyn = NDYN ? NDYN : rc.n_dyn;
    const int n_chunks = H >> 6;   // 64-column boxes per row block
    if (tid == 0) {
        prefetch_tmap(&tmap_x);
        for (int s = 0; s < kXStages; ++s) {
            mbar_init(x_full(s), 1);
            mbar_init(x_empty(s), 4);
        }
        fence_barrier_init();
    }
    __syncthreads();   // barriers initialised
    const long long t_begin = dbg ? clock64() : 0;
    int prod_it = 0;
    if (warp == 0) {
        // the producer puts the first stages in flight before anyone stages W_g
        for (int blk = blockIdx.x; blk < n_blocks && prod_it < kXStages; blk += gridDim.x, ++prod_it) {
            // one box per lane: a TMA issue costs the issuing thread ~50 cycles (tools/probe_stream.cu)
            if (lane == 0) mbar_expect_tx(x_full(prod_it), (uint32_t)(kRouterBlock * H * 2));
            __syncwarp();
            for (int c = lane; c < n_chunks; c += 32)
                tma_load_2d(xs + prod_it * kXStageBytes + c * (kRouterBlock * 128), &tmap_x, c * 64, blk * kRouterBlock,
                            x_full(prod_it));
            __syncwarp();
        }
    } else {
        // W_g -> smem in mma B-fragment order (done by the 20 non-producer warps while the first x blocks are in
        // flight): wf[k16][nt][lane] = {W[n][16 k16 + 2 tq .. +1], W[n][16 k16 + 8 + 2 tq .. +1]}, n = 8 nt + lane / 4,
        // tq = lane % 4; rows n >= E are zero.  One 16-byte load (8 consecutive k of one row) feeds four entries.
        // One item = (row n, k16): 32 contiguous bytes of W_g -> the four (tq) 8-byte entries of that row, which are
        // contiguous in the fragment array, written as two 128-bit stores.  Consecutive threads take consecutive
        // rows of one k16, so a warp writes 1 KB contiguous: no bank conflicts (the former 4-byte scatter was
        // 16-way conflicted and took 5,400 cycles = 15 % of the kernel).
        const int n_k16 = H >> 4;
        const int ng0 = E < 8 ? E : 8, ng1 = E > 8 ? E - 8 : 0;
        const int n0_items = ng0 * n_k16, n_items = (ng0 + ng1) * n_k16;
        constexpr int kStageUnroll = 2;
        for (int i0 = tid - 32; i0 < n_items; i0 += kStageUnroll * (kTmaThreads - 32)) {
            uint4 v0[kStageUnroll], v1[kStageUnroll];
            uint4* dst[kStageUnroll];
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int i = i0 + u * (kTmaThreads - 32);
                dst[u] = nullptr;
                if (i >= n_items) continue;
                int n, k16;
                if (i < n0_items) {
                    k16 = i / ng0;
                    n = i - k16 * ng0;
                    dst[u] = reinterpret_cast<uint4*>(wf0 + k16 * 32 + n * 4);
                } else {
                    const int i1 = i - n0_items;
                    k16 = i1 / ng1;
                    n = 8 + i1 - k16 * ng1;
                    dst[u] = reinterpret_cast<uint4*>(wf1 + k16 * L1 + (n - 8) * 4);
                }
                const __nv_bfloat16* src = wg + (int64_t)n * H + k16 * 16;
                v0[u] = ld_ca_v4(src);
                v1[u] = ld_ca_v4(src + 8);
            }
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                if (dst[u] == nullptr) continue;
                dst[u][0] = make_uint4(v0[u].x, v1[u].x, v0[u].y, v1[u].y);   // entries tq = 0, 1: {k 0..7 half, k 8..15 half}
                dst[u][1] = make_uint4(v0[u].z, v1[u].z, v0[u].w, v1[u].w);   // entries tq = 2, 3
            }
        }
        if (E < 8) {   // n-tile 0 rows E..7 must read as zero
            for (int i = tid - 32; i < (8 - E) * n_k16; i += kTmaThreads - 32) {
                const int k16 = i / (8 - E), n = E + i % (8 - E);
                uint4* d = reinterpret_cast<uint4*>(wf0 + k16 * 32 + n * 4);
                d[0] = d[1] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }

    // W_g fragments staged: a barrier of the 28 consumer warps only -- the producer is still issuing its first 64
    // boxes (~60 cycles each in the TMA unit) and needs nothing from this phase (barrier 0 is not used again)
    if (warp != 0) named_bar_sync(0, kTmaThreads - 32);
    const long long t_staged = dbg ? clock64() : 0;
    long long d0 = 0, d1 = 0, d2 = 0;
    constexpr int kFullCount = 128 + 256;   // gate warps + one routing group
    if (warp == 0) {
        // ================= TMA producer (continues after the stages issued above) =================
        int it = prod_it;
        for (int blk = blockIdx.x + prod_it * gridDim.x; blk < n_blocks; blk += gridDim.x, ++it) {
            const int st = it & (kXStages - 1);
            const uint32_t ph = (uint32_t)(it / kXStages) & 1u;
            const long long q0 = dbg ? clock64() : 0;
            mbar_wait(x_empty(st), ph ^ 1u);
            if (dbg) d0 += clock64() - q0;
            if (lane == 0) mbar_expect_tx(x_full(st), (uint32_t)(kRouterBlock * H * 2));
            __syncwarp();
            for (int c = lane; c < n_chunks; c += 32)
                tma_load_2d(xs + st * kXStageBytes + c * (kRouterBlock * 128), &tmap_x, c * 64, blk * kRouterBlock,
                            x_full(st));
            __syncwarp();
        }
    } else if (warp <= 4) {
        // ================= gate MMA warps =================
        const int wq = warp - 1;
        const int chunks_per_warp = n_chunks >> 2;
        const int g = lane >> 2, tq = lane & 3;
        const int mi = lane >> 3;                              // ldmatrix: lane -> (matrix, row)
        const int lrow = (lane & 7) + (mi & 1) * 8;
        int it = 0;
        for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
            const int st = it & (kXStages - 1);
            const uint32_t ph = (uint32_t)(it / kXStages) & 1u;
            const int rs = it % kRedStages;
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
            const long long q0 = dbg ? clock64() : 0;
            mbar_wait(x_full(st), ph);
            const long long q1 = dbg ? clock64() : 0;
            for (int cc = 0; cc < chunks_per_warp; ++cc) {
                const int c = wq * chunks_per_warp + cc;
                const uint32_t tile = xs + st * kXStageBytes + c * (kRouterBlock * 128);
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int lchunk = 2 * s4 + (mi >> 1);
                    uint32_t a0, a1, a2, a3;
                    ldmatrix_x4(tile + lrow * 128 + ((lchunk ^ (lrow & 7)) << 4), a0, a1, a2, a3);
                    const int k16 = c * 4 + s4;
                    const uint2 b0 = wf0[k16 * 32 + lane];
                    const uint2 b1 = lane < L1 ? wf1[k16 * L1 + lane] : make_uint2(0u, 0u);
                    mma_bf16_16816(c0, a0, a1, a2, a3, b0.x, b0.y);
                    mma_bf16_16816(c1, a0, a1, a2, a3, b1.x, b1.y);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(x_empty(st));           // this warp is done reading the x stage
            const long long q2 = dbg ? clock64() : 0;
            if (it >= kRedStages) named_bar_sync(7 + rs, kFullCount);
            if (dbg) { d0 += q1 - q0; d1 += q2 - q1; d2 += clock64() - q2; }
            float* r = red + ((rs * 4 + wq) * kRouterBlock) * 16;
            r[g * 16 + 2 * tq] = c0[0];
            r[g * 16 + 2 * tq + 1] = c0[1];
            r[(g + 8) * 16 + 2 * tq] = c0[2];
            r[(g + 8) * 16 + 2 * tq + 1] = c0[3];
            r[g * 16 + 8 + 2 * tq] = c1[0];
            r[g * 16 + 8 + 2 * tq + 1] = c1[1];
            r[(g + 8) * 16 + 8 + 2 * tq] = c1[2];
            r[(g + 8) * 16 + 8 + 2 * tq + 1] = c1[3];
            named_bar_arrive(1 + rs, kFullCount);
        }
        if (dbg && warp == 1 && lane == 0) {
            dbg[blockIdx.x * 16 + 2] = d0;
            dbg[blockIdx.x * 16 + 3] = d1;
            dbg[blockIdx.x * 16 + 4] = d2;
            dbg[blockIdx.x * 16 + 10] = it;
            dbg[blockIdx.x * 16 + 11] = clock64() - t_begin;   // gate warps done
        }
    } else {
        // ================= routing warps: group g = warps 5+8g .. 12+8g =================
        const int grp = (warp - 5) >> 3, rw_ = (warp - 5) & 7;
        const int half = lane >> 4, j = lane & 15;
        const int gtid = tid - (5 + grp * 8) * 32;
        int* cnt0 = s_cnt + grp * 2 * kRouterBlock * kMaxDyn;
        float* prob0 = s_prob + grp * 2 * kRouterBlock * kMaxDyn;
        int it = grp;
        for (int blk = blockIdx.x + grp * gridDim.x; blk < n_blocks; blk += kRouteGroups * gridDim.x, it += kRouteGroups) {
            const int rs = it % kRedStages;
            const int64_t tok0 = (int64_t)blk * kRouterBlock;
            const int tl = rw_ * 2 + half;
            const int64_t t = tok0 + tl;
            const bool valid = t < T;
            const long long q0 = dbg ? clock64() : 0;
            named_bar_sync(1 + rs, kFullCount);
            const long long q1 = dbg ? clock64() : 0;
            float l = 0.0f;
            if (j < E) {
                const float* r = red + (rs * 4 * kRouterBlock + tl) * 16 + j;
                l = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[kRouterBlock * 16]), r[2 * kRouterBlock * 16]), r[3 * kRouterBlock * 16]);
                l = bf16_round(l);
            }
            if ((int64_t)blk + (int64_t)kRedStages * gridDim.x < n_blocks) named_bar_arrive(7 + rs, kFullCount);
            const int am = (attn_mask != nullptr && valid) ? (attn_mask[t] != 0) : 1;
            int raw, mk;
            float gw, ga;
            route_token<true, NDYN, NE>(valid ? l : (j == 0 ? 8.0f : 0.0f), j, half, am, rc, raw, mk, gw, ga);
            const long long q2 = dbg ? clock64() : 0;
            if (valid && j < E) {
                logits_out[t * E + j] = __float2bfloat16_rn(l);
                gw_out[t * E + j] = __float2bfloat16_rn(gw);
                expert_mask[t * E + j] = mk;
                if (j == 0) top_k[t] = raw;
            }
            const int sbuf = (it / kRouteGroups) & 1;                 // double-buffered block statistics
            int* cnt = cnt0 + sbuf * kRouterBlock * kMaxDyn;
            float* prob = prob0 + sbuf * kRouterBlock * kMaxDyn;
            cnt[tl * kMaxDyn + j] = (valid && j < n_dyn) ? mk : 0;
            prob[tl * kMaxDyn + j] = (valid && j < n_dyn) ? ga : 0.0f;
            named_bar_sync(13 + grp, 256);
            if (gtid < n_dyn) {
                int c = 0;
                float pr = 0.0f;
#pragma unroll
                for (int r = 0; r < kRouterBlock; ++r) {
                    c += cnt[r * kMaxDyn + gtid];
                    pr = __fadd_rn(pr, prob[r * kMaxDyn + gtid]);
                }
                block_counts[(int64_t)blk * n_dyn + gtid] = c;
                block_probs[(int64_t)blk * n_dyn + gtid] = pr;
            }
            if (dbg) { d0 += q1 - q0; d1 += q2 - q1; d2 += clock64() - q2; }
        }
        if (dbg && warp == 5 && lane == 0) {
            dbg[blockIdx.x * 16 + 5] = d0;
            dbg[blockIdx.x * 16 + 6] = d1;
            dbg[blockIdx.x * 16 + 7] = d2;
            dbg[blockIdx.x * 16 + 8] = clock64() - t_begin;   // routing group 0 done
            dbg[blockIdx.x * 16 + 9] = t_staged - t_begin;
        }
    }
    if (dbg && warp == 0 && lane == 0) {
        dbg[blockIdx.x * 16 + 0] = d0;
        dbg[blockIdx.x * 16 + 1] = clock64() - t_begin;       // producer done
    }
}

// ------------------------------------------------------------------------------------------------
// Decode front end (bf16, T <= 64): router + plan + permute in ONE single-CTA launch.
// The generation loop calls the layer with T = 2N tokens (reference model.py:1149-1203); at that size the three
// kernels above are pure launch + latency (14.7 + 6.6 + 6.6 us measured at T = 2).  Here one 32-warp CTA does the
// gate projection (token block x K-eighth per warp, mma.sync), routes every token in a single round (32 warps x 2
// tokens, the same route_token<> -> identical bits), builds the plan in shared memory (counts, segment bases, tile
// table, aux) and gathers the selected rows into x_packed.
constexpr int kFrontMaxT = 64;

template <int NDYN, int NE>
__global__ void __launch_bounds__(1024, 1)
front_small_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wg,
                   const int32_t* __restrict__ attn_mask, int T, int H, RouteConsts rc, int n_real, int t_pad,
                   int max_mtiles, __nv_bfloat16* __restrict__ logits_out, int64_t* __restrict__ top_k,
                   int32_t* __restrict__ expert_mask, __nv_bfloat16* __restrict__ gw_out, PlanView pv,
                   __nv_bfloat16* __restrict__ x_packed, int32_t* __restrict__ slot_of, int32_t* __restrict__ row_token,
                   float* __restrict__ row_scale) {
    __shared__ float red[4][8][kRouterBlock][16];
    __shared__ unsigned char s_mask[kFrontMaxT][kMaxDyn];
    __shared__ float s_ga[kFrontMaxT][kMaxDyn];
    __shared__ float s_gw[kFrontMaxT][kMaxDyn];
    __shared__ int s_slot[kFrontMaxT][kMaxDyn];
    __shared__ int s_cnt[kMaxDyn];
    __shared__ int s_seg[kMaxDyn + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = NE ? NE : rc.E;
    const int n_dyn = NDYN ? NDYN : rc.n_dyn;
    grid_dep_wait();     // (launched programmatically dependent on whatever precedes it, e.g. dcmoe_rmsnorm)
    grid_dep_launch();   // the GEMM-1 CTAs may come up on the other SMs and run their prologue meanwhile
    const long long t_fs0 = rc.dbg ? clock64() : 0;
    // ---- phase 1: gate projection, warp = (token block of 16, K slice) ----
    // all 32 warps work whatever T is: K is cut 32 / 16 / 8 ways for <= 16 / 32 / 64 tokens (one round of loads per
    // warp at T <= 16: the phase is a DRAM round trip, not arithmetic); a slice is a multiple of 32 columns
    const int nblk2 = T <= 16 ? 1 : (T <= 32 ? 2 : 4);
    const int ksplit = min(32 / nblk2, H >> 5);
    {
        const int blk = warp / ksplit, ks = warp - blk * ksplit;
        if (blk < nblk2 && blk * kRouterBlock < T) {
            gate_slice16(x, wg, T, H, E, blk, ks, ksplit, lane, (&red[0][0])[blk * ksplit + ks]);
        }
    }
    __syncthreads();
    if (rc.dbg && tid == 0) rc.dbg[1] = clock64() - t_fs0;   // gate projection done
    // ---- phase 2: routing, warp = 2 tokens ----
    {
        const int half = lane >> 4, j = lane & 15;
        const int t = warp * 2 + half;
        const bool valid = t < T;
        float l = 0.0f;
        if (valid && j < E) {
            const int blk = t >> 4, tl = t & 15;
            float (*rb)[16][16] = &red[0][0] + blk * ksplit;
            l = rb[0][tl][j];
            for (int ks = 1; ks < ksplit; ++ks) l = __fadd_rn(l, rb[ks][tl][j]);
            l = bf16_round(l);
        }
        const int am = (attn_mask != nullptr && valid) ? (attn_mask[t] != 0) : 1;
        int raw = 0, mk = 0;
        float gw = 0.0f, ga = 0.0f;
        // Warps without a token skip the routing; the idle half of the last warp routes a one-hot row.  (All-zero
        // logits are the WORST case of route_token -- a nine-way tie, seven near-tie softmaxes -- and the 28 idle
        // warps of a T = 8 call used to hold the barrier below for 9,400 cycles after the real tokens were done.)
        if (warp * 2 < T) {
            const float lr = valid ? l : (j == 0 ? 8.0f : 0.0f);
            route_token<true, NDYN, NE>(lr, j, half, am, rc, raw, mk, gw, ga);
        }
        if (valid && j < E) {
            logits_out[(int64_t)t * E + j] = __float2bfloat16_rn(l);
            gw_out[(int64_t)t * E + j] = __float2bfloat16_rn(gw);
            expert_mask[(int64_t)t * E + j] = mk;
            if (j == 0) top_k[t] = raw;
        }
        if (t < kFrontMaxT) {
            s_mask[t][j] = (valid && j < E) ? mk : 0;
            s_ga[t][j] = (valid && j < n_dyn) ? ga : 0.0f;
            s_gw[t][j] = (valid && j < E) ? gw : 0.0f;
        }
    }
    __syncthreads();
    if (rc.dbg && tid == 0) rc.dbg[2] = clock64() - t_fs0;   // routing done
    // ---- phase 3: plan ----
    if (tid < n_real) {
        int c = 0;
        for (int t = 0; t < T; ++t) c += s_mask[t][tid];
        s_cnt[tid] = c;
        if (rc.dbg && tid == 0) rc.dbg[9] = clock64() - t_fs0;
        pv.counts[tid] = c;
        if (rc.dbg && tid == 0) rc.dbg[10] = clock64() - t_fs0;
    }
    if (rc.dbg && tid == 64) rc.dbg[11] = clock64() - t_fs0;
    __syncthreads();
    if (rc.dbg && tid == 0) rc.dbg[6] = clock64() - t_fs0;
    if (tid < 32) {
        // segment bases and the tile table: lane e lays out expert e (exclusive warp scan of the experts' tile counts);
        // the shared-expert tiles come first
        const int n_sh = t_pad / kTileM;
        const int cnt = tid < n_real ? s_cnt[tid] : 0;
        const int nt = (cnt + kTileM - 1) / kTileM;
        int incl = nt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(kFull, incl, off);
            if (tid >= off) incl += o;
        }
        const int excl = incl - nt;
        const int total_routed = __shfl_sync(kFull, incl, 31);
        if (tid < n_real) {
            const int row = t_pad + excl * kTileM;
            s_seg[tid] = row;
            pv.seg_base[tid] = row;
            for (int i = 0; i < nt; ++i) {
                const int tile = n_sh + excl + i;
                if (tile < max_mtiles) {
                    dcmoe_mtile mt;
                    mt.out_row = row + i * kTileM; mt.a_row = mt.out_row - t_pad; mt.group = tid;
                    mt.rows = min(kTileM, cnt - i * kTileM);
                    pv.mtiles[tile] = mt;
                }
            }
        }
        for (int i = tid; i < n_sh && i < max_mtiles; i += 32) {
            dcmoe_mtile mt;
            mt.a_row = i * kTileM; mt.out_row = i * kTileM; mt.group = n_real;
            mt.rows = min(kTileM, T - i * kTileM);
            pv.mtiles[i] = mt;
        }
        if (tid == 0) {
            const int end_row = t_pad + total_routed * kTileM;
            s_seg[n_real] = end_row;
            pv.seg_base[n_real] = end_row;
            *pv.n_mtiles = min(n_sh + total_routed, max_mtiles);
            *pv.overflow = (end_row > max_mtiles * kTileM) ? 1 : 0;
            if (rc.dbg) rc.dbg[7] = clock64() - t_fs0;
        }
    }
    if (tid >= 32 && tid < 32 + n_dyn) {
        // aux loss, same reduction shape as router + plan kernels: fp32 partials per block of 16 tokens, then fp64
        const int jx = tid - 32;
        double ps = 0.0;
        long long ts = 0;
        for (int b = 0; b * kRouterBlock < T; ++b) {
            float pr = 0.0f;
            int cn = 0;
            for (int r = 0; r < kRouterBlock; ++r) {
                const int t = b * kRouterBlock + r;
                if (t < T) { cn += s_mask[t][jx]; pr = __fadd_rn(pr, s_ga[t][jx]); }
                else pr = __fadd_rn(pr, 0.0f);
            }
            ps += (double)pr;
            ts += cn;
        }
        float tpe = (float)((double)ts / (double)T);
        float rp = bf16_round((float)(ps / (double)T));
        s_ga[0][jx] = tpe * rp;   // reuse as scratch (all reads of column jx are done by this thread)
        if (rc.dbg && jx == 0) rc.dbg[8] = clock64() - t_fs0;
    }
    __syncthreads();
    if (rc.dbg && tid == 0) rc.dbg[3] = clock64() - t_fs0;   // plan done
    if (tid == 0) {
        double acc = 0.0;
        for (int jx = 0; jx < n_dyn; ++jx) acc += (double)s_ga[0][jx];
        *pv.aux_loss = (float)acc * (float)n_dyn;
    }
    // ---- phase 4: slots, scales, row gather ----
    for (int i = tid; i < T * n_real; i += 1024) {
        const int t = i / n_real, e = i - t * n_real;
        int rank = 0;
        for (int q = 0; q < t; ++q) rank += s_mask[q][e];
        int slot = -1;
        if (s_mask[t][e]) {
            slot = s_seg[e] + rank;
            if (slot >= max_mtiles * kTileM) {
                slot = -1;     // row_capacity below the worst case and exceeded: dropped (plan.overflow = 1)
            } else {
                row_token[slot] = t;
                if (slot < DCMOE_SMALL_ROWS) pv.small_tokens[slot] = t;   // (the weight-streaming GEMM-1 gathers its rows from x)
                row_scale[2 * (int64_t)slot] = s_gw[t][e];
                row_scale[2 * (int64_t)slot + 1] = s_gw[t][e];
            }
        }
        s_slot[t][e] = slot;
        slot_of[i] = slot;
    }
    if (tid < T) {
        row_token[tid] = tid;
        row_scale[2 * tid] = s_gw[tid][n_dyn];
        row_scale[2 * tid + 1] = (E - n_dyn) > 1 ? s_gw[tid][n_dyn + 1] : 0.0f;
    }
    __syncthreads();
    if (rc.dbg && tid == 0) rc.dbg[4] = clock64() - t_fs0;   // permute staged / plan visible
    const int n_vec = H >> 3;   // 16-byte vectors per row
    // (x_packed == nullptr: more than 16 tokens -- one CTA would copy the rows at ~50 GB/s, 11 us for 40 tokens; the
    // launcher follows up with gather_rows_kernel on many CTAs instead)
    for (int p = warp; x_packed != nullptr && p < T * n_real; p += 32) {
        const int t = p / n_real, e = p - t * n_real;
        const int slot = s_slot[t][e];
        if (slot < 0) continue;
        const uint4* src = reinterpret_cast<const uint4*>(x + (int64_t)t * H);
        uint4* dst = reinterpret_cast<uint4*>(x_packed + (int64_t)(slot - t_pad) * H);
        for (int c0 = 0; c0 < n_vec; c0 += 256) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u * 32 + lane;
                if (c < n_vec) v[u] = ld_ca_v4(src + c);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u * 32 + lane;
                if (c < n_vec) st_na_v4(dst + c, v[u]);
            }
        }
    }
    if (rc.dbg) {
        __syncthreads();
        if (tid == 0) rc.dbg[5] = clock64() - t_fs0;   // rows gathered
    }
}

// row gather of the decode front end for 16 < T <= 64: one warp per (token, routed expert) pair
__global__ void __launch_bounds__(256) gather_rows_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ slot_of,
                                                          int n_pairs, int n_real, int H, int t_pad,
                                                          __nv_bfloat16* __restrict__ x_packed) {
    grid_dep_wait();
    grid_dep_launch();
    const int p = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= n_pairs) return;
    const int slot = slot_of[p];
    if (slot < 0) return;
    const int t = p / n_real, n_vec = H >> 3;
    const uint4* src = reinterpret_cast<const uint4*>(x + (int64_t)t * H);
    uint4* dst = reinterpret_cast<uint4*>(x_packed + (int64_t)(slot - t_pad) * H);
    for (int c0 = 0; c0 < n_vec; c0 += 256) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * 32 + lane;
            if (c < n_vec) v[u] = ld_ca_v4(src + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * 32 + lane;
            if (c < n_vec) st_na_v4(dst + c, v[u]);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Token drop, drop_policy == "probs" (core.py:305-314): per dynamic column, keep the `capacity` tokens with the largest
// logit among the tokens that selected the expert.  One CTA per column: the selected tokens' 64-bit keys
// (order-preserving image of the logit, then ~token so that ties go to the LOWER token index -- torch.topk leaves ties
// to the implementation) are compacted once, then an 8-pass MSB radix select finds the capacity-th largest key; the
// keys are unique, so exactly `capacity` tokens survive.  Integer work, exact.
__device__ __forceinline__ unsigned ordered_bits(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <bool LBF16>
__global__ void __launch_bounds__(1024) drop_select_kernel(const void* __restrict__ logits_, const int32_t* __restrict__ mask,
                                                           int64_t T, int E, long long capacity,
                                                           unsigned long long* __restrict__ keys_all,
                                                           uint8_t* __restrict__ keep) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_n;
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned long long s_k;
    const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    unsigned long long* keys = keys_all + (int64_t)e * T;
    auto key_of = [&](int64_t t) -> unsigned long long {
        const float l = LBF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(logits_)[t * E + e])
                              : static_cast<const float*>(logits_)[t * E + e];
        return ((unsigned long long)ordered_bits(l) << 32) | (unsigned long long)(0xffffffffu - (unsigned)t);
    };
    if (tid == 0) s_n = 0u;
    __syncthreads();
    for (int64_t t0 = 0; t0 < T; t0 += 1024) {
        const int64_t t = t0 + tid;
        const bool sel = t < T && mask[t * E + e] != 0;
        const unsigned bal = __ballot_sync(kFull, sel);
        unsigned base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_n, (unsigned)__popc(bal));
        base = __shfl_sync(kFull, base, 0);
        if (sel) keys[base + __popc(bal & ((1u << lane) - 1u))] = key_of(t);
    }
    __syncthreads();
    const unsigned n = s_n;
    unsigned long long threshold = 0ull;                  // keep every selected token
    if (capacity <= 0) {
        threshold = ~0ull;                                // torch.topk(k = 0): nothing survives (no key reaches all ones)
    } else if ((long long)n > capacity) {
        if (tid == 0) { s_prefix = 0ull; s_k = (unsigned long long)capacity; }
        for (int pass = 7; pass >= 0; --pass) {
            if (tid < 256) hist[tid] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            for (unsigned i = tid; i < n; i += 1024) {
                const unsigned long long k = keys[i];
                if (pass == 7 || (k >> (8 * (pass + 1))) == prefix) atomicAdd(&hist[(unsigned)(k >> (8 * pass)) & 255u], 1u);
            }
            __syncthreads();
            if (tid < 32) {
                // lane q owns bins 255 - 8q .. 248 - 8q (descending): find the bin that holds the k-th largest key
                unsigned own = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) own += hist[255 - (8 * lane + b)];
                unsigned incl = own;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned v = __shfl_up_sync(kFull, incl, off);
                    if (lane >= off) incl += v;
                }
                const unsigned long long k = s_k;
                const bool mine = (unsigned long long)(incl - own) < k && k <= (unsigned long long)incl;
                if (mine) {
                    unsigned long long rem = k - (incl - own);
                    int bin = 255 - 8 * lane;
                    while (rem > hist[bin]) { rem -= hist[bin]; --bin; }
                    s_prefix = (prefix << 8) | (unsigned long long)bin;
                    s_k = rem;
                }
            }
            __syncthreads();
        }
        threshold = s_prefix;
    }
    for (int64_t t = tid; t < T; t += 1024) {
        const bool sel = mask[t * E + e] != 0;
        keep[t * E + e] = (sel && key_of(t) >= threshold) ? 1 : 0;
    }
}

// shared columns of the keep mask are always 1 (core.py:313)
__global__ void drop_fill_shared_kernel(uint8_t* __restrict__ keep, int64_t T, int n_dyn, int E) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nf = E - n_dyn;
    if (i < T * nf) keep[(i / nf) * E + n_dyn + (i % nf)] = 1;
}

// ------------------------------------------------------------------------------------------------
// aux_balance_weight branch of the load-balancing loss (core.py:380-385): weighted means over tokens.
//   tokens_per_expert_e = sum_t float(mask[t,e]) * w_t / sum_t w_t                    (fp32 products)
//   router_prob_e       = sum_t P(ga[t,e] * w_t) / sum_t w_t   with ga = softmax_9(logits masked with finfo.min),
//                         products in the promoted dtype P (D for integer weights, fp32 for float weights)
// Pass 1: half-warp per token, per-block (16 tokens) partial sums in fixed order; pass 2: one CTA, fp64, fixed order.
template <bool BF16>
__global__ void __launch_bounds__(256) aux_weighted_partials_kernel(const void* __restrict__ logits_, const int32_t* __restrict__ mask,
                                                                    const float* __restrict__ w, int64_t T, int n_dyn, int E,
                                                                    float finfo_min, int prod_bf16, float* __restrict__ partials) {
    __shared__ float s_tok[kRouterBlock][kMaxDyn], s_prob[kRouterBlock][kMaxDyn], s_w[kRouterBlock];
    const int tid = threadIdx.x, lane = tid & 31, half = lane >> 4, j = lane & 15;
    const int tl = (tid >> 5) * 2 + half;
    const int64_t t = (int64_t)blockIdx.x * kRouterBlock + tl;
    const bool valid = t < T;
    float l = 0.0f;
    int mk = 0;
    if (valid && j < n_dyn) {
        l = BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(logits_)[t * E + j]) : static_cast<const float*>(logits_)[t * E + j];
        mk = mask[t * E + j];
    }
    const float ga = softmax_lanes<BF16>(j < n_dyn ? (mk ? l : finfo_min) : __int_as_float(0xff800000), j, n_dyn);
    const float wt = valid ? (w != nullptr ? w[t] : 1.0f) : 0.0f;      // w == nullptr: the plain means of core.py:378-379
    float pr = __fmul_rn(ga, wt);
    if (prod_bf16) pr = bf16_round(pr);
    s_tok[tl][j] = (valid && j < n_dyn) ? __fmul_rn((float)mk, wt) : 0.0f;
    s_prob[tl][j] = (valid && j < n_dyn) ? pr : 0.0f;
    if (j == 0) s_w[tl] = wt;
    __syncthreads();
    if (tid < 2 * n_dyn + 1) {
        float acc = 0.0f;
#pragma unroll
        for (int r = 0; r < kRouterBlock; ++r)
            acc = __fadd_rn(acc, tid < n_dyn ? s_tok[r][tid] : (tid < 2 * n_dyn ? s_prob[r][tid - n_dyn] : s_w[r]));
        partials[(int64_t)blockIdx.x * 32 + tid] = acc;
    }
}

__global__ void __launch_bounds__(1024) aux_weighted_finish_kernel(const float* __restrict__ partials, int64_t n_blocks, int n_dyn,
                                                                   int prob_bf16, float* __restrict__ aux_out) {
    __shared__ double s_col[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_cols = 2 * n_dyn + 1;
    for (int c = warp; c < n_cols; c += 32) {
        double acc = 0.0;
        for (int64_t b = lane; b < n_blocks; b += 32) acc += (double)partials[b * 32 + c];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
        if (lane == 0) s_col[c] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double den = s_col[2 * n_dyn];
        double acc = 0.0;
        for (int e = 0; e < n_dyn; ++e) {
            const float tpe = (float)(s_col[e] / den);
            float rp = (float)(s_col[n_dyn + e] / den);
            if (prob_bf16) rp = bf16_round(rp);
            acc += (double)(tpe * rp);
        }
        *aux_out = (float)acc * (float)n_dyn;
    }
}

}  // namespace

int launch_drop_select(const void* logits, bool logits_bf16, const int32_t* expert_mask, int64_t T, const dcmoe_config* cfg,
                       int64_t capacity, void* key_scratch, uint8_t* keep, cudaStream_t stream) {
    if (T == 0) return DCMOE_OK;
    const int n_dyn = cfg->n_real + cfg->n_null, E = n_dyn + cfg->n_fix;
    if (T > 0x7fffffffll) {
        set_error("dcmoe_drop_select: T must fit 31 bits");
        return DCMOE_ERR_INVALID;
    }
    if (logits_bf16)
        drop_select_kernel<true><<<n_dyn, 1024, 0, stream>>>(logits, expert_mask, T, E, (long long)capacity,
                                                             (unsigned long long*)key_scratch, keep);
    else
        drop_select_kernel<false><<<n_dyn, 1024, 0, stream>>>(logits, expert_mask, T, E, (long long)capacity,
                                                              (unsigned long long*)key_scratch, keep);
    if (cfg->n_fix > 0)
        drop_fill_shared_kernel<<<(unsigned)ceil_div(T * cfg->n_fix, 256), 256, 0, stream>>>(keep, T, n_dyn, E);
    return check_cuda(cudaGetLastError(), "drop_select_kernel launch");
}

int launch_aux_weighted(const void* logits, bool arith_bf16, const int32_t* expert_mask, const float* w, bool prod_bf16,
                        int64_t T, const dcmoe_config* cfg, float* scratch, float* aux_out, cudaStream_t stream) {
    const int n_dyn = cfg->n_real + cfg->n_null, E = n_dyn + cfg->n_fix;
    const int64_t n_blocks = ceil_div(T, kRouterBlock);
    if (T == 0) return DCMOE_OK;
    const float fmin_bf16 = -3.3895313892515355e38f, fmin_f32 = -3.4028234663852886e38f;
    if (arith_bf16)
        aux_weighted_partials_kernel<true><<<(unsigned)n_blocks, 256, 0, stream>>>(logits, expert_mask, w, T, n_dyn, E, fmin_bf16,
                                                                                  prod_bf16 ? 1 : 0, scratch);
    else
        aux_weighted_partials_kernel<false><<<(unsigned)n_blocks, 256, 0, stream>>>(logits, expert_mask, w, T, n_dyn, E, fmin_f32, 0,
                                                                                   scratch);
    aux_weighted_finish_kernel<<<1, 1024, 0, stream>>>(scratch, n_blocks, n_dyn, (arith_bf16 && prod_bf16) ? 1 : 0, aux_out);
    return check_cuda(cudaGetLastError(), "aux_weighted kernels launch");
}

namespace {
__global__ void exp_cr_test_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = mode == 0 ? exp_cr(x[i]) : (mode == 1 ? exp_cr_double(x[i]) : exp_sleef_u10(x[i]));
}
}  // namespace

int launch_exp_test(const float* x, float* y, int64_t n, int mode, cudaStream_t stream) {
    if (n <= 0) return DCMOE_OK;
    exp_cr_test_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(x, y, n, mode);
    return check_cuda(cudaGetLastError(), "exp_cr_test_kernel launch");
}

int launch_router(const void* x, const void* w_gate, const void* logits_in, const int32_t* attn_mask, const uint8_t* keep,
                  int flags, int64_t T, const dcmoe_config* cfg, void* logits_out, int64_t* top_k, int32_t* expert_mask,
                  void* global_weight, int32_t* block_counts, float* block_probs, cudaStream_t stream) {
    if (T == 0) return DCMOE_OK;
    // DCMOE_ROUTER_FP32_GATE on a bf16 layer: logits and routing arithmetic in fp32 (core.py:240-249)
    const bool mixed = (flags & DCMOE_ROUTER_FP32_GATE) && cfg->dtype == DCMOE_BF16;
    const bool bf16 = cfg->dtype == DCMOE_BF16 && !mixed;
    RouteConsts rc;
    rc.n_dyn = cfg->n_real + cfg->n_null;
    rc.E = rc.n_dyn + cfg->n_fix;
    // scalar operands are rounded to D exactly as torch does for a wrapped python scalar
    auto r = [&](float v) { return bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v; };
    rc.thr_p = r((float)cfg->top_p);
    rc.thr_eps = r((float)(2.0 * cfg->jitter_eps));
    rc.plus_eps = r(1e-6f);
    rc.finfo_min = bf16 ? -3.3895313892515355e38f : -3.4028234663852886e38f;
    {
        const char* dbg = getenv("DCMOE_ROUTER_ALWAYS_SOFTMAX");
        rc.always_softmax = (dbg && dbg[0] == '1') ? 1 : 0;
    }
    rc.dbg = nullptr;
    rc.fixed_k = cfg->top_p == 0.0 ? cfg->fixed_top_k : 0;
    const int64_t n_blocks = ceil_div(T, kRouterBlock);
    dim3 grid((unsigned)n_blocks), block(128);
#define DCMOE_LAUNCH_ROUTER(BF, ND, NE_, MX)                                                                                   \
    router_kernel<BF, ND, NE_, MX><<<grid, block, 0, stream>>>(x, w_gate, logits_in, attn_mask, keep, T, cfg->hidden_size, rc, \
                                                               logits_out, top_k, expert_mask, global_weight,                 \
                                                               block_counts, block_probs)
    const bool ref_shape = rc.n_dyn == 9 && rc.E == 11;   // utils/config.json: 8 routed + 1 null + 2 shared
    const int sms = device_sm_count();
    // (decode-sized calls were measured with the one-CTA-per-block kernel too: 21.8 us vs 14.7 us for the TMA-fed
    // kernel at T = 2, so the persistent kernel is used at every size)
    const int router_smem = router_smem_bytes(cfg->hidden_size, rc.E);
    if (bf16 && logits_in == nullptr && keep == nullptr && cfg->hidden_size <= 2048 && cfg->hidden_size % 256 == 0 &&
        router_smem <= 232448) {
        static PerDeviceOnce attr_once;
        if (attr_once.first()) {
            const int max_smem = router_smem_bytes(2048, 16) > 232448 ? 232448 : router_smem_bytes(2048, 16);
            int rc2 = check_cuda(cudaFuncSetAttribute(router_tma_kernel<9, 11, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem), "cudaFuncSetAttribute(router_tma<9,11>)");
            if (rc2) { attr_once.reset_current(); return rc2; }
            rc2 = check_cuda(cudaFuncSetAttribute(router_tma_kernel<0, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem), "cudaFuncSetAttribute(router_tma<0,0>)");
            if (!rc2) rc2 = check_cuda(cudaFuncSetAttribute(router_tma_kernel<9, 11, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem), "cudaFuncSetAttribute(router_tma<9,11,dbg>)");
            if (rc2) { attr_once.reset_current(); return rc2; }
        }
        CUtensorMap tmap;
        int rc2 = make_tensor_map_bf16(&tmap, x, T, cfg->hidden_size, kRouterBlock);
        if (rc2) return rc2;
        dim3 g2((unsigned)(n_blocks < sms ? n_blocks : sms)), b2(kTmaThreads);
        static unsigned long long* dbg = nullptr;
        const bool debug = getenv("DCMOE_ROUTER_DEBUG") != nullptr;
        if (debug) {
            if (!dbg) cudaMalloc(&dbg, 256 * 16 * sizeof(unsigned long long));
            cudaMemsetAsync(dbg, 0, 256 * 16 * sizeof(unsigned long long), stream);
            rc.dbg = dbg;
        }
        if (ref_shape && debug)     // the counters exist for the reference shape only
            router_tma_kernel<9, 11, true><<<g2, b2, router_smem, stream>>>(tmap, (const __nv_bfloat16*)w_gate, attn_mask, T,
                cfg->hidden_size, (int)n_blocks, rc, (__nv_bfloat16*)logits_out, top_k, expert_mask,
                (__nv_bfloat16*)global_weight, block_counts, block_probs);
        else if (ref_shape)
            router_tma_kernel<9, 11, false><<<g2, b2, router_smem, stream>>>(tmap, (const __nv_bfloat16*)w_gate, attn_mask, T,
                cfg->hidden_size, (int)n_blocks, rc, (__nv_bfloat16*)logits_out, top_k, expert_mask,
                (__nv_bfloat16*)global_weight, block_counts, block_probs);
        else
            router_tma_kernel<0, 0, false><<<g2, b2, router_smem, stream>>>(tmap, (const __nv_bfloat16*)w_gate, attn_mask, T,
                cfg->hidden_size, (int)n_blocks, rc, (__nv_bfloat16*)logits_out, top_k, expert_mask,
                (__nv_bfloat16*)global_weight, block_counts, block_probs);
        if (debug && ref_shape) {   // tuning only: synchronises
            static int printed = 0;
            cudaStreamSynchronize(stream);
            static unsigned long long host[256 * 16];
            cudaMemcpy(host, dbg, sizeof(host), cudaMemcpyDeviceToHost);
            if (printed++ < 4) {
                double a[16] = {0};
                for (unsigned c = 0; c < g2.x; ++c)
                    for (int k = 0; k < 16; ++k) a[k] += (double)host[c * 16 + k] / g2.x;
                fprintf(stderr, "router T=%lld blocks/CTA %.1f | staging %.0f | producer: wait-empty %.0f done@%.0f | gate warp: wait-x %.0f "
                        "mma %.0f wait-red %.0f done@%.0f | routing grp0: wait-logits %.0f route %.0f store+stats %.0f done@%.0f (cycles)\n",
                        (long long)T, a[10], a[9], a[0], a[1], a[2], a[3], a[4], a[11], a[5], a[6], a[7], a[8]);
            }
        }
        return check_cuda(cudaGetLastError(), "router_tma_kernel launch");
    }
    if (bf16) {
        if (ref_shape) DCMOE_LAUNCH_ROUTER(true, 9, 11, false); else DCMOE_LAUNCH_ROUTER(true, 0, 0, false);
    } else if (mixed) {
        if (ref_shape) DCMOE_LAUNCH_ROUTER(false, 9, 11, true); else DCMOE_LAUNCH_ROUTER(false, 0, 0, true);
    } else {
        if (ref_shape) DCMOE_LAUNCH_ROUTER(false, 9, 11, false); else DCMOE_LAUNCH_ROUTER(false, 0, 0, false);
    }
#undef DCMOE_LAUNCH_ROUTER
    return check_cuda(cudaGetLastError(), "router_kernel launch");
}

int launch_front_small(const void* x, const void* w_gate, const int32_t* attn_mask, int64_t T, const dcmoe_config* cfg,
                       const dcmoe_sizes& sz, PlanView pv, void* logits_out, int64_t* top_k, int32_t* expert_mask,
                       void* global_weight, void* x_packed, int32_t* slot_of, int32_t* row_token, float* row_scale,
                       bool gather_rows, cudaStream_t stream) {
    // gather_rows == false: x_packed is not filled -- the caller's GEMM-1 gathers its token rows from x (TMA gather4,
    // dcmoe_forward with T <= 32)
    if (T == 0) return DCMOE_OK;
    if (cfg->dtype != DCMOE_BF16 || T > kFrontMaxT) {
        set_error("dcmoe_front_small: bf16 and T <= %d only", kFrontMaxT);
        return DCMOE_ERR_INVALID;
    }
    RouteConsts rc;
    rc.n_dyn = cfg->n_real + cfg->n_null;
    rc.E = rc.n_dyn + cfg->n_fix;
    auto r = [&](float v) { return __bfloat162float(__float2bfloat16_rn(v)); };
    rc.thr_p = r((float)cfg->top_p);
    rc.thr_eps = r((float)(2.0 * cfg->jitter_eps));
    rc.plus_eps = r(1e-6f);
    rc.finfo_min = -3.3895313892515355e38f;
    rc.always_softmax = 0;
    rc.dbg = nullptr;
    rc.fixed_k = cfg->top_p == 0.0 ? cfg->fixed_top_k : 0;
    const dim3 grid(1), block(1024);
    cudaError_t err;
    void* const x_packed_all = x_packed;
    if (T > 16 || !gather_rows) x_packed = nullptr;     // rows are gathered by gather_rows_kernel below / by the GEMM
    static unsigned long long* dbg = nullptr;
    const bool debug = getenv("DCMOE_ROUTER_DEBUG") != nullptr;
    if (debug) {
        if (!dbg) cudaMalloc(&dbg, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(dbg, 0, 16 * sizeof(unsigned long long), stream);
        rc.dbg = dbg;
    }
    if (rc.n_dyn == 9 && rc.E == 11)
        err = launch_kernel(front_small_kernel<9, 11>, grid, block, 0, stream, pdl_enabled(), (const __nv_bfloat16*)x,
                            (const __nv_bfloat16*)w_gate, attn_mask, (int)T, cfg->hidden_size, rc, cfg->n_real, (int)sz.t_pad,
                            (int)sz.max_mtiles, (__nv_bfloat16*)logits_out, top_k, expert_mask, (__nv_bfloat16*)global_weight, pv,
                            (__nv_bfloat16*)x_packed, slot_of, row_token, row_scale);
    else
        err = launch_kernel(front_small_kernel<0, 0>, grid, block, 0, stream, pdl_enabled(), (const __nv_bfloat16*)x,
                            (const __nv_bfloat16*)w_gate, attn_mask, (int)T, cfg->hidden_size, rc, cfg->n_real, (int)sz.t_pad,
                            (int)sz.max_mtiles, (__nv_bfloat16*)logits_out, top_k, expert_mask, (__nv_bfloat16*)global_weight, pv,
                            (__nv_bfloat16*)x_packed, slot_of, row_token, row_scale);
    if (debug && err == cudaSuccess) {   // tuning only: synchronises
        static int printed = 0;
        cudaStreamSynchronize(stream);
        unsigned long long h[16];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        if (printed++ < 4)
            fprintf(stderr, "front_small T=%lld: gate done@%llu routing done@%llu plan done@%llu slots visible@%llu rows gathered@%llu | counts@%llu tile table@%llu aux@%llu | loop@%llu store@%llu warp2 at barrier@%llu (cycles)\n",
                    (long long)T, h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11]);
    }
    if (err == cudaSuccess && T > 16 && gather_rows) {
        const int n_pairs = (int)T * cfg->n_real;
        err = launch_kernel(gather_rows_kernel, dim3((unsigned)ceil_div(n_pairs, 8)), dim3(256), 0, stream, pdl_enabled(),
                            (const __nv_bfloat16*)x, (const int32_t*)slot_of, n_pairs, cfg->n_real, cfg->hidden_size,
                            (int)sz.t_pad, (__nv_bfloat16*)x_packed_all);
    }
    return check_cuda(err, "front_small_kernel launch");
}

}  // namespace dcmoe
