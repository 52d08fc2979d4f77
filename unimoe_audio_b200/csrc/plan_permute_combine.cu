// plan_permute_combine.cu -- the integer / byte-moving side of the DCMoE layer (sm_100a).
//
//   plan_kernel     (one launch, also finishes the aux loss) exact per-expert histogram (from the router's per-block counts) -> exclusive
//                   prefix sums over token blocks -> 128-row aligned expert segments -> m-tile table
//                   for the grouped GEMMs -> aux loss.  Replaces reference core.py:455 (capacity =
//                   mask.sum(0).max(); here exact counts, no padding to the max) and finishes
//                   core.py:376-389.  All integer work is exact; the aux reduction order is fixed.
//   permute_kernel  replaces core.py:459-462 + utils/UniMoE_Audio_utils.py:436-485 (compress_matrix,
//                   argsort + gather of a [T, 8, H] expansion): each token row is read ONCE with
//                   128-bit loads and written to the r_t packed rows it was routed to.  The slot of
//                   (token, expert) = segment base + block prefix + rank inside the 16-token block,
//                   i.e. the stable (ascending token id) permutation.
//   combine (ep.cu: ep_combine_kernel, one rank)  replaces core.py:486-488 + utils.py:488-523 (decompress_matrix, scatter into
//                   [T, 8, H] zeros, weighted einsum) and core.py:338-353 (shared-expert adds):
//                   per token, gather of the shared row + <= n_real routed rows (already weighted by
//                   the GEMM-1 epilogue), fp32 accumulation in fixed expert order, one store.
//                   HBM bound: (1 + r_t) * H * sizeof(D) read + H * sizeof(D) written per token.
#include "common.cuh"

namespace dcmoe {
namespace {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// One CTA, 16 warps.  The token blocks are cut into 16 contiguous chunks, one per warp:
//   phase 1  every warp sums its chunk per column (lane = block, 32 blocks per round, all loads of a round in flight:
//            counts as integers, the aux-softmax partials in fp64) -> chunk totals;
//   phase 2  16 threads prefix the 16 chunk totals per column -> chunk bases, expert counts, aux terms;
//   phase 3  every warp rescans its chunk (L1 / L2 hits) and writes the exclusive per-block offsets: warp scan across
//            the 32 blocks of a round + the running carry + the chunk base;
//   phase 4  segment bases (thread 0) and the tile table (all threads).
// The former version scanned an expert's blocks with ONE warp, 512 blocks per dependent step: 14 us at 1,024 blocks,
// 129 us at 8,192, 262 us at 16,384 (serial latency in front of the permute and the GEMMs).
// (kPlanWarps warps; CN = compile-time column bound: 9 for the reference's expert counts, where four rounds of loads
// -- 4 x 18 values per lane -- are in flight at once, else kMaxDyn with one round)
constexpr int kPlanWarps = 16;
template <bool BF16, int CN>
__global__ void __launch_bounds__(kPlanWarps * 32) plan_kernel(int n_blocks, int n_dyn, int n_real, int64_t T, int t_pad,
                                                                int max_mtiles, PlanView pv) {
    constexpr int U = CN <= 9 ? 4 : 1;             // rounds of 32 blocks whose loads are issued together
    __shared__ int s_chunk_cnt[kPlanWarps][kMaxDyn];
    __shared__ double s_chunk_prob[kPlanWarps][kMaxDyn];
    __shared__ int s_chunk_base[kPlanWarps][kMaxDyn];
    __shared__ int s_counts[kMaxDyn];
    __shared__ int s_seg[kMaxDyn + 1];
    __shared__ int s_tile0[kMaxDyn + 1];
    __shared__ double s_term[kMaxDyn];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = (n_blocks + kPlanWarps - 1) / kPlanWarps;
    const int b0 = min(n_blocks, warp * per_warp), b1 = min(n_blocks, b0 + per_warp);

    {   // ---- phase 1: chunk totals ----
        int cnt[CN];
        double pr[CN];
#pragma unroll
        for (int j = 0; j < CN; ++j) { cnt[j] = 0; pr[j] = 0.0; }
        for (int bb = b0 + lane; bb < b1; bb += 32 * U) {
            int c[U][CN];
            float q[U][CN];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int b = bb + 32 * u;
#pragma unroll
                for (int j = 0; j < CN; ++j) {
                    const bool ok = b < b1 && j < n_dyn;
                    c[u][j] = ok ? pv.block_counts[(int64_t)b * n_dyn + j] : 0;
                    q[u][j] = ok ? pv.block_probs[(int64_t)b * n_dyn + j] : 0.0f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int j = 0; j < CN; ++j) {
                    cnt[j] += c[u][j];
                    pr[j] += (double)q[u][j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < CN; ++j) {
            if (j < n_dyn) {                                   // (warp-uniform) fixed xor tree: bit-stable run to run
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    cnt[j] += __shfl_xor_sync(kFull, cnt[j], off);
                    pr[j] += __shfl_xor_sync(kFull, pr[j], off);
                }
                if (lane == 0) {
                    s_chunk_cnt[warp][j] = cnt[j];
                    s_chunk_prob[warp][j] = pr[j];
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < n_dyn) {   // ---- phase 2: chunk bases, totals, aux terms (core.py:376-389, aux_balance_weight = None) ----
        const int j = threadIdx.x;
        int base = 0;
        double ps = 0.0;
        for (int w = 0; w < kPlanWarps; ++w) {                 // fixed order over the chunks
            s_chunk_base[w][j] = base;
            base += s_chunk_cnt[w][j];
            ps += s_chunk_prob[w][j];
        }
        s_counts[j] = base;
        if (j < n_real) pv.counts[j] = base;
        float tpe = (float)((double)base / (double)T);         // torch.mean(expert_mask.float(), 0)
        float rp = (float)(ps / (double)T);                    // torch.mean(global_weight, 0) ...
        if (BF16) rp = bf16_round(rp);                         // ... is a D tensor (bf16 rounds here)
        s_term[j] = (double)(tpe * rp);
    }
    __syncthreads();
    {   // ---- phase 3: exclusive per-block offsets of the routed experts ----
        int carry[CN];
#pragma unroll
        for (int e = 0; e < CN; ++e) carry[e] = e < n_real ? s_chunk_base[warp][e] : 0;
        for (int bb = b0; bb < b1; bb += 32 * U) {
            int c[U][CN];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int b = bb + 32 * u + lane;
#pragma unroll
                for (int e = 0; e < CN; ++e) c[u][e] = (b < b1 && e < n_real) ? pv.block_counts[(int64_t)b * n_dyn + e] : 0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int b = bb + 32 * u + lane;
                if (bb + 32 * u >= b1) break;                  // (warp-uniform)
#pragma unroll
                for (int e = 0; e < CN; ++e) {
                    if (e < n_real) {                          // warp-uniform
                        int incl = c[u][e];
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const int o = __shfl_up_sync(kFull, incl, off);
                            if (lane >= off) incl += o;
                        }
                        if (b < b1) pv.block_offsets[(int64_t)b * n_real + e] = carry[e] + incl - c[u][e];
                        carry[e] += __shfl_sync(kFull, incl, 31);
                    }
                }
            }
        }
    }
    if (threadIdx.x == 0) {   // ---- phase 4 ----
        int row = t_pad;
        int tile = t_pad / kTileM;
        for (int e = 0; e < n_real; ++e) {
            s_seg[e] = row;
            s_tile0[e] = tile;
            pv.seg_base[e] = row;
            const int nt = (s_counts[e] + kTileM - 1) / kTileM;
            row += nt * kTileM;
            tile += nt;
        }
        s_seg[n_real] = row;
        s_tile0[n_real] = tile;
        pv.seg_base[n_real] = row;
        *pv.n_mtiles = tile < max_mtiles ? tile : max_mtiles;
        *pv.overflow = tile > max_mtiles ? 1 : 0;   // only possible when the caller chose a row_capacity below the worst case
        double acc = 0.0;
        for (int j = 0; j < n_dyn; ++j) acc += s_term[j];         // fixed left-to-right order over experts
        *pv.aux_loss = (float)acc * (float)n_dyn;
    }
    __syncthreads();
    // m-tile table: shared-expert tiles first (largest group first), then routed experts
    const int n_shared_tiles = t_pad / kTileM;
    const int total = s_tile0[n_real];
    for (int i = threadIdx.x; i < total && i < max_mtiles; i += blockDim.x) {
        dcmoe_mtile mt;
        if (i < n_shared_tiles) {
            mt.a_row = i * kTileM;
            mt.out_row = i * kTileM;
            mt.group = n_real;
            const int64_t left = T - (int64_t)i * kTileM;
            mt.rows = (int)(left < kTileM ? left : kTileM);
        } else {
            int e = 0;
            while (e + 1 < n_real && i >= s_tile0[e + 1]) ++e;
            const int local = i - s_tile0[e];
            mt.out_row = s_seg[e] + local * kTileM;
            mt.a_row = mt.out_row - t_pad;
            mt.group = e;
            const int left = s_counts[e] - local * kTileM;
            mt.rows = left < kTileM ? left : kTileM;
        }
        pv.mtiles[i] = mt;
    }
}

// ------------------------------------------------------------------------------------------------
template <int ESIZE>  // bytes per element
__global__ void __launch_bounds__(128) permute_kernel(const char* __restrict__ x, const int32_t* __restrict__ mask,
                                                      const char* __restrict__ gw, int64_t T, int H, int n_real,
                                                      int n_dyn, int n_fix, int t_pad, int row_limit, PlanView pv,
                                                      char* __restrict__ x_packed, int32_t* __restrict__ slot_of,
                                                      int32_t* __restrict__ row_token, float* __restrict__ row_scale) {
    __shared__ int s_m[kRouterBlock][kMaxDyn];
    __shared__ int s_slot[kRouterBlock][kMaxDyn];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = n_dyn + n_fix;
    const int64_t tok0 = (int64_t)blockIdx.x * kRouterBlock;
    auto load_gw = [&](int64_t t, int j) -> float {
        if (ESIZE == 2) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gw)[t * E + j]);
        return reinterpret_cast<const float*>(gw)[t * E + j];
    };
    for (int i = tid; i < kRouterBlock * n_real; i += 128) {
        const int tl = i / n_real, e = i % n_real;
        const int64_t t = tok0 + tl;
        s_m[tl][e] = (t < T) ? mask[t * E + e] : 0;
    }
    __syncthreads();
    for (int i = tid; i < kRouterBlock * n_real; i += 128) {
        const int tl = i / n_real, e = i % n_real;
        const int64_t t = tok0 + tl;
        int rank = 0;
        for (int q = 0; q < tl; ++q) rank += s_m[q][e];
        int slot = -1;
        if (s_m[tl][e]) {
            slot = pv.seg_base[e] + pv.block_offsets[(int64_t)blockIdx.x * n_real + e] + rank;
            if (slot >= row_limit) {
                slot = -1;     // row_capacity below the worst case and exceeded: the row is dropped (plan.overflow = 1)
            } else {
                row_token[slot] = (int32_t)t;
                if (T <= 64 && slot < DCMOE_SMALL_ROWS) pv.small_tokens[slot] = (int32_t)t;   // decode-sized: GEMM-1 gathers from x
                const float w = load_gw(t, e);
                row_scale[2 * (int64_t)slot] = w;
                row_scale[2 * (int64_t)slot + 1] = w;
            }
        }
        s_slot[tl][e] = slot;
        if (t < T) slot_of[t * n_real + e] = slot;
    }
    if (tid < kRouterBlock) {
        const int64_t t = tok0 + tid;
        if (t < T) {
            row_token[t] = (int32_t)t;
            row_scale[2 * t] = load_gw(t, n_dyn);
            row_scale[2 * t + 1] = n_fix > 1 ? load_gw(t, n_dyn + 1) : 0.0f;
        }
    }
    __syncthreads();
    // row copies: one warp per token, 8 x 128-bit in flight per lane
    const int n_vec = H * ESIZE / 16;
    for (int q = 0; q < kRouterBlock / 4; ++q) {
        const int tl = warp * (kRouterBlock / 4) + q;
        const int64_t t = tok0 + tl;
        if (t >= T) continue;
        int n_dst = 0;
        for (int e = 0; e < n_real; ++e) n_dst += s_slot[tl][e] >= 0;
        if (n_dst == 0) continue;
        const uint4* src = reinterpret_cast<const uint4*>(x + t * (int64_t)H * ESIZE);
        for (int c0 = 0; c0 < n_vec; c0 += 256) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u * 32 + lane;
                if (c < n_vec) v[u] = ld_nc_v4(src + c);
            }
            for (int e = 0; e < n_real; ++e) {
                const int slot = s_slot[tl][e];
                if (slot < 0) continue;
                uint4* dst = reinterpret_cast<uint4*>(x_packed + (int64_t)(slot - t_pad) * H * ESIZE);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = c0 + u * 32 + lane;
                    if (c < n_vec) st_na_v4(dst + c, v[u]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Weight packing: W13[group] rows = blocks of 64 gate rows followed by the matching 64 up rows;
// W2[group] = down_proj, shared experts concatenated along K.
template <int ESIZE>
__global__ void pack_kernel(const char* __restrict__ gate_proj, const char* __restrict__ up_proj,
                            const char* __restrict__ down_proj, int H, int I_part, int I_total, int part,
                            char* __restrict__ w13_group, char* __restrict__ w2_group) {
    const int n_vec_h = H * ESIZE / 16;
    // W13: one CTA row per source row r of gate/up
    for (int r = blockIdx.x; r < I_part; r += gridDim.x) {
        const int gcol = part * I_part + r;  // column in the packed intermediate dimension
        const int blk = gcol / 64, in = gcol % 64;
        const int64_t grow = (int64_t)blk * 128 + in, urow = grow + 64;
        const uint4* gs = reinterpret_cast<const uint4*>(gate_proj + (int64_t)r * H * ESIZE);
        const uint4* us = reinterpret_cast<const uint4*>(up_proj + (int64_t)r * H * ESIZE);
        uint4* gd = reinterpret_cast<uint4*>(w13_group + grow * H * ESIZE);
        uint4* ud = reinterpret_cast<uint4*>(w13_group + urow * H * ESIZE);
        for (int c = threadIdx.x; c < n_vec_h; c += blockDim.x) {
            gd[c] = gs[c];
            ud[c] = us[c];
        }
    }
    // W2: down_proj [H, I_part] -> columns [part*I_part, (part+1)*I_part) of [H, I_total]
    for (int n = blockIdx.x; n < H; n += gridDim.x) {
        const char* s = down_proj + (int64_t)n * I_part * ESIZE;
        char* d = w2_group + ((int64_t)n * I_total + (int64_t)part * I_part) * ESIZE;
        for (int c = threadIdx.x; c < I_part * ESIZE / 4; c += blockDim.x)
            reinterpret_cast<uint32_t*>(d)[c] = reinterpret_cast<const uint32_t*>(s)[c];
    }
}

}  // namespace

int launch_plan(int64_t T, const dcmoe_config* cfg, const dcmoe_sizes& sz, PlanView pv, cudaStream_t stream) {
    const int n_dyn = cfg->n_real + cfg->n_null;
#define DCMOE_LAUNCH_PLAN(BF, CN_) \
    plan_kernel<BF, CN_><<<1, kPlanWarps * 32, 0, stream>>>((int)sz.n_blocks, n_dyn, cfg->n_real, T, (int)sz.t_pad, (int)sz.max_mtiles, pv)
    const bool bf = cfg->dtype == DCMOE_BF16;
    if (n_dyn <= 9) { if (bf) DCMOE_LAUNCH_PLAN(true, 9); else DCMOE_LAUNCH_PLAN(false, 9); }
    else { if (bf) DCMOE_LAUNCH_PLAN(true, kMaxDyn); else DCMOE_LAUNCH_PLAN(false, kMaxDyn); }
#undef DCMOE_LAUNCH_PLAN
    return check_cuda(cudaGetLastError(), "plan_kernel launch");
}

int launch_permute(const void* x, const int32_t* expert_mask, const void* gw, int64_t T, const dcmoe_config* cfg,
                   const dcmoe_sizes& sz, PlanView pv, void* x_packed, int32_t* slot_of, int32_t* row_token,
                   float* row_scale, cudaStream_t stream) {
    if (T == 0) return DCMOE_OK;
    const int n_dyn = cfg->n_real + cfg->n_null;
    dim3 grid((unsigned)sz.n_blocks), block(128);
    if (cfg->dtype == DCMOE_BF16)
        permute_kernel<2><<<grid, block, 0, stream>>>((const char*)x, expert_mask, (const char*)gw, T, cfg->hidden_size,
                                                      cfg->n_real, n_dyn, cfg->n_fix, (int)sz.t_pad, (int)(sz.max_mtiles * kTileM), pv,
                                                      (char*)x_packed, slot_of, row_token, row_scale);
    else
        permute_kernel<4><<<grid, block, 0, stream>>>((const char*)x, expert_mask, (const char*)gw, T, cfg->hidden_size,
                                                      cfg->n_real, n_dyn, cfg->n_fix, (int)sz.t_pad, (int)(sz.max_mtiles * kTileM), pv,
                                                      (char*)x_packed, slot_of, row_token, row_scale);
    return check_cuda(cudaGetLastError(), "permute_kernel launch");
}

int launch_pack(const void* gate_proj, const void* up_proj, const void* down_proj, int group, int part,
                const dcmoe_config* cfg, void* w13, void* w2, cudaStream_t stream) {
    const int H = cfg->hidden_size, Id = cfg->dynamic_intermediate_size;
    const bool shared = group == cfg->n_real;
    const int I_part = shared ? cfg->shared_intermediate_size : Id;
    const int es = cfg->dtype == DCMOE_BF16 ? 2 : 4;
    char* w13g = (char*)w13 + (int64_t)group * 2 * Id * H * es;
    char* w2g = (char*)w2 + (int64_t)group * H * Id * es;
    if (es == 2)
        pack_kernel<2><<<512, 256, 0, stream>>>((const char*)gate_proj, (const char*)up_proj, (const char*)down_proj, H,
                                                I_part, Id, shared ? part : 0, w13g, w2g);
    else
        pack_kernel<4><<<512, 256, 0, stream>>>((const char*)gate_proj, (const char*)up_proj, (const char*)down_proj, H,
                                                I_part, Id, shared ? part : 0, w13g, w2g);
    return check_cuda(cudaGetLastError(), "pack_kernel launch");
}

}  // namespace dcmoe
