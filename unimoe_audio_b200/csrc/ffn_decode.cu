// ffn_decode.cu -- expert FFNs for decode-sized calls (T <= 64 tokens, bf16), sm_100a.
//
// With a handful of tokens (the generation loop calls the layer with T = 2N, reference model.py:1149-1203) the
// layer is bound by reading each hit expert's weights ONCE (304 MB per layer, 47 us at the measured HBM peak);
// 128-row tensor-core tiles would stream mostly padding next to them.  Here the WEIGHTS are the M operand:
//   GEMM-1: one warp per (row tile, 8 h columns): 8 gate rows + the matching 8 up rows of W13 form the 16 x K "A"
//           operand of mma.sync.m16n8k16, streamed straight from HBM with 128-bit no-allocate loads; the token rows
//           (<= 64, L1/L2 resident) are the "B" operand, 8 tokens per n-tile.  The K order inside an MMA step is
//           permuted identically for both operands, so each 16-byte load feeds two MMA steps with no shuffles.
//           SwiGLU and the routing weight are applied to the accumulators in registers (gate in c0/c1, up in c2/c3).
//   GEMM-2: one CTA per (row tile, 64 output features): warp = 16 features x one K half, the two halves are
//           reduced through shared memory.
// All 3096 (GEMM-1) / 2304 (GEMM-2) warps are resident at once, so the block scheduler needs no tile balancing
// and >= 9 MB of weight loads are in flight.  Same plan, row space, weight packs and outputs as the tcgen05 path.
#include "common.cuh"

namespace dcmoe {
namespace {

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

constexpr int kMaxNT = 8;   // n-tiles of 8 tokens -> up to 64 rows per group

// h[out_row + n][j] = silu(a_n . Wg_j) * (a_n . Wu_j) * scale      a_n = x row (shared group) or packed row
template <int NT>
__global__ void __launch_bounds__(256)
decode_ffn1_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x_packed,
                   const __nv_bfloat16* __restrict__ w13, const float* __restrict__ row_scale, int H, int Id, int n_real,
                   int I_s, const dcmoe_mtile* __restrict__ mtiles, const int32_t* __restrict__ n_mtiles,
                   __nv_bfloat16* __restrict__ h) {
    const int ti = blockIdx.y;
    if (ti >= *n_mtiles) return;
    const dcmoe_mtile mt = mtiles[ti];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j8 = blockIdx.x * 8 + warp;                 // 8 h columns per warp
    if (j8 * 8 >= Id) return;
    const int g8 = lane >> 2, tq = lane & 3;
    const int hcol = j8 * 8 + g8;
    const int blk = hcol >> 6, in = hcol & 63;            // W13 interleave: blocks of 64 gate rows then 64 up rows
    const __nv_bfloat16* wg = w13 + ((int64_t)mt.group * 2 * Id + (int64_t)blk * 128 + in) * H + tq * 8;
    const __nv_bfloat16* wu = wg + (int64_t)64 * H;
    const __nv_bfloat16* a_base = (mt.group == n_real ? x : x_packed) + (int64_t)mt.a_row * H + tq * 8;
    const int rows = mt.rows;
    float c[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const int n_chunks = H >> 5;                          // 32-column chunks
    for (int c0 = 0; c0 < n_chunks; c0 += 4) {
        uint4 va[4], vb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                     // 8 x 16 B of weights in flight per lane
            va[u] = ld_nc_v4(wg + (c0 + u) * 32);
            vb[u] = ld_nc_v4(wu + (c0 + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int n = nt * 8 + g8;
                const uint4 vx = n < rows ? ld_ca_v4(a_base + (int64_t)n * H + (c0 + u) * 32) : zero;
                mma16816(c[nt], va[u].x, vb[u].x, va[u].y, vb[u].y, vx.x, vx.y);
                mma16816(c[nt], va[u].z, vb[u].z, va[u].w, vb[u].w, vx.z, vx.w);
            }
        }
    }
    const bool shared_grp = mt.group == n_real;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int n = nt * 8 + 2 * tq + q;
            if (n >= rows) continue;
            const int64_t r = (int64_t)mt.out_row + n;
            const float sc = (shared_grp && hcol >= I_s) ? row_scale[2 * r + 1] : row_scale[2 * r];
            h[r * Id + hcol] = __float2bfloat16_rn(silu_mul(c[nt][q], c[nt][2 + q]) * sc);
        }
    }
}

// y[out_row + n][o] = sum_k h[out_row + n][k] * W2[group][o][k]
template <int NT>
__global__ void __launch_bounds__(256)
decode_ffn2_kernel(const __nv_bfloat16* __restrict__ hbuf, const __nv_bfloat16* __restrict__ w2, int H, int Id,
                   const dcmoe_mtile* __restrict__ mtiles, const int32_t* __restrict__ n_mtiles,
                   __nv_bfloat16* __restrict__ y) {
    const int ti = blockIdx.y;
    if (ti >= *n_mtiles) return;
    const dcmoe_mtile mt = mtiles[ti];
    __shared__ float red[4][NT][32][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = warp & 3, khalf = warp >> 2;         // 4 units of 16 output features, 2 K halves
    const int o0 = (blockIdx.x * 4 + unit) * 16;
    const int g8 = lane >> 2, tq = lane & 3;
    const int n_chunks = Id >> 5;                         // 86 for I_d = 2752
    const int cbeg = khalf ? (n_chunks >> 1) : 0, cend = khalf ? n_chunks : (n_chunks >> 1);
    const bool ok = o0 < H;
    const __nv_bfloat16* wa = w2 + ((int64_t)mt.group * H + (ok ? o0 : 0) + g8) * Id + tq * 8;
    const __nv_bfloat16* wb = wa + (int64_t)8 * Id;
    const __nv_bfloat16* a_base = hbuf + (int64_t)mt.out_row * Id + tq * 8;
    const int rows = mt.rows;
    float c[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int c0 = cbeg; c0 < cend; c0 += 4) {
        uint4 va[4], vb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool in = c0 + u < cend;
            va[u] = in ? ld_nc_v4(wa + (c0 + u) * 32) : zero;
            vb[u] = in ? ld_nc_v4(wb + (c0 + u) * 32) : zero;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (c0 + u >= cend) break;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int n = nt * 8 + g8;
                const uint4 vx = n < rows ? ld_ca_v4(a_base + (int64_t)n * Id + (c0 + u) * 32) : zero;
                mma16816(c[nt], va[u].x, vb[u].x, va[u].y, vb[u].y, vx.x, vx.y);
                mma16816(c[nt], va[u].z, vb[u].z, va[u].w, vb[u].w, vx.z, vx.w);
            }
        }
    }
    if (khalf == 1) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) red[unit][nt][lane][q] = c[nt][q];
    }
    __syncthreads();
    if (khalf == 0 && ok) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int n = nt * 8 + 2 * tq + q;
                if (n >= rows) continue;
                const int64_t r = (int64_t)mt.out_row + n;
                y[r * H + o0 + g8] = __float2bfloat16_rn(c[nt][q] + red[unit][nt][lane][q]);
                y[r * H + o0 + g8 + 8] = __float2bfloat16_rn(c[nt][2 + q] + red[unit][nt][lane][2 + q]);
            }
        }
    }
}

}  // namespace

int launch_ffn_decode(const void* x, const void* x_packed, const void* w13, const void* w2, const float* row_scale,
                      int64_t T, const dcmoe_config* cfg, const dcmoe_sizes& sz, PlanView pv, void* h, void* y, int phase,
                      cudaStream_t stream) {
    if (T == 0) return DCMOE_OK;
    if (cfg->dtype != DCMOE_BF16 || T > 8 * kMaxNT) {
        set_error("decode FFN: bf16 and T <= %d only (got T = %lld)", 8 * kMaxNT, (long long)T);
        return DCMOE_ERR_INVALID;
    }
    const int H = cfg->hidden_size, Id = cfg->dynamic_intermediate_size;
    // every group has at most T rows: one row tile per group, n_real + 1 tiles at most
    const int max_tiles = cfg->n_real + 1;
    dim3 block(256);
    dim3 g1((unsigned)ceil_div(Id, 64), (unsigned)max_tiles), g2((unsigned)ceil_div(H, 64), (unsigned)max_tiles);
    const int nt = (int)ceil_div(T, 8);
#define DCMOE_DECODE_LAUNCH(NT_)                                                                                         \
    do {                                                                                                                 \
        if (phase != 2)                                                                                                  \
            decode_ffn1_kernel<NT_><<<g1, block, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)x_packed,   \
                (const __nv_bfloat16*)w13, row_scale, H, Id, cfg->n_real, cfg->shared_intermediate_size, pv.mtiles,      \
                pv.n_mtiles, (__nv_bfloat16*)h);                                                                         \
        if (phase != 1)                                                                                                  \
            decode_ffn2_kernel<NT_><<<g2, block, 0, stream>>>((const __nv_bfloat16*)h, (const __nv_bfloat16*)w2, H, Id,  \
                pv.mtiles, pv.n_mtiles, (__nv_bfloat16*)y);                                                              \
    } while (0)
    if (nt <= 1) DCMOE_DECODE_LAUNCH(1);
    else if (nt <= 2) DCMOE_DECODE_LAUNCH(2);
    else if (nt <= 4) DCMOE_DECODE_LAUNCH(4);
    else DCMOE_DECODE_LAUNCH(8);
#undef DCMOE_DECODE_LAUNCH
    (void)sz;
    return check_cuda(cudaGetLastError(), "decode_ffn kernel launch");
}

}  // namespace dcmoe
