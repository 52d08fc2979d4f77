// ffn_tcgen05.cu -- grouped expert FFN on 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces the 24 + 6 cuBLAS calls of the reference (core.py:406-416, :48-49 routed experts on
// capacity-padded blocks; core.py:344-349, :30-31 shared experts) with two persistent grouped GEMMs
// over "row space" (include/dcmoe_b200.h):
//
//   GEMM-1  h[row, :] = silu(a W_gate^T) * (a W_up^T) * row_scale       a = x row or packed row
//           B operand = W13[group] : gate/up rows interleaved in blocks of 64, so one 128x256 fp32
//           accumulator tile in TMEM holds 128 gate and the matching 128 up columns and SwiGLU, the
//           routing weight (core.py:447) and the shared-expert weights (core.py:348) are applied in
//           the epilogue; h is written once, in bf16.
//   GEMM-2  y[row, :] = h[row, :] W2[group]^T
//
// Kernel anatomy (one CTA per SM, 256 threads, persistent over a flattened (m-tile, n-tile) list
// read from the device-side plan -- no host round trip, ragged expert sizes cost no padded FLOPs
// beyond the last 128-row tile of each expert):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (128B swizzle) A 128x64 + B 256x64 bf16 per stage
//   warp 1      MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16, M=128, N=256 (128 on
//               the half tile at the end of 2*I_d = 5504), K=16 x 4 per stage; tcgen05.commit frees
//               the smem stage / publishes the accumulator
//   warp 2      TMEM allocator (512 columns = 2 accumulator stages of 256 fp32 columns)
//   warps 4-7   epilogue: tcgen05.ld 32x32b.x32 -> registers -> SwiGLU/scale -> bf16 -> 128B-swizzled
//               smem slab -> per-warp TMA store (cp.async.bulk.tensor ... bulk_group), double buffered
//   pipelines   4-stage smem ring (full/empty mbarriers), 2-stage TMEM ring (tmem_full/tmem_empty)
#include <cuda.h>

#include <cstdio>

#include "common.cuh"
#include "ptx.cuh"

namespace dcmoe {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;             // 16 KB
constexpr int B_BYTES = BN * BK * 2;             // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KB
constexpr int EPI_SLAB = 32 * 128;               // 32 rows x 64 bf16
constexpr int EPI_BUFS = 1;                      // slabs per epilogue warp (1: leaves ~18 KB of smem per SM for a
                                                 // co-resident dispatch / combine CTA when expert parallelism overlaps them)
constexpr int EPI_BYTES = 4 * EPI_BUFS * EPI_SLAB;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 256;

// K-major, 128B-swizzled smem operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 |
// SBO = 1024 B (8 rows x 128 B) | version 1 | layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

struct GemmParams {
    int n_tiles;        // n-tiles per m-tile
    int n_last;         // accumulator columns of the last n-tile
    int num_kb;         // K / 64
    int w_rows;         // B rows per weight group
    int n_real;
    int split_col;      // GEMM-1: h column where the second shared expert starts (I_s)
    const dcmoe_mtile* mtiles;
    const int32_t* n_mtiles;
    const float* row_scale;
    int bn;             // accumulator columns per tile: 256, or 128 for small token counts (more, finer tiles so
                        // that every SM streams weights when the layer is weight-bandwidth bound)
    int m_begin;        // first m-tile of this launch
    int m_end;          // one past the last m-tile; < 0: read *n_mtiles
};

// ------------------------------------------------------------------ kernel
template <bool SWIGLU>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ffn_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a0,   // GEMM-1: x          GEMM-2: h
                const __grid_constant__ CUtensorMap tmap_a1,   // GEMM-1: x_packed   GEMM-2: h
                const __grid_constant__ CUtensorMap tmap_b,    // W13 / W2
                const __grid_constant__ CUtensorMap tmap_out,  // h / y   (box 64 x 32)
                const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
    const uint32_t bar_base = epi_base + EPI_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a0);
        prefetch_tmap(&tmap_a1);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int m_end = p.m_end >= 0 ? min(p.m_end, *p.n_mtiles) : *p.n_mtiles;
    const int total_tiles = max(m_end - p.m_begin, 0) * p.n_tiles;

    if (warp == 0) {
        // ================= TMA producer =================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const dcmoe_mtile mt = p.mtiles[p.m_begin + tile / p.n_tiles];
            const int nt = tile % p.n_tiles;
            const CUtensorMap* amap = SWIGLU ? (mt.group == p.n_real ? &tmap_a0 : &tmap_a1) : &tmap_a0;
            const int a_row = SWIGLU ? mt.a_row : mt.out_row;
            const int b_row = mt.group * p.w_rows + nt * p.bn;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (lane == 0) {
                    const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), (uint32_t)(A_BYTES + p.bn * BK * 2));
                    tma_load_2d(a_dst, amap, kb * BK, a_row, full_bar(stage));
                    tma_load_2d(a_dst + A_BYTES, &tmap_b, kb * BK, b_row, full_bar(stage));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles;
            const uint32_t idesc = make_idesc(nt == p.n_tiles - 1 ? p.n_last : p.bn);
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
                    const uint64_t adesc = make_smem_desc(a_addr);
                    const uint64_t bdesc = make_smem_desc(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                  (uint32_t)((kb | k) != 0));
                    umma_commit(empty_bar(stage));
                    if (kb == p.num_kb - 1) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int wq = warp - 4;  // TMEM lane quarter == warp_id % 4
        const uint32_t slab0 = epi_base + (uint32_t)(wq * EPI_BUFS * EPI_SLAB);
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t chunk_ctr = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const dcmoe_mtile mt = p.mtiles[p.m_begin + tile / p.n_tiles];
            const int nt = tile % p.n_tiles;
            const int n_acc = nt == p.n_tiles - 1 ? p.n_last : p.bn;
            const int n_chunks = SWIGLU ? n_acc / 128 : n_acc / 64;
            float sa = 1.0f, sb = 1.0f;
            if (SWIGLU) {
                const int64_t r = (int64_t)mt.out_row + wq * 32 + lane;
                sa = p.row_scale[2 * r];
                sb = p.row_scale[2 * r + 1];
            }
            const bool shared_grp = mt.group == p.n_real;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN);
            for (int j = 0; j < n_chunks; ++j) {
                const uint32_t slab = slab0 + (chunk_ctr % EPI_BUFS) * EPI_SLAB;
                if (lane == 0) tma_wait_read<EPI_BUFS - 1>();  // the store that last used this slab has drained
                __syncwarp();
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t packed[16];
                    if (SWIGLU) {
                        uint32_t g[32], u[32];
                        tmem_ld32(t_row + (uint32_t)(128 * j + 32 * hf), g);
                        tmem_ld32(t_row + (uint32_t)(128 * j + 64 + 32 * hf), u);
                        tmem_ld_wait();
                        const int hcol0 = nt * (p.bn / 2) + 64 * j + 32 * hf;
                        const float sc = (shared_grp && hcol0 >= p.split_col) ? sb : sa;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float v0 = silu_mul(__uint_as_float(g[2 * i]), __uint_as_float(u[2 * i])) * sc;
                            const float v1 = silu_mul(__uint_as_float(g[2 * i + 1]), __uint_as_float(u[2 * i + 1])) * sc;
                            packed[i] = pack_bf16(v0, v1);
                        }
                    } else {
                        uint32_t v[32];
                        tmem_ld32(t_row + (uint32_t)(64 * j + 32 * hf), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            packed[i] = pack_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                    }
                    // row `lane` of the slab: 128 B = 8 x 16 B chunks, 128B-swizzled (chunk ^ (row & 7))
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t chunk = (uint32_t)(4 * hf + q) ^ (uint32_t)(lane & 7);
                        const uint32_t addr = slab + (uint32_t)lane * 128u + chunk * 16u;
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[4 * q]),
                                     "r"(packed[4 * q + 1]), "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                                     : "memory");
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    const int col0 = SWIGLU ? nt * (p.bn / 2) + 64 * j : nt * p.bn + 64 * j;
                    tma_store_2d(&tmap_out, slab, col0, mt.out_row + wq * 32);
                    tma_commit_group();
                }
                ++chunk_ctr;
            }
            // accumulator stage fully read -> hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) tma_wait_all();
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

int launch_ffn_tcgen05(const void* x, const void* x_packed, const void* w13, const void* w2, const float* row_scale,
                       int64_t T, int64_t row_capacity, const dcmoe_config* cfg, const dcmoe_sizes& sz, PlanView pv,
                       void* h, void* y, int phase, int group_sel, int max_ctas, cudaStream_t stream) {
    if (T == 0) return DCMOE_OK;
    if (cfg->dtype != DCMOE_BF16) {
        set_error("tcgen05 FFN is bf16 only (fp32 layers use the CUDA-core path, impl = 1)");
        return DCMOE_ERR_INVALID;
    }
    const int H = cfg->hidden_size, Id = cfg->dynamic_intermediate_size;
    const int G = cfg->n_real + 1;
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        int rc = check_cuda(cudaFuncSetAttribute(ffn_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                            "cudaFuncSetAttribute(gemm1)");
        if (rc) { attr_once.reset_current(); return rc; }
        rc = check_cuda(cudaFuncSetAttribute(ffn_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                        "cudaFuncSetAttribute(gemm2)");
        if (rc) { attr_once.reset_current(); return rc; }
    }
    CUtensorMap m_x, m_xp, m_w13, m_h_st, m_h_ld, m_w2, m_y_st;
    int rc;
    const int64_t packed_rows = row_capacity - sz.t_pad;
    if ((rc = make_tensor_map_bf16(&m_x, x, T, H, BM))) return rc;
    if ((rc = make_tensor_map_bf16(&m_xp, x_packed, packed_rows > 0 ? packed_rows : 1, H, BM))) return rc;
    const int bn = BN;   // (decode sizes, T <= 64, run ffn_tcgen05_stream.cu instead: narrower tiles of this kernel
                         // were measured slower -- whole tiles per SM leave the chip 52-68 % busy)
    if ((rc = make_tensor_map_bf16(&m_w13, w13, (int64_t)G * 2 * Id, H, bn))) return rc;
    if ((rc = make_tensor_map_bf16(&m_h_st, h, row_capacity, Id, 32))) return rc;
    if ((rc = make_tensor_map_bf16(&m_h_ld, h, row_capacity, Id, BM))) return rc;
    if ((rc = make_tensor_map_bf16(&m_w2, w2, (int64_t)G * H, Id, bn))) return rc;
    if ((rc = make_tensor_map_bf16(&m_y_st, y, row_capacity, H, 32))) return rc;

    GemmParams p1, p2;
    p1.bn = bn;
    p1.n_tiles = (int)ceil_div(2 * Id, bn);
    p1.n_last = 2 * Id - (p1.n_tiles - 1) * bn;
    p1.num_kb = H / BK;
    p1.w_rows = 2 * Id;
    p1.n_real = cfg->n_real;
    p1.split_col = cfg->shared_intermediate_size;
    p1.mtiles = pv.mtiles;
    p1.n_mtiles = pv.n_mtiles;
    p1.row_scale = row_scale;
    // group_sel: 0 = every m-tile, 1 = shared-expert tiles only, 2 = routed tiles only (shared tiles come first):
    // expert parallelism runs the shared experts while remote rows / remote weights are still in flight
    const int n_shared_tiles = (int)(sz.t_pad / BM);
    p1.m_begin = group_sel == 2 ? n_shared_tiles : 0;
    p1.m_end = group_sel == 1 ? n_shared_tiles : -1;
    p2 = p1;
    p2.n_tiles = (int)ceil_div(H, bn);
    p2.n_last = H - (p2.n_tiles - 1) * bn;
    p2.num_kb = Id / BK;
    p2.w_rows = H;

    // DCMOE_FFN_MAX_CTAS (debug / tuning): run the persistent GEMMs on fewer SMs than the chip has, e.g. to leave
    // SMs to the expert-parallel dispatch / combine kernels that run concurrently
    int n_ctas = device_sm_count();
    if (max_ctas > 0 && max_ctas < n_ctas) n_ctas = max_ctas;
    dim3 grid((unsigned)n_ctas), block(NUM_THREADS);
    if (phase != 2) {
        ffn_gemm_kernel<true><<<grid, block, SMEM_BYTES, stream>>>(m_x, m_xp, m_w13, m_h_st, p1);
        if ((rc = check_cuda(cudaGetLastError(), "ffn_gemm_kernel<SwiGLU> launch"))) return rc;
    }
    if (phase != 1) ffn_gemm_kernel<false><<<grid, block, SMEM_BYTES, stream>>>(m_h_ld, m_h_ld, m_w2, m_y_st, p2);
    return check_cuda(cudaGetLastError(), "ffn_gemm_kernel<down> launch");
}

}  // namespace dcmoe
