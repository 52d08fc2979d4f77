// ffn_tcgen05_2cta.cu -- the grouped expert FFN on CTA pairs (tcgen05.mma.cta_group::2), sm_100a.
//
// Same maths, data layout, plan and epilogue as ffn_tcgen05.cu; what changes is how one tile is fed:
// two CTAs of a cluster (the two SMs of a TPC) cooperate on a 256-row x 256-column accumulator tile.
//   * each CTA TMA-loads its own 128 rows of A and ONE HALF (128 of the 256 rows) of the B tile per k-block:
//     32 KB per stage per SM instead of 48 KB, i.e. a third less L2->SM and shared-memory fill traffic;
//   * the leader CTA's MMA warp issues tcgen05.mma.cta_group::2 (M = 256, N = 256): both SMs' tensor cores run,
//     each reads its own A rows and both B halves, each accumulates its 128 rows in its own TMEM;
//   * 6-stage smem ring (6 x 32 KB); the TMA loads of both CTAs signal the LEADER's full barrier (peer bit of the
//     mbarrier address cleared), tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both
//     CTAs, and the peer's epilogue warps release the accumulator with a remote mbarrier arrive.
// Tiles are PAIRS of consecutive 128-row m-tiles of one weight group (plan: `pairs`); an odd last tile leaves the
// peer CTA's half unused (its MMA rows are garbage and are never stored).
#include <cuda.h>

#include <cstdio>

#include "common.cuh"
#include "ptx.cuh"

namespace dcmoe {
namespace {

constexpr int BM = 128, BN = 256, BNH = 128, BK = 64, STAGES = 6;
constexpr int A_BYTES = BM * BK * 2;             // 16 KB
constexpr int B_BYTES = BNH * BK * 2;            // 16 KB: this CTA's half of the B tile
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB per CTA
constexpr int EPI_SLAB = 32 * 128;
constexpr int EPI_BYTES = 4 * EPI_SLAB;          // one slab per epilogue warp
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 256;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta)
        : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar & kPeerBitMask)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit: arrive (once) on the barrier at this offset in BOTH CTAs of the pair when the prior MMAs retire
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3)
        : "memory");
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, K-major, M = 256 (pair), N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__device__ __forceinline__ float silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

struct Gemm2Params {
    int n_tiles, n_last, num_kb, w_rows, n_real, split_col;
    const dcmoe_mtile* mtiles;
    const int32_t* pairs;
    const int32_t* n_pairs;
    const float* row_scale;
    int p_begin, p_end;   // pair range of this launch (p_end < 0: read *n_pairs)
};

template <bool SWIGLU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
ffn_gemm2cta_kernel(const __grid_constant__ CUtensorMap tmap_a0,      // GEMM-1: x          GEMM-2: h
                    const __grid_constant__ CUtensorMap tmap_a1,      // GEMM-1: x_packed   GEMM-2: h
                    const __grid_constant__ CUtensorMap tmap_b,       // W13 / W2, box 64 x 128
                    const __grid_constant__ CUtensorMap tmap_b_half,  // W13, box 64 x 64 (the N = 128 tail tile)
                    const __grid_constant__ CUtensorMap tmap_out,     // h / y, box 64 x 32
                    const Gemm2Params p) {
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
    const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a0);
        prefetch_tmap(&tmap_a1);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_b_half);
        prefetch_tmap(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 2);      // leader: its own arrive.expect_tx + the peer producer's remote arrive
            mbar_init(empty_bar(s), 1);     // multicast commit from the leader's MMA warp
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);     // multicast commit
            mbar_init(tempty_bar(a), 8);    // leader: 4 epilogue warps of each CTA
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // the peer's barriers exist before anything signals them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int p_end = p.p_end >= 0 ? min(p.p_end, *p.n_pairs) : *p.n_pairs;
    const int total_tiles = max(p_end - p.p_begin, 0) * p.n_tiles;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int pr = p.pairs[p.p_begin + tile / p.n_tiles];
            const dcmoe_mtile mt = p.mtiles[pr & 0x3fffffff];
            const int nt = tile % p.n_tiles;
            const bool last = nt == p.n_tiles - 1 && p.n_last != BN;
            const CUtensorMap* amap = SWIGLU ? (mt.group == p.n_real ? &tmap_a0 : &tmap_a1) : &tmap_a0;
            const int a_row = (SWIGLU ? mt.a_row : mt.out_row) + (int)rank * BM;
            const int nh = last ? p.n_last / 2 : BNH;                    // B rows this CTA loads
            const int b_row = mt.group * p.w_rows + nt * BN + (int)rank * nh;
            const CUtensorMap* bmap = last ? &tmap_b_half : &tmap_b;
            const uint32_t bytes_cta = (uint32_t)(A_BYTES + nh * BK * 2);
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (lane == 0) {
                    const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * bytes_cta);
                    else mbar_arrive_cluster(full_bar(stage), 0);
                    tma_load_2d_2sm(a_dst, amap, kb * BK, a_row, full_bar(stage));
                    tma_load_2d_2sm(a_dst + A_BYTES, bmap, kb * BK, b_row, full_bar(stage));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ================= MMA issuer (leader CTA only) =================
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int nt = tile % p.n_tiles;
            const uint32_t idesc = make_idesc(nt == p.n_tiles - 1 ? p.n_last : BN);
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
                        umma_bf16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                      (uint32_t)((kb | k) != 0));
                    umma_commit_2sm(empty_bar(stage));
                    if (kb == p.num_kb - 1) umma_commit_2sm(tfull_bar(acc));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if (warp >= 4) {
        // ================= epilogue (both CTAs, own 128 rows) =================
        const int wq = warp - 4;
        const uint32_t slab = epi_base + (uint32_t)(wq * EPI_SLAB);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int pr = p.pairs[p.p_begin + tile / p.n_tiles];
            const dcmoe_mtile mt = p.mtiles[pr & 0x3fffffff];
            const bool active = rank == 0 || (pr & (1 << 30));            // the peer's half exists
            const int out_row = mt.out_row + (int)rank * BM;
            const int nt = tile % p.n_tiles;
            const int n_acc = nt == p.n_tiles - 1 ? p.n_last : BN;
            const int n_chunks = SWIGLU ? n_acc / 128 : n_acc / 64;
            float sa = 1.0f, sb = 1.0f;
            if (SWIGLU && active) {
                const int64_t r = (int64_t)out_row + wq * 32 + lane;
                sa = p.row_scale[2 * r];
                sb = p.row_scale[2 * r + 1];
            }
            const bool shared_grp = mt.group == p.n_real;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN);
            if (active) {
                for (int j = 0; j < n_chunks; ++j) {
                    if (lane == 0) tma_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t packed[16];
                        if (SWIGLU) {
                            uint32_t g[32], u[32];
                            tmem_ld32(t_row + (uint32_t)(128 * j + 32 * hf), g);
                            tmem_ld32(t_row + (uint32_t)(128 * j + 64 + 32 * hf), u);
                            tmem_ld_wait();
                            const int hcol0 = nt * (BN / 2) + 64 * j + 32 * hf;
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
                        const int col0 = SWIGLU ? nt * (BN / 2) + 64 * j : nt * BN + 64 * j;
                        tma_store_2d(&tmap_out, slab, col0, out_row + wq * 32);
                        tma_commit_group();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(tempty_bar(acc));
                else mbar_arrive_cluster(tempty_bar(acc), 0);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) tma_wait_all();
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // nobody leaves while the pair may still signal its barriers / read its smem
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

int make_map64(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows) {
    return make_tensor_map_bf16(map, base, rows, cols, box_rows);
}

int num_sms2() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

}  // namespace

int launch_ffn_tcgen05_2cta(const void* x, const void* x_packed, const void* w13, const void* w2, const float* row_scale,
                            int64_t T, int64_t row_capacity, const dcmoe_config* cfg, const dcmoe_sizes& sz, PlanView pv,
                            void* h, void* y, int phase, int group_sel, int max_ctas, cudaStream_t stream) {
    if (T == 0) return DCMOE_OK;
    const int H = cfg->hidden_size, Id = cfg->dynamic_intermediate_size;
    const int G = cfg->n_real + 1;
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        int rc = check_cuda(cudaFuncSetAttribute(ffn_gemm2cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                            "cudaFuncSetAttribute(gemm1 2cta)");
        if (rc) { attr_once.reset_current(); return rc; }
        rc = check_cuda(cudaFuncSetAttribute(ffn_gemm2cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                        "cudaFuncSetAttribute(gemm2 2cta)");
        if (rc) { attr_once.reset_current(); return rc; }
    }
    CUtensorMap m_x, m_xp, m_w13, m_w13h, m_h_st, m_h_ld, m_w2, m_y_st;
    int rc;
    const int64_t packed_rows = row_capacity - sz.t_pad;
    if ((rc = make_map64(&m_x, x, T, H, BM))) return rc;
    if ((rc = make_map64(&m_xp, x_packed, packed_rows > 0 ? packed_rows : 1, H, BM))) return rc;
    if ((rc = make_map64(&m_w13, w13, (int64_t)G * 2 * Id, H, BNH))) return rc;
    if ((rc = make_map64(&m_w13h, w13, (int64_t)G * 2 * Id, H, 64))) return rc;
    if ((rc = make_map64(&m_h_st, h, row_capacity, Id, 32))) return rc;
    if ((rc = make_map64(&m_h_ld, h, row_capacity, Id, BM))) return rc;
    if ((rc = make_map64(&m_w2, w2, (int64_t)G * H, Id, BNH))) return rc;
    if ((rc = make_map64(&m_y_st, y, row_capacity, H, 32))) return rc;

    Gemm2Params p1, p2;
    p1.n_tiles = (int)ceil_div(2 * Id, BN);
    p1.n_last = 2 * Id - (p1.n_tiles - 1) * BN;
    p1.num_kb = H / BK;
    p1.w_rows = 2 * Id;
    p1.n_real = cfg->n_real;
    p1.split_col = cfg->shared_intermediate_size;
    p1.mtiles = pv.mtiles;
    p1.pairs = pv.pairs;
    p1.n_pairs = pv.n_pairs;
    p1.row_scale = row_scale;
    const int n_shared_pairs = (int)((sz.t_pad / BM + 1) / 2);
    p1.p_begin = group_sel == 2 ? n_shared_pairs : 0;
    p1.p_end = group_sel == 1 ? n_shared_pairs : -1;
    p2 = p1;
    p2.n_tiles = (int)ceil_div(H, BN);
    p2.n_last = H - (p2.n_tiles - 1) * BN;
    p2.num_kb = Id / BK;
    p2.w_rows = H;
    if (p1.n_last != BN && (p1.n_last != 128)) {
        set_error("2-CTA GEMM-1: the last accumulator tile must be 128 or 256 columns wide (got %d)", p1.n_last);
        return DCMOE_ERR_UNSUPPORTED;
    }
    if (p2.n_last != BN) {
        set_error("2-CTA GEMM-2 needs hidden_size %% 256 == 0");
        return DCMOE_ERR_UNSUPPORTED;
    }
    int n_ctas = num_sms2() & ~1;
    if (max_ctas > 0 && max_ctas < n_ctas) n_ctas = max_ctas & ~1;
    if (n_ctas < 2) n_ctas = 2;
    dim3 grid((unsigned)n_ctas), block(NUM_THREADS);
    if (phase != 2) {
        ffn_gemm2cta_kernel<true><<<grid, block, SMEM_BYTES, stream>>>(m_x, m_xp, m_w13, m_w13h, m_h_st, p1);
        if ((rc = check_cuda(cudaGetLastError(), "ffn_gemm2cta_kernel<SwiGLU> launch"))) return rc;
    }
    if (phase != 1) ffn_gemm2cta_kernel<false><<<grid, block, SMEM_BYTES, stream>>>(m_h_ld, m_h_ld, m_w2, m_w2, m_y_st, p2);
    return check_cuda(cudaGetLastError(), "ffn_gemm2cta_kernel<down> launch");
}

}  // namespace dcmoe
