// ep.cu -- expert-parallel dispatch / combine over NVLink peer memory (sm_100a).
//
// Reference semantics: AudioMOELayer.forward with an expert-parallel group (core.py:446-493): every rank
// routes its own tokens, rank r owns routed experts [r*n_loc, (r+1)*n_loc), two all_to_all_single calls on
// buffers padded to the GLOBAL-MAX capacity (core.py:455-457, :467, :480) carry tokens to the owners and
// results back.  Here the exchange is fused into the permute and combine kernels:
//   * dispatch: the permute kernel stores each selected row DIRECTLY into the owner's packed buffer
//     (st.global on a cudaIpc-mapped peer pointer, 128-bit, one NVLink write per row per expert) at the
//     row-space slot computed from the all-gathered per-rank counts -- only real rows travel, nothing is
//     padded to a capacity, and the owner's buffer ends up in the canonical order (rank-major, then
//     ascending token id) with no receive-side regrouping;
//   * combine: the combine kernel gathers a token's <= n_real routed rows from the owners' y buffers with
//     128-bit peer loads (ld.global.nc over NVLink), adds the local shared-expert row in fp32, one store.
// The only host-visible collectives are an all-gather of (n_real + 1) int32 per rank and two 4-byte
// all-reduces used as stream-ordered barriers (NCCL, issued from Python).
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace dcmoe {

constexpr int kMaxRanks = 8;

struct EpPeers {
    char* x_packed[kMaxRanks];
    float* row_scale[kMaxRanks];
    const char* y[kMaxRanks];
    const float* aux_src;   // single-GPU combine: optional 4-byte copy of the plan's aux loss into a per-call output
    float* aux_dst;
};

// ep_meta (device int32): [0,16) dest_base[e]  row-space row ON THE OWNER where this rank's rows of expert e start
//                         [16,32) dest_tpad[e]  t_pad of the owner of expert e
static_assert(DCMOE_EP_META_INTS == 32, "ep_meta layout");

namespace {

constexpr unsigned kFull = 0xffffffffu;

// all_counts: [world][n_real + 1] int32 = per-rank routed counts per (global) expert, then that rank's token count
__global__ void __launch_bounds__(256) ep_plan_kernel(const int32_t* __restrict__ all_counts, int rank, int world,
                                                      int n_real, int n_loc, int64_t T, int max_mtiles, PlanView pv,
                                                      int32_t* __restrict__ ep_meta) {
    __shared__ int s_total[kMaxDyn];     // global rows per expert
    __shared__ int s_before[kMaxDyn];    // rows of expert e coming from ranks < rank
    __shared__ int s_seg[kMaxDyn + 1];   // local row-space segment bases (local experts)
    __shared__ int s_tile0[kMaxDyn + 1];
    const int stride = n_real + 1;
    const int t_pad = (int)((T + kTileM - 1) / kTileM) * kTileM;
    if (threadIdx.x < n_real) {
        const int e = threadIdx.x;
        int tot = 0, before = 0;
        for (int r = 0; r < world; ++r) {
            const int c = all_counts[r * stride + e];
            if (r < rank) before += c;
            tot += c;
        }
        s_total[e] = tot;
        s_before[e] = before;
    }
    __syncthreads();
    if (threadIdx.x < n_real) {
        // destination of MY rows of expert e on its owner
        const int e = threadIdx.x;
        const int owner = e / n_loc;
        const int owner_T = all_counts[owner * stride + n_real];
        const int owner_tpad = (owner_T + kTileM - 1) / kTileM * kTileM;
        int base = owner_tpad;
        for (int q = owner * n_loc; q < e; ++q) base += (s_total[q] + kTileM - 1) / kTileM * kTileM;
        ep_meta[e] = base + s_before[e];
        ep_meta[16 + e] = owner_tpad;
    }
    if (threadIdx.x == 0) {
        int row = t_pad, tile = t_pad / kTileM;
        for (int l = 0; l < n_loc; ++l) {
            const int e = rank * n_loc + l;
            s_seg[l] = row;
            s_tile0[l] = tile;
            pv.seg_base[l] = row;
            pv.counts[l] = s_total[e];
            const int nt = (s_total[e] + kTileM - 1) / kTileM;
            row += nt * kTileM;
            tile += nt;
        }
        s_seg[n_loc] = row;
        s_tile0[n_loc] = tile;
        pv.seg_base[n_loc] = row;
        *pv.n_mtiles = tile < max_mtiles ? tile : max_mtiles;
        *pv.overflow = tile > max_mtiles ? 1 : 0;   // only possible when the caller chose a row_capacity below the worst case
    }
    __syncthreads();
    const int n_shared_tiles = t_pad / kTileM;
    const int total = s_tile0[n_loc];
    for (int i = threadIdx.x; i < total && i < max_mtiles; i += blockDim.x) {
        dcmoe_mtile mt;
        if (i < n_shared_tiles) {
            mt.a_row = i * kTileM;
            mt.out_row = i * kTileM;
            mt.group = n_loc;
            const int64_t left = T - (int64_t)i * kTileM;
            mt.rows = (int)(left < kTileM ? left : kTileM);
        } else {
            int l = 0;
            while (l + 1 < n_loc && i >= s_tile0[l + 1]) ++l;
            const int local = i - s_tile0[l];
            mt.out_row = s_seg[l] + local * kTileM;
            mt.a_row = mt.out_row - t_pad;
            mt.group = l;
            const int left = s_total[rank * n_loc + l] - local * kTileM;
            mt.rows = left < kTileM ? left : kTileM;
        }
        pv.mtiles[i] = mt;
    }
}

// shared-expert row scales (gw[t, n_dyn], gw[t, n_dyn + 1]) of the local rows [0, T): written here, before the
// dispatch, so that the shared experts' GEMM-1 can run while the dispatch is still in flight
template <int ESIZE>
__global__ void ep_shared_scale_kernel(const char* __restrict__ gw, int64_t T, int n_dyn, int n_fix,
                                       float* __restrict__ row_scale) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int E = n_dyn + n_fix;
    auto load_gw = [&](int j) -> float {
        if (ESIZE == 2) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gw)[t * E + j]);
        return reinterpret_cast<const float*>(gw)[t * E + j];
    };
    row_scale[2 * t] = load_gw(n_dyn);
    row_scale[2 * t + 1] = n_fix > 1 ? load_gw(n_dyn + 1) : 0.0f;
}

template <int ESIZE>
__global__ void __launch_bounds__(128) ep_dispatch_kernel(const char* __restrict__ x, const int32_t* __restrict__ mask,
                                                          const char* __restrict__ gw, int64_t T, int H, int n_real,
                                                          int n_dyn, int n_fix, int n_loc, int rank,
                                                          const int32_t* __restrict__ block_offsets,
                                                          const int32_t* __restrict__ ep_meta, EpPeers peers,
                                                          int32_t* __restrict__ slot_of, int n_blocks, int row_limit) {
    __shared__ int s_m[kRouterBlock][kMaxDyn];
    __shared__ int s_slot[kRouterBlock][kMaxDyn];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = n_dyn + n_fix;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {   // grid may be capped (comm under compute)
    __syncthreads();
    const int64_t tok0 = (int64_t)blk * kRouterBlock;
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
        int rk = 0;
        for (int q = 0; q < tl; ++q) rk += s_m[q][e];
        int slot = -1;
        if (s_m[tl][e]) {
            slot = ep_meta[e] + block_offsets[(int64_t)blk * n_real + e] + rk;   // row on the owner
            if (slot >= row_limit) {
                slot = -1;     // beyond the owner's row capacity (every rank uses the same): dropped, plan.overflow = 1 there
            } else {
                const float w = load_gw(t, e);
                float* sc = peers.row_scale[e / n_loc];
                sc[2 * (int64_t)slot] = w;
                sc[2 * (int64_t)slot + 1] = w;
            }
        }
        s_slot[tl][e] = slot;
        if (t < T) slot_of[t * n_real + e] = slot;
    }
    (void)rank;
    __syncthreads();
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
                uint4* dst = reinterpret_cast<uint4*>(peers.x_packed[e / n_loc] +
                                                      (int64_t)(slot - ep_meta[16 + e]) * H * ESIZE);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = c0 + u * 32 + lane;
                    if (c < n_vec) st_na_v4(dst + c, v[u]);   // local HBM or NVLink peer write
                }
            }
        }
    }
    }  // block loop
}

// MODE 0: out = D(sum of routed rows (expert order) + shared row)          -- the whole combine
// MODE 1: partial[t] = fp32 sum of routed rows                              -- overlappable with the shared GEMM-2
// MODE 2: out = D(partial[t] + shared row)
// MODE 0 and MODE 1 + MODE 2 perform the same fp32 additions in the same order (bitwise equal outputs).
template <bool BF16, int MODE>
__global__ void __launch_bounds__(256, 3) ep_combine_kernel(const char* __restrict__ y_local, EpPeers peers,
                                                            const int32_t* __restrict__ slot_of, int64_t T, int H,
                                                            int n_real, int n_loc, float* __restrict__ partial,
                                                            const char* __restrict__ residual, char* __restrict__ out) {
    constexpr int ESIZE = BF16 ? 2 : 4;
    constexpr int PER = 16 / ESIZE;
    constexpr int U = 4;             // 128-bit vectors per lane per pass (x 2 source rows in flight)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_vec = H * ESIZE / 16;
    const int64_t row_bytes = (int64_t)H * ESIZE;
    grid_dep_wait();   // (decode-sized calls launch this kernel programmatically dependent on GEMM-2; else a no-op)
    if (MODE == 0 && peers.aux_dst != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *peers.aux_dst = *peers.aux_src;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < T; t += (int64_t)gridDim.x * 8) {   // grid may be capped
    // compact source list, in accumulation order: lane i < n_src holds the base pointer of source row i
    // (selected routed rows in expert order, then the shared row)
    const char* my_src = nullptr;
    int n_src = 0;
    {
        int slot = -1;
        if (MODE != 2 && lane < n_real) slot = slot_of[t * n_real + lane];
        const unsigned sel = __ballot_sync(kFull, slot >= 0);
        const char* p = slot >= 0 ? peers.y[lane / n_loc] + (int64_t)slot * row_bytes : nullptr;   // local or NVLink peer
        const int pos = __popc(sel & ((1u << lane) - 1u));
        n_src = __popc(sel);
        // scatter lane -> position pos: every destination lane i pulls from the i-th set bit of sel
        int srcl = 0;
        {
            unsigned m = sel;
            for (int i = 0; i < lane && m; ++i) m &= m - 1;          // drop the `lane` lowest set bits
            srcl = m ? __ffs(m) - 1 : 0;
        }
        (void)pos;
        const unsigned long long pv = __shfl_sync(kFull, (unsigned long long)p, srcl);
        my_src = lane < n_src ? (const char*)pv : nullptr;
        if (MODE != 1) {                                             // the shared row comes last
            if (lane == n_src) my_src = y_local + t * row_bytes;
            n_src += 1;
        }
    }
    auto add = [&](float (&acc)[U][PER], const uint4 (&v)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (BF16) {
                acc[u][0] += bf16lo(v[u].x); acc[u][1] += bf16hi(v[u].x);
                acc[u][2 % PER] += bf16lo(v[u].y); acc[u][3 % PER] += bf16hi(v[u].y);
                acc[u][4 % PER] += bf16lo(v[u].z); acc[u][5 % PER] += bf16hi(v[u].z);
                acc[u][6 % PER] += bf16lo(v[u].w); acc[u][7 % PER] += bf16hi(v[u].w);
            } else {
                acc[u][0] += __uint_as_float(v[u].x); acc[u][1] += __uint_as_float(v[u].y);
                acc[u][2] += __uint_as_float(v[u].z); acc[u][3] += __uint_as_float(v[u].w);
            }
        }
    };
    for (int c0 = 0; c0 < n_vec; c0 += 32 * U) {
        float acc[U][PER];
        if (MODE == 2) {
            // resume from the fp32 partial sums: element k of vector c lives at partial[t*H + c*PER + k]
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u * 32 + lane;
#pragma unroll
                for (int q = 0; q < PER / 4; ++q) {
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c < n_vec) f = *reinterpret_cast<const float4*>(partial + t * (int64_t)H + (int64_t)c * PER + 4 * q);
                    acc[u][4 * q] = f.x; acc[u][4 * q + 1] = f.y; acc[u][4 * q + 2] = f.z; acc[u][4 * q + 3] = f.w;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int i = 0; i < PER; ++i) acc[u][i] = 0.0f;
        }
        // two source rows (2 x U x 128-bit per lane) in flight; accumulation stays in list order
        for (int i = 0; i < n_src; i += 2) {
            const uint4* s0 = reinterpret_cast<const uint4*>((const char*)__shfl_sync(kFull, (unsigned long long)my_src, i));
            const bool two = i + 1 < n_src;
            const uint4* s1 = reinterpret_cast<const uint4*>((const char*)__shfl_sync(kFull, (unsigned long long)my_src, two ? i + 1 : i));
            uint4 v0[U], v1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u * 32 + lane;
                v0[u] = c < n_vec ? ld_nc_v4(s0 + c) : make_uint4(0, 0, 0, 0);
            }
            if (two) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u * 32 + lane;
                    v1[u] = c < n_vec ? ld_nc_v4(s1 + c) : make_uint4(0, 0, 0, 0);
                }
            }
            add(acc, v0);
            if (two) add(acc, v1);
        }
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u * 32 + lane;
                if (c >= n_vec) continue;
#pragma unroll
                for (int q = 0; q < PER / 4; ++q)
                    *reinterpret_cast<float4*>(partial + t * (int64_t)H + (int64_t)c * PER + 4 * q) =
                        make_float4(acc[u][4 * q], acc[u][4 * q + 1], acc[u][4 * q + 2], acc[u][4 * q + 3]);
            }
        } else {
            if (residual != nullptr) {                 // fused residual add of the decoder layer (model.py:242), added last
                uint4 rv[U];
                const uint4* rs = reinterpret_cast<const uint4*>(residual + t * row_bytes);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u * 32 + lane;
                    rv[u] = c < n_vec ? ld_nc_v4(rs + c) : make_uint4(0, 0, 0, 0);
                }
                add(acc, rv);
            }
            uint4* dst = reinterpret_cast<uint4*>(out + t * row_bytes);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u * 32 + lane;
                if (c >= n_vec) continue;
                uint4 o;
                if (BF16) {
                    o.x = pack_bf16(acc[u][0], acc[u][1]);
                    o.y = pack_bf16(acc[u][2 % PER], acc[u][3 % PER]);
                    o.z = pack_bf16(acc[u][4 % PER], acc[u][5 % PER]);
                    o.w = pack_bf16(acc[u][6 % PER], acc[u][7 % PER]);
                } else {
                    o.x = __float_as_uint(acc[u][0]); o.y = __float_as_uint(acc[u][1]);
                    o.z = __float_as_uint(acc[u][2]); o.w = __float_as_uint(acc[u][3]);
                }
                st_na_v4(dst + c, o);
            }
        }
    }
    }  // token loop
}

// Cross-GPU barrier over peer memory, optionally carrying a payload (replaces the NCCL collectives of the
// expert-parallel paths: the all-gather of the per-rank counts, the all-gather of decode-sized token rows and the
// 4-byte all-reduces used as stream-ordered barriers).
// Every rank owns a flag array int32[DCMOE_EP_FLAG_SLOTS][DCMOE_MAX_RANKS] mapped into all ranks.  CTA r of rank q
// copies the payload (if any) into rank r's buffer at offset q * payload_bytes, publishes `epoch` into rank r's
// flags[slot][q] (fence.sc.sys + st.release.sys: the payload AND everything this rank's EARLIER kernels on the stream
// wrote -- to local or peer memory -- is visible to whoever acquires the flag), then waits until its own
// flags[slot][r] has reached `epoch` (ld.acquire.sys).  Kernels launched after this one on the stream therefore see
// every rank's payload and every rank's writes from before its barrier call.  The spin is bounded (trap, no hang).
struct EpFlagPeers {
    int32_t* flags[kMaxRanks];
    char* payload_dst[kMaxRanks];
};

__global__ void __launch_bounds__(256) ep_barrier_kernel(EpFlagPeers peers, int rank, int world, int slot, int32_t epoch,
                                                         const char* __restrict__ payload, int64_t payload_bytes) {
    const int r = blockIdx.x;   // peer this CTA talks to
    if (payload != nullptr && payload_bytes > 0) {
        char* dst = peers.payload_dst[r] + (int64_t)rank * payload_bytes;
        if (((payload_bytes | (int64_t)(uintptr_t)dst | (int64_t)(uintptr_t)payload) & 15) == 0) {
            const uint4* s4 = reinterpret_cast<const uint4*>(payload);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            for (int64_t i = threadIdx.x; i < payload_bytes / 16; i += blockDim.x) d4[i] = s4[i];
        } else {
            const int32_t* s1 = reinterpret_cast<const int32_t*>(payload);
            int32_t* d1 = reinterpret_cast<int32_t*>(dst);
            for (int64_t i = threadIdx.x; i < payload_bytes / 4; i += blockDim.x) d1[i] = s1[i];
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    asm volatile("fence.sc.sys;" ::: "memory");
    int32_t* dstf = peers.flags[r] + slot * kMaxRanks + rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dstf), "r"(epoch) : "memory");
    const int32_t* src = peers.flags[rank] + slot * kMaxRanks + r;
    const long long t0 = clock64();
    for (;;) {
        int32_t v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        __nanosleep(64);
        if (clock64() - t0 > 40000000000ll) {   // ~20 s: a rank that never arrives must not hang the GPU
            printf("dcmoe: expert-parallel barrier timed out (rank %d waiting for rank %d, slot %d, epoch %d, saw %d)\n",
                   rank, r, slot, epoch, v);
            __trap();
        }
    }
}

}  // namespace

// single-GPU combine = the expert-parallel combine with one rank (peers.y[0] = y)
int launch_combine(const void* y, const int32_t* slot_of, int64_t T, const dcmoe_config* cfg, const void* residual,
                   void* out, const float* aux_src, float* aux_dst, cudaStream_t stream) {
    if (T == 0) {
        if (aux_dst) return check_cuda(cudaMemcpyAsync(aux_dst, aux_src, 4, cudaMemcpyDeviceToDevice, stream), "aux copy");
        return DCMOE_OK;
    }
    EpPeers peers{};
    peers.y[0] = (const char*)y;
    peers.aux_src = aux_src;
    peers.aux_dst = aux_dst;
    dim3 grid((unsigned)ceil_div(T, 8)), block(256);
    const bool pdl = pdl_enabled() && T <= 64;   // decode-sized chain: come up under the tail of GEMM-2
    if (cfg->dtype == DCMOE_BF16)
        return check_cuda(launch_kernel(ep_combine_kernel<true, 0>, grid, block, 0, stream, pdl, (const char*)y, peers, slot_of, T,
                                        cfg->hidden_size, cfg->n_real, cfg->n_real, (float*)nullptr, (const char*)residual,
                                        (char*)out),
                          "combine kernel launch");
    return check_cuda(launch_kernel(ep_combine_kernel<false, 0>, grid, block, 0, stream, pdl, (const char*)y, peers, slot_of, T,
                                    cfg->hidden_size, cfg->n_real, cfg->n_real, (float*)nullptr, (const char*)residual,
                                    (char*)out),
                      "combine kernel launch");
}

}  // namespace dcmoe

using namespace dcmoe;

// defined in api.cu
namespace dcmoe {
int ep_plan_view(const dcmoe_config* cfg, int64_t T, int64_t row_capacity, void* plan, dcmoe_sizes* sz, PlanView* pv);
}

extern "C" {

int dcmoe_ipc_alloc(int64_t bytes, void** ptr) {
    if (!ptr || bytes <= 0) { set_error("dcmoe_ipc_alloc: bad arguments"); return DCMOE_ERR_INVALID; }
    return check_cuda(cudaMalloc(ptr, (size_t)bytes), "cudaMalloc (ipc buffer)");
}
int dcmoe_ipc_free(void* ptr) { return check_cuda(cudaFree(ptr), "cudaFree (ipc buffer)"); }
int dcmoe_ipc_export(const void* ptr, uint8_t* handle64) {
    if (!ptr || !handle64) { set_error("dcmoe_ipc_export: NULL argument"); return DCMOE_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    int rc = check_cuda(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)), "cudaIpcGetMemHandle");
    if (rc) return rc;
    memcpy(handle64, &h, 64);
    return DCMOE_OK;
}
int dcmoe_ipc_import(const uint8_t* handle64, void** ptr) {
    if (!ptr || !handle64) { set_error("dcmoe_ipc_import: NULL argument"); return DCMOE_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return check_cuda(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
int dcmoe_ipc_close(void* ptr) { return check_cuda(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle"); }

int dcmoe_ep_plan(const int32_t* all_counts, int rank, int world, int64_t T, int64_t row_capacity,
                  const dcmoe_config* cfg, void* plan, int32_t* ep_meta, const void* global_weight,
                  float* row_scale_local, void* stream) {
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (!all_counts || !plan || !ep_meta || !global_weight || !row_scale_local || world < 1 || world > kMaxRanks ||
        rank < 0 || rank >= world || cfg->n_real % world != 0) {
        set_error("dcmoe_ep_plan: bad arguments (world=%d rank=%d n_real=%d)", world, rank, cfg->n_real);
        return DCMOE_ERR_INVALID;
    }
    dcmoe_sizes sz; PlanView pv;
    if ((rc = ep_plan_view(cfg, T, row_capacity, plan, &sz, &pv))) return rc;
    ep_plan_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(all_counts, rank, world, cfg->n_real, cfg->n_real / world, T,
                                                        (int)sz.max_mtiles, pv, ep_meta);
    if ((rc = check_cuda(cudaGetLastError(), "ep_plan_kernel launch"))) return rc;
    if (T > 0) {
        const int n_dyn = cfg->n_real + cfg->n_null;
        dim3 grid((unsigned)ceil_div(T, 256)), block(256);
        if (cfg->dtype == DCMOE_BF16)
            ep_shared_scale_kernel<2><<<grid, block, 0, (cudaStream_t)stream>>>((const char*)global_weight, T, n_dyn, cfg->n_fix, row_scale_local);
        else
            ep_shared_scale_kernel<4><<<grid, block, 0, (cudaStream_t)stream>>>((const char*)global_weight, T, n_dyn, cfg->n_fix, row_scale_local);
        rc = check_cuda(cudaGetLastError(), "ep_shared_scale_kernel launch");
    }
    return rc;
}

int dcmoe_ep_dispatch(const void* x, const int32_t* expert_mask, const void* global_weight, int64_t T,
                      int64_t row_capacity, const dcmoe_config* cfg, const void* plan, const int32_t* ep_meta,
                      int rank, int world, void* const* peer_x_packed, float* const* peer_row_scale, int32_t* slot_of,
                      int max_ctas, void* stream) {
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (T == 0) return DCMOE_OK;
    if (!x || !expert_mask || !global_weight || !plan || !ep_meta || !peer_x_packed || !peer_row_scale || !slot_of ||
        world < 1 || world > kMaxRanks || cfg->n_real % world != 0) {
        set_error("dcmoe_ep_dispatch: bad arguments");
        return DCMOE_ERR_INVALID;
    }
    dcmoe_sizes sz; PlanView pv;
    if ((rc = ep_plan_view(cfg, T, row_capacity, const_cast<void*>(plan), &sz, &pv))) return rc;
    EpPeers peers{};
    for (int r = 0; r < world; ++r) { peers.x_packed[r] = (char*)peer_x_packed[r]; peers.row_scale[r] = peer_row_scale[r]; }
    const int n_dyn = cfg->n_real + cfg->n_null;
    int64_t nb = sz.n_blocks;
    if (max_ctas > 0 && nb > max_ctas) nb = max_ctas;
    dim3 grid((unsigned)nb), block(128);
    if (cfg->dtype == DCMOE_BF16)
        ep_dispatch_kernel<2><<<grid, block, 0, (cudaStream_t)stream>>>((const char*)x, expert_mask, (const char*)global_weight,
            T, cfg->hidden_size, cfg->n_real, n_dyn, cfg->n_fix, cfg->n_real / world, rank, pv.block_offsets, ep_meta, peers, slot_of, (int)sz.n_blocks, (int)(sz.max_mtiles * kTileM));
    else
        ep_dispatch_kernel<4><<<grid, block, 0, (cudaStream_t)stream>>>((const char*)x, expert_mask, (const char*)global_weight,
            T, cfg->hidden_size, cfg->n_real, n_dyn, cfg->n_fix, cfg->n_real / world, rank, pv.block_offsets, ep_meta, peers, slot_of, (int)sz.n_blocks, (int)(sz.max_mtiles * kTileM));
    return check_cuda(cudaGetLastError(), "ep_dispatch_kernel launch");
}

int dcmoe_ep_combine(const void* y_local, const void* const* peer_y, const int32_t* slot_of, int64_t T,
                     const dcmoe_config* cfg, int world, int mode, float* partial, void* out, int max_ctas,
                     void* stream) {
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (T == 0) return DCMOE_OK;
    if (!y_local || !peer_y || !slot_of || world < 1 || world > kMaxRanks || cfg->n_real % world != 0 || mode < 0 ||
        mode > 2 || (mode != 1 && !out) || (mode != 0 && !partial)) {
        set_error("dcmoe_ep_combine: bad arguments");
        return DCMOE_ERR_INVALID;
    }
    EpPeers peers{};
    for (int r = 0; r < world; ++r) peers.y[r] = (const char*)peer_y[r];
    int64_t nb = ceil_div(T, 8);
    if (max_ctas > 0 && nb > max_ctas) nb = max_ctas;
    dim3 grid((unsigned)nb), block(256);
    const bool bf16 = cfg->dtype == DCMOE_BF16;
#define DCMOE_EP_COMBINE(BF, MODE_)                                                                                   \
    ep_combine_kernel<BF, MODE_><<<grid, block, 0, (cudaStream_t)stream>>>((const char*)y_local, peers, slot_of, T,  \
        cfg->hidden_size, cfg->n_real, cfg->n_real / world, partial, nullptr, (char*)out)
    if (mode == 0) { if (bf16) DCMOE_EP_COMBINE(true, 0); else DCMOE_EP_COMBINE(false, 0); }
    else if (mode == 1) { if (bf16) DCMOE_EP_COMBINE(true, 1); else DCMOE_EP_COMBINE(false, 1); }
    else { if (bf16) DCMOE_EP_COMBINE(true, 2); else DCMOE_EP_COMBINE(false, 2); }
#undef DCMOE_EP_COMBINE
    return check_cuda(cudaGetLastError(), "ep_combine_kernel launch");
}

int dcmoe_ep_barrier(int32_t* const* peer_flags, int rank, int world, int slot, int32_t epoch, const void* payload,
                     int64_t payload_bytes, void* const* peer_payload_dst, void* stream) {
    if (!peer_flags || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || slot < 0 || slot >= DCMOE_EP_FLAG_SLOTS ||
        payload_bytes < 0 || (payload_bytes & 3) != 0 || (payload_bytes > 0 && (!payload || !peer_payload_dst))) {
        set_error("dcmoe_ep_barrier: bad arguments (world=%d rank=%d slot=%d payload_bytes=%lld)", world, rank, slot,
                  (long long)payload_bytes);
        return DCMOE_ERR_INVALID;
    }
    EpFlagPeers peers{};
    for (int r = 0; r < world; ++r) {
        if (!peer_flags[r] || (payload_bytes > 0 && !peer_payload_dst[r])) {
            set_error("dcmoe_ep_barrier: NULL pointer for rank %d", r);
            return DCMOE_ERR_INVALID;
        }
        peers.flags[r] = peer_flags[r];
        peers.payload_dst[r] = payload_bytes > 0 ? (char*)peer_payload_dst[r] : nullptr;
    }
    ep_barrier_kernel<<<world, 256, 0, (cudaStream_t)stream>>>(peers, rank, world, slot, epoch,
                                                               payload_bytes > 0 ? (const char*)payload : nullptr, payload_bytes);
    return check_cuda(cudaGetLastError(), "ep_barrier_kernel launch");
}

int dcmoe_ep_fetch_weights(const void* const* peer_w13, const void* const* peer_w2, int rank, int world,
                           const dcmoe_config* cfg, void* w13_full, void* w2_full, void* stream) {
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (!peer_w13 || !peer_w2 || !w13_full || !w2_full || world < 1 || world > kMaxRanks || rank < 0 || rank >= world ||
        cfg->n_real % world != 0) {
        set_error("dcmoe_ep_fetch_weights: bad arguments (world=%d rank=%d n_real=%d)", world, rank, cfg->n_real);
        return DCMOE_ERR_INVALID;
    }
    const int64_t es = cfg->dtype == DCMOE_BF16 ? 2 : 4;
    const int n_loc = cfg->n_real / world;
    const int64_t g13 = 2ll * cfg->dynamic_intermediate_size * cfg->hidden_size * es;   // bytes of one weight group of W13
    const int64_t g2 = (int64_t)cfg->hidden_size * cfg->dynamic_intermediate_size * es; // ... of W2
    cudaStream_t st = (cudaStream_t)stream;
    auto copy = [&](void* dst, const void* src, int64_t bytes, const char* what) {
        return check_cuda(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, st), what);
    };
    // the shared pack (group n_loc of every rank's pack; taken from this rank's own) goes first: the shared experts'
    // row tiles head the tile list
    if ((rc = copy((char*)w13_full + cfg->n_real * g13, (const char*)peer_w13[rank] + n_loc * g13, g13, "copy of the shared W13"))) return rc;
    if ((rc = copy((char*)w2_full + cfg->n_real * g2, (const char*)peer_w2[rank] + n_loc * g2, g2, "copy of the shared W2"))) return rc;
    // then the ranks' routed experts, starting with this rank's own and walking the ring: at any moment every rank
    // is read by ONE other rank, so no GPU's NVLink egress is shared between readers
    for (int i = 0; i < world; ++i) {
        const int q = (rank + i) % world;
        if (!peer_w13[q] || !peer_w2[q]) { set_error("dcmoe_ep_fetch_weights: NULL weight pointer for rank %d", q); return DCMOE_ERR_INVALID; }
        if ((rc = copy((char*)w13_full + (int64_t)q * n_loc * g13, peer_w13[q], n_loc * g13, "peer copy of W13"))) return rc;
        if ((rc = copy((char*)w2_full + (int64_t)q * n_loc * g2, peer_w2[q], n_loc * g2, "peer copy of W2"))) return rc;
    }
    return DCMOE_OK;
}

}  // extern "C"
