// probe_stream.cu -- HBM read bandwidth of the access patterns the DCMoE kernels use, measured on one B200.
// Standalone (no torch):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/probe_stream tools/probe_stream.cu
// A 1 GiB bf16 matrix [R, 2048] (4 KB rows, like x and W13) is read once per run:
//   ldg_row       one warp per row, 512 B contiguous per load instruction (GEMV style)
//   ldg_tile      one CTA per 256-row tile, 128-byte K slices (the order a TMA box fetches)
//   tma(b,k,s)    TMA boxes of b rows x 128 B, k consecutive K slices per ring stage, s stages
// Prints GB/s for each; used to choose the weight-streaming layout of the decode path and the router's x ring.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kCols = 2048;          // bf16 per row -> 4096 B
constexpr int kRowBytes = kCols * 2;

__device__ __forceinline__ uint4 ld_nc(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(1024) ldg_row_kernel(const uint8_t* base, int64_t rows, uint32_t* sink) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    uint32_t acc = 0;
    for (int64_t r = warp; r < rows; r += n_warps) {
        const uint8_t* p = base + r * kRowBytes + lane * 16;
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ld_nc(p + i * 512);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void __launch_bounds__(256) ldg_tile_kernel(const uint8_t* base, int64_t rows, uint32_t* sink) {
    const int t = threadIdx.x;
    uint32_t acc = 0;
    for (int64_t tile = blockIdx.x; tile * 256 < rows; tile += gridDim.x) {
        const uint8_t* p = base + (tile * 256 + (t >> 3)) * kRowBytes + (t & 7) * 16;
        for (int kb = 0; kb < 32; ++kb) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = ld_nc(p + (int64_t)j * 32 * kRowBytes + kb * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (++spins > 2000000u) { printf("probe: mbarrier timeout\n"); __trap(); }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// warp 0 = producer (lanes issue boxes in parallel), warp 1 = consumer (waits full, releases)
// mode 0: tensor boxes [box_rows x 128 B], kper consecutive K slices per stage; mode 1: 1-D bulk copies of
// kper*128 contiguous bytes per row (box_rows rows per stage)
__global__ void __launch_bounds__(64) tma_kernel(const __grid_constant__ CUtensorMap map, const uint8_t* base, int64_t rows,
                                                  int box_rows, int kper, int stages, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = (uint32_t)box_rows * 128u * kper;
    const uint32_t bars = sbase + stages * stage_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (stages + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_tiles = mode == 2 ? rows / (box_rows * kper) : rows / box_rows;
    const int ksteps = mode == 2 ? 32 : 32 / kper;
    int st = 0; uint32_t ph = 0;
    if (warp == 0) {
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(bars + 8 * (stages + st), ph ^ 1u);
                if (lane == 0) mbar_expect_tx(bars + 8 * st, stage_bytes);
                __syncwarp();
                if (mode == 0) {
                    if (lane < kper)
                        tma_load_2d(sbase + st * stage_bytes + lane * box_rows * 128, &map, (ks * kper + lane) * 64,
                                    (int)(tile * box_rows), bars + 8 * st);
                } else if (mode == 2) {   // kper row groups of box_rows rows, one K slice (the decode GEMM's B tile)
                    if (lane < kper)
                        tma_load_2d(sbase + st * stage_bytes + lane * box_rows * 128, &map, ks * 64,
                                    (int)((tile * kper + lane) * box_rows), bars + 8 * st);
                } else {
                    for (int r = lane; r < box_rows; r += 32)
                        bulk_load_1d(sbase + st * stage_bytes + r * kper * 128,
                                     base + (tile * box_rows + r) * kRowBytes + (int64_t)ks * kper * 128, kper * 128, bars + 8 * st);
                }
                __syncwarp();
                if (++st == stages) { st = 0; ph ^= 1u; }
            }
    } else {
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(bars + 8 * st, ph);
                if (lane == 0) mbar_arrive(bars + 8 * (stages + st));
                __syncwarp();
                if (++st == stages) { st = 0; ph ^= 1u; }
            }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int64_t rows = 262144;   // 1 GiB
    uint8_t* buf;
    uint32_t* sink;
    CK(cudaMalloc(&buf, rows * kRowBytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 1, rows * kRowBytes));
    void* fn_ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn_ptr, cudaEnableDefault, &qres));
    EncodeFn encode = (EncodeFn)fn_ptr;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto report = [&](const char* name, float ms) {
        printf("%-34s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, rows * (double)kRowBytes / (ms * 1e-3) / 1e9);
    };
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        for (int wpb : {8, 16, 32}) {
            ldg_row_kernel<<<sms * (32 / wpb) * 2, wpb * 32>>>(buf, rows, sink);
            CK(cudaEventRecord(e0));
            ldg_row_kernel<<<sms * (32 / wpb) * 2, wpb * 32>>>(buf, rows, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char nm[64]; snprintf(nm, 64, "ldg_row warps/cta=%d", wpb);
            if (rep) report(nm, ms);
        }
        for (int cpsm : {2, 4, 8}) {
            CK(cudaEventRecord(e0));
            ldg_tile_kernel<<<sms * cpsm, 256>>>(buf, rows, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char nm[64]; snprintf(nm, 64, "ldg_tile ctas/sm=%d", cpsm);
            if (rep) report(nm, ms);
        }
        struct Cfg { int box_rows, kper, stages, mode, cpsm; };
        const Cfg cfgs[] = {
            {256, 1, 4, 0, 1}, {256, 1, 6, 0, 1}, {128, 1, 10, 0, 1}, {32, 1, 32, 0, 1}, {16, 1, 32, 0, 1},
            {16, 32, 2, 0, 1}, {16, 32, 3, 0, 1}, {32, 4, 12, 0, 1}, {32, 8, 6, 0, 1}, {64, 4, 6, 0, 1}, {8, 32, 6, 0, 1},
            {128, 1, 5, 0, 2}, {32, 8, 3, 0, 2}, {16, 32, 1, 0, 2}, {16, 8, 4, 0, 3},
            {16, 32, 3, 1, 1}, {32, 8, 6, 1, 1}, {48, 32, 1, 1, 1}, {16, 32, 1, 1, 3},
            {16, 8, 10, 2, 1}, {16, 16, 5, 2, 1}, {16, 32, 3, 2, 1}, {16, 8, 4, 2, 3}, {32, 4, 10, 2, 1}, {64, 2, 10, 2, 1},
            {16, 8, 10, 0, 1}, {16, 4, 20, 0, 1}, {16, 2, 32, 0, 1},
        };
        for (const Cfg& c : cfgs) {
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)kCols, (cuuint64_t)rows};
            cuuint64_t strides[1] = {(cuuint64_t)kRowBytes};
            cuuint32_t box[2] = {64u, (cuuint32_t)c.box_rows};
            cuuint32_t estr[2] = {1u, 1u};
            CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            const int smem = c.stages * c.box_rows * 128 * c.kper + 16 * c.stages + 1024 + 64;
            if (smem > 227 * 1024 / c.cpsm) { printf("skip (smem %d)\n", smem); continue; }
            CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CK(cudaEventRecord(e0));
            tma_kernel<<<sms * c.cpsm, 64, smem>>>(map, buf, rows, c.box_rows, c.kper, c.stages, c.mode);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char nm[96];
            snprintf(nm, 96, "%s box=%dx128B k/stage=%d stages=%d cta/sm=%d", c.mode == 1 ? "bulk1d" : (c.mode == 2 ? "tma2d-rowgroups" : "tma2d"), c.box_rows, c.kper,
                     c.stages, c.cpsm);
            if (rep) report(nm, ms);
        }
    }
    return 0;
}
