// probe_gather4.cu -- what does cp.async.bulk.tensor.2d ... tile::gather4 deliver on sm_100a, and which tensor-map box
// does it want?  Standalone:  nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_build/probe_gather4 tools/probe_gather4.cu -lcuda
// Tensor [64 rows, 256 cols] bf16 with value(row, col) = row * 256 + col (exact in fp32 after conversion of the index
// pattern: stored as uint16 row * 256 + col), 128B swizzle, box {64 cols, BOX_ROWS}; gathers rows {5, 17, 2, 40} at column 64.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, int r0, int r1, int r2, int r3, int col, uint32_t expect, uint16_t* out,
                      int* status) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    __shared__ uint64_t bar;
    uint32_t base = ((uint32_t)__cvta_generic_to_shared(sm_raw) + 1023u) & ~1023u;
    uint8_t* sm = sm_raw + (base - (uint32_t)__cvta_generic_to_shared(sm_raw));
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) ((uint16_t*)sm)[i] = 0xffff;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(expect));
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(base), "l"((uint64_t)&tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(b) : "memory");
        uint32_t done = 0, spins = 0;
        while (!done && spins < 2000000) {
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(b));
            ++spins;
        }
        *status = done;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = ((uint16_t*)sm)[i];
}

// ---- rate probe: how fast can one SM fill a 128-row x 128-byte A tile with 32 gather4 instructions (one per lane of a
// producer warp), compared with one 128-row box?  Ring of `stages` 16 KB stages, a consumer warp frees every stage as
// soon as it is full; each tile walks 32 K slices of its 128 rows (as the grouped GEMM does).
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
__global__ void __launch_bounds__(64) rate_kernel(const __grid_constant__ CUtensorMap tm_row, const __grid_constant__ CUtensorMap tm_box,
                                                  const int* __restrict__ rows, int n_tiles, int stages, int gather, long long* cycles, int n_rows) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    __shared__ uint64_t bars[32];
    const uint32_t base = ((uint32_t)__cvta_generic_to_shared(sm_raw) + 1023u) & ~1023u;
    const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(bars);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8 * s));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8 * (16 + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const long long t0 = clock64();
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int4 r = reinterpret_cast<const int4*>(rows)[tile * 32 + lane];
        for (int kb = 0; kb < 32; ++kb, ++it) {
            const int s = it % stages;
            const uint32_t ph = (uint32_t)(it / stages) & 1u;
            if (warp == 0) {
                while (!try_wait(b0 + 8 * (16 + s), ph ^ 1u)) {}
                if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0 + 8 * s), "r"(16384) : "memory");
                __syncwarp();
                const uint32_t dst = base + s * 16384;
                if (gather)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                                 ::"r"(dst + lane * 512), "l"((uint64_t)&tm_row), "r"(kb * 64), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w), "r"(b0 + 8 * s) : "memory");
                else if (lane == 0)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(dst), "l"((uint64_t)&tm_box), "r"(kb * 64), "r"((tile * 128) % n_rows), "r"(b0 + 8 * s) : "memory");
                __syncwarp();
            } else {
                while (!try_wait(b0 + 8 * s, ph)) {}
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0 + 8 * (16 + s)) : "memory");
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

static void rate_probe(int64_t R) {
    const int64_t C = 2048;                   // a [R, 2048] bf16 activation matrix (65536 rows = 256 MiB; 8192 = 32 MiB, L2 resident)
    uint16_t* x;
    cudaMalloc(&x, R * C * 2);
    cudaMemset(x, 0, R * C * 2);
    const int n_tiles = 148 * 8;
    std::vector<int> rows(n_tiles * 128);
    uint32_t seed = 12345u;
    for (auto& v : rows) { seed = seed * 1664525u + 1013904223u; v = (int)((seed >> 8) % R); }
    int* drows; cudaMalloc(&drows, rows.size() * 4);
    cudaMemcpy(drows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice);
    long long* dcyc; cudaMalloc(&dcyc, 148 * 8);
    CUtensorMap tm_row, tm_box;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t estr[2] = {1u, 1u};
    cuuint32_t box1[2] = {64u, 1u}, box128[2] = {64u, 128u};
    cuTensorMapEncodeTiled(&tm_row, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuTensorMapEncodeTiled(&tm_box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024);
    for (int gather : {0, 1}) {
        for (int stages : {2, 4, 8}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            rate_kernel<<<148, 64, 8 * 16384 + 1024>>>(tm_row, tm_box, drows, n_tiles, stages, gather, dcyc, (int)R);
            cudaEventRecord(e0);
            rate_kernel<<<148, 64, 8 * 16384 + 1024>>>(tm_row, tm_box, drows, n_tiles, stages, gather, dcyc, (int)R);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            std::vector<long long> cyc(148);
            cudaMemcpy(cyc.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
            const double bytes = (double)n_tiles * 32 * 16384;
            printf("[%lld source rows] %s, ring of %d x 16 KB: %s  %.3f ms  %.2f TB/s into shared memory, %.0f cycles per 16 KB stage per SM\n",
                   (long long)R, gather ? "32 x gather4 (random rows)" : "1 x 128-row box (contiguous)", stages, cudaGetErrorString(err), ms,
                   bytes / (ms * 1e-3) / 1e12, (double)mx / (8.0 * 32));
        }
    }
}

int main(int argc, char** argv) {
    cuInit(0);
    cudaSetDevice(0);
    if (argc > 1) { rate_probe(argc > 2 ? atoll(argv[2]) : 65536); return 0; }
    const int R = 64, C = 256;
    std::vector<uint16_t> h(R * C);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 256 + c);
    uint16_t *d, *dout; int* dst;
    cudaMalloc(&d, R * C * 2); cudaMalloc(&dout, 4096); cudaMalloc(&dst, 4);
    cudaMemcpy(d, h.data(), R * C * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    for (int box_rows : {1, 4}) {
        for (int sw : {0, 1}) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
            cuuint64_t strides[1] = {(cuuint64_t)C * 2};
            cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
            cuuint32_t estr[2] = {1u, 1u};
            CUresult cr = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                 sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) { printf("box_rows %d swizzle %d: encode failed %d\n", box_rows, sw, (int)cr); continue; }
            cudaMemset(dst, 0, 4);
            probe<<<1, 128, 8192>>>(tm, 5, 17, 2, 40, 64, 512u, dout, dst);
            cudaError_t e = cudaDeviceSynchronize();
            int st = -1; std::vector<uint16_t> o(2048);
            cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), dout, 4096, cudaMemcpyDeviceToHost);
            printf("box_rows %d swizzle %d: launch %s, barrier completed with 512 bytes: %d\n", box_rows, sw, cudaGetErrorString(e), st);
            if (e != cudaSuccess) return 1;
            for (int row = 0; row < 5; ++row) {
                printf("  smem row %d (128 B), 16-byte chunks hold source (row, col0):", row);
                for (int ch = 0; ch < 8; ++ch) {
                    uint16_t v = o[row * 64 + ch * 8];
                    if (v == 0xffff) printf(" [--]"); else printf(" [%d,%d]", v >> 8, v & 255);
                }
                printf("\n");
            }
        }
    }
    return 0;
}
