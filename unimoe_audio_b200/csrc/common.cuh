// common.cuh -- shared device/host helpers for libdcmoe_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcmoe_b200.h"

namespace dcmoe {

constexpr int kRouterBlock = DCMOE_ROUTER_BLOCK;  // tokens per router CTA
constexpr int kTileM = DCMOE_TILE_M;              // rows per FFN m-tile
constexpr int kMaxDyn = 16;                       // n_real + n_null upper bound (half-warp per token)

// host-side error plumbing (api.cu)
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t err, const char* what);
int validate_config(const dcmoe_config* cfg);
// 2-D bf16 row-major tensor [rows, cols] -> CUtensorMap (void*: CUtensorMap*) with box [box_rows, 64 cols] and
// 128B swizzle (api.cu)
int make_tensor_map_bf16(void* map, const void* base, int64_t rows, int64_t cols, int box_rows);

// Programmatic dependent launch (decode-sized chain front_small -> GEMM-1 -> GEMM-2 -> combine): the next kernel's
// CTAs may become resident and run their prologue (barrier init, TMEM allocation, tensor-map prefetch) while the
// previous kernel is still running; every such kernel executes grid_dep_wait() before its first global-memory
// access, which returns once the preceding grid has completed and its writes are visible.  DCMOE_PDL=0 disables.
bool pdl_enabled();   // api.cu
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                 Args... args) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = grid;
    lc.blockDim = block;
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...);
}
#endif

// "done once per device" flags for per-device function attributes (cudaFuncSetAttribute is per device; a process may
// drive several GPUs).  Returns true the first time it is called for the current device with this flag set.
struct PerDeviceOnce {
    bool done[64] = {};
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return true;
        const bool f = !done[dev];
        done[dev] = true;
        return f;
    }
    void reset_current() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64) done[dev] = false;
    }
};

// SM count of the CURRENT device (cached per device: a process may drive several GPUs) -- api.cu
int device_sm_count();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// plan accessors
struct PlanView {
    int32_t* block_counts;
    float* block_probs;
    int32_t* block_offsets;
    int32_t* counts;
    int32_t* seg_base;
    int32_t* n_mtiles;
    float* aux_loss;
    dcmoe_mtile* mtiles;
    int32_t* overflow;
    int32_t* small_tokens;
};

inline PlanView plan_view(void* plan, const dcmoe_plan_layout& l) {
    char* p = static_cast<char*>(plan);
    PlanView v;
    v.block_counts = reinterpret_cast<int32_t*>(p + l.block_counts);
    v.block_probs = reinterpret_cast<float*>(p + l.block_probs);
    v.block_offsets = reinterpret_cast<int32_t*>(p + l.block_offsets);
    v.counts = reinterpret_cast<int32_t*>(p + l.counts);
    v.seg_base = reinterpret_cast<int32_t*>(p + l.seg_base);
    v.n_mtiles = reinterpret_cast<int32_t*>(p + l.n_mtiles);
    v.aux_loss = reinterpret_cast<float*>(p + l.aux_loss);
    v.mtiles = reinterpret_cast<dcmoe_mtile*>(p + l.mtiles);
    v.overflow = reinterpret_cast<int32_t*>(p + l.overflow);
    v.small_tokens = reinterpret_cast<int32_t*>(p + l.small_tokens);
    return v;
}

// ---- device helpers ----
#ifdef __CUDACC__

__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// c10::BFloat16 rounding (round-to-nearest-even) on an fp32 value, kept in fp32
__device__ __forceinline__ float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_ca_v4(const void* p) {  // cacheable (L1-resident small operands)
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

#endif  // __CUDACC__

}  // namespace dcmoe
