// rmsnorm.cu -- the decoder layer's post_attention_layernorm in front of the MoE block.
//
// Reference: utils/UniMoE_Audio_model.py:239-240 (`residual = hidden_states; hidden_states =
// self.post_attention_layernorm(hidden_states)`), a transformers Qwen2RMSNorm (model.py:207, eps = rms_norm_eps):
//     v = mean(float(x)^2, -1);  n = D(float(x) * rsqrt(v + eps));  y = D(weight * n)         (D = layer dtype)
// One warp per token row; the row is read from HBM once (it stays in registers for H <= 2048 bf16 / 1024 fp32,
// otherwise the second pass re-reads it from L1), the sum of squares is reduced in fp32 (lane-strided partial sums,
// xor-shuffle tree), and the row is written once.  HBM-bound: 2 x T x H x sizeof(D) bytes.
#include <algorithm>

#include "common.cuh"

namespace dcmoe {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kRegVecs = 8;   // 16-byte vectors per lane kept in registers

template <bool BF16>
__device__ __forceinline__ float sumsq_vec(const uint4& v) {
    if (BF16) {
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = bf16lo(u[i]), b = bf16hi(u[i]);
            s = fmaf(a, a, s);
            s = fmaf(b, b, s);
        }
        return s;
    }
    const float f[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
    return fmaf(f[0], f[0], fmaf(f[1], f[1], fmaf(f[2], f[2], f[3] * f[3])));
}

template <bool BF16>
__device__ __forceinline__ uint4 norm_vec(const uint4& v, const uint4& w, float inv) {
    uint4 o;
    if (BF16) {
        const uint32_t xs[4] = {v.x, v.y, v.z, v.w}, ws[4] = {w.x, w.y, w.z, w.w};
        uint32_t r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // D(x * inv) then D(weight * .): both roundings of the reference
            const float n0 = bf16_round(__fmul_rn(bf16lo(xs[i]), inv)), n1 = bf16_round(__fmul_rn(bf16hi(xs[i]), inv));
            r[i] = pack_bf16(__fmul_rn(bf16lo(ws[i]), n0), __fmul_rn(bf16hi(ws[i]), n1));
        }
        o = make_uint4(r[0], r[1], r[2], r[3]);
    } else {
        o.x = __float_as_uint(__fmul_rn(__uint_as_float(w.x), __fmul_rn(__uint_as_float(v.x), inv)));
        o.y = __float_as_uint(__fmul_rn(__uint_as_float(w.y), __fmul_rn(__uint_as_float(v.y), inv)));
        o.z = __float_as_uint(__fmul_rn(__uint_as_float(w.z), __fmul_rn(__uint_as_float(v.z), inv)));
        o.w = __float_as_uint(__fmul_rn(__uint_as_float(w.w), __fmul_rn(__uint_as_float(v.w), inv)));
    }
    return o;
}

template <bool BF16>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const char* __restrict__ x, const char* __restrict__ weight, float eps,
                                                      int64_t T, int H, char* __restrict__ out) {
    constexpr int ESIZE = BF16 ? 2 : 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_vec = H * ESIZE / 16;          // 16-byte vectors per row (H % 256 == 0 -> a multiple of 32)
    const int per_lane = n_vec >> 5;
    const int64_t row_bytes = (int64_t)H * ESIZE;
    grid_dep_wait();     // decode-sized calls chain this kernel and the router front end programmatically
    grid_dep_launch();
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < T; t += (int64_t)gridDim.x * 8) {
        const char* xr = x + t * row_bytes + lane * 16;
        char* yr = out + t * row_bytes + lane * 16;
        uint4 buf[kRegVecs];
        float s = 0.0f;
        if (per_lane <= kRegVecs) {
#pragma unroll
            for (int i = 0; i < kRegVecs; ++i)
                if (i < per_lane) buf[i] = ld_nc_v4(xr + i * 512);
#pragma unroll
            for (int i = 0; i < kRegVecs; ++i)
                if (i < per_lane) s += sumsq_vec<BF16>(buf[i]);
        } else {
            for (int i = 0; i < per_lane; ++i) s += sumsq_vec<BF16>(ld_ca_v4(xr + i * 512));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
        const float inv = __frsqrt_rn(__fadd_rn(__fdiv_rn(s, (float)H), eps));
        if (per_lane <= kRegVecs) {
#pragma unroll
            for (int i = 0; i < kRegVecs; ++i)
                if (i < per_lane) st_na_v4(yr + i * 512, norm_vec<BF16>(buf[i], ld_ca_v4(weight + lane * 16 + i * 512), inv));
        } else {
            for (int i = 0; i < per_lane; ++i)
                st_na_v4(yr + i * 512, norm_vec<BF16>(ld_ca_v4(xr + i * 512), ld_ca_v4(weight + lane * 16 + i * 512), inv));
        }
    }
}

}  // namespace

int launch_rmsnorm(const void* x, const void* weight, double eps, int64_t T, const dcmoe_config* cfg, void* out,
                   cudaStream_t stream) {
    if (T == 0) return DCMOE_OK;
    const int64_t blocks = std::min<int64_t>(ceil_div(T, 8), 148 * 8 * 4);
    dim3 grid((unsigned)blocks), block(256);
    const bool pdl = pdl_enabled() && T <= 64;
    if (cfg->dtype == DCMOE_BF16)
        return check_cuda(launch_kernel(rmsnorm_kernel<true>, grid, block, 0, stream, pdl, (const char*)x, (const char*)weight,
                                        (float)eps, T, cfg->hidden_size, (char*)out),
                          "rmsnorm kernel launch");
    return check_cuda(launch_kernel(rmsnorm_kernel<false>, grid, block, 0, stream, pdl, (const char*)x, (const char*)weight,
                                    (float)eps, T, cfg->hidden_size, (char*)out),
                      "rmsnorm kernel launch");
}

}  // namespace dcmoe
