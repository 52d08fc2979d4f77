// api.cu -- extern "C" entry points of libdcmoe_b200.so (declared in include/dcmoe_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda.h>

#include "common.cuh"

namespace dcmoe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t err, const char* what) {
    if (err == cudaSuccess) return DCMOE_OK;
    set_error("%s: %s", what, cudaGetErrorString(err));
    return DCMOE_ERR_CUDA;
}

int validate_config(const dcmoe_config* cfg) {
    if (cfg == nullptr) {
        set_error("config is NULL");
        return DCMOE_ERR_INVALID;
    }
    const int n_dyn = cfg->n_real + cfg->n_null;
    if (cfg->dtype != DCMOE_F32 && cfg->dtype != DCMOE_BF16) {
        set_error("dtype must be DCMOE_F32 or DCMOE_BF16 (got %d)", cfg->dtype);
        return DCMOE_ERR_INVALID;
    }
    if (cfg->n_real < 1 || cfg->n_null < 0 || n_dyn > kMaxDyn || n_dyn + cfg->n_fix > kMaxDyn) {
        set_error("expert counts out of range: n_real=%d n_null=%d n_fix=%d (n_real+n_null+n_fix <= %d)", cfg->n_real,
                  cfg->n_null, cfg->n_fix, kMaxDyn);
        return DCMOE_ERR_INVALID;
    }
    if (cfg->n_fix < 1 || cfg->n_fix > 2 ||
        cfg->n_fix * cfg->shared_intermediate_size != cfg->dynamic_intermediate_size) {
        set_error("shared experts must pack into one routed-size group: n_fix (1 or 2) * shared_intermediate_size "
                  "== dynamic_intermediate_size (got %d * %d vs %d)",
                  cfg->n_fix, cfg->shared_intermediate_size, cfg->dynamic_intermediate_size);
        return DCMOE_ERR_UNSUPPORTED;
    }
    if (cfg->hidden_size % 256 != 0 || cfg->dynamic_intermediate_size % 64 != 0 ||
        cfg->shared_intermediate_size % 32 != 0) {
        set_error("hidden_size %% 256, dynamic_intermediate_size %% 64 and shared_intermediate_size %% 32 must be 0 "
                  "(got %d, %d, %d)",
                  cfg->hidden_size, cfg->dynamic_intermediate_size, cfg->shared_intermediate_size);
        return DCMOE_ERR_UNSUPPORTED;
    }
    if (!(cfg->top_p >= 0.0) || cfg->top_p > 1.0) {
        set_error("mlp_dynamic_top_p must lie in [0, 1] (got %g)", cfg->top_p);
        return DCMOE_ERR_INVALID;
    }
    if (cfg->top_p == 0.0 ? cfg->fixed_top_k < 1 : cfg->fixed_top_k != 0) {
        set_error("fixed_top_k must be >= 1 when top_p == 0 (core.py:256-257) and 0 otherwise (got top_p %g, fixed_top_k %d)",
                  cfg->top_p, cfg->fixed_top_k);
        return DCMOE_ERR_INVALID;
    }
    return DCMOE_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(err));
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

// 2-D bf16 row-major tensor [rows, cols], box [box_rows, 64 cols], 128B swizzle
int make_tensor_map_bf16(void* map_, const void* base, int64_t rows, int64_t cols, int box_rows) {
    CUtensorMap* map = static_cast<CUtensorMap*>(map_);
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return DCMOE_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d base=%p", (int)r, (long long)rows,
                  (long long)cols, box_rows, base);
        return DCMOE_ERR_CUDA;
    }
    return DCMOE_OK;
}


static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static int fill_sizes(const dcmoe_config* cfg, int64_t T, int64_t row_capacity_hint, dcmoe_sizes* sz,
                      dcmoe_plan_layout* l) {
    sz->n_blocks = ceil_div(T, kRouterBlock);
    sz->t_pad = round_up(T, kTileM);
    const int64_t worst = sz->t_pad + (int64_t)cfg->n_real * T + (int64_t)kTileM * cfg->n_real;
    sz->row_capacity = row_capacity_hint > 0 ? row_capacity_hint : worst;
    if (sz->row_capacity < sz->t_pad) {
        set_error("row_capacity %lld smaller than the shared-expert rows %lld", (long long)sz->row_capacity,
                  (long long)sz->t_pad);
        return DCMOE_ERR_INVALID;
    }
    sz->max_mtiles = sz->row_capacity / kTileM;
    int64_t off = 0;
    // sections are sized for kMaxDyn columns so that the layout does not depend on the expert counts
    // (expert parallelism runs the FFN with a per-rank config on the same plan buffer)
    l->block_counts = off; off = align_up(off + sz->n_blocks * kMaxDyn * 4, 16);
    l->block_probs = off;  off = align_up(off + sz->n_blocks * kMaxDyn * 4, 16);
    l->block_offsets = off; off = align_up(off + sz->n_blocks * kMaxDyn * 4, 16);
    l->counts = off;       off = align_up(off + kMaxDyn * 4, 16);
    l->seg_base = off;     off = align_up(off + (kMaxDyn + 1) * 4, 16);
    l->n_mtiles = off;     off = align_up(off + 4, 16);
    l->aux_loss = off;     off = align_up(off + 4, 16);
    l->mtiles = off;       off = align_up(off + sz->max_mtiles * (int64_t)sizeof(dcmoe_mtile), 16);
    l->overflow = off;     off = align_up(off + 4, 16);
    l->small_tokens = off; off = align_up(off + DCMOE_SMALL_ROWS * 4, 16);
    l->total = off;
    sz->plan_bytes = off;
    return DCMOE_OK;
}

// launchers implemented in the other translation units
int launch_router(const void*, const void*, const void*, const int32_t*, const uint8_t*, int, int64_t, const dcmoe_config*,
                  void*, int64_t*, int32_t*, void*, int32_t*, float*, cudaStream_t);
int launch_exp_test(const float*, float*, int64_t, int, cudaStream_t);
int launch_drop_select(const void*, bool, const int32_t*, int64_t, const dcmoe_config*, int64_t, void*, uint8_t*, cudaStream_t);
int launch_aux_weighted(const void*, bool, const int32_t*, const float*, bool, int64_t, const dcmoe_config*, float*, float*,
                        cudaStream_t);
int launch_plan(int64_t, const dcmoe_config*, const dcmoe_sizes&, PlanView, cudaStream_t);
int launch_front_small(const void*, const void*, const int32_t*, int64_t, const dcmoe_config*, const dcmoe_sizes&, PlanView,
                       void*, int64_t*, int32_t*, void*, void*, int32_t*, int32_t*, float*, bool, cudaStream_t);
bool ffn_stream_gathers_from_x(int64_t T);
int launch_permute(const void*, const int32_t*, const void*, int64_t, const dcmoe_config*, const dcmoe_sizes&, PlanView,
                   void*, int32_t*, int32_t*, float*, cudaStream_t);
int launch_combine(const void*, const int32_t*, int64_t, const dcmoe_config*, const void*, void*, const float*, float*,
                   cudaStream_t);
int launch_pack(const void*, const void*, const void*, int, int, const dcmoe_config*, void*, void*, cudaStream_t);
int ep_plan_view(const dcmoe_config* cfg, int64_t T, int64_t row_capacity, void* plan, dcmoe_sizes* sz, PlanView* pv);
int launch_ffn_simt(const void*, const void*, const void*, const void*, const float*, int64_t, const dcmoe_config*,
                    const dcmoe_sizes&, PlanView, void*, void*, int, int, cudaStream_t);
int launch_ffn_tcgen05(const void*, const void*, const void*, const void*, const float*, int64_t, int64_t,
                       const dcmoe_config*, const dcmoe_sizes&, PlanView, void*, void*, int, int, int, cudaStream_t);
int launch_rmsnorm(const void*, const void*, double, int64_t, const dcmoe_config*, void*, cudaStream_t);
bool ffn_stream_applicable(int64_t, const dcmoe_config*, const dcmoe_sizes&, int, int);
int launch_ffn_tcgen05_stream(const void*, const void*, const void*, const void*, const float*, int64_t, int64_t,
                              const dcmoe_config*, const dcmoe_sizes&, PlanView, void*, void*, int, int, int, int, cudaStream_t);

int device_sm_count() {
    static int cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("DCMOE_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

static int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); libdcmoe_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return DCMOE_ERR_CUDA;
    }
    return DCMOE_OK;
}

int ep_plan_view(const dcmoe_config* cfg, int64_t T, int64_t row_capacity, void* plan, dcmoe_sizes* sz, PlanView* pv) {
    dcmoe_plan_layout l;
    int rc = fill_sizes(cfg, T, row_capacity, sz, &l);
    if (rc) return rc;
    *pv = plan_view(plan, l);
    return DCMOE_OK;
}

}  // namespace dcmoe

using namespace dcmoe;

extern "C" {

const char* dcmoe_last_error(void) { return g_err; }
int dcmoe_abi_version(void) { return DCMOE_ABI_VERSION; }

int dcmoe_query_sizes(const dcmoe_config* cfg, int64_t T, int64_t row_capacity_hint, dcmoe_sizes* sizes,
                      dcmoe_plan_layout* layout) {
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (T < 0 || sizes == nullptr || layout == nullptr) {
        set_error("dcmoe_query_sizes: bad arguments (T=%lld)", (long long)T);
        return DCMOE_ERR_INVALID;
    }
    return fill_sizes(cfg, T, row_capacity_hint, sizes, layout);
}

#define DCMOE_PROLOGUE(T_)                                                      \
    int rc = validate_config(cfg);                                              \
    if (rc) return rc;                                                          \
    if ((T_) < 0) { set_error("negative token count"); return DCMOE_ERR_INVALID; } \
    if ((rc = require_device())) return rc;

int dcmoe_router_ex(const void* x, const void* w_gate, const void* logits_in, const int32_t* attn_mask, const uint8_t* keep,
                    int flags, int64_t T, const dcmoe_config* cfg, void* logits_out, int64_t* top_k, int32_t* expert_mask,
                    void* global_weight, void* plan, void* stream) {
    DCMOE_PROLOGUE(T)
    if (T > 0 && ((logits_in == nullptr && (x == nullptr || w_gate == nullptr)) || !logits_out || !top_k || !expert_mask ||
                  !global_weight || !plan)) {
        set_error("dcmoe_router: NULL pointer argument");
        return DCMOE_ERR_INVALID;
    }
    if (flags & ~DCMOE_ROUTER_FP32_GATE) { set_error("dcmoe_router_ex: unknown flag bits 0x%x", flags); return DCMOE_ERR_INVALID; }
    dcmoe_sizes sz; dcmoe_plan_layout l;
    if ((rc = fill_sizes(cfg, T, 0, &sz, &l))) return rc;
    PlanView pv = plan_view(plan, l);
    return launch_router(x, w_gate, logits_in, attn_mask, keep, flags, T, cfg, logits_out, top_k, expert_mask, global_weight,
                         pv.block_counts, pv.block_probs, (cudaStream_t)stream);
}

int dcmoe_router(const void* x, const void* w_gate, const void* logits_in, const int32_t* attn_mask, int64_t T,
                 const dcmoe_config* cfg, void* logits_out, int64_t* top_k, int32_t* expert_mask, void* global_weight,
                 void* plan, void* stream) {
    return dcmoe_router_ex(x, w_gate, logits_in, attn_mask, nullptr, 0, T, cfg, logits_out, top_k, expert_mask, global_weight,
                           plan, stream);
}

int dcmoe_expert_capacity(int64_t T, const dcmoe_config* cfg, double capacity_factor, int64_t min_capacity, int64_t* capacity) {
    int rc;
    if ((rc = validate_config(cfg))) return rc;
    if (!capacity || T < 0) { set_error("dcmoe_expert_capacity: bad argument"); return DCMOE_ERR_INVALID; }
    // core.py:172: the quotient is a Python float (double); product and ceil run on a 0-dim float32 tensor
    const float q = (float)((double)T / (double)(cfg->n_real + cfg->n_null));
    int64_t cap = (int64_t)ceilf(q * (float)capacity_factor);
    if (cap < min_capacity) cap = min_capacity;          // core.py:173-174
    if (cap > T) cap = T;                                // core.py:306-308
    *capacity = cap;
    return DCMOE_OK;
}

int dcmoe_drop_select(const void* logits, int logits_dtype, const int32_t* expert_mask, int64_t T, const dcmoe_config* cfg,
                      int64_t capacity, void* key_scratch, uint8_t* keep, void* stream) {
    DCMOE_PROLOGUE(T)
    if (T > 0 && (!logits || !expert_mask || !key_scratch || !keep)) { set_error("dcmoe_drop_select: NULL pointer argument"); return DCMOE_ERR_INVALID; }
    if (capacity < 0 || (logits_dtype != DCMOE_F32 && logits_dtype != DCMOE_BF16)) { set_error("dcmoe_drop_select: bad capacity / dtype"); return DCMOE_ERR_INVALID; }
    return launch_drop_select(logits, logits_dtype == DCMOE_BF16, expert_mask, T, cfg, capacity, key_scratch, keep, (cudaStream_t)stream);
}

int dcmoe_aux_weighted(const void* logits, int logits_dtype, const int32_t* expert_mask, const float* weight, int integer_weights,
                       int64_t T, const dcmoe_config* cfg, float* scratch, float* aux_out, void* stream) {
    DCMOE_PROLOGUE(T)
    if (!aux_out || (T > 0 && (!logits || !expert_mask || !scratch))) { set_error("dcmoe_aux_weighted: NULL pointer argument"); return DCMOE_ERR_INVALID; }
    if (logits_dtype != DCMOE_F32 && logits_dtype != DCMOE_BF16) { set_error("dcmoe_aux_weighted: bad dtype"); return DCMOE_ERR_INVALID; }
    return launch_aux_weighted(logits, logits_dtype == DCMOE_BF16, expert_mask, weight, integer_weights != 0, T, cfg, scratch, aux_out,
                               (cudaStream_t)stream);
}

// NOTE: the plan layout depends on (T, row_capacity); callers pass the same row_capacity to every call of a forward.
static int plan_for(const dcmoe_config* cfg, int64_t T, int64_t row_capacity, void* plan, dcmoe_sizes* sz, PlanView* pv) {
    dcmoe_plan_layout l;
    int rc = fill_sizes(cfg, T, row_capacity, sz, &l);
    if (rc) return rc;
    *pv = plan_view(plan, l);
    return DCMOE_OK;
}

int dcmoe_front_small(const void* x, const void* w_gate, const int32_t* attn_mask, int64_t T, int64_t row_capacity,
                      const dcmoe_config* cfg, void* logits_out, int64_t* top_k, int32_t* expert_mask, void* global_weight,
                      void* plan, void* x_packed, int32_t* slot_of, int32_t* row_token, float* row_scale, void* stream) {
    DCMOE_PROLOGUE(T)
    if (T > 0 && (!x || !w_gate || !logits_out || !top_k || !expert_mask || !global_weight || !plan || !x_packed || !slot_of ||
                  !row_token || !row_scale)) {
        set_error("dcmoe_front_small: NULL pointer argument");
        return DCMOE_ERR_INVALID;
    }
    dcmoe_sizes sz; PlanView pv;
    if ((rc = plan_for(cfg, T, row_capacity, plan, &sz, &pv))) return rc;
    return launch_front_small(x, w_gate, attn_mask, T, cfg, sz, pv, logits_out, top_k, expert_mask, global_weight, x_packed,
                              slot_of, row_token, row_scale, true, (cudaStream_t)stream);
}

int dcmoe_plan(int64_t T, int64_t row_capacity, const dcmoe_config* cfg, void* plan, void* stream) {
    DCMOE_PROLOGUE(T)
    if (!plan) { set_error("dcmoe_plan: NULL plan"); return DCMOE_ERR_INVALID; }
    dcmoe_sizes sz; PlanView pv;
    if ((rc = plan_for(cfg, T, row_capacity, plan, &sz, &pv))) return rc;
    return launch_plan(T, cfg, sz, pv, (cudaStream_t)stream);
}

int dcmoe_test_exp(const float* x, float* y, int64_t n, int mode, void* stream) {
    int rc;
    if ((rc = require_device())) return rc;
    if (n > 0 && (!x || !y)) { set_error("dcmoe_test_exp: NULL pointer argument"); return DCMOE_ERR_INVALID; }
    return launch_exp_test(x, y, n, mode, (cudaStream_t)stream);
}

int dcmoe_permute(const void* x, const int32_t* expert_mask, const void* global_weight, int64_t T, int64_t row_capacity,
                  const dcmoe_config* cfg, const void* plan, void* x_packed, int32_t* slot_of, int32_t* row_token,
                  float* row_scale, void* stream) {
    DCMOE_PROLOGUE(T)
    if (T > 0 && (!x || !expert_mask || !global_weight || !plan || !x_packed || !slot_of || !row_token || !row_scale)) {
        set_error("dcmoe_permute: NULL pointer argument");
        return DCMOE_ERR_INVALID;
    }
    dcmoe_sizes sz; PlanView pv;
    if ((rc = plan_for(cfg, T, row_capacity, const_cast<void*>(plan), &sz, &pv))) return rc;
    return launch_permute(x, expert_mask, global_weight, T, cfg, sz, pv, x_packed, slot_of, row_token, row_scale,
                          (cudaStream_t)stream);
}

int dcmoe_grouped_ffn(const void* x, const void* x_packed, const void* w13, const void* w2, const float* row_scale,
                      int64_t T, int64_t row_capacity, const dcmoe_config* cfg, const void* plan, void* h, void* y,
                      int impl, int phase, void* stream) {
    // phase: low 4 bits = 0 both GEMMs / 1 GEMM-1 / 2 GEMM-2; bits 4-5 = group selection (0 all tiles,
    // 1 shared-expert tiles only, 2 routed tiles only)
    const int group_sel = (phase >> 4) & 3;
    const int max_ctas = (phase >> 8) & 0xfff;   // bits 8-19: cap on the persistent grid (0 = one CTA per SM)
    const int ep_n_loc = (phase >> 21) & 15;     // bits 21-24 / 25-27 (impl 3 only): expert-parallel decode -- the plan covers
    const int ep_rank = (phase >> 25) & 7;       // all experts, w13 / w2 hold this rank's ep_n_loc routed experts + the shared pair
    const bool no_decode = (phase >> 20) & 1;    // bit 20: never pick the decode kernels (expert parallelism: a rank
                                                 // can own more rows than it has tokens)
    phase &= 15;
    DCMOE_PROLOGUE(T)
    if (T > 0 && (!x || !x_packed || !w13 || !w2 || !row_scale || !plan || !h || !y)) {
        set_error("dcmoe_grouped_ffn: NULL pointer argument");
        return DCMOE_ERR_INVALID;
    }
    if (phase < 0 || phase > 2) { set_error("dcmoe_grouped_ffn: phase must be 0, 1 or 2"); return DCMOE_ERR_INVALID; }
    dcmoe_sizes sz; PlanView pv;
    if ((rc = plan_for(cfg, T, row_capacity, const_cast<void*>(plan), &sz, &pv))) return rc;
    // decode-sized calls (T <= 64): weight-streaming tcgen05 GEMMs (ffn_tcgen05_stream.cu; bit-identical h and y).
    // impl 3 asks for them explicitly; impl 0 picks them unless the caller vetoes (bit 20), selects tile groups, or
    // DCMOE_FFN_STREAM=0 (A/B measurements)
    if (impl == 3) return launch_ffn_tcgen05_stream(x, x_packed, w13, w2, row_scale, T, sz.row_capacity, cfg, sz, pv, h, y,
                                                    phase, max_ctas, ep_n_loc, ep_rank, (cudaStream_t)stream);
    if (ep_n_loc != 0) { set_error("dcmoe_grouped_ffn: expert-parallel decode bits need impl 3"); return DCMOE_ERR_INVALID; }
    if (impl == 0 && !no_decode && group_sel == 0 && ffn_stream_applicable(T, cfg, sz, max_ctas, 0)) {
        const char* e = getenv("DCMOE_FFN_STREAM");
        if (!(e && e[0] == '0'))
            return launch_ffn_tcgen05_stream(x, x_packed, w13, w2, row_scale, T, sz.row_capacity, cfg, sz, pv, h, y, phase,
                                             max_ctas, 0, 0, (cudaStream_t)stream);
    }
    if (group_sel == 3) { set_error("dcmoe_grouped_ffn: tile group must be 0, 1 or 2"); return DCMOE_ERR_INVALID; }
    if (impl == 0) return launch_ffn_tcgen05(x, x_packed, w13, w2, row_scale, T, sz.row_capacity, cfg, sz, pv, h, y,
                                             phase, group_sel, max_ctas, (cudaStream_t)stream);
    if (impl == 1) {
        if (sz.max_mtiles > 65535) { set_error("CUDA-core FFN: too many row tiles (%lld)", (long long)sz.max_mtiles); return DCMOE_ERR_INVALID; }
        if (group_sel != 0) { set_error("CUDA-core FFN does not support tile-group selection"); return DCMOE_ERR_INVALID; }
        return launch_ffn_simt(x, x_packed, w13, w2, row_scale, T, cfg, sz, pv, h, y, phase, group_sel, (cudaStream_t)stream);
    }
    set_error("dcmoe_grouped_ffn: unknown impl %d", impl);
    return DCMOE_ERR_INVALID;
}

int dcmoe_combine(const void* y, const int32_t* slot_of, int64_t T, const dcmoe_config* cfg, const void* residual,
                  void* out, void* stream) {
    DCMOE_PROLOGUE(T)
    if (T > 0 && (!y || !slot_of || !out)) { set_error("dcmoe_combine: NULL pointer argument"); return DCMOE_ERR_INVALID; }
    return launch_combine(y, slot_of, T, cfg, residual, out, nullptr, nullptr, (cudaStream_t)stream);
}

int dcmoe_combine_aux(const void* y, const int32_t* slot_of, int64_t T, const dcmoe_config* cfg, const void* residual,
                      void* out, const float* aux_src, float* aux_dst, void* stream) {
    DCMOE_PROLOGUE(T)
    if ((T > 0 && (!y || !slot_of || !out)) || !aux_src || !aux_dst) {
        set_error("dcmoe_combine_aux: NULL pointer argument");
        return DCMOE_ERR_INVALID;
    }
    return launch_combine(y, slot_of, T, cfg, residual, out, aux_src, aux_dst, (cudaStream_t)stream);
}

int dcmoe_forward(const void* x, const void* w_gate, const int32_t* attn_mask, const void* w13, const void* w2, int64_t T,
                  int64_t row_capacity, const dcmoe_config* cfg, const dcmoe_workspace* ws, const void* residual, void* out,
                  void* logits_out, int64_t* top_k, int32_t* expert_mask, void* global_weight, float* aux_out, int impl,
                  void* stream) {
    // the calls below validate their own arguments; this entry point only sequences them (one host call per layer)
    if (!ws) { set_error("dcmoe_forward: NULL workspace"); return DCMOE_ERR_INVALID; }
    int rc;
    const bool small = cfg && cfg->dtype == DCMOE_BF16 && T > 0 && T <= 64;
    if (small) {
        // x_packed is only filled when the FFN that follows reads it: the weight-streaming GEMM-1 gathers its token rows
        // straight from x (T <= 32)
        dcmoe_sizes szf; PlanView pvf;
        if ((rc = validate_config(cfg))) return rc;
        if ((rc = require_device())) return rc;
        if (!x || !w_gate || !logits_out || !top_k || !expert_mask || !global_weight || !ws->plan || !ws->x_packed || !ws->slot_of ||
            !ws->row_token || !ws->row_scale) {
            set_error("dcmoe_forward: NULL pointer argument");
            return DCMOE_ERR_INVALID;
        }
        if ((rc = plan_for(cfg, T, row_capacity, ws->plan, &szf, &pvf))) return rc;
        const char* e = getenv("DCMOE_FFN_STREAM");
        const bool stream_ffn = impl == 0 && !(e && e[0] == '0') && ffn_stream_applicable(T, cfg, szf, 0, 0);
        if ((rc = launch_front_small(x, w_gate, attn_mask, T, cfg, szf, pvf, logits_out, top_k, expert_mask, global_weight,
                                     ws->x_packed, ws->slot_of, ws->row_token, ws->row_scale,
                                     !(stream_ffn && ffn_stream_gathers_from_x(T)), (cudaStream_t)stream)))
            return rc;
    } else {
        if ((rc = dcmoe_router(x, w_gate, nullptr, attn_mask, T, cfg, logits_out, top_k, expert_mask, global_weight, ws->plan,
                               stream)))
            return rc;
        if ((rc = dcmoe_plan(T, row_capacity, cfg, ws->plan, stream))) return rc;
        if (T > 0 && (rc = dcmoe_permute(x, expert_mask, global_weight, T, row_capacity, cfg, ws->plan, ws->x_packed,
                                         ws->slot_of, ws->row_token, ws->row_scale, stream)))
            return rc;
    }
    if (T > 0 && (rc = dcmoe_grouped_ffn(x, ws->x_packed, w13, w2, ws->row_scale, T, row_capacity, cfg, ws->plan, ws->h, ws->y,
                                         impl, 0, stream)))
        return rc;
    dcmoe_sizes sz; dcmoe_plan_layout lay;
    if ((rc = dcmoe_query_sizes(cfg, T, row_capacity, &sz, &lay))) return rc;
    const float* aux_src = reinterpret_cast<const float*>(static_cast<const char*>(ws->plan) + lay.aux_loss);
    if (aux_out) return dcmoe_combine_aux(ws->y, ws->slot_of, T, cfg, residual, out, aux_src, aux_out, stream);
    return dcmoe_combine(ws->y, ws->slot_of, T, cfg, residual, out, stream);
}

int dcmoe_rmsnorm(const void* x, const void* weight, double eps, int64_t T, const dcmoe_config* cfg, void* out,
                  void* stream) {
    DCMOE_PROLOGUE(T)
    if (T > 0 && (!x || !weight || !out)) { set_error("dcmoe_rmsnorm: NULL pointer argument"); return DCMOE_ERR_INVALID; }
    if (T > 0 && x == out) { set_error("dcmoe_rmsnorm: out must not alias x (x stays the residual)"); return DCMOE_ERR_INVALID; }
    if (!(eps >= 0.0)) { set_error("dcmoe_rmsnorm: eps must be >= 0"); return DCMOE_ERR_INVALID; }
    return launch_rmsnorm(x, weight, eps, T, cfg, out, (cudaStream_t)stream);
}

int dcmoe_pack_expert(const void* gate_proj, const void* up_proj, const void* down_proj, int group, int part,
                      const dcmoe_config* cfg, void* w13, void* w2, void* stream) {
    DCMOE_PROLOGUE(0)
    if (!gate_proj || !up_proj || !down_proj || !w13 || !w2) { set_error("dcmoe_pack_expert: NULL pointer argument"); return DCMOE_ERR_INVALID; }
    if (group < 0 || group > cfg->n_real || part < 0 || (group == cfg->n_real ? part >= cfg->n_fix : part != 0)) {
        set_error("dcmoe_pack_expert: bad group/part (%d, %d)", group, part);
        return DCMOE_ERR_INVALID;
    }
    return launch_pack(gate_proj, up_proj, down_proj, group, part, cfg, w13, w2, (cudaStream_t)stream);
}

}  // extern "C"
