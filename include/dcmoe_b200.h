/*
 * dcmoe_b200.h -- C ABI of the B200-native DCMoE layer forward (libdcmoe_b200.so).
 *
 * Drop-in boundary for ONE path of UniMoE-Audio: UniMoEAudioSparseMoeBlock.forward
 * (reference utils/UniMoE_Audio_core.py:236-358).  The reference has no FFI of its own (it is
 * pure PyTorch); each entry point below names the reference lines it replaces, and
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - `row_capacity` is the number of rows of the h / y buffers the caller allocated (from
 *     dcmoe_query_sizes); pass the same value to every call of one forward.
 *   - extern "C", plain device pointers + sizes, a cudaStream_t passed as void*.  No torch types.
 *   - every function returns 0 on success, a negative dcmoe_status otherwise;
 *     dcmoe_last_error() returns a thread-local description of the last failure.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - dtype: DCMOE_F32 or DCMOE_BF16 is the activation/weight dtype D of the layer
 *     (bf16 at inference, reference utils/UniMoE_Audio_mod.py:44).  Integer outputs keep the
 *     reference's dtypes: dynamic_top_k int64, expert_mask int32 (core.py:259, :165).
 *   - no function synchronises the stream or reads device data on the host: the whole layer is
 *     launch-only (CUDA-graph capturable); data-dependent sizes (rows per expert, tile lists)
 *     live in the device-side plan.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     DCMOE_ERR_CUDA.
 *
 * Row space.  Rows handed to the expert FFNs live in one "row space" of `row_capacity` rows:
 *     [0, T_pad)                     shared-expert rows, row t = token t          (T_pad = ceil128(T))
 *     [T_pad, T_pad + sum_e ceil128(count_e))   routed rows, expert-major, ascending token id inside
 *                                    an expert (the canonical stable permutation)
 * x_packed holds the routed part only (row - T_pad); h / y hold the whole row space.
 */
#ifndef DCMOE_B200_H_
#define DCMOE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCMOE_ABI_VERSION 3

typedef enum dcmoe_dtype { DCMOE_F32 = 0, DCMOE_BF16 = 1 } dcmoe_dtype;

typedef enum dcmoe_status {
    DCMOE_OK = 0,
    DCMOE_ERR_INVALID = -1,     /* bad argument / unsupported configuration        */
    DCMOE_ERR_CUDA = -2,        /* CUDA runtime / launch failure, or no CUDA device */
    DCMOE_ERR_UNSUPPORTED = -3  /* valid reference config this build does not implement */
} dcmoe_status;

/* Constructor contract: the fields UniMoEAudioSparseMoeBlock.__init__ reads from
 * utils/config.json["text_config"] (core.py:204-234, :24, :42). */
typedef struct dcmoe_config {
    int32_t hidden_size;               /* 2048  */
    int32_t n_real;                    /* mlp_dynamic_expert_num       8 */
    int32_t n_null;                    /* mlp_dynamic_null_expert_num  1 */
    int32_t n_fix;                     /* mlp_fixed_expert_num         2 */
    int32_t dynamic_intermediate_size; /* 2752  */
    int32_t shared_intermediate_size;  /* 1376  (n_fix * shared == dynamic is required) */
    int32_t dtype;                     /* dcmoe_dtype */
    int32_t fixed_top_k;               /* mlp_dynamic_top_k when top_p == 0 (fixed top-k routing, core.py:256-257), else 0 */
    double top_p;                      /* mlp_dynamic_top_p     0.7  (python float); 0 selects fixed_top_k experts per token */
    double jitter_eps;                 /* router_jitter_noise   0.01 (python float) */
} dcmoe_config;

/* Sizes of the device-side plan / workspace for T local tokens and `row_capacity` FFN rows. */
typedef struct dcmoe_sizes {
    int64_t n_blocks;          /* router token blocks (DCMOE_ROUTER_BLOCK tokens each)          */
    int64_t t_pad;             /* ceil128(T)                                                    */
    int64_t max_mtiles;        /* upper bound on 128-row tiles in row space                     */
    int64_t row_capacity;      /* rows of h / y  (worst case t_pad + n_real*T + 128*n_real)     */
    int64_t plan_bytes;        /* bytes of the plan buffer (counts, offsets, tile table, aux)   */
} dcmoe_sizes;

#define DCMOE_ROUTER_BLOCK 16
#define DCMOE_TILE_M 128

/* Offsets (in bytes) of the fields inside the plan buffer; all int32 unless noted. */
typedef struct dcmoe_plan_layout {
    int64_t block_counts;   /* [n_blocks][n_dyn] int32   per-block selected-token counts (router)  */
    int64_t block_probs;    /* [n_blocks][n_dyn] float   per-block sums of aux softmax (router)    */
    int64_t block_offsets;  /* [n_blocks][n_real] int32  exclusive prefix over blocks (plan)       */
    int64_t counts;         /* [n_real] int32            tokens per routed expert (core.py:455)    */
    int64_t seg_base;       /* [n_real+1] int32          row-space start of each routed segment;   */
                            /*                           [n_real] = end of used row space          */
    int64_t n_mtiles;       /* [1] int32                 number of valid entries in mtiles         */
    int64_t aux_loss;       /* [1] float                 audio_load_balancing_loss_func result     */
    int64_t mtiles;         /* [max_mtiles] dcmoe_mtile                                          */
    int64_t overflow;       /* [1] int32                 1 if the rows did not fit row_capacity (tiles dropped!)  */
    int64_t total;          /* == plan_bytes */
    int64_t small_tokens;   /* [DCMOE_SMALL_ROWS] int32  decode-sized calls (dcmoe_front_small): token of every routed row of   */
                            /*                           row space -- the weight-streaming GEMM-1 gathers its token rows        */
                            /*                           straight from x with TMA gather4 (T <= 32), x_packed is not read       */
} dcmoe_plan_layout;
#define DCMOE_SMALL_ROWS 2176   /* 128 + 16 x 128: row space of a call with T <= 64 tokens */

typedef struct dcmoe_mtile {
    int32_t a_row;     /* GEMM-1 A row: in x (group == n_real) or in x_packed (routed)  */
    int32_t out_row;   /* row-space row of this tile (h, y, scales)                     */
    int32_t group;     /* weight group: 0..n_real-1 routed expert, n_real = shared pack */
    int32_t rows;      /* valid rows in this tile (1..128)                              */
} dcmoe_mtile;

const char* dcmoe_last_error(void);
int dcmoe_abi_version(void);

/* Fill `sizes` / `layout` for T tokens.  row_capacity_hint <= 0 selects the worst case. Host only. */
int dcmoe_query_sizes(const dcmoe_config* cfg, int64_t T, int64_t row_capacity_hint, dcmoe_sizes* sizes,
                      dcmoe_plan_layout* layout);

/*
 * Router.  Replaces core.py:251 (gate Linear), :255 -> :157-167 (Top-P count), :259-284 (mixer loop ->
 * :94-154, normalise), :286-291 (padding mask, shared columns), :293-300 -> :361-389 (aux-loss partials),
 * :331-332 -> :178-193 (global weights).
 *   x            [T, H] D
 *   w_gate       [n_dyn + n_fix, H] D        (gate.weight)
 *   logits_in    [T, E] D or NULL.  When given, the gate projection is skipped and routing runs on
 *                these logits ("bit-exact given identical router logits").
 *   attn_mask    [T] int32 (0/1) or NULL      (padding_token_mask, model.py:241)
 * outputs
 *   logits_out   [T, E] D          full_router_logits
 *   top_k        [T] int64         dynamic_top_k
 *   expert_mask  [T, E] int32      shared columns = 1
 *   global_weight[T, E] D
 *   plan         block_counts / block_probs sections are written
 */
int dcmoe_router(const void* x, const void* w_gate, const void* logits_in, const int32_t* attn_mask, int64_t T,
                 const dcmoe_config* cfg, void* logits_out, int64_t* top_k, int32_t* expert_mask,
                 void* global_weight, void* plan, void* stream);

/*
 * Router with the branches the V2 training recipe turns on (UniMoEV2-Preview/script/training.sh:46-59).
 *   keep         [T, E] uint8 or NULL: token_drop's capacity mask from dcmoe_drop_select.  When given, a dynamic column
 *                survives only where keep != 0 (core.py:314-316), the dropped weights are zeroed and the weights are
 *                normalised a second time (core.py:328-329) before the global weights (core.py:331-332).  The per-block
 *                aux partials written to the plan are still those of the mask before the drop.
 *   flags        DCMOE_ROUTER_FP32_GATE: the training-mode fp32 gate (core.py:240-249) on a bf16 layer: x / w_gate are
 *                bf16 and widened on load, logits_in / logits_out are FLOAT32 and all routing arithmetic is fp32;
 *                global_weight is still written in D (core.py:339).  Ignored on an fp32 layer.
 */
#define DCMOE_ROUTER_FP32_GATE 1
int dcmoe_router_ex(const void* x, const void* w_gate, const void* logits_in, const int32_t* attn_mask, const uint8_t* keep,
                    int flags, int64_t T, const dcmoe_config* cfg, void* logits_out, int64_t* top_k, int32_t* expert_mask,
                    void* global_weight, void* plan, void* stream);

/*
 * Token drop, drop_policy == "probs".  Replaces core.py:304 -> :170-175 (capacity) and :305-314 (per-expert
 * torch.topk over tokens on the masked logits + scatter).  Host only: capacity = max(ceil(f32(T / n_dyn) *
 * f32(capacity_factor)), min_capacity), clamped to T.
 */
int dcmoe_expert_capacity(int64_t T, const dcmoe_config* cfg, double capacity_factor, int64_t min_capacity, int64_t* capacity);
/*   logits       [T, E] of logits_dtype (DCMOE_F32 with the fp32 gate, else D)
 *   expert_mask  [T, E] int32: the mask BEFORE the drop (dcmoe_router's output)
 *   key_scratch  n_dyn * T * 8 bytes
 *   keep         [T, E] uint8 out: per dynamic column the `capacity` selected tokens with the largest logit (all of them
 *                when at most `capacity` selected it); ties at the boundary go to the LOWER token index (torch.topk
 *                leaves them to the implementation).  Shared columns 1.  Exact integer work (radix select). */
int dcmoe_drop_select(const void* logits, int logits_dtype, const int32_t* expert_mask, int64_t T, const dcmoe_config* cfg,
                      int64_t capacity, void* key_scratch, uint8_t* keep, void* stream);

/*
 * aux_balance_weight branch of the load-balancing loss.  Replaces core.py:380-385 (+ :370-374, :387-389): weighted means
 * over tokens of the pre-drop mask and of softmax_9(logits masked with finfo.min).
 *   weight          [T] float (the caller converts the reference's [B, S] int64 / float tensor), or NULL: the plain means of
 *                   core.py:378-379 in the arithmetic of logits_dtype -- used with the fp32 gate on a bf16 layer, where
 *                   dcmoe_plan would round the mean probability to bf16
 *   integer_weights 1 when the reference tensor has an integer dtype: `global_weight * w` then stays in D (rounded to
 *                   bf16 on a bf16 layer); 0: promoted to fp32
 *   scratch         ceil(T / 16) * 32 floats;   aux_out [1] float
 */
int dcmoe_aux_weighted(const void* logits, int logits_dtype, const int32_t* expert_mask, const float* weight, int integer_weights,
                       int64_t T, const dcmoe_config* cfg, float* scratch, float* aux_out, void* stream);

/* Test hook: y[i] = the router's exponential of x[i].  mode 0: the production correctly rounded expf (float-pair fast
 * path + double-precision fallback, csrc/exp_fast.cuh), 1: (float)exp((double)x) (the definition), 2: the fp32 layers'
 * Sleef expf_u10.  tests/ compare 0 against 1 and against oracle/exp_fast.h. */
int dcmoe_test_exp(const float* x, float* y, int64_t n, int mode, void* stream);

/*
 * Plan: exact integer histogram -> exclusive prefix sums -> segment bases, tile table, aux loss.
 * Replaces core.py:455 (capacity = mask.sum(0).max(), here exact per-expert counts instead of a padded
 * capacity) and finishes core.py:376-389 (means over tokens, deterministic order).
 */
int dcmoe_plan(int64_t T, int64_t row_capacity, const dcmoe_config* cfg, void* plan, void* stream);

/*
 * Decode front end: dcmoe_router + dcmoe_plan + dcmoe_permute in one single-CTA launch for T <= 64 tokens (bf16):
 * the generation loop (reference model.py:1149-1203) calls the layer with a handful of tokens, where the three
 * launches are pure latency.  Same outputs, bit for bit (the block_counts / block_probs / block_offsets scratch
 * sections of the plan are not written).
 */
int dcmoe_front_small(const void* x, const void* w_gate, const int32_t* attn_mask, int64_t T, int64_t row_capacity,
                      const dcmoe_config* cfg, void* logits_out, int64_t* top_k, int32_t* expert_mask, void* global_weight,
                      void* plan, void* x_packed, int32_t* slot_of, int32_t* row_token, float* row_scale, void* stream);

/*
 * Permute (dispatch).  Replaces core.py:459-462 + utils/UniMoE_Audio_utils.py:436-485
 * (compress_matrix x2 + the 0/1 "ce,cem->ecm" einsum): rows of x are gathered into the packed
 * routed part of row space with 128-bit loads/stores.
 *   x_packed     [row_capacity - t_pad, H] D
 *   slot_of      [T, n_real] int32   row-space row of (token, expert) or -1          (combine side)
 *   row_token    [row_capacity] int32 token id of each routed row (-1 for padding / shared rows hold t)
 *   row_scale    [row_capacity, 2] float  weights folded into the FFN: routed row -> (gw[t,e], gw[t,e]);
 *                shared row t -> (gw[t,n_dyn], gw[t,n_dyn+1])        (core.py:447, :348)
 */
int dcmoe_permute(const void* x, const int32_t* expert_mask, const void* global_weight, int64_t T,
                  int64_t row_capacity, const dcmoe_config* cfg, const void* plan, void* x_packed, int32_t* slot_of,
                  int32_t* row_token, float* row_scale, void* stream);

/*
 * Expert FFNs (routed + shared) as two grouped GEMMs over row space.  Replaces core.py:475 ->
 * :406-416, :48-49 (24 cuBLAS calls on padded [C, H] blocks) and core.py:344-349 -> :30-31 (shared
 * experts; packed as weight group n_real because n_fix * I_s == I_d).
 *   w13          [n_real+1, 2*I_d, H] D   gate/up rows interleaved in blocks of 64
 *   w2           [n_real+1, H, I_d] D
 *   h            [row_capacity, I_d] D    silu(gate) * up * row_scale   (scratch)
 *   y            [row_capacity, H] D      already weighted expert outputs
 *   impl         0 = tcgen05/TMEM/TMA grouped GEMM, one CTA per 128 x 256 tile (bf16 only); 1 = CUDA-core
 *                fp32-accumulate GEMM (the fp32 layer path; also usable with bf16 for cross-checking); 3 = weight-streaming
 *                tcgen05 GEMMs for decode-sized calls (bf16, T <= 64: weights as the MMA M operand, one K pass per SM; h
 *                and y bit-equal to impl 0's tiles).  impl 0 selects them by itself when T <= 64 (see bit 20;
 *                DCMOE_FFN_STREAM=0 disables)
 *   phase        low 4 bits: 0 = both GEMMs, 1 = GEMM-1 only (x -> h), 2 = GEMM-2 only (h -> y); lets a caller
 *                put CUDA events between the two launches.  Bits 4-5 select the tile group (tcgen05 only):
 *                0 = every row tile, 1 = shared-expert tiles only, 2 = routed tiles only -- expert parallelism runs the
 *                shared experts while the dispatch / combine gather are in flight.  Bits 8-19: cap on the number of
 *                persistent CTAs (0 = one per SM), to leave SMs to concurrently running dispatch / combine kernels.
 *                Bits 21-24 / 25-27 (impl 3): expert-parallel decode -- n_loc and rank: the plan and x_packed cover all
 *                experts (replicated routing of the gathered tokens), w13 / w2 hold this rank's n_loc routed experts as
 *                groups 0..n_loc-1 plus the shared pair as group n_loc; only those row tiles are computed.
 *                Bit 20: never select the decode-sized kernels automatically (set by expert parallelism, where a
 *                rank can own more rows than it has tokens)
 */
int dcmoe_grouped_ffn(const void* x, const void* x_packed, const void* w13, const void* w2, const float* row_scale,
                      int64_t T, int64_t row_capacity, const dcmoe_config* cfg, const void* plan, void* h, void* y,
                      int impl, int phase, void* stream);

/* Device buffers of one forward, sized by dcmoe_query_sizes (all caller-owned; reused from call to call). */
typedef struct dcmoe_workspace {
    void* plan;        /* sizes.plan_bytes */
    void* x_packed;    /* [row_capacity - t_pad, H] D */
    int32_t* slot_of;  /* [T, n_real] */
    int32_t* row_token;/* [row_capacity] */
    float* row_scale;  /* [row_capacity, 2] (zero-initialised once) */
    void* h;           /* [row_capacity, I_d] D */
    void* y;           /* [row_capacity, H] D */
} dcmoe_workspace;

/*
 * One layer forward in one host call: UniMoEAudioSparseMoeBlock.forward (core.py:236-358) = router (or the fused front
 * end for T <= 64 in bf16) -> plan -> permute -> grouped FFN -> combine, launched back to back on `stream`; nothing is
 * read back.  Arguments as in the per-stage calls; `residual` and `aux_out` may be NULL (aux_out: the aux loss as a
 * per-call scalar, see dcmoe_combine_aux); impl as in dcmoe_grouped_ffn (0 for bf16, 1 for fp32).
 */
int dcmoe_forward(const void* x, const void* w_gate, const int32_t* attn_mask, const void* w13, const void* w2, int64_t T,
                  int64_t row_capacity, const dcmoe_config* cfg, const dcmoe_workspace* ws, const void* residual, void* out,
                  void* logits_out, int64_t* top_k, int32_t* expert_mask, void* global_weight, float* aux_out, int impl,
                  void* stream);

/*
 * Pre-MoE RMSNorm (decoder-layer glue in front of the block).  Replaces utils/UniMoE_Audio_model.py:240
 * `hidden_states = self.post_attention_layernorm(hidden_states)` (Qwen2RMSNorm, model.py:207, eps = rms_norm_eps):
 * y = D(weight * D(float(x) * rsqrt(mean(float(x)^2) + eps))).  One pass: 2 x T x H x sizeof(D) bytes of HBM traffic.
 * Together with `residual` of dcmoe_combine this covers model.py:239-242 around the MoE call.
 *   x, out       [T, H] D (out may not alias x: the caller keeps x as the residual)
 *   weight       [H] D
 */
int dcmoe_rmsnorm(const void* x, const void* weight, double eps, int64_t T, const dcmoe_config* cfg, void* out,
                  void* stream);

/*
 * Combine.  Replaces core.py:486-488 + utils/UniMoE_Audio_utils.py:488-523 (decompress_matrix + "se,sem->sm"
 * einsum), core.py:338-353 (zeros + adds of routed and shared outputs): per token, a gather of the shared
 * row and of its <= n_real routed rows, fp32 accumulate in fixed order (deterministic, no atomics).
 *   residual     [T, H] D or NULL: when given, out = D(residual + layer output), i.e. the decoder layer's residual add
 *                (reference utils/UniMoE_Audio_model.py:242) is fused into the same pass (added last, in fp32)
 *   out          [T, H] D
 */
int dcmoe_combine(const void* y, const int32_t* slot_of, int64_t T, const dcmoe_config* cfg, const void* residual,
                  void* out, void* stream);

/* dcmoe_combine that also copies the layer's auxiliary loss (a float in the plan buffer, written by dcmoe_plan /
 * dcmoe_front_small: core.py:361-389) into a caller-owned scalar inside the same launch -- the reference returns
 * aux_loss as a fresh tensor per call (core.py:358); this saves the separate 4-byte copy kernel per layer call.
 *   aux_src      &plan[layout.aux_loss]      aux_dst   [1] float */
int dcmoe_combine_aux(const void* y, const int32_t* slot_of, int64_t T, const dcmoe_config* cfg, const void* residual,
                      void* out, const float* aux_src, float* aux_dst, void* stream);

/*
 * Weight packing (one-off, at load time): from the reference's separate gate_proj / up_proj / down_proj
 * matrices (state-dict keys in SURVEY.md 8b) into the grouped layouts above.
 *   group        0..n_real-1: routed expert;  n_real: shared pack, `part` = shared expert index
 *   gate_proj, up_proj [I, H] D;  down_proj [H, I] D     with I = I_d (routed) or I_s (shared)
 */
int dcmoe_pack_expert(const void* gate_proj, const void* up_proj, const void* down_proj, int group, int part,
                      const dcmoe_config* cfg, void* w13, void* w2, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Expert parallelism (reference: AudioMOELayer.forward with ep_group, core.py:455-457, :467, :480;
 * group wiring core.py:505-520).  Rank r owns routed experts [r*n_loc, (r+1)*n_loc), n_loc = n_real/world;
 * gate and shared experts are replicated; every rank routes its own T tokens.  The two all_to_all_single
 * calls of the reference become peer-memory stores (dispatch) and peer-memory loads (combine) inside the
 * permute / combine kernels; buffers that peers touch are allocated with dcmoe_ipc_alloc and mapped with
 * dcmoe_ipc_export / dcmoe_ipc_import (cudaIpc*).  Up to 8 ranks on one NVSwitch node.
 * Call order per forward on each rank (all launch-only, same stream):
 *   dcmoe_router -> dcmoe_plan (local counts + block prefix sums)
 *   -> [all-gather of (counts[0..n_real), T) over ranks: NCCL, done by the caller]
 *   -> dcmoe_ep_plan -> dcmoe_ep_dispatch -> [barrier] -> dcmoe_grouped_ffn with a config whose
 *   n_real = n_loc (w13 / w2 hold the local experts + the shared pack) -> [barrier] -> dcmoe_ep_combine.
 */
#define DCMOE_MAX_RANKS 8
#define DCMOE_EP_META_INTS 32

int dcmoe_ipc_alloc(int64_t bytes, void** ptr);
int dcmoe_ipc_free(void* ptr);
int dcmoe_ipc_export(const void* ptr, uint8_t* handle64);          /* 64-byte cudaIpcMemHandle_t */
int dcmoe_ipc_import(const uint8_t* handle64, void** ptr);
int dcmoe_ipc_close(void* ptr);

/* all_counts: device int32 [world][n_real + 1] (per-rank routed rows per global expert, then the rank's T).
 * Writes this rank's row-space layout (seg_base, counts of the LOCAL experts, tile table) into `plan` and the
 * destinations of this rank's rows into ep_meta (device int32 [DCMOE_EP_META_INTS]); also writes the shared-expert
 * scales of the local rows (row_scale_local[t] = global_weight[t, n_dyn..]) so the shared experts can start early. */
int dcmoe_ep_plan(const int32_t* all_counts, int rank, int world, int64_t T, int64_t row_capacity,
                  const dcmoe_config* cfg, void* plan, int32_t* ep_meta, const void* global_weight,
                  float* row_scale_local, void* stream);

/* peer_x_packed / peer_row_scale: HOST arrays of `world` device pointers (entry `rank` = this rank's own
 * buffers).  slot_of [T, n_real] receives the row-space row ON THE OWNER of (token, expert), or -1.
 * max_ctas > 0 caps the grid (the kernels loop), for running them next to a persistent GEMM. */
int dcmoe_ep_dispatch(const void* x, const int32_t* expert_mask, const void* global_weight, int64_t T,
                      int64_t row_capacity, const dcmoe_config* cfg, const void* plan, const int32_t* ep_meta,
                      int rank, int world, void* const* peer_x_packed, float* const* peer_row_scale, int32_t* slot_of,
                      int max_ctas, void* stream);

/* peer_y: HOST array of `world` device pointers to the ranks' y buffers.  out [T, H] D.
 * mode 0: whole combine.  mode 1: partial[T, H] (fp32) = sum of the routed rows only (can run while the shared
 * experts' GEMM-2 is still computing); mode 2: out = D(partial + shared row).  0 and 1+2 give bitwise equal outputs. */
int dcmoe_ep_combine(const void* y_local, const void* const* peer_y, const int32_t* slot_of, int64_t T,
                     const dcmoe_config* cfg, int world, int mode, float* partial, void* out, int max_ctas,
                     void* stream);

/*
 * Stream-ordered barrier of the expert-parallel group over peer memory, optionally carrying a payload (instead of NCCL:
 * the all-gather of the per-rank counts, the all-gather of decode-sized token rows, the 4-byte all-reduces used as
 * barriers).  Every rank allocates int32 flags[DCMOE_EP_FLAG_SLOTS][DCMOE_MAX_RANKS] with dcmoe_ipc_alloc (zeroed), maps
 * the peers' arrays, and calls this with the same (slot, epoch) on every rank, epoch increasing from call to call.
 * Work enqueued on `stream` after the call sees everything every rank's stream wrote -- to its own or to peer memory --
 * before that rank's call, and (payload_bytes > 0) holds every rank's payload: rank q's `payload` lands at
 * peer_payload_dst[r] + q * payload_bytes on every rank r (an all-gather into buffers the ranks mapped beforehand).
 * peer_flags / peer_payload_dst: HOST arrays of `world` device pointers (entry `rank` = own buffers).
 * payload_bytes must be a multiple of 4.  The wait is bounded (~20 s), then the kernel traps.
 */
#define DCMOE_EP_FLAG_SLOTS 8
int dcmoe_ep_barrier(int32_t* const* peer_flags, int rank, int world, int slot, int32_t epoch, const void* payload,
                     int64_t payload_bytes, void* const* peer_payload_dst, void* stream);

/*
 * Weight-gather expert parallelism (large token counts): the experts stay sharded in HBM (core.py:505: rank r holds
 * experts [r*n_loc, (r+1)*n_loc) + the replicated shared pack as w13 / w2 packs of n_loc + 1 groups), but instead of
 * sending every routed ROW to its expert's owner and back (core.py:467, :480: ~ T_loc * r * 8 KB per layer), each rank
 * pulls the remote experts' packed WEIGHTS ((world-1)/world * 270 MB per layer) into a staging pack of all n_real + 1
 * groups and runs the single-GPU forward on its own tokens: no dispatch, no combine exchange, no load imbalance when
 * the routing is skewed.  One cudaMemcpyAsync per peer and matrix (copy engines over NVLink, no SM involved), enqueued
 * on `stream`; peers are visited in ring order so that every GPU is read by one peer at a time.
 *   peer_w13 / peer_w2   HOST arrays of `world` device pointers to the ranks' packs (cudaIpc-mapped; entry `rank` = own)
 *   w13_full / w2_full   [n_real + 1, 2*I_d, H] / [n_real + 1, H, I_d] D   (cfg describes the FULL layer)
 */
int dcmoe_ep_fetch_weights(const void* const* peer_w13, const void* const* peer_w2, int rank, int world,
                           const dcmoe_config* cfg, void* w13_full, void* w2_full, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCMOE_B200_H_ */
