/*
 * oracle/route_oracle.c -- CPU restatement (plain C) of the DCMoE Top-P router.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (unimoe_audio_b200/) may link,
 * load or call this file; it is the checker for the CUDA router kernel
 * (unimoe_audio_b200/csrc/router.cu), used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 * fixtures under tests/golden/ that were produced by running the UNMODIFIED reference
 * block (utils/UniMoE_Audio_core.py:196-358) in the build container
 * (tools/make_golden.py); see DESIGN.md section "Oracle".
 *
 * What it restates (reference file:line), per token, for logits l[0..E) in dtype D:
 *   - audio_dynamic_expert_selection            utils/UniMoE_Audio_core.py:157-167
 *   - audio_sparse_expert_mixer (eval branch)   utils/UniMoE_Audio_core.py:94-119, :139-154
 *   - scatter / one-hot / normalise             utils/UniMoE_Audio_core.py:259-291
 *   - calculate_audio_global_routing_weight     utils/UniMoE_Audio_core.py:178-193
 *   - audio_load_balancing_loss_func            utils/UniMoE_Audio_core.py:361-389  (both branches: plain means and
 *                                               the aux_balance_weight-weighted means of :380-385)
 *   - token drop, mask + renormalise step       utils/UniMoE_Audio_core.py:316, :326-329  (given the per-token keep mask;
 *                                               the capacity selection of :303-315 is oracle/dcmoe_oracle.py::drop_keep_mask)
 *
 * Arithmetic contract ("canonical arithmetic"): the *decisions*
 * (dynamic_top_k, expert_mask) depend on floating point only through the 9-way softmax,
 * the running sum and the >= top_p comparison.  Those are restated with the exact rounding
 * points of torch 2.11 CPU (AVX512 build), which is where the golden vectors come from:
 *   D = fp32: e_j = Sleef_expf_u10(l_j - max)  (what ATen's vectorised softmax calls),
 *             s = ((e_0 + e_1) + e_2) + ... (sequential), p_j = e_j * (1 / s),
 *             c_j = c_{j-1} + sorted_p_j in fp32, threshold (float)top_p.
 *   D = bf16: e_j = correctly rounded expf(l_j - max) in fp32, s sequential in fp32,
 *             p_j = bf16(e_j * (1 / s)); c_j = bf16(fp32 running sum); threshold bf16(top_p).
 * Every operation is a single IEEE-754 binary32 op (+,-,*,/,fma) so that the CUDA kernel can
 * reproduce it bit for bit with __fadd_rn/__fmul_rn/__fdiv_rn/__fmaf_rn.
 *
 * Build:  gcc -O2 -fPIC -shared -ffp-contract=off -mfma -o oracle/_build/libroute_oracle.so \
 *             oracle/route_oracle.c -lm        (done by oracle/build_oracle.py)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAX_E 32

/* ---- bf16 helpers (round-to-nearest-even, as c10::BFloat16) ---- */
static inline float bf16_round(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) { /* NaN */
        u = 0x7fc00000u;
    } else {
        uint32_t lsb = (u >> 16) & 1u;
        u += 0x7fffu + lsb;
        u &= 0xffff0000u;
    }
    memcpy(&f, &u, 4);
    return f;
}

static inline float rnd(float v, int bf16) { return bf16 ? bf16_round(v) : v; }

/* ---- exp variants ---- */
/* Sleef_expf16_u10 (what ATen's Vectorized<float>::exp() calls on AVX512/AVX2 builds; the fp32
 * vectorised softmax applies it to x - max).  SLEEF 3.6 xexpf, FMA flavour: every step below is
 * one binary32 operation, so the CUDA kernel reproduces it bit for bit. */
static inline float pow2if(int q) {
    uint32_t b = (uint32_t)(q + 127) << 23;
    float f;
    memcpy(&f, &b, 4);
    return f;
}
static inline float exp_sleef_u10(float d) {
    float qf = rintf(d * 1.442695040888963407359924681001892137426645954152985934135449406931f);
    int q = (int)qf;
    float s = fmaf(qf, -0.693145751953125f, d);
    s = fmaf(qf, -1.428606765330187045e-06f, s);
    float u = 0.000198527617612853646278381f;
    u = fmaf(u, s, 0.00139304355252534151077271f);
    u = fmaf(u, s, 0.00833336077630519866943359f);
    u = fmaf(u, s, 0.0416664853692054748535156f);
    u = fmaf(u, s, 0.166666671633720397949219f);
    u = fmaf(u, s, 0.5f);
    u = 1.0f + fmaf(s * s, u, s);
    u = u * pow2if(q >> 1) * pow2if(q - (q >> 1));
    if (d < -104.0f) u = 0.0f;
    if (d > 100.0f) u = INFINITY;
    return u;
}

/* correctly rounded expf: round-to-nearest of the double-precision exp */
static inline float exp_cr(float x) { return (float)exp((double)x); }

/* softmax over n entries of v (already in fp32 holding D values), result rounded to D.
 * Entries equal to -inf give exactly 0. */
static void softmax_D(const float* v, int n, int bf16, float* out) {
    float m = v[0];
    for (int j = 1; j < n; ++j) m = v[j] > m ? v[j] : m;
    float e[MAX_E];
    float s = 0.0f;
    if (bf16) {
        /* ATen reduced-precision path: scalar tail, std::exp, sum starts at 0 */
        for (int j = 0; j < n; ++j) {
            e[j] = exp_cr(v[j] - m);
            s = s + e[j];
        }
    } else {
        for (int j = 0; j < n; ++j) e[j] = exp_sleef_u10(v[j] - m);
        s = e[0];
        for (int j = 1; j < n; ++j) s = s + e[j];
    }
    float inv = 1.0f / s;
    for (int j = 0; j < n; ++j) out[j] = rnd(e[j] * inv, bf16);
}

/* torch.sum over an inner dimension of length n (fp32 accumulate): ATen's row_sum keeps
 * 8 interleaved partial sums (aten/src/ATen/native/cpu/SumKernel.cpp, ilp_factor = 8 in
 * torch 2.11), adds the tail into partial 0, then folds the partials left to right. */
static float row_sum8(const float* d, int n) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int n8 = n / 8;
    for (int i = 0; i < n8; ++i)
        for (int k = 0; k < 8; ++k) acc[k] = acc[k] + d[i * 8 + k];
    for (int i = n8 * 8; i < n; ++i) acc[0] = acc[0] + d[i];
    for (int k = 1; k < 8; ++k) acc[0] = acc[0] + acc[k];
    return acc[0];
}

/*
 * Route T tokens.
 *   logits      [T, E] fp32 storage holding D-representable values (E = n_dyn + n_fix)
 *   attn_mask   [T] int32 or NULL
 *   bf16        0: D = fp32, 1: D = bf16
 *   n_dyn       dynamic experts incl. null experts (9);  n_fix shared experts (2)
 *   top_p, eps  0.7, 0.01
 * outputs
 *   top_k       [T] int64
 *   mask        [T, E] int32
 *   gw          [T, E] fp32 storage (D-rounded values)
 *   aux_out     [1] fp32  -- audio_load_balancing_loss_func with aux_balance_weight = None
 */
/* fixed_k > 0: mlp_dynamic_top_p == 0 -- every token selects fixed_k (= mlp_dynamic_top_k) experts, core.py:256-257
 * keep   [T, E] uint8 or NULL: token_drop's capacity mask (core.py:313-316); a dynamic column survives only where keep
 *        is non-zero.  The aux loss is computed BEFORE the drop (core.py:293-300 precede :302), the routing weights are
 *        zeroed where the final mask is zero and normalised again (core.py:328-329), the global weights use the final mask.
 * aux_w  [T] float or NULL: aux_balance_weight (core.py:380-385), weighted means instead of plain means. */
int dcmoe_oracle_route_ex(const float* logits, const int32_t* attn_mask, const uint8_t* keep, const float* aux_w, int64_t T,
                          int n_dyn, int n_fix, int bf16, double top_p, double eps, int fixed_k, int64_t* top_k,
                          int32_t* mask, float* gw, float* aux_out) {
    const int E = n_dyn + n_fix;
    if (E > MAX_E || n_dyn < 1) return -1;
    const float thr_p = rnd((float)top_p, bf16);
    const float thr_eps = rnd((float)(2.0 * eps), bf16);
    const float plus_eps = rnd(1e-6f, bf16);
    const float finfo_min = bf16 ? -3.3895313892515355e38f : -3.4028234663852886e38f;
    double tok_sum[MAX_E], prob_sum[MAX_E], w_sum = 0.0;
    for (int j = 0; j < E; ++j) tok_sum[j] = prob_sum[j] = 0.0;

    for (int64_t t = 0; t < T; ++t) {
        const float* l = logits + t * E;
        float p[MAX_E], s[MAX_E];
        /* ---- Top-P count: core.py:162-166 ---- */
        softmax_D(l, n_dyn, bf16, p);
        memcpy(s, p, sizeof(float) * n_dyn);
        for (int a = 1; a < n_dyn; ++a) { /* insertion sort, descending (values only) */
            float key = s[a];
            int b = a - 1;
            while (b >= 0 && s[b] < key) { s[b + 1] = s[b]; --b; }
            s[b + 1] = key;
        }
        int raw = 1;
        float run = 0.0f;
        for (int j = 0; j < n_dyn; ++j) {
            run = run + s[j];                  /* fp32 running sum (acc_type) */
            float c = rnd(run, bf16);          /* each prefix rounded to D    */
            if (!(c >= thr_p)) ++raw;          /* (~(c >= p)).sum() + 1       */
        }
        if (fixed_k > 0) raw = fixed_k;        /* core.py:257: torch.full((T,), mlp_dynamic_top_k) */
        top_k[t] = raw;
        /* core.py:262 only visits groups 1..n_dyn: a token whose last prefix is still < p
         * (raw == n_dyn + 1, impossible for p <= 0.95) would select no expert at all. */
        int k = raw <= n_dyn ? raw : 0;

        /* ---- mixer: core.py:103-147 ---- */
        float rem[MAX_E], rw[MAX_E];
        int32_t* mk = mask + t * E;
        for (int j = 0; j < E; ++j) mk[j] = 0;
        for (int j = 0; j < n_dyn; ++j) { rem[j] = l[j]; rw[j] = 0.0f; }
        for (int it = 0; it < k; ++it) {
            int idx = 0;
            float thr = rem[0];
            for (int j = 1; j < n_dyn; ++j)
                if (rem[j] > thr) { thr = rem[j]; idx = j; }  /* first index on ties */
            float gates[MAX_E], sm[MAX_E];
            float athr = fabsf(thr);
            for (int j = 0; j < n_dyn; ++j) {
                float fac = fabsf(l[j]);
                fac = fac < athr ? athr : fac;                 /* clamp(min=|thr|) */
                float diff = rnd(thr - l[j], bf16);
                float ratio = rnd(diff / fac, bf16);
                int drop = ratio > thr_eps;
                gates[j] = drop ? -INFINITY : rem[j];
            }
            softmax_D(gates, n_dyn, bf16, sm);
            rw[idx] = sm[idx];
            mk[idx] = 1;
            rem[idx] = -INFINITY;
        }
        /* ---- normalise: core.py:284 ---- */
        float rs = rnd(row_sum8(rw, n_dyn), bf16);
        float den = rnd(rs + plus_eps, bf16);
        for (int j = 0; j < n_dyn; ++j) rw[j] = rnd(rw[j] / den, bf16);
        /* ---- attention mask + shared experts always on: core.py:286-291 ---- */
        if (attn_mask) {
            int32_t a = attn_mask[t];
            for (int j = 0; j < E; ++j) mk[j] *= a;
        }
        for (int j = n_dyn; j < E; ++j) mk[j] = 1;
        /* ---- aux-loss partials: core.py:370-379 ---- */
        {
            float ml[MAX_E], ga[MAX_E];
            for (int j = 0; j < n_dyn; ++j) ml[j] = mk[j] ? l[j] : finfo_min;
            softmax_D(ml, n_dyn, bf16, ga);
            if (aux_w) {   /* core.py:384-385: mask.float() * w (fp32) and global_weight * w (a D tensor) */
                const float w = aux_w[t];
                w_sum += (double)w;
                for (int j = 0; j < n_dyn; ++j) { tok_sum[j] += (double)((float)mk[j] * w); prob_sum[j] += (double)rnd(ga[j] * w, bf16); }
            } else {
                for (int j = 0; j < n_dyn; ++j) { tok_sum[j] += (double)mk[j]; prob_sum[j] += (double)ga[j]; }
            }
        }
        /* ---- token drop: core.py:316 (mask AND capacity mask), :328-329 (zero the dropped weights, normalise again) ---- */
        if (keep) {
            const uint8_t* kp = keep + t * E;
            for (int j = 0; j < n_dyn; ++j) {
                if (!kp[j]) mk[j] = 0;
                if (!mk[j]) rw[j] = 0.0f;            /* masked_fill(~expert_mask.bool(), 0): also where the padding mask cleared it */
            }
            float rs2 = rnd(row_sum8(rw, n_dyn), bf16);
            float den2 = rnd(rs2 + plus_eps, bf16);
            for (int j = 0; j < n_dyn; ++j) rw[j] = rnd(rw[j] / den2, bf16);
        }
        /* ---- global weights: core.py:188-192 ---- */
        {
            float ml[MAX_E], G[MAX_E];
            for (int j = 0; j < E; ++j) ml[j] = mk[j] ? l[j] : -INFINITY;
            softmax_D(ml, E, bf16, G);
            float dyn = rnd(row_sum8(G, n_dyn), bf16);
            float* g = gw + t * E;
            for (int j = 0; j < n_dyn; ++j) g[j] = rnd(rw[j] * dyn, bf16);
            for (int j = n_dyn; j < E; ++j) g[j] = G[j];
        }
    }
    if (aux_out) {
        double acc = 0.0;
        const double den = aux_w ? w_sum : (double)T;
        for (int j = 0; j < n_dyn; ++j) {
            float tpe = (float)(tok_sum[j] / den);
            float rp = rnd((float)(prob_sum[j] / den), bf16);
            acc += (double)(tpe * rp);
        }
        *aux_out = (float)acc * (float)n_dyn;
    }
    return 0;
}

/* Stand-alone helpers exported for unit tests of the arithmetic spec */
float dcmoe_oracle_exp_sleef(float x) { return exp_sleef_u10(x); }
float dcmoe_oracle_exp_cr(float x) { return exp_cr(x); }
/* the float-pair evaluation the CUDA router runs (oracle/exp_fast.h); *fallback = 1 when Ziv's test rejects it */
#include "exp_fast.h"
float dcmoe_oracle_exp_fast(float x, int* fallback) { return dcmoe_exp_fast(x, fallback, 0, 0, 0); }
float dcmoe_oracle_bf16_round(float x) { return bf16_round(x); }
void dcmoe_oracle_softmax(const float* v, int n, int bf16, float* out) { softmax_D(v, n, bf16, out); }

int dcmoe_oracle_route_k(const float* logits, const int32_t* attn_mask, int64_t T, int n_dyn, int n_fix,
                         int bf16, double top_p, double eps, int fixed_k, int64_t* top_k, int32_t* mask, float* gw,
                         float* aux_out) {
    return dcmoe_oracle_route_ex(logits, attn_mask, NULL, NULL, T, n_dyn, n_fix, bf16, top_p, eps, fixed_k, top_k, mask, gw,
                                 aux_out);
}

int dcmoe_oracle_route(const float* logits, const int32_t* attn_mask, int64_t T, int n_dyn, int n_fix,
                       int bf16, double top_p, double eps, int64_t* top_k, int32_t* mask, float* gw,
                       float* aux_out) {
    return dcmoe_oracle_route_k(logits, attn_mask, T, n_dyn, n_fix, bf16, top_p, eps, 0, top_k, mask, gw, aux_out);
}
