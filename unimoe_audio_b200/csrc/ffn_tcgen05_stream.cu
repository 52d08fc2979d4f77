// ffn_tcgen05_stream.cu -- expert FFNs for decode-sized calls (T <= 64 tokens, bf16): weight-streaming tcgen05 GEMMs.
//
// The generation loop calls the layer with T = 2N tokens (reference model.py:1149-1203).  Every m-tile then holds
// a handful of rows and the layer is bound by reading each hit expert's weights ONCE (up to 304 MB per layer).  The
// 128 x 256 tiles of ffn_tcgen05.cu fit that badly: 22 + 8 n-tiles per group leave the 148 SMs 52-68 % busy (whole
// tiles per SM).  What was measured on the way here (tools/probe_stream.cu, DCMOE_FFN_STREAM_DEBUG=1 counters):
//   * TMA streams HBM at 6.7-7.0 TB/s with >= 16 KB per ring stage; issuing a box costs the producer ~50 cycles
//     whatever its size, and a ring stage ~200 cycles of mbarrier handshake, so at this size the STAGE COUNT per SM
//     is the cost to balance, and boxes must be as large as the rows allow;
//   * tcgen05.mma instructions that accumulate into the same TMEM tile are a ~140-cycle dependent chain each.
// So here:
//   * the output columns of every m-tile are cut into 16-column granules (GEMM-1: 16 h columns = 16 gate + 16 up
//     rows of W13; GEMM-2: 16 y columns); the CTAs are dealt evenly to the hit weight groups and each group's
//     granules evenly to its CTAs: every SM makes exactly ONE pass over K and streams the same weight bytes (+-1
//     granule) whatever the number of hit experts;
//   * the WEIGHTS are the MMA's M operand (128 rows = 8 granules per M-tile) and the <= 64 token rows its N operand:
//     no tensor-core work or shared-memory read is padding; an M-tile that is only partly loaded computes lanes
//     that nobody reads.  Up to 16 granules per CTA = 2 (GEMM-2) / 4 (GEMM-1: gate and up) accumulators;
//   * runs of consecutive weight rows are fetched with the largest TMA boxes that tile them (16/32/64/128 rows),
//     one box per producer lane; the token box is 16 / 32 / 64 rows;
//   * the epilogue reads TMEM lane = output column, TMEM column = token, applies SwiGLU and the routing weight
//     (GEMM-1) and stores the valid tokens straight from registers.
// Same K order and fp32 accumulation per output element as ffn_tcgen05.cu: h is bit-identical, and so is y with
// DCMOE_FFN_STREAM_KSPLIT=0 (tested).  By default GEMM-2 spreads the four k16 steps of a stage over four accumulators
// that are summed in the epilogue -- (a0 + a1) + (a2 + a3), a fixed order: deterministic, fp32-reassociation-level
// differences from the large tiles -- because a single accumulator makes its K loop a dependent MMA chain
// (GEMM-2: 26 -> 23 us).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace dcmoe {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int GR = 16;                        // granule: 16 output columns
constexpr int BOX_BYTES = GR * BK * 2;        // one B box: 16 rows x 128 B
constexpr int RING_BYTES = 192 * 1024;
constexpr int MAX_STAGES = 16;
constexpr int SLACK_BYTES = BM * BK * 2;      // the MMA reads a full 128-row A tile from the last stage's small box
constexpr int BAR_BYTES = 512;                // full[16] empty[16] tfull[2] tempty[2] tmem slot
constexpr int SMEM_BYTES = RING_BYTES + SLACK_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 256;

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {   // K-major, SWIZZLE_128B, SBO 1024 B
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {   // kind::f16, D = f32, A = B = bf16, K-major, M = 128, N = n
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ float silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

// 32 lanes x 16 columns of fp32
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

struct StreamParams {
    int gpg;            // granules per m-tile (= per weight group): I_d / 16 (GEMM-1), H / 16 (GEMM-2)
    int max_gran;       // widest segment any CTA can get (sizes the ring stages)
    int num_kb;         // K / 64
    int w_rows;         // B rows per weight group
    int n_real;
    int split_col;      // GEMM-1: h column where the second shared expert starts (I_s)
    const dcmoe_mtile* mtiles;
    const int32_t* n_mtiles;
    const float* row_scale;
    int a_alloc;        // bytes of the token box (n_tok rows x 128)
    int n_tok;          // token rows per box = MMA N: 16 / 32 / 64
    int stage_bytes;    // a_alloc + widest segment's B boxes
    int stages;
    __nv_bfloat16* out; // h / y
    int ld_out;
    int ep_n_loc;       // expert-parallel decode: this rank owns routed experts [ep_base, ep_base + ep_n_loc) and packs
    int ep_base;        // them as weight groups 0..ep_n_loc-1 (+ the shared pair as group ep_n_loc); 0 = all groups local
    int ksplit;         // GEMM-2 only: 1 = the four k16 steps of a stage go to four accumulators (summed in the epilogue)
    unsigned long long* dbg;   // tuning (DCMOE_FFN_STREAM_DEBUG=1): per-CTA cycle counters, else nullptr
    const int32_t* small_tokens;   // GEMM-1, n_tok <= 32: token of every routed row (plan.small_tokens) -> the token box of a
                                   // routed m-tile is gathered from x with TMA gather4 (4 rows per instruction); else nullptr
};

// One segment per CTA: the CTAs are dealt to the m-tiles (= hit weight groups) as evenly as possible and the CTAs of
// one m-tile cut its granules evenly, so every CTA makes exactly ONE pass over K (stage count, the real cost at this
// size -- ~200 cycles of handshake + ~40 per TMA box -- is the same everywhere) and streams the same bytes +-1 granule.
struct Segment {
    int m;    // m-tile
    int g0;   // first granule inside the m-tile
    int ng;   // granules (0: no work)
};
__device__ __forceinline__ Segment cta_segment(int n_m, int gpg) {
    Segment s{0, 0, 0};
    const int grid = gridDim.x, c = blockIdx.x;
    if (n_m <= 0) return s;
    if (n_m > grid) n_m = grid;                     // (never with <= 17 groups; the host checks)
    const int base = grid / n_m, extra = grid % n_m;   // m-tiles [0, extra) get base + 1 CTAs
    int idx, n;
    if (c < extra * (base + 1)) {
        s.m = c / (base + 1);
        idx = c - s.m * (base + 1);
        n = base + 1;
    } else {
        const int c2 = c - extra * (base + 1);
        s.m = extra + c2 / base;
        idx = c2 - (c2 / base) * base;
        n = base;
    }
    s.g0 = (int)((int64_t)gpg * idx / n);
    s.ng = (int)((int64_t)gpg * (idx + 1) / n) - s.g0;
    return s;
}

template <bool SWIGLU>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ffn_stream_kernel(const __grid_constant__ CUtensorMap tmap_a0,   // GEMM-1: x          GEMM-2: h
                  const __grid_constant__ CUtensorMap tmap_a1,   // GEMM-1: x_packed   GEMM-2: h
                  const __grid_constant__ CUtensorMap tmap_b16,  // W13 / W2, boxes of 16 / 32 / 64 / 128 rows
                  const __grid_constant__ CUtensorMap tmap_b32, const __grid_constant__ CUtensorMap tmap_b64,
                  const __grid_constant__ CUtensorMap tmap_b128,
                  const __grid_constant__ CUtensorMap tmap_xrow,   // GEMM-1: x with a box of one row (gather4)
                  const StreamParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + RING_BYTES + SLACK_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * MAX_STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a0);
        prefetch_tmap(&tmap_a1);
        prefetch_tmap(&tmap_b16);
        prefetch_tmap(&tmap_b32);
        prefetch_tmap(&tmap_b64);
        prefetch_tmap(&tmap_b128);
        prefetch_tmap(&tmap_xrow);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // everything above touched no global memory: under programmatic dependent launch it overlapped the tail of the
    // previous kernel; now wait for it (plan / x_packed / h are its outputs) and let the next kernel's CTAs come up
    grid_dep_wait();
    grid_dep_launch();
    const long long t_start = p.dbg ? clock64() : 0;
    // m-tiles this launch works on: all of them, or (expert-parallel decode: the plan is replicated on every rank) the
    // ones whose weights this rank holds -- its routed experts and the shared pair
    __shared__ int s_local_m[kMaxDyn + 1];
    __shared__ int s_n_local;
    __shared__ dcmoe_mtile s_mt[kMaxDyn + 1];
    if (warp == 3) {
        // ONE round trip for the whole plan of a decode-sized call (<= 17 m-tiles + their count): the former
        // count -> groups -> own tile chain of dependent loads cost ~1 us at the head of each GEMM
        const int n_all = *p.n_mtiles;
        dcmoe_mtile mine = dcmoe_mtile{0, 0, 0, 0};
        if (lane <= kMaxDyn) mine = p.mtiles[lane];
        const int n_m = min(n_all, kMaxDyn + 1);
        const bool local = lane < n_m && (p.ep_n_loc == 0 || mine.group == p.n_real ||
                                         (mine.group >= p.ep_base && mine.group < p.ep_base + p.ep_n_loc));
        const unsigned bal = __ballot_sync(0xffffffffu, local);
        if (lane <= kMaxDyn) s_mt[lane] = mine;
        if (local) s_local_m[__popc(bal & ((1u << lane) - 1u))] = lane;
        if (lane == 0) s_n_local = __popc(bal);
    }
    __syncthreads();
    Segment sg = cta_segment(s_n_local, p.gpg);
    if (sg.ng > 0) sg.m = s_local_m[sg.m];
    // The WEIGHTS are the MMA's M operand (128 rows = 8 granules per M-tile) and the token rows its N operand
    // (N = the A box: 16 / 32 / 64 tokens): D[weight row, token].  Nothing of the tensor-core work or of its shared
    // memory reads is padding, and an M-tile that is only partly loaded just computes lanes nobody reads.
    // Sub-segment 0 = the first 8 granules, sub-segment 1 = the rest; GEMM-1 keeps gate and up rows of a
    // sub-segment in two M-tiles, so that TMEM lane i holds gate and up of the SAME h column (in different columns).
    // TMEM columns: accumulator a = 2 * sub + half (GEMM-1) / sub (GEMM-2) at [a * n_tok, (a + 1) * n_tok).
    constexpr int kSub = BM / GR;   // 8 granules = 128 weight rows = one MMA M-tile
    const int ng0 = min(sg.ng, kSub), ng1 = sg.ng - ng0;
    const int nb0 = SWIGLU ? 2 * ng0 : ng0, nb1 = SWIGLU ? 2 * ng1 : ng1;
    const int nb = nb0 + nb1;                         // B boxes per stage (<= 32, one per producer lane)

    if (sg.ng > 0 && warp == 0) {
        // ================= TMA producer: lane j -> B box j, lane 0 also the A box =================
        const dcmoe_mtile mt = s_mt[sg.m];
        const CUtensorMap* amap = SWIGLU ? (mt.group == p.n_real ? &tmap_a0 : &tmap_a1) : &tmap_a0;
        const int a_row = SWIGLU ? mt.a_row : mt.out_row;
        // GEMM-1, routed tile, <= 32 token rows: lane j gathers rows 4j .. 4j+3 of the token box from x (one
        // cp.async.bulk.tensor gather4 = 4 rows x 128 B, the same swizzled layout a row box produces; ~69 cycles of TMA
        // issue each, profiles/r02_probe_gather4.txt -- affordable for 4 - 8 per ~2000-cycle stage) instead of reading
        // a packed copy: the front end then skips the row gather.  Rows past mt.rows gather token 0 (never stored).
        const bool gather_a = SWIGLU && p.small_tokens != nullptr && mt.group != p.n_real;
        int tok[4] = {0, 0, 0, 0};
        if (gather_a && lane * 4 < p.n_tok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = lane * 4 + q;
                tok[q] = r < mt.rows ? p.small_tokens[mt.out_row + r] : 0;
            }
        }
        // The B tile of a stage is nb granule slots of 16 rows x 128 B.  Runs of consecutive weight rows are fetched
        // with the largest boxes that tile them (16 / 32 / 64 / 128 rows): issuing a TMA box costs ~50 cycles
        // whatever its size, and at one box per granule that issue rate, not HBM, bounds the kernel.  Lane 0 builds
        // the box list once; lane j issues box j of every stage.
        __shared__ int s_box[32][3];   // {first weight row, byte offset inside the B tile, log2(rows / 16)}
        __shared__ int s_nbox;
        if (lane == 0) {
            int n = 0;
            auto add_run = [&](int row0, int nrows, int dst) {
                while (nrows > 0) {
                    const int sz = min(128, 1 << (31 - __clz(nrows)));
                    s_box[n][0] = row0;
                    s_box[n][1] = dst;
                    s_box[n][2] = 31 - __clz(sz >> 4);
                    ++n;
                    row0 += sz;
                    dst += sz * 128;
                    nrows -= sz;
                }
            };
            const int wgrp = p.ep_n_loc == 0 ? mt.group : (mt.group == p.n_real ? p.ep_n_loc : mt.group - p.ep_base);
            const int wbase = wgrp * p.w_rows;
            int dst = 0;
            for (int sub = 0; sub < 2; ++sub) {
                const int gs = sg.g0 + sub * ng0, ngs = sub ? ng1 : ng0;
                if (ngs == 0) continue;
                if (SWIGLU) {
                    for (int half = 0; half < 2; ++half) {   // W13: blocks of 64 gate rows then 64 up rows
                        int c = gs * GR;
                        const int ce = (gs + ngs) * GR;
                        while (c < ce) {
                            const int run_end = min(ce, ((c >> 6) + 1) << 6);
                            add_run(wbase + (c >> 6) * 128 + (c & 63) + half * 64, run_end - c, dst);
                            dst += (run_end - c) * 128;
                            c = run_end;
                        }
                    }
                } else {
                    add_run(wbase + gs * GR, ngs * GR, dst);
                    dst += ngs * GR * 128;
                }
            }
            s_nbox = n;
        }
        __syncwarp();
        const int n_box = s_nbox;
        const int b_row = lane < n_box ? s_box[lane][0] : 0;
        const int b_off = lane < n_box ? s_box[lane][1] : 0;
        const int b_sel = lane < n_box ? s_box[lane][2] : 0;
        const CUtensorMap* bmap = b_sel == 0 ? &tmap_b16 : (b_sel == 1 ? &tmap_b32 : (b_sel == 2 ? &tmap_b64 : &tmap_b128));
        int stage = 0;
        uint32_t phase = 0;
        long long d_wait = 0, d_issue = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const long long t1 = p.dbg ? clock64() : 0;
            const uint32_t dst = smem_base + stage * p.stage_bytes;
            if (lane == 0) {
                mbar_expect_tx(full_bar(stage), (uint32_t)(p.a_alloc + nb * BOX_BYTES));
                if (!gather_a) tma_load_2d(dst, amap, kb * BK, a_row, full_bar(stage));
            }
            __syncwarp();
            if (gather_a && lane * 4 < p.n_tok)
                tma_gather4_2d(dst + lane * 512, &tmap_xrow, kb * BK, tok[0], tok[1], tok[2], tok[3], full_bar(stage));
            if (lane < n_box) tma_load_2d(dst + p.a_alloc + b_off, bmap, kb * BK, b_row, full_bar(stage));
            if (p.dbg) {
                d_wait += t1 - t0;
                d_issue += clock64() - t1;
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (p.dbg && lane == 0) {
            p.dbg[blockIdx.x * 8 + 0] = d_wait;
            p.dbg[blockIdx.x * 8 + 1] = d_issue;
            p.dbg[blockIdx.x * 8 + 5] = p.num_kb;
            p.dbg[blockIdx.x * 8 + 6] = clock64() - t_start;   // producer done
        }
    } else if (sg.ng > 0 && warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc(p.n_tok);
        int stage = 0;
        uint32_t phase = 0;
        long long d_wait = 0, d_issue = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const long long t1 = p.dbg ? clock64() : 0;
            if (lane == 0) {
                const uint32_t tok_addr = smem_base + stage * p.stage_bytes;
                const uint64_t tdesc = make_smem_desc(tok_addr);
                const uint32_t w_addr = tok_addr + p.a_alloc;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    if (!SWIGLU && p.ksplit) {
                        // MMAs into one accumulator are a ~140-cycle dependent chain each; GEMM-2 has a single M-tile
                        // per sub-segment, so its K loop is bound by that chain.  Four accumulators (one per k16 step
                        // of a stage), summed in the epilogue, make the steps independent.
                        const uint32_t acc_k = (uint32_t)(kb != 0);
                        umma_bf16(tmem_base + (uint32_t)(k * p.n_tok), make_smem_desc(w_addr) + (uint64_t)(2 * k),
                                  tdesc + (uint64_t)(2 * k), idesc, acc_k);
                        if (ng1 > 0)
                            umma_bf16(tmem_base + (uint32_t)((4 + k) * p.n_tok), make_smem_desc(w_addr + nb0 * BOX_BYTES) + (uint64_t)(2 * k),
                                      tdesc + (uint64_t)(2 * k), idesc, acc_k);
                        continue;
                    }
                    const uint32_t accum = (uint32_t)((kb | k) != 0);
                    // M-tiles in B-tile order: [sub 0: gate | up] [sub 1: gate | up]  (GEMM-2: [sub 0] [sub 1])
                    umma_bf16(tmem_base, make_smem_desc(w_addr) + (uint64_t)(2 * k), tdesc + (uint64_t)(2 * k), idesc, accum);
                    if (SWIGLU)
                        umma_bf16(tmem_base + p.n_tok, make_smem_desc(w_addr + ng0 * BOX_BYTES) + (uint64_t)(2 * k),
                                  tdesc + (uint64_t)(2 * k), idesc, accum);
                    if (ng1 > 0) {
                        const uint32_t w1 = w_addr + nb0 * BOX_BYTES;
                        umma_bf16(tmem_base + (SWIGLU ? 2 : 1) * p.n_tok, make_smem_desc(w1) + (uint64_t)(2 * k),
                                  tdesc + (uint64_t)(2 * k), idesc, accum);
                        if (SWIGLU)
                            umma_bf16(tmem_base + 3 * p.n_tok, make_smem_desc(w1 + ng1 * BOX_BYTES) + (uint64_t)(2 * k),
                                      tdesc + (uint64_t)(2 * k), idesc, accum);
                    }
                }
                umma_commit(empty_bar(stage));
                if (kb == p.num_kb - 1) umma_commit(tfull_bar);
                if (p.dbg) {
                    d_wait += t1 - t0;
                    d_issue += clock64() - t1;
                }
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (p.dbg && lane == 0) {
            p.dbg[blockIdx.x * 8 + 2] = d_wait;
            p.dbg[blockIdx.x * 8 + 3] = d_issue;
            p.dbg[blockIdx.x * 8 + 7] = clock64() - t_start;   // last MMA issued
        }
    } else if (sg.ng > 0 && warp >= 4) {
        // ================= epilogue: TMEM lane = output column, TMEM column = token =================
        const int wq = warp - 4;  // TMEM lane quarter == warp_id % 4
        const dcmoe_mtile mt = s_mt[sg.m];
        const bool shared_grp = mt.group == p.n_real;
        mbar_wait(tfull_bar, 0u);
        tc_fence_after();
        for (int sub = 0; sub < 2; ++sub) {
            const int ngs = sub ? ng1 : ng0;
            const int c_local = wq * 32 + lane;                       // column inside the sub-segment
            if (wq * 32 >= ngs * GR) continue;                        // (warp-uniform) no valid lane in this quarter
            const bool lane_ok = c_local < ngs * GR;
            const int col = (sg.g0 + sub * ng0) * GR + c_local;        // h / y column of this lane
            const bool ks4 = !SWIGLU && p.ksplit;
            const uint32_t t_acc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((SWIGLU ? 2 : (ks4 ? 4 : 1)) * sub * p.n_tok);
            const int sel = (SWIGLU && shared_grp && col >= p.split_col) ? 1 : 0;
            for (int n0 = 0; n0 < mt.rows; n0 += 16) {                // 16 tokens per TMEM load
                uint32_t g[16], u[16];
                tmem_ld16(t_acc + (uint32_t)n0, g);
                if (SWIGLU) tmem_ld16(t_acc + (uint32_t)(p.n_tok + n0), u);
                if (ks4) {   // (a0 + a1) + (a2 + a3), fixed order
                    uint32_t a2[16], a3[16];
                    tmem_ld16(t_acc + (uint32_t)(p.n_tok + n0), u);
                    tmem_ld16(t_acc + (uint32_t)(2 * p.n_tok + n0), a2);
                    tmem_ld16(t_acc + (uint32_t)(3 * p.n_tok + n0), a3);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        g[e] = __float_as_uint(__fadd_rn(__fadd_rn(__uint_as_float(g[e]), __uint_as_float(u[e])),
                                                         __fadd_rn(__uint_as_float(a2[e]), __uint_as_float(a3[e]))));
                }
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int n = n0 + e;
                    if (n < mt.rows && lane_ok) {
                        const int64_t r = (int64_t)mt.out_row + n;
                        float v = __uint_as_float(g[e]);
                        if (SWIGLU) v = silu_mul(v, __uint_as_float(u[e])) * p.row_scale[2 * r + sel];
                        p.out[r * p.ld_out + col] = __float2bfloat16_rn(v);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 4] = clock64() - t_start;
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

// GEMM-1 gathers its routed token rows straight from x (TMA gather4) when the token box has at most 32 rows
bool ffn_stream_gathers_from_x(int64_t T) {
    static const bool on = !(getenv("DCMOE_FFN_STREAM_GATHER") && getenv("DCMOE_FFN_STREAM_GATHER")[0] == '0');
    return on && T > 0 && T <= 32;
}

bool ffn_stream_applicable(int64_t T, const dcmoe_config* cfg, const dcmoe_sizes& sz, int max_ctas, int ep_n_loc) {
    // every m-tile must be a whole weight group with at most 64 rows: T <= 64 (one shared tile, one tile per hit
    // expert), and the widest segment a CTA can get must fit the two accumulators / 32 producer lanes
    if (!(cfg->dtype == DCMOE_BF16 && T > 0 && T <= 64 && sz.t_pad == BM && cfg->dynamic_intermediate_size % GR == 0 &&
          cfg->shared_intermediate_size % GR == 0 && cfg->hidden_size % GR == 0))
        return false;
    int n_ctas = device_sm_count();
    if (max_ctas > 0 && max_ctas < n_ctas) n_ctas = max_ctas;
    const int G = (ep_n_loc > 0 ? ep_n_loc : cfg->n_real) + 1;   // most weight groups one launch can work on
    if (n_ctas < G) return false;
    const int per_group = n_ctas / G;
    return ceil_div(cfg->dynamic_intermediate_size / GR, per_group) <= 16 && ceil_div(cfg->hidden_size / GR, per_group) <= 16;
}

int launch_ffn_tcgen05_stream(const void* x, const void* x_packed, const void* w13, const void* w2, const float* row_scale,
                              int64_t T, int64_t row_capacity, const dcmoe_config* cfg, const dcmoe_sizes& sz, PlanView pv,
                              void* h, void* y, int phase, int max_ctas, int ep_n_loc, int ep_rank, cudaStream_t stream) {
    if (ep_n_loc < 0 || (ep_n_loc > 0 && (cfg->n_real % ep_n_loc != 0 || ep_rank < 0 || ep_rank >= cfg->n_real / ep_n_loc))) {
        set_error("weight-streaming FFN: bad expert-parallel layout (%d local experts, rank %d)", ep_n_loc, ep_rank);
        return DCMOE_ERR_INVALID;
    }
    if (!ffn_stream_applicable(T, cfg, sz, max_ctas, ep_n_loc)) {
        set_error("weight-streaming FFN needs bf16, 1 <= T <= 64 and enough CTAs per weight group (got T = %lld)", (long long)T);
        return DCMOE_ERR_INVALID;
    }
    const int H = cfg->hidden_size, Id = cfg->dynamic_intermediate_size;
    const int G = (ep_n_loc > 0 ? ep_n_loc : cfg->n_real) + 1;   // weight groups in the packs handed in
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        int rc = check_cuda(cudaFuncSetAttribute(ffn_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                            "cudaFuncSetAttribute(stream gemm1)");
        if (rc) { attr_once.reset_current(); return rc; }
        rc = check_cuda(cudaFuncSetAttribute(ffn_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                        "cudaFuncSetAttribute(stream gemm2)");
        if (rc) { attr_once.reset_current(); return rc; }
    }
    const int a_box = T <= 16 ? 16 : (T <= 32 ? 32 : 64);
    CUtensorMap m_x, m_xp, m_h, m_xrow, m_w13[4], m_w2[4];
    int rc;
    const int64_t packed_rows = row_capacity - sz.t_pad;
    if ((rc = make_tensor_map_bf16(&m_x, x, T, H, a_box))) return rc;
    if ((rc = make_tensor_map_bf16(&m_xp, x_packed, packed_rows > 0 ? packed_rows : 1, H, a_box))) return rc;
    if ((rc = make_tensor_map_bf16(&m_h, h, row_capacity, Id, a_box))) return rc;
    if ((rc = make_tensor_map_bf16(&m_xrow, x, T, H, 1))) return rc;
    for (int i = 0; i < 4; ++i) {
        if ((rc = make_tensor_map_bf16(&m_w13[i], w13, (int64_t)G * 2 * Id, H, GR << i))) return rc;
        if ((rc = make_tensor_map_bf16(&m_w2[i], w2, (int64_t)G * H, Id, GR << i))) return rc;
    }

    int n_ctas = device_sm_count();
    if (max_ctas > 0 && max_ctas < n_ctas) n_ctas = max_ctas;

    StreamParams p1, p2;
    p1.gpg = Id / GR;
    const int ctas_per_group = n_ctas / G;   // fewest CTAs a hit group can get (all G groups hit)
    p1.max_gran = (int)ceil_div(p1.gpg, ctas_per_group);
    p1.num_kb = H / BK;
    p1.w_rows = 2 * Id;
    p1.n_real = cfg->n_real;
    p1.split_col = cfg->shared_intermediate_size;
    p1.mtiles = pv.mtiles;
    p1.n_mtiles = pv.n_mtiles;
    p1.row_scale = row_scale;
    p1.a_alloc = a_box * BK * 2;
    p1.n_tok = a_box;
    p1.stage_bytes = p1.a_alloc + 2 * p1.max_gran * BOX_BYTES;
    p1.stages = std::min(MAX_STAGES, RING_BYTES / p1.stage_bytes);
    p1.out = static_cast<__nv_bfloat16*>(h);
    p1.ld_out = Id;
    p2 = p1;
    p2.gpg = H / GR;
    p2.max_gran = (int)ceil_div(p2.gpg, ctas_per_group);
    p2.num_kb = Id / BK;
    p2.w_rows = H;
    p2.stage_bytes = p2.a_alloc + p2.max_gran * BOX_BYTES;
    p2.stages = std::min(MAX_STAGES, RING_BYTES / p2.stage_bytes);
    p2.out = static_cast<__nv_bfloat16*>(y);
    p2.ld_out = H;
    p1.ksplit = 0;
    p1.small_tokens = ffn_stream_gathers_from_x(T) ? pv.small_tokens : nullptr;
    p2.small_tokens = nullptr;
    p1.ep_n_loc = p2.ep_n_loc = ep_n_loc;
    p1.ep_base = p2.ep_base = ep_rank * ep_n_loc;
    {
        const char* e = getenv("DCMOE_FFN_STREAM_KSPLIT");   // 0: one accumulator (y bit-identical to the large tiles)
        p2.ksplit = (e && e[0] == '0') ? 0 : 1;
    }

    static unsigned long long* dbg = nullptr;
    const bool debug = getenv("DCMOE_FFN_STREAM_DEBUG") != nullptr;
    if (debug && !dbg) cudaMalloc(&dbg, 2 * 256 * 8 * sizeof(unsigned long long));
    p1.dbg = p2.dbg = nullptr;
    if (debug) {
        cudaMemsetAsync(dbg, 0, 2 * 256 * 8 * sizeof(unsigned long long), stream);
        p1.dbg = dbg;
        p2.dbg = dbg + 256 * 8;
    }
    dim3 grid((unsigned)n_ctas), block(NUM_THREADS);
    if (phase != 2) {
        if ((rc = check_cuda(launch_kernel(ffn_stream_kernel<true>, grid, block, SMEM_BYTES, stream, pdl_enabled(), m_x, m_xp,
                                           m_w13[0], m_w13[1], m_w13[2], m_w13[3], m_xrow, p1),
                             "ffn_stream_kernel<SwiGLU> launch")))
            return rc;
    }
    if (phase != 1 &&
        (rc = check_cuda(launch_kernel(ffn_stream_kernel<false>, grid, block, SMEM_BYTES, stream, pdl_enabled(), m_h, m_h,
                                       m_w2[0], m_w2[1], m_w2[2], m_w2[3], m_xrow, p2),
                         "ffn_stream_kernel<down> launch")))
        return rc;
    if (debug) {   // tuning only: synchronises
        static int printed = 0;
        cudaStreamSynchronize(stream);
        unsigned long long host[2 * 256 * 8];
        cudaMemcpy(host, dbg, sizeof(host), cudaMemcpyDeviceToHost);
        if (printed++ < 6)
            for (int g = 0; g < 2; ++g) {
                if ((g == 0 && phase == 2) || (g == 1 && phase == 1)) continue;
                double sum[8] = {0}, mx[8] = {0};
                for (int c = 0; c < n_ctas; ++c)
                    for (int k = 0; k < 8; ++k) {
                        const double v = (double)host[(g * 256 + c) * 8 + k];
                        sum[k] += v;
                        mx[k] = std::max(mx[k], v);
                    }
                fprintf(stderr, "stream gemm%d T=%lld stages=%d stage_bytes=%d max_gran=%d | per CTA avg/max: stages %.0f/%.0f  "
                        "total %.0f/%.0f cyc, producer done at %.0f, last MMA issued at %.0f | per stage avg: prod wait %.0f issue %.0f | mma wait %.0f issue %.0f cyc\n",
                        g + 1, (long long)T, g ? p2.stages : p1.stages, g ? p2.stage_bytes : p1.stage_bytes,
                        g ? p2.max_gran : p1.max_gran, sum[5] / n_ctas, mx[5], sum[4] / n_ctas, mx[4], sum[6] / n_ctas, sum[7] / n_ctas, sum[0] / sum[5],
                        sum[1] / sum[5], sum[2] / sum[5], sum[3] / sum[5]);
            }
    }
    return check_cuda(cudaGetLastError(), "ffn_stream_kernel<down> launch");
}

}  // namespace dcmoe
