// Tensor-core GEMM for the wide hidden layers: tcgen05.mma (kind::tf32) with TMEM accumulators, TMA operand
// staging through an mbarrier ring, warp-specialised roles.  sm_100a only.
//
//   C[g][m][n] = epi( sum_k A[g][m][k] * B[g][n][k] )          A, B row-major with k contiguous ("K-major")
//
// One CTA = one 128-row tile of A x all N (<= 256) columns, for one k-split.  The fp32 accumulator tile lives in
// TMEM (128 lanes x N columns).  Precision modes:
//   passes = 1 : one TF32 MMA per product (10-bit mantissa operands, fp32 accumulate)  -- "fast" mode
//   passes = 3 : lo = x - tf32_trunc(x) is written next to each raw tile in shared memory and three MMAs
//                (hi*hi + lo*hi + hi*lo) are accumulated: ~2^-21 relative error per product -- fp32-parity mode.
// Optional fused outputs: row-major C, transposed C^T (the K-major operand of the next weight-gradient GEMM),
// bias + ReLU / ReLU-mask epilogues, and row sums of A (bias gradients) computed by the tensor core against a
// tile of ones.
//
// Replaces the 256x256 nn.Linear forward / dgrad / wgrad GEMMs of both critics over the 7936-row CQL critic batch
// (reference: nets/mlp.py:22 forward; autograd backward of the same layers; cql.py:133-190).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "orlk_tcgen.cuh"
using namespace orlk;
using namespace orlk::tcg;

namespace {

constexpr int BM = 128, BN_MAX = 256, BK = 32;          // BK fp32 = 128 bytes = one SWIZZLE_128B row
constexpr int A_BYTES = BM * BK * 4;                      // 16 KB
constexpr int B_BYTES = BN_MAX * BK * 4;                  // 32 KB
constexpr int ONES_BYTES = 16 * BK * 4;                   // 2 KB
constexpr int TMEM_COLS = 512;
constexpr int ROWSUM_COL = 256;
constexpr int ATM_COL = 288;                              // A-from-TMEM mode: per stage [hi 32 cols | lo 32 cols] from here
constexpr int MAX_STAGES = 8;
constexpr int FIXED_SMEM = 8192;                          // ones tile, barriers, staged bias, mask bits
constexpr int NUM_THREADS = 320;                          // warp0 TMA, warp1 MMA, warps2-5 split + epilogue, warps6-9 mask

struct TcParams {
    float* C; int64_t ldc, c_gs, c_split_stride;
    float* CT; int64_t ldct, ct_gs;
    const float* bias; int64_t bias_gs;
    const float* aux; int64_t ldaux, aux_gs;
    float* rowsum; int64_t rowsum_gs, rowsum_split_stride;
    int M, N, K, G, epi, k_splits, slabs_per_split, tiles_m, tiles_n, NT;   // NT = output columns per CTA (n-tile)
    int stages, stage_bytes;   // depth of the TMA ring and bytes per stage (depend on NT and the precision mode)
    const float* gen_row; int64_t gen_row_gs;   // rank-1 operand generator (see OrlkTcGemm), NULL = off
    const float* gen_col; int64_t gen_col_gs;
    const float* Bm; int64_t ldbm, bm_gs;       // B_MANUAL: B rows are not TMA-able (pitch not a multiple of 16 bytes);
    int b_manual;                               //   K <= 32 then, and the splitter warps fill the single B tile themselves
    int a_shared, b_shared;                     // one A / B for all groups (group stride 0): that map has a single plane
    int a_mn, b_mn;            // operand stored [k][m] / [k][n] (MN-major) instead of [m][k] / [n][k]
    int r3;                    // 3-pass + A-from-TMEM: ONE A buffer (it is dead as soon as the splitter has moved it to TMEM)
                               // and a THREE-stage ring of [B | B lo]: the 2-stage ring was bound by its own chain latency
    int late_trigger;          // programmatic launch of the next kernel after the mainloop instead of at kernel start
    int a_tmem;                // the splitter moves A (and its lo part) into tensor memory; the MMAs read A from there
    int c_tma;                 // C leaves through TMA stores (tmC valid)
    int trace_mode;            // 0: slots 8..15 = k-slab landed, 1: slots 8..15 = TMA for k-slab issued
    unsigned long long* trace; // profiling aid (orlk_tc_set_trace): 16 clock stamps per CTA, NULL in normal operation
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TC_STAMP(slot)                                                       \
    do {                                                                     \
        if (p.trace != nullptr) p.trace[(int64_t)blockIdx.x * 16 + (slot)] = (unsigned long long)clock64(); \
    } while (0)


// lo = x - trunc_tf32(x) for float4s [i_lo, i_hi) of a raw operand tile, by 128 threads (t = 0..127)
__device__ __forceinline__ void split_range(const float4* __restrict__ raw, float4* __restrict__ lo, int i_lo, int i_hi, int t) {
    for (int i0 = i_lo; i0 < i_hi; i0 += 128 * 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + t + 128 * j;
            v[j] = i < i_hi ? raw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + t + 128 * j;
            if (i < i_hi) lo[i] = lo_tf32(v[j]);
        }
    }
}

template <int PASSES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmC, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment
    // (offset arithmetic on the __shared__ array keeps the address space known to the compiler: LDS/STS, not generic)
    uint8_t* fixed = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ones = fixed;                                               // [16 x 32] fp32 ones (bias-gradient MMA)
    uint64_t* bars = reinterpret_cast<uint64_t*>(fixed + ONES_BYTES);
    uint64_t* full = bars;                          // [MAX_STAGES]  TMA bytes landed
    uint64_t* splitb = bars + MAX_STAGES;           // [MAX_STAGES]  hi/lo split done (PASSES == 3)
    uint64_t* empty = bars + 2 * MAX_STAGES;        // [MAX_STAGES]  MMAs that read the stage have completed
    uint64_t* accum = bars + 3 * MAX_STAGES;        // accumulator tile complete
    uint64_t* maskbar = bars + 3 * MAX_STAGES + 1;  // mask tile complete (128 arrivals)
    uint64_t* a_full = bars + 3 * MAX_STAGES + 2;   // r3: the A tile of the current slab has landed
    uint64_t* a_free = bars + 3 * MAX_STAGES + 3;   // r3: the splitter has moved it to TMEM (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
    float* bias_s = reinterpret_cast<float*>(fixed + ONES_BYTES + 256);  // [BN_MAX] bias staged once per CTA
    uint32_t* mask_s = reinterpret_cast<uint32_t*>(bias_s + BN_MAX);     // [BM][BN_MAX/32] ReLU-mask bits of the tile
    uint8_t* smem = fixed + FIXED_SMEM;                                  // the operand ring
    const int STAGES = p.stages, STAGE_BYTES = p.stage_bytes;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);    // provably warp-uniform
    const int lane = threadIdx.x & 31;
    int idx = blockIdx.x;
    const int split = idx % p.k_splits;
    idx /= p.k_splits;
    const int tile_n = idx % p.tiles_n;
    idx /= p.tiles_n;
    const int tile_m = idx % p.tiles_m;
    const int g = idx / p.tiles_m;
    const int NT = p.NT;                       // this CTA owns output columns [n0, n0 + NT)
    const int n0 = tile_n * NT;
    const int n_lim = min(NT, p.N - n0);       // valid output columns of this tile (N need not be a multiple of NT)
    const int total_slabs = (p.K + BK - 1) / BK;
    const int slab0 = split * p.slabs_per_split;
    const int nslabs = min(p.slabs_per_split, total_slabs - slab0);
    const bool want_rowsum = p.rowsum != nullptr && tile_n == 0;
    const bool gen = p.gen_row != nullptr;
    const int y_from = (p.epi == ORLK_EPI_RELU_MASK) ? p.stages : 0;   // first k-slab the mask warps help to split
    const bool split_runs = PASSES == 3 || gen || p.b_manual != 0 || p.a_tmem != 0;     // the splitter warps process every stage
    constexpr int ATM_STRIDE = PASSES == 3 ? 64 : 32;

    const int b_bytes = NT * BK * 4;
    const bool r3 = p.r3 != 0;                 // ring layout: [A buffer] [stage 0: B | B lo] [stage 1] [stage 2]
    auto a_raw = [&](int s) { return r3 ? smem : smem + s * STAGE_BYTES; };
    auto b_raw = [&](int s) { return r3 ? smem + A_BYTES + s * STAGE_BYTES : smem + s * STAGE_BYTES + A_BYTES; };
    auto a_lo = [&](int s) { return smem + s * STAGE_BYTES + A_BYTES + b_bytes; };
    auto b_lo = [&](int s) { return r3 ? smem + A_BYTES + s * STAGE_BYTES + b_bytes : smem + s * STAGE_BYTES + 2 * A_BYTES + b_bytes; };

    // Prologue that touches no global data: runs while the previous kernel of the stream is still finishing.
    if (!p.late_trigger) orlk::pdl_trigger();
    if (threadIdx.x == 0) TC_STAMP(0);
    if (threadIdx.x == 32) {                   // descriptor fetch off the critical path (~0.3 us on first use)
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        if (!p.b_manual) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (p.c_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&splitb[s]), (PASSES == 3 && !p.b_manual) ? 256 : 128);     // both warp groups split B
            mbar_init(smem_u32(&empty[s]), 1);
        }
        mbar_init(smem_u32(accum), 1);
        mbar_init(smem_u32(a_full), 1);
        mbar_init(smem_u32(a_free), 128);
        mbar_init(smem_u32(maskbar), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2 && warp < 6 && want_rowsum) {
        float* o = reinterpret_cast<float*>(ones);
        for (int i = threadIdx.x - 64; i < ONES_BYTES / 4; i += 128) o[i] = 1.0f;
        fence_proxy_async();
    }
    orlk::pdl_wait();                          // operands, bias and aux come from earlier kernels
    if (threadIdx.x == 0) TC_STAMP(1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    if (threadIdx.x == 0) TC_STAMP(2);

    if (warp == 0) {
      if (elect_one()) {
        // ------------------------------------------------------------------ TMA producer
        const uint32_t tx_bytes = (r3 ? 0u : (uint32_t)A_BYTES) + (p.b_manual ? 0u : (uint32_t)NT * BK * 4);
        for (int it = 0; it < nslabs; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(smem_u32(&empty[s]), ph ^ 1);
            mbar_expect_tx(smem_u32(&full[s]), tx_bytes);
            const int k0 = (slab0 + it) * BK;
            if (r3) {           // B of this slab first (its stage is free), then A once the splitter is done with the previous A
                if (p.b_mn) {
                    for (int b = 0; b < NT / 32; ++b)
                        tma_load_3d(smem_u32(b_raw(s)) + b * 4096, &tmB, smem_u32(&full[s]), n0 + 32 * b, k0, p.b_shared ? 0 : g);
                } else tma_load_3d(smem_u32(b_raw(s)), &tmB, smem_u32(&full[s]), k0, n0, p.b_shared ? 0 : g);
                mbar_wait(smem_u32(a_free), (it & 1) ^ 1);
                mbar_expect_tx(smem_u32(a_full), A_BYTES);
                if (p.a_mn) {
#pragma unroll
                    for (int b = 0; b < BM / 32; ++b)
                        tma_load_3d(smem_u32(a_raw(0)) + b * 4096, &tmA, smem_u32(a_full), tile_m * BM + 32 * b, k0, p.a_shared ? 0 : g);
                } else tma_load_3d(smem_u32(a_raw(0)), &tmA, smem_u32(a_full), k0, tile_m * BM, p.a_shared ? 0 : g);
                continue;
            }
            if (p.a_mn) {       // MN-major: one 32 (m) x 32 (k) box per 4 KB block
#pragma unroll
                for (int b = 0; b < BM / 32; ++b)
                    tma_load_3d(smem_u32(a_raw(s)) + b * 4096, &tmA, smem_u32(&full[s]), tile_m * BM + 32 * b, k0, p.a_shared ? 0 : g);
            } else tma_load_3d(smem_u32(a_raw(s)), &tmA, smem_u32(&full[s]), k0, tile_m * BM, p.a_shared ? 0 : g);
            if (p.b_mn) {
                for (int b = 0; b < NT / 32; ++b)
                    tma_load_3d(smem_u32(b_raw(s)) + b * 4096, &tmB, smem_u32(&full[s]), n0 + 32 * b, k0, p.b_shared ? 0 : g);
            } else if (!p.b_manual) tma_load_3d(smem_u32(b_raw(s)), &tmB, smem_u32(&full[s]), k0, n0, p.b_shared ? 0 : g);
            if (p.trace_mode == 1 && it < 8) TC_STAMP(8 + it);
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        const uint32_t majors = (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);     // a_major / b_major bits
        const uint32_t idesc = instr_desc_tf32(BM, NT) | majors;
        const uint32_t idesc_rs = instr_desc_tf32(BM, 16) | (p.a_mn ? (1u << 15) : 0u);       // the ones tile is K-major
        // start-address step per UMMA_K = 8: 32 bytes along a K-major row, one 1024-byte k group of an MN-major tile
        const uint64_t a_kstep = p.a_mn ? 64u : 2u, b_kstep = p.b_mn ? 64u : 2u;
        const uint64_t ones_desc = smem_desc_sw128(smem_u32(ones));
        if (p.trace != nullptr && p.trace_mode == 2 && nslabs <= STAGES) {   // experiment: issue only once all slabs landed
            for (int it = 0; it < nslabs; ++it) mbar_wait(smem_u32(split_runs ? &splitb[it] : &full[it]), 0);
            TC_STAMP(3);
        }
        for (int it = 0; it < nslabs; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(smem_u32(split_runs ? &splitb[s] : &full[s]), ph);
            tc_fence_after();
            if (it == 0) TC_STAMP(4);
            if (it == nslabs - 1) TC_STAMP(5);
            if (PASSES == 1 && p.trace_mode == 0 && it < 8) TC_STAMP(8 + it);
            if (p.a_tmem) {
                const uint32_t at = tmem_base + ATM_COL + s * ATM_STRIDE;
                const uint64_t bd = p.b_mn ? smem_desc_sw128_mn(smem_u32(b_raw(s))) : smem_desc_sw128(smem_u32(b_raw(s)));
                const uint64_t bdl = p.b_mn ? smem_desc_sw128_mn(smem_u32(b_lo(s))) : smem_desc_sw128(smem_u32(b_lo(s)));
                const uint32_t idesc_ts = instr_desc_tf32(BM, NT) | (p.b_mn ? (1u << 16) : 0u);      // A from TMEM is never transposed
                const uint32_t idesc_rs_ts = instr_desc_tf32(BM, 16);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
                    const uint64_t kb = (uint64_t)k * b_kstep;
                    umma_tf32_ts(tmem_base, at + 8 * k, bd + kb, idesc_ts, acc);
                    if (PASSES == 3) {
                        umma_tf32_ts(tmem_base, at + 32 + 8 * k, bd + kb, idesc_ts, 1u);
                        umma_tf32_ts(tmem_base, at + 8 * k, bdl + kb, idesc_ts, 1u);
                    }
                    if (want_rowsum) {
                        umma_tf32_ts(tmem_base + ROWSUM_COL, at + 8 * k, ones_desc, idesc_rs_ts, acc);
                        if (PASSES == 3) umma_tf32_ts(tmem_base + ROWSUM_COL, at + 32 + 8 * k, ones_desc, idesc_rs_ts, 1u);
                    }
                }
                umma_commit(smem_u32(&empty[s]));
                continue;
            }
            const uint64_t ad = p.a_mn ? smem_desc_sw128_mn(smem_u32(a_raw(s))) : smem_desc_sw128(smem_u32(a_raw(s)));
            const uint64_t bd = p.b_mn ? smem_desc_sw128_mn(smem_u32(b_raw(s))) : smem_desc_sw128(smem_u32(b_raw(s)));
            const uint64_t adl = p.a_mn ? smem_desc_sw128_mn(smem_u32(a_lo(s))) : smem_desc_sw128(smem_u32(a_lo(s)));
            const uint64_t bdl = p.b_mn ? smem_desc_sw128_mn(smem_u32(b_lo(s))) : smem_desc_sw128(smem_u32(b_lo(s)));
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {              // UMMA_K = 8 for tf32
                const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
                const uint64_t ka = (uint64_t)k * a_kstep, kb = (uint64_t)k * b_kstep;
                umma_tf32(tmem_base, ad + ka, bd + kb, idesc, acc);
                if (PASSES == 3) {
                    umma_tf32(tmem_base, adl + ka, bd + kb, idesc, 1u);
                    umma_tf32(tmem_base, ad + ka, bdl + kb, idesc, 1u);
                }
                if (want_rowsum) {
                    umma_tf32(tmem_base + ROWSUM_COL, ad + ka, ones_desc, idesc_rs, acc);
                    if (PASSES == 3) umma_tf32(tmem_base + ROWSUM_COL, adl + ka, ones_desc, idesc_rs, 1u);
                }
            }
            umma_commit(smem_u32(&empty[s]));               // frees the stage when these MMAs have completed
        }
        umma_commit(smem_u32(accum));
      }
    } else if (warp >= 6) {
        // ------------------------------------------------------------------ ReLU-mask builder (overlaps the mainloop)
        // Each lane loads float4s (512 contiguous bytes per warp instruction, 8 rows in flight); component e of
        // 128-column group q gives one ballot word:  bit j of mask_s[row][4q+e]  <=>  aux[m][128q + 4j + e] > 0.
        if (p.trace != nullptr && p.trace_mode == 2 && warp == 6 && lane == 0) {     // experiment: landing observer
            for (int it = 0; it < nslabs && it < 8; ++it) {
                mbar_wait(smem_u32(&full[it % STAGES]), (it / STAGES) & 1);
                TC_STAMP(8 + it);
            }
        }
        if (PASSES == 3 && !p.b_manual) {       // their arrivals for the slabs they do not split (first phase of those stages)
            for (int it = 0; it < y_from && it < nslabs; ++it) mbar_arrive(smem_u32(&splitb[it]));
        }
        if (p.epi == ORLK_EPI_RELU_MASK) {
            const int w = warp - 6;
            const float* auxg = p.aux + (int64_t)g * p.aux_gs + n0;
            const bool vec = (p.ldaux % 4 == 0) && (p.aux_gs % 4 == 0) && aligned16(p.aux) && (n_lim % 4 == 0);
            constexpr int RB = 8;                                   // rows per batch
            for (int r0 = 0; r0 < 32; r0 += RB) {
                float4 a[RB][2];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const int m = tile_m * BM + w * 32 + r0 + rr;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int n = q * 128 + 4 * lane;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (m < p.M && n < n_lim) {
                            const float* src = auxg + (int64_t)m * p.ldaux + n;
                            if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
                            else {
                                v.x = __ldg(src);
                                if (n + 1 < n_lim) v.y = __ldg(src + 1);
                                if (n + 2 < n_lim) v.z = __ldg(src + 2);
                                if (n + 3 < n_lim) v.w = __ldg(src + 3);
                            }
                        }
                        a[rr][q] = v;
                    }
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const int row = w * 32 + r0 + rr;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const uint32_t b0 = __ballot_sync(0xffffffffu, a[rr][q].x > 0.f);
                        const uint32_t b1 = __ballot_sync(0xffffffffu, a[rr][q].y > 0.f);
                        const uint32_t b2 = __ballot_sync(0xffffffffu, a[rr][q].z > 0.f);
                        const uint32_t b3 = __ballot_sync(0xffffffffu, a[rr][q].w > 0.f);
                        if (lane == 0) {
                            uint4* dst = reinterpret_cast<uint4*>(mask_s + row * (BN_MAX / 32) + 4 * q);
                            *dst = make_uint4(b0, b1, b2, b3);
                        }
                    }
                }
            }
            mbar_arrive(smem_u32(maskbar));
        }
        if (PASSES == 3 && !p.b_manual) {
            // second splitter group: the larger share of every B tile (the operand split is what paces the mainloop:
            // one group of four warps needs ~1.1 us per k-slab for A + B, the tensor core 0.85 us)
            const int ty = threadIdx.x - 192;               // 0..127
            const int nB4 = NT * BK / 4, nBx = (nB4 / 4) & ~511;
            for (int it = y_from; it < nslabs; ++it) {
                const int s = it % STAGES;
                mbar_wait(smem_u32(&full[s]), (it / STAGES) & 1);
                split_range(reinterpret_cast<const float4*>(b_raw(s)), reinterpret_cast<float4*>(b_lo(s)), nBx, nB4, ty);
                fence_proxy_async();
                mbar_arrive(smem_u32(&splitb[s]));
            }
        }
    } else if (warp >= 2) {
        const int t = threadIdx.x - 64;                     // 0..127
        {   // bias of this tile -> shared memory (read again only in the epilogue, behind a barrier of the eight warps;
            // staged here and not before the CTA-wide sync so that the first TMA does not wait for this global load)
            const float* bg = p.bias ? p.bias + (int64_t)g * p.bias_gs + n0 : nullptr;
            for (int i = t; i < BN_MAX; i += 128) bias_s[i] = (bg != nullptr && i < n_lim) ? __ldg(bg + i) : 0.f;
        }
        if (split_runs) {
            // -------------------------------------------------------------- operand splitter (+ rank-1 generator)
            const int nB4 = NT * BK / 4;
            const int nBx = (nB4 / 4) & ~511;               // this group's share of B: it also has all of A to do
            // float4 number t + 128 j of the swizzled A tile sits in row t/8 + 16 j at chunk position t%8, i.e. it holds
            // k = 4c .. 4c+3 of that row with c = (t%8) ^ ((t/8) & 7) - the same c for all eight j.  So the generator
            // needs eight row factors (fixed for the whole CTA) and one float4 of column factors per k-slab.
            // MN-major A ([k][32 m] blocks of 256 float4): float4 t + 128 j is block j/2, k-row t/8 + 16 (j%2), m-chunk c.
            // There the ROW factors (per m) are four float4s fixed for the CTA and the column factors are two scalars
            // (k rows t/8 and t/8 + 16) per slab.
            float grow[8];
            float4 growv[4];
            const int gc = p.a_mn ? 2 * (((t & 7) >> 1) ^ ((t >> 3) & 3)) + (t & 1)       // 32-byte atoms, rows mod 4
                                  : (t & 7) ^ ((t >> 3) & 7);
            const float* gcol = nullptr;
            if (gen && !p.a_mn) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int m = tile_m * BM + (t >> 3) + 16 * j;
                    grow[j] = m < p.M ? __ldg(p.gen_row + (int64_t)g * p.gen_row_gs + m) : 0.f;
                }
                gcol = p.gen_col + (int64_t)g * p.gen_col_gs + 4 * gc;
            }
            if (gen && p.a_mn) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int m = tile_m * BM + 32 * b + 4 * gc;
                    const float* gr = p.gen_row + (int64_t)g * p.gen_row_gs + m;
                    growv[b] = make_float4(m < p.M ? __ldg(gr) : 0.f, m + 1 < p.M ? __ldg(gr + 1) : 0.f,
                                           m + 2 < p.M ? __ldg(gr + 2) : 0.f, m + 3 < p.M ? __ldg(gr + 3) : 0.f);
                }
                gcol = p.gen_col + (int64_t)g * p.gen_col_gs;
            }
            for (int it = 0; it < nslabs; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(smem_u32(r3 ? a_full : &full[s]), r3 ? (uint32_t)(it & 1) : ph);
                if (t == 0 && it == 0) TC_STAMP(3);
                if (t == 0 && p.trace_mode == 0 && it < 8) TC_STAMP(8 + it);
                const bool st4 = (t == 0 && p.trace_mode == 4 && it == 1);
                if (st4) TC_STAMP(8);
                float4* __restrict__ ar = reinterpret_cast<float4*>(a_raw(s));
                float4* __restrict__ al = reinterpret_cast<float4*>(a_lo(s));
                float4* __restrict__ br = reinterpret_cast<float4*>(b_raw(s));
                float4* __restrict__ bl = reinterpret_cast<float4*>(b_lo(s));
                if (p.a_tmem) {
                    // A-from-TMEM: thread -> one ROW m of the tile (warp w owns TMEM lanes 32 (w % 4) .. +31); it reads
                    // the row's 32 k values out of the swizzled tile, optionally turns them into the rank-1 gradient,
                    // and stores them (and their lo parts) into the stage's TMEM columns
                    const int row = 32 * (warp & 3) + lane;
                    const float* af = reinterpret_cast<const float*>(a_raw(s));
                    float x[32];
                    if (p.a_mn) {       // [k][32 m] blocks of 4 KB, 32-byte chunk c of row k at position c ^ (k & 3)
                        const float* blk = af + (row >> 5) * 1024;
                        const int c32 = (row & 31) >> 3, e = row & 7;
#pragma unroll
                        for (int k = 0; k < 32; ++k) x[k] = blk[k * 32 + ((c32 ^ (k & 3)) << 3) + e];
                    } else {            // [128 m][32 k], 16-byte chunk c of row m at position c ^ (m & 7)
                        const float4* r4 = reinterpret_cast<const float4*>(af + row * 32);
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 q = r4[c ^ (row & 7)];
                            x[4 * c] = q.x; x[4 * c + 1] = q.y; x[4 * c + 2] = q.z; x[4 * c + 3] = q.w;
                        }
                    }
                    if (gen) {
                        const int m = tile_m * BM + row;
                        const float gr = m < p.M ? __ldg(p.gen_row + (int64_t)g * p.gen_row_gs + m) : 0.f;
                        const float* gcl = p.gen_col + (int64_t)g * p.gen_col_gs + (int64_t)(slab0 + it) * BK;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const int k = (slab0 + it) * BK + 4 * c;
                            const float4 cv = k < p.K ? __ldg(reinterpret_cast<const float4*>(gcl) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                            x[4 * c] = x[4 * c] > 0.f ? gr * cv.x : 0.f;
                            x[4 * c + 1] = x[4 * c + 1] > 0.f ? gr * cv.y : 0.f;
                            x[4 * c + 2] = x[4 * c + 2] > 0.f ? gr * cv.z : 0.f;
                            x[4 * c + 3] = x[4 * c + 3] > 0.f ? gr * cv.w : 0.f;
                        }
                    }
                    const uint32_t ta = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + ATM_COL + s * ATM_STRIDE;
                    tmem_st32(ta, x);
                    if (PASSES == 3) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) x[k] = x[k] - __uint_as_float(__float_as_uint(x[k]) & 0xFFFFE000u);
                        tmem_st32(ta + 32, x);
                    }
                    tmem_wait_st();
                    if (r3) {       // the A buffer may be refilled; B of this slab is a separate barrier
                        mbar_arrive(smem_u32(a_free));
                        mbar_wait(smem_u32(&full[s]), ph);
                    }
                }
                // A: 1024 float4 -> 8 per thread, all loads issued before the first store
                if (!p.a_tmem) {
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = ar[t + 128 * j];
                    if (gen && p.a_mn) {
                        const int k = (slab0 + it) * BK + (t >> 3);
                        const float c0 = k < p.K ? __ldg(gcol + k) : 0.f, c1 = k + 16 < p.K ? __ldg(gcol + k + 16) : 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float cf = (j & 1) ? c1 : c0;
                            const float4 rv = growv[j >> 1];
                            v[j].x = v[j].x > 0.f ? rv.x * cf : 0.f;
                            v[j].y = v[j].y > 0.f ? rv.y * cf : 0.f;
                            v[j].z = v[j].z > 0.f ? rv.z * cf : 0.f;
                            v[j].w = v[j].w > 0.f ? rv.w * cf : 0.f;
                            ar[t + 128 * j] = v[j];
                        }
                    } else if (gen) {
                        const int k = (slab0 + it) * BK + 4 * gc;
                        const float4 cv = k < p.K ? __ldg(reinterpret_cast<const float4*>(gcol + (int64_t)(slab0 + it) * BK))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[j].x = v[j].x > 0.f ? grow[j] * cv.x : 0.f;
                            v[j].y = v[j].y > 0.f ? grow[j] * cv.y : 0.f;
                            v[j].z = v[j].z > 0.f ? grow[j] * cv.z : 0.f;
                            v[j].w = v[j].w > 0.f ? grow[j] * cv.w : 0.f;
                            ar[t + 128 * j] = v[j];
                        }
                    }
                    if (PASSES == 3) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) al[t + 128 * j] = lo_tf32(v[j]);
                    }
                }
                if (st4) TC_STAMP(9);
                if (p.b_manual) {
                    // B tile by hand (a single k-slab): 16-byte chunk c (k = 4c .. 4c+3) of row n of the K-major
                    // SWIZZLE_128B tile lives at n * 128 bytes + 16 * (c ^ (n & 7)); k >= K and n >= N are zeros
                    const float* Bg = p.Bm + (int64_t)g * p.bm_gs + (int64_t)n0 * p.ldbm;
                    for (int n = t; n < NT; n += 128) {
                        const float* row = Bg + (int64_t)n * p.ldbm;
                        const bool n_ok = n0 + n < p.N;
#pragma unroll
                        for (int c = 0; c < BK / 4; ++c) {
                            float4 v;
                            v.x = (n_ok && 4 * c + 0 < p.K) ? __ldg(row + 4 * c + 0) : 0.f;
                            v.y = (n_ok && 4 * c + 1 < p.K) ? __ldg(row + 4 * c + 1) : 0.f;
                            v.z = (n_ok && 4 * c + 2 < p.K) ? __ldg(row + 4 * c + 2) : 0.f;
                            v.w = (n_ok && 4 * c + 3 < p.K) ? __ldg(row + 4 * c + 3) : 0.f;
                            br[n * 8 + (c ^ (n & 7))] = v;
                            if (PASSES == 3) bl[n * 8 + (c ^ (n & 7))] = lo_tf32(v);
                        }
                    }
                }
                // the mask warps take [nBx, nB4) - except for the first y_from slabs of a launch whose mask they are building
                if (PASSES == 3 && !p.b_manual) split_range(br, bl, 0, it < y_from ? nB4 : nBx, t);
                if (st4) TC_STAMP(10);
                if (p.a_tmem) tc_fence_before();            // TMEM stores ordered before the MMA issuer's reads
                fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core
                if (st4) TC_STAMP(11);
                mbar_arrive(smem_u32(&splitb[s]));
                if (st4) TC_STAMP(12);
            }
        }
    }
    if (warp < 2 && p.late_trigger) orlk::pdl_trigger();
    if (warp >= 2) {
        // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
        // Eight warps: warp w may only touch TMEM lanes 32*(w%4) .. +31, so quadrant q is shared by warps 2+q' and 6+q''
        // which take alternate 32-column chunks.  A lane owns one accumulator ROW, so:
        //  * C (row-major) is parked in the now idle operand ring as 32x32 SWIZZLE_128B tiles and leaves through TMA
        //    stores (direct stores put 32 different lines in every instruction: measured 6 us for a 128x256 tile);
        //  * CT (transposed) is written directly, a warp instruction covering 32 consecutive floats of one CT row.
        const int t = threadIdx.x - 64;
        if (p.epi == ORLK_EPI_RELU_MASK) mbar_wait(smem_u32(maskbar), 0);
        mbar_wait(smem_u32(accum), 0);
        tc_fence_after();
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps: bias_s (and mask_s) complete
        if (p.late_trigger) orlk::pdl_trigger();        // mainloop done: only the epilogue is left of this CTA
        if (t == 0) TC_STAMP(6);
        const int q = warp & 3;
        const int half = warp >= 6 ? 1 : 0;
        const int row = q * 32 + lane;
        const int m = tile_m * BM + row;
        const bool row_ok = m < p.M;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        float* C = p.C ? p.C + (int64_t)g * p.c_gs + (int64_t)split * p.c_split_stride + (int64_t)m * p.ldc + n0 : nullptr;
        float* CT = p.CT ? p.CT + (int64_t)g * p.ct_gs + (int64_t)n0 * p.ldct + m : nullptr;
        const bool vec_ok = (n_lim % 4 == 0) && (p.ldc % 4 == 0) && aligned16(p.C) && (p.c_gs % 4 == 0) &&
                            (p.c_split_stride % 4 == 0);
        const bool c_tma = p.c_tma != 0;
        bool issued = false;
        for (int c0 = 32 * half; c0 < n_lim; c0 += 64) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            tmem_wait_ld();
            float x[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * j4);
                x[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + b4.x;
                x[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b4.y;
                x[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b4.z;
                x[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b4.w;
            }
            if (p.epi == ORLK_EPI_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
            } else if (p.epi == ORLK_EPI_SWISH) {       // dynamics_module.py:12-17 (inference: the pre-activation is not kept)
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = x[j] / (1.f + expf(-x[j]));
            } else if (p.epi == ORLK_EPI_RELU_MASK) {
                // columns c0 .. c0+31 live in 128-column group c0/128 at bits 8*((c0/32)%4) .. +7
                const uint4 mb = *reinterpret_cast<const uint4*>(mask_s + row * (BN_MAX / 32) + 4 * (c0 >> 7));
                const int sh = 8 * ((c0 >> 5) & 3);
                const uint32_t m0 = mb.x >> sh, m1 = mb.y >> sh, m2 = mb.z >> sh, m3 = mb.w >> sh;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    x[4 * jj + 0] = ((m0 >> jj) & 1u) ? x[4 * jj + 0] : 0.f;
                    x[4 * jj + 1] = ((m1 >> jj) & 1u) ? x[4 * jj + 1] : 0.f;
                    x[4 * jj + 2] = ((m2 >> jj) & 1u) ? x[4 * jj + 2] : 0.f;
                    x[4 * jj + 3] = ((m3 >> jj) & 1u) ? x[4 * jj + 3] : 0.f;
                }
            }
            if (c_tma) {
                // tile (q, c0/32): 32 rows x 128 bytes, 16-byte chunk j of row r at position j ^ (r & 7)
                uint8_t* tile = smem + (size_t)(q * ((NT + 31) >> 5) + (c0 >> 5)) * 4096;
                float4* trow = reinterpret_cast<float4*>(tile + lane * 128);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    trow[j4 ^ (lane & 7)] = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
                fence_proxy_async();
                __syncwarp();
                if (elect_one()) {
                    tma_store_4d(&tmC, smem_u32(tile), n0 + c0, tile_m * BM + q * 32, g, split);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                issued = true;
            } else if (row_ok && C != nullptr) {
                if (vec_ok) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        if (c0 + 4 * j4 + 3 < n_lim)
                            *reinterpret_cast<float4*>(C + c0 + 4 * j4) =
                                make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < n_lim) C[c0 + j] = x[j];
                }
            }
            if (row_ok && CT != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < n_lim) CT[(int64_t)(c0 + j) * p.ldct] = x[j];
            }
        }
        if (want_rowsum && half == 0) {
            const uint32_t r = tmem_ld1(taddr + ROWSUM_COL);
            tmem_wait_ld();
            if (row_ok) p.rowsum[(int64_t)g * p.rowsum_gs + (int64_t)split * p.rowsum_split_stride + m] = __uint_as_float(r);
        }
        if (issued) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the tiles must outlive the stores
    }
    if (threadIdx.x == 64) TC_STAMP(7);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}


}  // namespace

extern "C" int orlk_sizeof_tc_gemm(void) { return (int)sizeof(OrlkTcGemm); }

extern "C" int orlk_tc_set_trace(void* dev_buf) {
    orlk::set_trace_buffer((unsigned long long*)dev_buf);
    return 0;
}

constexpr int MAX_DYN_SMEM = 227 * 1024;

// ring geometry for (n-tile, precision): as many stages as fit next to the fixed region, at most MAX_STAGES
static void tc_ring(int NT, int passes, int* stages, int* stage_bytes) {
    *stage_bytes = (A_BYTES + NT * BK * 4) * (passes == 3 ? 2 : 1);
    int n = (MAX_DYN_SMEM - 1024 - FIXED_SMEM) / *stage_bytes;
    *stages = n > MAX_STAGES ? MAX_STAGES : (n < 1 ? 1 : n);
}

// Opt in to the full 227 KB of dynamic shared memory once, outside any stream capture.
extern "C" int orlk_tc_init(void) {
    int rc = check(cudaFuncSetAttribute(k_tc_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM), "smem attr <1>");
    if (rc) return rc;
    return check(cudaFuncSetAttribute(k_tc_gemm<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM), "smem attr <3>");
}

extern "C" int orlk_tc_gemm(const OrlkTcGemm* q, void* stream) {
    ORLK_REQUIRE(q != nullptr, "params");
    ORLK_REQUIRE(q->M > 0 && q->K > 0 && q->G > 0, "sizes");
    ORLK_REQUIRE(q->N >= 1 && q->N <= BN_MAX, "N must be in [1,256]");
    ORLK_REQUIRE(q->passes == 1 || q->passes == 3, "passes must be 1 or 3");
    ORLK_REQUIRE(q->lda % 4 == 0 && q->a_gs % 4 == 0 && aligned16(q->A), "A must be 16-byte aligned with strides that are multiples of 4 floats");
    // B rows that TMA cannot address (the K = obs+act wide first-layer weights, pitch 23 floats) are staged by hand
    const bool b_manual = q->ldb % 4 != 0 || q->b_gs % 4 != 0 || !aligned16(q->B);
    ORLK_REQUIRE(!b_manual || q->K <= BK, "B must be 16-byte aligned with strides that are multiples of 4 floats unless K <= 32");
    ORLK_REQUIRE(q->epi == ORLK_EPI_NONE || q->epi == ORLK_EPI_RELU || q->epi == ORLK_EPI_RELU_MASK || q->epi == ORLK_EPI_SWISH,
                 "epilogue");
    ORLK_REQUIRE(q->epi != ORLK_EPI_RELU_MASK || q->aux != nullptr, "mask epilogue needs aux");
    const int total_slabs = (q->K + BK - 1) / BK;
    int splits = q->k_splits < 1 ? 1 : q->k_splits;
    int per = (total_slabs + splits - 1) / splits;
    splits = (total_slabs + per - 1) / per;
    ORLK_REQUIRE(splits == 1 || (q->epi == ORLK_EPI_NONE && q->bias == nullptr), "split-K needs a linear epilogue");
    ORLK_REQUIRE(splits == (q->k_splits < 1 ? 1 : q->k_splits), "k_splits must divide the slab count evenly enough (use orlk_tc_effective_splits)");

    CUtensorMap tmA, tmB;
    const bool a_shared = q->a_gs == 0 && q->G > 1, b_shared = q->b_gs == 0 && q->G > 1;
    int rc = q->a_mn ? make_map_mn(&tmA, q->A, q->lda, q->a_gs, q->M, q->K, a_shared ? 1 : q->G)
                     : make_map(&tmA, q->A, q->lda, q->a_gs, q->M, q->K, a_shared ? 1 : q->G, BM);
    if (rc) return rc;
    // the MMA is always NT (a multiple of 16) columns wide; columns past N are zero operand rows and are not stored
    const int NT = (q->n_tile > 0) ? q->n_tile : (q->N + 15) / 16 * 16;
    ORLK_REQUIRE(NT >= 16 && NT <= BN_MAX && NT % 16 == 0, "n_tile must be a multiple of 16 in [16,256]");
    ORLK_REQUIRE(!q->b_mn || (NT % 32 == 0 && !b_manual), "an MN-major B needs 32-column n-tiles and 16-byte aligned rows");
    if (b_manual) memset(&tmB, 0, sizeof(tmB));
    else {
        rc = q->b_mn ? make_map_mn(&tmB, q->B, q->ldb, q->b_gs, q->N, q->K, b_shared ? 1 : q->G)
                     : make_map(&tmB, q->B, q->ldb, q->b_gs, q->N, q->K, b_shared ? 1 : q->G, NT);
        if (rc) return rc;
    }

    TcParams p;
    p.C = q->C; p.ldc = q->ldc; p.c_gs = q->c_gs; p.c_split_stride = q->c_split_stride;
    p.CT = q->CT; p.ldct = q->ldct; p.ct_gs = q->ct_gs;
    p.bias = q->bias; p.bias_gs = q->bias_gs;
    p.aux = q->aux; p.ldaux = q->ldaux; p.aux_gs = q->aux_gs;
    p.rowsum = q->rowsum; p.rowsum_gs = q->rowsum_gs; p.rowsum_split_stride = q->rowsum_split_stride;
    p.M = q->M; p.N = q->N; p.K = q->K; p.G = q->G; p.epi = q->epi;
    p.k_splits = splits; p.slabs_per_split = per; p.tiles_m = (q->M + BM - 1) / BM;
    p.NT = NT; p.tiles_n = (q->N + NT - 1) / NT;
    ORLK_REQUIRE((q->gen_row == nullptr) == (q->gen_col == nullptr), "gen_row and gen_col go together");
    ORLK_REQUIRE(q->gen_row == nullptr || (q->K % 4 == 0 && aligned16(q->gen_col) && q->gen_col_gs % 4 == 0),
                 "the operand generator needs K % 4 == 0 and 16-byte aligned column factors");
    p.a_mn = q->a_mn ? 1 : 0; p.b_mn = q->b_mn ? 1 : 0;
    {
        static int a_tmem = -1;
        if (a_tmem < 0) { const char* e = getenv("ORLK_TC_A_TMEM"); a_tmem = e ? atoi(e) : 1; }
        p.a_tmem = a_tmem;
        static int late = -1;
        if (late < 0) { const char* e = getenv("ORLK_TC_LATE_TRIGGER"); late = e ? atoi(e) : 1; }
        p.late_trigger = late;
    }
    p.Bm = q->B; p.ldbm = q->ldb; p.bm_gs = q->b_gs; p.b_manual = b_manual ? 1 : 0; p.a_shared = a_shared ? 1 : 0; p.b_shared = b_shared ? 1 : 0;
    p.gen_row = q->gen_row; p.gen_row_gs = q->gen_row_gs; p.gen_col = q->gen_col; p.gen_col_gs = q->gen_col_gs;
    p.trace = orlk::trace_buffer();
    { const char* e = getenv("ORLK_TC_TRACE_MODE"); p.trace_mode = e ? atoi(e) : 0; }

    tc_ring(NT, q->passes, &p.stages, &p.stage_bytes);
    if (p.a_tmem) {     // the A stages live in the TMEM columns behind the accumulator: 512 - 288 columns
        const int max_st = (TMEM_COLS - ATM_COL) / (q->passes == 3 ? 64 : 32);
        if (p.stages > max_st) p.stages = max_st;
    }
    size_t ring_bytes = (size_t)p.stages * p.stage_bytes;
    {
        static int r3 = -1;
        if (r3 < 0) { const char* e = getenv("ORLK_TC_R3"); r3 = e ? atoi(e) : 1; }
        p.r3 = (r3 && q->passes == 3 && p.a_tmem && !b_manual) ? 1 : 0;
    }
    if (p.r3) {         // one A buffer + three [B | B lo] stages
        p.stages = 3;
        p.stage_bytes = 2 * NT * BK * 4;
        ring_bytes = A_BYTES + (size_t)p.stages * p.stage_bytes;
    }
    CUtensorMap tmC;
    memset(&tmC, 0, sizeof(tmC));
    p.c_tma = 0;
    if (q->C != nullptr && NT % 32 == 0 && q->N % 4 == 0 && q->ldc % 4 == 0 && aligned16(q->C) && q->c_gs % 4 == 0 && q->c_split_stride % 4 == 0 &&
        (int64_t)NT * 512 <= (int64_t)ring_bytes) {
        rc = make_map_c(&tmC, q->C, q->ldc, q->c_gs, q->c_split_stride, q->M, q->N, q->G, splits);
        if (rc) return rc;
        p.c_tma = 1;
    }
    const size_t smem = 1024 + FIXED_SMEM + ring_bytes;
    const int grid = q->G * p.tiles_m * p.tiles_n * splits;
    cudaStream_t s = (cudaStream_t)stream;
    static int hp = -1;     // experiment (ORLK_TC_PRIORITY=1): the GEMMs' CTAs are dispatched before optimiser / SIMT blocks
    if (hp < 0) { const char* e = getenv("ORLK_TC_PRIORITY"); hp = e ? atoi(e) : 0; }
    if (hp) {
        if (q->passes == 3) orlk::launch_high_priority(k_tc_gemm<3>, grid, NUM_THREADS, smem, s, tmA, tmB, tmC, p);
        else orlk::launch_high_priority(k_tc_gemm<1>, grid, NUM_THREADS, smem, s, tmA, tmB, tmC, p);
    } else if (q->passes == 3) orlk::launch(k_tc_gemm<3>, grid, NUM_THREADS, smem, s, tmA, tmB, tmC, p);
    else orlk::launch(k_tc_gemm<1>, grid, NUM_THREADS, smem, s, tmA, tmB, tmC, p);
    return check_launch("k_tc_gemm");
}

// number of k-splits orlk_tc_gemm will accept for (K, want): chunks are whole 32-wide slabs
extern "C" int orlk_tc_effective_splits(int K, int want) {
    const int total_slabs = (K + BK - 1) / BK;
    if (want < 1) want = 1;
    const int per = (total_slabs + want - 1) / want;
    return (total_slabs + per - 1) / per;
}
