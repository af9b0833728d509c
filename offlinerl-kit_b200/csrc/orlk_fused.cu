// Whole critic passes in ONE launch on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 = fp32-grade).  sm_100a only.
//
// k_critic_fwd: q[g][m] = head( relu( ... relu( relu(X W0^T + b0) W1^T + b1) ... ) )  for every member g (twin critics)
// of a Linear+ReLU stack with equal hidden widths N <= 256 and a scalar head (reference: nets/mlp.py:22-28 forward inside
// modules/critic_module.py:25-33, called on the 7936-row CQL critic batch, policy/model_free/cql.py:133-160).
//
// One CTA owns a 128-row strip of ONE member for the whole pass: the activations never leave the SM between layers.
//   * accumulators: two 128 x N fp32 tiles in tensor memory (ping-pong over the layers, 2 x 256 columns);
//   * layer l's epilogue (8 warps) turns accumulator columns [32c, 32c+32) into relu(acc + b) and writes them - and their
//     lo = x - trunc_tf32(x) part - as the K-major SWIZZLE_128B A tile of k-slab c of layer l+1 (two-stage ring); the
//     same tile leaves for H[l] in global memory by TMA store (the backward pass needs it), so the next layer's MMAs
//     start on slab 0 while the epilogue is still draining slab 1..7;
//   * weights: [W | W lo] k-slabs through a two-stage TMA ring; the lo copies are kept by orlk_split_lo (one tiny launch
//     per step, off the critical path) instead of being recomputed by all 62 strips of a member;
//   * first layer (K0 = obs+act <= 32 columns, rows not TMA-addressable): X by TMA with zero fill, W0 staged by hand;
//   * scalar head: a dot product with the last epilogue's registers, one value per row.
// Roles: warp 0 TMA producer, warp 1 MMA issuer (one elected lane each), warps 2-5 / 6-9 epilogue of the even / odd
// 32-column chunks (warp w may touch TMEM lanes 32 (w % 4) .. +31).
#include "orlk_tcgen.cuh"
using namespace orlk;
using namespace orlk::tcg;

namespace {

constexpr int BM = 128, BK = 32, NMAX = 256, MAXL = ORLK_FUSED_MAX_LAYERS;
constexpr int A_TILE = BM * BK * 4;            // 16 KB
constexpr int A_STAGE = 2 * A_TILE;            // [hi | lo]
constexpr int B_TILE = NMAX * BK * 4;          // 32 KB
constexpr int B_STAGE = 2 * B_TILE;            // [hi | lo]
constexpr int RING = 2 * A_STAGE + 2 * B_STAGE;     // 192 KB
constexpr int FIXED = 8192;
constexpr int NUM_THREADS = 320;
constexpr int TMEM_COLS = 512;

enum { BAR_X = 0, BAR_AFULL = 1, BAR_AEMPTY = 4, BAR_ACC = 7, BAR_BFULL = 9, BAR_BEMPTY = 13, BAR_COUNT = 17 };

struct FwdMaps {
    CUtensorMap x;                // X [M][K0]: box 32 (k, zero filled past K0) x 128 rows
    CUtensorMap w0, w0lo;         // zero-padded first-layer weights [G][N][32] and their lo words: box 32 x N
    CUtensorMap w[MAXL - 1];      // W_l [G][N][N], l >= 1: box 32 (k) x N
    CUtensorMap wlo[MAXL - 1];
    CUtensorMap h[MAXL];          // H_l [G][M][N]: box 32 x 32 (store)
};

struct FwdParams {
    int64_t gs;                           // member stride of the bias / head tensors
    const float* bias[MAXL];
    const float* head_w; const float* head_b;
    float* out; int64_t out_gs;
    uint32_t* bits;                       // [L][G][8][M] ReLU-decision bits of every activation (bit j of word [c][m]: column 32c+j > 0), or NULL
    int M, N, G, L, tiles_m, store_h;
    int no_store;                         // experiment (ORLK_FUSED_NO_STORE=1): skip the H stores, results are then incomplete
    unsigned long long* trace;            // profiling aid (orlk_tc_set_trace): 128 clock stamps per CTA, NULL in normal operation
};

__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define FZ_GSTAMP(slot)                                                                                \
    do {                                                                                               \
        if (p.trace != nullptr) p.trace[(int64_t)blockIdx.x * 128 + (slot)] = gtimer_ns();             \
    } while (0)
#define FZ_STAMP(slot)                                                                                          \
    do {                                                                                                        \
        if (p.trace != nullptr) p.trace[(int64_t)blockIdx.x * 128 + (slot)] = (unsigned long long)clock64();    \
    } while (0)

__device__ __forceinline__ float lo_of(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Up to two independent passes ("jobs") share one launch: CTAs [0, ctas0) run job 0 (the online critics on the 7936-row
// batch), the rest job 1 (the target critics on the next-state rows, no activations stored) - side by side instead of
// one pass starving the other of SMs.
//
// PAIR = true: CTA pairs (thread-block clusters of two, tcgen05.mma.cta_group::2, M = 256 over two strips).  Each CTA keeps
// its own 128 rows of A and its own accumulators, but only HALF of every weight slab (N/2 rows: 32 KB instead of 64 KB
// per slab and SM; three A and three weight stages in the same shared memory); the leader CTA issues the MMAs for both, its barriers
// collect the A tiles of both CTAs (remote arrivals) and the bytes of both CTAs' TMA loads, and its commits are
// multicast to the stage / accumulator barriers of both.  What it buys: the L2 -> SM path of a TPC, which bounds the
// second SM of a pair in the single-CTA kernel, carries half the bytes.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of one CTA's part of a pair's operand: the bytes are counted on the LEADER's barrier (a shared::cluster address)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {        // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_critic_fwd_t(const __grid_constant__ FwdMaps maps0, const __grid_constant__ FwdMaps maps1,
             const __grid_constant__ FwdParams p0, const __grid_constant__ FwdParams p1, const int ctas0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // SWIZZLE_128B tiles: 1024-byte aligned
    // ring split of the 192 KB: single CTA: 2 A stages (64 KB) + 2 weight stages of 64 KB; pair: 3 A stages (96 KB: the A
    // round trip  MMA done -> multicast commit -> epilogue writes -> remote arrive -> issue  is longer than one slab's MMAs,
    // two stages left the mainloop waiting for A every other slab) + 3 weight stages of 32 KB
    constexpr int NSA = PAIR ? 3 : 2;
    uint8_t* b_base = base + NSA * A_STAGE;
    uint8_t* fixed = base + RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(fixed);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fixed + 192);
    float* bias_s = reinterpret_cast<float*>(fixed + 256);       // [MAXL][NMAX]
    float* headw_s = bias_s + MAXL * NMAX;                        // [NMAX]
    float* qpart_s = headw_s + NMAX;                              // [BM]
    auto a_hi = [&](int s) { return base + s * A_STAGE; };
    auto a_lo = [&](int s) { return base + s * A_STAGE + A_TILE; };
    constexpr int NSB = PAIR ? 3 : 2;                      // weight ring stages
    constexpr int BST = PAIR ? B_STAGE / 2 : B_STAGE;      // bytes per stage: [hi | lo] of this CTA's rows of the slab
    auto b_hi = [&](int s) { return b_base + s * BST; };
    auto b_lo = [&](int s) { return b_base + s * BST + BST / 2; };
    auto bar = [&](int i) { return smem_u32(&bars[i]); };

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const bool job1 = (int)blockIdx.x >= ctas0;
    const FwdMaps& maps = job1 ? maps1 : maps0;
    const FwdParams& p = job1 ? p1 : p0;
    const int cta = job1 ? (int)blockIdx.x - ctas0 : (int)blockIdx.x;
    const int g = cta / p.tiles_m;
    const int tile_m = cta - g * p.tiles_m;
    const int N = p.N, L = p.L;
    const int KS = N / BK;                      // k-slabs of a hidden layer = 32-column chunks of an accumulator
    const uint32_t rank = PAIR ? cluster_rank() : 0u;
    const int b_rows = PAIR ? N / 2 : N;        // weight rows this CTA keeps of every slab
    const bool store_h = p.store_h != 0 && p.no_store == 0;

    // ---------------------------------------------------------------- prologue (touches no global data)
    if (threadIdx.x == 0) { FZ_STAMP(0); FZ_GSTAMP(4); }
    if (threadIdx.x == 32) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w0lo) : "memory");
        for (int l = 1; l < L; ++l) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w[l - 1]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wlo[l - 1]) : "memory");
        }
        if (store_h)
            for (int l = 0; l < L; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.h[l]) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bar(BAR_X), 1);
        for (int s = 0; s < NSA; ++s) {
            mbar_init(bar(BAR_AFULL + s), PAIR ? 8 : 4);    // one arrival per epilogue warp of the chunk's group (of both CTAs)
            // free again = the MMAs that read the stage have completed (1, a commit) AND the four warps that filled it have
            // seen their TMA stores of its tiles read the shared memory (4; with three stages the next writer of a
            // stage is a warp of the OTHER group, whose own bulk-group wait says nothing about those stores)
            mbar_init(bar(BAR_AEMPTY + s), 5);
        }
        for (int s = 0; s < 2; ++s) mbar_init(bar(BAR_ACC + s), 1);
        for (int s = 0; s < NSB; ++s) {
            mbar_init(bar(BAR_BFULL + s), 1);
            mbar_init(bar(BAR_BEMPTY + s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    orlk::pdl_wait();                           // X and the weights come from earlier kernels of the step
    if (threadIdx.x == 0) { FZ_STAMP(1); FZ_GSTAMP(5); }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();               // the pair's barriers are initialised before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------------------ TMA producer
            mbar_expect_tx(bar(BAR_X), A_TILE);
            tma_load_3d(smem_u32(a_hi(0)), &maps.x, bar(BAR_X), 0, tile_m * BM, 0);
            const uint32_t tx = 2u * (uint32_t)N * BK * 4;      // a whole slab: hi + lo (of both CTAs' halves)
            const int n_fill = 1 + (L - 1) * KS;                // fill 0: the padded first-layer weights
            for (int bi = 0; bi < n_fill; ++bi) {
                const int s = bi % NSB, nf = bi / NSB;
                const int l = bi == 0 ? 0 : 1 + (bi - 1) / KS, j = bi == 0 ? 0 : (bi - 1) % KS;
                const CUtensorMap* mh = l == 0 ? &maps.w0 : &maps.w[l - 1];
                const CUtensorMap* ml = l == 0 ? &maps.w0lo : &maps.wlo[l - 1];
                mbar_wait(bar(BAR_BEMPTY + s), (nf & 1) ^ 1);
                if (bi < 32) FZ_STAMP(80 + bi);
                if (PAIR) {
                    const uint32_t lb = map_to_rank(bar(BAR_BFULL + s), 0);
                    if (rank == 0) mbar_expect_tx(bar(BAR_BFULL + s), tx);
                    tma_load_3d_pair(smem_u32(b_hi(s)), mh, lb, j * BK, (int)rank * b_rows, g);
                    tma_load_3d_pair(smem_u32(b_lo(s)), ml, lb, j * BK, (int)rank * b_rows, g);
                } else {
                    mbar_expect_tx(bar(BAR_BFULL + s), tx);
                    tma_load_3d(smem_u32(b_hi(s)), mh, bar(BAR_BFULL + s), j * BK, 0, g);
                    tma_load_3d(smem_u32(b_lo(s)), ml, bar(BAR_BFULL + s), j * BK, 0, g);
                }
            }
        }
        __syncwarp();
        orlk::pdl_trigger();
    } else if (warp == 1) {
        if ((!PAIR || rank == 0) && elect_one()) {
            // ------------------------------------------------------------ MMA issuer (the leader CTA's, for a pair)
            const uint32_t idesc = instr_desc_tf32(PAIR ? 2 * BM : BM, N);
            int bi = 0;                         // slabs issued so far = index of the A fill and of the B fill they consume
            for (int l = 0; l < L; ++l) {
                const uint32_t acc = tmem_base + (uint32_t)(NMAX * (l & 1));
                const int nsl = l == 0 ? 1 : KS;
                for (int j = 0; j < nsl; ++j, ++bi) {
                    const int sa = bi % NSA;
                    mbar_wait(bar(BAR_AFULL + sa), (bi / NSA) & 1);
                    if (bi < 32) FZ_STAMP(16 + bi);
                    const int sb = bi % NSB;
                    mbar_wait(bar(BAR_BFULL + sb), (bi / NSB) & 1);
                    if (bi < 32) FZ_STAMP(48 + bi);
                    tc_fence_after();
                    const uint64_t ad = smem_desc_sw128(smem_u32(a_hi(sa))), adl = smem_desc_sw128(smem_u32(a_lo(sa)));
                    const uint64_t bd = smem_desc_sw128(smem_u32(b_hi(sb))), bdl = smem_desc_sw128(smem_u32(b_lo(sb)));
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {          // UMMA_K = 8 for tf32: 32 bytes along a K-major row
                        const uint64_t ko = (uint64_t)(2 * k);
                        if (PAIR) {
                            umma_tf32_pair(acc, ad + ko, bd + ko, idesc, (j > 0 || k > 0) ? 1u : 0u);
                            umma_tf32_pair(acc, adl + ko, bd + ko, idesc, 1u);
                            umma_tf32_pair(acc, ad + ko, bdl + ko, idesc, 1u);
                        } else {
                            umma_tf32(acc, ad + ko, bd + ko, idesc, (j > 0 || k > 0) ? 1u : 0u);
                            umma_tf32(acc, adl + ko, bd + ko, idesc, 1u);
                            umma_tf32(acc, ad + ko, bdl + ko, idesc, 1u);
                        }
                    }
                    if (PAIR) {                                 // both stages are free (in both CTAs) once these MMAs have completed
                        umma_commit_pair(bar(BAR_AEMPTY + sa));
                        umma_commit_pair(bar(BAR_BEMPTY + sb));
                    } else {
                        umma_commit(bar(BAR_AEMPTY + sa));
                        umma_commit(bar(BAR_BEMPTY + sb));
                    }
                }
                if (PAIR) umma_commit_pair(bar(BAR_ACC + (l & 1)));
                else umma_commit(bar(BAR_ACC + (l & 1)));
            }
        }
        __syncwarp();
        orlk::pdl_trigger();
    } else {
        // ---------------------------------------------------------------- epilogue warps
        const int grp = warp >= 6 ? 1 : 0;          // A stage (and chunk parity) this warp group produces
        const int q = warp & 3;                     // TMEM lane quadrant
        const int t = threadIdx.x - 64;             // 0..255
        const int row = q * 32 + lane;
        const int m = tile_m * BM + row;
#pragma unroll
        for (int l = 0; l < MAXL; ++l)
            if (l < L) bias_s[l * NMAX + t] = t < N ? __ldg(p.bias[l] + (int64_t)g * p.gs + t) : 0.f;
        headw_s[t] = t < N ? __ldg(p.head_w + (int64_t)g * p.gs + t) : 0.f;
        if (grp == 0) {         // lo part of the X tile: this warp's 32 rows = 256 float4 (hi and lo tiles share the layout)
            mbar_wait(bar(BAR_X), 0);
            const float4* xh = reinterpret_cast<const float4*>(a_hi(0)) + q * 256;
            float4* xl = reinterpret_cast<float4*>(a_lo(0)) + q * 256;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = xh[i * 32 + lane];
#pragma unroll
            for (int i = 0; i < 8; ++i) xl[i * 32 + lane] = make_float4(lo_of(v[i].x), lo_of(v[i].y), lo_of(v[i].z), lo_of(v[i].w));
            fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) {
                if (PAIR && rank != 0) mbar_arrive_remote(map_to_rank(bar(BAR_AFULL + 0), 0));
                else mbar_arrive(bar(BAR_AFULL + 0));
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");     // bias_s / headw_s complete
        if (t == 0) FZ_STAMP(2);

        float qacc = 0.f;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        int prev_stage = grp == 0 ? 0 : -1;         // stage of this warp's previous A fill (group 0 wrote the X tile's lo words)
        for (int l = 0; l < L; ++l) {
            const bool last = l == L - 1;
            mbar_wait(bar(BAR_ACC + (l & 1)), (l >> 1) & 1);
            tc_fence_after();
            if (t == 0) FZ_STAMP(112 + l);
            if (last) {
                orlk::pdl_trigger();                // every MMA of this strip has completed
                // the last layer's store tiles take the upper 128 KB of the ring (all of it is idle now), which in pair mode
                // includes an A stage: every warp's earlier stores must have read their tiles first
                if (store_h && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const float* bl = bias_s + l * NMAX;
            for (int c = grp; c < KS; c += 2) {
                uint32_t v[32];
                tmem_ld32(tlane + (uint32_t)(NMAX * (l & 1) + 32 * c), v);
                tmem_wait_ld();
                float x[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bl + 32 * c + 4 * j4);
                    x[4 * j4 + 0] = fmaxf(__uint_as_float(v[4 * j4 + 0]) + b4.x, 0.f);
                    x[4 * j4 + 1] = fmaxf(__uint_as_float(v[4 * j4 + 1]) + b4.y, 0.f);
                    x[4 * j4 + 2] = fmaxf(__uint_as_float(v[4 * j4 + 2]) + b4.z, 0.f);
                    x[4 * j4 + 3] = fmaxf(__uint_as_float(v[4 * j4 + 3]) + b4.w, 0.f);
                }
                if (p.bits != nullptr && m < p.M) {
                    uint32_t wbits = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) wbits |= (x[j] > 0.f ? 1u : 0u) << j;
                    p.bits[(((int64_t)l * p.G + g) * 8 + c) * p.M + m] = wbits;       // a warp's 32 rows: one 128-byte store
                }
                if (last) {
                    // every MMA has completed: the whole B ring is idle, so each of this warp's (up to four) chunks gets
                    // its own 4 KB store tile there and no store ever waits for the previous one
                    const float* hw = headw_s + 32 * c;
#pragma unroll
                    for (int j = 0; j < 32; ++j) qacc = fmaf(x[j], hw[j], qacc);
                    if (store_h) {
                        uint8_t* tile = base + (RING - 32 * 4096) + ((warp - 2) * 4 + (c >> 1)) * 4096;
                        float4* hrow = reinterpret_cast<float4*>(tile + lane * 128);
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            hrow[j4 ^ (lane & 7)] = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_4d(&maps.h[l], smem_u32(tile), 32 * c, tile_m * BM + q * 32, g, 0);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                    continue;
                }
                // A fill number 1 + l * KS + c (fill 0 was the X tile) goes to stage fill % NSA; the stage's previous content
                // was read by the MMAs of the slab NSA fills earlier and by this warp's own TMA store
                const int af = 1 + l * KS + c, sa = af % NSA;
                uint8_t* my_hi = a_hi(sa) + q * 4096;       // this warp's 32 rows of the stage: a 32 x 32 SWIZZLE_128B store tile
                uint8_t* my_lo = a_lo(sa) + q * 4096;
                if (lane == 0 && prev_stage >= 0) {         // this warp's stores of its previous fill have read their tiles
                    if (store_h) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    mbar_arrive(bar(BAR_AEMPTY + prev_stage));
                }
                prev_stage = sa;
                if (af >= NSA) mbar_wait(bar(BAR_AEMPTY + sa), (af / NSA - 1) & 1);
                __syncwarp();
                float4* hrow = reinterpret_cast<float4*>(my_hi + lane * 128);
                float4* lrow = reinterpret_cast<float4*>(my_lo + lane * 128);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    hrow[j4 ^ (lane & 7)] = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    lrow[j4 ^ (lane & 7)] = make_float4(lo_of(x[4 * j4]), lo_of(x[4 * j4 + 1]), lo_of(x[4 * j4 + 2]),
                                                        lo_of(x[4 * j4 + 3]));
                tc_fence_before();                  // the TMEM reads above are ordered before the next layer's MMAs
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR && rank != 0) mbar_arrive_remote(map_to_rank(bar(BAR_AFULL + sa), 0));
                    else mbar_arrive(bar(BAR_AFULL + sa));
                    if (store_h) {
                        tma_store_4d(&maps.h[l], smem_u32(my_hi), 32 * c, tile_m * BM + q * 32, g, 0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
        }
        // scalar head: even-chunk partial (warps 2-5) + odd-chunk partial (warps 6-9) + bias, fixed order
        if (grp == 1) qpart_s[row] = qacc;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (grp == 0 && m < p.M) p.out[(int64_t)g * p.out_gs + m] = (qacc + qpart_s[row]) + __ldg(p.head_b + (int64_t)g * p.gs);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the tiles must outlive the stores
    }
    if (threadIdx.x == 64) { FZ_STAMP(3); FZ_GSTAMP(6); }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();               // neither CTA leaves (or frees tensor memory) while the other may still use it
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_critic_bwd: the input-gradient chain of the same stack behind its scalar head, one launch:
//   dZ_{L-1}[m][k] = dq[m] * w_head[k] * relu'(H_{L-1}[m][k])          (generated, never stored)
//   dZ_{l-1} = (dZ_l W_l) * relu'(H_{l-1})    for l = L-1 .. 1        (stored: the weight gradients need them)
// (autograd of modules/critic_module.py:25-33 / nets/mlp.py:22-28 w.r.t. the hidden activations).  Same strip-per-CTA
// structure as the forward pass: the generator and each layer's epilogue write the next GEMM's A tiles (hi | lo) into
// the two-stage ring, the B operand is the transposed weight W_l^T [in][out] (kept by the Adam kernel) and its lo words
// through TMA, accumulators ping-pong in tensor memory.  relu' comes from the decision bits the forward pass left
// (4 bytes per 32 activations instead of re-reading H).
struct BwdMaps {
    CUtensorMap wt[MAXL - 1];     // W_l^T [G][N (in)][N (out)], l >= 1: box 32 (k = out) x N
    CUtensorMap wtlo[MAXL - 1];
    CUtensorMap dz[MAXL - 1];     // dZ_l [G][M][N], l = 0 .. L-2: box 32 x 32 (store)
};

struct BwdParams {
    int64_t gs;
    const float* dq; int64_t dq_gs;       // [G][M] upstream gradient of the head output
    const float* head_w;                  // [G][N] (member stride gs)
    const uint32_t* bits;                 // [L][G][8][M]
    int M, N, G, L, tiles_m;
    unsigned long long* trace;
};

template <bool PAIR>       // CTA pairs as in k_critic_fwd_t
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_critic_bwd_t(const __grid_constant__ BwdMaps maps, const __grid_constant__ BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int NSA = PAIR ? 3 : 2;
    constexpr int NSB = PAIR ? 3 : 2;
    constexpr int BST = PAIR ? B_STAGE / 2 : B_STAGE;
    uint8_t* b_base = base + NSA * A_STAGE;
    uint8_t* fixed = base + RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(fixed);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fixed + 192);
    float* headw_s = reinterpret_cast<float*>(fixed + 256);       // [NMAX]
    auto a_hi = [&](int s) { return base + s * A_STAGE; };
    auto a_lo = [&](int s) { return base + s * A_STAGE + A_TILE; };
    auto b_hi = [&](int s) { return b_base + s * BST; };
    auto b_lo = [&](int s) { return b_base + s * BST + BST / 2; };
    auto bar = [&](int i) { return smem_u32(&bars[i]); };

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x / p.tiles_m;
    const int tile_m = blockIdx.x - g * p.tiles_m;
    const int N = p.N, L = p.L;
    const int KS = N / BK;
    const int NG = L - 1;                       // GEMM layers of the chain
    const uint32_t rank = PAIR ? cluster_rank() : 0u;
    const int b_rows = PAIR ? N / 2 : N;

    if (threadIdx.x == 0) { FZ_STAMP(0); FZ_GSTAMP(4); }
    if (threadIdx.x == 32) {
        for (int t = 0; t < NG; ++t) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wt[t]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wtlo[t]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.dz[t]) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSA; ++s) {
            mbar_init(bar(BAR_AFULL + s), PAIR ? 8 : 4);
            mbar_init(bar(BAR_AEMPTY + s), 5);      // the MMAs' commit + the four writer warps (their TMA stores have read the tiles)
        }
        for (int s = 0; s < NSB; ++s) {
            mbar_init(bar(BAR_BFULL + s), 1);
            mbar_init(bar(BAR_BEMPTY + s), 1);
        }
        for (int s = 0; s < 2; ++s) mbar_init(bar(BAR_ACC + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    // Only dq (the loss launch right in front) and the decision bits depend on earlier launches of the step; the weight
    // slabs and the head weights were final long before (last step's Adam, this step's fused_prep): the producer starts
    // its TMA loads and the epilogue warps stage the head weights while the predecessor is still finishing, and only the
    // warps that read dq / bits wait for it.
    if (threadIdx.x == 0) { FZ_STAMP(1); FZ_GSTAMP(5); }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------------------ TMA producer: W_l^T k-slabs, l = L-1 .. 1
            const uint32_t tx = 2u * (uint32_t)N * BK * 4;
            int bi = 0;
            for (int t = 0; t < NG; ++t) {
                const int l = L - 1 - t;
                for (int j = 0; j < KS; ++j, ++bi) {
                    const int s = bi % NSB;
                    mbar_wait(bar(BAR_BEMPTY + s), ((bi / NSB) & 1) ^ 1);
                    if (bi < 32) FZ_STAMP(80 + bi);
                    if (PAIR) {
                        const uint32_t lb = map_to_rank(bar(BAR_BFULL + s), 0);
                        if (rank == 0) mbar_expect_tx(bar(BAR_BFULL + s), tx);
                        tma_load_3d_pair(smem_u32(b_hi(s)), &maps.wt[l - 1], lb, j * BK, (int)rank * b_rows, g);
                        tma_load_3d_pair(smem_u32(b_lo(s)), &maps.wtlo[l - 1], lb, j * BK, (int)rank * b_rows, g);
                    } else {
                        mbar_expect_tx(bar(BAR_BFULL + s), tx);
                        tma_load_3d(smem_u32(b_hi(s)), &maps.wt[l - 1], bar(BAR_BFULL + s), j * BK, 0, g);
                        tma_load_3d(smem_u32(b_lo(s)), &maps.wtlo[l - 1], bar(BAR_BFULL + s), j * BK, 0, g);
                    }
                }
            }
        }
        __syncwarp();
        orlk::pdl_trigger();
    } else if (warp == 1) {
        if ((!PAIR || rank == 0) && elect_one()) {
            // ------------------------------------------------------------ MMA issuer (the leader CTA's, for a pair)
            const uint32_t idesc = instr_desc_tf32(PAIR ? 2 * BM : BM, N);
            int bi = 0;
            for (int t = 0; t < NG; ++t) {
                const uint32_t acc = tmem_base + (uint32_t)(NMAX * (t & 1));
                for (int j = 0; j < KS; ++j, ++bi) {
                    const int sa = bi % NSA;
                    mbar_wait(bar(BAR_AFULL + sa), (bi / NSA) & 1);
                    if (bi < 32) FZ_STAMP(16 + bi);
                    const int sb = bi % NSB;
                    mbar_wait(bar(BAR_BFULL + sb), (bi / NSB) & 1);
                    if (bi < 32) FZ_STAMP(48 + bi);
                    tc_fence_after();
                    const uint64_t ad = smem_desc_sw128(smem_u32(a_hi(sa))), adl = smem_desc_sw128(smem_u32(a_lo(sa)));
                    const uint64_t bd = smem_desc_sw128(smem_u32(b_hi(sb))), bdl = smem_desc_sw128(smem_u32(b_lo(sb)));
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t ko = (uint64_t)(2 * k);
                        if (PAIR) {
                            umma_tf32_pair(acc, ad + ko, bd + ko, idesc, (j > 0 || k > 0) ? 1u : 0u);
                            umma_tf32_pair(acc, adl + ko, bd + ko, idesc, 1u);
                            umma_tf32_pair(acc, ad + ko, bdl + ko, idesc, 1u);
                        } else {
                            umma_tf32(acc, ad + ko, bd + ko, idesc, (j > 0 || k > 0) ? 1u : 0u);
                            umma_tf32(acc, adl + ko, bd + ko, idesc, 1u);
                            umma_tf32(acc, ad + ko, bdl + ko, idesc, 1u);
                        }
                    }
                    if (PAIR) {
                        umma_commit_pair(bar(BAR_AEMPTY + sa));
                        umma_commit_pair(bar(BAR_BEMPTY + sb));
                    } else {
                        umma_commit(bar(BAR_AEMPTY + sa));
                        umma_commit(bar(BAR_BEMPTY + sb));
                    }
                }
                if (PAIR) umma_commit_pair(bar(BAR_ACC + (t & 1)));
                else umma_commit(bar(BAR_ACC + (t & 1)));
            }
        }
        __syncwarp();
        orlk::pdl_trigger();
    } else {
        // ---------------------------------------------------------------- generator + epilogue warps
        const int grp = warp >= 6 ? 1 : 0;
        const int q = warp & 3;
        const int t256 = threadIdx.x - 64;
        const int row = q * 32 + lane;
        const int m = tile_m * BM + row;
        const bool row_ok = m < p.M;
        headw_s[t256] = t256 < N ? __ldg(p.head_w + (int64_t)g * p.gs + t256) : 0.f;
        orlk::pdl_wait();
        // (dq and the decision bits are written by earlier launches of the step: ordinary loads, NOT __ldg - a read-only
        // load may be hoisted above the wait, and then reads what the loss launch has not written yet)
        const float dq_m = row_ok ? __ldcg(p.dq + (int64_t)g * p.dq_gs + m) : 0.f;
        // ReLU-decision words of this row's chunks (c = grp, grp + 2, ...; at most four) of layer l
        auto load_bits = [&](int l, uint32_t (&mb)[4]) {
            const uint32_t* bp = p.bits + ((int64_t)l * p.G + g) * 8 * p.M + (row_ok ? m : 0);
#pragma unroll
            for (int i = 0; i < 4; ++i) mb[i] = (row_ok && grp + 2 * i < KS) ? __ldcg(bp + (int64_t)(grp + 2 * i) * p.M) : 0u;
        };
        uint32_t mb[4];
        load_bits(L - 1, mb);
        asm volatile("bar.sync 1, 256;" ::: "memory");     // headw_s complete
        if (t256 == 0) FZ_STAMP(2);

        int prev_stage = -1;                        // stage of this warp's previous A fill
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        // A fill number af (sequential over the chain: GEMM t, chunk c -> t * KS + c) goes to stage af % NSA.  Before
        // writing: confirm that this warp's stores of its previous fill have read their tiles (an arrival on THAT stage's
        // free barrier), then wait until the target stage is free (MMAs done + its previous writers' confirmations).
        auto claim = [&](int af) -> int {
            const int sa = af % NSA;
            if (lane == 0 && prev_stage >= 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(bar(BAR_AEMPTY + prev_stage));
            }
            prev_stage = sa;
            if (af >= NSA) mbar_wait(bar(BAR_AEMPTY + sa), (af / NSA - 1) & 1);
            __syncwarp();
            return sa;
        };
        auto write_a = [&](int sa, const float (&x)[32]) {      // this warp's 32 rows of the stage's hi and lo tiles
            uint8_t* my_hi = a_hi(sa) + q * 4096;
            uint8_t* my_lo = a_lo(sa) + q * 4096;
            float4* hrow = reinterpret_cast<float4*>(my_hi + lane * 128);
            float4* lrow = reinterpret_cast<float4*>(my_lo + lane * 128);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                hrow[j4 ^ (lane & 7)] = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                lrow[j4 ^ (lane & 7)] = make_float4(lo_of(x[4 * j4]), lo_of(x[4 * j4 + 1]), lo_of(x[4 * j4 + 2]),
                                                    lo_of(x[4 * j4 + 3]));
        };
        // ---- the generated operand of the first GEMM: dZ_{L-1} chunk by chunk, paced by the ring
        for (int c = grp, i = 0; c < KS; c += 2, ++i) {
            float x[32];
            const float* hw = headw_s + 32 * c;
            const uint32_t bits = mb[i];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = ((bits >> j) & 1u) ? dq_m * hw[j] : 0.f;
            const int sa = claim(c);
            write_a(sa, x);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && rank != 0) mbar_arrive_remote(map_to_rank(bar(BAR_AFULL + sa), 0));
                else mbar_arrive(bar(BAR_AFULL + sa));
            }
        }
        // ---- epilogues: dZ_{l-1} = acc * relu'(H_{l-1}); stored, and (unless it is the last) the next GEMM's operand
        for (int t = 0; t < NG; ++t) {
            const int l = L - 1 - t;                // the GEMM multiplied by W_l; its output is the gradient of layer l-1
            const bool last = t == NG - 1;
            load_bits(l - 1, mb);
            mbar_wait(bar(BAR_ACC + (t & 1)), (t >> 1) & 1);
            tc_fence_after();
            if (t256 == 0) FZ_STAMP(112 + t);
            if (last) {
                orlk::pdl_trigger();
                // the last layer's store tiles take the upper 128 KB of the (now idle) ring: every warp's earlier stores first
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            for (int c = grp, i = 0; c < KS; c += 2, ++i) {
                uint32_t v[32];
                tmem_ld32(tlane + (uint32_t)(NMAX * (t & 1) + 32 * c), v);
                tmem_wait_ld();
                float x[32];
                const uint32_t bits = mb[i];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = ((bits >> j) & 1u) ? __uint_as_float(v[j]) : 0.f;
                if (last) {     // every MMA has completed: own store tiles in the idle B ring, no store waits for another
                    uint8_t* tile = base + (RING - 32 * 4096) + ((warp - 2) * 4 + i) * 4096;
                    float4* hrow = reinterpret_cast<float4*>(tile + lane * 128);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        hrow[j4 ^ (lane & 7)] = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&maps.dz[l - 1], smem_u32(tile), 32 * c, tile_m * BM + q * 32, g, 0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    continue;
                }
                const int sa = claim((t + 1) * KS + c);
                write_a(sa, x);
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR && rank != 0) mbar_arrive_remote(map_to_rank(bar(BAR_AFULL + sa), 0));
                    else mbar_arrive(bar(BAR_AFULL + sa));
                    tma_store_4d(&maps.dz[l - 1], smem_u32(a_hi(sa) + q * 4096), 32 * c, tile_m * BM + q * 32, g, 0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (threadIdx.x == 64) { FZ_STAMP(3); FZ_GSTAMP(6); }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// Once per step, off the critical path: lo[i] = x[i] - trunc_tf32(x[i]) (exact in fp32) over a whole parameter arena - the
// second operand word of the 3xTF32 products - and the zero-padded copy [G][N][32] (+ lo words) of the first layer's
// [N][K0] weights, whose rows TMA cannot address (pitch K0 floats).
__global__ void k_fused_prep(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int nb_split,
                             const float* __restrict__ W0, int64_t gs, int N, int K0, int G, float* __restrict__ w0pad) {
    orlk::pdl_enter();
    if ((int)blockIdx.x < nb_split) {
        const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
        if (i + 3 < n) {
            const float4 v = *reinterpret_cast<const float4*>(src + i);
            *reinterpret_cast<float4*>(dst + i) = make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
        } else {
            for (int64_t j = i; j < n; ++j) dst[j] = lo_of(src[j]);
        }
        return;
    }
    const int64_t total = (int64_t)G * N * BK;
    const int64_t i = ((int64_t)blockIdx.x - nb_split) * blockDim.x + threadIdx.x;
    if (i < total) {
        const int k = (int)(i % BK);
        const int64_t gn = i / BK;
        const int nn = (int)(gn % N), g = (int)(gn / N);
        const float v = k < K0 ? src[(W0 - src) + (int64_t)g * gs + (int64_t)nn * K0 + k] : 0.f;
        w0pad[i] = v;
        w0pad[total + i] = lo_of(v);
    }
}

// the same for up to four arenas in one launch (blockIdx.y = arena): the step's derived copies of the online, target and
// transposed critic weights cost one launch slot at the start of the step instead of three
struct PrepJobs {
    OrlkFusedPrep j[4];
    int nb_split[4];
};

__global__ void k_fused_prep_multi(const PrepJobs P) {
    orlk::pdl_enter();
    const OrlkFusedPrep& q = P.j[blockIdx.y];
    const int nb_split = P.nb_split[blockIdx.y];
    if ((int)blockIdx.x < nb_split) {
        const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
        if (i + 3 < q.n) {
            const float4 v = *reinterpret_cast<const float4*>(q.src + i);
            *reinterpret_cast<float4*>(q.dst_lo + i) = make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
        } else {
            for (int64_t j = i; j < q.n; ++j) q.dst_lo[j] = lo_of(q.src[j]);
        }
        return;
    }
    if (q.W0 == nullptr) return;
    const int64_t total = (int64_t)q.G * q.N * BK;
    const int64_t i = ((int64_t)blockIdx.x - nb_split) * blockDim.x + threadIdx.x;
    if (i < total) {
        const int k = (int)(i % BK);
        const int64_t gn = i / BK;
        const int nn = (int)(gn % q.N), g = (int)(gn / q.N);
        const float v = k < q.K0 ? q.W0[(int64_t)g * q.gs + (int64_t)nn * q.K0 + k] : 0.f;
        q.w0pad[i] = v;
        q.w0pad[total + i] = lo_of(v);
    }
}

constexpr size_t FWD_SMEM = 1024 + RING + FIXED;

int fill_fwd_job(const OrlkFusedFwd* q, FwdMaps* maps, FwdParams* p, bool pair) {
    ORLK_REQUIRE(q->M > 0 && q->G > 0, "sizes");
    ORLK_REQUIRE(q->n_hidden >= 2 && q->n_hidden <= MAXL, "2..4 hidden layers");
    // (an even number of 32-column chunks: the two epilogue groups then alternate on every layer, which the stage-free
    // protocol - a writer confirms its stores when it reaches its next chunk - relies on)
    ORLK_REQUIRE(q->N >= 32 && q->N <= NMAX && (q->N == 32 || q->N % 64 == 0), "hidden width must be 32 or a multiple of 64 up to 256");
    ORLK_REQUIRE(q->K0 >= 1 && q->K0 <= BK, "first-layer fan-in must be <= 32");
    ORLK_REQUIRE(q->ldx % 4 == 0 && aligned16(q->X), "X rows must be 16-byte aligned");
    ORLK_REQUIRE(q->gs % 4 == 0 && q->h_gs % 4 == 0, "member strides must be multiples of 4 floats");
    ORLK_REQUIRE(q->head_w != nullptr && q->head_b != nullptr && q->out != nullptr, "scalar head");
    ORLK_REQUIRE(q->W0pad != nullptr && q->W0pad_lo != nullptr && aligned16(q->W0pad) && aligned16(q->W0pad_lo), "padded first-layer weights");
    memset(maps, 0, sizeof(*maps));
    memset(p, 0, sizeof(*p));
    int rc = make_map(&maps->x, q->X, q->ldx, 0, q->M, q->K0, 1, BM);
    if (rc) return rc;
    const int wbox = pair ? q->N / 2 : q->N;        // weight rows per TMA box: a CTA of a pair loads half of every slab
    rc = make_map(&maps->w0, q->W0pad, BK, (int64_t)q->N * BK, q->N, BK, q->G, wbox);
    if (rc) return rc;
    rc = make_map(&maps->w0lo, q->W0pad_lo, BK, (int64_t)q->N * BK, q->N, BK, q->G, wbox);
    if (rc) return rc;
    const bool store = q->H[0] != nullptr;
    for (int l = 0; l < q->n_hidden; ++l) {
        ORLK_REQUIRE(q->bias[l] != nullptr, "bias pointers");
        ORLK_REQUIRE((q->H[l] != nullptr) == store, "H: all layers or none");
        if (l >= 1) {
            ORLK_REQUIRE(q->W[l] != nullptr && q->Wlo[l] != nullptr && aligned16(q->W[l]) && aligned16(q->Wlo[l]),
                         "hidden weights (and lo copies) must be 16-byte aligned");
            rc = make_map(&maps->w[l - 1], q->W[l], q->N, q->gs, q->N, q->N, q->G, wbox);
            if (rc) return rc;
            rc = make_map(&maps->wlo[l - 1], q->Wlo[l], q->N, q->gs, q->N, q->N, q->G, wbox);
            if (rc) return rc;
        }
        if (store) {
            ORLK_REQUIRE(aligned16(q->H[l]), "H must be 16-byte aligned");
            rc = make_map_c(&maps->h[l], q->H[l], q->N, q->h_gs, 0, q->M, q->N, q->G, 1);
            if (rc) return rc;
        }
        p->bias[l] = q->bias[l];
    }
    p->gs = q->gs;
    p->head_w = q->head_w; p->head_b = q->head_b;
    p->out = q->out; p->out_gs = q->out_gs;
    p->bits = q->relu_bits;
    p->M = q->M; p->N = q->N; p->G = q->G; p->L = q->n_hidden;
    p->tiles_m = (q->M + BM - 1) / BM;
    if (pair) p->tiles_m = (p->tiles_m + 1) & ~1;   // whole pairs per member: an odd last strip gets a partner that is all padding
    p->store_h = store ? 1 : 0;
    p->trace = orlk::trace_buffer();
    { const char* e = getenv("ORLK_FUSED_NO_STORE"); p->no_store = e ? atoi(e) : 0; }
    return 0;
}

}  // namespace

extern "C" int orlk_sizeof_fused_fwd(void) { return (int)sizeof(OrlkFusedFwd); }

extern "C" int orlk_sizeof_fused_bwd(void) { return (int)sizeof(OrlkFusedBwd); }

extern "C" int orlk_fused_init(void) {
    int rc = check(cudaFuncSetAttribute(k_critic_fwd_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM), "smem attr k_critic_fwd");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(k_critic_fwd_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM), "smem attr k_critic_fwd (pairs)");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(k_critic_bwd_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM), "smem attr k_critic_bwd");
    if (rc) return rc;
    return check(cudaFuncSetAttribute(k_critic_bwd_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM), "smem attr k_critic_bwd (pairs)");
}

extern "C" int orlk_critic_bwd_fused(const OrlkFusedBwd* q, void* stream) {
    ORLK_REQUIRE(q != nullptr && q->M > 0 && q->G > 0, "sizes");
    ORLK_REQUIRE(q->n_hidden >= 2 && q->n_hidden <= MAXL, "2..4 hidden layers");
    ORLK_REQUIRE(q->N >= 32 && q->N <= NMAX && (q->N == 32 || q->N % 64 == 0), "hidden width must be 32 or a multiple of 64 up to 256");
    ORLK_REQUIRE(q->gs % 4 == 0 && q->dz_gs % 4 == 0, "member strides must be multiples of 4 floats");
    ORLK_REQUIRE(q->dq != nullptr && q->head_w != nullptr && q->relu_bits != nullptr, "dq, head weights, ReLU bits");
    BwdMaps maps;
    BwdParams p;
    const bool pair = (q->flags & ORLK_FUSED_PAIRS) != 0 && q->N % 64 == 0;
    const int wbox = pair ? q->N / 2 : q->N;
    memset(&maps, 0, sizeof(maps));
    memset(&p, 0, sizeof(p));
    for (int l = 1; l < q->n_hidden; ++l) {
        ORLK_REQUIRE(q->WT[l] != nullptr && q->WTlo[l] != nullptr && aligned16(q->WT[l]) && aligned16(q->WTlo[l]),
                     "transposed weights (and lo copies) must be 16-byte aligned");
        ORLK_REQUIRE(q->dZ[l - 1] != nullptr && aligned16(q->dZ[l - 1]), "dZ must be 16-byte aligned");
        int rc = make_map(&maps.wt[l - 1], q->WT[l], q->N, q->gs, q->N, q->N, q->G, wbox);
        if (rc) return rc;
        rc = make_map(&maps.wtlo[l - 1], q->WTlo[l], q->N, q->gs, q->N, q->N, q->G, wbox);
        if (rc) return rc;
        rc = make_map_c(&maps.dz[l - 1], q->dZ[l - 1], q->N, q->dz_gs, 0, q->M, q->N, q->G, 1);
        if (rc) return rc;
    }
    p.gs = q->gs;
    p.dq = q->dq; p.dq_gs = q->dq_gs;
    p.head_w = q->head_w;
    p.bits = q->relu_bits;
    p.M = q->M; p.N = q->N; p.G = q->G; p.L = q->n_hidden;
    p.tiles_m = (q->M + BM - 1) / BM;
    if (pair) p.tiles_m = (p.tiles_m + 1) & ~1;
    p.trace = orlk::trace_buffer();
    if (pair) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(q->G * p.tiles_m);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = FWD_SMEM;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = orlk::pdl_enabled() ? 2 : 1;
        cudaLaunchKernelEx(&cfg, k_critic_bwd_t<true>, maps, p);
        return check_launch("k_critic_bwd (pairs)");
    }
    orlk::launch(k_critic_bwd_t<false>, dim3(q->G * p.tiles_m), dim3(NUM_THREADS), FWD_SMEM, (cudaStream_t)stream, maps, p);
    return check_launch("k_critic_bwd");
}

extern "C" int orlk_fused_prep(const float* src, float* dst_lo, int64_t n, const float* W0, int64_t gs, int N, int K0, int G,
                               float* w0pad, void* stream) {
    ORLK_REQUIRE(src != nullptr && dst_lo != nullptr && n > 0, "fused_prep arguments");
    ORLK_REQUIRE(aligned16(src) && aligned16(dst_lo), "fused_prep needs 16-byte aligned arrays");
    ORLK_REQUIRE(W0 == nullptr || (w0pad != nullptr && N > 0 && K0 > 0 && K0 <= BK && G > 0 && W0 >= src && W0 < src + n),
                 "first-layer weights must lie inside the arena");
    const int64_t vec = (n + 3) / 4;
    const int nb_split = (int)((vec + 255) / 256);
    const int nb_pad = W0 != nullptr ? (int)(((int64_t)G * N * BK + 255) / 256) : 0;
    orlk::launch(k_fused_prep, dim3((unsigned)(nb_split + nb_pad)), dim3(256), 0, (cudaStream_t)stream, src, dst_lo, n, nb_split, W0, gs,
                 N, K0, G, w0pad);
    return check_launch("k_fused_prep");
}

extern "C" int orlk_sizeof_fused_prep(void) { return (int)sizeof(OrlkFusedPrep); }

extern "C" int orlk_fused_prep_multi(const OrlkFusedPrep* jobs, int n_jobs, void* stream) {
    ORLK_REQUIRE(jobs != nullptr && n_jobs >= 1 && n_jobs <= 4, "1..4 arenas");
    PrepJobs P;
    memset(&P, 0, sizeof(P));
    int nbx = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const OrlkFusedPrep& q = jobs[j];
        ORLK_REQUIRE(q.src != nullptr && q.dst_lo != nullptr && q.n > 0 && aligned16(q.src) && aligned16(q.dst_lo), "fused_prep arguments");
        ORLK_REQUIRE(q.W0 == nullptr || (q.w0pad != nullptr && q.N > 0 && q.K0 > 0 && q.K0 <= BK && q.G > 0), "first-layer weights");
        P.j[j] = q;
        P.nb_split[j] = (int)(((q.n + 3) / 4 + 255) / 256);
        const int nb = P.nb_split[j] + (q.W0 != nullptr ? (int)(((int64_t)q.G * q.N * BK + 255) / 256) : 0);
        nbx = nb > nbx ? nb : nbx;
    }
    orlk::launch(k_fused_prep_multi, dim3((unsigned)nbx, (unsigned)n_jobs), dim3(256), 0, (cudaStream_t)stream, P);
    return check_launch("k_fused_prep_multi");
}

extern "C" int orlk_critic_fwd_fused(const OrlkFusedFwd* jobs, int n_jobs, void* stream) {
    ORLK_REQUIRE(jobs != nullptr && (n_jobs == 1 || n_jobs == 2), "one or two jobs");
    FwdMaps maps[2];
    FwdParams p[2];
    bool pair = (jobs[0].flags & ORLK_FUSED_PAIRS) != 0;
    for (int j = 0; j < n_jobs; ++j) pair = pair && jobs[j].N % 64 == 0;
    for (int j = 0; j < 2; ++j) {
        const int rc = fill_fwd_job(&jobs[j < n_jobs ? j : 0], &maps[j], &p[j], pair);
        if (rc) return rc;
    }
    const int ctas0 = p[0].G * p[0].tiles_m;
    const int grid = ctas0 + (n_jobs == 2 ? p[1].G * p[1].tiles_m : 0);
    if (pair) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = FWD_SMEM;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = orlk::pdl_enabled() ? 2 : 1;
        cudaLaunchKernelEx(&cfg, k_critic_fwd_t<true>, maps[0], maps[1], p[0], p[1], ctas0);
        return check_launch("k_critic_fwd (pairs)");
    }
    orlk::launch(k_critic_fwd_t<false>, dim3(grid), dim3(NUM_THREADS), FWD_SMEM, (cudaStream_t)stream, maps[0], maps[1], p[0], p[1], ctas0);
    return check_launch("k_critic_fwd");
}
