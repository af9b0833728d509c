// Fused small-row MLP chain: a whole forward (or backward) pass of a few-hundred-row batch through a stack of
// <= 256-wide layers in ONE launch.
//
// Standalone, each of those layers is a ~3 us kernel body behind ~2.5 us of launch / prologue / drain, and a CQL step
// has ~25 of them back to back.  Here a thread-block CLUSTER of 8 CTAs owns one 16-row strip of the batch for the whole
// chain: CTA r computes the 16 x 32 output tile of columns [32r, 32r+32) of every stage and stores it to global memory
// (the activations / gradients are needed by the weight-gradient launches anyway); after ONE hardware cluster barrier
// (barrier.cluster release / acquire, ~0.2 us) every CTA of the strip fetches the finished 16-row strip back from L2
// with 16-byte cp.async copies as the next stage's A operand.  A strip is 16 KB: re-reading it from L2 (~250 cycles of
// latency, ~64 B/cycle/SM) costs less than pushing it to the seven peers over distributed shared memory (~20 B/cycle/SM
// on this part; the first version of this kernel did that and lost 2.5 us per stage to it).  The weight tile of stage
// s+1 is prefetched into the second half of a double buffer while stage s is being multiplied.  Different strips and
// different networks (twin critics) never synchronise with each other; 256 rows = 16 strips = 128 CTAs per network.
// Arithmetic: warp-level TF32 MMAs (mma.sync m16n8k8) on fragments read straight from the staged tiles, 3xTF32 hi/lo
// split for fp32-grade results (passes == 3) or single pass; 16 warps = 4 column tiles x 4 k groups, fixed-order sum of
// the 4 partial tiles.  Stage descriptors are OrlkGemmDesc (same epilogues as orlk_gemm_grouped), A always row-major
// [m][k]; they travel in the kernel parameters.  Replaces the per-layer launches of nets/mlp.py:22,28 forward and
// autograd dgrad for small batches.
#include <stdlib.h>
#include "orlk_common.cuh"
using namespace orlk;

namespace {

constexpr int TM = 16, TN = 32;   // output tile of one CTA
constexpr int CL = 8;             // CTAs per cluster = column tiles per stage (N <= 256)
constexpr int KC = 256;           // max k of a stage
constexpr int KP = KC + 4;        // pitch of a k-contiguous tile row (floats): 260 % 32 == 4 -> conflict-free fragment loads
constexpr int PN = TN + 8;        // pitch of an n-contiguous B tile row: 40 floats (40 % 32 == 8 -> conflict-free)
constexpr int NTHR = 512;
constexpr int KG = 4;             // k groups (of four warps: one per 8-column tile)
constexpr int PR = TN + 1;        // pitch of a partial tile row
constexpr int MAX_DESC = 24;      // chains x stages per launch (kernel-parameter space)
constexpr int A_FLOATS = TM * KP;
constexpr int B_FLOATS = (TN * KP > KC * PN) ? TN * KP : KC * PN;
constexpr int R_FLOATS = KG * TM * PR;
static_assert(TM * TN == NTHR, "the epilogue gives every thread one element of the tile");

struct ChainArgs {
    OrlkGemmDesc d[MAX_DESC];     // [chain][stage]
    int n_chains, n_stages, tiles_m, passes, passes0;   // passes0: arithmetic of stage 0 (raw inputs may want 3 when the rest is 1)
    unsigned long long* trace;    // profiling aid (orlk_tc_set_trace): 16 clock stamps per CTA, NULL in normal operation
};
#define CHAIN_STAMP(slot)                                                                                     \
    do {                                                                                                      \
        if (P.trace != nullptr && threadIdx.x == 0 && (slot) < 16)                                            \
            P.trace[(int64_t)blockIdx.x * 16 + (slot)] = (unsigned long long)clock64();                      \
    } while (0)

__device__ __forceinline__ void cp_async16(float* dst, const float* src, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x) { return __float_as_uint(x - __uint_as_float(tf32_hi(x))); }

// Rows [t0, t0+ROWS) x k [0, 8 ceil(K/8)) of a k-contiguous operand (base[t*ld + k]) into S[t][KP], by all 16 warps:
// warp w takes rows w, w+16, ..., its lanes the 16-byte chunks of the row (no integer divisions: the address arithmetic of
// the first version cost more than the copies).  Outside the matrix (t >= T or k >= K): zeros.  The global reads bypass
// L1 (cp.async.cg / ld.global.cg): the A strip of a stage was written by the other CTAs of the cluster a barrier ago.
template <int ROWS>
__device__ __forceinline__ void stage_kc(float* S, const float* __restrict__ base, int64_t ld, int t0, int T, int K, int tid) {
    const int nb = 2 * ((K + 7) >> 3);                  // 16-byte chunks per row
    const int w = tid >> 5, lane = tid & 31;
    if (aligned16(base) && (ld % 4) == 0) {
#pragma unroll
        for (int r = w; r < ROWS; r += NTHR / 32) {
            const int t = t0 + r;
            const float* row = base + (int64_t)min(t, T - 1) * ld;
            for (int c = lane; c < nb; c += 32) {
                const int k = 4 * c;
                const int bytes = (t < T && k < K) ? 4 * min(4, K - k) : 0;
                cp_async16(S + r * KP + k, row + (bytes ? k : 0), bytes);
            }
        }
    } else {
#pragma unroll
        for (int r = w; r < ROWS; r += NTHR / 32) {
            const int t = t0 + r;
            for (int k = lane; k < 4 * nb; k += 32) S[r * KP + k] = (t < T && k < K) ? __ldcg(base + (int64_t)t * ld + k) : 0.f;
        }
    }
}
// k [0, 8 ceil(K/8)) x columns [n0, n0+TN) of an n-contiguous operand (base[k*ld + n]) into S[k][PN], by all threads:
// a warp covers 4 k rows x 8 chunks per sweep.
__device__ __forceinline__ void stage_nc(float* S, const float* __restrict__ base, int64_t ld, int n0, int N, int K, int tid) {
    const int nk = 8 * ((K + 7) >> 3);
    if (aligned16(base) && (ld % 4) == 0) {
        const int c = tid & 7;
        const int n = n0 + 4 * c;
        for (int k = tid >> 3; k < nk; k += NTHR / 8) {
            const int bytes = (k < K && n < N) ? 4 * min(4, N - n) : 0;
            cp_async16(S + k * PN + 4 * c, base + (int64_t)min(k, K - 1) * ld + (bytes ? n : 0), bytes);
        }
    } else {
        const int c = tid & 31;
        const int n = n0 + c;
        for (int k = tid >> 5; k < nk; k += NTHR / 32) S[k * PN + c] = (k < K && n < N) ? __ldcg(base + (int64_t)k * ld + n) : 0.f;
    }
}

__device__ __forceinline__ void stage_b(float* Bs, const OrlkGemmDesc& d, int n0, int tid) {
    if (n0 >= d.N) return;
    if (d.b_layout == 1) stage_kc<TN>(Bs, d.B, d.ldb, n0, d.N, d.K, tid);
    else stage_nc(Bs, d.B, d.ldb, n0, d.N, d.K, tid);
}

// Measured (profiles/chain_trace_r02.txt, 256 rows, 3xTF32): 3.0 us per stage = 0.35 us to issue the strip's 32 warp-wide
// 16-byte cp.async copies (the LSU takes 8 cycles per LDGSTS) + 0.3 us for them to land + 0.55 us to issue the next weight
// tile's 64 + 1.05 us of MMAs (legacy mma.sync issues once per 16 cycles per SM sub-partition: 0.8 us for the 96 of a
// sub-partition) + 0.2 us reduce / epilogue + 0.5 us for the tile's stores to be released + 0.1 us barrier.  Tried and
// rejected: cp.async.bulk for the 1 KB rows (per-lane addresses make the compiler serialise the UBLKCP issue over the
// lanes: 0.7 us for 16 rows) and slicing the weight prefetch between the k steps (the LDGSTS then stall each warp's own
// fragment loads: 2.1 us of MMA phase).  A four-stage pass costs 17-18 us as one launch against ~22 us as four
// PDL-chained small-row launches, but the passes that feed fused head / sampler / backward-entry launches lose those
// fusions, and twin-critic passes (256 CTAs, two per SM) wait 1.5-2.5 us per barrier for their co-resident CTA: the CQL
// step is 299.6 us with the chain against 291.8 us without, so it stays OFF by default (ORLK_CHAIN=1 enables it).

__global__ void __launch_bounds__(NTHR, 2)
k_chain_gemm(const __grid_constant__ ChainArgs P) {
    extern __shared__ float4 smem_f4[];
    float* As = reinterpret_cast<float*>(smem_f4);      // A strip of the current stage  [TM][KP]
    float* B0 = As + A_FLOATS;                          // weight tiles of even / odd stages: b_layout 1: [TN][KP], 0: [KC][PN]
    float* B1 = B0 + B_FLOATS;
    float* red = B1 + B_FLOATS;                         // [KG][TM][PR] partial tiles

    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int nt = wi & 3, kg = wi >> 2;                // 8-column tile, k group
    const int gid = lane >> 2, tig = lane & 3;
    const int rank = blockIdx.x % CL;                   // cluster dims (8,1,1): rank == %cluster_ctarank
    const int strip = blockIdx.x / CL;
    const int chain = strip / P.tiles_m, tm = strip - chain * P.tiles_m;
    const int m0 = tm * TM, n0 = rank * TN;
    const OrlkGemmDesc* D = P.d + chain * P.n_stages;
    const int er = tid >> 5, ec = tid & 31;             // my element of the tile in the epilogue
    CHAIN_STAMP(0);
    orlk::pdl_wait();                                   // nothing above touched global data
    CHAIN_STAMP(1);

    // stage 0 operands from global memory: the input strip (every CTA of the cluster) and my weight tile
    stage_kc<TM>(As, D[0].A, D[0].lda, m0, D[0].M, D[0].K, tid);
    stage_b(B0, D[0], n0, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int s = 0; s < P.n_stages; ++s) {
        const OrlkGemmDesc& d = D[s];
        const int M = d.M, N = d.N, K = d.K;
        const bool active = n0 < N;
        const bool last = s + 1 == P.n_stages;
        float* Bs = (s & 1) ? B1 : B0;
        // my output element: bias / mask operands prefetched off the critical path
        const int em = m0 + er, en = n0 + ec;
        const bool e_ok = active && em < M && en < N;
        const float e_bias = (e_ok && d.bias != nullptr) ? __ldg(d.bias + en) : 0.f;
        const float e_aux = (e_ok && d.aux != nullptr) ? __ldcg(d.aux + (int64_t)em * d.ldaux + en) : 0.f;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (s == 1) CHAIN_STAMP(8);
        __syncthreads();                                // A strip and this stage's weight tile landed
        if (s == 0) CHAIN_STAMP(2);
        if (s == 1) CHAIN_STAMP(9);
        if (s == 2) CHAIN_STAMP(14);
        if (!last) {                                    // next stage's weight tile streams in under this stage's MMAs
            stage_b((s & 1) ? B0 : B1, D[s + 1], n0, tid);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (s == 1) CHAIN_STAMP(7);
        if (active) {
            // ---- MMAs: k group kg owns 1/4 of the k steps, its four warps the four 8-column tiles
            // three independent accumulators (hi*hi, lo*hi, hi*lo): consecutive MMAs do not wait for each other
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
            const int nstep = (K + 7) >> 3, per = (nstep + KG - 1) / KG;
            const int s_lo = min(nstep, kg * per), s_hi = min(nstep, s_lo + per);
            const bool b_kc = d.b_layout == 1;
            const bool split3 = (s == 0 ? P.passes0 : P.passes) == 3;
            const int c = nt * 8 + gid;
            const float* ap = As + gid * KP + tig;
            const float* bp = b_kc ? Bs + c * KP + tig : Bs + tig * PN + c;
            const int bk = b_kc ? 1 : PN;               // stride of one k in the B tile
#pragma unroll 4
            for (int st = s_lo; st < s_hi; ++st) {
                const int k = 8 * st;
                const float a0 = ap[k], a1 = ap[8 * KP + k], a2 = ap[k + 4], a3 = ap[8 * KP + k + 4];
                const float b0 = bp[k * bk], b1 = bp[(k + 4) * bk];
                const uint32_t ah[4] = {tf32_hi(a0), tf32_hi(a1), tf32_hi(a2), tf32_hi(a3)};
                const uint32_t bh0 = tf32_hi(b0), bh1 = tf32_hi(b1);
                mma_tf32(c0, ah, bh0, bh1);
                if (split3) {
                    const uint32_t al[4] = {tf32_lo(a0), tf32_lo(a1), tf32_lo(a2), tf32_lo(a3)};
                    mma_tf32(c1, al, bh0, bh1);
                    mma_tf32(c2, ah, tf32_lo(b0), tf32_lo(b1));
                }
            }
            float cacc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) cacc[i] = c0[i] + (c1[i] + c2[i]);
            // partial tile -> shared memory
            float* r0 = red + (kg * TM + gid) * PR + nt * 8 + 2 * tig;
            r0[0] = cacc[0];
            r0[1] = cacc[1];
            r0[8 * PR] = cacc[2];
            r0[8 * PR + 1] = cacc[3];
        }
        if (last) orlk::pdl_trigger();                  // the next kernel's CTAs may take their SMs from here on
        __syncthreads();                                // partials complete; everybody is done with As and Bs
        if (s == 0) CHAIN_STAMP(3);
        if (s == 1) CHAIN_STAMP(10);
        if (s == 2) CHAIN_STAMP(15);
        if (active) {
            const int epi = d.epi;
            float v = 0.f;
#pragma unroll
            for (int g = 0; g < KG; ++g) v += red[(g * TM + er) * PR + ec];      // fixed order: bit-reproducible
            if (em < M && en < N) {
                v += e_bias;
                if (epi == ORLK_EPI_SWISH && d.C2 != nullptr) d.C2[(int64_t)em * d.ldc + en] = v;
                switch (epi) {
                    case ORLK_EPI_RELU: v = fmaxf(v, 0.f); break;
                    case ORLK_EPI_RELU_MASK: v = e_aux > 0.f ? v : 0.f; break;
                    case ORLK_EPI_SWISH: v = v / (1.f + expf(-v)); break;
                    case ORLK_EPI_DSWISH: {
                        const float sg = 1.f / (1.f + expf(-e_aux));
                        v = v * (sg * (1.f + e_aux * (1.f - sg)));
                        break;
                    }
                    default: break;
                }
                d.C[(int64_t)em * d.ldc + en] = v;
            }
        }
        if (s == 0) CHAIN_STAMP(4);
        if (s == 1) CHAIN_STAMP(11);
        if (!last) {
            // every CTA of the strip, working or not: my tile is in L2, the peers' tiles are visible after the barrier
            cluster_arrive();
            if (s == 1) CHAIN_STAMP(12);
            cluster_wait();
            if (s == 0) CHAIN_STAMP(5);
            if (s == 1) CHAIN_STAMP(13);
            const OrlkGemmDesc& dn = D[s + 1];
            stage_kc<TM>(As, dn.A, dn.lda, m0, dn.M, dn.K, tid);     // the strip all eight CTAs have just finished
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (s == 0) CHAIN_STAMP(6);
        }
    }
}

constexpr size_t chain_smem() { return sizeof(float) * (A_FLOATS + 2 * B_FLOATS + R_FLOATS); }

}  // namespace

extern "C" int orlk_gemm_chain_init(void) {
    int rc = check(cudaFuncSetAttribute(k_chain_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem()), "chain smem attr");
    if (rc) return rc;
    if (getenv("ORLK_GRAPH_DEBUG")) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(16 * CL);
        cfg.blockDim = dim3(NTHR);
        cfg.dynamicSmemBytes = chain_smem();
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_chain_gemm, &cfg);
        fprintf(stderr, "[orlk] chain kernel: max co-resident clusters of %d = %d (%s), %zu bytes smem\n", CL, n, cudaGetErrorString(e),
                chain_smem());
    }
    return 0;
}

// descs_host[chain * n_stages + stage]; every chain has the same M; stage s+1 reads what stage s wrote (A of s+1 == C of s).
// Requirements per stage: a_layout 0, K <= 256, N <= 256, k_splits <= 1, C != NULL, no row / column sums, no CT.
extern "C" int orlk_gemm_chain(const OrlkGemmDesc* descs_host, int n_chains, int n_stages, int passes, int passes_stage0,
                               void* stream) {
    ORLK_REQUIRE(descs_host != nullptr && n_chains > 0 && n_stages > 0 && n_chains * n_stages <= MAX_DESC, "1..24 stage descriptors");
    ORLK_REQUIRE((passes == 1 || passes == 3) && (passes_stage0 == 1 || passes_stage0 == 3),
                 "passes must be 1 or 3 (the fp32 FFMA mode launches the layers one by one)");
    ChainArgs args;
    args.n_chains = n_chains;
    args.n_stages = n_stages;
    args.passes = passes;
    args.passes0 = passes_stage0;
    args.trace = orlk::trace_buffer();
    const int M = descs_host[0].M;
    args.tiles_m = (M + TM - 1) / TM;
    for (int i = 0; i < n_chains * n_stages; ++i) {
        const OrlkGemmDesc& d = descs_host[i];
        ORLK_REQUIRE(d.M == M, "all stages of a chain launch share M");
        ORLK_REQUIRE(d.a_layout == 0, "chain stages take A row-major [m][k]");
        ORLK_REQUIRE(d.K >= 1 && d.K <= KC && d.N >= 1 && d.N <= CL * TN, "chain stages need K <= 256 and N <= 256");
        ORLK_REQUIRE(d.k_splits <= 1 && d.C != nullptr && d.rowsum == nullptr && d.colsum == nullptr && d.CT == nullptr,
                     "chain stages: no split-K, no sums, no transposed copy");
        args.d[i] = d;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(n_chains * args.tiles_m * CL));
    cfg.blockDim = dim3(NTHR);
    cfg.dynamicSmemBytes = chain_smem();
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    int rc = check(cudaLaunchKernelEx(&cfg, k_chain_gemm, args), "k_chain_gemm launch");
    if (rc) return rc;
    return check_launch("k_chain_gemm");
}
