// Fused small-row MLP chain: a whole forward (or backward) pass of a few-hundred-row batch through a stack of
// <= 256-wide layers in ONE launch.
//
// Standalone, each of those layers is a ~3 us kernel body behind ~3 us of launch / prologue / drain, and a CQL step has
// ~25 of them back to back.  Here a thread-block CLUSTER of 4 CTAs owns one 32-row strip of the batch for the whole
// chain: CTA r computes the 32 x 64 output tile of columns [64r, 64r+64) of every stage, stores it to global memory
// (the activations / gradients are needed by the weight-gradient launches anyway) and PUSHES it over distributed shared
// memory into the next stage's A strip of all four CTAs, so the next stage starts from shared memory after one cluster
// barrier - no trip through L2.  The next stage's weight tile is prefetched (cp.async) while the current stage is
// reduced and pushed.  Different strips and different networks (twin critics) never synchronise with each other.
// Arithmetic: warp-level TF32 MMAs (mma.sync m16n8k8) on fragments read straight from the staged tiles, 3xTF32 hi/lo
// split for fp32-grade results (passes == 3) or single pass; 8 k-groups of two warps (one per 32-column half of the
// tile), fixed-order sum of the 8 partial tiles.  Stage descriptors are OrlkGemmDesc (same epilogues as
// orlk_gemm_grouped), A always row-major [m][k]; they travel in the kernel parameters.  Replaces the per-layer launches
// of nets/mlp.py:22,28 forward and autograd dgrad for small batches.
#include <stdlib.h>
#include "orlk_common.cuh"
using namespace orlk;

namespace {

constexpr int TM = 32, TN = 64;   // output tile of one CTA
constexpr int CL = 4;             // CTAs per cluster = column tiles per stage (N <= 256)
constexpr int KC = 256;           // max k of a stage
constexpr int KP = KC + 4;        // pitch of a k-contiguous tile row (floats): 260 % 32 == 4 -> conflict-free fragment loads
constexpr int PN = TN + 8;        // pitch of an n-contiguous B tile row: 72 floats (72 % 32 == 8 -> conflict-free)
constexpr int NTHR = 512;
constexpr int KG = 8;             // k groups of two warps
constexpr int PR = TN + 1;        // pitch of a partial tile row
constexpr int MAX_DESC = 16;      // chains x stages per launch (kernel-parameter space)
constexpr int A_FLOATS = TM * KP;
constexpr int B_FLOATS = (TN * KP > KC * PN) ? TN * KP : KC * PN;
constexpr int R_FLOATS = KG * TM * PR;

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
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store into the same shared-memory location of CTA `rank` of this cluster
__device__ __forceinline__ void dsmem_store(float* local, int rank, float v) {
    const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x) { return __float_as_uint(x - __uint_as_float(tf32_hi(x))); }

// Rows [t0, t0+ROWS) x k [0, 8 ceil(K/8)) of a k-contiguous operand (base[t*ld + k]) into S[t][KP], by all threads.
// Outside the matrix (t >= T or k >= K): zeros.
template <int ROWS>
__device__ __forceinline__ void stage_kc(float* S, const float* __restrict__ base, int64_t ld, int t0, int T, int K, int tid) {
    const int nb = 2 * ((K + 7) >> 3);                  // 16-byte chunks per row
    if (aligned16(base) && (ld % 4) == 0) {
        for (int idx = tid; idx < ROWS * nb; idx += NTHR) {
            const int r = idx / nb, c = idx - r * nb;
            const int t = t0 + r, k = 4 * c;
            const int bytes = (t < T && k < K) ? 4 * min(4, K - k) : 0;
            cp_async16(S + r * KP + k, base + (int64_t)min(t, T - 1) * ld + (bytes ? k : 0), bytes);
        }
    } else {
        for (int q = tid; q < ROWS * nb * 4; q += NTHR) {
            const int r = q / (nb * 4), k = q - r * (nb * 4);
            const int t = t0 + r;
            S[r * KP + k] = (t < T && k < K) ? __ldg(base + (int64_t)t * ld + k) : 0.f;
        }
    }
}
// k [0, 8 ceil(K/8)) x columns [n0, n0+TN) of an n-contiguous operand (base[k*ld + n]) into S[k][PN], by all threads.
__device__ __forceinline__ void stage_nc(float* S, const float* __restrict__ base, int64_t ld, int n0, int N, int K, int tid) {
    const int nk = 8 * ((K + 7) >> 3);
    if (aligned16(base) && (ld % 4) == 0) {
        for (int idx = tid; idx < nk * (TN / 4); idx += NTHR) {
            const int k = idx / (TN / 4), c = idx - k * (TN / 4);
            const int n = n0 + 4 * c;
            const int bytes = (k < K && n < N) ? 4 * min(4, N - n) : 0;
            cp_async16(S + k * PN + 4 * c, base + (int64_t)min(k, K - 1) * ld + (bytes ? n : 0), bytes);
        }
    } else {
        for (int q = tid; q < nk * TN; q += NTHR) {
            const int k = q / TN, c = q - k * TN;
            const int n = n0 + c;
            S[k * PN + c] = (k < K && n < N) ? __ldg(base + (int64_t)k * ld + n) : 0.f;
        }
    }
}

__device__ __forceinline__ void stage_b(float* Bs, const OrlkGemmDesc& d, int n0, int tid) {
    if (n0 >= d.N) return;
    if (d.b_layout == 1) stage_kc<TN>(Bs, d.B, d.ldb, n0, d.N, d.K, tid);
    else stage_nc(Bs, d.B, d.ldb, n0, d.N, d.K, tid);
}

__global__ void __launch_bounds__(NTHR, 1)
k_chain_gemm(const __grid_constant__ ChainArgs P) {
    extern __shared__ float4 smem_f4[];
    float* A0 = reinterpret_cast<float*>(smem_f4);      // A strip of even stages  [TM][KP]
    float* A1 = A0 + A_FLOATS;                          // A strip of odd stages (filled by the cluster's pushes)
    float* Bs = A1 + A_FLOATS;                          // b_layout 1: [TN][KP]   b_layout 0: [KC][PN]
    float* red = Bs + B_FLOATS;                         // [KG][TM][PR] partial tiles

    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int kg = wi >> 1, wsub = wi & 1;              // k group, 32-column half of the tile
    const int gid = lane >> 2, tig = lane & 3;
    const int rank = blockIdx.x % CL;                   // cluster dims (4,1,1): rank == %cluster_ctarank
    const int strip = blockIdx.x / CL;
    const int chain = strip / P.tiles_m, tm = strip - chain * P.tiles_m;
    const int m0 = tm * TM, n0 = rank * TN;
    const OrlkGemmDesc* D = P.d + chain * P.n_stages;
    CHAIN_STAMP(0);
    orlk::pdl_enter();
    CHAIN_STAMP(1);

    // stage 0 operands from global memory: the whole input strip (every CTA of the cluster) and my weight tile
    stage_kc<TM>(A0, D[0].A, D[0].lda, m0, D[0].M, D[0].K, tid);
    stage_b(Bs, D[0], n0, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int s = 0; s < P.n_stages; ++s) {
        const OrlkGemmDesc& d = D[s];
        const int M = d.M, N = d.N, K = d.K;
        const bool active = n0 < N;
        const bool last = s + 1 == P.n_stages;
        float* As = (s & 1) ? A1 : A0;
        float* An = (s & 1) ? A0 : A1;                  // the next stage's strip
        // my four output elements: bias / mask operands prefetched off the critical path
        float e_bias[4], e_aux[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = tid + j * NTHR, er = e >> 6, ec = e & 63;
            const int em = m0 + er, en = n0 + ec;
            const bool ok = active && em < M && en < N;
            e_bias[j] = (ok && d.bias != nullptr) ? __ldg(d.bias + en) : 0.f;
            e_aux[j] = (ok && d.aux != nullptr) ? __ldg(d.aux + (int64_t)em * d.ldaux + en) : 0.f;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                // weight tile (and, for stage 0, the input strip) landed
        if (s < 3) CHAIN_STAMP(2 + 4 * s);
        if (active) {
            // ---- MMAs: k group kg owns 1/8 of the k steps, its two warps the two 32-column halves of the tile
            float cacc[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) cacc[mt][nt][0] = cacc[mt][nt][1] = cacc[mt][nt][2] = cacc[mt][nt][3] = 0.f;
            const int nstep = (K + 7) >> 3, per8 = (nstep + KG - 1) / KG;
            const int s_lo = min(nstep, kg * per8), s_hi = min(nstep, s_lo + per8);
            const bool b_kc = d.b_layout == 1;
            const bool split3 = (s == 0 ? P.passes0 : P.passes) == 3;
            for (int st = s_lo; st < s_hi; ++st) {
                const int k = 8 * st;
                uint32_t ah[2][4], al[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int r = mt * 16 + gid;
                    const float a0 = As[r * KP + k + tig], a1 = As[(r + 8) * KP + k + tig];
                    const float a2 = As[r * KP + k + tig + 4], a3 = As[(r + 8) * KP + k + tig + 4];
                    ah[mt][0] = tf32_hi(a0); ah[mt][1] = tf32_hi(a1); ah[mt][2] = tf32_hi(a2); ah[mt][3] = tf32_hi(a3);
                    al[mt][0] = tf32_lo(a0); al[mt][1] = tf32_lo(a1); al[mt][2] = tf32_lo(a2); al[mt][3] = tf32_lo(a3);
                }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int c = wsub * 32 + nt * 8 + gid;
                    float b0, b1;
                    if (b_kc) {
                        b0 = Bs[c * KP + k + tig];
                        b1 = Bs[c * KP + k + tig + 4];
                    } else {
                        b0 = Bs[(k + tig) * PN + c];
                        b1 = Bs[(k + tig + 4) * PN + c];
                    }
                    const uint32_t bh0 = tf32_hi(b0), bh1 = tf32_hi(b1);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma_tf32(cacc[mt][nt], ah[mt], bh0, bh1);
                        if (split3) {
                            mma_tf32(cacc[mt][nt], al[mt], bh0, bh1);
                            mma_tf32(cacc[mt][nt], ah[mt], tf32_lo(b0), tf32_lo(b1));
                        }
                    }
                }
            }
            // partial tiles -> shared memory
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    float* r0 = red + (kg * TM + mt * 16 + gid) * PR + wsub * 32 + nt * 8 + 2 * tig;
                    r0[0] = cacc[mt][nt][0];
                    r0[1] = cacc[mt][nt][1];
                    r0[8 * PR] = cacc[mt][nt][2];
                    r0[8 * PR + 1] = cacc[mt][nt][3];
                }
        }
        __syncthreads();                                // partials complete; everybody is done with Bs
        if (s < 3) CHAIN_STAMP(3 + 4 * s);
        if (!last) {                                    // next stage's weight tile streams in behind the reduction
            stage_b(Bs, D[s + 1], n0, tid);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (active) {
            const int epi = d.epi;
            const int kn = last ? 0 : 8 * ((N + 7) >> 3);     // k extent the next stage will read (zero padded)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = tid + j * NTHR, er = e >> 6, ec = e & 63;
                const int em = m0 + er, en = n0 + ec;
                float v = 0.f;
#pragma unroll
                for (int g = 0; g < KG; ++g) v += red[(g * TM + er) * PR + ec];      // fixed order: bit-reproducible
                if (en < N) {
                    v += e_bias[j];
                    if (epi == ORLK_EPI_SWISH && d.C2 != nullptr && em < M) d.C2[(int64_t)em * d.ldc + en] = v;
                    switch (epi) {
                        case ORLK_EPI_RELU: v = fmaxf(v, 0.f); break;
                        case ORLK_EPI_RELU_MASK: v = e_aux[j] > 0.f ? v : 0.f; break;
                        case ORLK_EPI_SWISH: v = v / (1.f + expf(-v)); break;
                        case ORLK_EPI_DSWISH: {
                            const float sg = 1.f / (1.f + expf(-e_aux[j]));
                            v = v * (sg * (1.f + e_aux[j] * (1.f - sg)));
                            break;
                        }
                        default: break;
                    }
                    if (em < M) d.C[(int64_t)em * d.ldc + en] = v;
                    else v = 0.f;                       // rows past M stay zero all the way down the chain
                } else v = 0.f;
                if (en < kn) {                          // push into every CTA's next-stage strip (column en = its k index)
#pragma unroll
                    for (int r = 0; r < CL; ++r) dsmem_store(An + er * KP + en, r, v);
                }
            }
        }
        if (s < 3) CHAIN_STAMP(4 + 4 * s);
        // every CTA of the strip, working or not: the pushes have landed and the shared-memory buffers may be reused
        if (!last) cluster_sync();
        if (s < 3) CHAIN_STAMP(5 + 4 * s);
    }
}

constexpr size_t chain_smem() { return sizeof(float) * (2 * A_FLOATS + B_FLOATS + R_FLOATS); }

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
    ORLK_REQUIRE(descs_host != nullptr && n_chains > 0 && n_stages > 0 && n_chains * n_stages <= MAX_DESC, "1..16 stage descriptors");
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
