// Grouped fp32 GEMM with fused epilogues (SIMT FFMA path).
//
// One launch covers a list of independent problems (twin critics, ensemble members, all wgrads of a
// network, ...); every CTA owns one BMxBN output tile of one problem and, for split-K problems, one k-chunk.
// This is the fp32-parity path of the engine: the reference runs full-FP32 SGEMM
// (torch.backends.cuda.matmul.allow_tf32 is False), so sums are accumulated with FFMA in
// ascending k.  Replaces nn.Linear / torch.einsum forward and the autograd dgrad / wgrad GEMMs
// (nets/mlp.py:22,28; nets/ensemble_linear.py:35,37).
#include "orlk_common.cuh"
using namespace orlk;

namespace {

constexpr int PAD = 4;

// A BT x BK operand tile: staged in registers (vector path: whole float4s, all loads in flight before the first
// store) and committed to shared memory as S[k][t]; ragged / unaligned tiles take a compact element-wise path that
// goes straight to shared memory.  KC: operand(t,k) = base[t*ld + k] (k contiguous), else base[k*ld + t].
// The code paths are kept small on purpose: these kernels run ~1 us per CTA and are instruction-fetch bound
// when the straight-line code exceeds the instruction cache (ncu: stall_no_inst).
template <int BT, int BK, int NT, bool KC>
struct TileLoader {
    static constexpr int E = BT * BK / NT;  // elements per thread
    static_assert(E % 4 == 0, "tile must give each thread whole float4s");
    static constexpr int V = E / 4;

    __device__ __forceinline__ static void fetch(float4 (&r)[V], const float* __restrict__ base, int64_t ld, int t0, int k0,
                                                 int tid) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int q = tid + i * NT;
            const float* p;
            if (KC) {       // consecutive lanes -> consecutive tile rows (conflict-free transposed stores below)
                const int tt = q % BT, k4 = (q / BT) * 4;
                p = base + (int64_t)(t0 + tt) * ld + (k0 + k4);
            } else {
                const int kk = q / (BT / 4), t4 = (q % (BT / 4)) * 4;
                p = base + (int64_t)(k0 + kk) * ld + (t0 + t4);
            }
            r[i] = __ldg(reinterpret_cast<const float4*>(p));
        }
    }

    __device__ __forceinline__ static void commit(const float4 (&r)[V], float (*S)[BT + PAD], int tid) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int q = tid + i * NT;
            if (KC) {
                const int tt = q % BT, k4 = (q / BT) * 4;
                S[k4 + 0][tt] = r[i].x; S[k4 + 1][tt] = r[i].y; S[k4 + 2][tt] = r[i].z; S[k4 + 3][tt] = r[i].w;
            } else {
                const int kk = q / (BT / 4), t4 = (q % (BT / 4)) * 4;
                *reinterpret_cast<float4*>(&S[kk][t4]) = r[i];
            }
        }
    }

    // element-wise, bounds-checked, zero-filling (not unrolled: rarely taken, keeps the kernel small)
    __device__ __noinline__ static void stage_scalar(float (*S)[BT + PAD], const float* __restrict__ base, int64_t ld, int t0,
                                                     int T, int k0, int kend, int tid) {
#pragma unroll 1
        for (int e = tid; e < BT * BK; e += NT) {
            int tt, kk;
            if (KC) { kk = e % BK; tt = e / BK; } else { tt = e % BT; kk = e / BT; }
            const int t = t0 + tt, k = k0 + kk;
            float v = 0.f;
            if (t < T && k < kend) v = KC ? __ldg(base + (int64_t)t * ld + k) : __ldg(base + (int64_t)k * ld + t);
            S[kk][tt] = v;
        }
    }
};

// Position of micro-tile element i of thread coordinate c inside a BT-wide tile.
template <int BT, int TT>
__device__ __forceinline__ int frag_pos(int c, int i) {
    if constexpr (TT == 8) return (i < 4) ? (c * 4 + i) : (BT / 2 + c * 4 + (i - 4));
    else return c * TT + i;
}

template <int BT, int TT>
__device__ __forceinline__ void load_frag(float (&f)[TT], const float* __restrict__ row, int c) {
    if constexpr (TT == 8) {
        const float4 lo = *reinterpret_cast<const float4*>(row + c * 4);
        const float4 hi = *reinterpret_cast<const float4*>(row + BT / 2 + c * 4);
        f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = lo.w; f[4] = hi.x; f[5] = hi.y; f[6] = hi.z; f[7] = hi.w;
    } else if constexpr (TT == 4) {
        const float4 v = *reinterpret_cast<const float4*>(row + c * 4);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    } else {
        static_assert(TT == 2, "micro-tile width");
        const float2 v = *reinterpret_cast<const float2*>(row + c * 2);
        f[0] = v.x; f[1] = v.y;
    }
}

__device__ __forceinline__ float epilogue(float v, int epi, float aux) {
    switch (epi) {
        case ORLK_EPI_RELU: return fmaxf(v, 0.f);
        case ORLK_EPI_RELU_MASK: return aux > 0.f ? v : 0.f;
        case ORLK_EPI_SWISH: return v / (1.f + expf(-v));
        case ORLK_EPI_DSWISH: {
            const float s = 1.f / (1.f + expf(-aux));
            return v * (s * (1.f + aux * (1.f - s)));
        }
        default: return v;
    }
}

// KG > 1: the CTA's threads form KG groups that each own the whole BM x BN tile for 1/KG of every k-chunk and are
// summed through shared memory at the end (k-parallel: for M <= 512 the layer is latency-bound, so the k chain is
// cut in KG pieces and the whole K extent is fetched in one shot, NBUF = 1).
template <int BM, int BN, int BK, int TM, int TN, int KG, int NBUF, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(KG * (BM / TM) * (BN / TN), (KG > 1 ? 1 : (BM >= 128 ? 2 : 3)))
k_gemm_grouped(const OrlkGemmDesc* __restrict__ descs, int n_descs) {
    orlk::pdl_enter();
    constexpr int NTG = (BM / TM) * (BN / TN);      // threads per k-group
    constexpr int NT = KG * NTG;
    constexpr int TX = BN / TN;
    constexpr int KSUB = BK / KG;
    using LA = TileLoader<BM, BK, NT, A_KC>;
    using LB = TileLoader<BN, BK, NT, B_KC>;

    extern __shared__ float4 smem_f4[];
    float* smem = reinterpret_cast<float*>(smem_f4);
    float (*As)[BK][BM + PAD] = reinterpret_cast<float (*)[BK][BM + PAD]>(smem);
    float (*Bs)[BK][BN + PAD] = reinterpret_cast<float (*)[BK][BN + PAD]>(smem + NBUF * BK * (BM + PAD));
    __shared__ OrlkGemmDesc sd;

    const int tid = threadIdx.x;
    if (tid == 0) {
        int p = 0;
        const int tile = blockIdx.x;
        while (p + 1 < n_descs && descs[p + 1].tile_start <= tile) ++p;
        sd = descs[p];
    }
    __syncthreads();
    const OrlkGemmDesc& d = sd;

    int t = blockIdx.x - d.tile_start;
    const int per = d.tiles_m * d.tiles_n;
    const int split = t / per;
    t -= split * per;
    const int tm = t / d.tiles_n, tn = t % d.tiles_n;
    const int m0 = tm * BM, n0 = tn * BN;
    const int kbeg = split * d.k_chunk;
    const int kend = min(d.K, kbeg + d.k_chunk);
    const int M = d.M, N = d.N;

    const bool chunks_full = ((kend - kbeg) % BK) == 0 && (kbeg % 4) == 0;
    const bool vecA = chunks_full && aligned16(d.A) && (d.lda % 4) == 0 && (m0 + BM <= M);
    const bool vecB = chunks_full && aligned16(d.B) && (d.ldb % 4) == 0 && (n0 + BN <= N);

    const int kg = tid / NTG, tg = tid % NTG;
    const int tx = tg % TX, ty = tg / TX;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float rs[TM], cs[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) rs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) cs[j] = 0.f;
    const bool do_rs = d.rowsum != nullptr && tn == 0;
    const bool do_cs = d.colsum != nullptr && tm == 0;

    float4 ra[LA::V], rb[LB::V];
    const int nk = (kend - kbeg + BK - 1) / BK;
    // software pipeline with ONE instance of every stage: iteration `it` fetches chunk it+1 into registers, computes
    // chunk it from shared memory, then commits chunk it+1 (it = -1 only stages chunk 0).
    for (int it = -1; it < nk; ++it) {
        const bool more = it + 1 < nk;
        const int k_next = kbeg + (it + 1) * BK;
        const int nxt = (NBUF == 2) ? ((it + 1) & 1) : 0;
        if (more) {
            if (vecA) LA::fetch(ra, d.A, d.lda, m0, k_next, tid);
            if (vecB) LB::fetch(rb, d.B, d.ldb, n0, k_next, tid);
        }
        if (it >= 0) {
            const int cur = (NBUF == 2) ? (it & 1) : 0;
            // this group's slice of the chunk, clipped to the valid k range (the rest of the tile holds zeros)
            const int kvalid = min(BK, kend - (kbeg + it * BK));
            const int k_lo = kg * KSUB, k_hi = min(k_lo + KSUB, kvalid);
#pragma unroll 4
            for (int kk = k_lo; kk < k_hi; ++kk) {
                float a[TM], b[TN];
                load_frag<BM, TM>(a, As[cur][kk], ty);
                load_frag<BN, TN>(b, Bs[cur][kk], tx);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                if (do_rs) {
#pragma unroll
                    for (int i = 0; i < TM; ++i) rs[i] += a[i];
                }
                if (do_cs) {
#pragma unroll
                    for (int j = 0; j < TN; ++j) cs[j] += b[j];
                }
            }
            if (NBUF == 1) __syncthreads();     // everyone is done reading the single buffer
        }
        if (more) {
            if (vecA) LA::commit(ra, As[nxt], tid); else LA::stage_scalar(As[nxt], d.A, d.lda, m0, M, k_next, kend, tid);
            if (vecB) LB::commit(rb, Bs[nxt], tid); else LB::stage_scalar(Bs[nxt], d.B, d.ldb, n0, N, k_next, kend, tid);
        }
        __syncthreads();
    }

    if (KG > 1) {
        // fixed-order sum of the KG partial tiles (the operand tiles are dead now: reuse their shared memory)
        constexpr int PER = TM * TN + TM + TN;
        float* red = smem;
        if (kg > 0) {
            float* r = red + ((kg - 1) * NTG + tg) * PER;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) r[i * TN + j] = acc[i][j];
#pragma unroll
            for (int i = 0; i < TM; ++i) r[TM * TN + i] = rs[i];
#pragma unroll
            for (int j = 0; j < TN; ++j) r[TM * TN + TM + j] = cs[j];
        }
        __syncthreads();
        if (kg > 0) return;
        for (int g2 = 1; g2 < KG; ++g2) {
            const float* r = red + ((g2 - 1) * NTG + tg) * PER;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] += r[i * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i) rs[i] += r[TM * TN + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) cs[j] += r[TM * TN + TM + j];
        }
    }

    // ---- epilogue
    const int slot = d.split_base + split;
    float* __restrict__ C = d.C + (int64_t)slot * d.c_split_stride;
    const int epi = d.epi;
    const bool has_aux = d.aux != nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + frag_pos<BM, TM>(ty, i);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + frag_pos<BN, TN>(tx, j);
            if (n >= N) continue;
            float v = acc[i][j];
            if (d.bias != nullptr) v += __ldg(d.bias + n);
            const float ax = has_aux ? __ldg(d.aux + (int64_t)m * d.ldaux + n) : 0.f;
            if (epi == ORLK_EPI_SWISH && d.C2 != nullptr) d.C2[(int64_t)m * d.ldc + n] = v;
            const float r = epilogue(v, epi, ax);
            C[(int64_t)m * d.ldc + n] = r;
            if (d.CT != nullptr) d.CT[(int64_t)n * d.ldct + m] = r;
        }
    }
    if (do_rs && tx == 0) {
        float* out = d.rowsum + (int64_t)slot * d.sum_split_stride;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = m0 + frag_pos<BM, TM>(ty, i);
            if (m < M) out[m] = rs[i];
        }
    }
    if (do_cs && ty == 0) {
        float* out = d.colsum + (int64_t)slot * d.sum_split_stride;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + frag_pos<BN, TN>(tx, j);
            if (n < N) out[n] = cs[j];
        }
    }
}

template <int BM, int BN, int BK, int NBUF>
constexpr size_t gemm_smem_bytes() { return sizeof(float) * NBUF * BK * ((BM + PAD) + (BN + PAD)); }

}  // namespace

template <int BM, int BN, int BK, int TM, int TN, int KG, int NBUF>
int launch_cfg(const OrlkGemmDesc* descs_dev, int n_descs, int total_tiles, int a_layout, int b_layout, cudaStream_t s) {
    constexpr int NT = KG * (BM / TM) * (BN / TN);
    constexpr size_t smem = gemm_smem_bytes<BM, BN, BK, NBUF>();
    const bool a_kc = a_layout == 0, b_kc = b_layout == 1;
    if (a_kc && b_kc) orlk::launch(k_gemm_grouped<BM, BN, BK, TM, TN, KG, NBUF, true, true>, total_tiles, NT, smem, s, descs_dev, n_descs);
    else if (a_kc) orlk::launch(k_gemm_grouped<BM, BN, BK, TM, TN, KG, NBUF, true, false>, total_tiles, NT, smem, s, descs_dev, n_descs);
    else if (b_kc) orlk::launch(k_gemm_grouped<BM, BN, BK, TM, TN, KG, NBUF, false, true>, total_tiles, NT, smem, s, descs_dev, n_descs);
    else orlk::launch(k_gemm_grouped<BM, BN, BK, TM, TN, KG, NBUF, false, false>, total_tiles, NT, smem, s, descs_dev, n_descs);
    return check_launch("k_gemm_grouped");
}

extern "C" int orlk_gemm_init(void) {
    // the k-parallel configuration stages a whole 32 x 256 slab of both operands: opt in to > 48 KB of shared memory
    constexpr int smem = (int)gemm_smem_bytes<32, 32, 256, 1>();
    int rc = check(cudaFuncSetAttribute(k_gemm_grouped<32, 32, 256, 4, 4, 4, 1, true, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attr");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(k_gemm_grouped<32, 32, 256, 4, 4, 4, 1, true, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attr");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(k_gemm_grouped<32, 32, 256, 4, 4, 4, 1, false, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attr");
    if (rc) return rc;
    return check(cudaFuncSetAttribute(k_gemm_grouped<32, 32, 256, 4, 4, 4, 1, false, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attr");
}

// All problems of one launch must share the operand layouts (a_layout, b_layout): the kernel is specialised on them.
extern "C" int orlk_gemm_grouped(const OrlkGemmDesc* descs_dev, int n_descs, int total_tiles, int cfg, int a_layout,
                                  int b_layout, void* stream) {
    ORLK_REQUIRE(descs_dev != nullptr && n_descs > 0, "descs");
    ORLK_REQUIRE(total_tiles > 0, "total_tiles");
    cudaStream_t s = (cudaStream_t)stream;
    switch (cfg) {
        case ORLK_CFG_BIG: return launch_cfg<128, 128, 16, 8, 8, 1, 2>(descs_dev, n_descs, total_tiles, a_layout, b_layout, s);
        case ORLK_CFG_MID: return launch_cfg<64, 64, 16, 4, 4, 1, 2>(descs_dev, n_descs, total_tiles, a_layout, b_layout, s);
        case ORLK_CFG_SMALL: return launch_cfg<32, 32, 32, 2, 2, 1, 2>(descs_dev, n_descs, total_tiles, a_layout, b_layout, s);
        case ORLK_CFG_KPAR: return launch_cfg<32, 32, 256, 4, 4, 4, 1>(descs_dev, n_descs, total_tiles, a_layout, b_layout, s);
        default: set_error("unknown gemm cfg %d", cfg); return ORLK_ERR_BAD_ARG;
    }
}
