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

template <int BT, int BK, int NT>
struct TileLoader {
    // A BT x BK operand tile staged in registers, then committed to smem as S[k][t].
    static constexpr int E = BT * BK / NT;  // elements per thread
    static_assert(E % 4 == 0, "tile must give each thread whole float4s");
    static constexpr int V = E / 4;

    // kcontig: operand(t,k) = base[t*ld + k]   else: operand(t,k) = base[k*ld + t]
    __device__ __forceinline__ static void fetch(float (&r)[E], const float* __restrict__ base, int64_t ld, bool kcontig,
                                                 bool vec, int t0, int T, int k0, int kend, int tid) {
        if (vec) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int q = tid + i * NT;
                const float* p;
                if (kcontig) {
                    const int tt = q / (BK / 4), k4 = (q % (BK / 4)) * 4;
                    p = base + (int64_t)(t0 + tt) * ld + (k0 + k4);
                } else {
                    const int kk = q / (BT / 4), t4 = (q % (BT / 4)) * 4;
                    p = base + (int64_t)(k0 + kk) * ld + (t0 + t4);
                }
                const float4 v = __ldg(reinterpret_cast<const float4*>(p));
                r[4 * i + 0] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int e = tid + i * NT;
                int tt, kk;
                if (kcontig) { kk = e % BK; tt = e / BK; } else { tt = e % BT; kk = e / BT; }
                const int t = t0 + tt, k = k0 + kk;
                float v = 0.f;
                if (t < T && k < kend) v = kcontig ? __ldg(base + (int64_t)t * ld + k) : __ldg(base + (int64_t)k * ld + t);
                r[i] = v;
            }
        }
    }

    __device__ __forceinline__ static void commit(const float (&r)[E], float (*S)[BT + PAD], bool kcontig, bool vec, int tid) {
        if (vec) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int q = tid + i * NT;
                if (kcontig) {
                    const int tt = q / (BK / 4), k4 = (q % (BK / 4)) * 4;
                    S[k4 + 0][tt] = r[4 * i + 0]; S[k4 + 1][tt] = r[4 * i + 1];
                    S[k4 + 2][tt] = r[4 * i + 2]; S[k4 + 3][tt] = r[4 * i + 3];
                } else {
                    const int kk = q / (BT / 4), t4 = (q % (BT / 4)) * 4;
                    *reinterpret_cast<float4*>(&S[kk][t4]) = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int e = tid + i * NT;
                int tt, kk;
                if (kcontig) { kk = e % BK; tt = e / BK; } else { tt = e % BT; kk = e / BT; }
                S[kk][tt] = r[i];
            }
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

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), (BM >= 128 ? 2 : 3))
k_gemm_grouped(const OrlkGemmDesc* __restrict__ descs, int n_descs) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TX = BN / TN;
    using LA = TileLoader<BM, BK, NT>;
    using LB = TileLoader<BN, BK, NT>;

    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
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

    const bool a_kc = d.a_layout == 0, b_kc = d.b_layout == 1;
    const bool chunks_full = ((kend - kbeg) % BK) == 0 && (kbeg % 4) == 0;
    const bool vecA = chunks_full && aligned16(d.A) && (d.lda % 4) == 0 && (m0 + BM <= M);
    const bool vecB = chunks_full && aligned16(d.B) && (d.ldb % 4) == 0 && (n0 + BN <= N);

    const int tx = tid % TX, ty = tid / TX;
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

    float ra[LA::E], rb[LB::E];
    const int nk = (kend - kbeg + BK - 1) / BK;
    if (nk > 0) {
        LA::fetch(ra, d.A, d.lda, a_kc, vecA, m0, M, kbeg, kend, tid);
        LB::fetch(rb, d.B, d.ldb, b_kc, vecB, n0, N, kbeg, kend, tid);
        LA::commit(ra, As[0], a_kc, vecA, tid);
        LB::commit(rb, Bs[0], b_kc, vecB, tid);
    }
    __syncthreads();
    for (int it = 0; it < nk; ++it) {
        const int cur = it & 1;
        const bool more = it + 1 < nk;
        if (more) {
            const int k0 = kbeg + (it + 1) * BK;
            LA::fetch(ra, d.A, d.lda, a_kc, vecA, m0, M, k0, kend, tid);
            LB::fetch(rb, d.B, d.ldb, b_kc, vecB, n0, N, k0, kend, tid);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
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
        if (more) {
            LA::commit(ra, As[cur ^ 1], a_kc, vecA, tid);
            LB::commit(rb, Bs[cur ^ 1], b_kc, vecB, tid);
        }
        __syncthreads();
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

}  // namespace

extern "C" int orlk_gemm_grouped(const OrlkGemmDesc* descs_dev, int n_descs, int total_tiles, int cfg, void* stream) {
    ORLK_REQUIRE(descs_dev != nullptr && n_descs > 0, "descs");
    ORLK_REQUIRE(total_tiles > 0, "total_tiles");
    cudaStream_t s = (cudaStream_t)stream;
    switch (cfg) {
        case ORLK_CFG_BIG: k_gemm_grouped<128, 128, 16, 8, 8><<<total_tiles, 256, 0, s>>>(descs_dev, n_descs); break;
        case ORLK_CFG_MID: k_gemm_grouped<64, 64, 16, 4, 4><<<total_tiles, 256, 0, s>>>(descs_dev, n_descs); break;
        case ORLK_CFG_SMALL: k_gemm_grouped<32, 32, 32, 2, 2><<<total_tiles, 256, 0, s>>>(descs_dev, n_descs); break;
        default: set_error("unknown gemm cfg %d", cfg); return ORLK_ERR_BAD_ARG;
    }
    return check_launch("k_gemm_grouped");
}
