// First-layer ("narrow-K") and head ("narrow-N") kernels for large row counts.
//
// The first critic layer has K = obs+act = 14..23 inputs and the heads have 1..12 outputs: these are not GEMM
// shaped (a 128x128x16 tile would be mostly padding), they are bandwidth-bound streaming kernels over the
// [M, 256] activation / gradient matrices.  One thread owns one of the 256 wide columns, the narrow operand of a
// row tile sits in shared memory and is broadcast.  Reference call sites: the first nn.Linear of
// Critic.backbone (critic_module.py:25-26 -> mlp.py:22) forward and weight gradient; Critic.last weight gradient.
#include "orlk_common.cuh"
using namespace orlk;

namespace {

constexpr int NK_MAX = 32;     // max narrow dimension
constexpr int ROWS_FWD = 16;   // rows per register sub-tile in the forward kernel
constexpr int SUB_FWD = 4;     // sub-tiles per block (64 rows per block: the layer's weights are staged once per block)
constexpr int XLD = NK_MAX + 4; // shared-memory row stride: 16-byte aligned rows for float4 broadcast loads
constexpr int ROWS_WG = 128;   // rows per block (= per partial slot) in the weight-gradient kernel

// Y[g][m][n] = act(b[g][n] + sum_k X[g][m][k] * W[g][n*ldw + k]),   K <= 32;  optional YT[g][n][m]
__global__ void __launch_bounds__(256, 2)
k_narrow_fwd(const float* __restrict__ X, int64_t ldx, int64_t x_gs, const float* __restrict__ W, int64_t ldw, int64_t w_gs,
             const float* __restrict__ b, int64_t b_gs, float* __restrict__ Y, int64_t ldy, int64_t y_gs,
             float* __restrict__ YT, int64_t ldyt, int64_t yt_gs, int M, int N, int K, int relu) {
    orlk::pdl_enter();
    __shared__ __align__(16) float xs[ROWS_FWD * SUB_FWD][XLD];
    __shared__ float ws[256 * (NK_MAX + 1)];           // this block's 256 weight rows, row stride K|1 (odd: conflict-free)
    const int g = blockIdx.z;
    const int mb = blockIdx.x * ROWS_FWD * SUB_FWD;
    const int n0 = blockIdx.y * 256;
    const int n = n0 + threadIdx.x;
    const float* Xg = X + g * x_gs;
    for (int i = threadIdx.x; i < ROWS_FWD * SUB_FWD * NK_MAX; i += 256) {
        const int r = i / NK_MAX, k = i % NK_MAX;          // zero-padded to 32 columns: the k loop runs in float4 steps
        xs[r][k] = (k < K && mb + r < M) ? __ldg(Xg + (int64_t)(mb + r) * ldx + k) : 0.f;
    }
    // The weight rows of the block are one contiguous run when ldw == K: read it with consecutive threads on consecutive
    // words (a per-thread row walk costs 32 sectors per instruction and was the bottleneck of this kernel).
    const int ks = K | 1;
    const int nrows = min(256, N - n0);
    const float* Wg = W + g * w_gs + (int64_t)n0 * ldw;
    if (ldw == K) {
        for (int i = threadIdx.x; i < nrows * K; i += 256) {
            const int r = i / K, k = i - r * K;
            ws[r * ks + k] = __ldg(Wg + i);
        }
    } else {
        for (int i = threadIdx.x; i < nrows * K; i += 256) {
            const int r = i / K, k = i - r * K;
            ws[r * ks + k] = __ldg(Wg + (int64_t)r * ldw + k);
        }
    }
    const bool n_ok = n < N;
    const float bias = (n_ok && b != nullptr) ? __ldg(b + g * b_gs + n) : 0.f;
    __syncthreads();
    float w[NK_MAX];
#pragma unroll
    for (int k = 0; k < NK_MAX; ++k) w[k] = (n_ok && k < K) ? ws[threadIdx.x * ks + k] : 0.f;
    const bool yt_vec = YT != nullptr && (ldyt % 4) == 0 && (yt_gs % 4) == 0 && aligned16(YT);
#pragma unroll 1
    for (int sub = 0; sub < SUB_FWD; ++sub) {
        const int m0 = mb + sub * ROWS_FWD;
        if (m0 >= M) break;
        float yt[ROWS_FWD];
#pragma unroll
        for (int r = 0; r < ROWS_FWD; ++r) {
            float acc = bias;
#pragma unroll
            for (int k4 = 0; k4 < NK_MAX / 4; ++k4) {
                if (4 * k4 < K) {       // one 16-byte broadcast load per 4 inputs (the kernel is LDS-issue bound otherwise)
                    const float4 xv = *reinterpret_cast<const float4*>(&xs[sub * ROWS_FWD + r][4 * k4]);
                    acc = fmaf(xv.x, w[4 * k4], fmaf(xv.y, w[4 * k4 + 1], fmaf(xv.z, w[4 * k4 + 2], fmaf(xv.w, w[4 * k4 + 3], acc))));
                }
            }
            if (relu) acc = fmaxf(acc, 0.f);
            yt[r] = acc;
            if (n_ok && m0 + r < M) Y[g * y_gs + (int64_t)(m0 + r) * ldy + n] = acc;
        }
        if (YT != nullptr && n_ok) {
            float* dst = YT + g * yt_gs + (int64_t)n * ldyt + m0;
            if (yt_vec && m0 + ROWS_FWD <= M) {
#pragma unroll
                for (int r4 = 0; r4 < ROWS_FWD / 4; ++r4)
                    reinterpret_cast<float4*>(dst)[r4] = make_float4(yt[4 * r4], yt[4 * r4 + 1], yt[4 * r4 + 2], yt[4 * r4 + 3]);
            } else {
#pragma unroll
                for (int r = 0; r < ROWS_FWD; ++r)
                    if (m0 + r < M) dst[r] = yt[r];
            }
        }
    }
}

// Partial sums over a 128-row chunk c (blockIdx.x):
//   out[g][c][ns][kw]  = sum_m Nar[g][m][ns] * Wide[g][m][kw]       (written at  ns*s_ns + kw*s_kw)
//   wide_sum[g][c][kw] = sum_m Wide[g][m][kw]                         (optional)
//   nar_sum[g][c][ns]  = sum_m Nar[g][m][ns]                          (optional)
// block = (64 threads x 4 wide columns each) x (4 row quarters).  A thread owns FOUR wide columns so that every 16-byte
// broadcast load of the narrow operand feeds 16 FMAs: with one column per thread the kernel was bound by the LDS issue
// rate (a 128-bit shared load occupies the load/store unit for 4 cycles whether or not it is a broadcast).
template <int NK>
__global__ void __launch_bounds__(256)
k_narrow_wgrad(const float* __restrict__ Wide, int64_t ldw, int64_t w_gs, const float* __restrict__ Nar, int64_t ldn,
               int64_t n_gs, float* __restrict__ out, int64_t s_ns, int64_t s_kw, int64_t o_gs, int64_t o_cs,
               float* __restrict__ wide_sum, int64_t ws_gs, int64_t ws_cs, float* __restrict__ nar_sum, int64_t ns_gs,
               int64_t ns_cs, int M, int KW, int NS) {
    orlk::pdl_enter();
    extern __shared__ float dsm[];
    constexpr int NLD = NK + 4;
    float (*ns_s)[NLD] = reinterpret_cast<float (*)[NLD]>(dsm);                     // [ROWS_WG][NK + 4], zero padded
    float* red = dsm + ROWS_WG * NLD;                                               // [3][4 * NK + 4][64]
    const int g = blockIdx.z, c = blockIdx.x;
    const int m0 = c * ROWS_WG;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 64 + tx;
    const int kw = blockIdx.y * 256 + 4 * tx;
    const int rows = min(ROWS_WG, M - m0);
    const float* Ng = Nar + g * n_gs;
    for (int i = tid; i < ROWS_WG * NK; i += 256) {
        const int r = i / NK, j = i % NK;
        ns_s[r][j] = (j < NS && r < rows) ? __ldg(Ng + (int64_t)(m0 + r) * ldn + j) : 0.f;
    }
    __syncthreads();
    float acc[4][NK];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NK; ++j) acc[q][j] = 0.f;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
    const int r_lo = ty * (ROWS_WG / 4), r_hi = min(rows, r_lo + ROWS_WG / 4);
    const bool vec = (ldw % 4) == 0 && (w_gs % 4) == 0 && aligned16(Wide) && kw + 3 < KW;
    if (kw < KW) {
        const float* wp = Wide + g * w_gs + (int64_t)m0 * ldw + kw;
        // The kernel is bound by the latency of the wide-operand loads (ncu: long-scoreboard stalls, 8 warps per SM), so
        // rows are fetched in batches of 8 with the NEXT batch already in flight while the current one is consumed.
        constexpr int RB = 8;
        float4 cur[RB], nxt[RB];
        auto fetch = [&](float4 (&dst)[RB], int r0) {
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const int r = r0 + i;
                if (r < r_hi) {
                    if (vec) dst[i] = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)r * ldw));
                    else {
                        const float* q = wp + (int64_t)r * ldw;
                        dst[i] = make_float4(__ldg(q), kw + 1 < KW ? __ldg(q + 1) : 0.f, kw + 2 < KW ? __ldg(q + 2) : 0.f,
                                             kw + 3 < KW ? __ldg(q + 3) : 0.f);
                    }
                } else dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        fetch(cur, r_lo);
        for (int r0 = r_lo; r0 < r_hi; r0 += RB) {
            fetch(nxt, r0 + RB);
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const int r = min(r0 + i, ROWS_WG - 1);     // rows past r_hi carry zeros in cur[]
                const float x[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) cs[q] += x[q];
#pragma unroll
                for (int j4 = 0; j4 < NK / 4; ++j4) {
                    const float4 nv = *reinterpret_cast<const float4*>(&ns_s[r][4 * j4]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[q][4 * j4] = fmaf(x[q], nv.x, acc[q][4 * j4]);
                        acc[q][4 * j4 + 1] = fmaf(x[q], nv.y, acc[q][4 * j4 + 1]);
                        acc[q][4 * j4 + 2] = fmaf(x[q], nv.z, acc[q][4 * j4 + 2]);
                        acc[q][4 * j4 + 3] = fmaf(x[q], nv.w, acc[q][4 * j4 + 3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < RB; ++i) cur[i] = nxt[i];
        }
    }
    // combine the four row quarters in a fixed order (deterministic): quarters 1..3 park their sums in smem
    constexpr int PER = 4 * NK + 4;
    if (ty > 0) {
        float* r = red + (size_t)(ty - 1) * PER * 64 + tx;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < NK; ++j) r[(q * NK + j) * 64] = acc[q][j];
            r[(4 * NK + q) * 64] = cs[q];
        }
    }
    __syncthreads();
    if (ty == 0 && kw < KW) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (kw + q >= KW) break;
            float* o = out + g * o_gs + (int64_t)c * o_cs + (int64_t)(kw + q) * s_kw;
#pragma unroll
            for (int j = 0; j < NK; ++j)
                if (j < NS) {
                    float v = acc[q][j];
                    for (int t = 0; t < 3; ++t) v += red[(size_t)t * PER * 64 + (q * NK + j) * 64 + tx];
                    o[(int64_t)j * s_ns] = v;
                }
            if (wide_sum != nullptr) {
                float v = cs[q];
                for (int t = 0; t < 3; ++t) v += red[(size_t)t * PER * 64 + (4 * NK + q) * 64 + tx];
                wide_sum[g * ws_gs + (int64_t)c * ws_cs + kw + q] = v;
            }
        }
    }
    if (nar_sum != nullptr && blockIdx.y == 0 && ty == 1 && tx < NS) {       // (a warp that is idle after parking its sums)
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += ns_s[r][tx];
        nar_sum[g * ns_gs + (int64_t)c * ns_cs + tx] = s;
    }
}

template <int NK>
constexpr size_t wgrad_smem() { return sizeof(float) * (ROWS_WG * (NK + 4) + 3 * (4 * NK + 4) * 64); }

}  // namespace

extern "C" {

int orlk_narrow_fwd(const float* X, int64_t ldx, int64_t x_gs, const float* W, int64_t ldw, int64_t w_gs, const float* b,
                    int64_t b_gs, float* Y, int64_t ldy, int64_t y_gs, float* YT, int64_t ldyt, int64_t yt_gs, int M, int N,
                    int K, int G, int relu, void* stream) {
    ORLK_REQUIRE(K >= 1 && K <= NK_MAX, "K must be in [1,32]");
    ORLK_REQUIRE(M > 0 && N > 0 && G > 0, "sizes");
    dim3 grid((M + ROWS_FWD * SUB_FWD - 1) / (ROWS_FWD * SUB_FWD), (N + 255) / 256, G);
    orlk::launch(k_narrow_fwd, grid, 256, 0, (cudaStream_t)stream, X, ldx, x_gs, W, ldw, w_gs, b, b_gs, Y, ldy, y_gs, YT, ldyt, yt_gs, M, N,
                                                        K, relu);
    return check_launch("k_narrow_fwd");
}

int orlk_narrow_wgrad_chunks(int M) { return (M + ROWS_WG - 1) / ROWS_WG; }

// opt in to the dynamic shared memory of the weight-gradient kernels; once, outside stream capture
int orlk_narrow_init(void) {
    int rc = check(cudaFuncSetAttribute(k_narrow_wgrad<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem<4>()), "smem attr narrow");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(k_narrow_wgrad<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem<8>()), "smem attr narrow");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(k_narrow_wgrad<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem<16>()), "smem attr narrow");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(k_narrow_wgrad<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem<24>()), "smem attr narrow");
    if (rc) return rc;
    return check(cudaFuncSetAttribute(k_narrow_wgrad<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem<32>()), "smem attr narrow");
}

int orlk_narrow_wgrad(const float* Wide, int64_t ldw, int64_t w_gs, const float* Nar, int64_t ldn, int64_t n_gs, float* out,
                      int64_t s_ns, int64_t s_kw, int64_t o_gs, int64_t o_cs, float* wide_sum, int64_t ws_gs, int64_t ws_cs,
                      float* nar_sum, int64_t ns_gs, int64_t ns_cs, int M, int KW, int NS, int G, void* stream) {
    ORLK_REQUIRE(NS >= 1 && NS <= NK_MAX, "NS must be in [1,32]");
    ORLK_REQUIRE(M > 0 && KW > 0 && G > 0, "sizes");
    dim3 grid((M + ROWS_WG - 1) / ROWS_WG, (KW + 255) / 256, G);
    cudaStream_t s = (cudaStream_t)stream;
#define ORLK_WG_LAUNCH(NKV)                                                                                               \
    orlk::launch(k_narrow_wgrad<NKV>, grid, dim3(64, 4), wgrad_smem<NKV>(), s, Wide, ldw, w_gs, Nar, ldn, n_gs, out, s_ns, s_kw, \
                 o_gs, o_cs, wide_sum, ws_gs, ws_cs, nar_sum, ns_gs, ns_cs, M, KW, NS)
    if (NS <= 4) ORLK_WG_LAUNCH(4);
    else if (NS <= 8) ORLK_WG_LAUNCH(8);
    else if (NS <= 16) ORLK_WG_LAUNCH(16);
    else if (NS <= 24) ORLK_WG_LAUNCH(24);
    else ORLK_WG_LAUNCH(32);
#undef ORLK_WG_LAUNCH
    return check_launch("k_narrow_wgrad");
}

}  // extern "C"
