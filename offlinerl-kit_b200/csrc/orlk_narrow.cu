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
constexpr int SUB_FWD = 2;     // sub-tiles per block (32 rows per block: the layer's weights are loaded once per block)
constexpr int XLD = NK_MAX + 4; // shared-memory row stride: 16-byte aligned rows for float4 broadcast loads
constexpr int ROWS_WG = 128;   // rows per block (= per partial slot) in the weight-gradient kernel

// Y[g][m][n] = act(b[g][n] + sum_k X[g][m][k] * W[g][n*ldw + k]),   K <= 32;  optional YT[g][n][m]
__global__ void __launch_bounds__(256, 2)
k_narrow_fwd(const float* __restrict__ X, int64_t ldx, int64_t x_gs, const float* __restrict__ W, int64_t ldw, int64_t w_gs,
             const float* __restrict__ b, int64_t b_gs, float* __restrict__ Y, int64_t ldy, int64_t y_gs,
             float* __restrict__ YT, int64_t ldyt, int64_t yt_gs, int M, int N, int K, int relu) {
    orlk::pdl_enter();
    __shared__ __align__(16) float xs[ROWS_FWD * SUB_FWD][XLD];
    const int g = blockIdx.z;
    const int mb = blockIdx.x * ROWS_FWD * SUB_FWD;
    const int n = blockIdx.y * 256 + threadIdx.x;
    const float* Xg = X + g * x_gs;
    for (int i = threadIdx.x; i < ROWS_FWD * SUB_FWD * NK_MAX; i += 256) {
        const int r = i / NK_MAX, k = i % NK_MAX;          // zero-padded to 32 columns: the k loop runs in float4 steps
        xs[r][k] = (k < K && mb + r < M) ? __ldg(Xg + (int64_t)(mb + r) * ldx + k) : 0.f;
    }
    float w[NK_MAX];
    const bool n_ok = n < N;
#pragma unroll
    for (int k = 0; k < NK_MAX; ++k) w[k] = (n_ok && k < K) ? __ldg(W + g * w_gs + (int64_t)n * ldw + k) : 0.f;
    const float bias = (n_ok && b != nullptr) ? __ldg(b + g * b_gs + n) : 0.f;
    __syncthreads();
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
__global__ void __launch_bounds__(1024)
k_narrow_wgrad(const float* __restrict__ Wide, int64_t ldw, int64_t w_gs, const float* __restrict__ Nar, int64_t ldn,
               int64_t n_gs, float* __restrict__ out, int64_t s_ns, int64_t s_kw, int64_t o_gs, int64_t o_cs,
               float* __restrict__ wide_sum, int64_t ws_gs, int64_t ws_cs, float* __restrict__ nar_sum, int64_t ns_gs,
               int64_t ns_cs, int M, int KW, int NS) {
    orlk::pdl_enter();
    // block = (256 wide columns) x (4 row quarters); the quarters are combined through shared memory in one shot
    extern __shared__ float dsm[];
    float (*ns_s)[XLD] = reinterpret_cast<float (*)[XLD]>(dsm);                             // [ROWS_WG][36], zero padded
    float* red = dsm + ROWS_WG * XLD;                                               // [3][NS+1][256]
    const int g = blockIdx.z, c = blockIdx.x;
    const int m0 = c * ROWS_WG;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 256 + tx;
    const int kw = blockIdx.y * 256 + tx;
    const int rows = min(ROWS_WG, M - m0);
    const float* Ng = Nar + g * n_gs;
    for (int i = tid; i < ROWS_WG * NK_MAX; i += 1024) {
        const int r = i / NK_MAX, j = i % NK_MAX;
        ns_s[r][j] = (j < NS && r < rows) ? __ldg(Ng + (int64_t)(m0 + r) * ldn + j) : 0.f;
    }
    __syncthreads();
    float acc[NK_MAX];
#pragma unroll
    for (int j = 0; j < NK_MAX; ++j) acc[j] = 0.f;
    float cs = 0.f;
    const int r_lo = ty * (ROWS_WG / 4), r_hi = min(rows, r_lo + ROWS_WG / 4);
    if (kw < KW) {
        const float* wp = Wide + g * w_gs + (int64_t)m0 * ldw + kw;
#pragma unroll 8
        for (int r = r_lo; r < r_hi; ++r) {
            const float x = __ldg(wp + (int64_t)r * ldw);
            cs += x;
#pragma unroll
            for (int j4 = 0; j4 < NK_MAX / 4; ++j4) {
                if (4 * j4 < NS) {
                    const float4 nv = *reinterpret_cast<const float4*>(&ns_s[r][4 * j4]);
                    acc[4 * j4] = fmaf(x, nv.x, acc[4 * j4]);
                    acc[4 * j4 + 1] = fmaf(x, nv.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(x, nv.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(x, nv.w, acc[4 * j4 + 3]);
                }
            }
        }
    }
    // combine the four row quarters in a fixed order (deterministic): quarters 1..3 park their sums in smem
    if (ty > 0) {
        float* r = red + (size_t)(ty - 1) * (NS + 1) * 256 + tx;
#pragma unroll
        for (int j = 0; j < NK_MAX; ++j)
            if (j < NS) r[j * 256] = acc[j];
        r[NS * 256] = cs;
    }
    __syncthreads();
    if (ty == 0 && kw < KW) {
        float* o = out + g * o_gs + (int64_t)c * o_cs + (int64_t)kw * s_kw;
#pragma unroll
        for (int j = 0; j < NK_MAX; ++j)
            if (j < NS) {
                float v = acc[j];
                for (int q = 0; q < 3; ++q) v += red[(size_t)q * (NS + 1) * 256 + j * 256 + tx];
                o[(int64_t)j * s_ns] = v;
            }
        if (wide_sum != nullptr) {
            float v = cs;
            for (int q = 0; q < 3; ++q) v += red[(size_t)q * (NS + 1) * 256 + NS * 256 + tx];
            wide_sum[g * ws_gs + (int64_t)c * ws_cs + kw] = v;
        }
    }
    if (nar_sum != nullptr && blockIdx.y == 0 && ty == 1 && tx < NS) {       // (a warp that is idle after parking its sums)
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += ns_s[r][tx];
        nar_sum[g * ns_gs + (int64_t)c * ns_cs + tx] = s;
    }
}

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

// opt in to the (up to 116 KB) dynamic shared memory of the weight-gradient kernel; once, outside stream capture
int orlk_narrow_init(void) {
    return check(cudaFuncSetAttribute(k_narrow_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(float) * (ROWS_WG * XLD + 3 * (NK_MAX + 1) * 256))), "smem attr narrow");
}

int orlk_narrow_wgrad(const float* Wide, int64_t ldw, int64_t w_gs, const float* Nar, int64_t ldn, int64_t n_gs, float* out,
                      int64_t s_ns, int64_t s_kw, int64_t o_gs, int64_t o_cs, float* wide_sum, int64_t ws_gs, int64_t ws_cs,
                      float* nar_sum, int64_t ns_gs, int64_t ns_cs, int M, int KW, int NS, int G, void* stream) {
    ORLK_REQUIRE(NS >= 1 && NS <= NK_MAX, "NS must be in [1,32]");
    ORLK_REQUIRE(M > 0 && KW > 0 && G > 0, "sizes");
    dim3 grid((M + ROWS_WG - 1) / ROWS_WG, (KW + 255) / 256, G);
    const size_t smem = sizeof(float) * (ROWS_WG * XLD + 3 * (NS + 1) * 256);
    orlk::launch(k_narrow_wgrad, grid, dim3(256, 4), smem, (cudaStream_t)stream, Wide, ldw, w_gs, Nar, ldn, n_gs, out, s_ns, s_kw, o_gs, o_cs, wide_sum,
                                                          ws_gs, ws_cs, nar_sum, ns_gs, ns_cs, M, KW, NS);
    return check_launch("k_narrow_wgrad");
}

}  // extern "C"
