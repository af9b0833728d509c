// Narrow layers, row assembly, the tanh-Gaussian policy head, loss epilogues and the fused optimiser.
// Everything here is latency-bound elementwise / reduction work on <= 8K rows: warp-shuffle reductions,
// one CTA for the scalar losses, no atomics (results are run-to-run deterministic).
#include <math.h>
#include <stdlib.h>
#include "orlk_common.cuh"
using namespace orlk;

namespace {

constexpr int MAX_NS = 16;

// ------------------------------------------------------------------------------------------ skinny layers
template <int NST>
__global__ void __launch_bounds__(256)
k_skinny_fwd(const float* __restrict__ X, int64_t ldx, int64_t x_gs, const float* __restrict__ W, int64_t ldw,
             int64_t w_sk, int64_t w_gs, const float* __restrict__ b, int64_t b_gs, float* __restrict__ Y, int64_t ldy,
             int64_t y_gs, int M, int K, int NS, int vec) {
    orlk::pdl_enter();
    const int g = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    const float* x = X + g * x_gs + (int64_t)m * ldx;
    const float* w = W + g * w_gs;
    float acc[NST];
#pragma unroll
    for (int n = 0; n < NST; ++n) acc[n] = 0.f;
    if (vec) {      // K % 4 == 0 and 16-byte aligned rows: 128-bit loads, all of a row's loads in flight at once
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const int K4 = K >> 2;
#pragma unroll 2
        for (int k = lane; k < K4; k += 32) {
            const float4 xv = x4[k];
#pragma unroll
            for (int n = 0; n < NST; ++n)
                if (n < NS) {
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (int64_t)n * ldw) + k);
                    acc[n] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[n]))));
                }
        }
    } else {
        for (int k = lane; k < K; k += 32) {
            const float xv = x[k];
#pragma unroll
            for (int n = 0; n < NST; ++n)
                if (n < NS) acc[n] = fmaf(xv, __ldg(w + (int64_t)n * ldw + (int64_t)k * w_sk), acc[n]);
        }
    }
#pragma unroll
    for (int n = 0; n < NST; ++n)
        if (n < NS) acc[n] = warp_sum(acc[n]);
    if (lane == 0) {
        float* y = Y + g * y_gs + (int64_t)m * ldy;
#pragma unroll
        for (int n = 0; n < NST; ++n)
            if (n < NS) y[n] = acc[n] + (b ? b[g * b_gs + n] : 0.f);
    }
}

// 32 x 32 (m x k) tile per block of 32 x 8 threads; the optional transposed copy goes through shared memory so
// that both stores are coalesced.
__global__ void k_skinny_dgrad(const float* __restrict__ dY, int64_t ldy, int64_t y_gs, const float* __restrict__ W,
                               int64_t ldw, int64_t w_gs, const float* __restrict__ mask, int64_t ldm, int64_t m_gs,
                               float* __restrict__ dX, int64_t ldx, int64_t x_gs, float* __restrict__ dXT, int64_t ldxt,
                               int64_t xt_gs, int M, int K, int NS) {
    orlk::pdl_enter();
    __shared__ float tile[32][33];
    __shared__ float dys[32][MAX_NS];
    const int g = blockIdx.z;
    const int m0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int i = ty * 32 + tx; i < 32 * NS; i += 256) {
        const int r = i / NS, n = i % NS;
        dys[r][n] = (m0 + r < M) ? dY[g * y_gs + (int64_t)(m0 + r) * ldy + n] : 0.f;
    }
    __syncthreads();
    for (int k0 = blockIdx.x * 32; k0 < K; k0 += gridDim.x * 32) {
        const int k = k0 + tx;
        float wk[MAX_NS];
#pragma unroll
        for (int n = 0; n < MAX_NS; ++n) wk[n] = (n < NS && k < K) ? __ldg(W + g * w_gs + (int64_t)n * ldw + k) : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = ty + 8 * r;
            const int m = m0 + row;
            float s = 0.f;
            if (m < M && k < K) {
#pragma unroll
                for (int n = 0; n < MAX_NS; ++n)
                    if (n < NS) s = fmaf(dys[row][n], wk[n], s);
                if (mask != nullptr && !(mask[g * m_gs + (int64_t)m * ldm + k] > 0.f)) s = 0.f;
                dX[g * x_gs + (int64_t)m * ldx + k] = s;
            }
            tile[row][tx] = s;
        }
        if (dXT != nullptr) {
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int kk = k0 + ty + 8 * r, m = m0 + tx;
                if (kk < K && m < M) dXT[g * xt_gs + (int64_t)kk * ldxt + m] = tile[tx][ty + 8 * r];
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------ row assembly
__global__ void k_concat_rows(const OrlkConcatSeg* __restrict__ segs, int n_segs, int total_rows) {
    orlk::pdl_enter();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    int s = 0;
    while (s + 1 < n_segs && segs[s + 1].row_start <= row) ++s;
    const OrlkConcatSeg sg = segs[s];
    const int m = row - sg.row_start;
    float* dst = sg.dst + (int64_t)m * sg.ld_dst;
    const float* a = sg.src1 + (int64_t)(m / sg.rep1) * sg.ld1;
    const float* b = sg.src2 + (int64_t)m * sg.ld2;
    for (int j = lane; j < sg.w1 + sg.w2; j += 32) dst[j] = j < sg.w1 ? a[j] : b[j - sg.w1];
}

// ------------------------------------------------------------------------------------------ member-sharded exchange
// After an equal-block all-gather over `world` ranks each rank holds [world][block_stride] floats, of which rank r's block
// carries counts[r] * per_member valid floats (ensembles split unevenly over the ranks: 10 critics on 4 GPUs = 3/3/2/2).
// Writes them densely, in rank order, to dst.  (edac.py:96-149 / ensemble_dynamics.py:178-223 with members sharded.)
struct CompactArgs { int counts[8]; };
__global__ void k_compact_blocks(const float* __restrict__ src, int64_t block_stride, float* __restrict__ dst, int world,
                                 int per_member, const __grid_constant__ CompactArgs C) {
    orlk::pdl_enter();
    const int r = blockIdx.y;
    int before = 0;
    for (int q = 0; q < r; ++q) before += C.counts[q];
    const int64_t n = (int64_t)C.counts[r] * per_member;
    const float* s = src + (int64_t)r * block_stride;
    float* d = dst + (int64_t)before * per_member;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = s[i];
}

// ------------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

__global__ void k_philox_fill(float* __restrict__ out, int64_t n_normal, int64_t n_uniform, float lo, float hi,
                              uint64_t seed, const unsigned long long* __restrict__ counter, const int* __restrict__ enable) {
    orlk::pdl_enter();
    if (enable != nullptr && *enable == 0) return;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one Philox call = 4 outputs
    const int64_t n = n_normal + n_uniform;
    if (q * 4 >= n) return;
    const unsigned long long ctr = counter ? *counter : 0ull;
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u0 = u01(c[0]), u1 = u01(c[1]), u2 = u01(c[2]), u3 = u01(c[3]);
    // Box-Muller on (u0,u1) and (u2,u3)
    const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.f * u1, &s0, &c0);
    sincospif(2.f * u3, &s1, &c1);
    const float nrm[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
    const float uni[4] = {u0, u1, u2, u3};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t i = q * 4 + j;
        if (i < n_normal) out[i] = nrm[j];
        else if (i < n) out[i] = lo + (hi - lo) * uni[j];
    }
}

// ------------------------------------------------------------------------------------------ tanh-Gaussian head
constexpr int MAX_A = 32;
constexpr float LOG_SIG_MIN = -5.f, LOG_SIG_MAX = 2.f;
constexpr float HALF_LOG_2PI = 0.91893853320467274178f;

__global__ void k_tanh_gauss_sample(const float* __restrict__ head, int64_t ld_head, int head_row_off, int rep,
                                    const float* __restrict__ eps, int M, int A, float* __restrict__ act, int64_t ld_act,
                                    float* __restrict__ logp, const float* __restrict__ obs, int64_t ld_obs, int obs_dim,
                                    float* __restrict__ xout, int64_t ld_x) {
    orlk::pdl_enter();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* h = head + (int64_t)(head_row_off + m / rep) * ld_head;
    float lp = 0.f, corr = 0.f;
    for (int i = 0; i < A; ++i) {
        const float mu = h[i];
        const float ls = fminf(fmaxf(h[A + i], LOG_SIG_MIN), LOG_SIG_MAX);
        const float sigma = expf(ls);
        const float u = eps ? fmaf(sigma, eps[(int64_t)m * A + i], mu) : mu;
        const float a = tanhf(u);
        const float d = u - mu;
        lp += -(d * d) / (2.f * sigma * sigma) - logf(sigma) - HALF_LOG_2PI;
        corr += logf((1.f - a * a) + 1e-6f);
        act[(int64_t)m * ld_act + i] = a;
    }
    if (logp) logp[m] = lp - corr;
    if (xout) {
        const float* o = obs + (int64_t)(m / rep) * ld_obs;
        for (int j = 0; j < obs_dim; ++j) xout[(int64_t)m * ld_x + j] = o[j];
    }
}

__global__ void k_tanh_gauss_bwd(const float* __restrict__ head, int64_t ld_head, const float* __restrict__ eps,
                                 const float* __restrict__ act, int64_t ld_act, const float* __restrict__ dA, int n_da,
                                 int64_t da_gs, int64_t ld_da, const float* __restrict__ glp, int M, int A,
                                 float* __restrict__ dhead, int64_t ld_dhead) {
    orlk::pdl_enter();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* h = head + (int64_t)m * ld_head;
    const float g = glp ? glp[m] : 0.f;
    for (int i = 0; i < A; ++i) {
        const float raw = h[A + i];
        const float ls = fminf(fmaxf(raw, LOG_SIG_MIN), LOG_SIG_MAX);
        const float sigma = expf(ls);
        const float e = eps[(int64_t)m * A + i];
        const float a = act[(int64_t)m * ld_act + i];
        const float om = 1.f - a * a;
        float da = 0.f;
        for (int j = 0; j < n_da; ++j) da += dA[j * da_gs + (int64_t)m * ld_da + i];    // sum over the critics / members
        const float t = 2.f * a * om / (om + 1e-6f);   // d logp / d u  (tanh correction only)
        const float du = da * om + g * t;              // through a = tanh(u) and through logp
        const float dmu = du;
        float draw = du * sigma * e - g;               // u = mu + sigma*eps ; -log sigma term of logp
        if (!(raw >= LOG_SIG_MIN && raw <= LOG_SIG_MAX)) draw = 0.f;
        dhead[(int64_t)m * ld_dhead + i] = dmu;
        dhead[(int64_t)m * ld_dhead + A + i] = draw;
    }
}

// Actor head + reparameterised sampling in one launch, one 4-warp block per head row m:
//   head[m] = X[m] . Wh^T + bh   (mu | log-std raw, 2A outputs), then for every use u whose row range contains m and every
//   repeat r:  a = tanh(mu + sigma * eps[o]),  logp[o],  X_out[o] = [obs[j] | a]   with j = m - r0, o = j * rep + r
// (exactly k_skinny_fwd followed by k_tanh_gauss_sample per use; CQL's critic phase has three uses of one head pass).
// The launch holds only a few hundred rows, so a warp has an SM scheduler to itself and runs at its dependent-issue
// latency (~8 cycles per instruction, measured): what counts is the LENGTH of a warp's instruction stream.  Hence four
// warps per row - each takes every fourth head output, then every fourth (use, pass of four repeats) - instead of one
// warp doing a row's ~1400 instructions alone (10 us -> see profiles/).  Arithmetic per element as before.
struct SampleUses {
    OrlkSampleUse u[4];
    int n;
};

constexpr int HS_WARPS = 4;

__global__ void __launch_bounds__(32 * HS_WARPS)
k_head_sample(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int64_t ldw, const float* __restrict__ b,
              float* __restrict__ head, int M, int K, int A, const __grid_constant__ SampleUses U) {
    orlk::pdl_enter();
    __shared__ float head_s[16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int m = blockIdx.x;
    const int NS = 2 * A;
    // ---- head outputs n = w, w + 4, ... of this row: lanes split k, then a warp sum
    const float4* x4 = reinterpret_cast<const float4*>(X + (int64_t)m * ldx);
    const int K4 = K >> 2;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < K4; k += 32) {
        const float4 xv = x4[k];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = w + HS_WARPS * i;
            if (n < NS) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * ldw) + k);
                acc[i] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[i]))));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = w + HS_WARPS * i;
        if (n < NS) {
            const float v = warp_sum(acc[i]) + (b ? __ldg(b + n) : 0.f);
            if (lane == 0) {
                head_s[n] = v;
                head[(int64_t)m * NS + n] = v;
            }
        }
    }
    __syncthreads();
    // ---- samplers: four repeats per pass, lane = 8 * (repeat within the pass) + action dimension (A <= 8);
    //      pass t (counted over all uses that contain this row) belongs to warp t % 4
    const int gi = lane & 7, gr = lane >> 3;
    const float mu_i = gi < A ? head_s[gi] : 0.f, raw_i = gi < A ? head_s[A + gi] : 0.f;
    const float ls = fminf(fmaxf(raw_i, LOG_SIG_MIN), LOG_SIG_MAX);
    const float sigma = expf(ls);
    const float lsig = logf(sigma);
    int t = 0;
    for (int q = 0; q < U.n; ++q) {
        const OrlkSampleUse& u = U.u[q];
        if (m < u.r0 || m >= u.r1) continue;            // block-uniform
        const int j = m - u.r0;
        for (int r0 = 0; r0 < u.rep; r0 += 4, ++t) {
            if ((t & (HS_WARPS - 1)) != w) continue;    // warp-uniform
            const int r = r0 + gr;
            const bool on = r < u.rep && gi < A;
            const int64_t o = (int64_t)j * u.rep + r;
            float tl = 0.f;
            if (on) {
                const float uu = u.eps ? fmaf(sigma, u.eps[o * A + gi], mu_i) : mu_i;
                const float a = tanhf(uu);
                const float d = uu - mu_i;
                tl = -(d * d) / (2.f * sigma * sigma) - lsig - HALF_LOG_2PI - logf((1.f - a * a) + 1e-6f);
                u.act[o * u.ld_act + gi] = a;
            }
            tl += __shfl_xor_sync(0xffffffffu, tl, 1);
            tl += __shfl_xor_sync(0xffffffffu, tl, 2);
            tl += __shfl_xor_sync(0xffffffffu, tl, 4);
            if (gi == 0 && r < u.rep && u.logp) u.logp[o] = tl;
            if (u.xout) {
                const float* ob = u.obs + (int64_t)j * u.ld_obs;
                for (int c = lane; c < u.obs_dim; c += 32) {
                    const float ov = ob[c];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr)
                        if (r0 + rr < u.rep) u.xout[((int64_t)j * u.rep + r0 + rr) * u.ld_x + c] = ov;
                }
            }
        }
    }
}

// Entry of the actor's backward pass in one launch (policy improvement step of SAC / CQL / MOPO), one 4-warp block per row m:
//   dL/da[m]      = sum_c dZ0_c[m][:] . W0_c[:, O:O+A]            (gradient of the critics w.r.t. the sampled action)
//   dhead[m]      = tanh-Gaussian backward                          (same maths as k_tanh_gauss_bwd)
//   dZlast[m][n]  = (dhead[m][:] . Wh[:, n]) * (Hlast[m][n] > 0)   (through the actor's head into its last hidden layer)
// i.e. the skinny d/da product, the sampler backward and the head dgrad, which otherwise are three 3-4 us launches.
// As in k_head_sample the launch is a few hundred rows, so the length of a warp's instruction stream is what costs:
// the k range of the d/da product and the columns of the head dgrad are split over the block's four warps (a row used to
// be one warp's ~1400 instructions behind a 24 KB shared-memory staging of the weights; they are L1 / L2 hits here).
constexpr int BE_WARPS = 4;

template <int AM>       // AM >= A: compile-time bound of the action dimension (8 or 32)
__global__ void __launch_bounds__(32 * BE_WARPS)
k_actor_bwd_entry(const float* __restrict__ dZ0, int64_t dz_gs, int Kc, int n_c, const float* __restrict__ W0, int64_t w0_gs,
                  int ld_w0, int col0, const float* __restrict__ head, const float* __restrict__ eps,
                  const float* __restrict__ act, int64_t ld_act, const float* __restrict__ glp, int M, int A,
                  float* __restrict__ dhead, const float* __restrict__ Wh, int Ka, const float* __restrict__ Hlast,
                  float* __restrict__ dZlast) {
    orlk::pdl_enter();
    __shared__ float da_s[BE_WARPS][AM];
    __shared__ float dh_s[2 * AM];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int m = blockIdx.x;
    // inputs of the later stages, requested up front
    float pre_g = 0.f, pre_raw = 0.f, pre_e = 0.f, pre_a = 0.f;
    if (w == 0 && lane < A) {
        pre_g = glp ? glp[m] : 0.f;
        pre_raw = head[(int64_t)m * 2 * A + A + lane];
        pre_e = eps[(int64_t)m * A + lane];
        pre_a = act[(int64_t)m * ld_act + lane];
    }
    // ---- d/da: the block's 128 threads split k, then warp sums and a fixed-order sum over the warps
    float da[AM];
#pragma unroll
    for (int a = 0; a < AM; ++a) da[a] = 0.f;
    for (int c = 0; c < n_c; ++c) {
        const float* z = dZ0 + c * dz_gs + (int64_t)m * Kc;
        const float* wc = W0 + c * w0_gs + col0;
        for (int k = threadIdx.x; k < Kc; k += 32 * BE_WARPS) {
            const float zv = __ldg(z + k);
            const float* wr = wc + (int64_t)k * ld_w0;
#pragma unroll
            for (int a = 0; a < AM; ++a)
                if (a < A) da[a] = fmaf(zv, __ldg(wr + a), da[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < AM; ++a)
        if (a < A) {
            const float v = warp_sum(da[a]);
            if (lane == 0) da_s[w][a] = v;
        }
    __syncthreads();
    // ---- sampler backward: lane i < A of warp 0 owns action dimension i
    if (w == 0 && lane < A) {
        const int i = lane;
        float dai = 0.f;
#pragma unroll
        for (int ww = 0; ww < BE_WARPS; ++ww) dai += da_s[ww][i];
        const float g = pre_g;
        const float raw = pre_raw;
        const float ls = fminf(fmaxf(raw, LOG_SIG_MIN), LOG_SIG_MAX);
        const float sigma = expf(ls);
        const float e = pre_e;
        const float a = pre_a;
        const float om = 1.f - a * a;
        const float t = 2.f * a * om / (om + 1e-6f);
        const float du = dai * om + g * t;
        float draw = du * sigma * e - g;
        if (!(raw >= LOG_SIG_MIN && raw <= LOG_SIG_MAX)) draw = 0.f;
        dhead[(int64_t)m * 2 * A + i] = du;
        dhead[(int64_t)m * 2 * A + A + i] = draw;
        dh_s[i] = du;
        dh_s[A + i] = draw;
    }
    __syncthreads();
    // ---- head dgrad: one output column per thread and pass
    for (int n = threadIdx.x; n < Ka; n += 32 * BE_WARPS) {
        const float hv = __ldg(Hlast + (int64_t)m * Ka + n);
        float acc = 0.f;
        for (int j = 0; j < A; ++j)
            acc = fmaf(dh_s[j], __ldg(Wh + (int64_t)j * Ka + n), fmaf(dh_s[A + j], __ldg(Wh + (int64_t)(A + j) * Ka + n), acc));
        dZlast[(int64_t)m * Ka + n] = hv > 0.f ? acc : 0.f;
    }
}

// ------------------------------------------------------------------------------------------ scalar Adam
__device__ float scalar_adam(float p, float g, float* mv, const OrlkAdamGroup& grp) {
    const float step_size = (float)((double)grp.lr / grp.bc1);      // bias corrections: kept by the host / orlk_step_end
    const float bc2_sqrt = grp.bc2_sqrt;
    float m = mv[0], v = mv[1];
    m = m + (g - m) * (1.f - grp.beta1);
    v = v * grp.beta2 + (1.f - grp.beta2) * g * g;
    mv[0] = m; mv[1] = v;
    const float denom = sqrtf(v) / bc2_sqrt + grp.eps;
    return p - step_size * (m / denom);
}

// ------------------------------------------------------------------------------------------ actor loss
__global__ void k_sac_actor_loss(const float* __restrict__ q, int64_t q_es, int E, const float* __restrict__ logp, int B,
                                 float* __restrict__ scalars, int auto_alpha, int clamp01, float target_entropy,
                                 const OrlkAdamGroup* __restrict__ groups, int alpha_group, float* __restrict__ alpha_mv,
                                 float* __restrict__ dq, int64_t dq_es, float* __restrict__ glp, float* __restrict__ out) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const float alpha = scalars[ORLK_SC_ALPHA];
    const float invB = 1.f / (float)B;
    float lsum = 0.f, lpsum = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float mn = q[b];
        int arg = 0;
        for (int e = 1; e < E; ++e) {
            const float v = q[e * q_es + b];
            if (v < mn) { mn = v; arg = e; }
        }
        const float lp = logp[b];
        lsum += alpha * lp - mn;
        lpsum += lp;
        if (E == 2) {   // torch.min(a, b): ties split the gradient evenly
            const float q0 = q[b], q1 = q[q_es + b];
            dq[b] = q0 < q1 ? -invB : (q0 == q1 ? -0.5f * invB : 0.f);
            dq[dq_es + b] = q1 < q0 ? -invB : (q0 == q1 ? -0.5f * invB : 0.f);
        } else {        // torch.min(dim=0): gradient to the returned (first minimal) index
            for (int e = 0; e < E; ++e) dq[e * dq_es + b] = (e == arg) ? -invB : 0.f;
        }
        glp[b] = alpha * invB;
    }
    lsum = block_sum(lsum, red);
    lpsum = block_sum(lpsum, red);
    if (threadIdx.x == 0) {
        out[0] = lsum * invB;
        float new_alpha = alpha, aloss = 0.f;
        if (auto_alpha) {
            const float la = scalars[ORLK_SC_LOG_ALPHA];
            const float mean_lp = lpsum * invB + target_entropy;
            aloss = -(la * mean_lp);
            const float la_new = scalar_adam(la, -mean_lp, alpha_mv, groups[alpha_group]);
            scalars[ORLK_SC_LOG_ALPHA] = la_new;
            new_alpha = expf(la_new);
            if (clamp01) new_alpha = fminf(fmaxf(new_alpha, 0.f), 1.f);
            scalars[ORLK_SC_ALPHA] = new_alpha;
        }
        out[1] = aloss;
        out[2] = new_alpha;
    }
}

// ------------------------------------------------------------------------------------------ twin scalar heads, policy step
// The critics' scalar heads, the gradient of the policy-improvement loss w.r.t. their outputs and that gradient pulled
// back through the heads, in ONE launch (k_skinny_fwd + the dq / glp part of k_sac_actor_loss + k_skinny_dgrad):
//   q_c[m] = H_c[m] . w_c + b_c;  dq_c[m] = d/dq_c mean_b(alpha logp - min(q_0, q_1)) = -1/B for the smaller one (ties: half each);
//   glp[m] = alpha / B;  dZ_c[m][k] = dq_c[m] * w_c[k] * (H_c[m][k] > 0)
// None of these needs a batch reduction; the loss VALUE and the temperature step do, and run beside the backward pass
// (k_sac_actor_loss on q, off the critical path).  One 4-warp block per row (sac.py:111-120, cql.py:93-100).
constexpr int TH_WARPS = 4;

__global__ void __launch_bounds__(32 * TH_WARPS)
k_twin_head_actor(const float* __restrict__ H, int64_t h_gs, const float* __restrict__ Wh, int64_t w_gs, const float* __restrict__ bh,
                  int64_t b_gs, const float* __restrict__ scalars, int B, int K, float* __restrict__ q, int64_t q_gs,
                  float* __restrict__ dq, int64_t dq_gs, float* __restrict__ glp, float* __restrict__ dZ, int64_t dz_gs) {
    orlk::pdl_enter();
    __shared__ float part[2][TH_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int m = blockIdx.x;
    const float* h0 = H + (int64_t)m * K;
    const float* h1 = H + h_gs + (int64_t)m * K;
    float a0 = 0.f, a1 = 0.f;
    for (int k = threadIdx.x; k < K; k += 32 * TH_WARPS) {
        a0 = fmaf(h0[k], __ldg(Wh + k), a0);
        a1 = fmaf(h1[k], __ldg(Wh + w_gs + k), a1);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) { part[0][w] = a0; part[1][w] = a1; }
    __syncthreads();
    float q0 = __ldg(bh), q1 = __ldg(bh + b_gs);
#pragma unroll
    for (int ww = 0; ww < TH_WARPS; ++ww) { q0 += part[0][ww]; q1 += part[1][ww]; }
    const float invB = 1.f / (float)B;
    const float d0 = q0 < q1 ? -invB : (q0 == q1 ? -0.5f * invB : 0.f);       // torch.min(a, b): ties split the gradient evenly
    const float d1 = q1 < q0 ? -invB : (q0 == q1 ? -0.5f * invB : 0.f);
    if (threadIdx.x == 0) {
        q[m] = q0; q[q_gs + m] = q1;
        dq[m] = d0; dq[dq_gs + m] = d1;
        glp[m] = scalars[ORLK_SC_ALPHA] * invB;
    }
    for (int k = threadIdx.x; k < K; k += 32 * TH_WARPS) {
        dZ[(int64_t)m * K + k] = h0[k] > 0.f ? d0 * __ldg(Wh + k) : 0.f;
        dZ[dz_gs + (int64_t)m * K + k] = h1[k] > 0.f ? d1 * __ldg(Wh + w_gs + k) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------ CQL critic loss
// Bootstrap value of row b.  tq_rep == 1: min over the two target critics at (s', a'), minus alpha * log pi(a'|s') unless
// the backup is deterministic (cql.py:121-132).  tq_rep == N > 1 (max_q_backup, cql.py:109-120): each target critic is
// maximised over N sampled next actions first, and no entropy term is subtracted.
__device__ __forceinline__ float next_q(const float* __restrict__ tq, int64_t tq_cs, const float* __restrict__ lp_next, int b,
                                        int tq_rep, int det_backup, float alpha) {
    if (tq_rep <= 1) {
        float nq = fminf(tq[b], tq[tq_cs + b]);
        if (!det_backup) nq -= alpha * lp_next[b];
        return nq;
    }
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int r = 0; r < tq_rep; ++r) {
        m0 = fmaxf(m0, tq[(int64_t)b * tq_rep + r]);
        m1 = fmaxf(m1, tq[tq_cs + (int64_t)b * tq_rep + r]);
    }
    return fminf(m0, m1);
}

// Many CTAs: the per-row upstream gradients dq need no reduction at all (their only scalar factor is the OLD Lagrange
// multiplier), so every CTA writes its rows' dq in one pass and leaves six partial sums in `scratch`; the CTA that
// finishes last adds the partials in CTA order (deterministic), writes the losses and takes the multiplier's Adam step.
constexpr int CQL_LOSS_THREADS = 128;

__global__ void __launch_bounds__(CQL_LOSS_THREADS)
k_cql_critic_loss(const float* __restrict__ q, int64_t q_cs, const float* __restrict__ tq, int64_t tq_cs,
                  const float* __restrict__ lp_next, const float* __restrict__ lp_pi, const float* __restrict__ lp_pn,
                  const float* __restrict__ rew, const float* __restrict__ term, int B, int n_qmean, int tq_rep, int R,
                  float log_u, float gamma, float w, float T, int det_backup, int with_lagrange, float thr, float* __restrict__ scalars,
                  const OrlkAdamGroup* __restrict__ groups, int cql_alpha_group, float* __restrict__ cql_alpha_mv,
                  float* __restrict__ dq, int64_t dq_cs, float* __restrict__ out, float* __restrict__ scratch) {
    orlk::pdl_enter();
    __shared__ float red[32];
    __shared__ int s_last;
    const float alpha = scalars[ORLK_SC_ALPHA];
    const float invB = 1.f / (float)B, invQ = 1.f / (float)n_qmean, invR = 1.f / (float)R, invT = 1.f / T;
    float scale = 1.f, ex = 0.f, la = 0.f;
    if (with_lagrange) {        // the multiplier that scales this step's gradients is the one from BEFORE its own update
        la = scalars[ORLK_SC_CQL_LOG_ALPHA];
        ex = expf(la);
        scale = fminf(fmaxf(ex, 0.f), 1e6f);
    }
    float td[2] = {0.f, 0.f}, qs[2] = {0.f, 0.f}, ls[2] = {0.f, 0.f};
    const int i = blockIdx.x * CQL_LOSS_THREADS + threadIdx.x;
    if (i < B) {
        const int b = i;
        const float nq = next_q(tq, tq_cs, lp_next, b, tq_rep, det_backup, alpha);
        const float y = rew[b] + gamma * (1.f - term[b]) * nq;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float qq = q[c * q_cs + b];
            const float df = qq - y;
            td[c] = df * df;
            if (b < n_qmean) qs[c] = qq;
            dq[c * dq_cs + b] = 2.f * df * invB - (b < n_qmean ? w * scale * invQ : 0.f);
        }
    } else if (i < B + R) {
        const int r = i - B;
        const float l1 = lp_pi[r], l2 = lp_pn[r];
        const float k = scale * w * invR;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float* qc = q + c * q_cs + B;
            float* dc = dq + c * dq_cs + B;
            const float z1 = (qc[r] - l1) * invT, z2 = (qc[R + r] - l2) * invT, z3 = (qc[2 * R + r] - log_u) * invT;
            const float mx = fmaxf(z1, fmaxf(z2, z3));
            const float e1 = expf(z1 - mx), e2 = expf(z2 - mx), e3 = expf(z3 - mx);
            const float sum = e1 + e2 + e3;
            ls[c] = mx + logf(sum);
            const float inv = k / sum;
            dc[r] = e1 * inv;
            dc[R + r] = e2 * inv;
            dc[2 * R + r] = e3 * inv;
        }
    }
    float part[6];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        part[3 * c + 0] = block_sum(td[c], red);
        part[3 * c + 1] = block_sum(qs[c], red);
        part[3 * c + 2] = block_sum(ls[c], red);
    }
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
    float* parts = scratch + 4;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) parts[blockIdx.x * 6 + j] = part[j];
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float tot[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int blk = threadIdx.x; blk < (int)gridDim.x; blk += CQL_LOSS_THREADS) {
#pragma unroll
        for (int j = 0; j < 6; ++j) tot[j] += __ldcg(parts + blk * 6 + j);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) tot[j] = block_sum(tot[j], red);
    if (threadIdx.x == 0) {
        float cons[2], tdm[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            tdm[c] = tot[3 * c] * invB;
            cons[c] = tot[3 * c + 2] * invR * w * T - tot[3 * c + 1] * invQ * w;
        }
        float closs = 0.f, calpha = 0.f;
        float c0 = cons[0], c1 = cons[1];
        if (with_lagrange) {
            calpha = scale;
            c0 = calpha * (cons[0] - thr);
            c1 = calpha * (cons[1] - thr);
            closs = -(c0 + c1) * 0.5f;
            const float gate = (ex >= 0.f && ex <= 1e6f) ? 1.f : 0.f;
            const float g = -0.5f * ((cons[0] - thr) + (cons[1] - thr)) * ex * gate;
            scalars[ORLK_SC_CQL_LOG_ALPHA] = scalar_adam(la, g, cql_alpha_mv, groups[cql_alpha_group]);
        }
        out[0] = tdm[0] + c0;
        out[1] = tdm[1] + c1;
        out[2] = closs;
        out[3] = calpha;
        *counter = 0u;          // ready for the next launch (graph replay)
    }
}

// ------------------------------------------------------------------------------------------ fused Adam + polyak
// 128-thread blocks (64 registers each): 8 K registers, so an optimiser block fits on an SM next to a resident tensor-core
// CTA (~51 K of the 64 K registers) and the per-layer updates really overlap the weight-gradient GEMMs of the other layers
constexpr int ADAM_BLOCK_ELEMS = 128;

__global__ void __launch_bounds__(ADAM_BLOCK_ELEMS, 8)
k_adam_step(const OrlkAdamDesc* __restrict__ descs, int n_descs, const OrlkAdamGroup* __restrict__ groups) {
    orlk::pdl_enter();
    __shared__ OrlkAdamDesc sd;
    if (threadIdx.x < 32) {
        // this block's tensor = the last descriptor whose first block is <= blockIdx.x; the candidates are fetched by the
        // lanes of one warp at once (a scalar scan is one dependent L2 round trip per descriptor)
        int p = 0;
        for (int base = 0; base < n_descs; base += 32) {
            const int i = base + (int)threadIdx.x;
            const bool le = i < n_descs && descs[i].block_start <= (int)blockIdx.x;
            const unsigned m = __ballot_sync(0xffffffffu, le);
            if (m == 0u) break;
            p = base + 31 - __clz(m);
            if (m != 0xffffffffu) break;
        }
        const uint32_t* src = reinterpret_cast<const uint32_t*>(descs + p);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sd);
        for (int w = threadIdx.x; w < (int)(sizeof(OrlkAdamDesc) / 4); w += 32) dst[w] = src[w];
    }
    __syncthreads();
    const OrlkAdamDesc& d = sd;
    const OrlkAdamGroup g = groups[d.group];
    const float s_step_size = (float)((double)g.lr / g.bc1), s_bc2_sqrt = g.bc2_sqrt;
    const int64_t base = (int64_t)(blockIdx.x - d.block_start) * ADAM_BLOCK_ELEMS;
#pragma unroll
    for (int j = 0; j < 1; ++j) {
        const int64_t i = base + threadIdx.x;
        if (i >= d.n) continue;
        float p = d.p[i];
        if (d.flags & ORLK_OPT_ADAM) {
            // fixed-order reduction of the split-K partials, four independent chains so the loads overlap
            // (all loads of a 16-slot batch - and m, v - are issued before the first add: the plain loop costs one L2 round
            // trip per four slots; the order of the additions, and so the result, is that of the plain loop)
            float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
            const float* gp = d.grad + i;
            const float m_old = d.m[i], v_old = d.v[i];
            const int full = d.g_splits & ~3;
            for (int s0 = 0; s0 < d.g_splits; s0 += 16) {
                float pv[16];
                const float* gq = gp + (int64_t)s0 * d.g_split_stride;
#pragma unroll
                for (int u = 0; u < 16; ++u) pv[u] = (s0 + u < d.g_splits) ? gq[(int64_t)u * d.g_split_stride] : 0.f;
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                    if (s0 + u + 4 <= full) { g0 += pv[u]; g1 += pv[u + 1]; g2 += pv[u + 2]; g3 += pv[u + 3]; }
                    else {      // the last, incomplete group of four goes to the first chain, slot by slot
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (s0 + u + e < d.g_splits) g0 += pv[u + e];
                    }
                }
            }
            float gr = (g0 + g1) + (g2 + g3);
            if (d.wd != 0.f) gr = fmaf(d.wd, p, gr);
            float m = m_old, v = v_old;
            m = m + (gr - m) * (1.f - g.beta1);
            v = v * g.beta2 + (1.f - g.beta2) * gr * gr;
            d.m[i] = m; d.v[i] = v;
            const float denom = sqrtf(v) / s_bc2_sqrt + g.eps;
            p = p - s_step_size * (m / denom);
            d.p[i] = p;
            if (d.pT != nullptr) {      // keep the transposed (K-major for dgrad) copy of the weight in sync
                const int64_t r = i / d.cols, c = i % d.cols;
                d.pT[c * (d.n / d.cols) + r] = p;
            }
        }
        if ((d.flags & ORLK_OPT_POLYAK) && d.tgt != nullptr) d.tgt[i] = d.tgt[i] * (1.f - g.tau) + p * g.tau;
    }
}

__global__ void k_step_end(OrlkAdamGroup* groups, unsigned int mask, unsigned long long* counter) {
    // The counters (and betas) are written by this kernel and by host uploads only, so the next step's bias corrections
    // - two double precision pow() - are evaluated while the optimiser launches in front of this one are still running.
    orlk::pdl_trigger();
    const int g = threadIdx.x;
    const bool on = g < 32 && ((mask >> g) & 1u);
    int step = 0;
    double bc1 = 1.0;
    float bc2_sqrt = 1.f;
    if (on) {
        step = groups[g].step + 1;
        const int t = step + 1;
        bc1 = 1.0 - pow((double)groups[g].beta1, (double)t);
        bc2_sqrt = (float)sqrt(1.0 - pow((double)groups[g].beta2, (double)t));
    }
    orlk::pdl_wait();
    if (on) {
        groups[g].step = step;
        groups[g].bc1 = bc1;
        groups[g].bc2_sqrt = bc2_sqrt;
    }
    if (g == 0 && counter != nullptr) *counter += 1ull;
}

}  // namespace

extern "C" {

int orlk_skinny_fwd(const float* X, int64_t ldx, int64_t x_gs, const float* W, int64_t ldw, int64_t w_sk, int64_t w_gs,
                    const float* b, int64_t b_gs, float* Y, int64_t ldy, int64_t y_gs, int M, int K, int NS, int G,
                    void* stream) {
    ORLK_REQUIRE(NS >= 1 && NS <= MAX_NS, "NS must be in [1,16]");
    ORLK_REQUIRE(M > 0 && K > 0 && G > 0, "sizes");
    const int wpb = 8;
    dim3 grid((M + wpb - 1) / wpb, G);
    const int vec = (w_sk == 1) && (K % 4 == 0) && (ldx % 4 == 0) && (ldw % 4 == 0) && (x_gs % 4 == 0) && (w_gs % 4 == 0) &&
                    aligned16(X) && aligned16(W);
    cudaStream_t s = (cudaStream_t)stream;
    if (NS == 1) orlk::launch(k_skinny_fwd<1>, grid, wpb * 32, 0, s, X, ldx, x_gs, W, ldw, w_sk, w_gs, b, b_gs, Y, ldy, y_gs, M, K, NS, vec);
    else if (NS <= 4) orlk::launch(k_skinny_fwd<4>, grid, wpb * 32, 0, s, X, ldx, x_gs, W, ldw, w_sk, w_gs, b, b_gs, Y, ldy, y_gs, M, K, NS, vec);
    else if (NS <= 8) orlk::launch(k_skinny_fwd<8>, grid, wpb * 32, 0, s, X, ldx, x_gs, W, ldw, w_sk, w_gs, b, b_gs, Y, ldy, y_gs, M, K, NS, vec);
    else orlk::launch(k_skinny_fwd<16>, grid, wpb * 32, 0, s, X, ldx, x_gs, W, ldw, w_sk, w_gs, b, b_gs, Y, ldy, y_gs, M, K, NS, vec);
    return check_launch("k_skinny_fwd");
}

int orlk_skinny_dgrad(const float* dY, int64_t ldy, int64_t y_gs, const float* W, int64_t ldw, int64_t w_gs,
                      const float* mask, int64_t ldm, int64_t m_gs, float* dX, int64_t ldx, int64_t x_gs, float* dXT,
                      int64_t ldxt, int64_t xt_gs, int M, int K, int NS, int G, void* stream) {
    ORLK_REQUIRE(NS >= 1 && NS <= MAX_NS, "NS must be in [1,16]");
    ORLK_REQUIRE(M > 0 && K > 0 && G > 0, "sizes");
    // one block per 32-row strip when there are plenty of strips, otherwise also split the k range across blocks
    const int strips = (M + 31) / 32;
    const int kblocks = strips * G >= 592 ? 1 : (K + 31) / 32;
    dim3 grid(kblocks, strips, G);
    orlk::launch(k_skinny_dgrad, grid, dim3(32, 8), 0, (cudaStream_t)stream, dY, ldy, y_gs, W, ldw, w_gs, mask, ldm, m_gs, dX, ldx, x_gs,
                                                                dXT, ldxt, xt_gs, M, K, NS);
    return check_launch("k_skinny_dgrad");
}

int orlk_concat_rows(const OrlkConcatSeg* segs_dev, int n_segs, int total_rows, void* stream) {
    ORLK_REQUIRE(segs_dev != nullptr && n_segs > 0 && total_rows > 0, "segments");
    const int wpb = 8;
    orlk::launch(k_concat_rows, (total_rows + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream, segs_dev, n_segs, total_rows);
    return check_launch("k_concat_rows");
}

int orlk_compact_blocks(const float* src, int64_t block_stride, float* dst, int world, int per_member, const int* counts_host,
                        void* stream) {
    ORLK_REQUIRE(src != nullptr && dst != nullptr && world >= 1 && world <= 8 && per_member > 0 && counts_host != nullptr, "args");
    CompactArgs C;
    int mx = 0;
    for (int r = 0; r < 8; ++r) {
        C.counts[r] = r < world ? counts_host[r] : 0;
        mx = C.counts[r] > mx ? C.counts[r] : mx;
    }
    ORLK_REQUIRE(mx > 0 && (int64_t)mx * per_member <= block_stride, "a block must hold its members");
    const int64_t n = (int64_t)mx * per_member;
    const int bx = (int)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64);
    orlk::launch(k_compact_blocks, dim3(bx, world), 256, 0, (cudaStream_t)stream, src, block_stride, dst, world, per_member, C);
    return check_launch("k_compact_blocks");
}

int orlk_philox_fill(float* out, int64_t n_normal, int64_t n_uniform, float lo, float hi, uint64_t seed,
                     unsigned long long* counter, const int* enable, void* stream) {
    const int64_t n = n_normal + n_uniform;
    ORLK_REQUIRE(out != nullptr && n > 0, "sizes");
    const int64_t calls = (n + 3) / 4;
    orlk::launch(k_philox_fill, (unsigned)((calls + 255) / 256), 256, 0, (cudaStream_t)stream, out, n_normal, n_uniform, lo, hi, seed,
                                                                                    counter, enable);
    return check_launch("k_philox_fill");
}

int orlk_tanh_gauss_sample(const float* head, int64_t ld_head, int head_row_off, int rep, const float* eps, int M, int A,
                           float* act, int64_t ld_act, float* logp, const float* obs, int64_t ld_obs, int obs_dim,
                           float* xout, int64_t ld_x, void* stream) {
    ORLK_REQUIRE(M > 0 && A > 0 && A <= MAX_A && rep >= 1, "sizes");
    ORLK_REQUIRE(xout == nullptr || obs != nullptr, "xout needs obs");
    orlk::launch(k_tanh_gauss_sample, (M + 127) / 128, 128, 0, (cudaStream_t)stream, head, ld_head, head_row_off, rep, eps, M, A, act,
                                                                          ld_act, logp, obs, ld_obs, obs_dim, xout, ld_x);
    return check_launch("k_tanh_gauss_sample");
}

int orlk_tanh_gauss_bwd(const float* head, int64_t ld_head, const float* eps, const float* act, int64_t ld_act,
                        const float* dA, int n_da, int64_t da_gs, int64_t ld_da, const float* glp, int M, int A, float* dhead,
                        int64_t ld_dhead, void* stream) {
    ORLK_REQUIRE(M > 0 && A > 0 && A <= MAX_A && eps != nullptr, "sizes");
    ORLK_REQUIRE(n_da == 0 || dA != nullptr, "dA");
    orlk::launch(k_tanh_gauss_bwd, (M + 127) / 128, 128, 0, (cudaStream_t)stream, head, ld_head, eps, act, ld_act, dA, n_da, da_gs, ld_da,
                                                                       glp, M, A, dhead, ld_dhead);
    return check_launch("k_tanh_gauss_bwd");
}

int orlk_head_sample(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* b, float* head, int M, int K, int A,
                     const OrlkSampleUse* uses_host, int n_uses, void* stream) {
    ORLK_REQUIRE(M > 0 && K > 0 && K % 4 == 0 && A > 0 && A <= 8, "sizes (K % 4 == 0, A <= 8)");
    ORLK_REQUIRE(ldx % 4 == 0 && ldw % 4 == 0 && aligned16(X) && aligned16(W), "16-byte aligned rows");
    ORLK_REQUIRE(uses_host != nullptr && n_uses >= 1 && n_uses <= 4 && head != nullptr, "1..4 uses");
    SampleUses U;
    U.n = n_uses;
    for (int i = 0; i < n_uses; ++i) {
        ORLK_REQUIRE(uses_host[i].rep >= 1 && uses_host[i].act != nullptr && (uses_host[i].xout == nullptr || uses_host[i].obs != nullptr),
                     "use");
        U.u[i] = uses_host[i];
    }
    orlk::launch(k_head_sample, M, 32 * HS_WARPS, 0, (cudaStream_t)stream, X, ldx, W, ldw, b, head, M, K, A, U);
    return check_launch("k_head_sample");
}

int orlk_actor_bwd_entry(const float* dZ0, int64_t dz_gs, int Kc, int n_c, const float* W0, int64_t w0_gs, int ld_w0, int col0,
                         const float* head, const float* eps, const float* act, int64_t ld_act, const float* glp, int M, int A,
                         float* dhead, const float* Wh, int Ka, const float* Hlast, float* dZlast, void* stream) {
    ORLK_REQUIRE(M > 0 && A > 0 && A <= MAX_A && Kc > 0 && Ka > 0 && n_c > 0, "sizes");
    ORLK_REQUIRE(dZ0 && W0 && head && eps && act && dhead && Wh && Hlast && dZlast, "pointers");
    if (A <= 8)
        orlk::launch(k_actor_bwd_entry<8>, M, 32 * BE_WARPS, 0, (cudaStream_t)stream, dZ0, dz_gs, Kc, n_c, W0, w0_gs, ld_w0, col0,
                     head, eps, act, ld_act, glp, M, A, dhead, Wh, Ka, Hlast, dZlast);
    else
        orlk::launch(k_actor_bwd_entry<MAX_A>, M, 32 * BE_WARPS, 0, (cudaStream_t)stream, dZ0, dz_gs, Kc, n_c, W0, w0_gs, ld_w0,
                     col0, head, eps, act, ld_act, glp, M, A, dhead, Wh, Ka, Hlast, dZlast);
    return check_launch("k_actor_bwd_entry");
}

int orlk_sac_actor_loss(const float* q, int64_t q_es, int E, const float* logp, int B, float* scalars, int auto_alpha,
                        int clamp01, float target_entropy, OrlkAdamGroup* groups, int alpha_group, float* alpha_mv,
                        float* dq, int64_t dq_es, float* glp, float* out_losses, void* stream) {
    ORLK_REQUIRE(E >= 1 && B > 0, "sizes");
    ORLK_REQUIRE(!auto_alpha || (groups != nullptr && alpha_mv != nullptr), "auto alpha needs its Adam state");
    orlk::launch(k_sac_actor_loss, 1, 256, 0, (cudaStream_t)stream, q, q_es, E, logp, B, scalars, auto_alpha, clamp01, target_entropy,
                                                         groups, alpha_group, alpha_mv, dq, dq_es, glp, out_losses);
    return check_launch("k_sac_actor_loss");
}

int orlk_twin_head_actor(const float* H, int64_t h_gs, const float* Wh, int64_t w_gs, const float* bh, int64_t b_gs,
                         const float* scalars, int B, int K, float* q, int64_t q_gs, float* dq, int64_t dq_gs, float* glp,
                         float* dZ, int64_t dz_gs, void* stream) {
    ORLK_REQUIRE(H && Wh && bh && scalars && q && dq && glp && dZ && B > 0 && K > 0, "twin_head_actor arguments");
    orlk::launch(k_twin_head_actor, B, 32 * TH_WARPS, 0, (cudaStream_t)stream, H, h_gs, Wh, w_gs, bh, b_gs, scalars, B, K, q, q_gs, dq,
                 dq_gs, glp, dZ, dz_gs);
    return check_launch("k_twin_head_actor");
}

int orlk_cql_critic_loss_scratch_floats(int B, int R) {
    return 4 + 6 * ((B + R + CQL_LOSS_THREADS - 1) / CQL_LOSS_THREADS);
}

int orlk_cql_critic_loss(const float* q, int64_t q_cs, const float* tq, int64_t tq_cs, const float* lp_next,
                         const float* lp_pi, const float* lp_pn, const float* rew, const float* term, int B, int n_qmean,
                         int tq_rep, int R, int A, float gamma, float cql_weight, float temperature, int deterministic_backup, int with_lagrange,
                         float lagrange_threshold, float* scalars, OrlkAdamGroup* groups, int cql_alpha_group,
                         float* cql_alpha_mv, float* dq, int64_t dq_cs, float* out_losses, float* scratch, void* stream) {
    ORLK_REQUIRE(B > 0 && R > 0 && A > 0 && n_qmean > 0 && n_qmean <= B && tq_rep >= 1, "sizes");
    ORLK_REQUIRE(tq_rep > 1 || deterministic_backup || lp_next != nullptr, "the entropy backup needs lp_next");
    ORLK_REQUIRE(!with_lagrange || (groups != nullptr && cql_alpha_mv != nullptr), "lagrange needs its Adam state");
    ORLK_REQUIRE(scratch != nullptr && aligned16(scratch), "scratch (orlk_cql_critic_loss_scratch_floats zero-initialised floats)");
    const float log_u = (float)log(pow(0.5, (double)A));   // cql.py:82: np.log(0.5 ** act_dim)
    const int blocks = (B + R + CQL_LOSS_THREADS - 1) / CQL_LOSS_THREADS;
    orlk::launch(k_cql_critic_loss, blocks, CQL_LOSS_THREADS, 0, (cudaStream_t)stream, q, q_cs, tq, tq_cs, lp_next, lp_pi, lp_pn, rew, term, B,
                 n_qmean, tq_rep, R, log_u, gamma, cql_weight, temperature, deterministic_backup, with_lagrange, lagrange_threshold,
                 scalars, groups, cql_alpha_group, cql_alpha_mv, dq, dq_cs, out_losses, scratch);
    return check_launch("k_cql_critic_loss");
}

int orlk_adam_step(const OrlkAdamDesc* descs_dev, int n_descs, int total_blocks, const OrlkAdamGroup* groups, void* stream) {
    ORLK_REQUIRE(descs_dev != nullptr && n_descs > 0 && total_blocks > 0 && groups != nullptr, "descs");
    static int hp = -1;
    // (measured: with priority the optimiser blocks crowd the SMs and keep waiting tensor-core CTAs out: 255 -> 275 us per CQL
    // step; off by default, the launch order in emit_wgrad_adam gives the overlap instead)
    if (hp < 0) { const char* e = getenv("ORLK_ADAM_PRIORITY"); hp = e ? atoi(e) : 0; }
    if (hp) orlk::launch_high_priority(k_adam_step, total_blocks, ADAM_BLOCK_ELEMS, 0, (cudaStream_t)stream, descs_dev, n_descs, groups);
    else orlk::launch(k_adam_step, total_blocks, ADAM_BLOCK_ELEMS, 0, (cudaStream_t)stream, descs_dev, n_descs, groups);
    return check_launch("k_adam_step");
}

int orlk_step_end(OrlkAdamGroup* groups, unsigned int mask, unsigned long long* philox_counter, void* stream) {
    ORLK_REQUIRE(groups != nullptr, "groups");
    orlk::launch(k_step_end, 1, 32, 0, (cudaStream_t)stream, groups, mask, philox_counter);
    return check_launch("k_step_end");
}

}  // extern "C"
