// Loss epilogues of SAC / TD3+BC / IQL (single-CTA reduction kernels: B <= a few thousand rows) and the
// deterministic-actor head.  Each kernel writes the scalar losses, the per-row upstream gradients for the
// hand-derived backward pass and, where an optimiser owns a scalar or a tiny
// vector (IQL's sigma_param), the gradient for it.
#include <math.h>
#include "orlk_common.cuh"
using namespace orlk;

namespace {

constexpr float HALF_LOG_2PI = 0.91893853320467274178f;

// y = r + gamma (1-d) [ min_e2 tq[e2] - alpha*lp_next ];  loss_e = mean (q_e - y)^2;  dq_e = 2 (q_e - y) / B
__global__ void __launch_bounds__(1024)
k_td_loss(const float* __restrict__ q, int64_t q_es, int E, const float* __restrict__ tq, int64_t tq_es, int E2,
          const float* __restrict__ lp_next, const float* __restrict__ scalars, int use_alpha, const float* __restrict__ rew,
          const float* __restrict__ term, int B, float gamma, float* __restrict__ dq, int64_t dq_es, float* __restrict__ y_out,
          float* __restrict__ out_losses, float* __restrict__ out_sum) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const float alpha = use_alpha ? scalars[ORLK_SC_ALPHA] : 0.f;
    const float invB = 1.f / (float)B;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float nq = tq[b];
        for (int e = 1; e < E2; ++e) nq = fminf(nq, tq[e * tq_es + b]);
        if (use_alpha) nq -= alpha * lp_next[b];
        const float y = rew[b] + gamma * (1.f - term[b]) * nq;
        if (y_out) y_out[b] = y;
        for (int e = 0; e < E; ++e) dq[e * dq_es + b] = 2.f * (q[e * q_es + b] - y) * invB;
    }
    __syncthreads();
    float total = 0.f;
    for (int e = 0; e < E; ++e) {
        float s = 0.f;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const float d = dq[e * dq_es + b] * (0.5f * (float)B);     // = q_e - y
            s += d * d;
        }
        s = block_sum(s, red) * invB;
        if (threadIdx.x == 0) out_losses[e] = s;
        total += s;
    }
    if (threadIdx.x == 0 && out_sum != nullptr) *out_sum = total;
}

// IQL value loss (iql.py:82-98): q = min(tq0, tq1); w = tau if q - v > 0 else 1 - tau; L = mean w (q-v)^2
__global__ void __launch_bounds__(1024)
k_iql_v_loss(const float* __restrict__ tq, int64_t tq_es, const float* __restrict__ v, int B, float expectile,
             float* __restrict__ dv, float* __restrict__ qmin, float* __restrict__ out_loss) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const float invB = 1.f / (float)B;
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float qq = fminf(tq[b], tq[tq_es + b]);
        const float d = qq - v[b];
        const float w = d > 0.f ? expectile : 1.f - expectile;
        s += w * d * d;
        dv[b] = -2.f * w * d * invB;
        qmin[b] = qq;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) out_loss[0] = s * invB;
}

// IQL advantage-weighted actor loss (iql.py:118-130) for DiagGaussian(unbounded=False, state-independent sigma):
//   mu = max_mu * tanh(z);  sigma_i = exp(sp_i);  logp_b = sum_i [ -(a-mu)^2 / (2 sigma^2) - sp_i - c ]
//   w_b = min(exp((q_b - v_b) * temp), 100);  L = -mean_b w_b logp_b
__global__ void __launch_bounds__(1024)
k_iql_actor_loss(const float* __restrict__ z, int64_t ldz, const float* __restrict__ sigma_param, const float* __restrict__ act,
                 int64_t lda, const float* __restrict__ qmin, const float* __restrict__ v, int B, int A, float temp,
                 float max_mu, float* __restrict__ dz, int64_t lddz, float* __restrict__ dsigma, float* __restrict__ out_loss) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const float invB = 1.f / (float)B;
    float loss = 0.f;
    float ds[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) ds[i] = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float w = fminf(expf((qmin[b] - v[b]) * temp), 100.f);
        float lp = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i < A) {
                const float sp = sigma_param[i];
                const float sigma = expf(sp);
                const float t = tanhf(z[(int64_t)b * ldz + i]);
                const float mu = max_mu * t;
                const float d = act[(int64_t)b * lda + i] - mu;
                const float var = sigma * sigma;
                lp += -(d * d) / (2.f * var) - logf(sigma) - HALF_LOG_2PI;
                // dL/dmu = -w (a-mu)/var / B ; through mu = max_mu tanh(z)
                dz[(int64_t)b * lddz + i] = -w * (d / var) * invB * max_mu * (1.f - t * t);
                ds[i] += -w * ((d * d) / var - 1.f) * invB;       // dL/d sigma_param
            }
        }
        loss += -w * lp;
    }
    loss = block_sum(loss, red);
    if (threadIdx.x == 0) out_loss[0] = loss * invB;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < A) {
            const float s = block_sum(ds[i], red);
            if (threadIdx.x == 0) dsigma[i] = s;
        }
    }
}

// Deterministic actor head (actor_module.py:46-50, td3bc.py:90-91):
//   a = max * tanh(z);  with eps: a = clamp(a + clamp(policy_noise*eps, -clip, clip), -max, max)
// writes a into act (may alias critic-input columns) and optionally the obs columns of the critic input row.
__global__ void k_det_actor_fwd(const float* __restrict__ z, int64_t ldz, const float* __restrict__ eps, int M, int A,
                                float max_action, float policy_noise, float noise_clip, float* __restrict__ act, int64_t ld_act,
                                const float* __restrict__ obs, int64_t ld_obs, int obs_dim, float* __restrict__ xout,
                                int64_t ld_x) {
    orlk::pdl_enter();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (int i = 0; i < A; ++i) {
        float a = max_action * tanhf(z[(int64_t)m * ldz + i]);
        if (eps != nullptr) {
            const float n = fminf(fmaxf(eps[(int64_t)m * A + i] * policy_noise, -noise_clip), noise_clip);
            a = fminf(fmaxf(a + n, -max_action), max_action);
        }
        act[(int64_t)m * ld_act + i] = a;
    }
    if (xout != nullptr)
        for (int j = 0; j < obs_dim; ++j) xout[(int64_t)m * ld_x + j] = obs[(int64_t)m * ld_obs + j];
}

// TD3+BC actor loss (td3bc.py:107-112): lambda = alpha / mean|q| (detached);  L = -lambda mean q + mean (a - a_data)^2
//   dq_b = -lambda / B ;  dabc[b,i] = 2 (a - a_data) / (B A)
__global__ void __launch_bounds__(1024)
k_td3bc_actor_loss(const float* __restrict__ q, const float* __restrict__ a, int64_t lda, const float* __restrict__ a_data,
                   int64_t ldd, int B, int A, float bc_alpha, float* __restrict__ dq, float* __restrict__ dabc, int64_t ldg,
                   float* __restrict__ out_loss) {
    orlk::pdl_enter();
    __shared__ float red[32];
    float sabs = 0.f, sq = 0.f, sbc = 0.f;
    const float invB = 1.f / (float)B, invBA = 1.f / ((float)B * (float)A);
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float qq = q[b];
        sabs += fabsf(qq);
        sq += qq;
        for (int i = 0; i < A; ++i) {
            const float d = a[(int64_t)b * lda + i] - a_data[(int64_t)b * ldd + i];
            sbc += d * d;
            dabc[(int64_t)b * ldg + i] = 2.f * d * invBA;
        }
    }
    sabs = block_sum(sabs, red);
    sq = block_sum(sq, red);
    sbc = block_sum(sbc, red);
    const float lambda = bc_alpha / (sabs * invB);
    for (int b = threadIdx.x; b < B; b += blockDim.x) dq[b] = -lambda * invB;
    if (threadIdx.x == 0) out_loss[0] = -lambda * (sq * invB) + sbc * invBA;
}

// dz = (dA0 + dA1) * max * (1 - tanh(z)^2) with tanh(z) = a / max
__global__ void k_det_actor_bwd(const float* __restrict__ a, int64_t lda, const float* __restrict__ dA0, int64_t ld0,
                                const float* __restrict__ dA1, int64_t ld1, int M, int A, float max_action,
                                float* __restrict__ dz, int64_t lddz) {
    orlk::pdl_enter();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (int i = 0; i < A; ++i) {
        const float t = a[(int64_t)m * lda + i] / max_action;
        float g = dA0[(int64_t)m * ld0 + i];
        if (dA1 != nullptr) g += dA1[(int64_t)m * ld1 + i];
        dz[(int64_t)m * lddz + i] = g * max_action * (1.f - t * t);
    }
}

// EDAC diversity loss and its gradient w.r.t. the input gradients g.  One thread per sample.
__global__ void __launch_bounds__(256)
k_edac_div(const float* __restrict__ g, int E, int B, int A, float eta, float* __restrict__ gbar, float* __restrict__ partial) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float Gb = 0.f;
    if (b < B) {
        float S[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) S[i] = 0.f;
        float sumsq = 0.f;
        for (int e = 0; e < E; ++e) {
            const float* ge = g + ((int64_t)e * B + b) * A;
            float n2 = 0.f;
            for (int i = 0; i < A; ++i) n2 += ge[i] * ge[i];
            const float inv = 1.f / (sqrtf(n2) + 1e-10f);
            for (int i = 0; i < A; ++i) {
                const float h = ge[i] * inv;
                S[i] += h;
                sumsq += h * h;
            }
        }
        float s2 = 0.f;
        for (int i = 0; i < A; ++i) s2 += S[i] * S[i];
        Gb = (s2 - sumsq) / (float)(E - 1);
        const float c = 2.f * eta / ((float)B * (float)(E - 1));
        for (int e = 0; e < E; ++e) {
            const float* ge = g + ((int64_t)e * B + b) * A;
            float* ob = gbar + ((int64_t)e * B + b) * A;
            float n2 = 0.f;
            for (int i = 0; i < A; ++i) n2 += ge[i] * ge[i];
            const float n = sqrtf(n2), d = n + 1e-10f;
            // hbar_i = c (S_i - ghat_i);  gbar = hbar/d - g <g,hbar> / (n d^2)
            float dot = 0.f;
            for (int i = 0; i < A; ++i) dot += ge[i] * (c * (S[i] - ge[i] / d));
            const float k = n > 0.f ? dot / (n * d * d) : 0.f;
            for (int i = 0; i < A; ++i) ob[i] = c * (S[i] - ge[i] / d) / d - ge[i] * k;
        }
    }
    Gb = block_sum(Gb, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = Gb;
}

__global__ void k_edac_div_final(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
    orlk::pdl_enter();
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += partial[i];
    out[0] = s * scale;
}

// out[g][b] = max_{r < rep} x[g][b * rep + r]   (max_q_backup: the best of `rep` sampled next actions per row)
__global__ void k_segment_max(const float* __restrict__ x, int64_t x_gs, int G, int B, int rep, float* __restrict__ out,
                              int64_t out_gs) {
    orlk::pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * B) return;
    const int g = i / B, b = i - g * B;
    const float* src = x + g * x_gs + (int64_t)b * rep;
    float m = -INFINITY;
    for (int r = 0; r < rep; ++r) m = fmaxf(m, src[r]);
    out[g * out_gs + b] = m;
}

}  // namespace

extern "C" {

int orlk_segment_max(const float* x, int64_t x_gs, int G, int B, int rep, float* out, int64_t out_gs, void* stream) {
    ORLK_REQUIRE(x != nullptr && out != nullptr && G >= 1 && B > 0 && rep >= 1, "sizes");
    orlk::launch(k_segment_max, (G * B + 255) / 256, 256, 0, (cudaStream_t)stream, x, x_gs, G, B, rep, out, out_gs);
    return check_launch("k_segment_max");
}

int orlk_td_loss(const float* q, int64_t q_es, int E, const float* tq, int64_t tq_es, int E2, const float* lp_next,
                 const float* scalars, int use_alpha, const float* rew, const float* term, int B, float gamma, float* dq,
                 int64_t dq_es, float* y_out, float* out_losses, float* out_sum, void* stream) {
    ORLK_REQUIRE(E >= 1 && E2 >= 1 && B > 0, "sizes");
    ORLK_REQUIRE(!use_alpha || (lp_next != nullptr && scalars != nullptr), "alpha term needs lp_next and scalars");
    orlk::launch(k_td_loss, 1, 1024, 0, (cudaStream_t)stream, q, q_es, E, tq, tq_es, E2, lp_next, scalars, use_alpha, rew, term, B, gamma,
                                                   dq, dq_es, y_out, out_losses, out_sum);
    return check_launch("k_td_loss");
}

int orlk_iql_v_loss(const float* tq, int64_t tq_es, const float* v, int B, float expectile, float* dv, float* qmin,
                    float* out_loss, void* stream) {
    ORLK_REQUIRE(B > 0, "sizes");
    orlk::launch(k_iql_v_loss, 1, 1024, 0, (cudaStream_t)stream, tq, tq_es, v, B, expectile, dv, qmin, out_loss);
    return check_launch("k_iql_v_loss");
}

int orlk_iql_actor_loss(const float* z, int64_t ldz, const float* sigma_param, const float* act, int64_t lda, const float* qmin,
                        const float* v, int B, int A, float temperature, float max_mu, float* dz, int64_t lddz, float* dsigma,
                        float* out_loss, void* stream) {
    ORLK_REQUIRE(B > 0 && A > 0 && A <= 32, "sizes");
    orlk::launch(k_iql_actor_loss, 1, 1024, 0, (cudaStream_t)stream, z, ldz, sigma_param, act, lda, qmin, v, B, A, temperature, max_mu, dz,
                                                          lddz, dsigma, out_loss);
    return check_launch("k_iql_actor_loss");
}

int orlk_det_actor_fwd(const float* z, int64_t ldz, const float* eps, int M, int A, float max_action, float policy_noise,
                       float noise_clip, float* act, int64_t ld_act, const float* obs, int64_t ld_obs, int obs_dim, float* xout,
                       int64_t ld_x, void* stream) {
    ORLK_REQUIRE(M > 0 && A > 0, "sizes");
    ORLK_REQUIRE(xout == nullptr || obs != nullptr, "xout needs obs");
    orlk::launch(k_det_actor_fwd, (M + 127) / 128, 128, 0, (cudaStream_t)stream, z, ldz, eps, M, A, max_action, policy_noise, noise_clip, act,
                                                                      ld_act, obs, ld_obs, obs_dim, xout, ld_x);
    return check_launch("k_det_actor_fwd");
}

int orlk_td3bc_actor_loss(const float* q, const float* a, int64_t lda, const float* a_data, int64_t ldd, int B, int A,
                          float bc_alpha, float* dq, float* dabc, int64_t ldg, float* out_loss, void* stream) {
    ORLK_REQUIRE(B > 0 && A > 0, "sizes");
    orlk::launch(k_td3bc_actor_loss, 1, 1024, 0, (cudaStream_t)stream, q, a, lda, a_data, ldd, B, A, bc_alpha, dq, dabc, ldg, out_loss);
    return check_launch("k_td3bc_actor_loss");
}

int orlk_edac_div(const float* g, int E, int B, int A, float eta, float* gbar, float* scratch, float* out_loss, void* stream) {
    ORLK_REQUIRE(E >= 2 && B > 0 && A > 0 && A <= 32, "sizes");
    ORLK_REQUIRE(scratch != nullptr, "scratch (ceil(B/256) floats)");
    const int blocks = (B + 255) / 256;
    orlk::launch(k_edac_div, blocks, 256, 0, (cudaStream_t)stream, g, E, B, A, eta, gbar, scratch);
    orlk::launch(k_edac_div_final, 1, 1, 0, (cudaStream_t)stream, scratch, blocks, eta / (float)B, out_loss);
    return check_launch("k_edac_div");
}

int orlk_det_actor_bwd(const float* a, int64_t lda, const float* dA0, int64_t ld0, const float* dA1, int64_t ld1, int M, int A,
                       float max_action, float* dz, int64_t lddz, void* stream) {
    ORLK_REQUIRE(M > 0 && A > 0 && dA0 != nullptr, "sizes");
    orlk::launch(k_det_actor_bwd, (M + 127) / 128, 128, 0, (cudaStream_t)stream, a, lda, dA0, ld0, dA1, ld1, M, A, max_action, dz, lddz);
    return check_launch("k_det_actor_bwd");
}

}  // extern "C"
