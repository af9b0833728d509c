// MOPO ensemble-dynamics kernels: input scaling, Gaussian-NLL training loss with the soft-clamped log-variance,
// holdout MSE, the imagination epilogue (elite pick, sampling, termination, uncertainty penalty), survivor
// compaction and bootstrap row gathers.  Reference: dynamics/ensemble_dynamics.py:28-79,178-217,
// modules/dynamics_module.py:19-29,87-94, utils/scaler.py:25-31, utils/termination_fns.py:10-30,63-75,
// policy/model_based/mopo.py:45-79.
#include <math.h>
#include "orlk_common.cuh"
using namespace orlk;

namespace {

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus (threshold 20)
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// X[s, :] = ([obs[s] | act[s]] - mu) / std        (scaler.transform on the concatenated input, fp32 like NumPy)
__global__ void k_dyn_input(const float* __restrict__ obs, int64_t ld_obs, const float* __restrict__ act, int64_t ld_act,
                            const float* __restrict__ mu, const float* __restrict__ sd, int S, int O, int A,
                            float* __restrict__ X, int64_t ldx) {
    orlk::pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int W = O + A;
    if (i >= (int64_t)S * W) return;
    const int s = (int)(i / W), j = (int)(i % W);
    const float v = j < O ? obs[(int64_t)s * ld_obs + j] : act[(int64_t)s * ld_act + (j - O)];
    X[(int64_t)s * ldx + j] = (v - mu[j]) / sd[j];
}

// dst[e][r][:] = src[idx[e*idx_ld + r0 + r]][:]   (per-member bootstrap batch, ensemble_dynamics.py:134,144,186-187)
__global__ void k_gather_rows(const float* __restrict__ src, int64_t ld_src, int w, const int64_t* __restrict__ idx,
                              int64_t idx_ld, int64_t r0, int E, int R, float* __restrict__ dst, int64_t ld_dst, int64_t dst_es) {
    orlk::pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (int64_t)E * R) return;
    const int e = (int)(row / R), r = (int)(row % R);
    const int64_t sidx = idx[(int64_t)e * idx_ld + r0 + r];
    const float* s = src + sidx * ld_src;
    float* d = dst + (int64_t)e * dst_es + (int64_t)r * ld_dst;
    for (int j = lane; j < w; j += 32) d[j] = s[j];
}

// sum of squares of [n] floats in 4096-element chunks (weight-decay term of the reported loss)
__global__ void __launch_bounds__(256)
k_sumsq(const float* __restrict__ x, int64_t n, float scale, float* __restrict__ partial) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const int64_t base = (int64_t)blockIdx.x * 4096;
    float s = 0.f;
    for (int j = threadIdx.x; j < 4096; j += 256) {
        const int64_t i = base + j;
        if (i < n) { const float v = x[i]; s += v * v; }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s * scale;
}

// Gaussian NLL with soft-clamped log-variance (ensemble_dynamics.py:193-201, dynamics_module.py:19-29,92-93).
//   out[e][b][0:D] = mean, out[e][b][D:2D] = raw logvar.   One CTA; E*b*D <= ~100k elements.
//   loss = sum_e mean_{b,d}[(mean-y)^2 e^{-lv}] + sum_e mean_{b,d} lv + decay + coef (sum max_lv - sum min_lv)
__global__ void __launch_bounds__(1024)
k_dyn_nll(const float* __restrict__ out, const float* __restrict__ y, int E, int Bn, int D, const float* __restrict__ max_lv,
          const float* __restrict__ min_lv, float coef, const float* __restrict__ decay_partials, int n_decay,
          float* __restrict__ dout, float* __restrict__ dmax, float* __restrict__ dmin, float* __restrict__ out_loss) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const float inv = 1.f / ((float)Bn * (float)D);
    float loss = 0.f;
    // thread <-> fixed d so that the per-d bound gradients reduce without atomics: d = tid % D
    const int n_rows = E * Bn;
    const int lanes_per_d = blockDim.x / D;            // threads sharing one d
    const int d = threadIdx.x % D, slot = threadIdx.x / D;
    float gmax = 0.f, gmin = 0.f;
    if (slot < lanes_per_d) {
        const float mx = max_lv[d], mn = min_lv[d];
        for (int r = slot; r < n_rows; r += lanes_per_d) {
            const float* o = out + (int64_t)r * 2 * D;
            const float mean = o[d], raw = o[D + d];
            const float t1 = mx - raw;
            const float u = mx - softplus_t(t1);
            const float t2 = u - mn;
            const float lv = mn + softplus_t(t2);
            const float iv = expf(-lv);
            const float diff = mean - y[(int64_t)r * D + d];
            loss += (diff * diff * iv + lv) * inv;
            const float dlv = (1.f - diff * diff * iv) * inv;
            const float s2 = sigmoid_f(t2), s1 = sigmoid_f(t1);
            dout[(int64_t)r * 2 * D + d] = 2.f * diff * iv * inv;
            dout[(int64_t)r * 2 * D + D + d] = dlv * s2 * s1;
            gmax += dlv * s2 * (1.f - s1);
            gmin += dlv * (1.f - s2);
        }
    }
    // reduce gmax/gmin over the threads that share d (fixed order -> deterministic)
    __shared__ float s_g[2][1024];
    s_g[0][threadIdx.x] = gmax;
    s_g[1][threadIdx.x] = gmin;
    __syncthreads();
    if (threadIdx.x < D) {
        float a = 0.f, b = 0.f;
        for (int s = 0; s < lanes_per_d; ++s) { a += s_g[0][s * D + threadIdx.x]; b += s_g[1][s * D + threadIdx.x]; }
        dmax[threadIdx.x] = a + coef;
        dmin[threadIdx.x] = b - coef;
    }
    float bound = 0.f;
    if (threadIdx.x < D) bound = coef * (max_lv[threadIdx.x] - min_lv[threadIdx.x]);
    float dec = 0.f;
    for (int i = threadIdx.x; i < n_decay; i += blockDim.x) dec += decay_partials[i];
    loss = block_sum(loss, red);
    bound = block_sum(bound, red);
    dec = block_sum(dec, red);
    if (threadIdx.x == 0) out_loss[0] = loss + dec + bound;
}

// Multi-CTA form of the same loss: 64 rows per CTA write dout and one partial record [loss | gmax[D] | gmin[D]]; a single
// small CTA sums the records in block order (deterministic) and adds the decay / bound terms.
constexpr int NLL_ROWS = 64;

__global__ void __launch_bounds__(256)
k_dyn_nll_part(const float* __restrict__ out, const float* __restrict__ y, int n_rows, int Bn, int D,
               const float* __restrict__ max_lv, const float* __restrict__ min_lv, float* __restrict__ dout,
               float* __restrict__ partial) {
    orlk::pdl_enter();
    __shared__ float red[32];
    __shared__ float s_g[2][256];
    const float inv = 1.f / ((float)Bn * (float)D);
    const int lanes_per_d = blockDim.x / D;
    const int d = threadIdx.x % D, slot = threadIdx.x / D;
    const int r0 = blockIdx.x * NLL_ROWS, r1 = min(n_rows, r0 + NLL_ROWS);
    float loss = 0.f, gmax = 0.f, gmin = 0.f;
    if (slot < lanes_per_d) {
        const float mx = max_lv[d], mn = min_lv[d];
        for (int r = r0 + slot; r < r1; r += lanes_per_d) {
            const float* o = out + (int64_t)r * 2 * D;
            const float mean = o[d], raw = o[D + d];
            const float t1 = mx - raw;
            const float u = mx - softplus_t(t1);
            const float t2 = u - mn;
            const float lv = mn + softplus_t(t2);
            const float iv = expf(-lv);
            const float diff = mean - y[(int64_t)r * D + d];
            loss += (diff * diff * iv + lv) * inv;
            const float dlv = (1.f - diff * diff * iv) * inv;
            const float s2 = sigmoid_f(t2), s1 = sigmoid_f(t1);
            dout[(int64_t)r * 2 * D + d] = 2.f * diff * iv * inv;
            dout[(int64_t)r * 2 * D + D + d] = dlv * s2 * s1;
            gmax += dlv * s2 * (1.f - s1);
            gmin += dlv * (1.f - s2);
        }
    }
    s_g[0][threadIdx.x] = gmax;
    s_g[1][threadIdx.x] = gmin;
    __syncthreads();
    float* rec = partial + (int64_t)blockIdx.x * (1 + 2 * D);
    if (threadIdx.x < D) {
        float a = 0.f, b = 0.f;
        for (int s = 0; s < lanes_per_d; ++s) { a += s_g[0][s * D + threadIdx.x]; b += s_g[1][s * D + threadIdx.x]; }
        rec[1 + threadIdx.x] = a;
        rec[1 + D + threadIdx.x] = b;
    }
    loss = block_sum(loss, red);
    if (threadIdx.x == 0) rec[0] = loss;
}

__global__ void __launch_bounds__(256)
k_dyn_nll_final(const float* __restrict__ partial, int n_blocks, int D, float coef, const float* __restrict__ max_lv,
                const float* __restrict__ min_lv, const float* __restrict__ decay_partials, int n_decay,
                float* __restrict__ dmax, float* __restrict__ dmin, float* __restrict__ out_loss) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const int W = 1 + 2 * D;
    if (threadIdx.x < 2 * D) {
        float a = 0.f;
        for (int b = 0; b < n_blocks; ++b) a += partial[(int64_t)b * W + 1 + threadIdx.x];
        if (threadIdx.x < D) dmax[threadIdx.x] = a + coef;
        else dmin[threadIdx.x - D] = a - coef;
    }
    float loss = 0.f;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) loss += partial[(int64_t)b * W];
    float bound = threadIdx.x < D ? coef * (max_lv[threadIdx.x] - min_lv[threadIdx.x]) : 0.f;
    float dec = 0.f;
    for (int i = threadIdx.x; i < n_decay; i += blockDim.x) dec += decay_partials[i];
    loss = block_sum(loss, red);
    bound = block_sum(bound, red);
    dec = block_sum(dec, red);
    if (threadIdx.x == 0) out_loss[0] = loss + dec + bound;
}

// per-member holdout MSE of the mean head (ensemble_dynamics.py:210-217): one CTA per member
__global__ void __launch_bounds__(256)
k_dyn_val_mse(const float* __restrict__ out, const float* __restrict__ y, int Bn, int D, float* __restrict__ mse) {
    orlk::pdl_enter();
    __shared__ float red[32];
    const int e = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < Bn * D; i += blockDim.x) {
        const int b = i / D, d = i % D;
        const float diff = out[((int64_t)e * Bn + b) * 2 * D + d] - y[(int64_t)b * D + d];
        s += diff * diff;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) mse[e] = s / ((float)Bn * (float)D);
}

__device__ __forceinline__ bool terminal_of(int kind, const float* nobs, int O) {
    // utils/termination_fns.py: 0 halfcheetah (:10-16), 1 hopper (:18-30, sic: only the upper bound on |obs[1:]| is
    // enforced because np.abs() is applied to a boolean array), 2 walker2d (:63-75), 3 never
    if (kind == 3) return false;
    if (kind == 1) {
        bool ok = true;
        for (int j = 0; j < O; ++j) ok = ok && isfinite(nobs[j]);
        for (int j = 1; j < O; ++j) ok = ok && (nobs[j] < 100.f);
        ok = ok && (nobs[0] > 0.7f) && (fabsf(nobs[1]) < 0.2f);
        return !ok;
    }
    bool ok = true;
    for (int j = 0; j < O; ++j) ok = ok && (nobs[j] > -100.f) && (nobs[j] < 100.f);
    if (kind == 2) ok = ok && (nobs[0] > 0.8f) && (nobs[0] < 2.0f) && (nobs[1] > -1.0f) && (nobs[1] < 1.0f);
    return !ok;
}

// Imagination epilogue (ensemble_dynamics.py:43-77): one thread per state.
//   out[e][s][0:D] mean (delta-obs | reward), out[e][s][D:2D] raw logvar -> soft clamp -> std = sqrt(exp(lv))
//   sample = (mean_e* + noise * std_e*) of the chosen elite e* = midx[s]  (float64 arithmetic, cast to fp32 like NumPy)
//   penalty = max_e ||std_e||_2 ;  reward -= coef * penalty
__global__ void k_dyn_step(const float* __restrict__ out, int E, int S, int D, const float* __restrict__ max_lv,
                           const float* __restrict__ min_lv, const float* __restrict__ obs, int64_t ld_obs,
                           const double* __restrict__ noise, const int* __restrict__ midx, const float* __restrict__ noise32,
                           const float* __restrict__ pick_u, const int* __restrict__ elites, int n_elites, int term_kind,
                           float penalty_coef, int unc_mode,
                           float* __restrict__ next_obs, float* __restrict__ reward, float* __restrict__ raw_reward,
                           float* __restrict__ penalty, unsigned char* __restrict__ terminal) {
    orlk::pdl_enter();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int O = D - 1;
    // elite member of this state: given (reference stream), or drawn from a device uniform in [0,1)
    const int em = midx != nullptr ? midx[s] : elites[min(n_elites - 1, (int)(pick_u[s] * (float)n_elites))];
    float pen = 0.f;
    float samp[64];
    for (int e = 0; e < E; ++e) {
        const float* o = out + ((int64_t)e * S + s) * 2 * D;
        float n2 = 0.f;
        for (int d = 0; d < D; ++d) {
            const float raw = o[D + d];
            const float u = max_lv[d] - softplus_t(max_lv[d] - raw);
            const float lv = min_lv[d] + softplus_t(u - min_lv[d]);
            const float sd = sqrtf(expf(lv));
            n2 += sd * sd;
            if (e == em) {
                float mean = o[d];
                if (d < O) mean += obs[(int64_t)s * ld_obs + d];
                const double z = noise != nullptr ? noise[((int64_t)e * S + s) * D + d] : (double)noise32[(int64_t)s * D + d];
                samp[d] = (float)((double)mean + z * (double)sd);
            }
        }
        pen = fmaxf(pen, sqrtf(n2));
    }
    if (unc_mode != 0) {
        // disagreement of the members' predicted next states m_e = obs + delta_e  (ensemble_dynamics.py:63-70):
        //   1 "pairwise-diff": max_e || m_e - mean_e m_e ||      2 "ensemble_std": sqrt( mean_d var_e m_e[d] )
        float mbar[64];
        for (int d = 0; d < O; ++d) mbar[d] = 0.f;
        for (int e = 0; e < E; ++e) {
            const float* o = out + ((int64_t)e * S + s) * 2 * D;
            for (int d = 0; d < O; ++d) mbar[d] += o[d] + obs[(int64_t)s * ld_obs + d];
        }
        for (int d = 0; d < O; ++d) mbar[d] /= (float)E;
        float worst = 0.f, var_sum = 0.f;
        for (int e = 0; e < E; ++e) {
            const float* o = out + ((int64_t)e * S + s) * 2 * D;
            float n2 = 0.f;
            for (int d = 0; d < O; ++d) {
                const float df = (o[d] + obs[(int64_t)s * ld_obs + d]) - mbar[d];
                n2 += df * df;
            }
            worst = fmaxf(worst, sqrtf(n2));
            var_sum += n2;
        }
        pen = unc_mode == 1 ? worst : sqrtf(var_sum / (float)E / (float)O);
    }
    for (int d = 0; d < O; ++d) next_obs[(int64_t)s * O + d] = samp[d];
    raw_reward[s] = samp[O];
    penalty[s] = pen;
    reward[s] = penalty_coef != 0.f ? samp[O] - penalty_coef * pen : samp[O];
    terminal[s] = terminal_of(term_kind, samp, O) ? 1 : 0;
}

// Stable compaction of the rows whose flag is 0 (mopo.py:69-73): single CTA, two-level scan.
__global__ void __launch_bounds__(1024)
k_compact_rows(const unsigned char* __restrict__ drop, int S, const float* __restrict__ src, int64_t ld_src, int w,
               float* __restrict__ dst, int64_t ld_dst, int* __restrict__ count_out) {
    orlk::pdl_enter();
    __shared__ int sums[1024];
    const int per = (S + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(S, lo + per);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += drop[i] ? 0 : 1;
    sums[threadIdx.x] = c;
    __syncthreads();
    // inclusive scan (Hillis-Steele) over the 1024 per-thread counts
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = threadIdx.x >= off ? sums[threadIdx.x - off] : 0;
        __syncthreads();
        sums[threadIdx.x] += v;
        __syncthreads();
    }
    int pos = sums[threadIdx.x] - c;
    if (threadIdx.x == 1023) *count_out = sums[1023];
    for (int i = lo; i < hi; ++i) {
        if (!drop[i]) {
            for (int j = 0; j < w; ++j) dst[(int64_t)pos * ld_dst + j] = src[(int64_t)i * ld_src + j];
            ++pos;
        }
    }
}

// Multi-CTA form of the same stable compaction, two launches over 256-row blocks:
//   count:   counts[b] = survivors of block b
//   scatter: block b starts at sum(counts[0..b)), a survivor's slot inside the block is its rank among the block's
//            survivors (warp ballots + a scan of the 8 warp totals); the last block publishes the total.
constexpr int CMP_ROWS = 256;

__global__ void __launch_bounds__(CMP_ROWS)
k_compact_count(const unsigned char* __restrict__ drop, int S, int* __restrict__ counts) {
    orlk::pdl_enter();
    const int i = blockIdx.x * CMP_ROWS + threadIdx.x;
    const int keep = (i < S && !drop[i]) ? 1 : 0;
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(CMP_ROWS)
k_compact_scatter(const unsigned char* __restrict__ drop, int S, const float* __restrict__ src, int64_t ld_src, int w,
                  float* __restrict__ dst, int64_t ld_dst, const int* __restrict__ counts, int* __restrict__ count_out) {
    orlk::pdl_enter();
    __shared__ int red[CMP_ROWS / 32];
    __shared__ int wsum[CMP_ROWS / 32];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += CMP_ROWS) part += counts[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    const int i = blockIdx.x * CMP_ROWS + threadIdx.x;
    const bool keep = i < S && !drop[i];
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = 0;
        for (int k = 0; k < CMP_ROWS / 32; ++k) b += red[k];
        base_s = b;
        if (blockIdx.x == gridDim.x - 1) {
            int tot = b;
            for (int k = 0; k < CMP_ROWS / 32; ++k) tot += wsum[k];
            *count_out = tot;
        }
    }
    __syncthreads();
    if (keep) {
        int pos = base_s + __popc(bal & ((1u << lane) - 1u));
        for (int k = 0; k < warp; ++k) pos += wsum[k];
        const float* s_row = src + (int64_t)i * ld_src;
        float* d_row = dst + (int64_t)pos * ld_dst;
        for (int j = 0; j < w; ++j) d_row[j] = s_row[j];
    }
}

// per-device scratch of the two-launch compaction (block counts), grown on demand outside stream capture
int* compact_scratch(int n_blocks) {
    static int* buf[64] = {nullptr};
    static int cap[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (cap[dev] < n_blocks) {
        int want = 1024;
        while (want < n_blocks) want *= 2;
        int* nb = nullptr;
        if (cudaMalloc(&nb, sizeof(int) * (size_t)want) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        if (buf[dev] != nullptr) cudaFree(buf[dev]);
        buf[dev] = nb;
        cap[dev] = want;
    }
    return buf[dev];
}

}  // namespace

extern "C" {

int orlk_dyn_input(const float* obs, int64_t ld_obs, const float* act, int64_t ld_act, const float* mu, const float* sd, int S,
                   int O, int A, float* X, int64_t ldx, void* stream) {
    ORLK_REQUIRE(S > 0 && O > 0 && A > 0, "sizes");
    const int64_t n = (int64_t)S * (O + A);
    orlk::launch(k_dyn_input, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, obs, ld_obs, act, ld_act, mu, sd, S, O, A, X, ldx);
    return check_launch("k_dyn_input");
}

int orlk_gather_rows(const float* src, int64_t ld_src, int w, const int64_t* idx, int64_t idx_ld, int64_t r0, int E, int R,
                     float* dst, int64_t ld_dst, int64_t dst_es, void* stream) {
    ORLK_REQUIRE(E > 0 && R > 0 && w > 0, "sizes");
    const int64_t rows = (int64_t)E * R;
    orlk::launch(k_gather_rows, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, src, ld_src, w, idx, idx_ld, r0, E, R, dst, ld_dst,
                                                                             dst_es);
    return check_launch("k_gather_rows");
}

int orlk_sumsq_chunks(int64_t n) { return (int)((n + 4095) / 4096); }

int orlk_sumsq(const float* x, int64_t n, float scale, float* partial, void* stream) {
    ORLK_REQUIRE(n > 0, "sizes");
    orlk::launch(k_sumsq, orlk_sumsq_chunks(n), 256, 0, (cudaStream_t)stream, x, n, scale, partial);
    return check_launch("k_sumsq");
}

int orlk_dyn_nll_scratch_floats(int E, int Bn, int D) {
    return ((E * Bn + NLL_ROWS - 1) / NLL_ROWS) * (1 + 2 * D);
}

int orlk_dyn_nll(const float* out, const float* y, int E, int Bn, int D, const float* max_lv, const float* min_lv, float coef,
                 const float* decay_partials, int n_decay, float* dout, float* dmax, float* dmin, float* out_loss,
                 float* scratch, void* stream) {
    ORLK_REQUIRE(E > 0 && Bn > 0 && D > 0 && D <= 64, "sizes (D <= 64)");
    if (scratch != nullptr) {       // multi-CTA form; scratch holds orlk_dyn_nll_scratch_floats(E, Bn, D) floats
        const int n_rows = E * Bn, n_blocks = (n_rows + NLL_ROWS - 1) / NLL_ROWS;
        orlk::launch(k_dyn_nll_part, n_blocks, 256, 0, (cudaStream_t)stream, out, y, n_rows, Bn, D, max_lv, min_lv, dout, scratch);
        orlk::launch(k_dyn_nll_final, 1, 256, 0, (cudaStream_t)stream, (const float*)scratch, n_blocks, D, coef, max_lv, min_lv,
                     decay_partials, n_decay, dmax, dmin, out_loss);
        return check_launch("k_dyn_nll_final");
    }
    orlk::launch(k_dyn_nll, 1, 1024, 0, (cudaStream_t)stream, out, y, E, Bn, D, max_lv, min_lv, coef, decay_partials, n_decay, dout, dmax,
                                                   dmin, out_loss);
    return check_launch("k_dyn_nll");
}

int orlk_dyn_val_mse(const float* out, const float* y, int E, int Bn, int D, float* mse, void* stream) {
    ORLK_REQUIRE(E > 0 && Bn > 0 && D > 0, "sizes");
    orlk::launch(k_dyn_val_mse, E, 256, 0, (cudaStream_t)stream, out, y, Bn, D, mse);
    return check_launch("k_dyn_val_mse");
}

int orlk_dyn_step(const float* out, int E, int S, int D, const float* max_lv, const float* min_lv, const float* obs,
                  int64_t ld_obs, const double* noise, const int* midx, const float* noise32, const float* pick_u,
                  const int* elites, int n_elites, int term_kind, float penalty_coef, int uncertainty_mode, float* next_obs,
                  float* reward, float* raw_reward, float* penalty, unsigned char* terminal, void* stream) {
    ORLK_REQUIRE(E > 0 && S > 0 && D > 1 && D <= 64, "sizes (D <= 64)");
    ORLK_REQUIRE(term_kind >= 0 && term_kind <= 3, "termination kind");
    ORLK_REQUIRE(uncertainty_mode >= 0 && uncertainty_mode <= 2, "uncertainty mode");
    ORLK_REQUIRE(noise != nullptr || noise32 != nullptr, "noise");
    ORLK_REQUIRE(midx != nullptr || (pick_u != nullptr && elites != nullptr && n_elites > 0), "elite selection");
    orlk::launch(k_dyn_step, (S + 127) / 128, 128, 0, (cudaStream_t)stream, out, E, S, D, max_lv, min_lv, obs, ld_obs, noise, midx, noise32,
                                                                 pick_u, elites, n_elites, term_kind, penalty_coef, uncertainty_mode,
                                                                 next_obs, reward, raw_reward, penalty, terminal);
    return check_launch("k_dyn_step");
}

int orlk_compact_rows(const unsigned char* drop, int S, const float* src, int64_t ld_src, int w, float* dst, int64_t ld_dst,
                      int* count_out, void* stream) {
    ORLK_REQUIRE(S > 0 && w > 0, "sizes");
    cudaStream_t st = (cudaStream_t)stream;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone;
    const int n_blocks = (S + CMP_ROWS - 1) / CMP_ROWS;
    int* counts = (S > 4096 && !capturing) ? compact_scratch(n_blocks) : nullptr;    // (the scratch may need a cudaMalloc)
    if (counts == nullptr) {
        orlk::launch(k_compact_rows, 1, 1024, 0, st, drop, S, src, ld_src, w, dst, ld_dst, count_out);
        return check_launch("k_compact_rows");
    }
    orlk::launch(k_compact_count, n_blocks, CMP_ROWS, 0, st, drop, S, counts);
    orlk::launch(k_compact_scatter, n_blocks, CMP_ROWS, 0, st, drop, S, src, ld_src, w, dst, ld_dst, (const int*)counts, count_out);
    return check_launch("k_compact_scatter");
}

}  // extern "C"
