// Latency-oriented fp32 GEMM for the small-row passes (M of a few hundred: actor / target / policy-improvement chains).
//
// Those layers are ~17 MFLOP each: what they cost is the dependent chain  launch -> operand fetch -> k loop -> store.
// So this kernel (a) takes its problem list BY VALUE in the kernel parameters (no descriptor fetch from global memory),
// (b) cuts the output into 32 x 16 tiles so that a 256 x 256 layer spreads over 128 SMs, (c) fetches the whole k extent
// of both operand tiles (up to 256 k at a time) with one burst of 16-byte cp.async copies in the operands' OWN layout -
// no register staging, no transposition, (d) splits k over four thread groups that are summed through shared memory,
// and (e) applies the epilogue with all 256 threads on the reduced tile so C and CT both leave in coalesced rows.
// Same math and the same descriptor as orlk_gemm_grouped (ascending-k FFMA inside each k group, fixed-order group sum).
// Replaces nn.Linear forward / autograd dgrad / wgrad for small batches (nets/mlp.py:22,28).
#include <stdlib.h>
#include "orlk_common.cuh"
using namespace orlk;

namespace {

constexpr int TM = 32;            // tile rows
constexpr int TN = 16;            // tile columns
constexpr int KC = 256;           // k extent staged per pass
constexpr int KP = KC + 4;        // row pitch (floats) of a k-contiguous operand tile: 260 % 32 == 4 -> rows sit 16 bytes apart in the banks
constexpr int PADMN = 8;          // pad of an mn-contiguous tile row: pitch 40 / 24 floats keeps 16-byte alignment and makes the
                                  // tensor-core fragment loads (4 k-rows x 8 columns per instruction) conflict-free
constexpr int NTHR = 512;
constexpr int KG = 8;             // k groups of 64 threads; group g owns the g-th eighth of every staged k extent
constexpr int MAXP = 16;          // problems per launch (kernel-parameter space: 16 x 176 bytes)
static_assert(TM * TN == NTHR, "the epilogue gives every thread one element of the tile");

struct TinyArgs {
    OrlkGemmDesc d[MAXP];
    int n;
    int late_trigger;             // let the next kernel of the stream in after the k loop instead of at the start: its CTAs
                                  // then do not sit on this kernel's SMs while it stages and multiplies (measured: 306 ->
                                  // 295 us per CQL step); ORLK_TINY_LATE_TRIGGER=0 restores the early trigger
    int passes;                   // MMA variant: 3 = hi/lo split operands, three TF32 MMAs per product (fp32-grade); 1 = plain TF32
    unsigned long long* trace;    // profiling aid (orlk_tc_set_trace): 16 clock stamps per CTA, NULL in normal operation
};
#define TINY_STAMP(slot)                                                                                      \
    do {                                                                                                      \
        if (P.trace != nullptr && threadIdx.x == 0) P.trace[(int64_t)blockIdx.x * 16 + (slot)] = (unsigned long long)clock64(); \
    } while (0)

__device__ __forceinline__ void cp_async16(float* dst, const float* src, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
// Named barrier 1 + g over the 64 threads of k group g (literal ids, so that the kernel reserves 9 barriers, not 16).
__device__ __forceinline__ void group_sync(int g) {
    switch (g) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        case 3: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
        case 4: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
        case 5: asm volatile("bar.sync 6, 64;" ::: "memory"); break;
        case 6: asm volatile("bar.sync 7, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 8, 64;" ::: "memory"); break;
    }
}

// One k group (64 threads, tg = 0..63) stages ITS OWN share of an operand tile: rows [t0, t0+BT) x 4-k blocks
// [b_lo, b_hi) (relative to k0), 16-byte cp.async copies in the operand's own layout.  KCONT: operand(t,k) =
// base[t*ld + k] -> S[t][KP]; otherwise operand(t,k) = base[k*ld + t] -> S[k][BT].  Outside the matrix: zeros.
template <int BT, bool KCONT>
__device__ __forceinline__ void stage(float* S, const float* __restrict__ base, int64_t ld, bool vec, int t0, int T, int k0,
                                      int b_lo, int b_hi, int kend, int tg) {
    const int nb = b_hi - b_lo;
    if (nb <= 0) return;
    if (KCONT) {
        if (vec) {
            if (nb == 8) {                              // the common case (256 k per pass): shifts instead of divisions
#pragma unroll
                for (int i = 0; i < BT * 8 / 64; ++i) {
                    const int idx = tg + 64 * i;
                    const int r = idx >> 3, c = b_lo + (idx & 7);
                    const int t = t0 + r, k = k0 + 4 * c;
                    const int bytes = (t < T && k < kend) ? 4 * min(4, kend - k) : 0;
                    cp_async16(S + r * KP + 4 * c, base + (int64_t)min(t, T - 1) * ld + (bytes ? k : 0), bytes);
                }
            } else {
                for (int idx = tg; idx < BT * nb; idx += 64) {
                    const int r = idx / nb, c = b_lo + (idx - r * nb);
                    const int t = t0 + r, k = k0 + 4 * c;
                    const int bytes = (t < T && k < kend) ? 4 * min(4, kend - k) : 0;
                    cp_async16(S + r * KP + 4 * c, base + (int64_t)min(t, T - 1) * ld + (bytes ? k : 0), bytes);
                }
            }
        } else {
            for (int q = tg; q < BT * nb * 4; q += 64) {
                const int r = q / (nb * 4), kk = 4 * b_lo + (q - r * (nb * 4));
                const int t = t0 + r, k = k0 + kk;
                S[r * KP + kk] = (t < T && k < kend) ? __ldg(base + (int64_t)t * ld + k) : 0.f;
            }
        }
    } else {
        constexpr int C4 = BT / 4;
        if (vec) {
            for (int idx = tg; idx < 4 * nb * C4; idx += 64) {
                const int kr = idx / C4, c = idx - kr * C4;
                const int kk = 4 * b_lo + kr;
                const int k = k0 + kk, t = t0 + 4 * c;
                const int bytes = (k < kend && t < T) ? 4 * min(4, T - t) : 0;
                cp_async16(S + kk * (BT + PADMN) + 4 * c, base + (int64_t)min(k, kend - 1) * ld + (t < T ? t : 0), bytes);
            }
        } else {
            for (int q = tg; q < 4 * nb * BT; q += 64) {
                const int kr = q / BT, r = q - kr * BT;
                const int kk = 4 * b_lo + kr;
                const int k = k0 + kk, t = t0 + r;
                S[kk * (BT + PADMN) + r] = (t < T && k < kend) ? __ldg(base + (int64_t)k * ld + t) : 0.f;
            }
        }
    }
}

// A_KC: A(m,k) = A[m*lda + k] (a_layout 0); B_KC: B(k,n) = B[n*ldb + k] (b_layout 1).
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x) { return __float_as_uint(x - __uint_as_float(tf32_hi(x))); }

// MMA = false: fp32 FFMA micro-kernel (the `fp32` mode, and launches that want row / column sums).
// MMA = true : warp-level TF32 tensor-core MMAs (mma.sync m16n8k8) on fragments read straight from the staged tiles, with
//   the same 3xTF32 operand split as the big-pass kernel (P.passes == 3) or single-pass TF32 (P.passes == 1); every one
//   of the 16 warps owns the whole 32 x 16 tile for 1/16 of k.  The k loop drops from 1.4 us (LDS-issue bound) to ~0.2 us.
template <bool A_KC, bool B_KC, bool MMA>
__global__ void __launch_bounds__(NTHR, 2)
k_tiny_gemm(const __grid_constant__ TinyArgs P) {
    TINY_STAMP(0);
    extern __shared__ float4 smem_f4[];
    float* As = reinterpret_cast<float*>(smem_f4);                  // A_KC ? [TM][KP] : [KC][TM]
    float* Bs = As + (A_KC ? TM * KP : KC * (TM + PADMN));          // B_KC ? [TN][KP] : [KC][TN + 8]

    const int tid = threadIdx.x;
    int p = 0;
    while (p + 1 < P.n && P.d[p + 1].tile_start <= (int)blockIdx.x) ++p;
    const OrlkGemmDesc& d = P.d[p];
    const int t = blockIdx.x - d.tile_start;
    const int tm = t / d.tiles_n, tn = t - tm * d.tiles_n;
    const int m0 = tm * TM, n0 = tn * TN;
    const int M = d.M, N = d.N, K = d.K;
    const bool vecA = aligned16(d.A) && (d.lda % 4) == 0, vecB = aligned16(d.B) && (d.ldb % 4) == 0;
    if (P.late_trigger) orlk::pdl_wait(); else orlk::pdl_enter();      // nothing above touched global data
    TINY_STAMP(1);

    // this thread's element of the tile in the epilogue: fetch its bias / aux operands now, off the critical path
    const int er = tid / TN, ec = tid - er * TN;
    const int em = m0 + er, en = n0 + ec;
    const bool e_ok = em < M && en < N;
    const float e_bias = (e_ok && d.bias != nullptr) ? __ldg(d.bias + en) : 0.f;
    const float e_aux = (e_ok && d.aux != nullptr) ? __ldg(d.aux + (int64_t)em * d.ldaux + en) : 0.f;

    const int kg = tid >> 6, tg = tid & 63;
    const int tx = tg & 7, ty = tg >> 3;                // 8 x 8 threads per k group, micro-tile 4 (m) x 2 (n)
    float acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
    float cacc[2][2][4];                                // MMA variant: 2 x 2 m16n8 accumulator tiles per warp
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) cacc[mt][nt][0] = cacc[mt][nt][1] = cacc[mt][nt][2] = cacc[mt][nt][3] = 0.f;
    float rs[4] = {0.f, 0.f, 0.f, 0.f}, cs[2] = {0.f, 0.f};
    const bool do_rs = d.rowsum != nullptr && tn == 0;
    const bool do_cs = d.colsum != nullptr && tm == 0;

    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = min(KC, K - k0);
        if (k0 > 0) __syncthreads();                    // the previous pass is done with the tiles
        // group g owns a contiguous share of the staged k extent: whole 4-k blocks (FFMA) or whole 8-k MMA steps
        // (a ragged tail is zero-filled by the staging code)
        int b_lo, b_hi;
        if (MMA) {
            const int nstep = (kc + 7) >> 3, per8 = (nstep + KG - 1) / KG;
            const int s_lo = min(nstep, kg * per8), s_hi = min(nstep, s_lo + per8);
            b_lo = 2 * s_lo;
            b_hi = 2 * s_hi;
        } else {
            const int nblk = (kc + 3) >> 2, per = (nblk + KG - 1) / KG;
            b_lo = min(nblk, kg * per);
            b_hi = min(nblk, b_lo + per);
        }
        // every k group fetches and waits for its own eighth of the two tiles: the groups never wait for each other
        stage<TM, A_KC>(As, d.A, d.lda, vecA, m0, M, k0, b_lo, b_hi, K, tg);
        stage<TN, B_KC>(Bs, d.B, d.ldb, vecB, n0, N, k0, b_lo, b_hi, K, tg);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (k0 == 0) TINY_STAMP(2);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        group_sync(kg);                                 // named barrier of this group's two warps
        if (k0 == 0) TINY_STAMP(3);
        if (MMA) {
            // the two warps of the group split its MMA steps; fragment element positions per PTX m16n8k8 (.tf32):
            // A: (row gid | gid+8, k tig | tig+4)   B: (k tig | tig+4, col gid)   C: (row gid | gid+8, col 2 tig | 2 tig+1)
            const int lane = tid & 31, wsub = (tid >> 5) & 1, gid = lane >> 2, tig = lane & 3;
            const int s_lo = b_lo >> 1, s_hi = b_hi >> 1;
            const int half = (s_hi - s_lo + 1) >> 1;
            const int w_lo = s_lo + wsub * half, w_hi = min(s_hi, w_lo + half);
            const bool split3 = P.passes == 3;
            for (int st = w_lo; st < w_hi; ++st) {
                const int k = 8 * st;
                float af[2][4], bf[2][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int r = mt * 16 + gid;
                    if (A_KC) {
                        af[mt][0] = As[r * KP + k + tig];
                        af[mt][1] = As[(r + 8) * KP + k + tig];
                        af[mt][2] = As[r * KP + k + tig + 4];
                        af[mt][3] = As[(r + 8) * KP + k + tig + 4];
                    } else {
                        af[mt][0] = As[(k + tig) * (TM + PADMN) + r];
                        af[mt][1] = As[(k + tig) * (TM + PADMN) + r + 8];
                        af[mt][2] = As[(k + tig + 4) * (TM + PADMN) + r];
                        af[mt][3] = As[(k + tig + 4) * (TM + PADMN) + r + 8];
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int c = nt * 8 + gid;
                    if (B_KC) {
                        bf[nt][0] = Bs[c * KP + k + tig];
                        bf[nt][1] = Bs[c * KP + k + tig + 4];
                    } else {
                        bf[nt][0] = Bs[(k + tig) * (TN + PADMN) + c];
                        bf[nt][1] = Bs[(k + tig + 4) * (TN + PADMN) + c];
                    }
                }
                uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ah[mt][i] = tf32_hi(af[mt][i]);
                        al[mt][i] = tf32_lo(af[mt][i]);
                    }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        bh[nt][i] = tf32_hi(bf[nt][i]);
                        bl[nt][i] = tf32_lo(bf[nt][i]);
                    }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        mma_tf32(cacc[mt][nt], ah[mt], bh[nt][0], bh[nt][1]);
                        if (split3) {
                            mma_tf32(cacc[mt][nt], al[mt], bh[nt][0], bh[nt][1]);
                            mma_tf32(cacc[mt][nt], ah[mt], bl[nt][0], bl[nt][1]);
                        }
                    }
            }
        }
#pragma unroll 2
        for (int blk = b_lo; !MMA && blk < b_hi; ++blk) {
            const int k = 4 * blk;
            float a[4][4], b[4][2];                     // a[i][e] = A(m_i, k+e), b[e][j] = B(k+e, n_j)
            if (A_KC) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {           // rows ty + 8i: the 4 rows a warp touches are 16 bytes apart in the banks
                    const float4 v = *reinterpret_cast<const float4*>(As + (ty + 8 * i) * KP + k);
                    a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {           // rows 4ty .. 4ty+3
                    const float4 v = *reinterpret_cast<const float4*>(As + (k + e) * (TM + PADMN) + 4 * ty);
                    a[0][e] = v.x; a[1][e] = v.y; a[2][e] = v.z; a[3][e] = v.w;
                }
            }
            if (B_KC) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {           // columns tx + 8j
                    const float4 v = *reinterpret_cast<const float4*>(Bs + (tx + 8 * j) * KP + k);
                    b[0][j] = v.x; b[1][j] = v.y; b[2][j] = v.z; b[3][j] = v.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {           // columns 2tx, 2tx+1
                    const float2 v = *reinterpret_cast<const float2*>(Bs + (k + e) * (TN + PADMN) + 2 * tx);
                    b[e][0] = v.x; b[e][1] = v.y;
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(a[i][e], b[e][0], acc[i][0]);
                    acc[i][1] = fmaf(a[i][e], b[e][1], acc[i][1]);
                }
            if (do_rs) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int i = 0; i < 4; ++i) rs[i] += a[i][e];
            }
            if (do_cs) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { cs[0] += b[e][0]; cs[1] += b[e][1]; }
            }
        }
    }

    // ---- k-group partial tiles -> shared memory (the operand tiles are dead), then every thread finishes one element
    if (P.late_trigger == 1) orlk::pdl_trigger();
    TINY_STAMP(4);
    __syncthreads();
    constexpr int NRED = MMA ? NTHR / 32 : KG;          // partial tiles: one per warp (MMA) or per k group (FFMA)
    float* red = As;                                    // [NRED][TM][TN + 1]
    float* rsum = red + NRED * TM * (TN + 1);           // [KG][TM]
    float* csum = rsum + KG * TM;                       // [KG][TN]
    if (MMA) {
        const int lane = tid & 31, wi = tid >> 5, gid = lane >> 2, tig = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float* r0 = red + (wi * TM + mt * 16 + gid) * (TN + 1) + nt * 8 + 2 * tig;
                r0[0] = cacc[mt][nt][0];
                r0[1] = cacc[mt][nt][1];
                r0[8 * (TN + 1)] = cacc[mt][nt][2];
                r0[8 * (TN + 1) + 1] = cacc[mt][nt][3];
            }
    }
#pragma unroll
    for (int i = 0; !MMA && i < 4; ++i) {
        const int r = A_KC ? ty + 8 * i : 4 * ty + i;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = B_KC ? tx + 8 * j : 2 * tx + j;
            red[(kg * TM + r) * (TN + 1) + c] = acc[i][j];
        }
        if (do_rs && tx == 0) rsum[kg * TM + r] = rs[i];
    }
    if (!MMA && do_cs && ty == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) csum[kg * TN + (B_KC ? tx + 8 * j : 2 * tx + j)] = cs[j];
    }
    __syncthreads();
    TINY_STAMP(5);
    if (P.late_trigger == 2) orlk::pdl_trigger();

    const int slot = d.split_base;
    const int epi = d.epi;
    // consecutive threads -> consecutive n: row-major C / C2 leave in 64-byte rows
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < NRED; ++g) v += red[(g * TM + er) * (TN + 1) + ec];    // fixed order: bit-reproducible
    if (e_ok) {
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
        if (d.C != nullptr) d.C[(int64_t)slot * d.c_split_stride + (int64_t)em * d.ldc + en] = v;
    }
    if (d.CT != nullptr) {
        __syncthreads();                                // everyone has read its partials
        red[er * (TN + 1) + ec] = v;
        __syncthreads();
        // consecutive threads -> consecutive m: CT leaves in 128-byte rows
        const int c = tid / TM, r = tid - c * TM;
        if (m0 + r < M && n0 + c < N) d.CT[(int64_t)(n0 + c) * d.ldct + m0 + r] = red[r * (TN + 1) + c];
    }
    TINY_STAMP(6);
    if (!MMA && do_rs && tid < TM && m0 + tid < M) {
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < KG; ++g) sum += rsum[g * TM + tid];
        d.rowsum[(int64_t)slot * d.sum_split_stride + m0 + tid] = sum;
    }
    if (!MMA && do_cs && tid < TN && n0 + tid < N) {
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < KG; ++g) sum += csum[g * TN + tid];
        d.colsum[(int64_t)slot * d.sum_split_stride + n0 + tid] = sum;
    }
}

template <bool A_KC, bool B_KC>
constexpr size_t tiny_smem() {
    return sizeof(float) * ((A_KC ? TM * KP : KC * (TM + PADMN)) + (B_KC ? TN * KP : KC * (TN + PADMN)));
}

template <bool A_KC, bool B_KC>
int tiny_launch(const TinyArgs& args, int total_tiles, bool mma, cudaStream_t s) {
    static int pdl = -1;
    if (pdl < 0) { const char* e = getenv("ORLK_PDL_TINY"); pdl = (e && e[0] == '0') ? 0 : 1; }
    if (mma) orlk::launch_opt(pdl != 0, k_tiny_gemm<A_KC, B_KC, true>, total_tiles, NTHR, tiny_smem<A_KC, B_KC>(), s, args);
    else orlk::launch_opt(pdl != 0, k_tiny_gemm<A_KC, B_KC, false>, total_tiles, NTHR, tiny_smem<A_KC, B_KC>(), s, args);
    return check_launch("k_tiny_gemm");
}

template <bool A_KC, bool B_KC>
int tiny_attr() {
    int rc = check(cudaFuncSetAttribute(k_tiny_gemm<A_KC, B_KC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)tiny_smem<A_KC, B_KC>()), "tiny smem attr");
    if (rc) return rc;
    return check(cudaFuncSetAttribute(k_tiny_gemm<A_KC, B_KC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)tiny_smem<A_KC, B_KC>()), "tiny smem attr");
}

}  // namespace

// Set the shared-memory opt-in of all layout / arithmetic variants once, outside stream capture.
extern "C" int orlk_gemm_tiny_init(void) {
    int rc = tiny_attr<true, true>();
    if (rc) return rc;
    rc = tiny_attr<true, false>();
    if (rc) return rc;
    rc = tiny_attr<false, true>();
    if (rc) return rc;
    return tiny_attr<false, false>();
}

// descs_host: HOST array (copied into the kernel parameters); tiles are 32 x 16, k_splits must be 1.
// passes: 0 = fp32 FFMA, 3 = 3xTF32 tensor-core MMAs (fp32-grade), 1 = single-pass TF32.  Launches that ask for row or
// column sums always take the FFMA kernel.
extern "C" int orlk_gemm_tiny(const OrlkGemmDesc* descs_host, int n_descs, int total_tiles, int a_layout, int b_layout,
                              int passes, void* stream) {
    ORLK_REQUIRE(descs_host != nullptr && n_descs > 0 && n_descs <= MAXP, "1..16 problems per launch");
    ORLK_REQUIRE(total_tiles > 0, "total_tiles");
    ORLK_REQUIRE(passes == 0 || passes == 1 || passes == 3, "passes must be 0, 1 or 3");
    TinyArgs args;
    args.n = n_descs;
    args.passes = passes;
    { static int lt = -1; if (lt < 0) { const char* e = getenv("ORLK_TINY_LATE_TRIGGER"); lt = e ? atoi(e) : 1; } args.late_trigger = lt; }
    args.trace = orlk::trace_buffer();
    bool mma = passes != 0;
    for (int i = 0; i < n_descs; ++i) {
        const OrlkGemmDesc& d = descs_host[i];
        ORLK_REQUIRE(d.k_splits <= 1, "the small-row kernel does not split k");
        ORLK_REQUIRE(d.a_layout == a_layout && d.b_layout == b_layout, "operand layouts must match the launch");
        ORLK_REQUIRE(d.tiles_m == (d.M + TM - 1) / TM && d.tiles_n == (d.N + TN - 1) / TN, "tiles must be 32 x 16");
        if (d.rowsum != nullptr || d.colsum != nullptr) mma = false;
        args.d[i] = d;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const bool a_kc = a_layout == 0, b_kc = b_layout == 1;
    if (a_kc && b_kc) return tiny_launch<true, true>(args, total_tiles, mma, s);
    if (a_kc) return tiny_launch<true, false>(args, total_tiles, mma, s);
    if (b_kc) return tiny_launch<false, true>(args, total_tiles, mma, s);
    return tiny_launch<false, false>(args, total_tiles, mma, s);
}
