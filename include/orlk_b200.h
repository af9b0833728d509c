/*
 * orlk_b200.h -- C ABI of the B200-native engine for OfflineRL-Kit's offline
 * actor-critic gradient step (liborlk_b200.so, built from offlinerl-kit_b200/csrc).
 *
 * The reference (zhaoyizhou1123/OfflineRL-Kit) is pure Python on PyTorch: it has
 * no FFI or operator registry.  The "reference interface each entry point
 * replaces" is therefore the Python call site whose ATen/cuBLAS work the entry
 * point takes over; those are cited per function as  file:line  relative to the
 * reference tree.  INTEGRATION.md shows the ctypes binding a maintainer would
 * add on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the name ends in _host (pinned host memory);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - every function returns 0 on success, otherwise a non-zero code (a
 *     cudaError_t, or ORLK_ERR_*); orlk_last_error() gives the message.  Nothing
 *     throws and nothing falls back to the CPU;
 *   - all launchers are asynchronous on `stream` and CUDA-graph capturable.
 */
#ifndef ORLK_B200_H
#define ORLK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORLK_ABI_VERSION 30
#define ORLK_ERR_BAD_ARG 10001
#define ORLK_ERR_UNSUPPORTED 10002

/* ------------------------------------------------------------------ runtime */
int orlk_abi_version(void);
const char* orlk_last_error(void);
/* sizeof() of the descriptor structs below, so a binding can verify its mirror */
int orlk_sizeof_gemm_desc(void);
int orlk_sizeof_adam_desc(void);
int orlk_sizeof_adam_group(void);
int orlk_sizeof_concat_seg(void);
/* device properties: out[0]=SM count, out[1]=cc major, out[2]=cc minor, out[3]=max smem/block optin */
int orlk_device_info(int device, int* out4);

/* CUDA-graph capture of a launch sequence on `stream` (replaces the ~4 300
 * per-step ATen launches of one policy.learn). */
int orlk_graph_begin(void* stream);
int orlk_graph_end(void* stream, void** graph_exec_out);
int orlk_graph_launch(void* graph_exec, void* stream);
int orlk_capture_status(void* stream); /* diagnostic (ORLK_GRAPH_DEBUG): 0 not capturing, 1 capturing, 2 invalidated; returned as the value */
int orlk_graph_launch_sync(void* graph_exec, void* stream); /* launch, then wait for the stream */
/* launch, then wait only for `ev`, which a node inside the graph records (orlk_event_record_external under capture): the
 * step's loss block has reached pinned host memory - policy.learn's Dict[str, float] contract (base_policy.py:25-26) - while
 * the rest of the step (the critics' backward pass) is still running and the caller already draws the next batch */
int orlk_graph_launch_wait_event(void* graph_exec, void* stream, void* ev);
int orlk_event_record_external(void* ev, void* stream); /* under capture: an external event-record node; else cudaEventRecord */
int orlk_graph_destroy(void* graph_exec);
int orlk_stream_sync(void* stream);
/* side streams + ordering events: independent launches of a step (e.g. the weight gradients of different layers)
 * are forked onto side streams inside the captured graph and joined before the optimiser */
int orlk_stream_create(void** stream_out);
int orlk_stream_destroy(void* stream);
int orlk_stream_wait_event(void* stream, void* ev);
int orlk_event_create_notiming(void** ev_out);
int orlk_memcpy_h2d_async(void* dst, const void* src_host, size_t bytes, void* stream);
int orlk_memcpy_d2h_async(void* dst_host, const void* src, size_t bytes, void* stream);
int orlk_memcpy_d2d_async(void* dst, const void* src, size_t bytes, void* stream);
int orlk_memset_async(void* dst, int value, size_t bytes, void* stream);
/* event timing on `stream`: returns elapsed ms between two recorded events */
int orlk_event_create(void** ev_out);
int orlk_event_record(void* ev, void* stream);
int orlk_event_sync(void* ev);
int orlk_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms_out);
int orlk_event_destroy(void* ev);

/* ------------------------------------------------------------------- replay */
/* ReplayBuffer device mirror: one row-major table, row = [obs | next_obs | act | rew | term | pad],
 * row_w floats per row (multiple of 4).  Replaces the five host arrays of buffer/buffer.py:26-30. */
int orlk_replay_pack(const float* obs, const float* next_obs, const float* act, const float* rew, const float* term,
                     int64_t n, int obs_dim, int act_dim, float* table, int row_w, int64_t row_offset, void* stream);
/* ReplayBuffer.sample gather (buffer/buffer.py:100-106): bit-exact copy of rows idx[0..n) into
 *   obs2  [2n, obs_dim]  rows [0,n) = observations, rows [n,2n) = next_observations
 *   act   [n, act_dim], rew [n], term [n]. */
int orlk_replay_gather(const float* table, int64_t n_rows, int row_w, int obs_dim, int act_dim, const int64_t* idx,
                       int n, float* obs2, float* act, float* rew, float* term, void* stream);

/* The same gather with separate destinations for observations and next_observations, so that the draws of several
 * buffers land in consecutive row blocks of one batch: the real + model-buffer mix of the model-based policies
 * (mopo.py:81-84, combo.py:110-112 -- there a torch.cat of two sampled dicts). */
int orlk_replay_gather_into(const float* table, int64_t n_rows, int row_w, int obs_dim, int act_dim, const int64_t* idx,
                            int n, float* obs, float* next_obs, float* act, float* rew, float* term, void* stream);

/* The whole of ReplayBuffer.sample (buffer/buffer.py:97-106) in one host call: waits for slot_event if event_armed
 * (the previous upload out of this pinned slot), copies idx_host[0..n) into idx_pinned, uploads it to idx_dev,
 * records slot_event and launches orlk_replay_gather. */
int orlk_replay_sample(const float* table, int64_t n_rows, int row_w, int obs_dim, int act_dim, const int64_t* idx_host,
                       int64_t* idx_pinned, int64_t* idx_dev, void* slot_event, int event_armed, int n, float* obs2,
                       float* act, float* rew, float* term, void* stream);

/* --------------------------------------------------------------- dense GEMM */
/* One problem of a grouped fp32 GEMM launch:  C[M,N] = epi( sum_k A(m,k) * B(k,n) ).
 * Replaces nn.Linear / einsum forward, dgrad and wgrad GEMMs (nets/mlp.py:22,28;
 * nets/ensemble_linear.py:35,37 and their autograd backward). */
enum {
    ORLK_EPI_NONE = 0,      /* C = acc (+ bias[n])                                    */
    ORLK_EPI_RELU = 1,      /* C = relu(acc + bias[n])                                */
    ORLK_EPI_RELU_MASK = 2, /* C = acc * (aux(m,n) > 0)        ReLU backward          */
    ORLK_EPI_SWISH = 3,     /* z = acc + bias; C = z*sigmoid(z); C2 = z               */
    ORLK_EPI_DSWISH = 4     /* C = acc * swish'(aux(m,n))      Swish backward, aux=z  */
};
/* tile configurations: 128x128x16, 64x64x16, 32x32x32, and 32x32x256 with 4 k-parallel thread groups (small-M layers) */
enum { ORLK_CFG_BIG = 0, ORLK_CFG_MID = 1, ORLK_CFG_SMALL = 2, ORLK_CFG_KPAR = 3, ORLK_CFG_TINY = 4 /* orlk_gemm_tiny only */ };

typedef struct OrlkGemmDesc {
    const float* A;
    const float* B;
    float* C;
    float* C2;         /* second output (ORLK_EPI_SWISH) or NULL                       */
    const float* bias; /* [N] or NULL                                                  */
    const float* aux;  /* epilogue operand indexed [m*ldaux + n] or NULL               */
    float* rowsum;     /* optional [M]: sum_k A(m,k), written by the n-tile-0 CTAs     */
    float* colsum;     /* optional [N]: sum_k B(k,n), written by the m-tile-0 CTAs     */
    float* CT;         /* optional transposed copy of the output: CT[n*ldct + m]       */
    int64_t lda, ldb, ldc, ldaux, ldct;
    int64_t c_split_stride;   /* split s writes C + (split_base+s)*c_split_stride      */
    int64_t sum_split_stride; /* same for rowsum / colsum                              */
    int32_t M, N, K;
    int32_t a_layout; /* 0: A(m,k)=A[m*lda+k]   1: A(m,k)=A[k*lda+m]                  */
    int32_t b_layout; /* 0: B(k,n)=B[k*ldb+n]   1: B(k,n)=B[n*ldb+k]                  */
    int32_t epi;
    int32_t k_splits; /* >=1 ; split s covers k in [s*k_chunk, min(K,(s+1)*k_chunk))  */
    int32_t k_chunk;  /* multiple of the config's BK                                   */
    int32_t split_base;
    int32_t tile_start; /* first linear tile id of this problem in the launch          */
    int32_t tiles_m, tiles_n;
} OrlkGemmDesc;

int orlk_gemm_init(void); /* once per process, outside stream capture (shared-memory opt-in) */
/* a_layout / b_layout: the operand layouts shared by ALL problems of the launch (the kernel is specialised on them;
 * the per-problem fields of the descriptors must agree). */
int orlk_gemm_grouped(const OrlkGemmDesc* descs_dev, int n_descs, int total_tiles, int cfg, int a_layout, int b_layout,
                      void* stream);
/* Small-row variant (M of a few hundred, latency-bound layers): same descriptor and math, 32 x 16 output tiles
 * (ORLK_CFG_TINY), 512 threads in 8 k-groups that each fetch their own share of the operand tiles with cp.async and sync
 * only among themselves.  passes: 0 = fp32 FFMA, 3 = 3xTF32 warp-level tensor-core MMAs on hi/lo split operands
 * (fp32-grade), 1 = single-pass TF32; launches that ask for row / column sums always take the FFMA kernel.
 * descs_host is a HOST array of 1..16 problems with k_splits == 1 (it travels in the kernel parameters). */
int orlk_gemm_tiny_init(void); /* once per process, outside stream capture */
int orlk_gemm_tiny(const OrlkGemmDesc* descs_host, int n_descs, int total_tiles, int a_layout, int b_layout, int passes,
                   void* stream);

/* Fused small-row layer chain: ONE launch runs n_stages dependent GEMM stages (stage s+1 multiplies what stage s
 * wrote: A of s+1 == C of s) for n_chains independent chains with the same M (twin critics).  A thread-block cluster of
 * 8 CTAs owns a 16-row strip for the whole chain (one 16 x 32 column tile per CTA and stage); between stages: one
 * hardware cluster barrier, then every CTA re-reads the finished 16 KB strip from L2 (cp.async.cg).  Arithmetic as
 * orlk_gemm_tiny's tensor-core variant (passes 1 or 3; passes_stage0 for the first stage, which may see raw
 * observations).  descs_host[chain * n_stages + stage] is a HOST array of at most 24 descriptors: a_layout 0, K <= 256,
 * N <= 256, no split-K / sums / transposed copy.
 * Replaces the per-layer launches of an MLP forward (nets/mlp.py:22,28) or its autograd input-gradient pass. */
int orlk_gemm_chain_init(void); /* once per process, outside stream capture */
int orlk_gemm_chain(const OrlkGemmDesc* descs_host, int n_chains, int n_stages, int passes, int passes_stage0, void* stream);

/* Tensor-core GEMM (tcgen05.mma kind::tf32, TMEM accumulators, TMA operand ring) for the wide hidden layers:
 *   C[g][m][n] = epi( sum_k A[g][m][k] * B[g][n][k] ),   A and B row-major with k contiguous, N <= 256, N % 16 == 0.
 * a_gs == 0 shares one A between all groups (twin critics on the same batch).  A needs 16-byte aligned rows (TMA);
 * B may have any row pitch when K <= 32 (the obs+act wide first layer: its single B tile is then staged by the
 * kernel's own warps instead of TMA).
 * passes = 1: single TF32 MMA per product ("fast" mode); passes = 3: hi/lo operand split in shared memory and
 * three MMAs per product (fp32-grade, the parity mode).  Outputs (each optional): row-major C (per k-split slot),
 * transposed CT[n][m], rowsum[m] = sum_k A[g][m][k] (bias gradients; also per k-split slot).  The struct is read
 * on the HOST (it is not a device pointer).  Same reference call sites as orlk_gemm_grouped. */
typedef struct OrlkTcGemm {
    const float* A; int64_t lda, a_gs;
    const float* B; int64_t ldb, b_gs;
    float* C; int64_t ldc, c_gs, c_split_stride;
    float* CT; int64_t ldct, ct_gs;
    const float* bias; int64_t bias_gs;
    const float* aux; int64_t ldaux, aux_gs;
    float* rowsum; int64_t rowsum_gs, rowsum_split_stride;
    int32_t M, N, K, G;
    int32_t epi;      /* ORLK_EPI_NONE | ORLK_EPI_RELU | ORLK_EPI_RELU_MASK | ORLK_EPI_SWISH (C only, no pre-activation copy) */
    int32_t k_splits; /* as returned by orlk_tc_effective_splits */
    int32_t passes;   /* 1 or 3 */
    int32_t n_tile;   /* output columns per CTA (multiple of 16 dividing N); 0 = N.  Small-M layers use 32 so that
                         (M/128) x (N/32) CTAs share the work */
    /* Rank-1 operand generator (both NULL = off).  When set, the kernel does not multiply A itself but
     *   A'[g][m][k] = gen_row[g][m] * gen_col[g][k] * (A[g][m][k] > 0),
     * built in shared memory from the A tile TMA just delivered.  This is the gradient that flows back through a
     * scalar head into the last ReLU layer: dZ[m][k] = dq[m] * w_head[k] * relu'(H[m][k]) (dgrad: A = H, gen_row = dq,
     * gen_col = w_head) and its transpose (wgrad: A = H^T, gen_row = w_head, gen_col = dq), so dZ never exists in
     * global memory (autograd of modules/critic_module.py:25-33's last Linear).  Requires K % 4 == 0. */
    const float* gen_row; int64_t gen_row_gs;
    const float* gen_col; int64_t gen_col_gs;
    /* MN-major operands: a_mn != 0 means A is stored [g][k][m] (m contiguous, lda = pitch of a k row) instead of
     * [g][m][k]; likewise b_mn for B as [g][k][n].  This is how the weight gradient dW[o][i] = sum_m dZ[m][o] H[m][i]
     * reads the row-major activations / gradients directly (no transposed copies).  b_mn needs n_tile % 32 == 0. */
    int32_t a_mn, b_mn;
} OrlkTcGemm;
int orlk_tc_init(void);
int orlk_tc_gemm(const OrlkTcGemm* params_host, void* stream);
int orlk_tc_effective_splits(int K, int want);
/* Profiling aid: when dev_buf is non-NULL every CTA of later orlk_tc_gemm launches writes 16 %globaltimer stamps
 * (ns) to dev_buf[16 * blockIdx.x ..]: 0 start, 1 predecessor complete, 2 prologue done, 3 first tile landed,
 * 4 first MMA issued, 5 last MMA issued, 6 accumulator ready, 7 epilogue stores issued, 8..15 k-slab 0..7 landed.  NULL turns it off. */
int orlk_tc_set_trace(void* dev_buf);
int orlk_sizeof_tc_gemm(void);

/* Fused critic forward pass: ONE launch evaluates, for every member g (twin critics),
 *   H_0 = relu(X W_0^T + b_0), H_l = relu(H_{l-1} W_l^T + b_l) (l < n_hidden), out[g][m] = H_last[m] . head_w[g] + head_b[g]
 * on the tensor cores (3xTF32, fp32-grade).  A CTA owns a 128-row strip of one member for the whole pass: accumulators
 * ping-pong in tensor memory, layer l's epilogue writes the A operand tiles of layer l+1 straight into shared memory and
 * stores H_l (needed by the backward pass; H[] all NULL = not stored) with TMA from the same tiles; the head is a dot
 * product in the last epilogue.  Up to two such passes ("jobs": the online critics on the 7936-row batch and the target
 * critics on the next-state rows) share one launch, each with its own rows, weights and outputs.
 * Shapes: hidden widths all N (32 or a multiple of 64, <= 256), K0 <= 32 input columns, X [M][ldx] shared by all members
 * (16-byte aligned rows), 'oi' weights W_l [N][N] (l >= 1) with one member stride `gs` for every parameter tensor, Wlo[l] =
 * W[l] - trunc_tf32(W[l]); W0pad / W0pad_lo = the first layer's [N][K0] weights zero-padded to [G][N][32] and their lo
 * words (all three kept by orlk_fused_prep).  Replaces the per-layer launches + head of
 * modules/critic_module.py:25-33 / nets/mlp.py:22-28 on CQL's critic batch and target rows
 * (policy/model_free/cql.py:108-160).  The structs are read on the HOST. */
#define ORLK_FUSED_MAX_LAYERS 4
/* CTA pairs (clusters of two; needs N % 64 == 0): M = 256 MMAs over two 128-row strips, each SM loads half of every weight
 * slab.  Halves the weight bytes per SM (the L2 -> SM path of a TPC bounds the single-CTA kernel when both SMs of a TPC
 * hold a strip), costs ~2 us of pair synchronisation per pass: pays from a few dozen strips on. */
#define ORLK_FUSED_PAIRS 1
typedef struct OrlkFusedFwd {
    const float* X; int64_t ldx;
    const float* W0pad; const float* W0pad_lo;    /* [G][N][32] */
    const float* W[ORLK_FUSED_MAX_LAYERS];        /* [0] unused */
    const float* Wlo[ORLK_FUSED_MAX_LAYERS];      /* [0] unused */
    const float* bias[ORLK_FUSED_MAX_LAYERS];
    float* H[ORLK_FUSED_MAX_LAYERS];              /* [G][M][N] each, or all NULL */
    int64_t gs, h_gs;                             /* member strides (floats) of the parameters / of H */
    const float* head_w; const float* head_b;
    float* out; int64_t out_gs;                   /* [G][M] */
    uint32_t* relu_bits;                          /* optional [n_hidden][G][8][M]: bit j of word [c][m] = (H_l[g][m][32c+j] > 0), for
                                                     orlk_critic_bwd_fused */
    int32_t M, N, K0, G, n_hidden;
    int32_t flags;                                /* ORLK_FUSED_PAIRS (of jobs[0]): CTA pairs, tcgen05.mma.cta_group::2 */
} OrlkFusedFwd;
int orlk_fused_init(void); /* once per process, outside stream capture */
int orlk_critic_fwd_fused(const OrlkFusedFwd* jobs_host, int n_jobs, void* stream);
int orlk_sizeof_fused_fwd(void);
/* Fused input-gradient chain of the same stack behind its scalar head, ONE launch:
 *   dZ_{L-1}[m][k] = dq[g][m] * head_w[g][k] * relu'(H_{L-1})   (generated in shared memory, never stored)
 *   dZ_{l-1} = (dZ_l W_l) * relu'(H_{l-1}),  l = L-1 .. 1        (stored: operands of the weight gradients)
 * with relu' read from the decision bits orlk_critic_fwd_fused left (relu_bits), WT[l] = W_l^T [N (in)][N (out)] (the
 * transposed copies the Adam kernel keeps) and WTlo[l] their lo words (orlk_fused_prep on that arena); member stride gs.
 * Replaces autograd's backward through modules/critic_module.py:25-33 for the hidden activations (the two dgrad
 * launches + the head's backward of CQL's critic update, policy/model_free/cql.py:190-205).  Host struct. */
typedef struct OrlkFusedBwd {
    const float* dq; int64_t dq_gs;               /* [G][M] */
    const float* head_w;                          /* [G][N], member stride gs */
    const uint32_t* relu_bits;                    /* [n_hidden][G][8][M] */
    const float* WT[ORLK_FUSED_MAX_LAYERS];       /* [0] unused */
    const float* WTlo[ORLK_FUSED_MAX_LAYERS];
    float* dZ[ORLK_FUSED_MAX_LAYERS];             /* dZ[l], l = 0 .. n_hidden-2: [G][M][N] */
    int64_t gs, dz_gs;
    int32_t M, N, G, n_hidden;
    int32_t flags, pad_;                          /* ORLK_FUSED_PAIRS */
} OrlkFusedBwd;
int orlk_critic_bwd_fused(const OrlkFusedBwd* params_host, void* stream);
int orlk_sizeof_fused_bwd(void);

/* Derived operand copies of a parameter arena for the fused passes, one launch per step:
 *   dst_lo[i] = src[i] - trunc_tf32(src[i]) for the whole arena (n floats, 16-byte aligned), and - when W0 != NULL -
 *   w0pad[0][g][o][k] = W0[g*gs + o*K0 + k] (k < K0, else 0), w0pad[1] = its lo words  (W0 inside [src, src + n)). */
int orlk_fused_prep(const float* src, float* dst_lo, int64_t n, const float* W0, int64_t gs, int N, int K0, int G,
                    float* w0pad, void* stream);
/* The same for up to four arenas in ONE launch (a step's online, target and transposed critic weights). Host array. */
typedef struct OrlkFusedPrep {
    const float* src; float* dst_lo; int64_t n;
    const float* W0; int64_t gs; float* w0pad;     /* W0 == NULL: lo words only */
    int32_t N, K0, G, pad_;
} OrlkFusedPrep;
int orlk_fused_prep_multi(const OrlkFusedPrep* jobs_host, int n_jobs, void* stream);
int orlk_sizeof_fused_prep(void);

/* Narrow-output linear layers (N <= 16: Critic.last, dist_net.mu/sigma, Actor.last;
 * modules/critic_module.py:15,26, dist_module.py:57-60, actor_module.py:44,49).
 *   fwd :  Y[g][m,n] = b[g][n] + sum_k X[g][m,k] * W[g][n*ldw + k*w_sk]       (one warp per row; w_sk = 1 for a
 *          row-major [N,K] weight, w_sk = leading dimension to read a column block of a [K, .] matrix, e.g. dQ/da)
 *   dgrad: dX[g][m,k] = (sum_n dY[g][m,n] * W[g][n*ldw+k]) * (mask ? mask[g][m,k] > 0 : 1)
 * groups g = 0..G-1 are addressed with the *_gs element strides. */
int orlk_skinny_fwd(const float* X, int64_t ldx, int64_t x_gs, const float* W, int64_t ldw, int64_t w_sk, int64_t w_gs,
                    const float* b, int64_t b_gs, float* Y, int64_t ldy, int64_t y_gs, int M, int K, int NS, int G,
                    void* stream);
int orlk_skinny_dgrad(const float* dY, int64_t ldy, int64_t y_gs, const float* W, int64_t ldw, int64_t w_gs,
                      const float* mask, int64_t ldm, int64_t m_gs, float* dX, int64_t ldx, int64_t x_gs, float* dXT,
                      int64_t ldxt, int64_t xt_gs, int M, int K, int NS, int G, void* stream);
/*   dXT (optional): transposed copy dXT[g][k*ldxt + m], the K-major operand of the tensor-core weight gradient. */

/* Streaming kernels for large row counts where one operand is narrow (<= 32 wide): the first critic / dynamics layer
 * (K = obs+act inputs) forward, and the weight gradients of that layer and of the narrow heads.
 *   fwd  : Y[g][m][n] = act(b[g][n] + sum_k X[g][m][k] W[g][n*ldw+k]), K <= 32; optional transposed copy YT[g][n][m]
 *   wgrad: per 128-row chunk c:  out[g][c][ns*s_ns + kw*s_kw] = sum_m Nar[g][m][ns] * Wide[g][m][kw]  (NS <= 32),
 *          wide_sum[g][c][kw] = sum_m Wide[g][m][kw],  nar_sum[g][c][ns] = sum_m Nar[g][m][ns]  (both optional);
 *          the chunks are summed by orlk_adam_step (g_splits = orlk_narrow_wgrad_chunks(M)). */
int orlk_narrow_fwd(const float* X, int64_t ldx, int64_t x_gs, const float* W, int64_t ldw, int64_t w_gs, const float* b,
                    int64_t b_gs, float* Y, int64_t ldy, int64_t y_gs, float* YT, int64_t ldyt, int64_t yt_gs, int M, int N,
                    int K, int G, int relu, void* stream);
int orlk_narrow_wgrad_chunks(int M);
int orlk_narrow_init(void); /* once per process, outside stream capture */
int orlk_narrow_wgrad(const float* Wide, int64_t ldw, int64_t w_gs, const float* Nar, int64_t ldn, int64_t n_gs, float* out,
                      int64_t s_ns, int64_t s_kw, int64_t o_gs, int64_t o_cs, float* wide_sum, int64_t ws_gs, int64_t ws_cs,
                      float* nar_sum, int64_t ns_gs, int64_t ns_cs, int M, int KW, int NS, int G, void* stream);

/* Row assembly for critic inputs: dst[row_off+m, 0:w1) = src1[(m / rep1), 0:w1); dst[.., w1:w1+w2) = src2[m, 0:w2)
 * (replaces torch.cat / repeat in critic_module.py:25 and cql.py:142-147). */
typedef struct OrlkConcatSeg {
    float* dst;
    const float* src1;
    const float* src2;
    int64_t ld_dst, ld1, ld2;
    int32_t M, w1, w2, rep1;
    int32_t row_start; /* prefix sum of M over segments */
    int32_t pad_;
} OrlkConcatSeg;
int orlk_concat_rows(const OrlkConcatSeg* segs_dev, int n_segs, int total_rows, void* stream);
/* Member-sharded ensembles (BASELINE.json configs[2], configs[4]): after an equal-block all-gather
 * every rank holds src[world][block_stride]; rank r's block carries counts_host[r] * per_member valid floats (uneven
 * splits: 10 critics on 4 GPUs = 3/3/2/2).  Writes them densely in rank order to dst [sum(counts)][per_member] -- the
 * [E, B] layout the reference's ensemble-wide reductions work on: the min over critics of the actor step and of the TD
 * target (edac.py:96-102, :124-131), the sum over critics of the normalised input gradients (:136-149), the per-member
 * holdout losses (ensemble_dynamics.py:145-168).  world <= 8. */
int orlk_compact_blocks(const float* src, int64_t block_stride, float* dst, int world, int per_member, const int* counts_host,
                        void* stream);

/* ------------------------------------------------------ stochastic policy head */
/* Philox4x32-10 fill: out[i] ~ N(0,1) for i < n_normal, then U[lo,hi) for the next n_uniform
 * (replaces the per-step device randn of dist_module.py:39-42 and the CPU uniform_ of cql.py:138-140
 * in performance mode; parity tests inject noise instead and set *enable = 0). counter[0] selects the
 * Philox stream offset; orlk_step_end advances it so that graph replays draw fresh numbers. */
int orlk_philox_fill(float* out, int64_t n_normal, int64_t n_uniform, float lo, float hi, uint64_t seed,
                     unsigned long long* counter, const int* enable, void* stream);

/* TanhDiagGaussian / SAC actforward (sac.py:66-77, dist_module.py:21-27,39-42,117-127).
 * head rows are [mu(0:A) | raw_log_sigma(A:2A)];  row m reads head[(head_row_off + m / rep)*ld_head].
 *   u = mu + exp(clamp(raw,-5,2)) * eps ; a = tanh(u) (eps == NULL -> mode) ; logp as in the reference.
 * Outputs: act[m*ld_act + i] (may alias into a critic-input row), logp[m],
 * and optionally xobs: copies obs[(m / rep)*ld_obs + j] into xout[m*ld_x + j], j < obs_dim. */
int orlk_tanh_gauss_sample(const float* head, int64_t ld_head, int head_row_off, int rep, const float* eps, int M, int A,
                           float* act, int64_t ld_act, float* logp, const float* obs, int64_t ld_obs, int obs_dim,
                           float* xout, int64_t ld_x, void* stream);
/* Backward of the above with eps fixed:
 *   dL/da = sum_{j<n_da} dA[j*da_gs + m*ld_da + i] (one slab per critic / ensemble member);
 *   dhead[m, 0:A) = dmu, dhead[m, A:2A) = draw (clamp-gated). glp[m] = dLoss/dlogp[m]. */
/* Actor head + reparameterised sampling in one launch (ActorProb.forward + TanhNormal.rsample / log_prob,
 * modules/actor_module.py + dist_module.py; cql.py:108-140 uses one head pass three times):
 *   head[m] = X[m] . W^T + b  (2A outputs: mu | raw log-std), then for every use whose head-row range [r0, r1) contains m and
 *   every repeat r < rep:  o = (m - r0) * rep + r,  a = tanh(mu + sigma * eps[o]),  act[o], logp[o], xout[o] = [obs[m - r0] | .]
 * (xout row o gets the obs_dim observation columns; act usually points at column obs_dim of the same row). */
typedef struct OrlkSampleUse {
    int32_t r0, r1, rep, obs_dim;
    const float* eps;     /* [rows * rep, A] or NULL (deterministic: a = tanh(mu)) */
    float* act; int64_t ld_act;
    float* logp;          /* [rows * rep] or NULL */
    const float* obs; int64_t ld_obs;
    float* xout; int64_t ld_x;
} OrlkSampleUse;
int orlk_sizeof_sample_use(void);
int orlk_head_sample(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* b, float* head, int M, int K, int A,
                     const OrlkSampleUse* uses_host, int n_uses, void* stream);
/* Entry of the actor's backward pass in one launch (sac.py:111-119 / cql.py:93-99 autograd): per row m
 *   dL/da[m] = sum_{c<n_c} dZ0[c][m][:] . W0[c][:, col0:col0+A]   (W0[c] is [Kc][ld_w0], the critics' first layers),
 *   dhead[m] = backward of orlk_tanh_gauss_sample (as orlk_tanh_gauss_bwd),
 *   dZlast[m][n] = (dhead[m][:] . Wh[:, n]) * (Hlast[m][n] > 0)     (Wh is the actor head [2A][Ka]).
 * Replaces orlk_skinny_fwd (d/da) + orlk_tanh_gauss_bwd + orlk_skinny_dgrad. */
int orlk_actor_bwd_entry(const float* dZ0, int64_t dz_gs, int Kc, int n_c, const float* W0, int64_t w0_gs, int ld_w0, int col0,
                         const float* head, const float* eps, const float* act, int64_t ld_act, const float* glp, int M, int A,
                         float* dhead, const float* Wh, int Ka, const float* Hlast, float* dZlast, void* stream);
int orlk_tanh_gauss_bwd(const float* head, int64_t ld_head, const float* eps, const float* act, int64_t ld_act,
                        const float* dA, int n_da, int64_t da_gs, int64_t ld_da, const float* glp, int M, int A,
                        float* dhead, int64_t ld_dhead, void* stream);
/* EDAC ensemble-diversity term (edac.py:136-149): g [E][B][A] = dQ_e/da; ghat = g / (|g| + 1e-10);
 *   G = mean_b sum_{i != j} <ghat_i, ghat_j> / (E-1);  writes gbar [E][B][A] = eta * dG/dg and out_loss[0] = eta * G. */
int orlk_edac_div(const float* g, int E, int B, int A, float eta, float* gbar, float* scratch /* ceil(B/256) floats */,
                  float* out_loss, void* stream);

/* ------------------------------------------------------------------ losses */
/* scalars block shared by the loss kernels of one learner (device memory, floats):
 *   [0] log_alpha  [1] alpha (value in use)  [2] cql_log_alpha  [3..] reserved */
enum { ORLK_SC_LOG_ALPHA = 0, ORLK_SC_ALPHA = 1, ORLK_SC_CQL_LOG_ALPHA = 2, ORLK_SC_COUNT = 8 };

/* Actor loss of SAC / CQL / EDAC (sac.py:111-126, cql.py:93-106, edac.py:96-110) for q[e][b], e < E:
 *   loss = mean_b(alpha*logp_b - min_e q_eb);  dq[e][b] = -1/B at the arg-min (E==2: ties split, torch.min(a,b);
 *   E>2: first index, torch.min(dim)); glp[b] = alpha/B.  If auto_alpha: loss_alpha = -mean(log_alpha*(logp+H)),
 *   one Adam step on log_alpha (group `alpha_group`), alpha <- exp(log_alpha) (clamped to [0,1] if clamp01).
 * out_losses[0] = actor loss, [1] = alpha loss, [2] = alpha (new). */
typedef struct OrlkAdamGroup {
    float lr, beta1, beta2, eps;
    float tau; /* polyak coefficient for descs with a target            */
    int32_t step; /* number of optimiser steps already applied             */
    /* bias corrections of the NEXT step t = step + 1, kept next to the counter so that no kernel evaluates a double
     * precision pow() on its critical path: set by the host when a group is created / restored, advanced by
     * orlk_step_end.  step_size = (float)((double)lr / bc1) as torch.optim.Adam computes it. */
    double bc1;      /* 1 - beta1^t */
    float bc2_sqrt;  /* sqrt(1 - beta2^t) */
    int32_t pad_;
} OrlkAdamGroup;

int orlk_sac_actor_loss(const float* q, int64_t q_es, int E, const float* logp, int B, float* scalars, int auto_alpha,
                        int clamp01, float target_entropy, OrlkAdamGroup* groups, int alpha_group, float* alpha_mv,
                        float* dq, int64_t dq_es, float* glp, float* out_losses, void* stream);

/* Policy-improvement step, twin critics with scalar heads, everything that needs no batch reduction in one launch:
 *   q[c][m] = H[c][m] . Wh[c] + bh[c]  (H [2][B][K], member strides h_gs / w_gs / b_gs),
 *   dq[c][m] = d/dq_c mean_b(alpha logp - min(q_0, q_1))  (as orlk_sac_actor_loss writes it for E = 2),  glp[m] = alpha / B,
 *   dZ[c][m][k] = dq[c][m] * Wh[c][k] * (H[c][m][k] > 0)   (the heads' input gradient, as orlk_skinny_dgrad with H as mask).
 * The loss value and the temperature step (orlk_sac_actor_loss on q) then run beside the backward pass.
 * Replaces Critic.last forward + autograd through it in sac.py:111-120 / cql.py:93-100. */
int orlk_twin_head_actor(const float* H, int64_t h_gs, const float* Wh, int64_t w_gs, const float* bh, int64_t b_gs,
                         const float* scalars, int B, int K, float* q, int64_t q_gs, float* dq, int64_t dq_gs, float* glp,
                         float* dZ, int64_t dz_gs, void* stream);

/* CQL critic phase loss (cql.py:108-205) for both critics; also COMBO's (combo.py:133-208), whose TD rows are the
 * real+fake mix (B), whose `- w * mean Q` term runs over the first n_qmean (= real) rows only (combo.py:196-203) and
 * whose R conservative rows come from the mix or from the fake rows alone (rho_s, combo.py:162-166).  CQL: n_qmean = B.
 *   q[c] : [B + 3R] rows = data | pi | pi_next | random  (c = 0,1; stride q_cs)
 *   tq[c]: [B * tq_rep] target-critic values on (s', a');  lp_next [B];  lp_pi, lp_pn [R]
 *   tq_rep = 1, or N with max_q_backup (cql.py:109-120): row b's N sampled next actions are consecutive, each target
 *   critic is maximised over them before the min over critics, and lp_next is not used.
 * Computes the TD target, the 3-way logsumexp per repeat row (the reference's quirk), the optional Lagrange
 * multiplier step, the per-row upstream gradients dq[c][.] and the losses
 *   out_losses[0..1] = critic1/2 loss, [2] = cql_alpha loss, [3] = cql_alpha (old, clamped).
 * One CTA per 128 rows writes dq directly (it needs no reduction); the last CTA to finish adds the per-CTA partial sums
 * in CTA order, writes the losses and takes the multiplier's Adam step.  scratch: orlk_cql_critic_loss_scratch_floats(B, R)
 * floats, zero-initialised once (the kernel leaves its counter at zero). */
int orlk_cql_critic_loss_scratch_floats(int B, int R);
int orlk_cql_critic_loss(const float* q, int64_t q_cs, const float* tq, int64_t tq_cs, const float* lp_next,
                         const float* lp_pi, const float* lp_pn, const float* rew, const float* term, int B, int n_qmean,
                         int tq_rep, int R, int A, float gamma, float cql_weight, float temperature, int deterministic_backup, int with_lagrange,
                         float lagrange_threshold, float* scalars, OrlkAdamGroup* groups, int cql_alpha_group,
                         float* cql_alpha_mv, float* dq, int64_t dq_cs, float* out_losses, float* scratch, void* stream);

/* Generic TD loss (sac.py:93-108, td3bc.py:87-104, iql.py:101-115, edac.py:124-134):
 *   y = r + gamma (1-d) [ min_{e2<E2} tq[e2] - (use_alpha ? alpha * lp_next : 0) ]
 *   out_losses[e] = mean_b (q[e][b]-y)^2,  *out_sum = sum_e (optional),  dq[e][b] = 2 (q[e][b]-y)/B,  y_out optional. */
int orlk_td_loss(const float* q, int64_t q_es, int E, const float* tq, int64_t tq_es, int E2, const float* lp_next,
                 const float* scalars, int use_alpha, const float* rew, const float* term, int B, float gamma, float* dq,
                 int64_t dq_es, float* y_out, float* out_losses, float* out_sum, void* stream);
/* out[g][b] = max over r < rep of x[g][b * rep + r]: EDAC's max_q_backup (edac.py:113-122) keeps, per target critic, the
 * best of `rep` sampled next actions; the result feeds orlk_td_loss as tq with use_alpha = 0. */
int orlk_segment_max(const float* x, int64_t x_gs, int G, int B, int rep, float* out, int64_t out_gs, void* stream);
/* IQL expectile value loss (iql.py:82-98): q = min(tq[0], tq[1]); writes dv[b], qmin[b], out_loss[0]. */
int orlk_iql_v_loss(const float* tq, int64_t tq_es, const float* v, int B, float expectile, float* dv, float* qmin,
                    float* out_loss, void* stream);
/* IQL advantage-weighted actor loss (iql.py:118-130; bounded DiagGaussian with state-independent sigma,
 * dist_module.py:65-76): z = pre-tanh mu head [B,A]; writes dz [B,A], dsigma [A] (gradient of sigma_param), loss. */
int orlk_iql_actor_loss(const float* z, int64_t ldz, const float* sigma_param, const float* act, int64_t lda, const float* qmin,
                        const float* v, int B, int A, float temperature, float max_mu, float* dz, int64_t lddz, float* dsigma,
                        float* out_loss, void* stream);
/* Deterministic actor head (actor_module.py:46-50) with TD3 target-policy smoothing (td3bc.py:90-91) when eps != NULL. */
int orlk_det_actor_fwd(const float* z, int64_t ldz, const float* eps, int M, int A, float max_action, float policy_noise,
                       float noise_clip, float* act, int64_t ld_act, const float* obs, int64_t ld_obs, int obs_dim, float* xout,
                       int64_t ld_x, void* stream);
/* TD3+BC actor loss (td3bc.py:107-112): writes dq[b] = -lambda/B, dabc[b,i] = 2(a-a_data)/(B A), out_loss[0]. */
int orlk_td3bc_actor_loss(const float* q, const float* a, int64_t lda, const float* a_data, int64_t ldd, int B, int A,
                          float bc_alpha, float* dq, float* dabc, int64_t ldg, float* out_loss, void* stream);
int orlk_det_actor_bwd(const float* a, int64_t lda, const float* dA0, int64_t ld0, const float* dA1, int64_t ld1, int M, int A,
                       float max_action, float* dz, int64_t lddz, void* stream);

/* ------------------------------------------------------- MOPO ensemble dynamics */
/* scaler.transform on [obs | act] (utils/scaler.py:25-31): X[s] = ([obs[s] | act[s]] - mu) / std. */
int orlk_dyn_input(const float* obs, int64_t ld_obs, const float* act, int64_t ld_act, const float* mu, const float* sd, int S,
                   int O, int A, float* X, int64_t ldx, void* stream);
/* Per-member bootstrap batch: dst[e][r][0:w) = src[idx[e*idx_ld + r0 + r]][0:w) (ensemble_dynamics.py:134,144,186-187). */
int orlk_gather_rows(const float* src, int64_t ld_src, int w, const int64_t* idx, int64_t idx_ld, int64_t r0, int E, int R,
                     float* dst, int64_t ld_dst, int64_t dst_es, void* stream);
/* partial[c] = scale * sum of squares of chunk c (4096 elements): the weight-decay term of the reported loss. */
int orlk_sumsq_chunks(int64_t n);
int orlk_sumsq(const float* x, int64_t n, float scale, float* partial, void* stream);
/* Gaussian NLL + soft-clamped logvar (ensemble_dynamics.py:193-201, dynamics_module.py:19-29): out [E][Bn][2D] = mean|raw,
 * y [E][Bn][D].  Writes dout (same shape as out), dmax[D], dmin[D] (incl. the +-coef terms) and out_loss[0]. */
int orlk_dyn_nll(const float* out, const float* y, int E, int Bn, int D, const float* max_lv, const float* min_lv, float coef,
                 const float* decay_partials, int n_decay, float* dout, float* dmax, float* dmin, float* out_loss,
                 float* scratch, void* stream);
/* scratch: NULL = one CTA; otherwise orlk_dyn_nll_scratch_floats(E, Bn, D) floats for the 64-rows-per-CTA form (partial
 * records summed in block order by a second small launch: deterministic). */
int orlk_dyn_nll_scratch_floats(int E, int Bn, int D);
/* Holdout MSE of the mean head per member (ensemble_dynamics.py:210-217); y [Bn][D] is shared by the members. */
int orlk_dyn_val_mse(const float* out, const float* y, int E, int Bn, int D, float* mse, void* stream);
/* Imagination epilogue (ensemble_dynamics.py:43-77): term_kind 0 halfcheetah, 1 hopper, 2 walker2d, 3 never
 * (utils/termination_fns.py).  Reference-stream mode: noise [E][S][D] float64 and midx [S] are the reference's two
 * NumPy draws.  Device mode (noise == NULL / midx == NULL): noise32 [S][D] normals for the chosen member and
 * pick_u [S] uniforms in [0,1) selecting elites[floor(u * n_elites)].
 * uncertainty_mode (ensemble_dynamics.py:60-72): 0 "aleatoric" max_e ||sigma_e||, 1 "pairwise-diff"
 * max_e ||m_e - mean_e m_e||, 2 "ensemble_std" sqrt(mean_d var_e m_e[d]) over the members' predicted next states m_e. */
int orlk_dyn_step(const float* out, int E, int S, int D, const float* max_lv, const float* min_lv, const float* obs,
                  int64_t ld_obs, const double* noise, const int* midx, const float* noise32, const float* pick_u,
                  const int* elites, int n_elites, int term_kind, float penalty_coef, int uncertainty_mode, float* next_obs,
                  float* reward, float* raw_reward, float* penalty, unsigned char* terminal, void* stream);
/* Stable compaction of rows with drop[i] == 0 (mopo.py:69-73); *count_out = number of survivors. */
int orlk_compact_rows(const unsigned char* drop, int S, const float* src, int64_t ld_src, int w, float* dst, int64_t ld_dst,
                      int* count_out, void* stream);

/* ---------------------------------------------------------------- optimiser */
/* Fused (split-K partial reduction) + Adam + polyak over a list of tensors (torch.optim.Adam as constructed in
 * run_example/run_cql.py:92-94; _sync_weight sac.py:60-64).  For element i of tensor d:
 *   g = sum_{s<g_splits} grad[s*g_split_stride + i] + wd * p[i]
 *   m <- m + (g-m)(1-b1);  v <- b2 v + (1-b2) g^2;  p <- p - (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 *   tgt <- tgt*(1-tau) + p*tau   (if tgt != NULL),  with t = groups[group].step + 1. */
typedef struct OrlkAdamDesc {
    float* p;
    float* m;
    float* v;
    float* tgt;
    const float* grad;
    int64_t n;
    int64_t g_split_stride;
    int32_t g_splits;
    int32_t group;
    float wd;
    int32_t block_start; /* prefix sum of ceil(n / 128) */
    int32_t flags;       /* ORLK_OPT_ADAM | ORLK_OPT_POLYAK */
    int32_t cols;        /* with pT: the tensor is [n/cols, cols] row-major ...                     */
    float* pT;           /* ... and pT receives its transpose [cols, n/cols] (K-major dgrad operand) */
} OrlkAdamDesc;
enum { ORLK_OPT_ADAM = 1, ORLK_OPT_POLYAK = 2 };
int orlk_adam_step(const OrlkAdamDesc* descs_dev, int n_descs, int total_blocks, const OrlkAdamGroup* groups, void* stream);
/* Last node of a step graph: groups[g].step += 1 for every g with bit g set in mask, and
 * philox_counter[0] += 1 when philox_counter != NULL. */
int orlk_step_end(OrlkAdamGroup* groups, unsigned int mask, unsigned long long* philox_counter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ORLK_B200_H */
