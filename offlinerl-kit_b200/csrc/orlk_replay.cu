// ReplayBuffer device mirror: row table packing and the index gather of ReplayBuffer.sample
// (reference: buffer/buffer.py:26-30 storage, :96-106 sample).  Pure copies -> bit-exact.
#include <string.h>
#include "orlk_common.cuh"
using namespace orlk;

// One warp per transition row: lanes stride over the row_w floats of the row.
__global__ void k_replay_pack(const float* __restrict__ obs, const float* __restrict__ nobs, const float* __restrict__ act,
                              const float* __restrict__ rew, const float* __restrict__ term, int64_t n, int O, int A,
                              float* __restrict__ table, int row_w, int64_t row_offset) {
    orlk::pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n; r += nwarps) {
        float* dst = table + (row_offset + r) * row_w;
        for (int j = lane; j < row_w; j += 32) {
            float v = 0.f;
            if (j < O) v = obs[r * O + j];
            else if (j < 2 * O) v = nobs[r * O + (j - O)];
            else if (j < 2 * O + A) v = act[r * A + (j - 2 * O)];
            else if (j == 2 * O + A) v = rew[r];
            else if (j == 2 * O + A + 1) v = term[r];
            dst[j] = v;
        }
    }
}

// One warp per sampled index: a coalesced read of the (<= 256 B) table row, scattered into the SoA batch.
__global__ void k_replay_gather(const float* __restrict__ table, int64_t n_rows, int row_w, int O, int A,
                                const int64_t* __restrict__ idx, int n, float* __restrict__ obs, float* __restrict__ nobs,
                                float* __restrict__ act, float* __restrict__ rew, float* __restrict__ term) {
    orlk::pdl_enter();
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= n) return;
    int64_t r = idx[warp];
    if (r < 0) r = 0;
    if (r >= n_rows) r = n_rows - 1;
    const float* src = table + r * row_w;
    const int used = 2 * O + A + 2;
    for (int j = lane; j < used; j += 32) {
        const float v = __ldg(src + j);
        if (j < O) obs[(int64_t)warp * O + j] = v;
        else if (j < 2 * O) nobs[(int64_t)warp * O + (j - O)] = v;
        else if (j < 2 * O + A) act[(int64_t)warp * A + (j - 2 * O)] = v;
        else if (j == 2 * O + A) rew[warp] = v;
        else term[warp] = v;
    }
}

extern "C" {

int orlk_replay_pack(const float* obs, const float* next_obs, const float* act, const float* rew, const float* term,
                     int64_t n, int obs_dim, int act_dim, float* table, int row_w, int64_t row_offset, void* stream) {
    ORLK_REQUIRE(n >= 0 && obs_dim > 0 && act_dim > 0, "sizes");
    ORLK_REQUIRE(row_w >= 2 * obs_dim + act_dim + 2, "row_w too small");
    if (n == 0) return 0;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    orlk::launch(k_replay_pack, (unsigned)blocks, 256, 0, (cudaStream_t)stream, obs, next_obs, act, rew, term, n, obs_dim, act_dim,
                                                                      table, row_w, row_offset);
    return check_launch("k_replay_pack");
}

int orlk_replay_gather_into(const float* table, int64_t n_rows, int row_w, int obs_dim, int act_dim, const int64_t* idx,
                            int n, float* obs, float* next_obs, float* act, float* rew, float* term, void* stream) {
    ORLK_REQUIRE(n >= 0 && n_rows > 0, "sizes");
    ORLK_REQUIRE(row_w >= 2 * obs_dim + act_dim + 2, "row_w too small");
    if (n == 0) return 0;
    const int wpb = 4;  // warps per block: 64 blocks for a 256 batch, spread over SMs
    orlk::launch(k_replay_gather, (n + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream, table, n_rows, row_w, obs_dim, act_dim,
                                                                                idx, n, obs, next_obs, act, rew, term);
    return check_launch("k_replay_gather");
}

int orlk_replay_gather(const float* table, int64_t n_rows, int row_w, int obs_dim, int act_dim, const int64_t* idx, int n,
                       float* obs2, float* act, float* rew, float* term, void* stream) {
    return orlk_replay_gather_into(table, n_rows, row_w, obs_dim, act_dim, idx, n, obs2, obs2 + (int64_t)n * obs_dim, act, rew,
                                   term, stream);
}

// ReplayBuffer.sample in ONE host call (the public-API step is host-latency sensitive: every ctypes round trip shows):
// wait until the pinned slot's previous upload has finished, copy the freshly drawn indices into it, upload, re-arm
// the slot's event, gather.
int orlk_replay_sample(const float* table, int64_t n_rows, int row_w, int obs_dim, int act_dim, const int64_t* idx_host,
                       int64_t* idx_pinned, int64_t* idx_dev, void* slot_event, int event_armed, int n, float* obs2,
                       float* act, float* rew, float* term, void* stream) {
    ORLK_REQUIRE(n > 0 && n_rows > 0, "sizes");
    ORLK_REQUIRE(idx_host != nullptr && idx_pinned != nullptr && idx_dev != nullptr && slot_event != nullptr, "index buffers");
    cudaStream_t s = (cudaStream_t)stream;
    if (event_armed) {
        int rc = check(cudaEventSynchronize((cudaEvent_t)slot_event), "cudaEventSynchronize");
        if (rc) return rc;
    }
    memcpy(idx_pinned, idx_host, sizeof(int64_t) * (size_t)n);
    int rc = check(cudaMemcpyAsync(idx_dev, idx_pinned, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, s), "index upload");
    if (rc) return rc;
    rc = check(cudaEventRecord((cudaEvent_t)slot_event, s), "cudaEventRecord");
    if (rc) return rc;
    return orlk_replay_gather(table, n_rows, row_w, obs_dim, act_dim, idx_dev, n, obs2, act, rew, term, stream);
}

}  // extern "C"
