// tcgen05 / TMA / mbarrier PTX wrappers, shared-memory matrix descriptors and tensor-map builders shared by the
// tensor-core kernels (orlk_tc.cu: one GEMM per launch; orlk_fused.cu: whole critic passes per launch).  sm_100a only.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "orlk_common.cuh"

namespace orlk {
namespace tcg {

constexpr int TCG_BK = 32;          // fp32 k-slab = 128 bytes = one SWIZZLE_128B row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// One lane of a converged warp.  Issuing TMA / tcgen05 instructions under elect.sync inside a warp-uniform branch lets
// the compiler keep descriptors and addresses in uniform registers; under a plain `lane == 0` test it wraps every
// UTCHMMA in an R2UR "waterfall" loop (~70 ns per instruction, measured).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, P;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell):
// 8-row groups of 128-byte rows, SBO = 1024 bytes between groups, LBO unused (1), layout_type 2.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major, N>>3 at [17,23), M>>4 at [24,29).
// MN-major TF32 operands: the only shared-memory layout the tensor core accepts is SWIZZLE_128B with 32-byte atoms
// (cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem layout";
// cute Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over 4 k-rows of 128 bytes).  A tile is stored [k][32 mn-elements]:
// 128-byte rows along M or N, the 32-byte chunk c of row r at chunk position c ^ (r & 3)  (what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), one 4 KB block per 32 mn-elements.  Canonical form in 16-byte units:
// ((8,n),(4,k)) : ((1,LBO),(8,SBO))  ->  LBO = bytes between mn blocks (4096), SBO = bytes between 4-row k groups (512).
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
           (1ull << 61);
}
__device__ __forceinline__ uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (lane = row m, column = k), B from a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
        "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]),
        "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]),
        "f"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    return v;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// The tensor core reads an fp32 word as TF32 by IGNORING the 13 low mantissa bits (measured: a kernel that rewrites
// hi = x & 0xFFFFE000 in shared memory and one that leaves x untouched give bit-identical results, while
// lo = x - cvt.rna.tf32(x) is off by 2^-11).  So the raw tile already is the "hi" operand and only
// lo = x - trunc_tf32(x) (exact in fp32) has to be written.
// On the truncation bias (round 2): with truncated splits every product comes out smaller by a relative 2^-22..2^-20
// (the dropped lo*lo term and the hardware's truncation of the lo words), i.e. a dot product is SHRUNK by ~7e-7 of
// itself (measured: 2.4e-6 absolute on pre-activations of scale 3) plus a random part of ~1e-7 -- the same size as the
// fp32 FFMA kernel's rounding noise.  Near a ReLU's zero only the random part matters, so rounding the split to nearest
// (cvt.rna, tried: +20 us per CQL step on the splitter's critical path) does not make mask decisions more stable.
__device__ __forceinline__ float4 lo_tf32(const float4& v) {
    return make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u),
                       v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                       v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u),
                       v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 3-D tensor map over [G][rows][K] fp32 (k contiguous), box = 32 (k) x box_rows x 1, 128-byte swizzle.
static inline int make_map(CUtensorMap* map, const float* base, int64_t ld, int64_t gs, int rows, int K, int G, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ORLK_ERR_UNSUPPORTED;
    }
    if (gs <= 0) gs = (int64_t)rows * ld;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)G};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)gs * 4};
    cuuint32_t box[3] = {(cuuint32_t)TCG_BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (ld=%lld gs=%lld rows=%d K=%d G=%d)", (int)r, (long long)ld,
                  (long long)gs, rows, K, G);
        return ORLK_ERR_BAD_ARG;
    }
    return 0;
}

// 3-D tensor map over an MN-major operand [G][K][MN] fp32 (mn contiguous), box = 32 (mn) x 32 (k) x 1, 128-byte swizzle
// with 32-byte atoms.
static inline int make_map_mn(CUtensorMap* map, const float* base, int64_t ld, int64_t gs, int MN, int K, int G) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ORLK_ERR_UNSUPPORTED;
    }
    if (gs <= 0) gs = (int64_t)K * ld;
    cuuint64_t dims[3] = {(cuuint64_t)MN, (cuuint64_t)K, (cuuint64_t)G};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)gs * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)TCG_BK, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (MN-major) failed with CUresult %d (ld=%lld gs=%lld MN=%d K=%d G=%d)", (int)r,
                  (long long)ld, (long long)gs, MN, K, G);
        return ORLK_ERR_BAD_ARG;
    }
    return 0;
}

// 4-D tensor map over the row-major output [splits][G][M][N] (n contiguous), box = 32 (n) x 32 (m), 128-byte swizzle.
static inline int make_map_c(CUtensorMap* map, float* base, int64_t ldc, int64_t gs, int64_t ss, int M, int N, int G, int S) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ORLK_ERR_UNSUPPORTED;
    }
    if (gs <= 0) gs = (int64_t)M * ldc;
    if (ss <= 0) ss = (int64_t)G * gs;
    cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)G, (cuuint64_t)S};
    cuuint64_t strides[3] = {(cuuint64_t)ldc * 4, (cuuint64_t)gs * 4, (cuuint64_t)ss * 4};
    cuuint32_t box[4] = {32, 32, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (C) failed with CUresult %d (ldc=%lld gs=%lld ss=%lld M=%d N=%d G=%d S=%d)", (int)r,
                  (long long)ldc, (long long)gs, (long long)ss, M, N, G, S);
        return ORLK_ERR_BAD_ARG;
    }
    return 0;
}

}  // namespace tcg
}  // namespace orlk
