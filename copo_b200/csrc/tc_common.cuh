// tc_common.cuh - tcgen05 / TMA / mbarrier building blocks shared by the tensor-core translation units (mlp_tc.cu,
// mlp_fused.cu): tile constants, PTX wrappers, the SWIZZLE_128B shared-memory descriptors, the rational tanh of the
// epilogues and the host-side tensor-map builder.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "b2c_internal.h"
#include "rng.cuh"

namespace b2c {
namespace tc {

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, UMMA_K = 16, STAGES = 2;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB per half (hi or lo)
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;      // 32 KB per half
constexpr int STAGE_BYTES = 2 * A_STAGE_BYTES + 2 * B_STAGE_BYTES;     // A_hi, A_lo, W_hi, W_lo: 96 KB
constexpr int HEAD_MAX = 4;                               // fused narrow output layer: up to 4 outputs
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias*/ +
                           HEAD_MAX * BLOCK_N * 4 /*head weights*/ + 3 * BLOCK_M * HEAD_MAX * 4 /*head partials*/;
// "resident weights" mode (reduction length <= 128, no fused output layer - the first layer of every network): the
// whole prepared weight matrix (hi | lo, 64 KB per reduction block) stays in shared memory for the life of the CTA and
// only the A blocks (32 KB) stream through a deeper ring, so the weights cross L2 -> shared memory once per CTA instead
// of once per 128-row tile.
constexpr int MAX_STAGES = 4;
constexpr int A_PAIR_BYTES = 2 * A_STAGE_BYTES;           // A_hi + A_lo of one reduction block: 32 KB
constexpr int W_PAIR_BYTES = 2 * B_STAGE_BYTES;           // W_hi + W_lo of one reduction block: 64 KB
constexpr int SMEM_MAX_OPTIN = 232448;
__host__ __device__ constexpr int resident_smem_bytes(int kp_blocks, int a_stages) {
    return kp_blocks * W_PAIR_BYTES + a_stages * A_PAIR_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias*/;
}
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = f32, A = B = bf16, both K-major,
// N = 256 (bits 17..22 = N >> 3), M = 128 (bits 24..28 = M >> 4)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((BLOCK_N >> 3) << 17) | ((BLOCK_M >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// SWIZZLE_128B, K-major shared-memory matrix descriptor (cute SmemDescriptor): start >> 4, LBO = 1 (unused for
// swizzled K-major), SBO = 1024 B (8 rows x 128 B) >> 4, version 1 (Blackwell), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// clamp to [-9, 9] in one instruction: min(|x|, 9) with the sign of x
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float clamp9(float x) {
    float y;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(9.0f));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_tanh(float x) {
    // odd rational minimax (13/6) on [-9, 9]: relative error < 4e-7 down to the smallest arguments (an exp-based
    // form loses relative accuracy near 0, which the value heads then amplify); checked in tests/test_tc_gpu.py.
    // The denominator lies in [4.8e-3, 1.7], so the plain reciprocal approximation (no range scaling) is exact enough;
    // fast_tanh2 below performs the same operations on two values at once and returns the same bits.
    x = clamp9(x);
    const float x2 = x * x;
    float p = -2.76076847742355e-16f;
    p = fmaf(p, x2, 2.00018790482477e-13f);
    p = fmaf(p, x2, -8.60467152213735e-11f);
    p = fmaf(p, x2, 5.12229709037114e-08f);
    p = fmaf(p, x2, 1.48572235717979e-05f);
    p = fmaf(p, x2, 6.37261928875436e-04f);
    p = fmaf(p, x2, 4.89352455891786e-03f);
    p *= x;
    float q = 1.19825839466702e-06f;
    q = fmaf(q, x2, 1.18534705686654e-04f);
    q = fmaf(q, x2, 2.26843463243900e-03f);
    q = fmaf(q, x2, 4.89352518554385e-03f);
    return p * rcp_approx(q);
}
// 32-byte global store (sm_100: STG.256): one thread fills a whole 32-byte sector, so the row-per-thread epilogue
// writes full sectors instead of two half-filled ones per 16-byte store pair
__device__ __forceinline__ void st_global_256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                              uint32_t a5, uint32_t a6, uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
                 "r"(a4), "r"(a5), "r"(a6), "r"(a7)
                 : "memory");
}
// ---- packed fp32 pairs (sm_100: fma / mul / add .f32x2, SASS FFMA2 / FMUL2 / FADD2) -----------------------------------
// The epilogues' bias add, rational tanh, [hi | lo] split and fused output layer on two columns per instruction.  Same
// operations per element as the scalar forms, so the hidden layers come out bit-identical; the fused output layer sums
// even and odd columns separately (last-bit differences).
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t bc2(float x) { return pk2(x, x); }
__device__ __forceinline__ uint64_t fast_tanh2(uint64_t x) {
    float x0, x1;
    upk2(x, x0, x1);
    x = pk2(clamp9(x0), clamp9(x1));
    const uint64_t x2 = mul2(x, x);
    uint64_t p = bc2(-2.76076847742355e-16f);
    p = fma2(p, x2, bc2(2.00018790482477e-13f));
    p = fma2(p, x2, bc2(-8.60467152213735e-11f));
    p = fma2(p, x2, bc2(5.12229709037114e-08f));
    p = fma2(p, x2, bc2(1.48572235717979e-05f));
    p = fma2(p, x2, bc2(6.37261928875436e-04f));
    p = fma2(p, x2, bc2(4.89352455891786e-03f));
    p = mul2(p, x);
    uint64_t q = bc2(1.19825839466702e-06f);
    q = fma2(q, x2, bc2(1.18534705686654e-04f));
    q = fma2(q, x2, bc2(2.26843463243900e-03f));
    q = fma2(q, x2, bc2(4.89352518554385e-03f));
    float q0, q1;
    upk2(q, q0, q1);
    return mul2(p, pk2(rcp_approx(q0), rcp_approx(q1)));
}
// two fp32 values -> bf16 pairs hi = bf16(v), lo = bf16(v - hi) (the [hi | lo] operand split; v - hi is exact)
__device__ __forceinline__ void split2(uint64_t v, uint32_t& hi, uint32_t& lo) {
    float v0, v1, r0, r1;
    upk2(v, v0, v1);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<uint32_t*>(&h2);
    const uint64_t r = fma2(pk2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u)), bc2(-1.0f), v);
    upk2(r, r0, r1);
    __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
    lo = *reinterpret_cast<uint32_t*>(&l2);
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 row-major [rows][cols] matrix, box = 64 columns x box_rows rows, 128-byte swizzle
// (K-major operand tiles use tall boxes; the MN-major tiles of the weight gradient use 32-row boxes)
static inline int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return b2c_set_error(B2C_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return b2c_set_error(B2C_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return B2C_OK;
}

}  // namespace tc
}  // namespace b2c
