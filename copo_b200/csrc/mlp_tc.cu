// mlp_tc.cu - tcgen05 tensor-core path for the 256-wide layers of the policy / value MLPs (K5 of SURVEY.md 2.2).
//
//   y[M x 256] = epilogue( A[M x K] * W[256 x K]^T )            M = scenes x slots x steps (10^5 .. 10^7 rows)
//
// fp32 accuracy on bf16 tensor cores: every fp32 operand x is split into hi = bf16(x), lo = bf16(x - hi) and the
// product is taken as hi*hi + lo*hi + hi*lo + lo*lo with fp32 accumulation in TMEM ("split bf16") - value
// predictions feed GAE, which must match the fp32 reference to 1e-4 (SURVEY.md 7 "hard parts").  Activations are
// stored as [M][hi(Kp) | lo(Kp)], weights as [256][hi(Kp) | lo(Kp)]; one pipeline stage holds the hi and lo blocks
// of both operands for one 64-wide reduction block, and the MMA thread issues the four products from it, so every
// operand byte crosses L2 -> shared memory once per output tile.
//
// Kernel: persistent, one CTA per SM, 128 x 256 output tile.
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B) of A_hi, A_lo, W_hi, W_lo blocks, 2-stage ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16), accumulators in
//               TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warp 2      TMEM allocation
//   warps 4-11  epilogue: tcgen05.ld 32 lanes x 32 columns, bias (from shared memory) + tanh (or * (1 - h^2) for
//               the input gradient), writes fp32 and / or the [hi | lo] bf16 operand of the next layer
//
// Replaces: SlimFC hidden layers of CCModel / CoPOModel forward (torch_copo/algo_ccppo.py:108-170, 201-219;
// algo_copo.py:138-153) and their input-gradient GEMM in backward.
#include "tc_common.cuh"

namespace b2c {
namespace tc {

struct LinearArgs {
    const float* bias;        // [256] or null
    const float* dtanh_src;   // [M][ld_src] or null: multiply the result by (1 - h^2)
    const uint16_t* dtanh_split;   // [M][512] bf16 [hi | lo] or null: the same factor with h = hi + lo (the layer's own
                                   // tensor-core operand: the forward pass need not keep an fp32 copy of h)
    float* out_f32;           // [M][ld_out] or null
    uint16_t* out_split;      // [M][512] bf16 (hi | lo) or null
    int M, kp_blocks;         // kp_blocks = Kp / 64 reduction blocks, four products each
    int ld_out, ld_src, act;  // act: 0 none, 1 tanh
    int wide_f32, wide_split; // 1: the output rows are 32-byte aligned (256-bit stores)
    int resident, stages;     // resident weights mode; ring depth
    int probe;                // measurements only (B2C_TC_PROBE): 1 = the epilogue stops after its tcgen05.ld
    int products;             // 4: hi*hi + lo*hi + hi*lo + lo*lo; 3: without lo*lo (B2C_TC_PRODUCTS, measurements)
    // fused narrow output layer on the activated result: head_out[m][j] = head_b[j] + sum_n y[m][n] head_w[j][n]
    const float* head_w;      // [head_n][256] or null
    const float* head_b;      // [head_n] or null
    float* head_out;          // [M][head_n]
    int head_n;               // 1 (value) or 4 (policy logits)
    // Gaussian sample on the 4 logits (TorchDiagGaussian): actions [M][2], logp [M]; eps from (seed, step, row)
    float* actions;
    float* logp;
    uint32_t seed, step;
};

// EPI = epilogue warps (8, 12 or 16: two, three or four per scheduler).  The 8 column chunks of a tile are dealt to
// EPI / 4 column groups; a warp reads the TMEM lanes of its quarter (warp % 4) and the chunks of its group.
template <int EPI, bool PACKED = false>
__global__ void __launch_bounds__(128 + EPI * 32, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ LinearArgs args) {
    constexpr int NT = 128 + EPI * 32;
    constexpr int GROUPS = EPI / 4;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic (not an integer round trip): the compiler keeps the shared address space
    // and emits LDS / STS instead of generic loads and stores for everything derived from `smem`
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const bool resident = args.resident != 0;
    const int n_stages = args.stages;
    const int stage_bytes = resident ? A_PAIR_BYTES : STAGE_BYTES;
    uint8_t* s_w = smem;                                         // resident mode: [kp_blocks][W_hi | W_lo]
    uint8_t* ring = smem + (resident ? args.kp_blocks * W_PAIR_BYTES : 0);
    uint8_t* misc = ring + n_stages * stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(misc);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tmem_full = empty + MAX_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* w_full = tmem_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* s_bias = reinterpret_cast<float*>(misc + 256);
    float* s_head_w = s_bias + BLOCK_N;                          // [HEAD_MAX][256]
    float* s_part = s_head_w + HEAD_MAX * BLOCK_N;               // [2 halves... upper half's partials][128][HEAD_MAX]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (args.M + BLOCK_M - 1) / BLOCK_M;
    const int num_kb = args.kp_blocks;
    const int kp = args.kp_blocks * BLOCK_K;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPI); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < BLOCK_N; i += NT) s_bias[i] = args.bias ? args.bias[i] : 0.0f;
    if (args.head_w)
        for (int i = threadIdx.x; i < args.head_n * BLOCK_N; i += NT) s_head_w[i] = args.head_w[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            if (resident) {
                mbar_expect_tx(w_full, (uint32_t)(num_kb * W_PAIR_BYTES));
                for (int kb = 0; kb < num_kb; ++kb) {
                    tma_load_2d(s_w + kb * W_PAIR_BYTES, &map_w, kb * BLOCK_K, 0, w_full);                       // W_hi
                    tma_load_2d(s_w + kb * W_PAIR_BYTES + B_STAGE_BYTES, &map_w, kp + kb * BLOCK_K, 0, w_full);  // W_lo
                }
            }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    uint8_t* st = ring + stage * stage_bytes;
                    mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
                    tma_load_2d(st, &map_a, kb * BLOCK_K, tile * BLOCK_M, &full[stage]);                       // A_hi
                    tma_load_2d(st + A_STAGE_BYTES, &map_a, kp + kb * BLOCK_K, tile * BLOCK_M, &full[stage]);  // A_lo
                    if (!resident) {
                        tma_load_2d(st + 2 * A_STAGE_BYTES, &map_w, kb * BLOCK_K, 0, &full[stage]);            // W_hi
                        tma_load_2d(st + 2 * A_STAGE_BYTES + B_STAGE_BYTES, &map_w, kp + kb * BLOCK_K, 0, &full[stage]);
                    }
                    if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            if (resident) mbar_wait(w_full, 0u);                 // the weights have landed (once per CTA)
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring + stage * stage_bytes);
                    const uint32_t sw = resident ? smem_u32(s_w + kb * W_PAIR_BYTES) : sa + 2 * A_STAGE_BYTES;
                    const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_STAGE_BYTES);
                    const uint64_t w_hi = make_desc(sw);
                    const uint64_t w_lo = make_desc(sw + B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance the start address by 32 B (16 bf16) inside the 128 B swizzle row
                        const uint64_t o = (uint64_t)(k * 2);
                        umma_bf16(tmem_d, a_hi + o, w_hi + o, IDESC, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_d, a_lo + o, w_hi + o, IDESC, 1u);
                        umma_bf16(tmem_d, a_hi + o, w_lo + o, IDESC, 1u);
                        if (args.products == 4) umma_bf16(tmem_d, a_lo + o, w_lo + o, IDESC, 1u);
                    }
                    umma_commit(&empty[stage]);                  // frees the smem slot when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (one output row per thread), columns by half =====
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;                         // column group
        const int c_begin = (8 * grp + GROUPS - 1) / GROUPS, c_end = (8 * (grp + 1) + GROUPS - 1) / GROUPS;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = tile * BLOCK_M + q * 32 + lane;
            const bool live = row < args.M;
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            float hacc[HEAD_MAX] = {0.0f, 0.0f, 0.0f, 0.0f};
            uint64_t hacc2[HEAD_MAX] = {0ull, 0ull, 0ull, 0ull};         // PACKED: (even, odd) column partial sums
#pragma unroll 1
            for (int c = c_begin; c < c_end; ++c) {              // 32-column chunks of this warp's group
                uint32_t r[32];
                tmem_ld32(taddr0 + (uint32_t)(c * 32), r);
                if (args.probe == 1) { hacc[0] += __uint_as_float(r[0]) + __uint_as_float(r[31]); continue; }
                float v[32];
                if constexpr (PACKED) {
                    uint64_t vv[16];
                    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(s_bias + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        ulonglong2 bb = b2[j];
                        vv[2 * j] = add2(pk2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), bb.x);
                        vv[2 * j + 1] = add2(pk2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), bb.y);
                    }
                    if (args.act == 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) vv[j] = fast_tanh2(vv[j]);
                    }
                    if (args.head_w) {
                        for (int hj = 0; hj < args.head_n; ++hj) {
                            const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(s_head_w + hj * BLOCK_N + c * 32);
                            uint64_t a = hacc2[hj];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                ulonglong2 ww = w2[j];
                                a = fma2(vv[2 * j], ww.x, a);
                                a = fma2(vv[2 * j + 1], ww.y, a);
                            }
                            hacc2[hj] = a;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) upk2(vv[j], v[2 * j], v[2 * j + 1]);
                } else {
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 bb = b4[j];
                    v[4 * j] = __uint_as_float(r[4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
                    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
                }
                if (args.act == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fast_tanh(v[j]);
                }
                if (args.head_w) {
                    for (int hj = 0; hj < args.head_n; ++hj) {
                        const float4* w4 = reinterpret_cast<const float4*>(s_head_w + hj * BLOCK_N + c * 32);
                        float a = hacc[hj];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 ww = w4[j];
                            a = fmaf(v[4 * j], ww.x, a); a = fmaf(v[4 * j + 1], ww.y, a);
                            a = fmaf(v[4 * j + 2], ww.z, a); a = fmaf(v[4 * j + 3], ww.w, a);
                        }
                        hacc[hj] = a;
                    }
                }
                }
                if (live) {
                    if (args.dtanh_split) {
                        const uint4* hh = reinterpret_cast<const uint4*>(args.dtanh_split + (size_t)row * 512 + c * 32);
                        const uint4* hl = reinterpret_cast<const uint4*>(args.dtanh_split + (size_t)row * 512 + 256 + c * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {            // 8 columns per 16-byte load of each half
                            const uint4 a = hh[j], b = hl[j];
                            const uint32_t ah[4] = {a.x, a.y, a.z, a.w}, al[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float h0 = __uint_as_float(ah[e] << 16) + __uint_as_float(al[e] << 16);
                                const float h1 = __uint_as_float(ah[e] & 0xffff0000u) + __uint_as_float(al[e] & 0xffff0000u);
                                v[8 * j + 2 * e] *= (1.0f - h0 * h0);
                                v[8 * j + 2 * e + 1] *= (1.0f - h1 * h1);
                            }
                        }
                    }
                    if (args.dtanh_src) {
                        const float4* hs = reinterpret_cast<const float4*>(args.dtanh_src + (size_t)row * args.ld_src + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 h = hs[j];
                            v[4 * j] *= (1.0f - h.x * h.x); v[4 * j + 1] *= (1.0f - h.y * h.y);
                            v[4 * j + 2] *= (1.0f - h.z * h.z); v[4 * j + 3] *= (1.0f - h.w * h.w);
                        }
                    }
                    if (args.out_f32) {
                        float* op = args.out_f32 + (size_t)row * args.ld_out + c * 32;
                        if (args.wide_f32) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                st_global_256(op + 8 * j, __float_as_uint(v[8 * j]), __float_as_uint(v[8 * j + 1]),
                                              __float_as_uint(v[8 * j + 2]), __float_as_uint(v[8 * j + 3]),
                                              __float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]),
                                              __float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7]));
                        } else {
                            float4* o = reinterpret_cast<float4*>(op);
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                    if (args.out_split) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                            uint32_t hb = *reinterpret_cast<uint32_t*>(&h2);
                            float r0 = v[2 * j] - __uint_as_float(hb << 16);
                            float r1 = v[2 * j + 1] - __uint_as_float(hb & 0xffff0000u);
                            __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
                            hi[j] = hb; lo[j] = *reinterpret_cast<uint32_t*>(&l2);
                        }
                        uint16_t* ohp = args.out_split + (size_t)row * 512 + c * 32;
                        if (args.wide_split) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                st_global_256(ohp + 16 * j, hi[8 * j], hi[8 * j + 1], hi[8 * j + 2], hi[8 * j + 3],
                                              hi[8 * j + 4], hi[8 * j + 5], hi[8 * j + 6], hi[8 * j + 7]);
                                st_global_256(ohp + 256 + 16 * j, lo[8 * j], lo[8 * j + 1], lo[8 * j + 2], lo[8 * j + 3],
                                              lo[8 * j + 4], lo[8 * j + 5], lo[8 * j + 6], lo[8 * j + 7]);
                            }
                        } else {
                            uint4* oh = reinterpret_cast<uint4*>(ohp);
                            uint4* ol = reinterpret_cast<uint4*>(ohp + 256);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                oh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                                ol[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if constexpr (PACKED) {
#pragma unroll
                for (int hj = 0; hj < HEAD_MAX; ++hj) {
                    float e, o;
                    upk2(hacc2[hj], e, o);
                    hacc[hj] += e + o;
                }
            }
            if (args.head_w) {
                // the column groups of a row live in warps w, w + 4, ...: groups 1.. hand their partial sums to group 0
                // through shared memory (named barrier 1 = the epilogue warps); summed in group order
                const int rloc = q * 32 + lane;
                if (grp > 0) {
#pragma unroll
                    for (int hj = 0; hj < HEAD_MAX; ++hj) s_part[((grp - 1) * BLOCK_M + rloc) * HEAD_MAX + hj] = hacc[hj];
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI * 32) : "memory");
                if (grp == 0 && live) {
                    float o[HEAD_MAX];
#pragma unroll
                    for (int hj = 0; hj < HEAD_MAX; ++hj) {
                        float t = hacc[hj];
#pragma unroll
                        for (int g = 1; g < GROUPS; ++g) t += s_part[((g - 1) * BLOCK_M + rloc) * HEAD_MAX + hj];
                        o[hj] = (hj < args.head_n) ? t + args.head_b[hj] : 0.0f;
                    }
                    if (args.head_n == 4) {
                        reinterpret_cast<float4*>(args.head_out)[row] = make_float4(o[0], o[1], o[2], o[3]);
                        if (args.actions) {
                            float e0, e1;
                            normal2(args.seed, args.step, (uint32_t)row, e0, e1);
                            float s0 = expf(o[2]), s1 = expf(o[3]);
                            float a0 = o[0] + s0 * e0, a1 = o[1] + s1 * e1;
                            float z0 = (a0 - o[0]) / s0, z1 = (a1 - o[1]) / s1;
                            reinterpret_cast<float2*>(args.actions)[row] = make_float2(a0, a1);
                            if (args.logp) args.logp[row] = -0.5f * (z0 * z0 + z1 * z1) - 1.8378770664093453f - (o[2] + o[3]);
                        }
                    } else {
                        for (int hj = 0; hj < args.head_n; ++hj) args.head_out[(size_t)row * args.head_n + hj] = o[hj];
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI * 32) : "memory");
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// x fp32 [M][K] -> [M][2*Kp] bf16: hi in columns [0, Kp), lo in [Kp, 2Kp), zero padding beyond K
// ones_col: column K (the first padding column) holds 1.0 - the weight-gradient kernel then returns the bias gradient in
// that column (dz^T [x | 1]); the forward weights are zero there, so the forward pass does not see it
__global__ void split_rows_kernel(const float* __restrict__ x, int ldx, uint16_t* __restrict__ out, int M, int K, int Kp,
                                  int ones_col) {
    const size_t total = (size_t)M * Kp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t m = i / Kp; int k = (int)(i - m * Kp);
        float v = (k < K) ? x[m * ldx + k] : ((ones_col && k == K) ? 1.0f : 0.0f);
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        out[m * 2 * Kp + k] = __bfloat16_as_ushort(h);
        out[m * 2 * Kp + Kp + k] = __bfloat16_as_ushort(l);
    }
}
// W fp32 [N][K] (or its transpose) -> [rows][2*Kp] bf16: [hi | lo]
__global__ void prep_weight_kernel(const float* __restrict__ W, uint16_t* __restrict__ out, int N, int K, int Kp, int rows,
                                   int transpose) {
    // transpose = 0: out row n, reduction index k reads W[n][k]   (forward, rows = N, reduction K)
    // transpose = 1: out row k, reduction index n reads W[n][k]   (input gradient, rows = K, reduction N)
    const int red = transpose ? N : K;
    const size_t total = (size_t)rows * Kp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / Kp), c = (int)(i - (size_t)r * Kp);
        float v = 0.0f;
        if (c < red) v = transpose ? W[(size_t)c * K + r] : W[(size_t)r * K + c];
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        uint16_t* o = out + (size_t)r * 2 * Kp;
        o[c] = __bfloat16_as_ushort(h);
        o[Kp + c] = __bfloat16_as_ushort(l);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  dW[256 x K] += dz^T[256 x M] * x[M x K]   (reduction over the batch rows)
//
// Both operands are "MN-major" for this product (the reduction index m is the slow one in memory), so the same
// [hi | lo] row tensors the forward pass made are read again, now through 64-column x 32-row TMA boxes: a box is
// exactly one SWIZZLE_128B MN-major atom stack (32 reduction rows of 128 B).  Each CTA owns a contiguous slice of
// rows, accumulates its 256 x Kp partial in TMEM (2 halves of 128 output rows x Kp columns) over the whole slice and
// writes it to a workspace; a second kernel adds the partials into dW in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 32;                                // reduction rows per stage
constexpr int WG_BOX_BYTES = WG_ROWS * 128;                // one 64-column box: 4 KB
constexpr int WG_OPERAND_BYTES = 4 * WG_BOX_BYTES;         // up to 256 columns per operand half: 16 KB
constexpr int WG_STAGE_BYTES = 4 * WG_OPERAND_BYTES;       // dz_hi, dz_lo, x_hi, x_lo: 64 KB
constexpr int WG_STAGES = 3;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

// MN-major SWIZZLE_128B descriptor: LBO = byte distance between 64-element atoms along M/N (one box), SBO = byte
// distance between 8-row groups along the reduction (1024 B)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(WG_BOX_BYTES >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct WgradArgs {
    float* partial;        // [gridDim.x][256][kp]
    int M, kp, rows_per_cta;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ WgradArgs args) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic (not an integer round trip): the compiler keeps the shared address space
    // and emits LDS / STS instead of generic loads and stores for everything derived from `smem`
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* done = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kp = args.kp, n_xbox = kp / 64;
    const int row0 = blockIdx.x * args.rows_per_cta;
    int rows = args.M - row0;
    rows = rows > args.rows_per_cta ? args.rows_per_cta : rows;
    const int num_kb = rows > 0 ? (rows + WG_ROWS - 1) / WG_ROWS : 0;
    // instruction descriptor: D f32, A = B = bf16, A and B MN-major (bits 15, 16), N = kp, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(kp >> 3) << 17) | ((128u >> 4) << 24);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t stage_tx = (uint32_t)(2 * WG_OPERAND_BYTES + 2 * n_xbox * WG_BOX_BYTES);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1u);
                uint8_t* st = smem + stage * WG_STAGE_BYTES;
                mbar_expect_tx(&full[stage], stage_tx);
                const int r = row0 + kb * WG_ROWS;
                for (int b = 0; b < 4; ++b) {
                    tma_load_2d(st + b * WG_BOX_BYTES, &map_dz, b * 64, r, &full[stage]);                          // dz_hi
                    tma_load_2d(st + WG_OPERAND_BYTES + b * WG_BOX_BYTES, &map_dz, 256 + b * 64, r, &full[stage]);  // dz_lo
                }
                for (int b = 0; b < n_xbox; ++b) {
                    tma_load_2d(st + 2 * WG_OPERAND_BYTES + b * WG_BOX_BYTES, &map_x, b * 64, r, &full[stage]);     // x_hi
                    tma_load_2d(st + 3 * WG_OPERAND_BYTES + b * WG_BOX_BYTES, &map_x, kp + b * 64, r, &full[stage]);
                }
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h) {                    // output rows 128 h .. 128 h + 127 (two 64-wide atoms)
                    const uint64_t dz_hi = make_desc_mn(sa + h * 2 * WG_BOX_BYTES);
                    const uint64_t dz_lo = make_desc_mn(sa + WG_OPERAND_BYTES + h * 2 * WG_BOX_BYTES);
                    const uint64_t x_hi = make_desc_mn(sa + 2 * WG_OPERAND_BYTES);
                    const uint64_t x_lo = make_desc_mn(sa + 3 * WG_OPERAND_BYTES);
                    const uint32_t tmem_d = tmem_base + (uint32_t)(h * 256);
#pragma unroll
                    for (int k = 0; k < WG_ROWS / UMMA_K; ++k) {
                        const uint64_t o = (uint64_t)((k * UMMA_K * 128) >> 4);      // 16 reduction rows = 2048 B
                        umma_bf16(tmem_d, dz_hi + o, x_hi + o, idesc, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_d, dz_lo + o, x_hi + o, idesc, 1u);
                        umma_bf16(tmem_d, dz_hi + o, x_lo + o, idesc, 1u);
                        umma_bf16(tmem_d, dz_lo + o, x_lo + o, idesc, 1u);
                    }
                }
                umma_commit(&empty[stage]);
                if (kb == num_kb - 1) umma_commit(done);
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int out_row = h * 128 + q * 32 + lane;
        float* dst = args.partial + ((size_t)blockIdx.x * 256 + out_row) * kp;
        if (num_kb > 0) {
            mbar_wait(done, 0);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 256);
            for (int c = 0; c < kp / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(taddr0 + (uint32_t)(c * 32), r);
                float4* o = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
        } else {
            for (int c = 0; c < kp / 4; ++c) reinterpret_cast<float4*>(dst)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// dW[n][k] += sum_c partial[c][n][k]  (k < K), fixed summation order; with db: column K of the partials (the ones column
// of x, split_rows_kernel) is the bias gradient: db[n] += sum_c partial[c][n][K].  Four lanes share one output element
// (partials c = q, q + 4, ...; then q0 + q1 + q2 + q3 by two shuffles): 4x the loads in flight of one thread walking all
// the partials, same result on every run.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, float* __restrict__ db,
                                    int parts, int kp, int K) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, q = threadIdx.x & 3;
    const int W = db ? K + 1 : K;
    const bool live = idx < 256 * W;
    const int n = live ? idx / W : 0, k = live ? idx - n * W : 0;
    float s = 0.0f;
    if (live)
        for (int c = q; c < parts; c += 4) s += partial[((size_t)c * 256 + n) * kp + k];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (live && q == 0) {
        if (k < K) dW[n * K + k] += s; else db[n] += s;
    }
}

}  // namespace tc
}  // namespace b2c

using namespace b2c::tc;

extern "C" {

int b2c_tc_padded_k(int K) { return (K + BLOCK_K - 1) / BLOCK_K * BLOCK_K; }

int b2c_tc_split_rows(const float* x, int ldx, uint16_t* out, int M, int K, int Kp, void* stream) {
    return b2c_tc_split_rows_ones(x, ldx, out, M, K, Kp, 0, stream);
}

int b2c_tc_split_rows_ones(const float* x, int ldx, uint16_t* out, int M, int K, int Kp, int ones_col, void* stream) {
    if (M == 0) return B2C_OK;
    if (!x || !out || K < 1 || Kp < K || Kp % BLOCK_K) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_split_rows: bad argument");
    if (ones_col && K >= Kp) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_split_rows_ones: no padding column to hold the ones");
    size_t total = (size_t)M * Kp;
    int grid = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    split_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, out, M, K, Kp, ones_col);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_tc_prep_weight(const float* W, uint16_t* out, int N, int K, int Kp, int transpose, void* stream) {
    int red = transpose ? N : K, rows = transpose ? K : N;
    if (!W || !out || Kp < red || Kp % BLOCK_K) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_prep_weight: bad argument");
    size_t total = (size_t)rows * Kp;
    prep_weight_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, out, N, K, Kp, rows, transpose);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

static int tc_linear_launch(const uint16_t* a_split, const uint16_t* w_prep, const float* bias, const float* dtanh_src,
                            int ld_src, float* out_f32, int ld_out, uint16_t* out_split, int M, int Kp, int act,
                            const b2c_tc_head* head, void* stream, const uint16_t* dtanh_split = nullptr);

int b2c_tc_linear(const uint16_t* a_split, const uint16_t* w_prep, const float* bias, const float* dtanh_src, int ld_src,
                  float* out_f32, int ld_out, uint16_t* out_split, int M, int Kp, int act, void* stream) {
    if (M == 0) return B2C_OK;
    if (!out_f32 && !out_split) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear: no output requested");
    return tc_linear_launch(a_split, w_prep, bias, dtanh_src, ld_src, out_f32, ld_out, out_split, M, Kp, act, nullptr,
                            stream);
}

int b2c_tc_linear_head(const uint16_t* a_split, const uint16_t* w_prep, const float* bias, float* out_f32, int ld_out,
                       int M, int Kp, int act, const b2c_tc_head* head, void* stream) {
    if (M == 0) return B2C_OK;
    if (!head || !head->weight || !head->bias || !head->out || (head->n != 1 && head->n != 4))
        return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear_head: the fused output layer needs weight, bias, out and n in {1, 4}");
    if (head->actions && head->n != 4) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear_head: sampling needs the 4 policy logits");
    if (((uintptr_t)head->out | (uintptr_t)head->actions) & 15)
        return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear_head: outputs must be 16-byte aligned");
    return tc_linear_launch(a_split, w_prep, bias, nullptr, 0, out_f32, ld_out, nullptr, M, Kp, act, head, stream);
}

int b2c_tc_linear_dgrad(const uint16_t* dz_split, const uint16_t* wT_prep, const uint16_t* h_split, float* out_f32,
                        int ld_out, uint16_t* out_split, int M, int Kp, void* stream) {
    if (M == 0) return B2C_OK;
    if (!out_f32 && !out_split) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear_dgrad: no output requested");
    if (!h_split || ((uintptr_t)h_split & 15)) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear_dgrad: h_split must be 16-byte aligned");
    return tc_linear_launch(dz_split, wT_prep, nullptr, nullptr, 0, out_f32, ld_out, out_split, M, Kp, 0, nullptr, stream,
                            h_split);
}

static int tc_linear_launch(const uint16_t* a_split, const uint16_t* w_prep, const float* bias, const float* dtanh_src,
                            int ld_src, float* out_f32, int ld_out, uint16_t* out_split, int M, int Kp, int act,
                            const b2c_tc_head* head, void* stream, const uint16_t* dtanh_split) {
    if (M == 0) return B2C_OK;
    if (!a_split || !w_prep || Kp < BLOCK_K || Kp % BLOCK_K || M < 0)
        return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear: bad argument");
    if (((uintptr_t)a_split | (uintptr_t)w_prep | (uintptr_t)out_f32 | (uintptr_t)out_split | (uintptr_t)dtanh_src) & 15)
        return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear: pointers must be 16-byte aligned");
    if ((out_f32 && (ld_out & 3)) || (dtanh_src && (ld_src & 3)))
        return b2c_set_error(B2C_ERR_ARG, "b2c_tc_linear: row strides must be multiples of 4 floats");
    static int attr_set = 0;
    static int num_sms = 0;
    if (!attr_set) {
        B2C_CUDA(cudaFuncSetAttribute(tc_linear_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX_OPTIN));
        B2C_CUDA(cudaFuncSetAttribute(tc_linear_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX_OPTIN));
        B2C_CUDA(cudaFuncSetAttribute(tc_linear_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX_OPTIN));
        B2C_CUDA(cudaFuncSetAttribute(tc_linear_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX_OPTIN));
        B2C_CUDA(cudaFuncSetAttribute(tc_linear_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX_OPTIN));
        int dev = 0;
        B2C_CUDA(cudaGetDevice(&dev));
        B2C_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_set = 1;
    }
    CUtensorMap map_a, map_w;
    int rc = make_map(&map_a, a_split, (uint64_t)M, (uint64_t)2 * Kp, BLOCK_M);
    if (rc) return rc;
    rc = make_map(&map_w, w_prep, (uint64_t)BLOCK_N, (uint64_t)2 * Kp, BLOCK_N);
    if (rc) return rc;
    LinearArgs a;
    a.bias = bias; a.dtanh_src = dtanh_src; a.dtanh_split = dtanh_split; a.out_f32 = out_f32; a.out_split = out_split; a.M = M;
    a.kp_blocks = Kp / BLOCK_K; a.ld_out = ld_out; a.ld_src = ld_src; a.act = act;
    a.wide_f32 = (out_f32 && ((uintptr_t)out_f32 & 31) == 0 && (ld_out & 7) == 0) ? 1 : 0;
    a.wide_split = (out_split && ((uintptr_t)out_split & 31) == 0) ? 1 : 0;
    static const bool narrow = getenv("B2C_TC_NARROW_STORES") != nullptr;    // A/B switches for measurements
    static const bool streamed = getenv("B2C_TC_STREAM_WEIGHTS") != nullptr;
    if (narrow) a.wide_f32 = a.wide_split = 0;
    static const int probe = getenv("B2C_TC_PROBE") ? atoi(getenv("B2C_TC_PROBE")) : 0;
    static const int products = getenv("B2C_TC_PRODUCTS") ? atoi(getenv("B2C_TC_PRODUCTS")) : 4;
    // epilogue warps: with the fused output layer (next to no stores) four warps per scheduler hide the tcgen05.ld and
    // FMA latencies best; with a full-width output two per scheduler are faster - more warps only spread the row-strided
    // stores over more concurrent streams (profiles/r01_g_tc_probe_epi*.json)
    static const int epi_env = getenv("B2C_TC_EPI_WARPS") ? atoi(getenv("B2C_TC_EPI_WARPS")) : 0;
    const int epi = epi_env ? epi_env : (head ? 16 : 8);
    // f32x2 epilogue math (two columns per FFMA2 / FMUL2 / FADD2): 4-6 % faster than the scalar form on B200
    // (profiles/r02_b_tc_packed.md), same bits for the hidden layers; B2C_TC_PACKED=0 selects the scalar epilogue
    static const bool packed = !(getenv("B2C_TC_PACKED") && atoi(getenv("B2C_TC_PACKED")) == 0);
    a.probe = probe;
    a.products = products == 3 ? 3 : 4;
    a.resident = 0; a.stages = STAGES;
    int smem_bytes = SMEM_BYTES;
    if (!head && a.kp_blocks <= 2 && !streamed) {
        a.resident = 1;
        a.stages = MAX_STAGES;
        while (a.stages > 2 && resident_smem_bytes(a.kp_blocks, a.stages) > SMEM_MAX_OPTIN) a.stages -= 1;
        smem_bytes = resident_smem_bytes(a.kp_blocks, a.stages);
    }
    a.head_w = nullptr; a.head_b = nullptr; a.head_out = nullptr; a.head_n = 0; a.actions = nullptr; a.logp = nullptr;
    a.seed = 0; a.step = 0;
    if (head) {
        a.head_w = head->weight; a.head_b = head->bias; a.head_out = head->out; a.head_n = head->n;
        a.actions = head->actions; a.logp = head->logp; a.seed = head->seed; a.step = head->step;
    }
    int tiles = (M + BLOCK_M - 1) / BLOCK_M;
    int grid = tiles < num_sms ? tiles : num_sms;
    if (packed && epi == 16) tc_linear_kernel<16, true><<<grid, 128 + 16 * 32, smem_bytes, (cudaStream_t)stream>>>(map_a, map_w, a);
    else if (packed) tc_linear_kernel<8, true><<<grid, 128 + 8 * 32, smem_bytes, (cudaStream_t)stream>>>(map_a, map_w, a);
    else if (epi == 16) tc_linear_kernel<16><<<grid, 128 + 16 * 32, smem_bytes, (cudaStream_t)stream>>>(map_a, map_w, a);
    else if (epi == 12) tc_linear_kernel<12><<<grid, 128 + 12 * 32, smem_bytes, (cudaStream_t)stream>>>(map_a, map_w, a);
    else tc_linear_kernel<8><<<grid, 128 + 8 * 32, smem_bytes, (cudaStream_t)stream>>>(map_a, map_w, a);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_tc_wgrad_parts(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return -1;
    return sms;
}

int b2c_tc_wgrad(const uint16_t* dz_split, const uint16_t* x_split, float* workspace, float* dW, int M, int K, int Kp,
                 void* stream) {
    return b2c_tc_wgrad_bias(dz_split, x_split, workspace, dW, nullptr, M, K, Kp, stream);
}

int b2c_tc_wgrad_bias(const uint16_t* dz_split, const uint16_t* x_split, float* workspace, float* dW, float* db, int M, int K,
                      int Kp, void* stream) {
    if (M == 0) return B2C_OK;
    if (db && K >= Kp) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_wgrad_bias: x has no ones column (K == Kp)");
    if (!dz_split || !x_split || !workspace || !dW || Kp < BLOCK_K || Kp > 256 || Kp % BLOCK_K || K > Kp || M < 0)
        return b2c_set_error(B2C_ERR_ARG, "b2c_tc_wgrad: bad argument (Kp must be 64..256)");
    static int attr_set = 0;
    static int num_sms = 0;
    if (!attr_set) {
        B2C_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
        num_sms = b2c_tc_wgrad_parts();
        attr_set = 1;
    }
    CUtensorMap map_dz, map_x;
    int rc = make_map(&map_dz, dz_split, (uint64_t)M, 512, WG_ROWS);
    if (rc) return rc;
    rc = make_map(&map_x, x_split, (uint64_t)M, (uint64_t)2 * Kp, WG_ROWS);
    if (rc) return rc;
    WgradArgs a;
    a.partial = workspace; a.M = M; a.kp = Kp;
    int per = (M + num_sms - 1) / num_sms;
    a.rows_per_cta = (per + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
    cudaStream_t s = (cudaStream_t)stream;
    tc_wgrad_kernel<<<num_sms, NUM_THREADS, WG_SMEM_BYTES, s>>>(map_dz, map_x, a);
    B2C_CUDA(cudaGetLastError());
    wgrad_reduce_kernel<<<(256 * (K + 1) * 4 + 255) / 256, 256, 0, s>>>(workspace, dW, db, num_sms, Kp, K);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

}  // extern "C"
