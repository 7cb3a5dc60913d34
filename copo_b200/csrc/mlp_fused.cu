// mlp_fused.cu - one kernel for the whole inference pass of a policy / value network (K5 of SURVEY.md 2.2):
//
//   out[M x n] = head( tanh( tanh( A[M x K] W1^T + b1 ) W2^T + b2 ) )          n = 4 (policy logits) or 1 (a value)
//
// and, for the rollout step, the Gaussian action sample and its log-probability on the 4 logits.  The two 256-wide
// hidden layers never leave the SM: the first layer's accumulator is read from TMEM, biased, squashed, split into
// bf16 [hi | lo] and written straight into the shared-memory ring as the A operand of the second layer's reduction
// blocks (the SWIZZLE_128B K-major image a TMA load would have produced), while the second layer's weight blocks
// arrive by TMA next to it.  Compared with tc_linear + tc_linear_head this removes the [M][512] bf16 hidden operand
// from HBM (write + read, 1 KB per row each way) and one launch.  Same arithmetic, same order of the MMAs and the
// same epilogue functions as the two-kernel path: the hidden layers are bit-identical to it, the 256-term sum of the
// output layer is grouped differently (last-bit differences).
//
// Persistent, one CTA per SM, 128-row tiles, 640 threads:
//   warp 0       TMA producer: per tile, kp1/64 stages {A_hi, A_lo, W1_hi, W1_lo} then 4 stages {W2_hi, W2_lo}
//   warp 1       MMA issuer (one thread): layer 1 -> TMEM columns 0..255, layer 2 -> columns 256..511
//   warp 2       TMEM allocation
//   warp 3       finaliser: sums the output layer's partial sums, adds the bias, samples, writes the rows
//   warps 4-19   epilogue (4 TMEM lane quarters x 4 column groups), per tile: first D1 -> +b1 -> tanh -> [hi | lo] ->
//                ring, all 16 warps on one reduction block at a time (the tensor core idles until the first block is
//                there: measured timeline in profiles/r02_c_fused_trace.md); then D2 -> +b2 -> tanh -> partial sums
// so the first layer of tile i+1 runs on the tensor cores while the epilogue warps are still on D2 of tile i.
//
// Replaces: CCModel / CoPOModel forward without gradients (torch_copo/algo_ccppo.py:108-170, 201-219;
// algo_copo.py:138-153) - compute_actions in the rollout, value predictions in postprocess_trajectory.
#include "tc_common.cuh"

namespace b2c {
namespace tc {

constexpr uint32_t IDESC_N128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((BLOCK_M >> 4) << 24);
constexpr int F_EPI = 16;                                 // epilogue warps: 4 TMEM lane quarters x 4 column groups
constexpr int F_GROUPS = F_EPI / 4;
constexpr int F_THREADS = 128 + F_EPI * 32;
constexpr int F_SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * BLOCK_N * 4 /*biases*/ +
                             HEAD_MAX * BLOCK_N * 4 /*head weights*/ +
                             2 * F_GROUPS * BLOCK_M * HEAD_MAX * 4 /*output-layer partial sums, two tiles*/;

struct FusedArgs {
    const float* b1;          // [256]
    const float* b2;          // [256]
    const float* head_w;      // [head_n][256]
    const float* head_b;      // [head_n]
    float* head_out;          // [M][head_n]
    float* actions;           // [M][2] or null (head_n = 4)
    float* logp;              // [M] or null
    int M, kp1_blocks, head_n, products;
    uint32_t seed, step;
    long long* trace;         // diagnostics (b2c_tc_mlp2_set_trace): CTA 0 stamps clock64() at 16 pipeline events per tile
    // training forward (TRAIN = true): what the backward pass reads again leaves the SM as a by-product of the epilogues
    uint16_t* h1_split;       // [M][512] bf16: the first hidden layer as [hi | lo] (layer 2's A operand; 1 - h1^2 in dgrad)
    float* h2;                // [M][256] fp32: the second hidden layer (head_backward)
};
#define F_TRACE(k) do { if (args.trace && blockIdx.x == 0) args.trace[t_local * 16 + (k)] = clock64(); } while (0)

// epilogue warps wait with a back-off: a failed probe sleeps instead of spinning on issue slots the other epilogue's
// warps could use
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tnanosleep.u32 64;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <bool TRAIN>
__global__ void __launch_bounds__(F_THREADS, 1)
tc_mlp2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w1,
               const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ FusedArgs args) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic (not an integer round trip): the compiler keeps the shared address space
    // and emits LDS / STS instead of generic loads and stores for everything derived from `smem`
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = smem;
    uint8_t* misc = ring + STAGES * STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(misc);          // [STAGES] TMA bytes landed
    uint64_t* empty = full + STAGES;                             // [STAGES] the stage's MMAs have retired
    uint64_t* hfull = empty + STAGES;                            // [STAGES] the hidden block is in the stage's A half
    uint64_t* d1a_full = hfull + STAGES;                         // hidden columns 0..127 of layer 1 are in TMEM
    uint64_t* d1b_full = d1a_full + 1;                           // ... and 128..255
    uint64_t* d1_empty = d1b_full + 1;
    uint64_t* d2_full = d1_empty + 1;
    uint64_t* d2_empty = d2_full + 1;
    uint64_t* part_full = d2_empty + 1;                          // [2] the 16 epilogue warps posted their partial sums
    uint64_t* part_empty = part_full + 2;                        // [2] the finaliser has consumed them
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(part_empty + 2);
    float* s_b1 = reinterpret_cast<float*>(misc + 256);
    float* s_b2 = s_b1 + BLOCK_N;
    float* s_head_w = s_b2 + BLOCK_N;                            // [HEAD_MAX][256]
    float* s_part = s_head_w + HEAD_MAX * BLOCK_N;               // [2 tiles][4 column groups][128][HEAD_MAX] partial sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (args.M + BLOCK_M - 1) / BLOCK_M;
    const int nb1 = args.kp1_blocks;
    const int nb = nb1 + 4;                                      // reduction blocks (= ring stages used) per tile
    const int kp1 = nb1 * BLOCK_K;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&hfull[s], F_EPI); }
        mbar_init(d1a_full, 1); mbar_init(d1b_full, 1); mbar_init(d1_empty, F_EPI);
        mbar_init(d2_full, 1); mbar_init(d2_empty, F_EPI);
        for (int k = 0; k < 2; ++k) { mbar_init(&part_full[k], F_EPI); mbar_init(&part_empty[k], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // everything above touches no global data; what follows may read what the previous kernel in the stream wrote
    // (observations, freshly updated weights).  Launched with programmatic serialisation this kernel's prologue
    // overlaps the previous kernel's tail; the next kernel may start its own prologue from here on.
    griddep_wait();
    griddep_trigger();
    for (int i = threadIdx.x; i < BLOCK_N; i += F_THREADS) { s_b1[i] = args.b1[i]; s_b2[i] = args.b2[i]; }
    for (int i = threadIdx.x; i < args.head_n * BLOCK_N; i += F_THREADS) s_head_w[i] = args.head_w[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + (uint32_t)BLOCK_N;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t g = 0, t_local = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_local) {
                for (int b = 0; b < nb; ++b, ++g) {
                    const int s = (int)(g & 1u);
                    mbar_wait(&empty[s], ((g >> 1) & 1u) ^ 1u);
                    if (b == 0) F_TRACE(14);
                    if (b == nb1) F_TRACE(15);
                    uint8_t* st = ring + s * STAGE_BYTES;
                    uint8_t* sw = st + 2 * A_STAGE_BYTES;
                    if (b < nb1) {
                        mbar_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
                        tma_load_2d(st, &map_a, b * BLOCK_K, tile * BLOCK_M, &full[s]);                          // A_hi
                        tma_load_2d(st + A_STAGE_BYTES, &map_a, kp1 + b * BLOCK_K, tile * BLOCK_M, &full[s]);    // A_lo
                        tma_load_2d(sw, &map_w1, b * BLOCK_K, 0, &full[s]);                                      // W1_hi
                        tma_load_2d(sw + B_STAGE_BYTES, &map_w1, kp1 + b * BLOCK_K, 0, &full[s]);                // W1_lo
                    } else {
                        const int c = b - nb1;
                        mbar_expect_tx(&full[s], (uint32_t)W_PAIR_BYTES);
                        tma_load_2d(sw, &map_w2, c * BLOCK_K, 0, &full[s]);                                      // W2_hi
                        tma_load_2d(sw + B_STAGE_BYTES, &map_w2, BLOCK_N + c * BLOCK_K, 0, &full[s]);            // W2_lo
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t g = 0, t_local = 0;
            uint32_t hcnt0 = 0u, hcnt1 = 0u;                     // layer-2 uses of each stage so far (hfull phase)
            auto issue_block = [&](int s, uint32_t tmem_d, bool first_block) {
                const uint32_t sa = smem_u32(ring + s * STAGE_BYTES);
                const uint32_t sw = sa + 2 * A_STAGE_BYTES;
                const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_STAGE_BYTES);
                const uint64_t w_hi = make_desc(sw), w_lo = make_desc(sw + B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t o = (uint64_t)(k * 2);        // 32 B inside the 128 B swizzle row
                    umma_bf16(tmem_d, a_hi + o, w_hi + o, IDESC, (first_block && k == 0) ? 0u : 1u);
                    umma_bf16(tmem_d, a_lo + o, w_hi + o, IDESC, 1u);
                    umma_bf16(tmem_d, a_hi + o, w_lo + o, IDESC, 1u);
                    if (args.products == 4) umma_bf16(tmem_d, a_lo + o, w_lo + o, IDESC, 1u);
                }
                umma_commit(&empty[s]);                          // frees the stage when these MMAs retire
            };
            // one 128-column half of a first-layer block: N = 128 MMAs on rows 128 half .. of the W tile (16 KB further on)
            auto issue_half = [&](uint32_t s, uint32_t tmem_d, uint32_t half, bool first_block) {
                const uint32_t sa = smem_u32(ring + s * STAGE_BYTES);
                const uint32_t sw = sa + 2 * A_STAGE_BYTES + half * (uint32_t)(128 * 128);
                const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_STAGE_BYTES);
                const uint64_t w_hi = make_desc(sw), w_lo = make_desc(sw + B_STAGE_BYTES);
                const uint32_t td = tmem_d + half * 128u;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t o = (uint64_t)(k * 2);
                    umma_bf16(td, a_hi + o, w_hi + o, IDESC_N128, (first_block && k == 0) ? 0u : 1u);
                    umma_bf16(td, a_lo + o, w_hi + o, IDESC_N128, 1u);
                    umma_bf16(td, a_hi + o, w_lo + o, IDESC_N128, 1u);
                    if (args.products == 4) umma_bf16(td, a_lo + o, w_lo + o, IDESC_N128, 1u);
                }
            };
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_local) {
                // ---- layer 1 -> D1 (epilogue 1 is done with the previous tile's D1) ----
                mbar_wait(d1_empty, (t_local & 1u) ^ 1u);
                F_TRACE(13);
                tc_fence_after();
                if (nb1 <= STAGES) {
                    // every first-layer block is resident at once: hidden columns 0..127 first (over all blocks), then
                    // 128..255 - epilogue 1 starts on reduction blocks 0 and 1 of layer 2 while the tensor core is still
                    // on the second half, so layer 2 can follow layer 1 without a gap (profiles/r02_c_fused_trace.md)
                    for (int b = 0; b < nb1; ++b) {
                        const uint32_t gb = g + (uint32_t)b;
                        mbar_wait(&full[gb & 1u], (gb >> 1) & 1u);
                        if (b < 2) F_TRACE(b);
                        tc_fence_after();
                        issue_half(gb & 1u, tmem_d1, 0u, b == 0);
                    }
                    umma_commit(d1a_full);
                    for (int b = 0; b < nb1; ++b) {
                        const uint32_t gb = g + (uint32_t)b;
                        issue_half(gb & 1u, tmem_d1, 1u, b == 0);
                        umma_commit(&empty[gb & 1u]);            // frees the stage when both halves have read it
                    }
                    umma_commit(d1b_full);
                    g += (uint32_t)nb1;
                } else {
                    for (int b = 0; b < nb1; ++b, ++g) {
                        const int s = (int)(g & 1u);
                        mbar_wait(&full[s], (g >> 1) & 1u);
                        if (b < 2) F_TRACE(b);
                        tc_fence_after();
                        issue_block(s, tmem_d1, b == 0);
                    }
                    umma_commit(d1a_full);
                    umma_commit(d1b_full);
                }
                // ---- layer 2 -> D2 (epilogue 2 is done with the previous tile's D2) ----
                mbar_wait(d2_empty, (t_local & 1u) ^ 1u);
                F_TRACE(2);
                tc_fence_after();
                for (int c = 0; c < 4; ++c, ++g) {
                    const int s = (int)(g & 1u);
                    mbar_wait(&full[s], (g >> 1) & 1u);          // W2 block landed
                    mbar_wait(&hfull[s], (s ? hcnt1 : hcnt0) & 1u);      // hidden block written by epilogue 1
                    if (s) hcnt1 += 1u; else hcnt0 += 1u;
                    F_TRACE(3 + c);
                    tc_fence_after();
                    issue_block(s, tmem_d2, c == 0);
                }
                umma_commit(d2_full);
            }
        }
    } else if (warp == 3) {
        // ===== finaliser: output-layer sums of the four column groups -> bias -> (sample) -> global.  Off the epilogue
        // warps' critical path: they post their partial sums and go straight to the next tile =====
        uint32_t t_local = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_local) {
            const int buf = (int)(t_local & 1u);
            mbar_wait_backoff(&part_full[buf], (t_local >> 1) & 1u);
            const float4* part = reinterpret_cast<const float4*>(s_part + buf * F_GROUPS * BLOCK_M * HEAD_MAX);
#pragma unroll 1
            for (int pass = 0; pass < BLOCK_M / 32; ++pass) {
                const int rloc = pass * 32 + lane;
                const int row = tile * BLOCK_M + rloc;
                float4 t = part[rloc];
#pragma unroll
                for (int gq = 1; gq < F_GROUPS; ++gq) {          // fixed order: group 0 + 1 + 2 + 3
                    const float4 u = part[gq * BLOCK_M + rloc];
                    t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
                if (row < args.M) {
                    if (args.head_n == 4) {
                        const float o0 = t.x + args.head_b[0], o1 = t.y + args.head_b[1];
                        const float o2 = t.z + args.head_b[2], o3 = t.w + args.head_b[3];
                        reinterpret_cast<float4*>(args.head_out)[row] = make_float4(o0, o1, o2, o3);
                        if (args.actions) {
                            float e0, e1;
                            normal2(args.seed, args.step, (uint32_t)row, e0, e1);
                            float s0 = expf(o2), s1 = expf(o3);
                            float a0 = o0 + s0 * e0, a1 = o1 + s1 * e1;
                            float z0 = (a0 - o0) / s0, z1 = (a1 - o1) / s1;
                            reinterpret_cast<float2*>(args.actions)[row] = make_float2(a0, a1);
                            if (args.logp) args.logp[row] = -0.5f * (z0 * z0 + z1 * z1) - 1.8378770664093453f - (o2 + o3);
                        }
                    } else {
                        args.head_out[row] = t.x + args.head_b[0];               // head_n = 1
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&part_empty[buf]);
            if (lane == 0) F_TRACE(12);
        }
    } else if (warp >= 4) {
        // ===== epilogue warps (16 = 4 TMEM lane quarters x 4 column groups) =====
        // Two jobs per tile: E1(t) = D1 -> +b1 -> tanh -> [hi | lo] A operand of layer 2 (all warps on one reduction
        // block at a time - the tensor core idles until the first block is there), and E2(t) = D2 -> +b2 -> tanh ->
        // partial sums of the output layer.  They are interleaved so that neither the tensor core nor these warps wait
        // for the other longer than necessary (measured timelines: profiles/r02_c_fused_trace.md):
        //   iteration i:  E2(i-1) first half | read E2(i-1) second half into registers, hand D2 back |
        //                 E1(i), all four blocks | E2(i-1) second half from the registers, post the partial sums
        // D2 is free again ~half an E2 after it filled, and E1(i)'s first block is not queued behind a whole E2.
        const int q = warp & 3;                                  // TMEM lane quarter = rows 32 q .. 32 q + 31 of the tile
        const int sub = (warp - 4) >> 2;                         // column group 0..3
        const int r = q * 32 + lane;                             // row in the tile
        const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const uint32_t swz = (uint32_t)(r & 7);
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

        // bias + tanh on 32 raw accumulator columns, then the output layer's (even, odd) partial sums
        auto e2_half = [&](const uint32_t* rr, int c, uint64_t* hacc2, float* h2_row) {
            uint64_t vv[16];
            const ulonglong2* bp = reinterpret_cast<const ulonglong2*>(s_b2 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const ulonglong2 bb = bp[j];
                vv[2 * j] = fast_tanh2(add2(pk2(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1])), bb.x));
                vv[2 * j + 1] = fast_tanh2(add2(pk2(__uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3])), bb.y));
            }
            if constexpr (TRAIN) {
                if (h2_row) {                                    // 32 columns of this row: four 32-byte stores
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float f[8];
#pragma unroll
                        for (int u = 0; u < 4; ++u) upk2(vv[4 * j + u], f[2 * u], f[2 * u + 1]);
                        st_global_256(h2_row + c * 32 + 8 * j, __float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                                      __float_as_uint(f[3]), __float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]),
                                      __float_as_uint(f[7]));
                    }
                }
            }
            if (args.head_n == 4) {
#pragma unroll
                for (int hj = 0; hj < 4; ++hj) {
                    const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(s_head_w + hj * BLOCK_N + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const ulonglong2 ww = wp[j];
                        hacc2[hj] = fma2(vv[2 * j], ww.x, hacc2[hj]);
                        hacc2[hj] = fma2(vv[2 * j + 1], ww.y, hacc2[hj]);
                    }
                }
            } else {
                const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(s_head_w + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const ulonglong2 ww = wp[j];
                    hacc2[0] = fma2(vv[2 * j], ww.x, hacc2[0]);
                    hacc2[0] = fma2(vv[2 * j + 1], ww.y, hacc2[0]);
                }
            }
        };

#pragma unroll 1
        for (int i = 0; i <= my_tiles; ++i) {
            const uint32_t t_local = (uint32_t)i;                // tile of E1 in this iteration (F_TRACE index)
            const uint32_t tp = (uint32_t)(i - 1);               // tile of E2 in this iteration
            uint64_t hacc2[HEAD_MAX] = {0ull, 0ull, 0ull, 0ull};
            uint32_t rb[32];                                     // second half of E2's columns, raw, held across E1
            float* h2_row = nullptr;                             // TRAIN: this thread's row of h2 for tile i - 1
            if constexpr (TRAIN) {
                const long long row2 = ((long long)blockIdx.x + (long long)(i - 1) * gridDim.x) * BLOCK_M + r;
                if (i > 0 && row2 < args.M) h2_row = args.h2 + row2 * BLOCK_N;
            }
            if (i > 0) {
                mbar_wait_backoff(d2_full, tp & 1u);
                if (warp == 4 && lane == 0) { if (args.trace && blockIdx.x == 0) args.trace[tp * 16 + 10] = clock64(); }
                tc_fence_after();
                uint32_t ra[32];
                tmem_ld32(tmem_d2 + lane_base + (uint32_t)(sub * 64), ra);
                e2_half(ra, 2 * sub, hacc2, h2_row);
                tmem_ld32(tmem_d2 + lane_base + (uint32_t)(sub * 64 + 32), rb);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d2_empty);            // layer 2 of the next tile may start
                if (warp == 4 && lane == 0) { if (args.trace && blockIdx.x == 0) args.trace[tp * 16 + 11] = clock64(); }
            }
            if (i < my_tiles) {
                // ---- E1(i): 16 columns of every reduction block ----
                const uint32_t g_base = t_local * (uint32_t)nb + (uint32_t)nb1;
                uint16_t* h1_row = nullptr;                      // TRAIN: this thread's row of the [hi | lo] hidden operand
                if constexpr (TRAIN) {
                    const long long row1 = ((long long)blockIdx.x + (long long)i * gridDim.x) * BLOCK_M + r;
                    if (row1 < args.M) h1_row = args.h1_split + row1 * (2 * BLOCK_N);
                }
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {                    // reduction block of layer 2 = hidden columns 64 c ..
                    if (c == 0) {
                        mbar_wait_backoff(d1a_full, t_local & 1u);
                        if (warp == 4 && lane == 0) F_TRACE(7);
                        tc_fence_after();
                    } else if (c == 2) {
                        mbar_wait_backoff(d1b_full, t_local & 1u);
                        tc_fence_after();
                    }
                    const uint32_t g = g_base + (uint32_t)c;
                    const int s = (int)(g & 1u);
                    uint8_t* a_hi = ring + s * STAGE_BYTES + row_off;
                    uint8_t* a_lo = a_hi + A_STAGE_BYTES;
                    uint32_t rr[16];
                    tmem_ld16(tmem_d1 + lane_base + (uint32_t)(c * 64 + sub * 16), rr);
                    if (c == 3) {                                // this warp's last read of D1: hand it back early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(d1_empty);
                    }
                    // bias, tanh and the [hi | lo] split on column pairs (f32x2: two columns per instruction)
                    uint32_t hi[8], lo[8];
                    const ulonglong2* bp = reinterpret_cast<const ulonglong2*>(s_b1 + c * 64 + sub * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const ulonglong2 bb = bp[j];
                        const uint64_t t0 = fast_tanh2(add2(pk2(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1])), bb.x));
                        const uint64_t t1 = fast_tanh2(add2(pk2(__uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3])), bb.y));
                        split2(t0, hi[2 * j], lo[2 * j]);
                        split2(t1, hi[2 * j + 1], lo[2 * j + 1]);
                    }
                    if constexpr (TRAIN) {
                        if (h1_row) {                            // the same 16 columns to HBM: 32 bytes into each half of the row
                            uint16_t* dst = h1_row + c * 64 + sub * 16;
                            st_global_256(dst, hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7]);
                            st_global_256(dst + BLOCK_N, lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[6], lo[7]);
                        }
                    }
                    mbar_wait_backoff(&empty[s], ((g >> 1) & 1u) ^ 1u);          // the stage's previous MMAs have retired
#pragma unroll
                    for (int j = 0; j < 2; ++j) {                // 16-byte chunks 2 sub + j of the row, XOR-swizzled
                        const uint32_t off = ((uint32_t)(2 * sub + j) ^ swz) << 4;
                        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                    fence_async_shared();                        // generic-proxy stores -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&hfull[s]);
                    if (warp == 4 && lane == 0 && c < 2) F_TRACE(8 + c);
                }
            }
            if (i > 0) {
                // ---- E2(i-1), second half, and the partial sums for the finaliser (double-buffered by tile parity) ----
                e2_half(rb, 2 * sub + 1, hacc2, h2_row);
                const int buf = (int)(tp & 1u);
                float hs[HEAD_MAX];
#pragma unroll
                for (int hj = 0; hj < HEAD_MAX; ++hj) {
                    float e, o;
                    upk2(hacc2[hj], e, o);
                    hs[hj] = e + o;
                }
                mbar_wait_backoff(&part_empty[buf], ((tp >> 1) & 1u) ^ 1u);      // the finaliser is done with tile tp - 2
                reinterpret_cast<float4*>(s_part + (buf * F_GROUPS + sub) * BLOCK_M * HEAD_MAX)[r] =
                    make_float4(hs[0], hs[1], hs[2], hs[3]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&part_full[buf]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc
}  // namespace b2c

using namespace b2c::tc;

static long long* g_trace = nullptr;

extern "C" {

/* diagnostics: CTA 0 of the following b2c_tc_mlp2_head launches writes clock64() stamps of 16 pipeline events per tile
 * into dev_buffer ([tiles of CTA 0][16]); null turns it off */
int b2c_tc_mlp2_set_trace(long long* dev_buffer) { g_trace = dev_buffer; return B2C_OK; }

static int launch_mlp2(const uint16_t* a_split, int Kp1, const uint16_t* w1_prep, const float* b1, const uint16_t* w2_prep,
                       const float* b2, const b2c_tc_head* head, uint16_t* h1_split, float* h2, int M, void* stream,
                       const char* who) {
    if (M == 0) return B2C_OK;
    if (!a_split || !w1_prep || !w2_prep || !b1 || !b2 || M < 0 || Kp1 < BLOCK_K || Kp1 % BLOCK_K)
        return b2c_set_error(B2C_ERR_ARG, "%s: bad argument", who);
    if (!head || !head->weight || !head->bias || !head->out || (head->n != 1 && head->n != 4))
        return b2c_set_error(B2C_ERR_ARG, "%s: the output layer needs weight, bias, out and n in {1, 4}", who);
    if (head->actions && head->n != 4) return b2c_set_error(B2C_ERR_ARG, "%s: sampling needs the 4 policy logits", who);
    if (((uintptr_t)a_split | (uintptr_t)w1_prep | (uintptr_t)w2_prep | (uintptr_t)head->out | (uintptr_t)head->actions) & 15)
        return b2c_set_error(B2C_ERR_ARG, "%s: pointers must be 16-byte aligned", who);
    const bool train = h1_split != nullptr;
    if (train && (!h2 || (((uintptr_t)h1_split | (uintptr_t)h2) & 31)))
        return b2c_set_error(B2C_ERR_ARG, "%s: h1_split and h2 must both be given, 32-byte aligned", who);
    static int attr_set = 0;
    static int num_sms = 0;
    if (!attr_set) {
        B2C_CUDA(cudaFuncSetAttribute(tc_mlp2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
        B2C_CUDA(cudaFuncSetAttribute(tc_mlp2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
        int dev = 0;
        B2C_CUDA(cudaGetDevice(&dev));
        B2C_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_set = 1;
    }
    CUtensorMap map_a, map_w1, map_w2;
    int rc = make_map(&map_a, a_split, (uint64_t)M, (uint64_t)2 * Kp1, BLOCK_M);
    if (rc) return rc;
    rc = make_map(&map_w1, w1_prep, (uint64_t)BLOCK_N, (uint64_t)2 * Kp1, BLOCK_N);
    if (rc) return rc;
    rc = make_map(&map_w2, w2_prep, (uint64_t)BLOCK_N, (uint64_t)2 * BLOCK_N, BLOCK_N);
    if (rc) return rc;
    static const int products = getenv("B2C_TC_PRODUCTS") ? atoi(getenv("B2C_TC_PRODUCTS")) : 4;
    static const bool pdl = !(getenv("B2C_TC_PDL") && atoi(getenv("B2C_TC_PDL")) == 0);
    FusedArgs a;
    a.b1 = b1; a.b2 = b2; a.head_w = head->weight; a.head_b = head->bias; a.head_out = head->out;
    a.actions = head->actions; a.logp = head->logp; a.M = M; a.kp1_blocks = Kp1 / BLOCK_K; a.head_n = head->n;
    a.products = products == 3 ? 3 : 4;
    a.seed = head->seed; a.step = head->step;
    a.trace = g_trace;
    a.h1_split = h1_split; a.h2 = h2;
    const int tiles = (M + BLOCK_M - 1) / BLOCK_M;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(tiles < num_sms ? tiles : num_sms));
    lc.blockDim = dim3((unsigned)F_THREADS);
    lc.dynamicSmemBytes = F_SMEM_BYTES;
    lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
    if (train) B2C_CUDA(cudaLaunchKernelEx(&lc, tc_mlp2_kernel<true>, map_a, map_w1, map_w2, a));
    else B2C_CUDA(cudaLaunchKernelEx(&lc, tc_mlp2_kernel<false>, map_a, map_w1, map_w2, a));
    return B2C_OK;
}

int b2c_tc_mlp2_head(const uint16_t* a_split, int Kp1, const uint16_t* w1_prep, const float* b1, const uint16_t* w2_prep,
                     const float* b2, const b2c_tc_head* head, int M, void* stream) {
    return launch_mlp2(a_split, Kp1, w1_prep, b1, w2_prep, b2, head, nullptr, nullptr, M, stream, "b2c_tc_mlp2_head");
}

int b2c_tc_mlp2_train(const uint16_t* a_split, int Kp1, const uint16_t* w1_prep, const float* b1, const uint16_t* w2_prep,
                      const float* b2, const b2c_tc_head* head, uint16_t* h1_split, float* h2, int M, void* stream) {
    if (!h1_split || !h2) return b2c_set_error(B2C_ERR_ARG, "b2c_tc_mlp2_train: h1_split and h2 are required");
    return launch_mlp2(a_split, Kp1, w1_prep, b1, w2_prep, b2, head, h1_split, h2, M, stream, "b2c_tc_mlp2_train");
}

}  // extern "C"
