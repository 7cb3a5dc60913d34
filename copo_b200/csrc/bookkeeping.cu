// bookkeeping.cu - CoPO rollout bookkeeping on the device (rows a8-a13, a15 of SURVEY.md section 8) + optimiser.
//
//   gae3                three reverse scans (native / neighbourhood / global) over the [T][N] rollout columns
//                       (rllib compute_advantages; torch_copo/algo_copo.py:189-204, 473-502; bootstrap rule
//                       algo_ccppo.py:362-365)
//   lcf_mix_stats/apply LCF-mixed advantage + whole-batch standardisation (algo_copo.py:539-551)
//   cc_obs_fuse         centralized-critic observation: mean-field / concat / none (algo_ccppo.py:225-355)
//   gather_rows         minibatch assembly (rllib minibatches(): shuffled row subsets)
//   adam_step, dot      torch.optim.Adam update (PPO lr 3e-4, LCF lr 1e-4) and <g_new, g_old> (algo_copo.py:274-278)
// All HBM-bound: one pass over their inputs, coalesced along the slot/row dimension.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "b2c_internal.h"

namespace b2c {

constexpr int FLAG_VALID = 1, FLAG_DONE = 2;

struct GaeArgs {
    const uint8_t* flags;          // [T][N]
    const float* rew[3];           // [T][N] (head 2 may be per scene: index n / g_div)
    const float* val[3];           // [T][N]
    float* adv[3];
    float* tgt[3];
    int T, N, heads, g_div;        // g_div: slots per scene when the global reward is stored per scene, else 1
    int g_ld;                      // row length of rew[2]
    float gamma, lambda_;
    const float* boot[3];          // optional [N]: value of the observation after the fragment's last row
};

__global__ void gae3_kernel(const GaeArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.N) return;
    double nv[3] = {0, 0, 0}, nadv[3] = {0, 0, 0};
    bool has_next = false;
    for (int t = a.T - 1; t >= 0; --t) {
        const size_t idx = (size_t)t * a.N + n;
        const int f = a.flags[idx];
        if (!(f & FLAG_VALID)) {
            for (int h = 0; h < a.heads; ++h) { a.adv[h][idx] = 0.0f; a.tgt[h][idx] = 0.0f; }
            continue;
        }
        const bool done = f & FLAG_DONE;
        for (int h = 0; h < a.heads; ++h) {
            const double g = (h == 2) ? 1.0 : (double)a.gamma;              // global head: gamma = 1 (algo_copo.py:498)
            const float r = (h == 2) ? a.rew[2][(size_t)t * a.g_ld + n / a.g_div] : a.rew[h][idx];
            const double v = (double)a.val[h][idx];
            double next_v = nv[h], next_adv = nadv[h];
            if (done) { next_v = 0.0; next_adv = 0.0; }
            else if (!has_next) {                                            // fragment cut: bootstrap with own value
                next_v = (a.boot[h] && t == a.T - 1) ? (double)a.boot[h][n] : v;   // (IPPO: value of the next obs)
                next_adv = 0.0;
            }
            const double delta = (double)r + g * next_v - v;
            const double adv = delta + g * (double)a.lambda_ * next_adv;
            a.adv[h][idx] = (float)adv;
            a.tgt[h][idx] = (float)(adv + v);
            nv[h] = v; nadv[h] = adv;
        }
        has_next = true;
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float lcf_mix(float adv, float nei, float lcf) {
    float phi = lcf * 3.14159274f * 0.5f;                                  // step_lcf * np.pi / 2 in float32
    return cosf(phi) * adv + sinf(phi) * nei;
}

// out[0..2] = sum x, sum x^2, count over valid rows of the mixed advantage; out[3..4] = sum, sum^2 of global adv
__global__ void lcf_mix_stats_kernel(const uint8_t* __restrict__ flags, const float* __restrict__ adv,
                                     const float* __restrict__ nei, const float* __restrict__ lcf,
                                     const float* __restrict__ gadv, size_t rows, double* __restrict__ out) {
    double s = 0, s2 = 0, c = 0, g = 0, g2 = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x) {
        if (flags && !(flags[i] & FLAG_VALID)) continue;
        float x = nei ? lcf_mix(adv[i], nei[i], lcf[i]) : adv[i];
        s += x; s2 += (double)x * x; c += 1.0;
        if (gadv) { float y = gadv[i]; g += y; g2 += (double)y * y; }
    }
    // warp -> CTA -> one atomic per CTA and statistic (float64 atomics to five addresses are the serial part)
    __shared__ double red[5][8];
    s = warp_sum_d(s); s2 = warp_sum_d(s2); c = warp_sum_d(c); g = warp_sum_d(g); g2 = warp_sum_d(g2);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][w] = s; red[1][w] = s2; red[2][w] = c; red[3][w] = g; red[4][w] = g2; }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[threadIdx.x][k];
        atomicAdd(&out[threadIdx.x], t);
    }
}

__global__ void lcf_mix_apply_kernel(const uint8_t* __restrict__ flags, const float* __restrict__ adv,
                                     const float* __restrict__ nei, const float* __restrict__ lcf,
                                     float* __restrict__ gadv, float* __restrict__ norm_adv, size_t rows, float mean,
                                     float inv_std, float gmean, float ginv_std) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x) {
        bool valid = !flags || (flags[i] & FLAG_VALID);
        float x = nei ? lcf_mix(adv[i], nei[i], lcf[i]) : adv[i];
        norm_adv[i] = valid ? (x - mean) * inv_std : 0.0f;
        if (gadv) gadv[i] = valid ? (gadv[i] - gmean) * ginv_std : 0.0f;
    }
}

// one warp per (t, slot) row.  mode 0: none (copy), 1: mean field, 2: concat of the 4 nearest
__global__ void cc_obs_fuse_kernel(const float* __restrict__ obs, const float* __restrict__ act,
                                   const uint8_t* __restrict__ flags, const unsigned long long* __restrict__ mf_mask,
                                   const int8_t* __restrict__ nei_list, float* __restrict__ cobs, size_t rows, int A,
                                   int D, int AD, int C, int mode, int counterfactual) {
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float* out = cobs + row * C;
    const float* own = obs + row * D;
    const bool valid = flags[row] & FLAG_VALID;
    for (int d = lane; d < C; d += 32) out[d] = (d < D && valid) ? own[d] : 0.0f;
    if (!valid || mode == 0) return;
    const size_t scene_row0 = row - (row % A);            // first slot of this (t, scene)
    if (mode == 1) {
        // neighbours that have a row at this t: lane j checks slot j and j + 32, the warp votes the mask together
        const unsigned long long m = mf_mask[row];
        const bool v0 = ((m >> lane) & 1ull) && lane < A && (flags[scene_row0 + lane] & FLAG_VALID);
        const bool v1 = ((m >> (lane + 32)) & 1ull) && lane + 32 < A && (flags[scene_row0 + lane + 32] & FLAG_VALID);
        const unsigned long long vm = (unsigned long long)__ballot_sync(0xffffffffu, v0) |
                                      ((unsigned long long)__ballot_sync(0xffffffffu, v1) << 32);
        const int cnt = __popcll(vm);
        if (cnt == 0) return;
        const float inv = 1.0f / (float)cnt;
        const int W = D + (counterfactual ? AD : 0);
        // lanes own columns (lane, lane + 32, ...: up to 8 per lane, W <= 256); the neighbours are walked once, in slot
        // order - the mask is the same on every lane, so the walk is uniform - and every lane adds its columns of the
        // neighbour's row (coalesced 128-byte reads)
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.0f;
        uint32_t lo = (uint32_t)vm, hi = (uint32_t)(vm >> 32);
        int base = 0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t w = half ? hi : lo;
            while (w) {
                const int j = base + __ffs((int)w) - 1;
                w &= w - 1u;
                const float* orow = obs + (scene_row0 + j) * D;
                const float* arow = act ? act + (scene_row0 + j) * AD : nullptr;       // only read when counterfactual
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int d = lane + 32 * u;
                    if (d < W) acc[u] += (d < D) ? orow[d] : arow[d - D];
                }
            }
            base = 32;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int d = lane + 32 * u;
            if (d < W) out[D + d] = acc[u] * inv;
        }
    } else {
        const int W = D + (counterfactual ? AD : 0);
        for (int k = 0; k < 4; ++k) {
            int j = nei_list[row * 4 + k];
            if (j < 0) continue;
            size_t r = scene_row0 + j;
            if (!(flags[r] & FLAG_VALID)) continue;
            for (int d = lane; d < W; d += 32)
                out[D + k * W + d] = (d < D) ? obs[r * D + d] : act[r * AD + (d - D)];
        }
    }
}

// Mean-field fusion, one CTA per (t, scene): the scene's observation rows (contiguous in HBM) and actions are staged in
// shared memory once, then every warp takes rows of the scene - the neighbour mask is warp-uniform, lanes own NU
// columns each (lane + 32 u < W = D + AD), so a neighbour costs NU shared loads and adds instead of eight predicated
// global loads; the row leaves as [own | mean] through coalesced stores.  Same sums in the same (slot) order as
// cc_obs_fuse_kernel, mode 1.  HBM traffic is the algorithmic 4 * (D + AD) in + 4 * (2 D + AD) out per row.
template <int NU, bool SPLIT>
__global__ void __launch_bounds__(128)
cc_obs_fuse_mf_scene_kernel(const float* __restrict__ obs, const float* __restrict__ act,
                            const uint8_t* __restrict__ flags, const unsigned long long* __restrict__ mf_mask,
                            float* __restrict__ cobs, uint16_t* __restrict__ cobs_split, int kp, int A, int D, int AD,
                            int C, int counterfactual) {
    extern __shared__ __align__(16) float s_fuse[];
    float* s_obs = s_fuse;                                       // [A][D]
    float* s_act = s_fuse + ((A * D + 3) & ~3);                  // [A][AD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const size_t row0 = (size_t)blockIdx.x * A;                  // first slot of this (t, scene)
    {
        const float* g = obs + row0 * D;
        const int n = A * D;
        if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
            const float4* g4 = reinterpret_cast<const float4*>(g);
            float4* s4 = reinterpret_cast<float4*>(s_obs);
            for (int k = tid; k < (n >> 2); k += blockDim.x) s4[k] = g4[k];
        } else {
            for (int k = tid; k < n; k += blockDim.x) s_obs[k] = g[k];
        }
        if (counterfactual) {
            const float* ga = act + row0 * AD;
            for (int k = tid; k < A * AD; k += blockDim.x) s_act[k] = ga[k];
        }
    }
    // slots of the scene that have a row at this t (the same on every warp)
    const bool f0 = lane < A && (flags[row0 + lane] & FLAG_VALID);
    const bool f1 = lane + 32 < A && (flags[row0 + lane + 32] & FLAG_VALID);
    const unsigned long long have = (unsigned long long)__ballot_sync(0xffffffffu, f0) |
                                    ((unsigned long long)__ballot_sync(0xffffffffu, f1) << 32);
    __syncthreads();
    const int W = D + (counterfactual ? AD : 0);
    // this lane's columns of a neighbour's [obs | act] row: base pointer and row stride per column slot
    const float* cbase[NU];
    int cstride[NU];
    bool con[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int d = lane + 32 * u;
        con[u] = d < W;
        cbase[u] = (d < D) ? s_obs + d : s_act + (d - D);
        cstride[u] = (d < D) ? D : AD;
    }
    for (int i = warp; i < A; i += n_warps) {
        float* out = cobs + (row0 + i) * C;
        const bool valid = (have >> i) & 1ull;
        const unsigned long long vm = valid ? (mf_mask[row0 + i] & have) : 0ull;
        const int cnt = __popcll(vm);
        float acc[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) acc[u] = 0.0f;
        uint32_t w = (uint32_t)vm;
        int base = 0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            while (w) {
                const int j = base + __ffs((int)w) - 1;
                w &= w - 1u;
#pragma unroll
                for (int u = 0; u < NU; ++u)
                    if (con[u]) acc[u] += cbase[u][j * cstride[u]];
            }
            w = (uint32_t)(vm >> 32);
            base = 32;
        }
        const float inv = cnt > 0 ? 1.0f / (float)cnt : 0.0f;
        const float* own = s_obs + i * D;
        // optionally the same row once more as the value network's tensor-core operand: bf16 [hi | lo], kp columns each,
        // zero padded (tc::split_rows_kernel's bits) - the critic's first layer then needs no conversion pass over cobs
        uint16_t* sp = SPLIT ? cobs_split + (row0 + i) * (size_t)(2 * kp) : nullptr;
        auto emit = [&](int d, float x) {
            out[d] = x;
            if constexpr (SPLIT) {
                const __nv_bfloat16 h = __float2bfloat16_rn(x);
                sp[d] = __bfloat16_as_ushort(h);
                sp[kp + d] = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
            }
        };
        for (int d = lane; d < D; d += 32) emit(d, valid ? own[d] : 0.0f);
#pragma unroll
        for (int u = 0; u < NU; ++u)
            if (con[u]) emit(D + lane + 32 * u, acc[u] * inv);
        if constexpr (SPLIT) for (int d = C + lane; d < kp; d += 32) { sp[d] = 0; sp[kp + d] = 0; }
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, size_t ld_src, const int64_t* __restrict__ idx,
                                   float* __restrict__ dst, size_t ld_dst, size_t rows, int width) {
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* s = src + (size_t)idx[row] * ld_src;
    float* d = dst + row * ld_dst;
    for (int k = lane; k < width; k += 32) d[k] = s[k];
}

// dst[c][r] = src[idx[r]][c]: the scalar columns of a minibatch, gathered from the row-major stack of the rollout's
// columns and laid out column by column, so that every column of the minibatch is a contiguous vector (no per-column
// copy afterwards).  One thread per row: a contiguous read of the row, coalesced writes along r.
__global__ void gather_cols_kernel(const float* __restrict__ src, size_t ld_src, const int64_t* __restrict__ idx,
                                   float* __restrict__ dst, size_t rows, int width) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* s = src + (size_t)idx[r] * ld_src;
    for (int c = 0; c < width; ++c) dst[(size_t)c * rows + r] = s[c];
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float step_size, float inv_sqrt_bc2, float b1, float b2,
                            float eps, float grad_scale) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        float mi = b1 * m[i] + (1.0f - b1) * gi;
        float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

__global__ void dot_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, double* __restrict__ out) {
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        s += (double)a[i] * (double)b[i];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

}  // namespace b2c

using namespace b2c;

static int grid_for(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

extern "C" {

int b2c_gae3(const b2c_gae_args* p, void* stream) {
    if (!p || !p->flags || p->heads < 1 || p->heads > 3 || p->T < 1 || p->N < 1)
        return b2c_set_error(B2C_ERR_ARG, "b2c_gae3: bad argument");
    GaeArgs a;
    a.flags = p->flags; a.T = p->T; a.N = p->N; a.heads = p->heads; a.gamma = p->gamma; a.lambda_ = p->lambda_;
    a.g_div = p->global_reward_per_scene > 0 ? p->global_reward_per_scene : 1;
    a.g_ld = p->global_reward_per_scene > 0 ? p->N / p->global_reward_per_scene : p->N;
    for (int h = 0; h < 3; ++h) {
        a.rew[h] = p->rewards[h]; a.val[h] = p->values[h]; a.adv[h] = p->advantages[h]; a.tgt[h] = p->targets[h];
        a.boot[h] = (h < p->heads) ? p->bootstrap[h] : nullptr;
        if (h < p->heads && (!a.rew[h] || !a.val[h] || !a.adv[h] || !a.tgt[h]))
            return b2c_set_error(B2C_ERR_ARG, "b2c_gae3: head %d has a null column", h);
    }
    gae3_kernel<<<(p->N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_lcf_mix_stats(const uint8_t* flags, const float* adv, const float* nei_adv, const float* step_lcf,
                      const float* global_adv, size_t rows, double* out5, void* stream) {
    if (rows == 0) return B2C_OK;
    if (!adv || !out5 || (nei_adv && !step_lcf)) return b2c_set_error(B2C_ERR_ARG, "b2c_lcf_mix_stats: bad argument");
    lcf_mix_stats_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)stream>>>(flags, adv, nei_adv, step_lcf, global_adv,
                                                                               rows, out5);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_lcf_mix_apply(const uint8_t* flags, const float* adv, const float* nei_adv, const float* step_lcf,
                      float* global_adv, float* normalized_adv, size_t rows, float mean, float std, float gmean,
                      float gstd, void* stream) {
    if (rows == 0) return B2C_OK;
    if (!adv || !normalized_adv) return b2c_set_error(B2C_ERR_ARG, "b2c_lcf_mix_apply: bad argument");
    float s = std > 1e-4f ? std : 1e-4f, gs = gstd > 1e-4f ? gstd : 1e-4f;      // max(1e-4, std): rllib standardized
    lcf_mix_apply_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)stream>>>(flags, adv, nei_adv, step_lcf, global_adv,
                                                                               normalized_adv, rows, mean, 1.0f / s,
                                                                               gmean, 1.0f / gs);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_cc_obs_fuse(const float* obs, const float* actions, const uint8_t* flags, const uint64_t* mf_mask,
                    const int8_t* nei_list, float* cobs, size_t rows, int slots, int obs_dim, int act_dim, int cobs_dim,
                    int mode, int counterfactual, void* stream) {
    return b2c_cc_obs_fuse_split(obs, actions, flags, mf_mask, nei_list, cobs, nullptr, 0, rows, slots, obs_dim, act_dim,
                                 cobs_dim, mode, counterfactual, stream);
}

int b2c_cc_obs_fuse_split(const float* obs, const float* actions, const uint8_t* flags, const uint64_t* mf_mask,
                          const int8_t* nei_list, float* cobs, uint16_t* cobs_split, int kp, size_t rows, int slots,
                          int obs_dim, int act_dim, int cobs_dim, int mode, int counterfactual, void* stream) {
    if (rows == 0) return B2C_OK;
    if (!obs || !flags || !cobs || slots < 1 || mode < 0 || mode > 2)
        return b2c_set_error(B2C_ERR_ARG, "b2c_cc_obs_fuse: bad argument");
    int n_other = mode == 0 ? 0 : (mode == 1 ? 1 : 4);
    int want = obs_dim + n_other * (obs_dim + (counterfactual ? act_dim : 0));
    if (cobs_dim != want) return b2c_set_error(B2C_ERR_ARG, "b2c_cc_obs_fuse: cobs_dim %d, expected %d (algo_ccppo.py:55-71)", cobs_dim, want);
    if ((mode == 1 && !mf_mask) || (mode == 2 && !nei_list) || (mode && counterfactual && !actions))
        return b2c_set_error(B2C_ERR_ARG, "b2c_cc_obs_fuse: missing neighbour columns");
    if (mode == 1 && obs_dim + (counterfactual ? act_dim : 0) > 256)
        return b2c_set_error(B2C_ERR_ARG, "b2c_cc_obs_fuse: mean-field rows of more than 256 columns are not supported");
    const int W = obs_dim + (counterfactual ? act_dim : 0);
    const size_t fuse_smem = (size_t)(((slots * obs_dim + 3) & ~3) + slots * act_dim) * sizeof(float);
    const bool scene_wise = mode == 1 && slots <= 64 && rows % (size_t)slots == 0 && fuse_smem <= 48 * 1024 &&
                            !getenv("B2C_FUSE_ROWWISE");
    if (cobs_split && (!scene_wise || kp < cobs_dim || kp % 64))
        return b2c_set_error(B2C_ERR_ARG, "b2c_cc_obs_fuse_split: the operand output needs the mean-field mode on whole scenes and kp = cobs_dim padded to 64");
    if (scene_wise) {
        // one CTA per (t, scene); rows are (t, scene, slot) with the slot fastest, as every caller lays them out
        const unsigned grid = (unsigned)(rows / (size_t)slots);
        cudaStream_t st = (cudaStream_t)stream;
#define B2C_MF(n) case n: if (cobs_split) cc_obs_fuse_mf_scene_kernel<n, true><<<grid, 128, fuse_smem, st>>>(obs, actions, flags, (const unsigned long long*)mf_mask, cobs, cobs_split, kp, slots, obs_dim, act_dim, cobs_dim, counterfactual); \
        else cc_obs_fuse_mf_scene_kernel<n, false><<<grid, 128, fuse_smem, st>>>(obs, actions, flags, (const unsigned long long*)mf_mask, cobs, nullptr, 0, slots, obs_dim, act_dim, cobs_dim, counterfactual); break;
        switch ((W + 31) / 32) { B2C_MF(1) B2C_MF(2) B2C_MF(3) B2C_MF(4) B2C_MF(5) B2C_MF(6) B2C_MF(7) B2C_MF(8) }
#undef B2C_MF
        B2C_CUDA(cudaGetLastError());
        return B2C_OK;
    }
    size_t blocks = (rows + 7) / 8;
    cc_obs_fuse_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(obs, actions, flags, (const unsigned long long*)mf_mask,
                                                                          nei_list, cobs, rows, slots, obs_dim, act_dim,
                                                                          cobs_dim, mode, counterfactual);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_gather_rows(const float* src, size_t ld_src, const int64_t* idx, float* dst, size_t ld_dst, size_t rows, int width,
                    void* stream) {
    if (rows == 0) return B2C_OK;
    if (!src || !idx || !dst || width < 1) return b2c_set_error(B2C_ERR_ARG, "b2c_gather_rows: bad argument");
    size_t blocks = (rows + 7) / 8;
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, idx, dst, ld_dst, rows, width);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_gather_cols(const float* src, size_t ld_src, const int64_t* idx, float* dst, size_t rows, int width, void* stream) {
    if (rows == 0) return B2C_OK;
    if (!src || !idx || !dst || width < 1) return b2c_set_error(B2C_ERR_ARG, "b2c_gather_cols: bad argument");
    gather_cols_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, idx, dst, rows, width);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, void* stream) {
    if (n == 0) return B2C_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) return b2c_set_error(B2C_ERR_ARG, "b2c_adam_step: bad argument");
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1),
                                                                   (float)(1.0 / sqrt(bc2)), beta1, beta2, eps, grad_scale);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_dot(const float* a, const float* b, size_t n, double* out, void* stream) {
    if (n == 0) return B2C_OK;
    if (!a || !b || !out) return b2c_set_error(B2C_ERR_ARG, "b2c_dot: null argument");
    dot_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

}  // extern "C"
