// heads.cu - narrow output layers and per-row loss heads of the policy / value networks.
//
//   head_forward / head_backward   logits (4) and value (1) layers: 256 -> n, n <= 8 (memory-bound, no GEMM tile)
//   gaussian_sample                TorchDiagGaussian sample + logp (rllib; used at algo_ccppo.py:201-208 rollouts)
//   ppo_head                       per-row forward + backward of {IPPO,CCPPO,CoPO}Policy.loss
//                                  (algo_ippo.py:79-172, algo_ccppo.py:376-472, algo_copo.py:311-424) and of the two
//                                  policy losses inside CoPOPolicy.meta_update (algo_copo.py:250-272)
//   lcf_meta_terms                 LCF side of the meta-gradient (algo_copo.py:280-287 with model.compute_coordinated
//                                  algo_copo.py:155-161)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "b2c_internal.h"
#include "rng.cuh"

namespace b2c {

constexpr int HEAD_MAX_N = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y[m][j] = b[j] + sum_k h[m][k] W[j][k]; one warp per row (strided), W staged in shared memory
template <int NOUT>
__global__ void head_forward_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ W,
                                    const float* __restrict__ b, float* __restrict__ y, int ldy, int M, int K) {
    extern __shared__ float sW[];
    for (int idx = threadIdx.x; idx < NOUT * K; idx += blockDim.x) sW[idx] = W[idx];
    __syncthreads();
    const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    for (int m = blockIdx.x * warps + (threadIdx.x >> 5); m < M; m += gridDim.x * warps) {
        float acc[NOUT];
#pragma unroll
        for (int j = 0; j < NOUT; ++j) acc[j] = 0.0f;
        const float* row = h + (size_t)m * ldh;
        for (int k = lane; k < K; k += 32) {
            float hv = row[k];
#pragma unroll
            for (int j = 0; j < NOUT; ++j) acc[j] = fmaf(hv, sW[j * K + k], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < NOUT; ++j) acc[j] = warp_sum(acc[j]);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NOUT; ++j) y[(size_t)m * ldy + j] = acc[j] + (b ? b[j] : 0.0f);
        }
    }
}

// thread k of a CTA owns column k: dz[m][k] = (sum_j dy[m][j] W[j][k]) * (1 - h[m][k]^2), dW[j][k] += sum_m dy h,
// db[j] += sum_m dy[m][j], optionally dz_colsum[k] += sum_m dz[m][k] (bias gradient of the hidden layer below).  A CTA walks a contiguous chunk of rows, four rows in flight per thread; dz can also
// leave as the [hi | lo] bf16 operand of the tensor-core kernels (dz_split [M][2 * K], K = 256).
template <int NOUT>
__global__ void head_backward_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ h, int ldh,
                                     const float* __restrict__ W, float* __restrict__ dz, int ldz,
                                     uint16_t* __restrict__ dz_split, float* __restrict__ dW, float* __restrict__ db,
                                     float* __restrict__ dz_colsum, int M, int K, int rows, int dtanh) {
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    const int m_begin = blockIdx.x * rows, m_end = min(M, m_begin + rows);
    float w[NOUT], accw[NOUT], accb = 0.0f, accz = 0.0f;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) { w[j] = (k < K) ? W[j * K + k] : 0.0f; accw[j] = 0.0f; }
    constexpr int U = 4;
    for (int m0 = m_begin; m0 < m_end; m0 += U) {
        float g[U][NOUT], hv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = m0 + u;
            const bool in = m < m_end;
#pragma unroll
            for (int j = 0; j < NOUT; ++j) g[u][j] = in ? dy[(size_t)m * ldy + j] : 0.0f;
            hv[u] = (in && k < K) ? h[(size_t)m * ldh + k] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = m0 + u;
            if (m < m_end && k < K) {
                float s = 0.0f;
#pragma unroll
                for (int j = 0; j < NOUT; ++j) { s = fmaf(g[u][j], w[j], s); accw[j] = fmaf(g[u][j], hv[u], accw[j]); }
                const float z = dtanh ? s * (1.0f - hv[u] * hv[u]) : s;
                accz += z;
                if (dz) dz[(size_t)m * ldz + k] = z;
                if (dz_split) {
                    __nv_bfloat16 hi = __float2bfloat16_rn(z);
                    __nv_bfloat16 lo = __float2bfloat16_rn(z - __bfloat162float(hi));
                    dz_split[(size_t)m * 2 * K + k] = __bfloat16_as_ushort(hi);
                    dz_split[(size_t)m * 2 * K + K + k] = __bfloat16_as_ushort(lo);
                }
            }
            if (blockIdx.y == 0 && threadIdx.x < NOUT && m < m_end) accb += g[u][threadIdx.x];
        }
    }
    if (k < K && dW) {
#pragma unroll
        for (int j = 0; j < NOUT; ++j) atomicAdd(&dW[j * K + k], accw[j]);
    }
    if (db && blockIdx.y == 0 && threadIdx.x < NOUT) atomicAdd(&db[threadIdx.x], accb);
    // column sums of dz = the bias gradient of the layer that produced h (the thread owns the column: no reduction)
    if (dz_colsum && k < K) atomicAdd(&dz_colsum[k], accz);
}

// logits[m] = (mu0, mu1, ls0, ls1); action = mu + exp(ls) * eps; logp as TorchDiagGaussian.logp
__global__ void gaussian_sample_kernel(const float* __restrict__ logits, const float* __restrict__ eps_in,
                                       float* __restrict__ actions, float* __restrict__ logp,
                                       float* __restrict__ eps_out, int M, uint32_t seed, uint32_t step,
                                       int deterministic) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float4 l = reinterpret_cast<const float4*>(logits)[m];
    float e0, e1;
    if (deterministic) { e0 = 0.0f; e1 = 0.0f; }
    else if (eps_in) { e0 = eps_in[2 * m]; e1 = eps_in[2 * m + 1]; }
    else normal2(seed, step, (uint32_t)m, e0, e1);
    float s0 = expf(l.z), s1 = expf(l.w);
    float a0 = l.x + s0 * e0, a1 = l.y + s1 * e1;
    float z0 = (a0 - l.x) / s0, z1 = (a1 - l.y) / s1;
    reinterpret_cast<float2*>(actions)[m] = make_float2(a0, a1);
    if (logp) logp[m] = -0.5f * (z0 * z0 + z1 * z1) - 1.8378770664093453f - (l.z + l.w);
    if (eps_out) { eps_out[2 * m] = e0; eps_out[2 * m + 1] = e1; }
}

struct PpoHeadArgs {
    const float* logits;        // [M][4] current policy output
    const float* actions;       // [M][2]
    const float* old_logp;      // [M]
    const float* old_logits;    // [M][4] behaviour distribution inputs (may be null when kl_coeff == 0)
    const float* adv;           // [M]
    const float* v_cur[3];      // value heads: current prediction, prediction at sampling time, target
    const float* v_old[3];
    const float* v_tgt[3];
    float* dlogits;             // [M][4]
    float* dv[3];               // [M]
    double* stats;              // [8]: sum(-surr), sum vloss0..2, sum entropy, sum kl, sum logp, rows
    int M, n_heads, mode;       // mode 0: PPO loss, 1: mean(logp) (old-policy term of the meta-gradient)
    float clip, vf_clip, vf_coeff, ent_coeff, kl_coeff, inv_rows;
    int plain_vf;               // old_value_loss=False: clamp((v - t)^2, 0, vf_clip)
    const float* dyn_coeffs;    // device [2] = {kl_coeff, entropy_coeff} read at run time (a captured launch keeps working
                                // when the adaptive KL coefficient changes), or null: the by-value fields above
};

__global__ void ppo_head_kernel(const PpoHeadArgs a) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (m < a.M) {
        float4 l = reinterpret_cast<const float4*>(a.logits)[m];
        float2 act = reinterpret_cast<const float2*>(a.actions)[m];
        float s0 = expf(l.z), s1 = expf(l.w);
        float z0 = (act.x - l.x) / s0, z1 = (act.y - l.y) / s1;
        float logp = -0.5f * (z0 * z0 + z1 * z1) - 1.8378770664093453f - (l.z + l.w);
        float g_mu0, g_mu1, g_ls0, g_ls1;
        const float s = a.inv_rows;
        st[6] = logp; st[7] = 1.0;
        if (a.mode == 1) {
            g_mu0 = s * z0 / s0; g_mu1 = s * z1 / s1;
            g_ls0 = s * (z0 * z0 - 1.0f); g_ls1 = s * (z1 * z1 - 1.0f);
        } else {
            const float kl_coeff = a.dyn_coeffs ? a.dyn_coeffs[0] : a.kl_coeff;
            const float ent_coeff = a.dyn_coeffs ? a.dyn_coeffs[1] : a.ent_coeff;
            float ratio = expf(logp - a.old_logp[m]);
            float adv = a.adv[m];
            float lo = 1.0f - a.clip, hi = 1.0f + a.clip;
            float rc = fminf(fmaxf(ratio, lo), hi);
            float t1 = adv * ratio, t2 = adv * rc;
            float surr = fminf(t1, t2);
            bool inside = (ratio >= lo) && (ratio <= hi);
            float dsurr = (inside || t1 < t2) ? adv : 0.0f;       // d surr / d ratio
            float c = -s * dsurr * ratio;                           // d(total)/d logp
            g_mu0 = c * z0 / s0; g_mu1 = c * z1 / s1;
            g_ls0 = c * (z0 * z0 - 1.0f) - s * ent_coeff;
            g_ls1 = c * (z1 * z1 - 1.0f) - s * ent_coeff;
            st[0] = -surr;
            st[4] = (l.z + l.w) + 2.8378770664093453f;             // entropy: sum(ls) + log(2 pi e)
            if (kl_coeff > 0.0f) {
                float4 o = reinterpret_cast<const float4*>(a.old_logits)[m];
                float so0 = expf(o.z), so1 = expf(o.w);
                float d0 = o.x - l.x, d1 = o.y - l.y;
                float q0 = (so0 * so0 + d0 * d0) / (s0 * s0), q1 = (so1 * so1 + d1 * d1) / (s1 * s1);
                st[5] = (l.z - o.z + 0.5f * q0 - 0.5f) + (l.w - o.w + 0.5f * q1 - 0.5f);
                float k = s * kl_coeff;
                g_mu0 += k * (l.x - o.x) / (s0 * s0); g_mu1 += k * (l.y - o.y) / (s1 * s1);
                g_ls0 += k * (1.0f - q0); g_ls1 += k * (1.0f - q1);
            }
            for (int hd = 0; hd < a.n_heads; ++hd) {
                float v = a.v_cur[hd][m], vo = a.v_old[hd][m], t = a.v_tgt[hd][m];
                float l1 = (v - t) * (v - t);
                if (a.plain_vf) {
                    // torch.clamp passes the gradient where 0 <= x <= max (inclusive)
                    a.dv[hd][m] = (l1 <= a.vf_clip) ? s * a.vf_coeff * 2.0f * (v - t) : 0.0f;
                    st[1 + hd] = fminf(l1, a.vf_clip);
                    continue;
                }
                float dvc = fminf(fmaxf(v - vo, -a.vf_clip), a.vf_clip);
                float vc = vo + dvc;
                float l2 = (vc - t) * (vc - t);
                bool pass = (v - vo >= -a.vf_clip) && (v - vo <= a.vf_clip);
                float g;
                if (l1 > l2) g = 2.0f * (v - t);
                else if (l2 > l1) g = pass ? 2.0f * (vc - t) : 0.0f;
                else g = (v - t) + (pass ? (vc - t) : 0.0f);      // tie: autograd splits the gradient in half
                a.dv[hd][m] = s * a.vf_coeff * g;
                st[1 + hd] = fmaxf(l1, l2);
            }
        }
        reinterpret_cast<float4*>(a.dlogits)[m] = make_float4(g_mu0, g_mu1, g_ls0, g_ls1);
    }
    // block reduction of the statistics (double, one atomic per warp)
#pragma unroll
    for (int q = 0; q < 8; ++q) st[q] = warp_sum_d(st[q]);
    if ((threadIdx.x & 31) == 0 && a.stats) {
#pragma unroll
        for (int q = 0; q < 8; ++q) if (st[q] != 0.0) atomicAdd(&a.stats[q], st[q]);
    }
}

// sums over rows of: c = cos(phi) adv + sin(phi) nei, d = (-sin(phi) adv + cos(phi) nei) * pi/2, d * eps
// with phi = (mean + std * eps) * pi/2
__global__ void lcf_meta_terms_kernel(const float* __restrict__ adv, const float* __restrict__ nei,
                                      const float* __restrict__ eps, int M, float mean, float std,
                                      const float* __restrict__ params, double* __restrict__ out,
                                      const float* __restrict__ gadv = nullptr) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0, d = 0, de = 0, ga = 0;
    if (params) {
        // CoPOModel.lcf_mean / lcf_std from the raw parameters (algo_copo.py:171-177)
        mean = fminf(fmaxf(tanhf(params[0]), -1.0f + 1e-6f), 1.0f - 1e-6f);
        std = expf(fminf(fmaxf(params[1], -20.0f), 2.0f));
    }
    if (m < M) {
        float e = eps[m];
        float phi = (mean + std * e) * 1.57079632679489661923f;
        float sn, cs;
        sincosf(phi, &sn, &cs);
        float a = adv[m], n = nei[m];
        c = cs * a + sn * n;
        float dd = (-sn * a + cs * n) * 1.57079632679489661923f;
        d = dd; de = dd * e;
        if (gadv) ga = gadv[m];
    }
    c = warp_sum_d(c); d = warp_sum_d(d); de = warp_sum_d(de);
    if (gadv) ga = warp_sum_d(ga);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], c); atomicAdd(&out[1], d); atomicAdd(&out[2], de);
        if (gadv) atomicAdd(&out[3], ga);                        // sum of the global advantage (a logged statistic)
    }
}

// Everything CoPOPolicy.meta_update does after the two policy gradients and their dot product (algo_copo.py:264-309), in
// one thread: the LCF advantage loss, d(loss)/d(lcf_parameters) through lcf_mean = clamp(tanh(p0)) and
// lcf_std = exp(clamp(p1)) (algo_copo.py:171-177), the Adam step on the two parameters and the 13 logged statistics.
// Float widths follow the torch expression it replaces: sums and losses in double, the parameters, their derivatives
// and Adam in float.
struct LcfFinishArgs {
    const double* grad_value; const double* st_new; const double* st_old; const double* sums;
    double rows, raw_mean, raw_std;
    float* params; float* m; float* v; float* grad; double* stats;
    float step_size, inv_sqrt_bc2, beta1, beta2, eps;
};
__global__ void lcf_meta_finish_kernel(const LcfFinishArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double gv = a.grad_value[0];
    const double coordinated = a.sums[0] / a.rows, t1 = a.sums[1] / a.rows, t2 = a.sums[2] / a.rows;
    const double adv_loss = (coordinated - a.raw_mean) / a.raw_std;
    const float p0 = a.params[0], p1 = a.params[1];
    const float th = tanhf(p0);
    const float dmean = (fabsf(th) < (float)(1.0 - 1e-6)) ? 1.0f - th * th : 0.0f;
    const float std_t = expf(fminf(fmaxf(p1, -20.0f), 2.0f));
    const float dstd = (p1 > -20.0f && p1 < 2.0f) ? std_t : 0.0f;
    const double dl0 = t1 * (double)dmean / a.raw_std, dl1 = t2 * (double)dstd / a.raw_std;
    const float g[2] = {(float)(gv * dl0), (float)(gv * dl1)};
    float pn[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {                                // adam_kernel's arithmetic
        a.grad[k] = g[k];
        const float mi = a.beta1 * a.m[k] + (1.0f - a.beta1) * g[k];
        const float vi = a.beta2 * a.v[k] + (1.0f - a.beta2) * g[k] * g[k];
        a.m[k] = mi; a.v[k] = vi;
        const float denom = sqrtf(vi) * a.inv_sqrt_bc2 + a.eps;
        pn[k] = a.params[k] - a.step_size * (mi / denom);
        a.params[k] = pn[k];
    }
    const float lm = fminf(fmaxf(tanhf(pn[0]), -1.0f + 1e-6f), 1.0f - 1e-6f);
    const float ls = expf(fminf(fmaxf(pn[1], -20.0f), 2.0f));
    double* o = a.stats;
    o[0] = a.st_new[0] / a.rows;  o[1] = a.st_old[6] / a.rows;  o[2] = adv_loss;  o[3] = gv * adv_loss;  o[4] = gv;
    o[5] = lm;  o[6] = (double)(lm * 90.0f);  o[7] = pn[0];  o[8] = coordinated;  o[9] = a.sums[3] / a.rows;
    o[10] = ls;  o[11] = (double)(ls * 90.0f);  o[12] = pn[1];
}

}  // namespace b2c

using namespace b2c;

extern "C" {

int b2c_head_forward(const float* h, int ldh, const float* W, const float* b, float* y, int ldy, int M, int K, int N,
                     void* stream) {
    if (M == 0) return B2C_OK;
    if (!h || !W || !y || N < 1 || N > HEAD_MAX_N || K < 1 || K > 4096)
        return b2c_set_error(B2C_ERR_ARG, "b2c_head_forward: N must be in [1, 8], K in [1, 4096]");
    int grid = (M + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    size_t smem = (size_t)N * K * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
#define B2C_HF(n) case n: head_forward_kernel<n><<<grid, 256, smem, s>>>(h, ldh, W, b, y, ldy, M, K); break;
    switch (N) { B2C_HF(1) B2C_HF(2) B2C_HF(3) B2C_HF(4) B2C_HF(5) B2C_HF(6) B2C_HF(7) B2C_HF(8) }
#undef B2C_HF
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_head_backward(const float* dy, int ldy, const float* h, int ldh, const float* W, float* dz, int ldz, float* dW,
                      float* db, int M, int K, int N, int dtanh, void* stream) {
    return b2c_head_backward_split(dy, ldy, h, ldh, W, dz, ldz, nullptr, dW, db, M, K, N, dtanh, stream);
}

int b2c_head_backward_split(const float* dy, int ldy, const float* h, int ldh, const float* W, float* dz, int ldz,
                            uint16_t* dz_split, float* dW, float* db, int M, int K, int N, int dtanh, void* stream) {
    return b2c_head_backward_tc(dy, ldy, h, ldh, W, dz, ldz, dz_split, dW, db, nullptr, M, K, N, dtanh, stream);
}

int b2c_head_backward_tc(const float* dy, int ldy, const float* h, int ldh, const float* W, float* dz, int ldz,
                         uint16_t* dz_split, float* dW, float* db, float* dz_colsum, int M, int K, int N, int dtanh,
                         void* stream) {
    if (M == 0) return B2C_OK;
    if (!dy || !h || !W || N < 1 || N > HEAD_MAX_N) return b2c_set_error(B2C_ERR_ARG, "b2c_head_backward: bad argument");
    if (dz_split && (K % 64)) return b2c_set_error(B2C_ERR_ARG, "b2c_head_backward: dz_split needs K to be a multiple of 64");
    // row chunk per CTA: enough CTAs to fill the machine several times over (the kernel streams h in and dz out; with
    // 256-row chunks a 65 536-row minibatch gave 1.7 CTAs per SM and the loads' latency was exposed)
    int rows = (M + 148 * 8 - 1) / (148 * 8);
    rows = (rows + 3) & ~3;
    rows = rows < 32 ? 32 : (rows > 256 ? 256 : rows);
    dim3 grid((M + rows - 1) / rows, (K + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
#define B2C_HB(n) case n: head_backward_kernel<n><<<grid, 256, 0, s>>>(dy, ldy, h, ldh, W, dz, ldz, dz_split, dW, db, dz_colsum, M, K, rows, dtanh); break;
    switch (N) { B2C_HB(1) B2C_HB(2) B2C_HB(3) B2C_HB(4) B2C_HB(5) B2C_HB(6) B2C_HB(7) B2C_HB(8) }
#undef B2C_HB
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_gaussian_sample(const float* logits, const float* eps_in, float* actions, float* logp, float* eps_out, int M,
                        uint32_t seed, uint32_t step, int deterministic, void* stream) {
    if (M == 0) return B2C_OK;
    if (!logits || !actions) return b2c_set_error(B2C_ERR_ARG, "b2c_gaussian_sample: null argument");
    gaussian_sample_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(logits, eps_in, actions, logp, eps_out, M,
                                                                              seed, step, deterministic);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_ppo_head(const b2c_ppo_head_args* p, void* stream) {
    if (p && p->rows == 0) return B2C_OK;
    if (!p || !p->logits || !p->actions || !p->dlogits || p->n_heads < 0 || p->n_heads > 3)
        return b2c_set_error(B2C_ERR_ARG, "b2c_ppo_head: bad argument");
    if (p->mode == 0 && (!p->old_logp || !p->adv)) return b2c_set_error(B2C_ERR_ARG, "b2c_ppo_head: missing columns");
    if (p->mode == 0 && (p->kl_coeff > 0.0f || p->dyn_coeffs) && !p->old_logits)
        return b2c_set_error(B2C_ERR_ARG, "b2c_ppo_head: kl_coeff > 0 needs the behaviour distribution inputs");
    PpoHeadArgs a;
    a.logits = p->logits; a.actions = p->actions; a.old_logp = p->old_logp; a.old_logits = p->old_logits; a.adv = p->adv;
    for (int h = 0; h < 3; ++h) {
        a.v_cur[h] = p->v_cur[h]; a.v_old[h] = p->v_old[h]; a.v_tgt[h] = p->v_tgt[h]; a.dv[h] = p->dv[h];
        if (h < p->n_heads && (!a.v_cur[h] || !a.v_old[h] || !a.v_tgt[h] || !a.dv[h]))
            return b2c_set_error(B2C_ERR_ARG, "b2c_ppo_head: value head %d has a null column", h);
    }
    a.dlogits = p->dlogits; a.stats = p->stats; a.M = p->rows; a.n_heads = p->n_heads; a.mode = p->mode;
    a.clip = p->clip_param; a.vf_clip = p->vf_clip_param; a.vf_coeff = p->vf_loss_coeff; a.ent_coeff = p->entropy_coeff;
    a.kl_coeff = p->kl_coeff; a.inv_rows = 1.0f / (float)(p->norm_rows > 0 ? p->norm_rows : p->rows);
    a.plain_vf = p->plain_value_loss ? 1 : 0;
    a.dyn_coeffs = p->dyn_coeffs;
    ppo_head_kernel<<<(p->rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_lcf_meta_terms(const float* adv, const float* nei_adv, const float* eps, int rows, float lcf_mean, float lcf_std,
                       double* out3, void* stream) {
    if (rows == 0) return B2C_OK;
    if (!adv || !nei_adv || !eps || !out3) return b2c_set_error(B2C_ERR_ARG, "b2c_lcf_meta_terms: null argument");
    lcf_meta_terms_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(adv, nei_adv, eps, rows, lcf_mean, lcf_std,
                                                                               nullptr, out3);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_lcf_meta_sums(const float* adv, const float* nei_adv, const float* eps, const float* global_adv, int rows,
                      const float* lcf_parameters, double* out4, void* stream) {
    if (rows == 0) return B2C_OK;
    if (!adv || !nei_adv || !eps || !global_adv || !out4 || !lcf_parameters)
        return b2c_set_error(B2C_ERR_ARG, "b2c_lcf_meta_sums: null argument");
    lcf_meta_terms_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(adv, nei_adv, eps, rows, 0.0f, 1.0f,
                                                                               lcf_parameters, out4, global_adv);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_lcf_meta_finish(const b2c_lcf_meta_finish_args* p, void* stream) {
    if (!p || !p->grad_value || !p->st_new || !p->st_old || !p->sums || !p->lcf_parameters || !p->exp_avg || !p->exp_avg_sq ||
        !p->lcf_grad || !p->stats || p->step < 1 || !(p->rows > 0.0) || !(p->raw_std > 0.0))
        return b2c_set_error(B2C_ERR_ARG, "b2c_lcf_meta_finish: bad argument");
    LcfFinishArgs a;
    a.grad_value = p->grad_value; a.st_new = p->st_new; a.st_old = p->st_old; a.sums = p->sums;
    a.rows = p->rows; a.raw_mean = p->raw_mean; a.raw_std = p->raw_std;
    a.params = p->lcf_parameters; a.m = p->exp_avg; a.v = p->exp_avg_sq; a.grad = p->lcf_grad; a.stats = p->stats;
    const double bc1 = 1.0 - pow((double)p->beta1, p->step), bc2 = 1.0 - pow((double)p->beta2, p->step);
    a.step_size = (float)(p->lr / bc1); a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    a.beta1 = p->beta1; a.beta2 = p->beta2; a.eps = p->eps;
    lcf_meta_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_lcf_meta_terms_params(const float* adv, const float* nei_adv, const float* eps, int rows,
                              const float* lcf_parameters, double* out3, void* stream) {
    if (rows == 0) return B2C_OK;
    if (!adv || !nei_adv || !eps || !out3 || !lcf_parameters)
        return b2c_set_error(B2C_ERR_ARG, "b2c_lcf_meta_terms_params: null argument");
    lcf_meta_terms_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(adv, nei_adv, eps, rows, 0.0f, 1.0f,
                                                                               lcf_parameters, out3);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

}  // extern "C"
