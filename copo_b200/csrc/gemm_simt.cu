// gemm_simt.cu - fp32 CUDA-core GEMMs for the 256-wide layers of the policy / value MLPs (rows a7, a14, a16-a18 of
// SURVEY.md section 8): forward  y = act(x W^T + b), input gradient  dx = (dy W) * (1 - h^2), weight gradient
// dW += dy^T x (reduction over the batch, split over CTAs, fp32 atomics), bias gradient db += colsum(dy).
// These are the exact-fp32 path (parity with the torch-CPU oracle to ~1e-6); mlp_tc.cu holds the tcgen05 path.
//
// Replaces: SlimFC layers of CCModel / CoPOModel (torch_copo/algo_ccppo.py:108-170, algo_copo.py:138-147) and
// their autograd backward inside {IPPO,CCPPO,CoPO}Policy.loss / meta_update.
#include <cuda_runtime.h>
#include <stdint.h>
#include "b2c_internal.h"

namespace b2c {

// C[M x N] = opA(A) * opB(B), tile 128 x 128 x 8, 256 threads, 8 x 8 outputs per thread.
//   TA = false: A is [M x K] row-major (lda);  TA = true: A is [K x M] row-major (reduction dim slow)
//   TB = false: B is [K x N] row-major (ldb);  TB = true: B is [N x K] row-major
// grid.z splits the reduction dimension; with SPLIT the partial tiles are added with atomics.
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_TANH = 2, EPI_DTANH = 3, EPI_ATOMIC = 4 };

constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8, NTHREADS = 256;

template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(NTHREADS)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
             int M, int N, int K, const float* __restrict__ bias, const float* __restrict__ H, int ldh, int k_chunk) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int tx = tid % 16, ty = tid / 16;          // 16 x 16 threads, each 8 x 8

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    // global -> register staging: each thread moves 4 A values and 4 B values per k-tile
    float ra[4], rb[4];
    auto load_tiles = [&](int k0) {
        if (!TA) {   // A[m][k]: thread -> (row = tid / 2, k = (tid % 2) * 4 .. +3)
            int r = tid >> 1, kk = (tid & 1) * 4;
            int m = m0 + r;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int k = k0 + kk + q;
                ra[q] = (m < M && k < k_end) ? A[(size_t)m * lda + k] : 0.0f;
            }
        } else {     // A[k][m]: thread -> (k = tid / 32, m = (tid % 32) * 4 .. +3)
            int kk = tid >> 5, mm = (tid & 31) * 4;
            int k = k0 + kk;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int m = m0 + mm + q;
                ra[q] = (m < M && k < k_end) ? A[(size_t)k * lda + m] : 0.0f;
            }
        }
        if (TB) {    // B[n][k]
            int r = tid >> 1, kk = (tid & 1) * 4;
            int n = n0 + r;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int k = k0 + kk + q;
                rb[q] = (n < N && k < k_end) ? B[(size_t)n * ldb + k] : 0.0f;
            }
        } else {     // B[k][n]
            int kk = tid >> 5, nn = (tid & 31) * 4;
            int k = k0 + kk;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int n = n0 + nn + q;
                rb[q] = (n < N && k < k_end) ? B[(size_t)k * ldb + n] : 0.0f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
        if (!TA) {
            int r = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) As[buf][kk + q][r] = ra[q];
        } else {
            int kk = tid >> 5, mm = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&As[buf][kk][mm]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
        }
        if (TB) {
            int r = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) Bs[buf][kk + q][r] = rb[q];
        } else {
            int kk = tid >> 5, nn = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        }
    };

    int buf = 0;
    if (k_begin < k_end) {
        load_tiles(k_begin);
        store_tiles(0);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = (k0 + BK) < k_end;
        if (more) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    // epilogue: rows ty*4..+3 and 64+ty*4..+3, cols tx*4..+3 and 64+tx*4..+3
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= N) continue;
            float v = acc[i][j];
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_TANH) v += bias[n];
            if (EPI == EPI_BIAS_TANH) v = tanhf(v);
            if (EPI == EPI_DTANH) { float h = H[(size_t)m * ldh + n]; v *= (1.0f - h * h); }
            if (EPI == EPI_ATOMIC) atomicAdd(&C[(size_t)m * ldc + n], v);
            else C[(size_t)m * ldc + n] = v;
        }
    }
}

// db[n] += sum_m dy[m][n]   (N <= 1024; one CTA per row chunk, coalesced over n)
__global__ void colsum_kernel(const float* __restrict__ dy, int ldy, float* __restrict__ db, int M, int N, int rows) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int m_begin = blockIdx.y * rows, m_end = min(M, m_begin + rows);
    if (n >= N) return;
    // four independent partial sums: four loads in flight per thread
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int m = m_begin;
    for (; m + 3 < m_end; m += 4) {
        s0 += dy[(size_t)m * ldy + n];
        s1 += dy[(size_t)(m + 1) * ldy + n];
        s2 += dy[(size_t)(m + 2) * ldy + n];
        s3 += dy[(size_t)(m + 3) * ldy + n];
    }
    for (; m < m_end; ++m) s0 += dy[(size_t)m * ldy + n];
    atomicAdd(&db[n], (s0 + s1) + (s2 + s3));
}

}  // namespace b2c

using namespace b2c;

extern "C" {

int b2c_linear_forward(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int M, int K, int N,
                       int act, void* stream) {
    if (M == 0) return B2C_OK;
    if (!x || !W || !y || M < 0 || K < 1 || N < 1) return b2c_set_error(B2C_ERR_ARG, "b2c_linear_forward: bad argument");
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
    cudaStream_t s = (cudaStream_t)stream;
    if (act == 1 && b) sgemm_kernel<false, true, EPI_BIAS_TANH><<<grid, NTHREADS, 0, s>>>(x, ldx, W, K, y, ldy, M, N, K, b, nullptr, 0, K);
    else if (act == 0 && b) sgemm_kernel<false, true, EPI_BIAS><<<grid, NTHREADS, 0, s>>>(x, ldx, W, K, y, ldy, M, N, K, b, nullptr, 0, K);
    else if (act == 0) sgemm_kernel<false, true, EPI_NONE><<<grid, NTHREADS, 0, s>>>(x, ldx, W, K, y, ldy, M, N, K, nullptr, nullptr, 0, K);
    else return b2c_set_error(B2C_ERR_ARG, "b2c_linear_forward: tanh needs a bias");
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_linear_backward_input(const float* dy, int ldy, const float* W, const float* h_prev, int ldh, float* dx, int ldx,
                              int M, int K, int N, void* stream) {
    if (M == 0) return B2C_OK;
    if (!dy || !W || !dx || M < 0) return b2c_set_error(B2C_ERR_ARG, "b2c_linear_backward_input: bad argument");
    // dx[M x K] = dy[M x N] * W[N x K]   (reduction over N)
    dim3 grid((K + BN - 1) / BN, (M + BM - 1) / BM, 1);
    cudaStream_t s = (cudaStream_t)stream;
    if (h_prev) sgemm_kernel<false, false, EPI_DTANH><<<grid, NTHREADS, 0, s>>>(dy, ldy, W, K, dx, ldx, M, K, N, nullptr, h_prev, ldh, N);
    else sgemm_kernel<false, false, EPI_NONE><<<grid, NTHREADS, 0, s>>>(dy, ldy, W, K, dx, ldx, M, K, N, nullptr, nullptr, 0, N);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_linear_backward_weight(const float* dy, int ldy, const float* x, int ldx, float* dW, float* db, int M, int K,
                               int N, void* stream) {
    if (M == 0) return B2C_OK;
    if (!dy || !x || !dW || M < 0) return b2c_set_error(B2C_ERR_ARG, "b2c_linear_backward_weight: bad argument");
    // dW[N x K] += dy^T[N x M] * x[M x K]   (reduction over M, split over grid.z, atomics into dW)
    int chunk = 2048;
    int splits = (M + chunk - 1) / chunk;
    if (splits > 65535) { chunk = (M + 65534) / 65535; chunk = (chunk + BK - 1) / BK * BK; splits = (M + chunk - 1) / chunk; }
    dim3 grid((K + BN - 1) / BN, (N + BM - 1) / BM, splits);
    cudaStream_t s = (cudaStream_t)stream;
    sgemm_kernel<true, false, EPI_ATOMIC><<<grid, NTHREADS, 0, s>>>(dy, ldy, x, ldx, dW, K, N, K, M, nullptr, nullptr, 0, chunk);
    B2C_CUDA(cudaGetLastError());
    if (db) {
        int rows = 64;                                   // many short CTAs: the sum is bandwidth-bound
        dim3 g2((N + 127) / 128, (M + rows - 1) / rows);
        colsum_kernel<<<g2, 128, 0, s>>>(dy, ldy, db, M, N, rows);
        B2C_CUDA(cudaGetLastError());
    }
    return B2C_OK;
}

int b2c_colsum(const float* dy, int ldy, float* db, int M, int N, void* stream) {
    if (M == 0) return B2C_OK;
    if (!dy || !db || M < 0 || N < 1) return b2c_set_error(B2C_ERR_ARG, "b2c_colsum: bad argument");
    int rows = 64;
    dim3 g2((N + 127) / 128, (M + rows - 1) / rows);
    colsum_kernel<<<g2, 128, 0, (cudaStream_t)stream>>>(dy, ldy, db, M, N, rows);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

}  // extern "C"
