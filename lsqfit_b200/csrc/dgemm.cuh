// FP64 GEMM on the DMMA path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) for the dense side of the
// hot path: W.[G|delta], J^T J, the block-Jacobi whitening updates and the fit.p propagation
// D = (cov J^T) W, cov(p) = (D C) D^T  (reference src/lsqfit/__init__.py:897-922 does these with
// Python loops over GVars; src/lsqfit/_utilities.pyx:20-36 `dot` row by row).
//
// tcgen05/TMEM have no FP64 kind, so this is the fastest FP64 matrix path on sm_100a; the roofline
// is the shared FP64 pipe (37 TFLOP/s measured, profiles/fp64_peak_r01.json).
//
//   C[b] (M x N) = alpha * opA(A[b]) . opB(B[b]) + beta * C[b]          all row-major, fp64
//   A_KC = true : A is stored [M][K] (k contiguous)      false: A is stored [K][M]  (C = A^T ...)
//   B_KC = true : B is stored [N][K] (k contiguous, i.e. C = ... B^T)    false: B is stored [K][N]
//
// CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 DMMA tiles (64 accumulator
// doubles per lane), 3-stage cp.async (LDGSTS, 16 B) pipeline.  Shared-memory pitches are chosen so
// that every fragment load touches each bank pair exactly twice (the minimum for 256 B):
//   k-contiguous operand  : [128][16 + 4]   (pitch == 4 mod 16)
//   mn-contiguous operand : [16][128 + 8]   (pitch == 8 mod 16)
// Requirements of the fast path: leading dimensions even and base pointers 16-byte aligned (the
// launcher falls back to a scalar-copy variant otherwise).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200lm {

constexpr int GM = 128, GN = 128, GK = 16, GSTAGES = 3, GTHREADS = 256;
constexpr int PITCH_KC = GK + 4;        // 20
constexpr int PITCH_MN = GM + 8;        // 136
constexpr int TILE_KC = GM * PITCH_KC;  // 2560 doubles
constexpr int TILE_MN = GK * PITCH_MN;  // 2176 doubles

struct GemmArgs {
    int M, N, K;
    double alpha, beta;
    const double* A; long long sA; int lda;
    const double* B; long long sB; int ldb;
    double* C; long long sC; int ldc;
    int lower;   // 1: skip CTA tiles strictly above the diagonal (symmetric rank-k updates)
};

__device__ __forceinline__ void gemm_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

// Copy one operand tile into shared memory.  KC: rows = m (or n) index, 16 k per row.
// MN: rows = k index, 128 m (or n) per row.  ALIGNED: 16-byte cp.async with zero fill, else scalar.
template <bool KC, bool ALIGNED>
__device__ __forceinline__ void load_tile(double* dst, const double* src, int ld, int mn0, int k0,
                                          int MN, int K, int tid) {
    if (KC) {
        // 128 rows x 8 chunks of 2 doubles
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int c = tid + it * GTHREADS;
            const int r = c >> 3, ch = c & 7;
            const int mn = mn0 + r, k = k0 + 2 * ch;
            double* d = dst + r * PITCH_KC + 2 * ch;
            if (ALIGNED) {
                int bytes = 0;
                if (mn < MN) bytes = k + 1 < K ? 16 : (k < K ? 8 : 0);
                const double* s = bytes ? src + (size_t)mn * ld + k : src;
                cp_async16(d, s, bytes);
            } else {
                d[0] = (mn < MN && k < K) ? src[(size_t)mn * ld + k] : 0.0;
                d[1] = (mn < MN && k + 1 < K) ? src[(size_t)mn * ld + k + 1] : 0.0;
            }
        }
    } else {
        // 16 rows x 64 chunks of 2 doubles
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int c = tid + it * GTHREADS;
            const int r = c >> 6, ch = c & 63;
            const int k = k0 + r, mn = mn0 + 2 * ch;
            double* d = dst + r * PITCH_MN + 2 * ch;
            if (ALIGNED) {
                int bytes = 0;
                if (k < K) bytes = mn + 1 < MN ? 16 : (mn < MN ? 8 : 0);
                const double* s = bytes ? src + (size_t)k * ld + mn : src;
                cp_async16(d, s, bytes);
            } else {
                d[0] = (k < K && mn < MN) ? src[(size_t)k * ld + mn] : 0.0;
                d[1] = (k < K && mn + 1 < MN) ? src[(size_t)k * ld + mn + 1] : 0.0;
            }
        }
    }
}

template <bool A_KC, bool B_KC, bool ALIGNED>
__global__ void __launch_bounds__(GTHREADS, 1) dgemm_kernel(const __grid_constant__ GemmArgs g) {
    extern __shared__ double gsm[];
    constexpr int TA = A_KC ? TILE_KC : TILE_MN;
    constexpr int TB = B_KC ? TILE_KC : TILE_MN;
    double* As = gsm;
    double* Bs = gsm + GSTAGES * TA;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;              // 2 x 4 warps
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    if (g.lower && n0 > m0 + GM - 1) return;
    const double* A = g.A + (size_t)blockIdx.z * g.sA;
    const double* B = g.B + (size_t)blockIdx.z * g.sB;
    double* C = g.C + (size_t)blockIdx.z * g.sC;
    const int nk = (g.K + GK - 1) / GK;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // prologue: stages 0 .. GSTAGES-2
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
        if (s < nk) {
            load_tile<A_KC, ALIGNED>(As + s * TA, A, g.lda, m0, s * GK, g.M, g.K, tid);
            load_tile<B_KC, ALIGNED>(Bs + s * TB, B, g.ldb, n0, s * GK, g.N, g.K, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        // prefetch tile kt + GSTAGES - 1 into the stage consumed at iteration kt - 1
        {
            const int kn = kt + GSTAGES - 1;
            if (kn < nk) {
                const int s = kn % GSTAGES;
                load_tile<A_KC, ALIGNED>(As + s * TA, A, g.lda, m0, kn * GK, g.M, g.K, tid);
                load_tile<B_KC, ALIGNED>(Bs + s * TB, B, g.ldb, n0, kn * GK, g.N, g.K, tid);
            }
            cp_async_commit();
        }
        const double* as = As + (kt % GSTAGES) * TA;
        const double* bs = Bs + (kt % GSTAGES) * TB;
#pragma unroll
        for (int s4 = 0; s4 < GK; s4 += 4) {
            double af[8], bf[4];
            // A fragment: A[m = 64 wm + 8 i + lane/4][k = s4 + lane%4]
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = 64 * wm + 8 * i + (lane >> 2), k = s4 + (lane & 3);
                af[i] = A_KC ? as[m * PITCH_KC + k] : as[k * PITCH_MN + m];
            }
            // B fragment: B[k = s4 + lane%4][n = 32 wn + 8 j + lane/4]
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 32 * wn + 8 * j + (lane >> 2), k = s4 + (lane & 3);
                bf[j] = B_KC ? bs[n * PITCH_KC + k] : bs[k * PITCH_MN + n];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) gemm_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    // epilogue: C[m][n], m = 64 wm + 8 i + lane/4, n = 32 wn + 8 j + 2 (lane%4) + {0,1}
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + 64 * wm + 8 * i + (lane >> 2);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 32 * wn + 8 * j + 2 * (lane & 3);
            double* cp = C + (size_t)m * g.ldc + n;
            if (n < g.N) cp[0] = g.beta != 0.0 ? fma(g.alpha, acc[i][j][0], g.beta * cp[0]) : g.alpha * acc[i][j][0];
            if (n + 1 < g.N) cp[1] = g.beta != 0.0 ? fma(g.alpha, acc[i][j][1], g.beta * cp[1]) : g.alpha * acc[i][j][1];
        }
    }
}

// Many small products (batched fit.p propagation: M = np <= 32 rows per fit): the 128 x 128 tile above would
// be > 85 % padding.  CTA tile 16 x 64 x 16, 4 warps (warp w: both 8-row tiles x column tiles 2w, 2w+1),
// plain loads through shared memory -- thousands of CTAs per launch hide the latency.
constexpr int SM_M = 16, SM_N = 64, SM_K = 16, SM_THREADS = 128;
static __global__ void __launch_bounds__(SM_THREADS) dgemm_small_kernel(const __grid_constant__ GemmArgs g, int transA, int transB) {
    __shared__ double As[SM_M][SM_K + 4];          // [m][k]
    __shared__ double Bs[SM_K][SM_N + 8];          // [k][n]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int m0 = blockIdx.y * SM_M, n0 = blockIdx.x * SM_N;
    const double* A = g.A + (size_t)blockIdx.z * g.sA;
    const double* B = g.B + (size_t)blockIdx.z * g.sB;
    double* C = g.C + (size_t)blockIdx.z * g.sC;
    double acc[2][2][2] = {};
    for (int k0 = 0; k0 < g.K; k0 += SM_K) {
        for (int e = tid; e < SM_M * SM_K; e += SM_THREADS) {
            const int m = e / SM_K, k = e % SM_K;
            const int gm = m0 + m, gk = k0 + k;
            As[m][k] = (gm < g.M && gk < g.K) ? (transA ? A[(size_t)gk * g.lda + gm] : A[(size_t)gm * g.lda + gk]) : 0.0;
        }
        for (int e = tid; e < SM_K * SM_N; e += SM_THREADS) {
            int k, n;
            if (transB) { n = e / SM_K; k = e % SM_K; } else { k = e / SM_N; n = e % SM_N; }   // contiguous global reads
            const int gk = k0 + k, gn = n0 + n;
            Bs[k][n] = (gk < g.K && gn < g.N) ? (transB ? B[(size_t)gn * g.ldb + gk] : B[(size_t)gk * g.ldb + gn]) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int s4 = 0; s4 < SM_K; s4 += 4) {
            double af[2], bf[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = As[8 * i + (lane >> 2)][s4 + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[j] = Bs[s4 + (lane & 3)][16 * w + 8 * j + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) gemm_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + 8 * i + (lane >> 2);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = n0 + 16 * w + 8 * j + 2 * (lane & 3);
            double* cp = C + (size_t)m * g.ldc + n;
            if (n < g.N) cp[0] = g.beta != 0.0 ? fma(g.alpha, acc[i][j][0], g.beta * cp[0]) : g.alpha * acc[i][j][0];
            if (n + 1 < g.N) cp[1] = g.beta != 0.0 ? fma(g.alpha, acc[i][j][1], g.beta * cp[1]) : g.alpha * acc[i][j][1];
        }
    }
}

template <bool A_KC, bool B_KC>
inline size_t dgemm_smem_bytes() {
    return (size_t)GSTAGES * ((A_KC ? TILE_KC : TILE_MN) + (B_KC ? TILE_KC : TILE_MN)) * sizeof(double);
}

// transA: C = A^T . (...) with A stored [K][M];  transB: C = (...) . B^T with B stored [N][K]
cudaError_t dgemm(bool transA, bool transB, int batch, int M, int N, int K, double alpha,
                  const double* A, long long sA, int lda, const double* B, long long sB, int ldb,
                  double beta, double* C, long long sC, int ldc, cudaStream_t stream, bool lower = false);

}  // namespace b200lm
