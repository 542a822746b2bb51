// svdcut / eps whitening of the correlated blocks of cov(y (+) prior), batched over
// blocks: one CTA per block, parallel (round-robin) two-sided Jacobi eigen-solver with
// the matrix and the eigenvectors resident in shared memory (blocks up to 112x112;
// blocks up to 512 can run the same code out of a global-memory workspace -- used only for the eps
// regulator; svdcut blocks above 112 go to the multi-CTA block-Jacobi solver of whiten_large.cu).
//
// Replaces gvar.PDF / gvar.regulate / gvar.svd as called by the reference at
// src/lsqfit/__init__.py:1895,1898 (third-party gvar >= 13.1.5, not vendored); semantics
// per doc/source/overview.rst:1546-1611 and the layout proven by
// tests/test_lsqfit.py:923-943.  The CPU restatement is oracle/whiten.py.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>
#include <string>
#include "../../include/b200lm.h"
#include "handle.h"
#include "jacobi_core.cuh"

namespace b200lm {

constexpr int WH_THREADS = 256;
constexpr int WH_SMEM_NMAX = 112;
constexpr int WH_NMAX = 512;

struct WhitenArgs {
    const double* cov;      // concatenated blocks
    double* w;
    double* cov_out;
    int* nout;
    int* nmod;
    double* logdet;
    const int* n;           // [nblk] device copy
    const long long* off;   // [nblk] element offset of each block
    double* work;           // global workspace (2 * ld * nmax doubles per block) or null
    long long work_stride;
    double svdcut, eps;
    int use_eps;
};

__global__ void __launch_bounds__(WH_THREADS) whiten_kernel(const __grid_constant__ WhitenArgs a) {
    extern __shared__ double sm[];
    const int blk = blockIdx.x;
    const int n = a.n[blk];
    const long long off = a.off[blk];
    const double* cov = a.cov + off;
    double* Wout = a.w + off;
    double* Cout = a.cov_out + off;
    const int tid = threadIdx.x;
    const int ld = n | 1;
    __shared__ double s_c[WH_NMAX / 2], s_s[WH_NMAX / 2];
    __shared__ int s_flag[2];
    __shared__ int s_pq[WH_NMAX + 2];
    __shared__ double s_red[WH_THREADS];
    double *A, *V;
    if (n <= WH_SMEM_NMAX) { A = sm; V = sm + (size_t)n * ld; }
    else { A = a.work + (size_t)blk * a.work_stride; V = A + (size_t)n * ld; }
    // vectors (always in shared memory, after the matrices when those are in smem)
    double* vecs = (n <= WH_SMEM_NMAX) ? sm + 2 * (size_t)n * ld : sm;
    double* Dv = vecs;            // D = diag^-1/2
    double* val = vecs + n;       // eigenvalues (signed), then the values used for W
    int* order = (int*)(vecs + 2 * n);   // order[r] = eigen index of r-th largest

    for (int i = tid; i < n; i += WH_THREADS) {
        double d = fabs(cov[(size_t)i * n + i]);
        if (d == 0.0) d = 1.0;
        Dv[i] = rsqrt(d);
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += WH_THREADS) {
        const int i = e / n, j = e % n;
        A[i * ld + j] = cov[e] * Dv[i] * Dv[j];
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();

    if (a.use_eps) {
        // ---- eps regulator: corr += eps*norm_inf(corr) I ; Cholesky; W = L^-1 D ----------
        double rs = 0.0;
        for (int i = tid; i < n; i += WH_THREADS) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s += fabs(A[i * ld + j]);
            rs = fmax(rs, s);
        }
        s_red[tid] = rs;
        __syncthreads();
        for (int o = WH_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) s_red[tid] = fmax(s_red[tid], s_red[tid + o]);
            __syncthreads();
        }
        const double shift = a.eps > 0.0 ? a.eps * s_red[0] : 0.0;
        __syncthreads();
        for (int i = tid; i < n; i += WH_THREADS) A[i * ld + i] += shift;
        __syncthreads();
        // right-looking Cholesky, lower triangle of A
        __shared__ int s_bad;
        if (tid == 0) s_bad = 0;
        for (int j = 0; j < n; ++j) {
            if (tid == 0) {
                const double piv = A[j * ld + j];
                // a non-positive pivot (eps <= 0, or a shift too small for an indefinite / rank-deficient block) would
                // put NaNs into W, logdet and the corrected covariance: flag it (nmod = -(first bad pivot + 1))
                if (!(piv > 0.0) && s_bad == 0) s_bad = j + 1;
                A[j * ld + j] = sqrt(piv);
            }
            __syncthreads();
            const double ljj = A[j * ld + j];
            for (int i = j + 1 + tid; i < n; i += WH_THREADS) A[i * ld + j] /= ljj;
            __syncthreads();
            const int m = n - j - 1;
            for (int e = tid; e < m * m; e += WH_THREADS) {
                const int i = j + 1 + e / m, k = j + 1 + e % m;
                if (k <= i) A[i * ld + k] -= A[i * ld + j] * A[k * ld + j];
            }
            __syncthreads();
        }
        // column c of W = L^-1 diag(D): forward substitution, one thread per column
        for (int c = tid; c < n; c += WH_THREADS) {
            for (int i = 0; i < n; ++i) {
                double s = (i == c) ? Dv[c] : 0.0;
                if (i < c) { V[i * ld + c] = 0.0; continue; }
                for (int k = c; k < i; ++k) s -= A[i * ld + k] * V[k * ld + c];
                V[i * ld + c] = s / A[i * ld + i];
            }
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += WH_THREADS) {
            const int i = e / n, j = e % n;
            Wout[e] = V[i * ld + j];
            Cout[e] = cov[e] + ((i == j) ? shift / (Dv[i] * Dv[i]) : 0.0);
        }
        double lg = 0.0;
        for (int i = tid; i < n; i += WH_THREADS) lg += 2.0 * log(A[i * ld + i]) - 2.0 * log(Dv[i]);
        s_red[tid] = lg;
        __syncthreads();
        for (int o = WH_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) s_red[tid] += s_red[tid + o];
            __syncthreads();
        }
        if (tid == 0) {
            a.logdet[blk] = s_red[0];
            a.nout[blk] = n;
            a.nmod[blk] = s_bad ? -s_bad : (shift > 0.0 ? n : 0);
        }
        return;
    }

    // ---- parallel cyclic Jacobi on the correlation matrix ---------------------------
    jacobi_diagonalize<WH_THREADS>(A, V, n, ld, 60, s_c, s_s, s_pq, s_flag);
    // eigenvalues (signed), descending order by rank sort
    for (int i = tid; i < n; i += WH_THREADS) val[i] = A[i * ld + i];
    __syncthreads();
    for (int i = tid; i < n; i += WH_THREADS) {
        const double vi = val[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double vj = val[j];
            rank += (vj > vi) || (vj == vi && j < i);
        }
        order[rank] = i;
    }
    __syncthreads();
    const double vmax = val[order[0]];
    const double valmin = fabs(a.svdcut) * vmax;
    const bool cut = a.svdcut != 0.0;
    // count modified modes (they are the trailing ones in descending order)
    int cnt = 0;
    for (int i = tid; i < n; i += WH_THREADS) cnt += (cut && val[i] < valmin) ? 1 : 0;
    s_red[tid] = (double)cnt;
    __syncthreads();
    for (int o = WH_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) s_red[tid] += s_red[tid + o];
        __syncthreads();
    }
    const int nmod = (int)(s_red[0] + 0.5);
    __syncthreads();
    const bool drop = cut && a.svdcut < 0.0;
    const int nkeep = drop ? n - nmod : n;
    // corrected covariance (uses the SIGNED eigenvalues: oracle/whiten.py, pinned by
    // the reference's examples/y-vs-x.out nexp=1)
    for (int e = tid; e < n * n; e += WH_THREADS) {
        const int i = e / n, j = e % n;
        double s = 0.0;
        if (cut && nmod > 0) {
            if (!drop) {
                for (int r = n - nmod; r < n; ++r) {
                    const int m = order[r];
                    s += (valmin - val[m]) * V[i * ld + m] * V[j * ld + m];
                }
                s = cov[e] + s / (Dv[i] * Dv[j]);
            } else {
                for (int r = 0; r < nkeep; ++r) {
                    const int m = order[r];
                    s += val[m] * V[i * ld + m] * V[j * ld + m];
                }
                s = s / (Dv[i] * Dv[j]);
            }
        } else {
            s = cov[e];
        }
        Cout[e] = s;
    }
    // W rows, largest eigenvalue first
    double lg = 0.0;
    for (int e = tid; e < n * n; e += WH_THREADS) {
        const int r = e / n, j = e % n;
        double w = 0.0;
        if (r < nkeep) {
            const int m = order[r];
            double v = val[m];
            if (cut && !drop && v < valmin) v = valmin;
            v = fabs(v);
            w = V[j * ld + m] * Dv[j] * rsqrt(v);
            if (j == 0) lg += log(v);
        }
        Wout[e] = w;
    }
    for (int i = tid; i < n; i += WH_THREADS) lg -= 2.0 * log(Dv[i]);
    s_red[tid] = lg;
    __syncthreads();
    for (int o = WH_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) s_red[tid] += s_red[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        a.logdet[blk] = s_red[0];
        a.nout[blk] = nkeep;
        a.nmod[blk] = nmod;
    }
}

}  // namespace b200lm

using namespace b200lm;

namespace b200lm {
int whiten_large(int device, int n, const double* d_cov, double svdcut, double* d_w, double* d_cov_out,
                 int* d_nout, int* d_nmod, double* d_logdet, cudaStream_t stream);
int whiten_large_eps(int device, int n, const double* d_cov, double eps, double* d_w, double* d_cov_out,
                     int* d_nout, int* d_nmod, double* d_logdet, cudaStream_t stream);
}

extern "C" int b200lm_whiten(int device, int nblk, const int* h_n, const double* d_cov,
                             double svdcut, double eps, int use_eps,
                             double* d_w, double* d_cov_out, int* d_nout, int* d_nmod, double* d_logdet,
                             void* stream) {
    if (nblk < 0 || (nblk > 0 && (!h_n || !d_cov || !d_w || !d_cov_out || !d_nout || !d_nmod || !d_logdet)))
        return set_error(nullptr, B200LM_EINVAL, "NULL argument");
    if (nblk == 0) return B200LM_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    cudaStream_t s = (cudaStream_t)stream;
    // blocks that do not fit the shared-memory Jacobi go to the multi-CTA block-Jacobi solver (measured:
    // n = 128: 14 ms vs 307 ms, n = 512: 72 ms vs 19.6 s for the single CTA working out of global memory);
    // the single-CTA kernel still takes blocks up to WH_NMAX for the eps (Cholesky) regulator
    int large_min = getenv("B200LM_WL_MIN") ? atoi(getenv("B200LM_WL_MIN")) : WH_SMEM_NMAX;
    if (large_min > WH_NMAX) large_min = WH_NMAX;
    int nmax = 0;
    bool any_large = false;
    std::vector<long long> off(nblk);
    long long o = 0;
    for (int k = 0; k < nblk; ++k) {
        if (h_n[k] < 1) return set_error(nullptr, B200LM_EINVAL, "block size must be positive");
        if (h_n[k] > (use_eps ? WH_NMAX : large_min)) any_large = true;
        nmax = std::max(nmax, h_n[k]);
        off[k] = o;
        o += (long long)h_n[k] * h_n[k];
    }
    if (any_large) {
        // blocks beyond the single-CTA kernel: one at a time through the block-Jacobi solver
        for (int k = 0; k < nblk; ++k) {
            int rc;
            if (h_n[k] > (use_eps ? WH_NMAX : large_min)) {
                if (use_eps)
                    rc = whiten_large_eps(device, h_n[k], d_cov + off[k], eps, d_w + off[k], d_cov_out + off[k],
                                          d_nout + k, d_nmod + k, d_logdet + k, s);
                else
                rc = whiten_large(device, h_n[k], d_cov + off[k], svdcut, d_w + off[k], d_cov_out + off[k],
                                  d_nout + k, d_nmod + k, d_logdet + k, s);
            } else {
                const int one = h_n[k];
                rc = b200lm_whiten(device, 1, &one, d_cov + off[k], svdcut, eps, use_eps, d_w + off[k],
                                   d_cov_out + off[k], d_nout + k, d_nmod + k, d_logdet + k, stream);
            }
            if (rc) return rc;
        }
        return B200LM_OK;
    }
    int* d_n = nullptr; long long* d_off = nullptr; double* d_work = nullptr;
    e = cudaMalloc((void**)&d_n, nblk * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_off, nblk * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, h_n, nblk * sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, off.data(), nblk * sizeof(long long), cudaMemcpyHostToDevice, s);
    const int ldmax = nmax | 1;
    WhitenArgs a;
    a.cov = d_cov; a.w = d_w; a.cov_out = d_cov_out; a.nout = d_nout; a.nmod = d_nmod; a.logdet = d_logdet;
    a.n = d_n; a.off = d_off; a.work = nullptr; a.work_stride = 0;
    a.svdcut = svdcut; a.eps = eps; a.use_eps = use_eps;
    size_t smem;
    if (nmax <= WH_SMEM_NMAX) {
        smem = (2 * (size_t)nmax * ldmax + 3 * (size_t)nmax + 2) * sizeof(double);
    } else {
        a.work_stride = 2LL * nmax * ldmax;
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_work, (size_t)a.work_stride * nblk * sizeof(double));
        a.work = d_work;
        smem = (3 * (size_t)nmax + 2) * sizeof(double);
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(whiten_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        whiten_kernel<<<nblk, WH_THREADS, smem, s>>>(a);
        e = cudaGetLastError();
    }
    // the temporaries are freed after the kernel has run
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_n); cudaFree(d_off); cudaFree(d_work);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "whiten");
    return B200LM_OK;
}
