// Bootstrap / simulated copies of the mean vector on the device:  mean_b = mean + L z_b,  z_b ~ N(0, 1)
// (what gvar.bootstrap_iter / gvar.raniter produce for the reference at src/lsqfit/__init__.py:
// 1532-1535, 1615-1624, one copy per Python iteration).  Counter-based Philox4x32-10 (Salmon et al.,
// SC'11) keyed by the seed, counter = index of the normal PAIR in the flat [copy][component] stream,
// so any sub-range of copies can be generated independently (multi-GPU shards, chunked batches) and
// is bit-identical to the same rows of one big call.  Normals by Box-Muller on 53-bit uniforms.
// The product Z L^T runs on the DMMA GEMM.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/b200lm.h"
#include "handle.h"
#include "dgemm.cuh"

namespace b200lm {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// element g of the stream = normal (g & 1) of pair (g >> 1)
__global__ void normals_kernel(long long g0, long long count, unsigned long long seed, double* __restrict__ z,
                               uint32_t* __restrict__ raw) {
    const long long p0 = g0 >> 1, p1 = (g0 + count - 1) >> 1;
    const long long pair = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pair > p1) return;
    uint32_t c[4] = {(uint32_t)pair, (uint32_t)((unsigned long long)pair >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    if (raw) {                                  // known-answer hook: the four raw words of this counter
        uint32_t* o = raw + 4 * (pair - p0);
        o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = c[3];
    }
    // two uniforms in (0, 1) with 53 random bits each
    const double u1 = ((double)(((unsigned long long)(c[0] >> 5) << 26) | (c[1] >> 6)) + 0.5) * 0x1p-53;
    const double u2 = ((double)(((unsigned long long)(c[2] >> 5) << 26) | (c[3] >> 6)) + 0.5) * 0x1p-53;
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    const long long e0 = 2 * pair - g0, e1 = e0 + 1;
    if (e0 >= 0 && e0 < count) z[e0] = r * cs;
    if (e1 >= 0 && e1 < count) z[e1] = r * sn;
}

__global__ void broadcast_rows_kernel(long long B, int N, const double* __restrict__ mean, double* __restrict__ out,
                                      long long stride) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * N) return;
    const long long b = e / N;
    const int i = (int)(e % N);
    out[b * stride + i] = mean[i];
}

cudaError_t normals(long long g0, long long count, unsigned long long seed, double* d_z, uint32_t* d_raw, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const long long npair = ((g0 + count - 1) >> 1) - (g0 >> 1) + 1;
    normals_kernel<<<(unsigned)((npair + 255) / 256), 256, 0, s>>>(g0, count, seed, d_z, d_raw);
    return cudaGetLastError();
}

}  // namespace b200lm

using namespace b200lm;

extern "C" int b200lm_normals(int device, long long first, long long count, unsigned long long seed, double* d_z,
                              unsigned int* d_raw, void* stream) {
    if (first < 0 || count < 0 || !d_z) return set_error(nullptr, B200LM_EINVAL, "bad normals argument");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    e = normals(first, count, seed, d_z, d_raw, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "normals");
    return B200LM_OK;
}

extern "C" int b200lm_bootstrap_means(int device, long long B, long long first, int N, int M, const double* d_mean,
                                      const double* d_L, int ldl, unsigned long long seed, double* d_z,
                                      double* d_out, long long out_stride, void* stream) {
    if (B < 0 || first < 0 || N <= 0 || M <= 0 || !d_mean || !d_L || !d_z || !d_out || ldl < M || out_stride < N ||
        B > 0x7fffffffLL)
        return set_error(nullptr, B200LM_EINVAL, "bad bootstrap_means argument");
    if (B == 0) return B200LM_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    cudaStream_t s = (cudaStream_t)stream;
    e = normals(first * M, B * M, seed, d_z, nullptr, s);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "normals");
    broadcast_rows_kernel<<<(unsigned)((B * N + 255) / 256), 256, 0, s>>>(B, N, d_mean, d_out, out_stride);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "broadcast_rows");
    // out[b][i] += sum_j Z[b][j] L[i][j]
    e = dgemm(false, true, 1, (int)B, N, M, 1.0, d_z, 0, M, d_L, 0, ldl, 1.0, d_out, 0, (int)out_stride, s);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "bootstrap gemm");
    return B200LM_OK;
}
