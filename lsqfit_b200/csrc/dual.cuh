// Forward-mode dual numbers held in registers (value + N partial derivatives).
//
// This is the device counterpart of gvar's automatic differentiation, which the
// reference uses to obtain Jacobians (reference src/lsqfit/_scipy.py:144-154:
// f(valder + x) and fx[i].der).  All loops are over a compile-time N and are
// fully unrolled, so a Dual<N> lives in 2(N+1) 32-bit registers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace b200lm {

template <int N>
struct Dual {
    double v;
    double d[N];

    __device__ __forceinline__ Dual() {}
    __device__ __forceinline__ Dual(double c) : v(c) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = 0.0;
    }
    // independent variable number k
    __device__ __forceinline__ static Dual variable(double val, int k) {
        Dual r;
        r.v = val;
#pragma unroll
        for (int i = 0; i < N; ++i) r.d[i] = (i == k) ? 1.0 : 0.0;
        return r;
    }
};

// chain rule helper: f(a) with derivative fp
template <int N>
__device__ __forceinline__ Dual<N> chain(const Dual<N>& a, double f, double fp) {
    Dual<N> r;
    r.v = f;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = fp * a.d[i];
    return r;
}

template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, double b) {
    Dual<N> r = a; r.v += b; return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator+(double b, const Dual<N>& a) { return a + b; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, double b) {
    Dual<N> r = a; r.v -= b; return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(double b, const Dual<N>& a) {
    Dual<N> r; r.v = b - a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = fma(a.d[i], b.v, a.v * b.d[i]);
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, double b) {
    Dual<N> r; r.v = a.v * b;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
// NOTE: every .v below is computed exactly like the plain-double expression would be, so that
// Body::eval<double> and Body::eval<Dual<N>>().v agree bit for bit.
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    const double ib = 1.0 / b.v;
    Dual<N> r; r.v = a.v / b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, double b) {
    const double ib = 1.0 / b;
    Dual<N> r; r.v = a.v / b;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * ib;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator/(double a, const Dual<N>& b) {
    const double ib = 1.0 / b.v;
    const double v = a / b.v;
    return chain(b, v, -v * ib);
}

template <int N> __device__ __forceinline__ Dual<N> exp(const Dual<N>& a) {
    const double e = ::exp(a.v); return chain(a, e, e);
}
template <int N> __device__ __forceinline__ Dual<N> log(const Dual<N>& a) {
    return chain(a, ::log(a.v), 1.0 / a.v);
}
template <int N> __device__ __forceinline__ Dual<N> sqrt(const Dual<N>& a) {
    const double s = ::sqrt(a.v); return chain(a, s, 0.5 / s);
}
template <int N> __device__ __forceinline__ Dual<N> sin(const Dual<N>& a) {
    return chain(a, ::sin(a.v), ::cos(a.v));
}
template <int N> __device__ __forceinline__ Dual<N> cos(const Dual<N>& a) {
    return chain(a, ::cos(a.v), -::sin(a.v));
}
template <int N> __device__ __forceinline__ Dual<N> atan(const Dual<N>& a) {
    return chain(a, ::atan(a.v), 1.0 / (1.0 + a.v * a.v));
}
// a ** c, constant exponent
template <int N> __device__ __forceinline__ Dual<N> pow(const Dual<N>& a, double c) {
    const double v = ::pow(a.v, c);
    return chain(a, v, a.v != 0.0 ? c * v / a.v : c * ::pow(a.v, c - 1.0));
}
// a ** b, both dual (a > 0)
template <int N> __device__ __forceinline__ Dual<N> pow(const Dual<N>& a, const Dual<N>& b) {
    const double v = ::pow(a.v, b.v);
    const double da = b.v * v / a.v, db = v * ::log(a.v);
    Dual<N> r; r.v = v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = fma(da, a.d[i], db * b.d[i]);
    return r;
}
// c ** b, constant base c > 0
template <int N> __device__ __forceinline__ Dual<N> pow(double c, const Dual<N>& b) {
    const double v = ::pow(c, b.v);
    return chain(b, v, v * ::log(c));
}
template <int N> __device__ __forceinline__ Dual<N> sqr(const Dual<N>& a) { return a * a; }

// plain-double overloads so that functor bodies are generic in T (the templates
// above hide the global math functions inside this namespace)
__device__ __forceinline__ double sqr(double a) { return a * a; }
__device__ __forceinline__ double exp(double a) { return ::exp(a); }
__device__ __forceinline__ double log(double a) { return ::log(a); }
__device__ __forceinline__ double sqrt(double a) { return ::sqrt(a); }
__device__ __forceinline__ double sin(double a) { return ::sin(a); }
__device__ __forceinline__ double cos(double a) { return ::cos(a); }
__device__ __forceinline__ double atan(double a) { return ::atan(a); }
__device__ __forceinline__ double pow(double a, double b) { return ::pow(a, b); }

}  // namespace b200lm
