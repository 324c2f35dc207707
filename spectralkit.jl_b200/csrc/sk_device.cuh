// Device-side numerics shared by all kernels: coordinate transformations and jet arithmetic.
//
// This translation unit is compiled with -fmad=false: nvcc never contracts a*b+c on its own, so
// plain `*`, `+`, `-` reproduce the reference's (Julia's) unfused IEEE operations bit for bit, and
// every fused multiply-add in the fast kernels is an explicit fma().  Division and sqrt are the
// IEEE-correct double-precision versions (nvcc default for FP64).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "spectralkit_b200.h"

namespace skd {

// Julia Base.Math._hypot, native-fma branch (call site: src/transformations.jl:314).  Same
// algorithm as the reference's dependency; documented there as correctly rounded.
__device__ __forceinline__ double hypot_julia(double x, double y) {
    double ax = fabs(x), ay = fabs(y);
    if (isinf(ax) || isinf(ay)) return CUDART_INF;
    if (ay > ax) { double t = ax; ax = ay; ay = t; }
    if (isnan(ax) || isnan(ay)) return CUDART_NAN;
    const double eps = 2.220446049250313e-16;
    if (ay <= ax * 1.0536712127723509e-08 /* sqrt(eps/2) */) return ax;
    double scale = eps * 1.4916681462400413e-154 /* sqrt(floatmin) */;
    if (ax > 9.480751908109176e+153 /* sqrt(floatmax/2) */) { ax = ax * scale; ay = ay * scale; scale = 1.0 / scale; }
    else if (ay < 1.4916681462400413e-154) { ax = ax / scale; ay = ay / scale; }
    else scale = 1.0;
    double h = sqrt(fma(ax, ax, ay * ay));
    double hsq = h * h, axsq = ax * ax;
    h -= (fma(-ay, ay, hsq - axsq) + fma(h, h, -hsq) - fma(ax, ax, -axsq)) / (2 * h);
    return h * scale;
}

// transform_to on a real (src/transformations.jl:183-186, 246-255, 311-320)
__device__ __forceinline__ double to_pm1(const sk_transform &t, double y) {
    switch (t.kind) {
    case SK_TR_BOUNDED_LINEAR: return (y - t.p0) / t.p1;
    case SK_TR_SEMI_INF_RATIONAL: {
        double z = y - t.p0;
        double x = (z - t.p1) / (z + t.p1);
        return isinf(y) ? 1.0 : x;                 // +1 at both infinities, as the reference
    }
    case SK_TR_INF_RATIONAL: {
        double z = y - t.p0;
        double x = z / hypot_julia(z, t.p1);
        return isinf(y) ? (y > 0 ? 1.0 : -1.0) : x;
    }
    default: return y;
    }
}

// transform_to of the seeded jet (y, 1, 0, …) of order DP1-1 (src/transformations.jl:188-195,
// 257-267, 322-333).  Order 2 through the rational maps is the documented extension
// (x2 = x''(y)); the plan rejects it unless the caller opted in.
template <int DP1>
__device__ __forceinline__ void to_pm1_jet(const sk_transform &t, double y, double (&x)[DP1]) {
    x[0] = to_pm1(t, y);
    if (DP1 == 1) return;
#pragma unroll
    for (int k = 1; k < DP1; k++) x[k] = 0.0;
    switch (t.kind) {
    case SK_TR_BOUNDED_LINEAR:
        x[1] = 1.0 / t.p1;
#pragma unroll
        for (int k = 2; k < DP1; k++) x[k] = 0.0 / t.p1;
        break;
    case SK_TR_SEMI_INF_RATIONAL: {
        double L = t.p1, x0 = x[0];
        double Q = (x0 - 1) * (x0 - 1);
        x[1] = (1.0 * Q) / (2 * L);
        if (DP1 > 2) {
            double omx = 1 - x0;
            double xpp = -((omx * omx) * omx) / ((2 * L) * L);
            x[2] = xpp * (1.0 * 1.0) + (Q / (2 * L)) * 0.0;
        }
        break;
    }
    case SK_TR_INF_RATIONAL: {
        double L = t.p1, x0 = x[0];
        double Q = 1 - x0 * x0;
        double sQ = sqrt(Q);
        x[1] = ((1.0 * Q) * sQ) / L;
        if (DP1 > 2) {
            double xpp = -(((3 * x0) * Q) * Q) / (L * L);
            x[2] = xpp * (1.0 * 1.0) + ((Q * sQ) / L) * 0.0;
        }
        break;
    }
    default:
        x[1] = 1.0;
        break;
    }
}

// transform_from (src/transformations.jl:178-181, 244, 309)
__device__ __forceinline__ double from_pm1(const sk_transform &t, double x) {
    switch (t.kind) {
    case SK_TR_BOUNDED_LINEAR: return x * t.p1 + t.p0;
    case SK_TR_SEMI_INF_RATIONAL: return t.p0 + (t.p1 * (1 + x)) / (1 - x);
    case SK_TR_INF_RATIONAL: return t.p0 + (t.p1 * x) / sqrt(1 - x * x);
    default: return x;
    }
}

// ---- jets of raw derivatives, reference operation order (src/derivatives.jl:113-134)
__host__ __device__ constexpr int binom(int n, int k) {
    return (k < 0 || k > n) ? 0 : (k == 0 || k == n) ? 1 : binom(n - 1, k - 1) + binom(n - 1, k);
}

template <int DP1>
struct Jet {
    double c[DP1];
};

template <int DP1>
__device__ __forceinline__ Jet<DP1> jet_one() {
    Jet<DP1> r;
#pragma unroll
    for (int k = 0; k < DP1; k++) r.c[k] = k == 0 ? 1.0 : 0.0;
    return r;
}
// one Chebyshev step in the reference's order: f = (2x)·fp − fpp with jet×jet Leibniz products
// out_k = Σ_i ((binom(k,i)·t_i)·fp_{k−i}), summed left to right (src/chebyshev.jl:65-70)
template <int DP1>
__device__ __forceinline__ Jet<DP1> cheb_step_exact(const Jet<DP1> &x, const Jet<DP1> &fp, const Jet<DP1> &fpp) {
    Jet<DP1> t, f;
#pragma unroll
    for (int k = 0; k < DP1; k++) t.c[k] = 2.0 * x.c[k];
#pragma unroll
    for (int k = 0; k < DP1; k++) {
        double s = ((double)binom(k, 0) * t.c[0]) * fp.c[k];
#pragma unroll
        for (int i = 1; i <= k; i++) s = s + ((double)binom(k, i) * t.c[i]) * fp.c[k - i];
        f.c[k] = s - fpp.c[k];
    }
    return f;
}

}  // namespace skd
