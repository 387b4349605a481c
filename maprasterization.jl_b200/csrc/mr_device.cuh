// mr_device.cuh -- exact Float64 per-pixel categorisation on the device.
//
// Arithmetic contract: every operation is one correctly
// rounded binary64 operation, written with the __d*_rn intrinsics so that nvcc can never
// contract a*b+c into an FMA (the file is also compiled with -fmad=false).
//
// Follows /root/reference/src/categories.jl:
//   :76-82   per-pixel categorise (errs, findmin -> first minimum, accept ? best : 0)
//   :91-95   colour error = sum of squared channel differences, left fold
//   :106-121 accept test: known-category override, alpha rule, per-channel box
//   :131-133 v >= lo && v <= hi with lo/hi precomputed by the host
#pragma once
#include "mr_internal.h"

__device__ __forceinline__ double mr_cbrt(double t) {
    // bit-hack seed on the high word + five Newton steps y <- (2y + t/y^2)/3 (Appendix A.4)
    unsigned int hi = (unsigned int)__double2hiint(t);
    hi = hi / 3u + 715094163u;
    double y = __hiloint2double((int)hi, 0);
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        double yy = __dmul_rn(y, y);
        double q = __ddiv_rn(t, yy);
        double s = __dadd_rn(y, y);
        s = __dadd_rn(s, q);
        y = __ddiv_rn(s, 3.0);
    }
    return y;
}

__device__ __forceinline__ double mr_lab_f(double v) {
    const double eps = 216.0 / 24389.0;
    const double kappa = 24389.0 / 27.0;
    if (v > eps) return mr_cbrt(v);
    double t = __dmul_rn(kappa, v);
    t = __dadd_rn(t, 16.0);
    return __ddiv_rn(t, 116.0);
}

__device__ __forceinline__ void mr_rgb8_to_lab(const double* __restrict__ lin, int r8, int g8, int b8,
                                               double& L, double& A, double& B) {
    double r = __ldg(lin + r8), g = __ldg(lin + g8), b = __ldg(lin + b8);
    double X = __dadd_rn(__dadd_rn(__dmul_rn(0.4124564390896921, r), __dmul_rn(0.357576077643909, g)),
                         __dmul_rn(0.18043748326639894, b));
    double Y = __dadd_rn(__dadd_rn(__dmul_rn(0.21267285140562248, r), __dmul_rn(0.715152155287818, g)),
                         __dmul_rn(0.07217499330655958, b));
    double Z = __dadd_rn(__dadd_rn(__dmul_rn(0.019333895582329317, r), __dmul_rn(0.119192025881303, g)),
                         __dmul_rn(0.9503040785363677, b));
    double fx = mr_lab_f(__ddiv_rn(X, 0.95047));
    double fy = mr_lab_f(Y);
    double fz = mr_lab_f(__ddiv_rn(Z, 1.08883));
    L = __dsub_rn(__dmul_rn(116.0, fy), 16.0);
    A = __dmul_rn(500.0, __dsub_rn(fx, fy));
    B = __dmul_rn(200.0, __dsub_rn(fy, fz));
}

// categorise one colour given as Float64 channels (x3 = alpha, used when pal.C == 4).
// known: 0 = ordinary pixel; k != 0 = clicked point of category k (categories.jl:110-112);
// transparent: the alpha == 0 rule of categories.jl:116 applies to this pixel.
__device__ __forceinline__ int mr_eval_f64(const MrPalette& pal, double x0, double x1, double x2, double x3,
                                           bool transparent, int known) {
    const int C = pal.C;
    int best = 0;
    double beste = 0.0;
    for (int c = 0; c < pal.K; ++c) {
        const double* m = pal.mean + (size_t)c * C;
        double d0 = __dsub_rn(x0, __ldg(m + 0));
        double d1 = __dsub_rn(x1, __ldg(m + 1));
        double d2 = __dsub_rn(x2, __ldg(m + 2));
        double e = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
        if (C == 4) {
            double d3 = __dsub_rn(x3, __ldg(m + 3));
            e = __dadd_rn(e, __dmul_rn(d3, d3));
        }
        if (c == 0 || e < beste) { beste = e; best = c; }   // first minimum
    }
    if (known != 0) return (best + 1 == known) ? best + 1 : 0;
    if (transparent) return 0;                               // categories.jl:116
    const double* l = pal.lo + (size_t)best * C;
    const double* h = pal.hi + (size_t)best * C;
    bool ok = (x0 >= __ldg(l + 0)) && (x0 <= __ldg(h + 0)) && (x1 >= __ldg(l + 1)) &&
              (x1 <= __ldg(h + 1)) && (x2 >= __ldg(l + 2)) && (x2 <= __ldg(h + 2));
    return ok ? best + 1 : 0;
}

// the same for a pixel given as N0f8 bytes
__device__ __forceinline__ int mr_eval_px(const MrPalette& pal, int r8, int g8, int b8, int a8,
                                          bool has_alpha, int known) {
    double x0, x1, x2, x3 = 0.0;
    if (pal.colourspace == 1) {
        mr_rgb8_to_lab(pal.lin, r8, g8, b8, x0, x1, x2);
    } else {
        x0 = __ddiv_rn((double)r8, 255.0);   // Float64(::N0f8) = correctly rounded i/255
        x1 = __ddiv_rn((double)g8, 255.0);
        x2 = __ddiv_rn((double)b8, 255.0);
        if (pal.C == 4) x3 = __ddiv_rn((double)a8, 255.0);
    }
    return mr_eval_f64(pal, x0, x1, x2, x3, has_alpha && a8 == 0, known);
}
