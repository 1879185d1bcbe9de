// cosf / sinf exactly as glibc 2.39 computes them on x86-64 (sysdeps/ieee754/flt-32/s_cosf.c, s_sinf.c,
// sincosf.h, sincosf_data.c -- the ARM optimized-routines algorithm; third-party code of the reference's
// `(float)cos(angle)` / `(float)sin(angle)` in computeOrbDescriptor, ORBextractor.cc:111-112, where the
// `using namespace std` overload set resolves the float argument to cosf / sinf).
//
// The device cannot call the host's libm, and neither `cosf()` of libdevice nor `(float)cos((double)x)`
// round like glibc for every argument: of the 1 086 931 827 floats in [0, 2*pi] 1 466 624 differ in the
// last bit and 88 of those move a rotated pattern coordinate of rBRIEF (tests/test_glibc_sincosf.py counts
// them).  So this header restates the algorithm operation by operation: float -> double, |x| < pi/4 uses the
// polynomials directly, otherwise one multiply by 2/pi * 2^24, an integer quadrant n, x - n * pi/2, and a
// degree-7 / degree-8 polynomial whose coefficients depend on n & 2.  glibc ships the same C source twice
// for x86-64 (sysdeps/x86_64/fpu/multiarch/s_sinf-fma.c, ifunc-fma.h): compiled with -mfma -mavx2 every
// `a + b * c` of the source is one fused operation, the SSE2 build rounds the product first.  `Fused`
// selects the variant; the engine picks the one the host's libm dispatches to (orb_extract.cu).
// Constants are the published table `__sincosf_table`.  Domain: |x| < 120 (the extractor only passes
// [0, 2*pi]); outside it the functions return NaN so that a misuse cannot go unnoticed.
//
// tests/test_glibc_sincosf.py runs this header on the host against the libm of the box over EVERY float of
// [0, 2*pi] (both functions), and the GPU test compares the device results with the host's.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GSC_HD __host__ __device__ __forceinline__
#else
#define GSC_HD inline
#endif

namespace glibc_sincosf {

// a + b * c, fused (one rounding) or not
template <bool Fused>
GSC_HD double mad(double b, double c, double a) {
#if defined(__CUDA_ARCH__)
  return Fused ? __fma_rn(b, c, a) : __dadd_rn(__dmul_rn(b, c), a);
#else
  if (Fused) return fma(b, c, a);
  volatile double p = b * c;  // keep the product rounded even under -ffp-contract=fast
  return p + a;
#endif
}
GSC_HD double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}

struct Coeffs { double c0, c1, s1, c2, s2, c3, s3, c4; };

// cos polynomial c0 + c1 x^2 + c2 x^4 + c3 x^6 + c4 x^8 in the operation order of sinf_poly (odd n)
template <bool Fused>
GSC_HD float poly_cos(double x2, double sgn) {
  // table 1 (n & 2) is table 0 with the cosine coefficients negated
  const double c0 = sgn * 0x1p0, c1 = sgn * -0x1.ffffffd0c621cp-2, c2 = sgn * 0x1.55553e1068f19p-5,
               c3 = sgn * -0x1.6c087e89a359dp-10, c4 = sgn * 0x1.99343027bf8c3p-16;
  const double x4 = mul(x2, x2);
  const double q1 = mad<Fused>(x2, c1, c0);   // c1' = c0 + x2 * c1
  const double q2 = mad<Fused>(x2, c4, c3);   // c2' = c3 + x2 * c4
  const double x6 = mul(x4, x2);
  const double c = mad<Fused>(x4, c2, q1);    // c = c1' + x4 * c2
  return (float)mad<Fused>(x6, q2, c);        // c + x6 * c2'
}

// sine polynomial x + s1 x^3 + s2 x^5 + s3 x^7 in the operation order of sinf_poly (even n)
template <bool Fused>
GSC_HD float poly_sin(double x, double x2) {
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  const double x3 = mul(x, x2);
  const double q = mad<Fused>(x2, s3, s2);    // s1' = s2 + x2 * s3
  const double x5 = mul(x3, x2);
  const double s = mad<Fused>(x3, s1, x);     // s = x + x3 * s1
  return (float)mad<Fused>(x5, q, s);         // s + x5 * s1'
}

GSC_HD uint32_t abstop12(float x) {
  uint32_t u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(x);
#else
  memcpy(&u, &x, 4);
#endif
  return (u >> 20) & 0x7ff;
}

// reduce_fast: x - n * pi/2 with n = round(x * 2/pi) from a 24-bit fixed-point product
template <bool Fused>
GSC_HD double reduce_fast(double x, int* np) {
  const double r = mul(x, 0x1.45F306DC9C883p+23);
  const int n = ((int32_t)r + 0x800000) >> 24;  // the conversion truncates (cvttsd2si)
  *np = n;
  return mad<Fused>(-(double)n, 0x1.921FB54442D18p0, x);  // x - n * hpi (vfnmadd: -(n*hpi) + x)
}

// sign[n & 3] of the table: {1, -1, -1, 1}
GSC_HD double quadrant_sign(int n) { return ((n ^ (n >> 1)) & 1) ? -1.0 : 1.0; }

GSC_HD float bad() {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(0x7fc00000u);
#else
  return NAN;
#endif
}

template <bool Fused>
GSC_HD float cosf_exact(float y) {
  const double x = (double)y;
  const uint32_t t = abstop12(y);
  if (t < 0x3f4) {  // |y| < pi/4
    if (t < 0x398) return 1.0f;  // |y| < 2^-12
    return poly_cos<Fused>(mul(x, x), 1.0);
  }
  if (t >= 0x42f) return bad();  // |y| >= 120: not needed by the extractor
  int n;
  const double xr = reduce_fast<Fused>(x, &n);
  const double x2 = mul(xr, xr);
  if (n & 1) return poly_sin<Fused>(mul(xr, quadrant_sign(n)), x2);
  return poly_cos<Fused>(x2, (n & 2) ? -1.0 : 1.0);
}

template <bool Fused>
GSC_HD float sinf_exact(float y) {
  const double x = (double)y;
  const uint32_t t = abstop12(y);
  if (t < 0x3f4) {
    if (t < 0x398) return y;
    return poly_sin<Fused>(x, mul(x, x));
  }
  if (t >= 0x42f) return bad();
  int n;
  const double xr = reduce_fast<Fused>(x, &n);
  const double x2 = mul(xr, xr);
  if ((n & 1) == 0) return poly_sin<Fused>(mul(xr, quadrant_sign(n)), x2);
  return poly_cos<Fused>(x2, (n & 2) ? -1.0 : 1.0);
}

}  // namespace glibc_sincosf
