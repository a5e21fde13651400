// ddgi_math.cuh — fp32 vector algebra and pinned transcendental functions shared by every kernel.
//
// GLSL semantics restated for CUDA.  Every operation here is a
// single correctly-rounded IEEE-754 fp32 (or fp64) operation in a fixed order; the
// library is compiled with --fmad=false so nvcc never contracts a*b+c, and explicit
// fused operations are written as __fmaf_rn where a proof shows the result is unchanged.
//
// The header is host/device clean (DDGI_HD) so tests/hostsim can compile the same
// per-ray logic with g++ for CPU-side unit tests; the shipped library only ever
// instantiates it in device code.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define DDGI_HD __host__ __device__ __forceinline__
#define DDGI_D __device__ __forceinline__
#else
#define DDGI_HD inline
#define DDGI_D inline
#endif

#if defined(__GNUC__) || defined(__CUDACC__)
#define DDGI_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define DDGI_UNLIKELY(x) (x)
#endif

namespace ddgi {

struct v3 {
    float x, y, z;
};

DDGI_HD v3 V3(float x, float y, float z)
{
    v3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
DDGI_HD v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
DDGI_HD v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
DDGI_HD v3 operator*(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
DDGI_HD v3 operator*(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
DDGI_HD v3 operator/(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
// dot(a,b) = (ax*bx + ay*by) + az*bz  (glm / SPIR-V OpDot expansion order)
DDGI_HD float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
DDGI_HD float length(v3 a) { return sqrtf(dot(a, a)); }
DDGI_HD v3 cross(v3 a, v3 b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// RN(1/x): the IEEE reciprocal.  rcp.rn returns exactly what 1.0f / x does for every input
// (zeros, denormals, Inf, NaN included) with a shorter sequence than the general division.
DDGI_HD float rcp_exact(float x)
{
#ifdef __CUDA_ARCH__
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
// normalize(v) = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)  (glm 0.9.9.8 form)
DDGI_HD v3 normalize(v3 a)
{
    float inv = rcp_exact(sqrtf(dot(a, a)));
    return a * inv;
}

// GLSL min/max with a NaN operand return the other operand (IEEE minNum/maxNum).
// fminf/fmaxf have exactly that contract on both CUDA (FMNMX) and the host.
DDGI_HD float gmax(float a, float b) { return fmaxf(a, b); }
DDGI_HD float gmin(float a, float b) { return fminf(a, b); }
DDGI_HD float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
DDGI_HD float gfract(float x) { return x - floorf(x); }
DDGI_HD float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
DDGI_HD float gmix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
DDGI_HD float gmod(float x, float y) { return x - y * floorf(x / y); }

// int(float): truncate; NaN -> 0; saturating (the F2I.TRUNC behaviour the reference ran on)
DDGI_HD int f2i(float x)
{
#ifdef __CUDA_ARCH__
    return __float2int_rz(x);
#else
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
#endif
}

// ---------------------------------------------------------------------------------
// Pinned sin/cos: Cody-Waite reduction by pi/2 in four double pieces followed by the
// classic degree-13/14 minimax kernels, evaluated in fp64 without contraction and
// rounded once to fp32.  Deterministic on every IEEE machine; |x| < 1e7 on this path.
// ---------------------------------------------------------------------------------
// The constants are one table: literals on the host; on the device a __constant__ array, so that each
// fp64 operation takes its constant straight from the constant bank instead of building it in a register
// pair first (an fp64 literal costs two moves per use: 36 of the 184 instructions of a scatter in round
// 1's profile).
#define DDGI_SINCOS_CONSTANTS(X)                                                                                      \
    X(two_over_pi, 6.36619772367581382433e-01)                                                                        \
    X(p1, 1.57079632673412561417e+00) X(p2, 6.07710050630396597660e-11) X(p3, 2.02226624871116645580e-21)             \
    X(p3t, 8.47842766036889956997e-32)                                                                                \
    X(S1, -1.66666666666666324348e-01) X(S2, 8.33333333332248946124e-03) X(S3, -1.98412698298579493134e-04)           \
    X(S4, 2.75573137070700676789e-06) X(S5, -2.50507602534068634195e-08) X(S6, 1.58969099521155010221e-10)            \
    X(C1, 4.16666666666666019037e-02) X(C2, -1.38888888888741095749e-03) X(C3, 2.48015872894767294178e-05)            \
    X(C4, -2.75573143513906633035e-07) X(C5, 2.08757232129817482790e-09) X(C6, -1.13596475577881948265e-11)
#ifdef __CUDACC__
#define DDGI_X_VALUE(name, value) value,
static __constant__ double kSinCosTable[] = {DDGI_SINCOS_CONSTANTS(DDGI_X_VALUE)};
#undef DDGI_X_VALUE
#endif
#define DDGI_X_INDEX(name, value) kSinCos_##name,
enum { DDGI_SINCOS_CONSTANTS(DDGI_X_INDEX) kSinCosCount };
#undef DDGI_X_INDEX

DDGI_HD void pin_sincos(float xf, float* s_out, float* c_out)
{
#ifdef __CUDA_ARCH__
#define DDGI_X_LOCAL(name, value) const double name = kSinCosTable[kSinCos_##name];
#else
#define DDGI_X_LOCAL(name, value) const double name = value;
#endif
    DDGI_SINCOS_CONSTANTS(DDGI_X_LOCAL)
#undef DDGI_X_LOCAL
    double x = (double)xf;
    if (!(fabs(x) < 1.0e15)) {
        *s_out = NAN;
        *c_out = NAN;
        return;
    }
    double k = rint(x * two_over_pi);
    double r = x - k * p1;
    r = r - k * p2;
    r = r - k * p3;
    r = r - k * p3t;
    double z = r * r;
    double ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    double sn = r + (z * r) * (S1 + z * ps);
    double pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double cs = 1.0 - (0.5 * z - z * pc);
    long long q = (long long)k;
    double sv, cv;
    switch ((int)(q & 3)) {
        case 0: sv = sn; cv = cs; break;
        case 1: sv = cs; cv = -sn; break;
        case 2: sv = -sn; cv = -cs; break;
        default: sv = -cs; cv = sn; break;
    }
    *s_out = (float)sv;
    *c_out = (float)cv;
}
DDGI_HD float pin_sin(float x)
{
    float s, c;
    pin_sincos(x, &s, &c);
    return s;
}
DDGI_HD float pin_cos(float x)
{
    float s, c;
    pin_sincos(x, &s, &c);
    return c;
}

// Pinned acos: the fdlibm e_acos.c algorithm (rational minimax on [0,0.5], sqrt
// identities elsewhere) in fp64, rounded once to fp32.  Written out so the device
// and the oracle evaluate the identical operation sequence.
DDGI_HD float pin_acos(float xf)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pi = 3.14159265358979311600e+00;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    double x = (double)xf;
    if (x != x) return NAN;
    double ax = fabs(x);
    if (ax >= 1.0) {
        if (ax == 1.0) return x > 0.0 ? 0.0f : (float)(pi + 2.0 * pio2_lo);
        return NAN;
    }
    if (ax < 0.5) {
        if (ax <= 5.55111512312578270212e-17) return (float)(pio2_hi + pio2_lo);
        double z = x * x;
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double r = p / q;
        return (float)(pio2_hi - (x - (pio2_lo - x * r)));
    }
    if (x < 0.0) {
        double z = (1.0 + x) * 0.5;
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double s = sqrt(z);
        double r = p / q;
        double w = r * s - pio2_lo;
        return (float)(pi - 2.0 * (s + w));
    }
    {
        double z = (1.0 - x) * 0.5;
        double s = sqrt(z);
        // df = s with the low 32 bits cleared
        union {
            double d;
            unsigned long long u;
        } cv;
        cv.d = s;
        cv.u &= 0xffffffff00000000ull;
        double df = cv.d;
        double c = (z - df * df) / (s + df);
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double r = p / q;
        double w = r * s + c;
        return (float)(2.0 * (df + w));
    }
}

// RGBA8 UNORM store rule: round(clamp(x,0,1)*255), NaN -> 0
DDGI_HD uint32_t unorm8(float x)
{
    if (x != x) return 0u;
    float c = gclamp(x, 0.0f, 1.0f);
    return (uint32_t)floorf(c * 255.0f + 0.5f);
}
DDGI_HD uint32_t pack_rgba8(float r, float g, float b, float a)
{
    return unorm8(r) | (unorm8(g) << 8) | (unorm8(b) << 16) | (unorm8(a) << 24);
}
// imageLoad of an rgba8 texel: byte / 255.0f.  The IEEE division is replaced by the FMA-corrected
// reciprocal form (q = b*r, q += (b - 255 q) r with r = RN(1/255)), which returns the same float for each
// of the 256 possible bytes - checked exhaustively by tests/test_oracle_math.py; the pixel pass unpacks
// up to 8 x 26 x 3 channels per pixel.
DDGI_HD float unorm8_to_float(uint32_t b)
{
    const float r = 0.0039215688593685627f;  // RN(1 / 255)
    float x = (float)b;
    float q = x * r;
    float e = fmaf(-255.0f, q, x);
    return fmaf(e, r, q);
}
DDGI_HD v3 unpack_rgb8(uint32_t v)
{
    return V3(unorm8_to_float(v & 255u), unorm8_to_float((v >> 8) & 255u), unorm8_to_float((v >> 16) & 255u));
}

}  // namespace ddgi
