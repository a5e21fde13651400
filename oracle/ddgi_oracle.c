/*
 * ddgi_oracle.c — CPU restatement of the reference's DDGI probe-field hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the parity checker for the CUDA
 * engine and the reported CPU baseline.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (dynamic-diffuse-global-illumination-minecraft_b200/) never links, imports or
 * calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference (helenl9098/Dynamic-Diffuse-Global-Illumination-
 * Minecraft @ 6e02028) ships no tests, golden vectors or fixtures for this path,
 * and its GLSL cannot be compiled or run here (no glslang / Vulkan / lavapipe;
 * its host code needs glm + network FetchContent).  This restatement follows the
 * GLSL / C++ line by line (each function cites the lines it follows) and freezes
 * every piece of undefined or platform-dependent behaviour as listed under
 * "PINS" below.  Golden vectors under tests/golden/ are outputs of THIS file.
 *
 * All paths are relative to /root/reference:
 *   G  = assets/shaders/intersection.glsl     P = assets/shaders/probe_pass.comp
 *   I  = assets/shaders/integrators.glsl      C = assets/shaders/compute_pass.comp
 *   K  = assets/shaders/camera.glsl           S = assets/shaders/structs.glsl
 *   R  = src/rvpt/rvpt.cpp
 *
 * PINS:
 *  1. int(NaN) -> 0; int(x) saturates at INT_MIN/INT_MAX (NVIDIA F2I).
 *  2. Uninitialised Isect.type / Material -> zero.
 *  3. GLSL min/max with a NaN operand -> IEEE minNum/maxNum (non-NaN operand).
 *  4. getColorAt fall-through for an unknown type -> (0,0,0,0).
 *  5. libc rand(): glibc, seed 1, the two calls of R:1161-1162 evaluated left to
 *     right (x jitter first).
 *  6. Dispatch overshoot (gid.x >= W or gid.y >= H) is not executed.
 *  7. RGBA8 store: floor(clamp(x,0,1)*255 + 0.5), NaN -> 0.
 *  8. fp32 everywhere, no FMA contraction (build with -ffp-contract=off);
 *     normalize(v) = v * (1.0f / sqrtf(dot(v,v))) (glm's form);
 *     dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z;
 *     sin/cos = pin_sin/pin_cos below: a double-precision Cody-Waite +
 *     fdlibm-polynomial evaluation rounded once to fp32 (deterministic on any
 *     IEEE-754 machine, within 1 fp32 ulp of libm; tests/test_oracle_math.py);
 *     acos = pin_acos (fdlibm algorithm in fp64, rounded once to fp32);
 *     tan (host-side camera uniform only) = (float) of the double libm function.
 *  9. The pixel pass writes only gid < (floor(w/16)*16, floor(h/16)*16).
 * 10. The pixel pass sees the fully updated probe texture of the same frame.
 *
 * Extensions beyond the reference (each reduces exactly to the reference when
 * not used): stored-voxel scene (scene_mode 1: uint8 block types + palette,
 * out-of-grid = empty), rectangular rx x ry ray tiles, caller-supplied lights.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* vector helpers (GLSL semantics)                                            */
/* ------------------------------------------------------------------------- */

typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 vdivs(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 vneg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float vlen(v3 a) { return sqrtf(vdot(a, a)); }
static inline v3 vcross(v3 a, v3 b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* PIN 8: glm/GLSL normalize = v * inversesqrt(dot(v,v)) */
static inline v3 vnormalize(v3 a)
{
    float inv = 1.0f / sqrtf(vdot(a, a));
    return vscale(a, inv);
}
static inline float vget(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* PIN 3 */
static inline float gmax(float a, float b)
{
    if (b != b) return a;
    if (a != a) return b;
    return a < b ? b : a;
}
static inline float gmin(float a, float b)
{
    if (b != b) return a;
    if (a != a) return b;
    return b < a ? b : a;
}
static inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
static inline float gfract(float x) { return x - floorf(x); }
static inline float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline float gmix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float gmod(float x, float y) { return x - y * floorf(x / y); }

/* PIN 1 */
static inline int f2i(float x)
{
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
}

/* ------------------------------------------------------------------------- */
/* pinned transcendental functions (PIN 8)                                    */
/* ------------------------------------------------------------------------- */

/* fdlibm constants: pi/2 split in 33-bit pieces, kernel sin/cos polynomials */
static const double PIN_2OPI = 6.36619772367581382433e-01;
static const double PIN_PIO2_1 = 1.57079632673412561417e+00;
static const double PIN_PIO2_2 = 6.07710050630396597660e-11;
static const double PIN_PIO2_3 = 2.02226624871116645580e-21;
static const double PIN_PIO2_3T = 8.47842766036889956997e-32;
static const double PIN_S1 = -1.66666666666666324348e-01, PIN_S2 = 8.33333333332248946124e-03,
                    PIN_S3 = -1.98412698298579493134e-04, PIN_S4 = 2.75573137070700676789e-06,
                    PIN_S5 = -2.50507602534068634195e-08, PIN_S6 = 1.58969099521155010221e-10;
static const double PIN_C1 = 4.16666666666666019037e-02, PIN_C2 = -1.38888888888741095749e-03,
                    PIN_C3 = 2.48015872894767294178e-05, PIN_C4 = -2.75573143513906633035e-07,
                    PIN_C5 = 2.08757232129817482790e-09, PIN_C6 = -1.13596475577881948265e-11;

/* sin and cos of a float argument, computed in double, rounded once to float.
   Valid (accurate to < 1e-9 absolute) for |x| < 1e7; deterministic everywhere. */
static void pin_sincos(float xf, float* s_out, float* c_out)
{
    double x = (double)xf;
    if (!(fabs(x) < 1.0e15)) { /* NaN, Inf or |x| >= 1e15 (never reached on this path): NaN */
        *s_out = NAN;
        *c_out = NAN;
        return;
    }
    double k = rint(x * PIN_2OPI);
    double r = x - k * PIN_PIO2_1;
    r = r - k * PIN_PIO2_2;
    r = r - k * PIN_PIO2_3;
    r = r - k * PIN_PIO2_3T;
    double z = r * r;
    double ps = PIN_S2 + z * (PIN_S3 + z * (PIN_S4 + z * (PIN_S5 + z * PIN_S6)));
    double sn = r + (z * r) * (PIN_S1 + z * ps);
    double pc = z * (PIN_C1 + z * (PIN_C2 + z * (PIN_C3 + z * (PIN_C4 + z * (PIN_C5 + z * PIN_C6)))));
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
static inline float pin_sin(float x) { float s, c; pin_sincos(x, &s, &c); return s; }
static inline float pin_cos(float x) { float s, c; pin_sincos(x, &s, &c); return c; }
/* acos: the fdlibm e_acos.c algorithm evaluated in fp64 and rounded once to fp32
   (deterministic; within 1 fp32 ulp of libm, tests/test_oracle_math.py). */
static double pin_acos_pq(double z)
{
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return p / q;
}
static float pin_acos(float xf)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pi = 3.14159265358979311600e+00;
    double x = (double)xf;
    if (x != x) return NAN;
    double ax = fabs(x);
    if (ax >= 1.0) {
        if (ax == 1.0) return x > 0.0 ? 0.0f : (float)(pi + 2.0 * pio2_lo);
        return NAN;
    }
    if (ax < 0.5) {
        if (ax <= 5.55111512312578270212e-17) return (float)(pio2_hi + pio2_lo);
        double r = pin_acos_pq(x * x);
        return (float)(pio2_hi - (x - (pio2_lo - x * r)));
    }
    if (x < 0.0) {
        double z = (1.0 + x) * 0.5;
        double r = pin_acos_pq(z);
        double s = sqrt(z);
        double w = r * s - pio2_lo;
        return (float)(pi - 2.0 * (s + w));
    }
    double z = (1.0 - x) * 0.5;
    double s = sqrt(z);
    union { double d; uint64_t u; } cv;
    cv.d = s;
    cv.u &= 0xffffffff00000000ull;
    double df = cv.d;
    double c = (z - df * df) / (s + df);
    double r = pin_acos_pq(z);
    double w = r * s + c;
    return (float)(2.0 * (df + w));
}
ORC_API void orc_pin_acos(const float* x, int n, float* out)
{
    for (int i = 0; i < n; i++) out[i] = pin_acos(x[i]);
}

ORC_API void orc_pin_sincos(const float* x, int n, float* s, float* c)
{
    for (int i = 0; i < n; i++) pin_sincos(x[i], &s[i], &c[i]);
}

/* ------------------------------------------------------------------------- */
/* public structs (ctypes mirrors these in oracle/oracle.py)                  */
/* ------------------------------------------------------------------------- */

/* S:54-59 */
typedef struct { float intensity; float col[3]; float pos[3]; } OrcLight;

/* S:22-27 / src/rvpt/probe.h:5-20 — 3 x alignas(16) vec3 = 48 bytes */
typedef struct {
    float origin[3]; float _p0;
    float direction[3]; float _p1;
    float probe_info[3]; float _p2;
} OrcProbeRay;

typedef struct {
    /* scene */
    int32_t scene_mode;  /* 0: procedural getBlockAt(scene) G:699-826; 1: stored voxel grid */
    int32_t scene;       /* 0 cave, 1 Cornell, 2 house */
    int32_t color_mode;  /* 0: literal getColorAt G:872-1047; 1: palette[type] */
    int32_t n_lights;
    OrcLight lights[8];
    int32_t vdim[3];     /* voxel grid size (x,y,z) */
    int32_t vorg[3];     /* voxel id of grid cell (0,0,0); voxel c covers (c-1,c] per axis */
    const uint8_t* vox;  /* linear, x fastest: vox[(z*dimy + y)*dimx + x] */
    const float* palette; /* 256 x 3 floats */
    /* irradiance field (src/rvpt/rvpt.h:82-90) */
    int32_t probe_count[3];
    int32_t side_length;
    int32_t rx, ry;      /* ray-tile shape; the reference has rx = ry = sqrt_rays_per_probe */
    float field_origin[3];
    /* render settings (src/rvpt/rvpt.h:70-80) */
    int32_t max_bounces;
    int32_t screen_width, screen_height;
} OrcParams;

/* ------------------------------------------------------------------------- */
/* light tables (S:61-89) and get_light (G:1131-1150)                         */
/* ------------------------------------------------------------------------- */

ORC_API int orc_default_lights(int scene, OrcLight* out)
{
    /* S:61 num_lights = {1,1,2}; S:63, S:81, S:88-89 */
    static const OrcLight l0[1] = {{100.f, {1.f, 1.f, 1.f}, {4.f, 17.5f, 8.5f}}};
    static const OrcLight l1[1] = {{15.f, {1.f, 1.f, 1.f}, {0.f, 8.f, 13.f}}};
    static const OrcLight l2[2] = {{1.f, {1.f, 1.f, 1.f}, {5.f, 9.3f, 36.5f}},
                                   {1.f, {1.f, 1.f, 1.f}, {0.f, 0.f, 0.f}}};
    if (scene == 0) { out[0] = l0[0]; return 1; }
    if (scene == 1) { out[0] = l1[0]; return 1; }
    if (scene == 2) { out[0] = l2[0]; out[1] = l2[1]; return 2; }
    return 0;
}

/* The commented 4-light cave table S:65-69 (BASELINE config 4). */
ORC_API int orc_cave_lights4(OrcLight* out)
{
    static const OrcLight t[4] = {{20.f, {1.f, 1.f, 1.f}, {4.f, 17.5f, 8.5f}},
                                  {10.f, {1.f, 0.5f, 0.1f}, {0.f, 2.f, 0.f}},
                                  {10.f, {0.1f, 1.1f, 1.f}, {5.f, 0.f, 0.f}},
                                  {10.f, {1.1f, 0.f, 1.1f}, {0.f, 5.f, 0.f}}};
    for (int i = 0; i < 4; i++) out[i] = t[i];
    return 4;
}

/* update_lights, cave branch (P:219-235): positions as a function of time,
   applied to the *initial* table (GLSL globals are re-initialised per invocation).
   sin/cos are the pinned versions. */
ORC_API void orc_update_lights_cave(const OrcLight* base, int n, float time, OrcLight* out)
{
    for (int i = 0; i < n; i++) {
        out[i] = base[i];
        float t = 0.05f * time;
        if (i == 0) {
            out[i].pos[2] = base[i].pos[2] + 10.0f * pin_cos(t * 0.1f);
            continue;
        }
        float x = base[i].pos[0] + (float)((i + 1) * 2) * pin_sin(t * 0.5f);
        float y = base[i].pos[1] + (float)((i / 2) * 4) * pin_sin(t * 0.5f);
        float z = base[i].pos[2] + (float)((i + 1) * 2) * pin_cos(t * 0.5f);
        out[i].pos[0] = x; out[i].pos[1] = y; out[i].pos[2] = z;
    }
}

/* ------------------------------------------------------------------------- */
/* noise (G:400-499)                                                          */
/* ------------------------------------------------------------------------- */

static float random1(v3 p) /* G:400 */
{
    float d = (p.x * 127.1f + p.y * 311.7f) + p.z * 191.999f;
    return gfract(pin_sin(d) * 43758.5453f);
}
static float noise2D(float px, float py) /* G:402 */
{
    float d = px * 127.1f + py * 311.7f;
    return gfract(pin_sin(d) * 43758.5453f);
}
static float interpNoise2D(float x, float y) /* G:404-419 */
{
    int intX = f2i(floorf(x));
    float fractX = gfract(x);
    int intY = f2i(floorf(y));
    float fractY = gfract(y);
    float v1 = noise2D((float)intX, (float)intY);
    float v2 = noise2D((float)(intX + 1), (float)intY);
    float v3_ = noise2D((float)intX, (float)(intY + 1));
    float v4 = noise2D((float)(intX + 1), (float)(intY + 1));
    float i1 = gmix(v1, v2, fractX);
    float i2 = gmix(v3_, v4, fractX);
    return gmix(i1, i2, fractY);
}
static float fbm(float x, float y) /* G:421-435 */
{
    float total = 0.0f;
    for (int i = 1; i <= 8; i++) {
        float freq = ldexpf(1.0f, i);  /* pow(2.f, i), exact */
        float amp = ldexpf(1.0f, -i);  /* pow(0.5, i), exact */
        total += interpNoise2D(x * freq, y * freq) * amp;
    }
    return total;
}
static float noise1(float i) /* G:437-439: fract(sin(vec2(203.311*i, ...))).x */
{
    return gfract(pin_sin(203.311f * i));
}
static float interpNoise1D(float x) /* G:441-448 */
{
    float intX = floorf(x);
    float fractX = gfract(x);
    float v1 = noise1(intX);
    float v2 = noise1(intX + 1.0f);
    return gmix(v1, v2, fractX);
}
static float fbm1D(float x) /* G:450-463 */
{
    float total = 0.0f;
    for (int i = 0; i < 8; i++) {
        float freq = ldexpf(1.0f, i);
        float amp = ldexpf(1.0f, -i);
        total += interpNoise1D(x * freq) * amp;
    }
    return total;
}
#define CELL_SIZE 5.0f
static v2 generate_point(v2 cell) /* G:467-471 */
{
    /* fract(sin(vec2(dot(p,a), dot(p,b) * 43758.5453))) — note the parenthesisation */
    float a = cell.x * 127.1f + cell.y * 311.7f;
    float b = (cell.x * 269.5f + cell.y * 183.3f) * 43758.5453f;
    v2 p;
    p.x = (cell.x + gfract(pin_sin(a))) * CELL_SIZE;
    p.y = (cell.y + gfract(pin_sin(b))) * CELL_SIZE;
    return p;
}
static float len2(float x, float y) { return sqrtf(x * x + y * y); }
static float worleyNoise(v2 pixel) /* G:473-499 */
{
    v2 cell;
    cell.x = floorf(pixel.x / CELL_SIZE);
    cell.y = floorf(pixel.y / CELL_SIZE);
    v2 point = generate_point(cell);
    float shortest = len2(pixel.x - point.x, pixel.y - point.y);
    for (float i = -1.0f; i <= 1.0f; i += 1.0f) {
        float ncx = cell.x + i;
        for (float j = -1.0f; j <= 1.0f; j += 1.0f) {
            float ncy = cell.y + j;
            v2 nc = {ncx, ncy};
            v2 np = generate_point(nc);
            float d = len2(pixel.x - np.x, pixel.y - np.y);
            if (d < shortest) shortest = d;
        }
    }
    return shortest / CELL_SIZE;
}

/* ------------------------------------------------------------------------- */
/* procedural scene: getBlockAt (G:538-826)                                   */
/* ------------------------------------------------------------------------- */

static float sdSphere(v3 p, float s) { return vlen(p) - s; } /* G:321 */

static float sdRoundBox(v3 p, v3 b, float r) /* G:538-542 */
{
    v3 q = V3(fabsf(p.x) - b.x, fabsf(p.y) - b.y, fabsf(p.z) - b.z);
    v3 qm = V3(gmax(q.x, 0.0f), gmax(q.y, 0.0f), gmax(q.z, 0.0f));
    return vlen(qm) + gmin(gmax(q.x, gmax(q.y, q.z)), 0.0f) - r;
}
static int tiny_mushroom(v3 p) /* G:544-552 */
{
    if (sdRoundBox(p, V3(1.0f, 0.5f, 1.0f), 0.0f) <= 0) return 7;
    if (p.x == 0 && p.z == 0 && p.y < 0) return 9;
    return 0;
}
static int small_mushroom(v3 p) /* G:554-570 */
{
    if (sdRoundBox(p, V3(1.0f, 0.5f, 1.0f), 1.0f) <= 0) {
        if (p.y > 0) return 8;
        if (p.y == 0) return 7;
        if (p.y < 0) return 6;
    }
    if (p.x == 0 && p.z == 0 && p.y < 0) return 9;
    return 0;
}
static int medium_mushroom(v3 p) /* G:572-594 */
{
    if (sdRoundBox(p, V3(2.0f, 0.5f, 2.0f), 1.0f) <= 0) {
        if (p.y > 0) return 6;
        if (p.y == 0) return 7;
        if (p.y < 0) return 8;
    }
    if (p.x == 0 && p.z == 0 && p.y < 0 && p.y > -7) return 9;
    if (p.x == 1 && p.z == 0 && p.y < -5 && p.y > -12) return 9;
    if (p.x == 2 && p.z == 0 && p.y < -10) return 9;
    return 0;
}
static int large_mushroom(v3 p, int dir) /* G:596-618 */
{
    if (sdRoundBox(p, V3(3.0f, 0.5f, 3.0f), 1.5f) <= 0) {
        if (p.y > 0) return 6;
        if (p.y == 0) return 8;
        if (p.y < 0) return 7;
    }
    if (p.x == 0 && p.z == 0 && p.y < 0 && p.y > -9) return 9;
    if (p.x == 0 && p.z == (float)dir && p.y < -7 && p.y > -18) return 9;
    if (p.x == 0 && p.z == (float)(2 * dir) && p.y < -16) return 9;
    return 0;
}
static int all_mushrooms(v3 c) /* G:630-697 */
{
    if (c.x < 0 && c.z > 0) {
        if (c.x < -16) {
            if (c.z > 20) return tiny_mushroom(vsub(c, V3(-19, -12, 22)));
            if (c.z < 4) return tiny_mushroom(vsub(c, V3(-18, -12, 2)));
            int check = large_mushroom(vsub(c, V3(-22, 3, 8)), -1);
            if (check != 0) return check;
            check = medium_mushroom(vsub(c, V3(-27, -4, 16)));
            if (check != 0) return check;
            return 0;
        } else {
            if (c.z > 10 && c.x > -6) return tiny_mushroom(vsub(c, V3(-4, -14, 12)));
            if (c.z < 14) return medium_mushroom(vsub(c, V3(-4, -1, 6)));
            return small_mushroom(vsub(c, V3(-10, -8, 18)));
        }
    }
    if (c.x < 0 && c.z < 0) {
        if (c.x < -16) {
            if (c.x < -28) {
                if (c.z < -16) return tiny_mushroom(vsub(c, V3(-32, -14, -20)));
                return tiny_mushroom(vsub(c, V3(-30, -12, -12)));
            }
            if (c.z > -10) return small_mushroom(vsub(c, V3(-25, -7, -4)));
            return medium_mushroom(vsub(c, V3(-20, -3, -20)));
        } else {
            if (c.x < -12 && c.z > -12) return tiny_mushroom(vsub(c, V3(-14, -15, -10)));
            if (c.z > -10 && c.x > -4) return tiny_mushroom(vsub(c, V3(-2, -12, -2)));
            if (c.z < -10) return small_mushroom(vsub(c, V3(-5, -9, -14)));
            return large_mushroom(vsub(c, V3(-8, 8, -6)), 1);
        }
    }
    if (c.x > 0 && c.z < 0) {
        if (c.z > -5) return tiny_mushroom(vsub(c, V3(6, -14, -3)));
        if (c.z < -14) {
            if (c.x > 18) return tiny_mushroom(vsub(c, V3(20, -7, -16)));
            return large_mushroom(vsub(c, V3(14, 10, -20)), -1);
        }
        return medium_mushroom(vsub(c, V3(6, -6, -10)));
    }
    return 0;
}

static int getBlockAt_procedural(v3 c, int scene) /* G:699-826 */
{
    if (scene == 0) { /* cave, G:720-756 */
        if (c.y > 17.0f) return 0;
        if (c.y < -15) {
            if (c.y < -18) {
                float r = fbm(c.x * 0.3f, c.z * 0.3f);
                int d = f2i(floorf(r * 2.0f));
                if (d == 0) return 12;
            }
            float r = fbm(c.x * 0.058f, c.z * 0.058f);
            int d = f2i(floorf(r * 5.0f));
            if ((float)(-21 + d) >= c.y) {
                if (c.y == -18) return 13;
                return 11;
            }
        }
        if (sdSphere(c, 20.0f) > 0.0f)
            if (sdSphere(vadd(c, V3(16, 8, -10)), 20.0f) > 0.0f)
                if (sdSphere(vadd(c, V3(-13, -1, 19)), 18.0f) > 0.0f)
                    if (sdSphere(vadd(c, V3(20, 15, 15)), 21.0f) > 0.0f) return 10;
        return all_mushrooms(c);
    } else if (scene == 1) { /* Cornell, G:758-791 */
        if (c.x == -10)
            if (fabsf(c.y) < 10 && fabsf(c.z - 15) < 10) return 2;
        if (c.x == 10)
            if (fabsf(c.y) < 10 && fabsf(c.z - 15) < 10) return 3;
        if (fabsf(c.y) == 10)
            if (fabsf(c.x) < 10 && fabsf(c.z - 15) < 10) return 5;
        if (c.z == 25)
            if (fabsf(c.x) < 10 && fabsf(c.y) < 10) return 5;
        if (fabsf(c.x + 3) < 3 && fabsf(c.y + 7) < 3 && fabsf(c.z - 13) < 3) return 5;
        if (fabsf(c.x - 4) < 3 && fabsf(c.y + 4) < 6 && fabsf(c.z - 16) < 3) return 5;
    } else if (scene == 2) { /* house, G:793-820 */
        if (c.y == -5) return 1;
        if (fabsf(c.x) == 25)
            if (fabsf(c.y) < 5 && fabsf(c.z) < 15) return 2;
        if (c.y == 5)
            if (fabsf(c.x) < 25 && fabsf(c.z) < 15) return 5;
        if (c.z == -15)
            if (fabsf(c.x) < 25 && fabsf(c.y) < 5) return 3;
        if (c.z == 15) {
            if (fabsf(c.x - 10) < 2 && fabsf(c.y + 1) < 4) return 0;
            if (fabsf(c.x) < 25 && fabsf(c.y) < 5) return 3;
        }
    } else {
        return 0;
    }
    return 0;
}

/* stored-voxel lookup (extension): c is integer-valued (ceil of a position);
   anything outside the grid (or NaN) is empty. */
static int getBlockAt_voxels(const OrcParams* P, v3 c, int* oob)
{
    float lox = (float)P->vorg[0], loy = (float)P->vorg[1], loz = (float)P->vorg[2];
    float hix = (float)(P->vorg[0] + P->vdim[0] - 1);
    float hiy = (float)(P->vorg[1] + P->vdim[1] - 1);
    float hiz = (float)(P->vorg[2] + P->vdim[2] - 1);
    if (!(c.x >= lox && c.x <= hix && c.y >= loy && c.y <= hiy && c.z >= loz && c.z <= hiz)) {
        if (oob) (*oob)++;
        return 0;
    }
    int gx = (int)c.x - P->vorg[0], gy = (int)c.y - P->vorg[1], gz = (int)c.z - P->vorg[2];
    return P->vox[((size_t)gz * P->vdim[1] + gy) * P->vdim[0] + gx];
}

/* per-invocation counters */
typedef struct { uint32_t lookups; uint32_t oob; } Counters;

static int getBlockAt(const OrcParams* P, v3 c, Counters* cnt)
{
    cnt->lookups++;
    if (P->scene_mode == 0) return getBlockAt_procedural(c, P->scene);
    int o = 0;
    int b = getBlockAt_voxels(P, c, &o);
    cnt->oob += (uint32_t)o;
    return b;
}

/* ------------------------------------------------------------------------- */
/* getColorAt (G:828-1047)                                                    */
/* ------------------------------------------------------------------------- */

static v2 getUVs(v3 point, v3 normal) /* G:828-863 */
{
    v2 uv = {0, 0};
    if (normal.y == 0) {
        if (normal.x == 0) {
            if (gsign(normal.z) > 0) {
                uv.x = ceilf(point.x) - point.x;
                uv.y = point.y - floorf(point.y);
            } else {
                uv.x = point.x - floorf(point.x);
                uv.y = point.y - floorf(point.y);
            }
        } else {
            if (gsign(normal.x) < 1) {
                uv.x = ceilf(point.z) - point.z;
                uv.y = point.y - floorf(point.y);
            } else {
                uv.x = point.z - floorf(point.z);
                uv.y = point.y - floorf(point.y);
            }
        }
    } else {
        if (gsign(normal.y) < 0) {
            uv.x = point.x - floorf(point.x);
            uv.y = ceilf(point.z) - point.z;
        } else {
            uv.x = point.x - floorf(point.x);
            uv.y = point.z - floorf(point.z);
        }
    }
    return uv;
}
static float dotsPattern(v2 point, float radius, float cellSize) /* G:865-870 */
{
    float c = 4.0f * radius * cellSize;
    float h = c / 2.0f;
    float px = gmod(point.x + h, c) - h;
    float py = gmod(point.y + h, c) - h;
    return len2(px, py) - radius;
}
static v3 mix3(v3 a, v3 b, float t) { return V3(gmix(a.x, b.x, t), gmix(a.y, b.y, t), gmix(a.z, b.z, t)); }
static v3 ceil3(v3 a) { return V3(ceilf(a.x), ceilf(a.y), ceilf(a.z)); }

/* The flat-colour variant (README.md:266 "fill the cave with flat colors"):
   types 2-5 are the reference's own flat colours (G:908-919); the others take
   the base / commented-out flat colour of their branch (G:894-906, 921, 929,
   938, 955, 965, 1008, 1023, 1036). */
ORC_API void orc_default_palette(float* pal /* 256*3 */)
{
    static const float t[14][3] = {
        {0.f, 0.f, 0.f},          /* 0 empty */
        {0.99f, 0.3f, 0.3f},      /* 1 noise (G:906, r = 0.3) */
        {.95f, 0.f, 0.f},         /* 2 red   G:909 */
        {0.f, .95f, 0.f},         /* 3 green G:912 */
        {0.f, 0.f, .95f},         /* 4 blue  G:915 */
        {0.95f, 0.95f, 0.95f},    /* 5 white G:918 */
        {1.f, 0.2f, 0.f},         /* 6 mushroom 1: orange G:921 */
        {1.f, 0.f, 0.011f},       /* 7 mushroom 2: dark orange G:929 */
        {1.f, 0.5f, 0.f},         /* 8 mushroom 3: G:938 (commented flat) */
        {1.f, 0.5f, 0.f},         /* 9 stem: G:955 (commented flat) */
        {1.f, 0.5f, 0.f},         /* 10 cave wall: G:965 (commented flat) */
        {1.f, 0.f, 0.f},          /* 11 cave ground: G:1008 (commented flat) */
        {0.619f, 1.f, 0.278f},    /* 12 moss: G:1023 (commented flat) */
        {0.356f, 1.f, 0.101f},    /* 13 mold: G:1036 (commented flat) */
    };
    memset(pal, 0, 256 * 3 * sizeof(float));
    memcpy(pal, t, sizeof(t));
}

static v3 getColorAt_literal(v3 point, int block_type, v3 normal) /* G:872-1047 */
{
    if (block_type == 1) {
        float r = 0.3f; /* G:890-891: random1 result is overwritten */
        if (point.x < 0 && point.z > 0) {
            if (point.x < -16) return V3(0.8f, 0.4f, 0.2f);
            return V3(0.1f, r, 0.2f);
        }
        if (point.x < 0 && point.z < 0) {
            if (point.x < -16) return V3(0.4f, 0.8f, 0.2f);
            return V3(0.99f, r, r);
        }
        if (point.x > 0 && point.z < 0) return V3(0.1f, r, 0.5f);
        return V3(0.99f, r, r);
    } else if (block_type == 2) {
        return V3(.95f, 0, 0);
    } else if (block_type == 3) {
        return V3(0, .95f, 0);
    } else if (block_type == 4) {
        return V3(0, 0, .95f);
    } else if (block_type == 5) {
        return V3(0.95f, 0.95f, 0.95f);
    } else if (block_type == 6) {
        v2 pz = {point.x, point.z};
        float w = worleyNoise(pz);
        if (w < 0.35f) return V3(1, 0, 0.223f);
        return V3(1, 0.2f, 0);
    } else if (block_type == 7) {
        v3 dark_orange = V3(1, 0, 0.011f);
        v3 green = V3(0.8f, 1, 0);
        v2 pz = {point.x + 5.0f, point.z + 5.0f};
        float w = worleyNoise(pz);
        if (w < 0.25f) {
            /* green - (w * (vec3(0.5) - green)) */
            return V3(green.x - (w * (0.5f - green.x)), green.y - (w * (0.5f - green.y)),
                      green.z - (w * (0.5f - green.z)));
        }
        return dark_orange;
    } else if (block_type == 8) {
        v3 light_orange = V3(1, 0.313f, 0);
        v3 dark_purple = V3(1, 0, 0.223f);
        v2 g = getUVs(point, normal);
        /* mat2(0.707,-0.707,0.707,0.707) * g, column major */
        v2 uv;
        uv.x = 0.707f * g.x + 0.707f * g.y;
        uv.y = -0.707f * g.x + 0.707f * g.y;
        float radius = 0.05f;
        float dist = dotsPattern(uv, radius, 1.8f);
        float circle = (radius - dist) * 100.0f;
        float alpha = gclamp(circle, 0.0f, 1.0f);
        return mix3(light_orange, dark_purple, alpha);
    } else if (block_type == 9) {
        v2 uvs = getUVs(point, normal);
        float val = fbm(uvs.x * 5.0f, point.z);
        val += 0.5f * fbm1D(point.x);
        val = gclamp(val, 0.0f, 1.0f);
        return mix3(V3(0.3f, 0.1f, 0.3f), V3(0.9f, 0.9f, 0.9f), val);
    } else if (block_type == 10) {
        v3 color = V3(0.568f, 0.133f, 0.439f);
        if (point.y < -8) color = V3(0.349f, 0.133f, 0.427f);
        else if (point.y < -6) color = V3(0.568f, 0.133f, 0.439f);
        else if (point.y < -5) color = V3(0.639f, 0.176f, 0.725f);
        else if (point.y < 0) color = V3(0.274f, 0.188f, 0.772f);
        else if (point.y < 4) color = V3(0.341f, 0.270f, 0.768f);
        else if (point.y < 6) color = V3(0.368f, 0.203f, 0.415f);
        else if (point.y < 11) color = V3(0.470f, 0.270f, 0.729f);
        v2 uv = getUVs(point, normal);
        float r = fbm(0.05f, (uv.y + point.y) * 0.3f);
        v3 wallColor = V3(0, 0.666f, 1);
        if (point.x < -1) {
            wallColor = V3(0.294f, 0.007f, 0.152f);
        } else if (point.x < 6 && point.x >= -1) {
            float gradient = point.x / 7.0f;
            float r2 = random1(ceil3(point));
            if (r2 < gradient) wallColor = V3(0, 0.666f, 1);
            else wallColor = V3(0.294f, 0.007f, 0.152f);
        }
        return mix3(wallColor, color, r);
    } else if (block_type == 11) {
        v3 color = V3(0.294f, 0.007f, 0.152f);
        v3 moldcolor = V3(0.901f, 0.992f, 0.427f);
        float r = random1(ceil3(point)) / 3.0f;
        v3 combined = mix3(color, moldcolor, r);
        v2 uv = getUVs(point, normal);
        r = fbm(uv.x * 2.0f, uv.y * 2.0f);
        return mix3(combined, V3(0.294f, 0.007f, 0.152f), r / 2.0f);
    } else if (block_type == 12 || block_type == 13) {
        v2 uv = getUVs(point, normal);
        v3 base_green = block_type == 12 ? V3(0.356f, 1, 0.101f) : V3(0.803f, 1, 0.341f);
        v3 base_purple = V3(0.619f, 1, 0.278f);
        float ax = uv.x - 0.5f, ay = uv.y - 0.5f;
        float inv = 1.0f / sqrtf(ax * ax + ay * ay); /* normalize(vec2) */
        ax = ax * inv; ay = ay * inv;
        float r = interpNoise2D(ax, ay);
        float dist = len2(uv.x - 0.5f, uv.y - 0.5f);
        return mix3(base_green, base_purple, 2.f * dist + r * 0.3f);
    }
    return V3(0, 0, 0); /* PIN 4 */
}

static v3 getColorAt(const OrcParams* P, v3 point, int type, v3 normal)
{
    if (P->color_mode == 0) return getColorAt_literal(point, type, normal);
    const float* c = P->palette + 3 * (type & 255);
    return V3(c[0], c[1], c[2]);
}

/* ------------------------------------------------------------------------- */
/* intersection (G:59-121, 1051-1100, 1244-1301)                              */
/* ------------------------------------------------------------------------- */

typedef struct { v3 origin, direction; } Ray;

typedef struct {
    float t;
    v3 pos;
    v3 normal;
    v3 base_color; /* mat.base_color */
    v3 emissive;   /* mat.emissive */
    int type;      /* 2 light, 3 block; PIN 2: zero otherwise */
} Isect;

static const float ORC_INF = INFINITY;

static int intersect_sphere(Ray ray, float mint, float maxt, Isect* info) /* G:78-121 */
{
    float A = vdot(ray.direction, ray.direction);
    float B = -vdot(ray.direction, ray.origin);
    float C = vdot(ray.origin, ray.origin) - 1.0f;
    float D = B * B - A * C;
    D = D > 0 ? sqrtf(D) : ORC_INF;
    float t1 = (B - D) / A;
    float t2 = (B + D) / A;
    t1 = (mint < t1 && t1 < maxt) ? t1 : ORC_INF;
    t2 = (mint < t2 && t2 < maxt) ? t2 : ORC_INF;
    info->t = gmin(t1, t2);
    info->pos = vadd(ray.origin, vscale(ray.direction, info->t));
    info->normal = info->pos;
    return info->t < ORC_INF;
}

static int grid_march(const OrcParams* P, Ray ray, Isect* info, Counters* cnt) /* G:1051-1100 */
{
    v3 ray_origin = ray.origin;
    v3 ray_dir = vnormalize(ray.direction);
    float curr_t = 0.0f;
    for (int i = 0; i < 125; i++) {
        v3 f = V3(gfract(ray_origin.x), gfract(ray_origin.y), gfract(ray_origin.z));
        v3 t2;
        t2.x = gmax((-f.x) / ray_dir.x, (1.0f - f.x) / ray_dir.x);
        t2.y = gmax((-f.y) / ray_dir.y, (1.0f - f.y) / ray_dir.y);
        t2.z = gmax((-f.z) / ray_dir.z, (1.0f - f.z) / ray_dir.z);
        float min_val = gmin(gmin(t2.x, t2.y), t2.z) + 0.0001f;
        curr_t += min_val;
        ray_origin = vadd(ray.origin, vscale(ray_dir, curr_t));
        v3 cell = ceil3(ray_origin);
        v3 pi = V3(cell.x - 0.5f, cell.y - 0.5f, cell.z - 0.5f);
        int block_type = getBlockAt(P, cell, cnt);
        if (block_type > 0) {
            info->t = curr_t;
            v3 diff = vnormalize(vsub(ray_origin, pi));
            v3 normal = V3(0, 0, 0);
            float mx = 0.0f;
            for (int a = 0; a < 3; a++) {
                float da = vget(diff, a);
                if (fabsf(da) > mx) {
                    mx = fabsf(da);
                    normal = V3(0, 0, 0);
                    float sg = gsign(da) * 1.0f;
                    if (a == 0) normal.x = sg; else if (a == 1) normal.y = sg; else normal.z = sg;
                }
            }
            info->normal = vnormalize(normal);
            info->base_color = getColorAt(P, ray_origin, block_type, vnormalize(normal));
            info->emissive = V3(0, 0, 0); /* PIN 2 */
            return 1;
        }
    }
    return 0;
}

static int intersect_scene(const OrcParams* P, Ray ray, Isect* info, Counters* cnt) /* G:1244-1301 */
{
    const float mint = 0.0f;
    float closest_t = ORC_INF;
    memset(info, 0, sizeof(*info)); /* PIN 2 */
    info->t = closest_t;
    Isect temp;
    memset(&temp, 0, sizeof(temp));
    for (int i = 0; i < P->n_lights; i++) {
        const OrcLight* l = &P->lights[i];
        Ray tr;
        tr.origin = vdivs(vsub(ray.origin, V3(l->pos[0], l->pos[1], l->pos[2])), 0.1f);
        tr.direction = vdivs(ray.direction, 0.1f);
        intersect_sphere(tr, mint, closest_t, &temp);
        if (temp.t < closest_t) {
            *info = temp;
            info->base_color = V3(0, 0, 0); /* PIN 2 */
            info->emissive = V3(l->col[0], l->col[1], l->col[2]);
            info->type = 2;
        }
        closest_t = gmin(temp.t, closest_t);
    }
    if (grid_march(P, ray, &temp, cnt)) {
        if (temp.t < closest_t) {
            /* info = temp_isect: pos is stale (overwritten below) */
            info->t = temp.t;
            info->normal = temp.normal;
            info->base_color = temp.base_color;
            info->emissive = temp.emissive;
            closest_t = info->t;
            info->type = 3;
        }
    }
    info->normal = closest_t < ORC_INF ? vnormalize(info->normal) : V3(0, 0, 0);
    info->pos = closest_t < ORC_INF ? vadd(ray.origin, vscale(ray.direction, info->t)) : V3(0, 0, 0);
    info->pos = vadd(info->pos, vscale(info->normal, 0.001f));
    return closest_t < ORC_INF;
}

/* ------------------------------------------------------------------------- */
/* probe pass (P:45-71, 139-215, 253-303)                                     */
/* ------------------------------------------------------------------------- */

static uint32_t wang_hash(uint32_t seed) /* P:45-53 */
{
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
static uint32_t rand_xorshift(uint32_t* st) /* P:59-66 */
{
    *st ^= (*st << 13);
    *st ^= (*st >> 17);
    *st ^= (*st << 5);
    return *st;
}
static float rng_rand(uint32_t* st) /* P:68-71 */
{
    return (float)rand_xorshift(st) / 4294967296.0f;
}

ORC_API uint32_t orc_wang_hash(uint32_t s) { return wang_hash(s); }
ORC_API void orc_rand_sequence(uint32_t seed_index, int n, float* out)
{
    uint32_t st = wang_hash(seed_index);
    for (int i = 0; i < n; i++) out[i] = rng_rand(&st);
}

#define ORC_TWO_PI 6.2831853071795864769252867665590057683943f   /* P:147 */
#define ORC_SQRT_OF_ONE_THIRD 0.5773502691896257645091487805019574556476f /* P:148 */

static v3 calculate_random_dir_hemisphere(v3 normal, uint32_t* st) /* P:150-178 */
{
    float up = sqrtf(rng_rand(st));
    float over = sqrtf(1.0f - up * up);
    float around = rng_rand(st) * ORC_TWO_PI;
    v3 dnn;
    if (fabsf(normal.x) < ORC_SQRT_OF_ONE_THIRD) dnn = V3(1, 0, 0);
    else if (fabsf(normal.y) < ORC_SQRT_OF_ONE_THIRD) dnn = V3(0, 1, 0);
    else dnn = V3(0, 0, 1);
    v3 p1 = vnormalize(vcross(normal, dnn));
    v3 p2 = vnormalize(vcross(normal, p1));
    float sn, cs;
    pin_sincos(around, &sn, &cs);
    v3 a = vscale(normal, up);
    v3 b = vscale(p1, cs * over);
    v3 c = vscale(p2, sn * over);
    return vadd(vadd(a, b), c);
}

ORC_API void orc_hemisphere(const float* normal, uint32_t seed_index, float* out3)
{
    uint32_t st = wang_hash(seed_index);
    v3 d = calculate_random_dir_hemisphere(V3(normal[0], normal[1], normal[2]), &st);
    out3[0] = d.x; out3[1] = d.y; out3[2] = d.z;
}

static v3 light_pos(const OrcLight* l) { return V3(l->pos[0], l->pos[1], l->pos[2]); }
static v3 light_col(const OrcLight* l) { return V3(l->col[0], l->col[1], l->col[2]); }

static v3 get_direct_lighting(const OrcParams* P, const Isect* info, Counters* cnt) /* P:180-215 */
{
    v3 direct = V3(0, 0, 0);
    int num_visible = 0;
    for (int i = 0; i < P->n_lights; i++) {
        const OrcLight* l = &P->lights[i];
        v3 lp = light_pos(l);
        Ray feeler;
        feeler.origin = info->pos;
        feeler.direction = vnormalize(vsub(lp, info->pos));
        Isect ti;
        if (intersect_scene(P, feeler, &ti, cnt)) {
            float lambert = gclamp(vdot(vnormalize(info->normal), vnormalize(vsub(lp, info->pos))), 0.0f, 1.0f);
            if (ti.type == 2) {
                float dist = vlen(vsub(lp, info->pos));
                /* lambert * l.col * l.intensity / dist */
                v3 c = vscale(light_col(l), lambert);
                c = vscale(c, l->intensity);
                c = vdivs(c, dist);
                direct = vadd(direct, c);
            } else {
                return vscale(vscale(info->base_color, 0.2f), lambert); /* 0.2 * base * lambert */
            }
            num_visible++;
        }
    }
    if (num_visible != 0) {
        return vdivs(vmul(info->base_color, direct), (float)num_visible);
    }
    return V3(0, 0, 0);
}

/* PIN 7 */
static uint32_t unorm8(float x)
{
    if (x != x) return 0;
    float c = gclamp(x, 0.0f, 1.0f);
    return (uint32_t)floorf(c * 255.0f + 0.5f);
}
static uint32_t pack_rgba8(float r, float g, float b, float a)
{
    return unorm8(r) | (unorm8(g) << 8) | (unorm8(b) << 16) | (unorm8(a) << 24);
}

/* One probe_pass invocation (P:253-303) for linear ray index k. */
static void probe_invocation(const OrcParams* P, const OrcProbeRay* rays, uint32_t k, int W,
                             uint32_t* tex_albedo, uint32_t* tex_dist, float* tex_f32,
                             uint32_t* steps, uint32_t* oob)
{
    Counters cnt = {0, 0};
    uint32_t rng = wang_hash(k); /* P:55-57 with PIN 6: p_idx = gid.x + gid.y*W = k */
    const OrcProbeRay* pr = &rays[k];
    Ray r;
    r.origin = V3(pr->origin[0], pr->origin[1], pr->origin[2]);
    r.direction = V3(pr->direction[0], pr->direction[1], pr->direction[2]);

    int probe_index = f2i(pr->probe_info[0]);
    /* get_texture_coords_of_probe_index P:139-145 */
    int texture_width = P->probe_count[0] * P->probe_count[2];
    int y_probe = probe_index / texture_width;
    int x_probe = probe_index - (y_probe * texture_width);
    int tx = x_probe * P->rx + f2i(pr->probe_info[1]);
    int ty = y_probe * P->ry + f2i(pr->probe_info[2]);

    v3 color = V3(0, 0, 0);
    Ray indirect = r;
    Isect hit;
    for (int i = 0; i < P->max_bounces; i++) {
        if (intersect_scene(P, indirect, &hit, &cnt)) {
            color = vadd(color, get_direct_lighting(P, &hit, &cnt));
        } else {
            break;
        }
        indirect.origin = vadd(hit.pos, vscale(hit.normal, 0.0001f));
        indirect.direction = calculate_random_dir_hemisphere(hit.normal, &rng);
    }
    color = vdivs(color, (float)P->max_bounces);

    size_t t = (size_t)ty * W + tx;
    tex_albedo[t] = pack_rgba8(color.x, color.y, color.z, 1.0f);
    if (tex_dist) tex_dist[t] = pack_rgba8(0.0f, 0.0f, 0.0f, 0.0f); /* P:276,302 */
    if (tex_f32) { tex_f32[4 * t + 0] = color.x; tex_f32[4 * t + 1] = color.y; tex_f32[4 * t + 2] = color.z; tex_f32[4 * t + 3] = 1.0f; }
    if (steps) steps[k] = cnt.lookups;
    if (oob) oob[k] = cnt.oob;
}

/* Runs invocations k in [k0,k1).  Textures are W x H, W = X*Z*rx, H = Y*ry. */
ORC_API void orc_probe_update(const OrcParams* P, const OrcProbeRay* rays, uint32_t k0, uint32_t k1,
                              uint32_t* tex_albedo, uint32_t* tex_dist, float* tex_f32,
                              uint32_t* steps, uint32_t* oob, int num_threads)
{
    int W = P->probe_count[0] * P->probe_count[2] * P->rx;
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = (int64_t)k0; k < (int64_t)k1; k++) {
        probe_invocation(P, rays, (uint32_t)k, W, tex_albedo, tex_dist, tex_f32, steps, oob);
    }
}

/* ------------------------------------------------------------------------- */
/* host ray generation (R:1145-1224)                                          */
/* ------------------------------------------------------------------------- */

#define ORC_HOST_PI 3.1415926 /* R:1145 (a double literal) */

/* generate_samples R:1147-1173, generalised to an rx x ry tile.
   reseed != 0 calls srand(1) first (= the never-seeded state of a fresh process). */
ORC_API void orc_generate_samples(int rx, int ry, int reseed, float* out /* rx*ry*3 */)
{
    if (reseed) srand(1);
    float inv_x = 1.f / (float)rx;
    float inv_y = 1.f / (float)ry;
    int i = 0;
    for (int y = 0; y < ry; y++) {
        for (int x = 0; x < rx; x++) {
            float j1 = (float)rand() / (float)RAND_MAX; /* PIN 5: x jitter first */
            float j2 = (float)rand() / (float)RAND_MAX;
            float sx = ((float)x + j1) * inv_x;
            float sy = ((float)y + j2) * inv_y;
            float z = 1 - (2 * sx);
            /* cosf(2.0f * PI * sample.y): double product, narrowed at the call */
            float ang = (float)(2.0f * ORC_HOST_PI * sy);
            out[3 * i + 0] = cosf(ang) * sqrtf(1 - (z * z));
            out[3 * i + 1] = sinf(ang) * sqrtf(1 - (z * z));
            out[3 * i + 2] = z;
            i++;
        }
    }
}

/* RVPT::generate_probe_rays R:1177-1224 */
ORC_API void orc_generate_probe_rays(const OrcParams* P, const float* samples, OrcProbeRay* out)
{
    int X = P->probe_count[0], Y = P->probe_count[1], Z = P->probe_count[2];
    int n = P->rx * P->ry;
    int num_probes = X * Y * Z;
    size_t o = 0;
    for (int p = 0; p < num_probes; p++) {
        int py = p / (X * Z);
        int leftover = p - (py * X * Z);
        int pz = leftover / X;
        int px = leftover - pz * X;
        /* probe_index_3d - ((dim - 1) / 2), integer division */
        v3 org = V3((float)(px - (X - 1) / 2), (float)(py - (Y - 1) / 2), (float)(pz - (Z - 1) / 2));
        org = vscale(org, (float)P->side_length);
        org = vadd(org, V3(P->field_origin[0], P->field_origin[1], P->field_origin[2]));
        int x = 0, y = 0;
        for (int i = 0; i < n; i++) {
            v3 d = vnormalize(V3(samples[3 * i], samples[3 * i + 1], samples[3 * i + 2]));
            OrcProbeRay* r = &out[o++];
            memset(r, 0, sizeof(*r));
            r->origin[0] = org.x; r->origin[1] = org.y; r->origin[2] = org.z;
            r->direction[0] = d.x; r->direction[1] = d.y; r->direction[2] = d.z;
            r->probe_info[0] = (float)p; r->probe_info[1] = (float)x; r->probe_info[2] = (float)y;
            x++;
            if (x >= P->rx) { x = 0; y++; }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* pixel pass (G:1152-1240, 1306-1409; I:27-106; K:29-51; C:162-191)          */
/* ------------------------------------------------------------------------- */

#define ORC_SHADER_PI 3.1415926535897932384626433832795f /* C:5 / P:4 */

static void get_text_coord_from_probe_number(const OrcParams* P, int probe_number, int* ox, int* oy) /* G:1152-1174 */
{
    int x_dim = P->probe_count[0] * P->probe_count[2];
    *ox = -1; *oy = -1;
    if (probe_number >= x_dim * P->probe_count[1]) return;
    if (probe_number < 0 || x_dim < 0) return;
    int r0 = f2i(gmod((float)probe_number, (float)x_dim));
    int r1 = probe_number / x_dim; /* int(floor(int / int)) */
    if (r1 >= P->probe_count[1]) return;
    *ox = r0 * P->rx;
    *oy = r1 * P->ry;
}

static v3 load_rgba8(const uint32_t* tex, int W, int x, int y) /* imageLoad rgba8 -> byte/255 */
{
    uint32_t v = tex[(size_t)y * W + x];
    return V3((float)(v & 255u) / 255.0f, (float)((v >> 8) & 255u) / 255.0f, (float)((v >> 16) & 255u) / 255.0f);
}

static v3 sample_probe_albedo(const OrcParams* P, const uint32_t* tex, int W, int probe_number, v3 dir) /* G:1176-1240, texture_to_sample = 0 */
{
    int cx, cy;
    get_text_coord_from_probe_number(P, probe_number, &cx, &cy);
    if (cx == -1 && cy == -1) return V3(1, 0, 1);
    v3 d = vnormalize(dir);
    int relx = f2i(((-1.0f * (d.z - 1.0f)) / 2.0f) * (float)P->rx);
    if (relx == P->rx) relx = 0;
    float sqrt_z = sqrtf(1.0f - (d.z * d.z));
    int rely = f2i((pin_acos(d.x / sqrt_z) / (2.0f * ORC_SHADER_PI)) * (float)P->ry);
    int sx = cx + relx, sy = cy + rely;
    v3 result = load_rgba8(tex, W, sx, sy);
    int count = 0;
    for (int x = -2; x <= 2; x++) {
        int temp = sx + x;
        if (temp < cx || temp >= cx + P->rx) continue;
        for (int y = -2; y <= 2; y++) {
            int ry_ = sy + y;
            if (ry_ < cy || ry_ >= cy + P->ry) continue;
            count++;
            result = vadd(result, load_rgba8(tex, W, temp, ry_));
        }
    }
    return vdivs(result, (float)count);
}

ORC_API void orc_sample_probe(const OrcParams* P, const uint32_t* tex, int probe_number, const float* dir, float* out3)
{
    int W = P->probe_count[0] * P->probe_count[2] * P->rx;
    v3 r = sample_probe_albedo(P, tex, W, probe_number, V3(dir[0], dir[1], dir[2]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

static v3 get_diffuse_gi(const OrcParams* P, const uint32_t* tex, int W, const Isect* info) /* G:1306-1409 */
{
    v3 pos = info->pos;
    v3 N = vnormalize(info->normal);
    v3 fo = V3(P->field_origin[0], P->field_origin[1], P->field_origin[2]);
    float side = (float)P->side_length;
    v3 q = vdivs(vsub(pos, fo), side);
    int base[3] = {f2i(floorf(q.x)), f2i(floorf(q.y)), f2i(floorf(q.z))};
    /* G:1316-1323: int(vec3) takes the x component for every axis */
    int lo = f2i(-floorf((float)P->probe_count[0] / 2.0f));
    int hi = f2i(floorf((float)P->probe_count[0] / 2.0f) - 1.0f);
    for (int i = 0; i < 3; i++)
        if (base[i] < lo || base[i] > hi) return V3(1, 0, 1);

    v3 base_world = vadd(V3((float)(base[0] * P->side_length), (float)(base[1] * P->side_length),
                            (float)(base[2] * P->side_length)), fo);
    v3 irradiance = V3(0, 0, 0);
    float sum_weight = 0.0f;
    v3 a = vdivs(vsub(pos, base_world), side);
    v3 alpha = V3(gclamp(a.x, 0.0f, 1.0f), gclamp(a.y, 0.0f, 1.0f), gclamp(a.z, 0.0f, 1.0f));
    int X = P->probe_count[0], Y = P->probe_count[1], Z = P->probe_count[2];
    for (int i = 0; i < 8; i++) {
        int off[3] = {(i >> 2) & 1, (i >> 1) & 1, i & 1};
        int sh[3] = {base[0] + off[0] + X / 2, base[1] + off[1] + Y / 2, base[2] + off[2] + Z / 2};
        int probe_index_1d = sh[1] * X * Z + sh[2] * X + sh[0];
        if (probe_index_1d < 0 || probe_index_1d >= X * Y * Z) return V3(1, 0, 1);
        v3 offf = V3((float)off[0], (float)off[1], (float)off[2]);
        v3 one_m = V3(1.0f - alpha.x, 1.0f - alpha.y, 1.0f - alpha.z);
        v3 tri = V3(gmix(one_m.x, alpha.x, offf.x), gmix(one_m.y, alpha.y, offf.y), gmix(one_m.z, alpha.z, offf.z));
        v3 probe_pos = vadd(base_world, vscale(offf, side));
        v3 dir = vnormalize(vsub(probe_pos, pos));
        float temp = gmax(0.0001f, (vdot(dir, N) + 1.0f) * 0.5f);
        float weight = temp * temp + 0.2f;
        /* G:1367-1383: Chebyshev term is computed and discarded (no effect) */
        weight = gmax(0.000001f, weight);
        const float crush = 0.2f;
        if (weight < crush) weight *= weight * weight * (1.f / (crush * crush));
        weight *= tri.x * tri.y * tri.z;
        v3 s = sample_probe_albedo(P, tex, W, probe_index_1d, N);
        irradiance = vadd(irradiance, vscale(s, weight));
        sum_weight += weight;
    }
    return vdivs(irradiance, sum_weight);
}

static v3 integrator_DDGI(const OrcParams* P, const uint32_t* tex, int W, Ray ray, Counters* cnt) /* I:27-106 */
{
    Isect info;
    int hit = intersect_scene(P, ray, &info, cnt);
    if (!hit) return V3(0.898f, 0.968f, 1.0f);
    if (info.type == 2) return info.emissive;
    v3 indirect = get_diffuse_gi(P, tex, W, &info);
    v3 direct = V3(0, 0, 0);
    int num_visible = 0;
    for (int i = 0; i < P->n_lights; i++) {
        const OrcLight* l = &P->lights[i];
        v3 lp = light_pos(l);
        Ray feeler;
        feeler.origin = info.pos;
        feeler.direction = vnormalize(vsub(lp, info.pos));
        Isect ti;
        if (intersect_scene(P, feeler, &ti, cnt)) {
            if (ti.type == 2) {
                float lambert = gclamp(vdot(vnormalize(info.normal), vnormalize(vsub(lp, info.pos))), 0.0f, 1.0f);
                float dist = vlen(vsub(lp, info.pos));
                v3 c = vscale(light_col(l), lambert);
                c = vscale(c, l->intensity);
                c = vdivs(c, dist);
                direct = vadd(direct, c);
                num_visible++;
            }
        }
    }
    if (num_visible != 0) {
        /* 0.5 * base * (direct / n) + 0.5 * base * indirect */
        v3 hb = vscale(info.base_color, 0.5f);
        v3 a = vmul(hb, vdivs(direct, (float)num_visible));
        v3 b = vmul(hb, indirect);
        return vadd(a, b);
    }
    return vmul(vscale(indirect, 0.5f), info.base_color);
}

/* camera_pinhole_ray K:29-51.  cam = 16 floats column-major matrix + (aspect, hfov, scale, 0) */
static Ray camera_pinhole_ray(const float* cam, float x, float y)
{
    float aspect = cam[16], hfov = cam[17];
    float u = aspect * (2.0f * x - 1.0f);
    float v = 2.0f * y - 1.0f;
    float w = 1.0f / (float)tan((double)(0.5f * hfov)); /* PIN 8 */
    Ray r;
    r.origin = V3(cam[12], cam[13], cam[14]);
    v3 d;
    d.x = ((cam[0] * u + cam[4] * v) + cam[8] * w) + cam[12] * 0.0f;
    d.y = ((cam[1] * u + cam[5] * v) + cam[9] * w) + cam[13] * 0.0f;
    d.z = ((cam[2] * u + cam[6] * v) + cam[10] * w) + cam[14] * 0.0f;
    r.direction = vnormalize(d);
    return r;
}

/* compute_pass main C:162-191, render_mode 0, camera_mode 0.  frame is w x h,
   only pixels gid < (floor(w/16)*16, floor(h/16)*16) are written (PIN 9). */
ORC_API void orc_render_frame(const OrcParams* P, const float* cam, const uint32_t* tex_albedo,
                              uint32_t* frame, float* frame_f32, uint32_t* steps, int num_threads)
{
    int w = P->screen_width, h = P->screen_height;
    int W = P->probe_count[0] * P->probe_count[2] * P->rx;
    int wx = (w / 16) * 16, hy = (h / 16) * 16;
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int gy = 0; gy < hy; gy++) {
        for (int gx = 0; gx < wx; gx++) {
            Counters cnt = {0, 0};
            float cx = (float)gx / (float)w;
            float cy = (float)gy / (float)h;
            cy = 1.0f - cy;
            Ray ray = camera_pinhole_ray(cam, cx, cy);
            v3 s = integrator_DDGI(P, tex_albedo, W, ray, &cnt);
            s = vadd(V3(0, 0, 0), s); /* sampled += ... */
            size_t o = (size_t)gy * w + gx;
            frame[o] = pack_rgba8(s.x, s.y, s.z, 1.0f);
            if (frame_f32) { frame_f32[4 * o] = s.x; frame_f32[4 * o + 1] = s.y; frame_f32[4 * o + 2] = s.z; frame_f32[4 * o + 3] = 1.0f; }
            if (steps) steps[o] = cnt.lookups;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* unit-level entry points for tests                                          */
/* ------------------------------------------------------------------------- */

ORC_API int orc_get_block_at(const OrcParams* P, const float* c)
{
    Counters cnt = {0, 0};
    return getBlockAt(P, V3(c[0], c[1], c[2]), &cnt);
}

ORC_API void orc_get_color_at(const OrcParams* P, const float* point, int type, const float* normal, float* out3)
{
    v3 c = getColorAt(P, V3(point[0], point[1], point[2]), type, V3(normal[0], normal[1], normal[2]));
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}

/* Bakes getBlockAt_procedural over the grid described by P->vorg / P->vdim. */
ORC_API void orc_bake_scene(int scene, const int32_t* vorg, const int32_t* vdim, uint8_t* out, int num_threads)
{
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = 0; z < vdim[2]; z++)
        for (int y = 0; y < vdim[1]; y++)
            for (int x = 0; x < vdim[0]; x++) {
                v3 c = V3((float)(x + vorg[0]), (float)(y + vorg[1]), (float)(z + vorg[2]));
                out[((size_t)z * vdim[1] + y) * vdim[0] + x] = (uint8_t)getBlockAt_procedural(c, scene);
            }
}

/* intersect_scene for one ray: out = {hit, t, pos3, normal3, base3, type, lookups} (13 floats) */
ORC_API void orc_intersect_scene(const OrcParams* P, const float* origin, const float* dir, float* out)
{
    Counters cnt = {0, 0};
    Ray r;
    r.origin = V3(origin[0], origin[1], origin[2]);
    r.direction = V3(dir[0], dir[1], dir[2]);
    Isect info;
    int hit = intersect_scene(P, r, &info, &cnt);
    out[0] = (float)hit; out[1] = info.t;
    out[2] = info.pos.x; out[3] = info.pos.y; out[4] = info.pos.z;
    out[5] = info.normal.x; out[6] = info.normal.y; out[7] = info.normal.z;
    out[8] = info.base_color.x; out[9] = info.base_color.y; out[10] = info.base_color.z;
    out[11] = (float)info.type; out[12] = (float)cnt.lookups;
}

ORC_API int orc_sizeof_params(void) { return (int)sizeof(OrcParams); }
ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
