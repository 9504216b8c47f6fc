/* oracle/netphys_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see netphys_oracle.h).
 *
 * Scalar C restatement of the reference's per-step rigid-body path.  Every function names the
 * reference lines it follows (paths relative to the reference's engine/ directory).  All
 * arithmetic is IEEE binary32 evaluated exactly in the reference's operation order; build with
 * -ffp-contract=off (oracle/Makefile).  The only double arithmetic is inside npo_sinf/npo_cosf,
 * which restate the float sin/cos the reference's unqualified sin()/cos() calls resolve to.
 */
#include "netphys_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ============================================================================================
 * sinf / cosf
 * The reference calls sin(float)/cos(float) (matrix.h:283-288, physics_shape.cpp:53-63).  On the
 * oracle platform (x86-64, glibc 2.39, FMA-capable CPU) that is glibc's sinf/cosf, whose published
 * algorithm (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h: double-precision polynomial on a
 * quadrant-reduced argument, every a+b*c fused in the -mfma ifunc variant) is restated here.
 * glibc is third-party and absent from /root/reference; tests/test_sincos.py pins this restatement
 * against the box's libm (strided in the CPU suite, exhaustively via oracle/sincos_exhaustive.c).
 * ============================================================================================ */
typedef struct {
    double sign[4];
    double hpi_inv, hpi;
    double c0, c1, c2, c3, c4;
    double s1, s2, s3;
} sc_tab;

static const sc_tab SC[2] = {
    { { 1.0, -1.0, -1.0, 1.0 }, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0,
      0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16,
      -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13 },
    { { 1.0, -1.0, -1.0, 1.0 }, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0,
      -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16,
      -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13 },
};
/* 4/pi to 192 bits, in overlapping 32-bit windows */
static const uint32_t INV_PIO4[24] = {
    0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041,
};
static const double PI63 = 0x1.921FB54442D18p-62;

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ff; }

/* odd n -> cosine polynomial, even n -> sine polynomial; result rounded once to float */
static inline float sc_poly(double x, double x2, const sc_tab* p, int n)
{
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = fma(x2, p->s3, p->s2);
        double x7 = x3 * x2;
        double s = fma(x3, p->s1, x);
        return (float)fma(x7, s1, s);
    } else {
        double x4 = x2 * x2;
        double c2 = fma(x2, p->c4, p->c3);
        double c1 = fma(x2, p->c1, p->c0);
        double x6 = x4 * x2;
        double c = fma(x4, p->c2, c1);
        return (float)fma(x6, c2, c);
    }
}
static inline double sc_reduce_fast(double x, const sc_tab* p, int* np)
{
    double r = x * p->hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return fma(-(double)n, p->hpi, x);
}
static inline double sc_reduce_large(uint32_t xi, int* np)
{
    const uint32_t* arr = &INV_PIO4[(xi >> 26) & 15];
    int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = xi * arr[0];
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return x * PI63;
}
static float sc_eval(float y, int want_cos)
{
    double x = y;
    int n;
    const sc_tab* p = &SC[0];
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return want_cos ? 1.0f : y;
        return sc_poly(x, x2, p, want_cos);
    } else if (abstop12(y) < abstop12(120.0f)) {
        x = sc_reduce_fast(x, p, &n);
        double s = p->sign[n & 3];
        if (n & 2) p = &SC[1];
        return sc_poly(x * s, x * x, p, n ^ want_cos);
    } else if (abstop12(y) < abstop12(INFINITY)) {
        uint32_t xi = f2u(y);
        int sign = (int)(xi >> 31);
        x = sc_reduce_large(xi, &n);
        double s = p->sign[(n + sign) & 3];
        if ((n + sign) & 2) p = &SC[1];
        return sc_poly(x * s, x * x, p, n ^ want_cos);
    }
    return (y - y) / (y - y);
}
float npo_sinf(float x) { return sc_eval(x, 0); }
float npo_cosf(float x) { return sc_eval(x, 1); }

/* ============================================================================================
 * L0 math: vector.h, lib.h, matrix.h
 * ============================================================================================ */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v_ld(const float* p) { return V(p[0], p[1], p[2]); }
static inline void v_st(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3 v_neg(v3 a) { return V(-a.x, -a.y, -a.z); }                              /* vector.h operator-() */
static inline v3 v_sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v_add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v_mul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline float v_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }        /* vector.h dot */
static inline v3 v_cross(v3 a, v3 b)                                                     /* vector.h cross */
{ return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float v_magsq(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline float v_mag(v3 a) { return sqrtf(v_magsq(a)); }
static inline v3 v_normalize(v3 a) { float m = v_mag(a); return V(a.x / m, a.y / m, a.z / m); }  /* three divisions */

/* lib.h:13-15,32 -- max is a MACRO there; keep its NaN behaviour */
#define NPO_MAX(a, b) (((a) > (b)) ? (a) : (b))
static inline int feq(float a, float b, float tol)
{
    float fa = fabsf(a), fb = fabsf(b);
    float inner = NPO_MAX(fa, fb);
    return fabsf(a - b) < (tol * NPO_MAX(1.0f, inner));
}
#define TOL_DEFAULT 0.00001f                                                             /* lib.h:29 */
static inline int fle(float a, float b) { return a < b || feq(a, b, TOL_DEFAULT); }       /* lib.h:33 (tol arg ignored) */
static inline int fge(float a, float b) { return a > b || feq(a, b, TOL_DEFAULT); }       /* lib.h:34 */
static inline int v_eq_tol(v3 a, v3 b, float tol) { return feq(a.x, b.x, tol) && feq(a.y, b.y, tol) && feq(a.z, b.z, tol); }
static inline int v_eq(v3 a, v3 b) { return v_eq_tol(a, b, TOL_DEFAULT); }                /* vector3::operator== */
static inline int v_isnone(v3 a) { return feq(a.x, 0.0f, TOL_DEFAULT) && feq(a.y, 0.0f, TOL_DEFAULT) && feq(a.z, 0.0f, TOL_DEFAULT); }

int npo_float_equals(float a, float b, float tol) { return feq(a, b, tol); }

typedef struct { float m[9]; } m3;    /* x1 x2 x3 / y1 y2 y3 / z1 z2 z3 */
typedef struct { float m[16]; } m4;   /* x1..x4 / y1..y4 / z1..z4 / w1..w4 : rows */

static inline v3 m3_mul(const m3* a, v3 b)                                               /* matrix.h:59-66 */
{
    const float* m = a->m;
    return V(m[0] * b.x + m[1] * b.y + m[2] * b.z, m[3] * b.x + m[4] * b.y + m[5] * b.z, m[6] * b.x + m[7] * b.y + m[8] * b.z);
}
static inline float det2(float x1, float x2, float y1, float y2) { return x1 * y2 - x2 * y1; }  /* matrix.h:27-30 */
void npo_matrix3_inv(const float in[9], float out[9])                                    /* matrix.h:83-103 */
{
    float x1 = in[0], x2 = in[1], x3 = in[2], y1 = in[3], y2 = in[4], y3 = in[5], z1 = in[6], z2 = in[7], z3 = in[8];
    float d = x1 * y2 * z3 - x1 * y3 * z2 + x2 * y3 * z1 - x2 * y1 * z3 + x3 * y1 * z2 - x3 * y2 * z1;
    out[0] = det2(y2, y3, z2, z3) / d; out[1] = det2(x3, x2, z3, z2) / d; out[2] = det2(x2, x3, y2, y3) / d;
    out[3] = det2(y3, y1, z3, z1) / d; out[4] = det2(x1, x3, z1, z3) / d; out[5] = det2(x3, x1, y3, y1) / d;
    out[6] = det2(y1, y2, z1, z2) / d; out[7] = det2(x2, x1, z2, z1) / d; out[8] = det2(x1, x2, y1, y2) / d;
}
/* matrix4 * vector3 with w = 1 (matrix.h:145-157); the w row is computed by the reference and dropped */
static inline v3 m4_mul_point(const float* m, v3 b)
{
    return V(m[0] * b.x + m[1] * b.y + m[2] * b.z + m[3] * 1.0f,
             m[4] * b.x + m[5] * b.y + m[6] * b.z + m[7] * 1.0f,
             m[8] * b.x + m[9] * b.y + m[10] * b.z + m[11] * 1.0f);
}
/* Physics::GetTransform (physics.h:70-76): identity, translate (matrix.h:297-303), rotate (matrix.h:283-288) */
void npo_transform(const float pos[3], const float rot[3], float o[16])
{
    float sx = npo_sinf(rot[0]), cx = npo_cosf(rot[0]);
    float sy = npo_sinf(rot[1]), cy = npo_cosf(rot[1]);
    float sz = npo_sinf(rot[2]), cz = npo_cosf(rot[2]);
    o[3] = 0.0f + pos[0]; o[7] = 0.0f + pos[1]; o[11] = 0.0f + pos[2];
    o[12] = 0.0f; o[13] = 0.0f; o[14] = 0.0f; o[15] = 1.0f;
    o[0] = cy * cz;  o[1] = sx * sy * cz - cx * sz;  o[2] = cx * sy * cz + sx * sz;
    o[4] = cy * sz;  o[5] = sx * sy * sz + cx * cz;  o[6] = cx * sy * sz - sx * cz;
    o[8] = -sy;      o[9] = sx * cy;                 o[10] = cx * cy;
}

/* ============================================================================================
 * util.h closest-point solvers and search directions
 * ============================================================================================ */
enum { LINE_AB = 0, LINE_A, LINE_B };
enum { TRI_ABC = 0, TRI_A, TRI_B, TRI_C, TRI_AB, TRI_AC, TRI_BC };
enum { TET_NONE = -1, TET_ABCD = 0, TET_A, TET_B, TET_C, TET_D, TET_AB, TET_AC, TET_AD, TET_BC, TET_BD, TET_CD,
       TET_ABC, TET_ABD, TET_ACD, TET_BCD };

static int closest_line(v3 p, v3 a, v3 b, float* u)                                      /* util.h:18-37 */
{
    v3 ap = v_sub(p, a), ab = v_sub(b, a);
    *u = v_dot(ap, ab) / v_dot(ab, ab);
    if (*u <= 0.0f) { *u = 0.0f; return LINE_A; }
    if (*u >= 1.0f) { *u = 1.0f; return LINE_B; }
    return LINE_AB;
}
static int closest_triangle(v3 p, v3 va, v3 vb, v3 vc, float* pu, float* pv)             /* util.h:144-213 */
{
    v3 A = va, B = v_sub(vb, va), C = v_sub(vc, va);
    float a = v_dot(B, B);
    float b = v_dot(C, C);
    float c = 2 * (v_dot(B, C));
    float d = 2 * (v_dot(A, B) - v_dot(p, B));
    float e = 2 * (v_dot(A, C) - v_dot(p, C));
    float denom = 4 * a * b - c * c;
    float u = (e * c - 2 * b * d) / denom;
    float v = (c * d - 2 * e * a) / denom;
    int r;
    if (fle(u, 0.0f) && fle(v, 0.0f)) { u = 0.0f; v = 0.0f; r = TRI_A; }
    else if (fle(v, 0.0f) && fge(u, 1.0f)) { u = 1.0f; v = 0.0f; r = TRI_B; }
    else if (fle(u, 0.0f) && fge(v, 1.0f)) { u = 0.0f; v = 1.0f; r = TRI_C; }
    else if (fle(v, 0.0f) && u > 0.0f && fle(u, 1.0f)) { v = 0.0f; r = TRI_AB; }
    else if (fle(u, 0.0f) && v > 0.0f && fle(v, 1.0f)) { u = 0.0f; r = TRI_AC; }
    else if (u + v > 1.0f) { float k = 1.0f / (u + v); u *= k; v *= k; r = TRI_BC; }
    else r = TRI_ABC;
    *pu = u; *pv = v;
    return r;
}
static int closest_tetra(v3 P, v3 va, v3 vb, v3 vc, v3 vd, float* pu, float* pv, float* pw)  /* util.h:312-457 */
{
    v3 A = va, B = v_sub(vb, va), C = v_sub(vc, va), D = v_sub(vd, va);
    float a = v_dot(B, B), b = v_dot(C, C), c = v_dot(D, D);
    float d = v_dot(B, C), e = v_dot(B, D), f = v_dot(C, D);
    float g = v_dot(A, B) - v_dot(P, B);
    float h = v_dot(A, C) - v_dot(P, C);
    float i = v_dot(A, D) - v_dot(P, D);
    float denom = (a * b * c - a * f * f - b * e * e - c * d * d + 2 * d * e * f);
    if (feq(denom, 0.f, TOL_DEFAULT)) return TET_NONE;      /* u,v,w left untouched */
    float u = (-b * c * g + b * e * i + c * d * h - d * f * i - e * h * f + g * f * f) / denom;
    float v = (-a * c * h + a * f * i + c * d * g - d * e * i + h * e * e - e * g * f) / denom;
    float w = (-a * b * i + a * h * f + b * e * g + i * d * d - d * e * h - d * g * f) / denom;
    int r;
    if (u <= 0.0f && v <= 0.0f && w <= 0.0f) { u = 0.0f; v = 0.0f; w = 0.0f; r = TET_A; }
    else if (u >= 1.0f && v <= 0.0f && w <= 0.0f) { u = 1.0f; v = 0.0f; w = 0.0f; r = TET_B; }
    else if (u <= 0.0f && v >= 1.0f && w <= 0.0f) { u = 0.0f; v = 1.0f; w = 0.0f; r = TET_C; }
    else if (u <= 0.0f && v <= 0.0f && w >= 1.0f) { u = 0.0f; v = 0.0f; w = 1.0f; r = TET_D; }
    else if (u >= 0.0f && u <= 1.0f && v <= 0.0f && w <= 0.0f) { v = 0.0f; w = 0.0f; r = TET_AB; }
    else if (u <= 0.0f && v >= 0.0f && v <= 1.0f && w <= 0.0f) { u = 0.0f; w = 0.0f; r = TET_AC; }
    else if (u <= 0.0f && v <= 0.0f && w >= 0.0f && w <= 1.0f) { u = 0.0f; v = 0.0f; r = TET_AD; }
    else if (u >= 0.0f && v >= 0.0f && w <= 0.0f) {
        w = 0.0f;
        if (u + v >= 1.0f) { float k = u + v; u = u / k; v = v / k; r = TET_BC; } else r = TET_ABC;
    }
    else if (u >= 0.0f && v <= 0.0f && w >= 0.0f) {
        v = 0.0f;
        if (u + w >= 1.0f) { float k = u + w; u = u / k; w = w / k; r = TET_BD; } else r = TET_ABD;
    }
    else if (u <= 0.0f && v >= 0.0f && w >= 0.0f) {
        u = 0.0f;
        if (v + w > 1.0f) { float k = v + w; v = v / k; w = w / k; r = TET_CD; } else r = TET_ACD;
    }
    else if (u + v + w > 1.0f) { float k = u + v + w; u = u / k; v = v / k; w = w / k; r = TET_BCD; }
    else r = TET_ABCD;
    *pu = u; *pv = v; *pw = w;
    return r;
}
/* the two degenerate fallbacks shared by util.h:473-483 and :496-506 */
static v3 dir_fallback(v3 ab)
{
    v3 abt = v_cross(ab, V(0, 0, 1));
    if (v_magsq(abt) == 0.0f) return v_cross(ab, V(0, 1, 0));
    return abt;
}
static v3 dir1(v3 a) { return v_neg(a); }                                                /* util.h:461-464 */
static v3 dir2(v3 a, v3 b)                                                               /* util.h:466-486 */
{
    v3 ab = v_sub(b, a), ao = v_neg(b);
    v3 t = v_cross(ab, ao);
    v3 bt = v_cross(t, ab);
    float sign = v_dot(bt, ao);
    if (feq(sign, 0.0f, TOL_DEFAULT)) return dir_fallback(ab);
    return (sign < 0.0f) ? v_neg(bt) : bt;
}
static v3 dir3(v3 a, v3 b, v3 c)                                                         /* util.h:488-509 */
{
    v3 ab = v_sub(b, a), ac = v_sub(c, a), ao = v_neg(a);
    v3 n = v_cross(ab, ac);
    float sign = v_dot(n, ao);
    if (feq(sign, 0.0f, TOL_DEFAULT)) return dir_fallback(ab);
    return (sign < 0.0f) ? v_neg(n) : n;
}

int npo_closest_line(const float p[3], const float a[3], const float b[3], float* u) { return closest_line(v_ld(p), v_ld(a), v_ld(b), u); }
int npo_closest_triangle(const float p[3], const float a[3], const float b[3], const float c[3], float uv[2])
{ return closest_triangle(v_ld(p), v_ld(a), v_ld(b), v_ld(c), &uv[0], &uv[1]); }
int npo_closest_tetra(const float p[3], const float a[3], const float b[3], const float c[3], const float d[3], float uvw[3])
{ return closest_tetra(v_ld(p), v_ld(a), v_ld(b), v_ld(c), v_ld(d), &uvw[0], &uvw[1], &uvw[2]); }
void npo_dir_to_origin(int n, const float* pts, float out[3])
{
    v3 r = n == 1 ? dir1(v_ld(pts)) : n == 2 ? dir2(v_ld(pts), v_ld(pts + 3)) : dir3(v_ld(pts), v_ld(pts + 3), v_ld(pts + 6));
    v_st(out, r);
}

/* ============================================================================================
 * shapes: physics_shape.cpp
 * ============================================================================================ */
typedef struct { v3* v; int n, cap; } vlist;
static void vl_push(vlist* l, v3 p)
{
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 64; l->v = (v3*)realloc(l->v, sizeof(v3) * (size_t)l->cap); }
    l->v[l->n++] = p;
}
/* Mesh::AddVertex (physics_shape.cpp:23-37): tolerant de-duplication, first match wins */
static void mesh_add(vlist* mesh, v3 p)
{
    for (int i = 0; i < mesh->n; i++) if (v_eq(mesh->v[i], p)) return;
    vl_push(mesh, p);
}
static int mesh_out(vlist* mesh, float* xyz, int cap)
{
    int n = mesh->n;
    for (int i = 0; i < n && i < cap; i++) v_st(xyz + 3 * i, mesh->v[i]);
    free(mesh->v);
    return n;
}
int npo_make_box(float width, float depth, float height, float* xyz, int cap)            /* physics_shape.cpp:163-217 */
{
    (void)depth;                                    /* :166 half_depth = width / 2 -- depth is ignored */
    float hw = width / 2.0f, hd = width / 2.0f, hh = height / 2.0f;
    v3 c[8] = { V(hw, -hh, -hd), V(hw, -hh, hd), V(hw, hh, -hd), V(hw, hh, hd), V(-hw, -hh, -hd), V(-hw, -hh, hd), V(-hw, hh, -hd), V(-hw, hh, hd) };
    static const int tri[12][3] = { { 0, 1, 3 }, { 0, 3, 2 }, { 1, 5, 7 }, { 1, 7, 3 }, { 5, 4, 6 }, { 5, 6, 7 },
                                    { 4, 0, 2 }, { 4, 2, 6 }, { 2, 3, 7 }, { 2, 7, 6 }, { 4, 5, 1 }, { 4, 1, 0 } };
    vlist mesh = { 0, 0, 0 };
    for (int t = 0; t < 12; t++) for (int k = 0; k < 3; k++) mesh_add(&mesh, c[tri[t][k]]);
    return mesh_out(&mesh, xyz, cap);
}
int npo_make_sphere(float r, float* xyz, int cap)                                        /* physics_shape.cpp:42-162, :313-316 */
{
    const float PIf = 3.1415926535897932384626433832795028841971693993751f;             /* lib.h:10 */
    const float theta = 26.56505117707799f * PIf / 180.0f;
    vlist vl = { 0, 0, 0 };
    for (int i = 0; i < 12; i++) vl_push(&vl, V(0, 0, 0));
    float phi = 0;
    for (int i = 1; i < 6; i++) {
        vl.v[i].x = r * npo_cosf(theta) * npo_cosf(phi);
        vl.v[i].y = r * npo_cosf(theta) * npo_sinf(phi);
        vl.v[i].z = r * -npo_sinf(theta);
        phi = (float)((double)phi + (double)(2 * PIf) / 5.0);                            /* :56 float += double */
    }
    phi = 0;
    for (int i = 6; i < 11; i++) {
        vl.v[i].x = r * npo_cosf(theta) * npo_cosf(phi);
        vl.v[i].y = r * npo_cosf(theta) * npo_sinf(phi);
        vl.v[i].z = r * npo_sinf(theta);
        phi = (float)((double)phi + (double)(2 * PIf) / 5.0);
    }
    vl.v[0] = V(0.0f, 0.0f, -r);
    vl.v[11] = V(0.0f, 0.0f, r);
    static const int ico[20][3] = { { 0, 2, 1 }, { 0, 3, 2 }, { 0, 4, 3 }, { 0, 5, 4 }, { 0, 1, 5 }, { 1, 2, 7 }, { 2, 3, 8 }, { 3, 4, 9 },
                                    { 4, 5, 10 }, { 5, 1, 6 }, { 1, 7, 6 }, { 2, 8, 7 }, { 3, 9, 8 }, { 4, 10, 9 }, { 5, 6, 10 },
                                    { 6, 7, 11 }, { 7, 8, 11 }, { 8, 9, 11 }, { 9, 10, 11 }, { 10, 6, 11 } };
    int ntri = 20;
    int* idx = (int*)malloc(sizeof(int) * 3 * (size_t)ntri);
    memcpy(idx, ico, sizeof ico);
    for (int round = 1; round < 3; round++) {                                            /* CreateIcosahadron(r, 3): loop runs twice */
        vlist nv = { 0, 0, 0 };
        int* nidx = (int*)malloc(sizeof(int) * 3 * 4 * (size_t)ntri);
        for (int j = 0; j < ntri; j++) {
            v3 a = vl.v[idx[3 * j]], b = vl.v[idx[3 * j + 1]], c = vl.v[idx[3 * j + 2]];
            v3 x = v_mul(v_normalize(v_add(a, b)), r);
            v3 y = v_mul(v_normalize(v_add(a, c)), r);
            v3 z = v_mul(v_normalize(v_add(b, c)), r);
            v3 seq[12] = { a, x, y, x, b, z, z, y, x, z, c, y };                        /* axy, xbz, zyx, zcy */
            int base = nv.n;
            for (int k = 0; k < 12; k++) { vl_push(&nv, seq[k]); nidx[12 * j + k] = base + k; }
        }
        free(vl.v); free(idx);
        vl = nv; idx = nidx; ntri *= 4;
    }
    vlist mesh = { 0, 0, 0 };
    for (int i = 0; i < vl.n; i++) mesh_add(&mesh, vl.v[i]);
    free(vl.v); free(idx);
    return mesh_out(&mesh, xyz, cap);
}

/* ============================================================================================
 * support mapping: physics_util.cpp:5-35 (strict '>', first maximum wins, result starts at 0)
 * ============================================================================================ */
static v3 support(const float* xyz, int n, v3 dir, const float* xf)
{
    v3 result = V(0, 0, 0);
    float best = -INFINITY;
    for (int i = 0; i < n; i++) {
        v3 wp = m4_mul_point(xf, v_ld(xyz + 3 * i));
        float dp = v_dot(dir, wp);
        if (dp > best) { best = dp; result = wp; }
    }
    return result;
}
void npo_support(const float* xyz, int n, const float dir[3], const float xf[16], float out[3]) { v_st(out, support(xyz, n, v_ld(dir), xf)); }

/* ============================================================================================
 * simplex.h + collision_detection.cpp
 * ============================================================================================ */
typedef struct { v3 p, A, B; } spoint;
typedef struct { int i0, i1, i2; v3 normal; } sface;
typedef struct {
    int contains_origin;
    int nv; spoint v[4 + 24];        /* GJK <= 4; EPA appends at most one per iteration (<= 20) */
    int nf, fcap; sface* f;          /* faces: unbounded in the reference (std::vector) */
} simplex;

typedef struct { const float* va; int na; const float* xa; const float* vb; int nb; const float* xb; } cparams;

static void sx_add_face(simplex* s, int va, int vb, int vc, npo_counters* ctr)          /* simplex.h:37-53 */
{
    if (s->nf == s->fcap) { s->fcap = s->fcap ? 2 * s->fcap : 16; s->f = (sface*)realloc(s->f, sizeof(sface) * (size_t)s->fcap); }
    sface* f = &s->f[s->nf++];
    f->i0 = va; f->i1 = vb; f->i2 = vc;
    v3 a = s->v[va].p, b = s->v[vb].p, c = s->v[vc].p;
    v3 n = v_normalize(v_cross(v_sub(b, a), v_sub(c, a)));
    f->normal = (v_dot(n, a) != 0.0f) ? n : v_neg(n);                                    /* :52 float-as-bool */
    if (ctr) { ctr->epa_faces_added++; if (s->nf > ctr->epa_max_faces) ctr->epa_max_faces = s->nf; }
}

/* Solve* (collision_detection.cpp:18-259): keep the vertices named by the region, order preserved */
static void keep(simplex* s, int n, int i0, int i1, int i2)
{
    spoint a = s->v[i0], b = s->v[i1 < 0 ? 0 : i1], c = s->v[i2 < 0 ? 0 : i2];
    s->v[0] = a; if (n > 1) s->v[1] = b; if (n > 2) s->v[2] = c;
    s->nv = n;
}
static void solve_line(simplex* s)
{
    float u;
    switch (closest_line(V(0, 0, 0), s->v[0].p, s->v[1].p, &u)) {
    case LINE_A: keep(s, 1, 0, -1, -1); break;
    case LINE_B: keep(s, 1, 1, -1, -1); break;
    default: break;
    }
}
static void solve_triangle(simplex* s)
{
    float u, v;
    switch (closest_triangle(V(0, 0, 0), s->v[0].p, s->v[1].p, s->v[2].p, &u, &v)) {
    case TRI_A: keep(s, 1, 0, -1, -1); break;
    case TRI_B: keep(s, 1, 1, -1, -1); break;
    case TRI_C: keep(s, 1, 2, -1, -1); break;
    case TRI_AB: keep(s, 2, 0, 1, -1); break;
    case TRI_AC: keep(s, 2, 0, 2, -1); break;
    case TRI_BC: keep(s, 2, 1, 2, -1); break;
    default: break;
    }
}
static void solve_tetra(simplex* s)
{
    float u, v, w;
    switch (closest_tetra(V(0, 0, 0), s->v[0].p, s->v[1].p, s->v[2].p, s->v[3].p, &u, &v, &w)) {
    case TET_A: keep(s, 1, 0, -1, -1); break;
    case TET_B: keep(s, 1, 1, -1, -1); break;
    case TET_C: keep(s, 1, 2, -1, -1); break;
    case TET_D: keep(s, 1, 3, -1, -1); break;
    case TET_AB: keep(s, 2, 0, 1, -1); break;
    case TET_AC: keep(s, 2, 0, 2, -1); break;
    case TET_AD: keep(s, 2, 0, 3, -1); break;
    case TET_BC: keep(s, 2, 1, 2, -1); break;
    case TET_BD: keep(s, 2, 1, 3, -1); break;
    case TET_CD: keep(s, 2, 2, 3, -1); break;
    case TET_ABC: keep(s, 3, 0, 1, 2); break;
    case TET_ABD: keep(s, 3, 0, 1, 3); break;
    case TET_ACD: keep(s, 3, 0, 2, 3); break;
    case TET_BCD: keep(s, 3, 1, 2, 3); break;
    default: break;            /* ABCD, and None under NDEBUG (:255): all four stay => reported as overlap */
    }
}

enum { RES_NONE = 0, RES_OVERLAP, RES_NO_OVERLAP, RES_CONTINUE };

static int gjk_step(const cparams* p, int is3D, simplex* s, npo_counters* ctr)           /* collision_detection.cpp:620-692 */
{
    spoint old[4]; int nold = s->nv;
    memcpy(old, s->v, sizeof(spoint) * (size_t)nold);
    switch (s->nv) {
    case 2: solve_line(s); if (ctr) ctr->solve_line++; break;
    case 3: solve_triangle(s); if (ctr) ctr->solve_tri++; break;
    case 4: solve_tetra(s); if (ctr) ctr->solve_tet++; break;
    default: break;
    }
    if (s->nv == 4 || (!is3D && s->nv == 3)) { s->contains_origin = 1; return RES_OVERLAP; }
    v3 d;
    if (s->nv == 1) { d = dir1(s->v[0].p); if (ctr) ctr->dir1++; }
    else if (s->nv == 2) { d = dir2(s->v[0].p, s->v[1].p); if (ctr) ctr->dir2++; }
    else { d = dir3(s->v[0].p, s->v[1].p, s->v[2].p); if (ctr) ctr->dir3++; }
    if (feq(v_magsq(d), 0.0f, TOL_DEFAULT)) return RES_NO_OVERLAP;
    v3 sa = support(p->va, p->na, d, p->xa);
    v3 sb = support(p->vb, p->nb, v_neg(d), p->xb);
    v3 sp = v_sub(sa, sb);
    if (ctr) ctr->gjk_supports++;
    if (v_dot(d, sp) < 0) return RES_NO_OVERLAP;
    for (int i = 0; i < nold; i++) if (v_eq(sa, old[i].A) && v_eq(sb, old[i].B)) return RES_NO_OVERLAP;
    s->v[s->nv].A = sa; s->v[s->nv].B = sb; s->v[s->nv].p = sp; s->nv++;
    return RES_CONTINUE;
}

/* GetClosestTriangleToOrigin (:292-325). face index defaults to 0 when no face qualifies (every
 * distance NaN): that is what the reference does when its uninitialised local is pinned to zero. */
static void closest_face(const simplex* s, float* distance, int* face_index, npo_counters* ctr)
{
    float best = FLT_MAX; int bi = 0;
    for (int i = 0; i < s->nf; i++) {
        v3 a = s->v[s->f[i].i0].p, b = s->v[s->f[i].i1].p, c = s->v[s->f[i].i2].p;
        float u, v;
        closest_triangle(V(0, 0, 0), a, b, c, &u, &v);
        float dsq = v_magsq(v_add(v_add(a, v_mul(v_sub(b, a), u)), v_mul(v_sub(c, a), v)));
        if (dsq < best) { best = dsq; bi = i; }
    }
    if (ctr) ctr->epa_tri_tests += s->nf;
    *distance = sqrtf(best); *face_index = bi;
}
/* GetClosestEdgeToOrigin (:261-290), 2-D EPA */
static void closest_edge(const simplex* s, float* distance, float* u, int* si, int* ei)
{
    float best = FLT_MAX, bu = 0.0f; *si = 0; *ei = 0;
    for (int i = 0; i < s->nv; i++) {
        int e = (i == s->nv - 1) ? 0 : i + 1;
        v3 a = s->v[i].p, b = s->v[e].p; float iu;
        closest_line(V(0, 0, 0), a, b, &iu);
        float dsq = v_magsq(v_add(a, v_mul(v_sub(b, a), iu)));
        if (dsq < best) { best = dsq; *si = i; *ei = e; bu = iu; }
    }
    *distance = sqrtf(best); *u = bu;
}

/* FindIntersectionPointsStep (:412-547). returns 1 when converged (out filled), 0 to continue,
 * -1 when the polytope has no faces left (undefined in the reference; defined here as failure). */
static int epa_step(const cparams* p, simplex* s, int is3D, npo_contact* out, npo_counters* ctr)
{
    const float TOL = 0.01f;
    float distance = 0.0f; v3 normal; int fi = 0, ea = 0, eb = 0;
    if (ctr) ctr->epa_iters++;
    if (is3D) {
        if (s->nf == 0) return -1;
        closest_face(s, &distance, &fi, ctr);
        normal = s->f[fi].normal;                                                        /* :408 stored normal */
    } else {
        float u; closest_edge(s, &distance, &u, &ea, &eb);                               /* :364-384 */
        v3 a = s->v[ea].p, b = s->v[eb].p;
        v3 cp = v_add(a, v_mul(v_sub(b, a), u));
        v3 edge = v_sub(b, a);
        v3 n = V(edge.y, -edge.x, 0.0f);
        normal = v_dot(cp, n) > 0.0f ? n : v_neg(n);
    }
    if (distance > TOL) {
        v3 sa = support(p->va, p->na, normal, p->xa);
        v3 sb = support(p->vb, p->nb, v_neg(normal), p->xb);
        v3 sp = v_sub(sa, sb);
        if (ctr) ctr->epa_supports++;
        int close = 0;
        for (int i = 0; i < s->nv; i++) if (v_eq_tol(sa, s->v[i].A, TOL) && v_eq_tol(sb, s->v[i].B, TOL)) { close = 1; break; }
        if (!close) {
            spoint np_; np_.A = sa; np_.B = sb; np_.p = sp;
            if (!is3D) {                                                                 /* simplex.h:59-75 insert at eb */
                for (int i = s->nv; i > eb; --i) s->v[i] = s->v[i - 1];
                s->v[eb] = np_; s->nv++;
            } else {
                /* :469-510 drop faces whose stored normal faces the support POINT; horizon = edges
                 * whose reverse was not also added, in insertion order */
                int ne = 0, ecap = 3 * s->nf + 3;
                int (*edges)[2] = (int (*)[2])malloc(sizeof(int[2]) * (size_t)ecap);
                int keepn = 0;
                for (int i = 0; i < s->nf; i++) {
                    sface f = s->f[i];
                    if (v_dot(f.normal, sp) > 0) {
                        int e3[3][2] = { { f.i0, f.i1 }, { f.i1, f.i2 }, { f.i2, f.i0 } };
                        for (int k = 0; k < 3; k++) {
                            int va = e3[k][0], vb = e3[k][1], found = -1;
                            for (int q = 0; q < ne; q++) if (edges[q][0] == vb && edges[q][1] == va) { found = q; break; }
                            if (found >= 0) { memmove(&edges[found], &edges[found + 1], sizeof(int[2]) * (size_t)(ne - found - 1)); ne--; }
                            else { edges[ne][0] = va; edges[ne][1] = vb; ne++; }
                        }
                    } else s->f[keepn++] = f;
                }
                s->nf = keepn;
                int ni = s->nv; s->v[s->nv++] = np_;
                for (int i = 0; i < ne; i++) sx_add_face(s, edges[i][0], edges[i][1], ni, ctr);
                free(edges);
            }
            if (ctr && s->nv > ctr->epa_max_verts) ctr->epa_max_verts = s->nv;
            return 0;
        }
    }
    v_st(out->pen_dir, v_neg(v_normalize(normal)));                                      /* :530 */
    out->depth = distance;
    return 1;
}
/* FindIntersectionPoints (:549-608) with the :567 guard of oracle/Makefile (the 3-D non-overlap
 * branch never runs: simplex.faces is always empty there) */
static int epa(const cparams* p, simplex* s, int max_steps, int is3D, npo_contact* out, npo_counters* ctr)
{
    if (s->contains_origin) {
        for (int it = 0; it < max_steps; it++) {
            int r = epa_step(p, s, is3D, out, ctr);
            if (r > 0) return 1;
            if (r < 0) return 0;
        }
    } else if (!is3D && s->nv >= 2) {                                                    /* :590-605 (note swapped u/v names there) */
        float dist, u; int si, ei;
        closest_edge(s, &dist, &u, &si, &ei);
        v3 a = s->v[si].p, b = s->v[ei].p;
        v_st(out->pen_dir, v_add(a, v_mul(v_sub(b, a), u)));
        out->depth = dist;
        return 1;
    }
    return 0;
}

int npo_detect(const float* va, int na, const float xa[16], const float* vb, int nb, const float xb[16],
               int is3D, npo_contact* out, npo_counters* ctr)                            /* :694-732 */
{
    cparams p = { va, na, xa, vb, nb, xb };
    simplex s; memset(&s, 0, sizeof s);
    s.nv = 1;
    s.v[0].A = support(va, na, V(1, 1, 1), xa);
    s.v[0].B = support(vb, nb, V(-1, -1, -1), xb);
    s.v[0].p = v_sub(s.v[0].B, s.v[0].A);                                                /* :704  B - A */
    if (ctr) ctr->gjk_supports++;
    int iter = 0, result = RES_CONTINUE;
    while (result == RES_CONTINUE && iter++ < 20) { result = gjk_step(&p, is3D, &s, ctr); if (ctr) ctr->gjk_steps++; }
    if (result == RES_OVERLAP && is3D) {                                                 /* simplex.h:29-36 */
        sx_add_face(&s, 0, 1, 2, ctr); sx_add_face(&s, 0, 2, 3, ctr); sx_add_face(&s, 3, 1, 0, ctr); sx_add_face(&s, 1, 3, 2, ctr);
    }
    if (out) {
        memset(out, 0, sizeof *out);             /* normals/penDir zero (vector3()), depth pinned to 0, success=false */
        out->success = epa(&p, &s, 20, is3D, out, ctr);
    }
    free(s.f);
    return result == RES_OVERLAP;
}

/* ============================================================================================
 * physics.cpp: bodies, integrate, all-pairs with break, impulse response
 * ============================================================================================ */
typedef struct {
    v3 gravity, init_pos, init_rot;
    float mass, elasticity, sfric;
    m3 inv_inertia;
    int response;
    int shape;
    v3 pos, rot, L, H;
    /* adapter extension (no counterpart in engine/): a one-step external force and an enable flag, see npo_world_add_force */
    v3 force; int disabled;
} body;
typedef struct { float* xyz; int n; } shape;
struct npo_world {
    body* b; int nb, bcap;
    shape* s; int ns, scap;
    long long pair_tests, hits;
    npo_counters sum;
    int friction;        /* npo_world_set_friction: run the block the reference keeps commented out at physics.cpp:49-72 */
};
npo_world* npo_world_create(void) { return (npo_world*)calloc(1, sizeof(npo_world)); }
void npo_world_destroy(npo_world* w)
{
    if (!w) return;
    for (int i = 0; i < w->ns; i++) free(w->s[i].xyz);
    free(w->s); free(w->b); free(w);
}
int npo_world_add_shape(npo_world* w, const float* xyz, int n)
{
    if (w->ns == w->scap) { w->scap = w->scap ? 2 * w->scap : 8; w->s = (shape*)realloc(w->s, sizeof(shape) * (size_t)w->scap); }
    shape* s = &w->s[w->ns];
    s->n = n; s->xyz = (float*)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    memcpy(s->xyz, xyz, sizeof(float) * 3 * (size_t)n);
    return w->ns++;
}
int npo_world_add_body(npo_world* w, int shp, const float* sd, int response)             /* physics.cpp:11-18 */
{
    if (w->nb == w->bcap) { w->bcap = w->bcap ? 2 * w->bcap : 64; w->b = (body*)realloc(w->b, sizeof(body) * (size_t)w->bcap); }
    body* b = &w->b[w->nb];
    memset(b, 0, sizeof *b);
    b->gravity = v_ld(sd); b->init_pos = v_ld(sd + 3); b->init_rot = v_ld(sd + 6);
    b->mass = sd[9]; b->elasticity = sd[10]; b->sfric = sd[11];
    memcpy(b->inv_inertia.m, sd + 23, 36);
    b->response = response; b->shape = shp;
    b->pos = b->init_pos; b->rot = b->init_rot;       /* momenta start at vector3() == 0 */
    return w->nb++;
}
int npo_world_count(const npo_world* w) { return w->nb; }
/* Adapter extension for the ODE-shaped World facade; the reference engine has gravity only.  The force acts during
 * the next step only, as L += F*dt right after the gravity term of Physics::Update; a disabled body is not integrated. */
void npo_world_add_force(npo_world* w, int i, const float* f) { if (i >= 0 && i < w->nb) w->b[i].force = v_add(w->b[i].force, v_ld(f)); }
void npo_world_set_enabled(npo_world* w, int i, int enabled) { if (i >= 0 && i < w->nb) w->b[i].disabled = !enabled; }
void npo_world_get_state(const npo_world* w, float* o)
{ for (int i = 0; i < w->nb; i++) { v_st(o + 12 * i, w->b[i].pos); v_st(o + 12 * i + 3, w->b[i].rot); v_st(o + 12 * i + 6, w->b[i].L); v_st(o + 12 * i + 9, w->b[i].H); } }
void npo_world_set_state(npo_world* w, const float* in)
{ for (int i = 0; i < w->nb; i++) { w->b[i].pos = v_ld(in + 12 * i); w->b[i].rot = v_ld(in + 12 * i + 3); w->b[i].L = v_ld(in + 12 * i + 6); w->b[i].H = v_ld(in + 12 * i + 9); } }
void npo_world_counters(const npo_world* w, long long* pt, long long* hits, npo_counters* sum)
{ if (pt) *pt = w->pair_tests; if (hits) *hits = w->hits; if (sum) *sum = w->sum; }

static void body_update(body* b, float dt)                                               /* physics.cpp:175-194 */
{
    if (b->disabled) return;                                                             /* extension: dBodyIsEnabled == 0 */
    if (!v_isnone(b->gravity)) b->L = v_add(b->L, v_mul(v_mul(b->gravity, b->mass), dt));
    if (b->force.x != 0.0f || b->force.y != 0.0f || b->force.z != 0.0f) {                /* extension: dBodyAddForce, consumed by this step */
        b->L = v_add(b->L, v_mul(b->force, dt));
        b->force.x = b->force.y = b->force.z = 0.0f;
    }
    if (!feq(b->mass, 0.0f, TOL_DEFAULT)) {
        v3 v = v_mul(b->L, 1.0f / b->mass);
        v3 w = m3_mul(&b->inv_inertia, b->H);
        b->pos = v_add(b->pos, v_mul(v, dt));
        b->rot = v_add(b->rot, v_mul(w, dt));
    }
}
static void body_impulse(body* b, const npo_contact* c, v3 n, int friction)              /* physics.cpp:27-46 */
{
    v3 r = v_ld(c->pen_dir);
    v3 v = v_mul(b->L, 1.0f / b->mass);
    v3 w = m3_mul(&b->inv_inertia, b->H);
    v3 vp = v_add(v, v_cross(w, r));
    float i = -(1.0f + b->elasticity) * v_dot(vp, n);
    float k = (1.0f / b->mass) + v_dot(v_cross(m3_mul(&b->inv_inertia, v_cross(r, n)), r), n);
    float j = i / k;
    b->L = v_add(b->L, v_mul(n, j));
    b->H = v_add(b->H, v_mul(v_cross(r, n), j));
    float adjust = c->depth * 1.01f;
    b->pos = v_add(b->pos, v_mul(n, adjust));
    /* physics.cpp:49-72, commented out in the reference; restated as written there.  Pinned against the reference compiled with that
     * block uncommented (oracle/Makefile: _ref/libnetphys_ref_friction.so; tests/golden/friction.npz).  Notes: `w` is the angular
     * velocity from BEFORE the impulse; `vt.normalize();` (:57) discards its result (vector.h:211-219 is const and returns a copy), so
     * vt keeps its length; clamp(x,a,b) is the macro max(a, min(x, b)) of lib.h:13-23. */
    if (friction && b->sfric) {
        v = v_mul(b->L, 1.0f / b->mass);
        vp = v_add(v, v_cross(w, r));
        float rel = v_dot(vp, n);
        if (!feq(rel, 0.0f, TOL_DEFAULT)) {
            v3 vt = v_sub(vp, v_mul(n, rel));
            float it = v_dot(vp, vt);
            float kt = (1.0f / b->mass) + v_dot(v_cross(m3_mul(&b->inv_inertia, v_cross(r, vt)), r), vt);
            float fc = b->sfric * j;
            float x = -it / kt, lo = -fc;
            float mn = (x < fc) ? x : fc;
            float jt = (lo > mn) ? lo : mn;
            b->L = v_add(b->L, v_mul(vt, jt));
            b->H = v_add(b->H, v_mul(v_cross(r, vt), jt));
        }
    }
}
void npo_world_set_friction(npo_world* w, int on) { w->friction = on; }
/* Physics::ApplyImpulseResponse on one body with a caller's contact and normal (the world step only ever passes the zero normals) */
void npo_world_apply_impulse(npo_world* w, int i, const float* pen_dir, float depth, const float* n)
{
    npo_contact c; memset(&c, 0, sizeof c);
    memcpy(c.pen_dir, pen_dir, 12); c.depth = depth;
    if (i >= 0 && i < w->nb) body_impulse(&w->b[i], &c, v_ld(n), w->friction);
}
static void acc(npo_counters* s, const npo_counters* c)
{
    s->gjk_steps += c->gjk_steps; s->gjk_supports += c->gjk_supports; s->solve_line += c->solve_line; s->solve_tri += c->solve_tri;
    s->solve_tet += c->solve_tet; s->dir1 += c->dir1; s->dir2 += c->dir2; s->dir3 += c->dir3; s->epa_iters += c->epa_iters;
    s->epa_supports += c->epa_supports; s->epa_tri_tests += c->epa_tri_tests; s->epa_faces_added += c->epa_faces_added;
    if (c->epa_max_faces > s->epa_max_faces) s->epa_max_faces = c->epa_max_faces;
    if (c->epa_max_verts > s->epa_max_verts) s->epa_max_verts = c->epa_max_verts;
}
int npo_world_step(npo_world* w, float dt, int* oa, int* ob, float* odata, int cap)      /* physics.cpp:148-163 */
{
    int n = w->nb, nc = 0;
    for (int i = 0; i < n; i++) body_update(&w->b[i], dt);
    /* GetCollisions (:116-145): transforms are rebuilt per pair there; they only depend on the body */
    float* xf = (float*)malloc(sizeof(float) * 16 * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { float p[3], r[3]; v_st(p, w->b[i].pos); v_st(r, w->b[i].rot); npo_transform(p, r, xf + 16 * i); }
    int* ca = (int*)malloc(sizeof(int) * (size_t)(n + 1)); int* cb = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    npo_contact* cd = (npo_contact*)malloc(sizeof(npo_contact) * (size_t)(n + 1));
    for (int i = 0; i < n; i++) {
        const shape* sa = &w->s[w->b[i].shape];
        for (int j = i + 1; j < n; j++) {
            const shape* sb = &w->s[w->b[j].shape];
            npo_contact c; npo_counters ctr; memset(&ctr, 0, sizeof ctr);
            int hit = npo_detect(sa->xyz, sa->n, xf + 16 * i, sb->xyz, sb->n, xf + 16 * j, 1, &c, &ctr);
            w->pair_tests++; acc(&w->sum, &ctr);
            if (hit) { w->hits++; ca[nc] = i; cb[nc] = j; cd[nc] = c; nc++; break; }     /* :137-141 first hit, break */
        }
    }
    for (int k = 0; k < nc; k++) {                                                       /* OnCollision :93-114 */
        body* a = &w->b[ca[k]]; body* b = &w->b[cb[k]];
        if (a->response == 1) body_impulse(a, &cd[k], v_ld(cd[k].b_normal), w->friction);
        if (b->response == 1) body_impulse(b, &cd[k], v_ld(cd[k].a_normal), w->friction);
        if (k < cap && oa) {
            oa[k] = ca[k]; ob[k] = cb[k];
            float* o = odata + 11 * k;
            memcpy(o, cd[k].a_normal, 12); memcpy(o + 3, cd[k].b_normal, 12); memcpy(o + 6, cd[k].pen_dir, 12);
            o[9] = cd[k].depth; o[10] = cd[k].success ? 1.0f : 0.0f;
        }
    }
    free(xf); free(ca); free(cb); free(cd);
    return nc;
}

/* ============================================================================================
 * candidate-pair mode (CPU baseline for pair-tests/s): DetectCollision over an explicit list
 * ============================================================================================ */
typedef struct {
    int lo, hi;
    const float* shape_xyz; const int* off; const int* cnt;
    const int* sa; const float* pa; const int* sb; const float* pb;
    int* hit; float* out; npo_counters ctr;
} batch_job;
static void* batch_run(void* arg)
{
    batch_job* j = (batch_job*)arg;
    memset(&j->ctr, 0, sizeof j->ctr);
    for (int i = j->lo; i < j->hi; i++) {
        float xa[16], xb[16];
        npo_transform(j->pa + 6 * i, j->pa + 6 * i + 3, xa);
        npo_transform(j->pb + 6 * i, j->pb + 6 * i + 3, xb);
        npo_contact c; npo_counters ctr; memset(&ctr, 0, sizeof ctr);
        j->hit[i] = npo_detect(j->shape_xyz + 3 * j->off[j->sa[i]], j->cnt[j->sa[i]], xa,
                               j->shape_xyz + 3 * j->off[j->sb[i]], j->cnt[j->sb[i]], xb, 1, &c, &ctr);
        acc(&j->ctr, &ctr);
        memcpy(j->out + 5 * i, c.pen_dir, 12); j->out[5 * i + 3] = c.depth; j->out[5 * i + 4] = c.success ? 1.0f : 0.0f;
    }
    return 0;
}
void npo_detect_batch(int n, const float* shape_xyz, const int* shape_off, const int* shape_cnt,
                      const int* sa, const float* posea, const int* sb, const float* poseb,
                      int* hit, float* out, npo_counters* ctr_sum, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    batch_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < nthreads; t++) {
        batch_job* j = &jobs[t];
        j->lo = (int)((long long)n * t / nthreads); j->hi = (int)((long long)n * (t + 1) / nthreads);
        j->shape_xyz = shape_xyz; j->off = shape_off; j->cnt = shape_cnt;
        j->sa = sa; j->pa = posea; j->sb = sb; j->pb = poseb; j->hit = hit; j->out = out;
        if (nthreads == 1) batch_run(j); else pthread_create(&th[t], 0, batch_run, j);
    }
    if (ctr_sum) memset(ctr_sum, 0, sizeof *ctr_sum);
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], 0);
        if (ctr_sum) acc(ctr_sum, &jobs[t].ctr);
    }
}

/* ============================================================================================
 * exhaustive sin/cos checks.  Inputs are enumerated by their 32-bit pattern.
 *   npo_sincos_vs_libm  : mismatches of npo_sinf/npo_cosf against THIS box's libm sinf/cosf
 *   npo_sincos_hash     : order-independent weighted checksum of a block, compared with the same
 *                         checksum computed on the GPU (tests/test_gpu_sincos.py)
 * ============================================================================================ */
static inline uint32_t canon(float f) { uint32_t u = f2u(f); return ((u & 0x7fffffffu) > 0x7f800000u) ? 0x7fc00000u : u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
long long npo_sincos_vs_libm(unsigned start, unsigned long long count, unsigned* first_bad)
{
    long long bad = 0;
    for (unsigned long long k = 0; k < count; k++) {
        uint32_t u = (uint32_t)(start + k);
        float x = u2f(u);
        if (canon(npo_sinf(x)) != canon(sinf(x)) || canon(npo_cosf(x)) != canon(cosf(x))) { if (!bad && first_bad) *first_bad = u; bad++; }
    }
    return bad;
}
void npo_sincos_hash(unsigned start, unsigned long long count, unsigned long long* hs, unsigned long long* hc)
{
    uint64_t s = 0, c = 0;
    for (unsigned long long k = 0; k < count; k++) {
        uint32_t u = (uint32_t)(start + k);
        float x = u2f(u);
        uint64_t wgt = 2ull * u + 1ull;
        s += wgt * (uint64_t)canon(npo_sinf(x));
        c += wgt * (uint64_t)canon(npo_cosf(x));
    }
    *hs = s; *hc = c;
}
