// netphys_b200/csrc/np_math.cuh -- bit-exact scalar building blocks of the netphys hot path.
//
// Everything here is IEEE binary32 evaluated in the operation order of the reference engine
// (dougfrazer/netphys engine/vector.h, lib.h, matrix.h, util.h).  The translation unit MUST be compiled
// with  -fmad=false -prec-div=true -prec-sqrt=true -ftz=false  (see __graft_entry__.build): with those
// flags nvcc emits one correctly rounded FADD/FMUL/div/sqrt per C operator, which is what SSE2 code
// generated from the reference does.  The only double arithmetic is np_sincosf (libm-compatible).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define NP_HD __host__ __device__ __forceinline__
#define NP_D __device__ __forceinline__

namespace np {

struct V3 { float x, y, z; };

NP_HD V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
NP_HD V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
NP_HD V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
NP_HD V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
NP_HD V3 mul(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
NP_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                       // vector.h dot: left to right
NP_HD V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
NP_HD float magsq(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
NP_HD float mag(V3 a) { return sqrtf(magsq(a)); }
NP_HD V3 normalize(V3 a) { float m = mag(a); return mk(a.x / m, a.y / m, a.z / m); }             // three divisions, never rsqrt

// lib.h:13-15,32: max() there is a MACRO -- (a > b) ? a : b -- whose NaN behaviour differs from fmaxf.
NP_HD float macro_max(float a, float b) { return (a > b) ? a : b; }
NP_HD bool feq(float a, float b, float tol) { return fabsf(a - b) < (tol * macro_max(1.0f, macro_max(fabsf(a), fabsf(b)))); }
#define NP_TOL 0.00001f                                                                            // KINDA_CLOSE_ENOUGH, lib.h:29
NP_HD bool feq(float a, float b) { return feq(a, b, NP_TOL); }
NP_HD bool fle(float a, float b) { return a < b || feq(a, b); }                                   // lib.h:33 (tolerance argument ignored there)
NP_HD bool fge(float a, float b) { return a > b || feq(a, b); }                                   // lib.h:34
NP_HD bool veq(V3 a, V3 b, float tol) { return feq(a.x, b.x, tol) && feq(a.y, b.y, tol) && feq(a.z, b.z, tol); }
NP_HD bool veq(V3 a, V3 b) { return veq(a, b, NP_TOL); }                                          // vector3::operator==
NP_HD bool is_none(V3 a) { return feq(a.x, 0.0f) && feq(a.y, 0.0f) && feq(a.z, 0.0f); }           // vector3::IsNone

// matrix3 * vector3, matrix.h:59-66.  m = x1 x2 x3 y1 y2 y3 z1 z2 z3
NP_HD V3 m3mul(const float* m, V3 b)
{
    return mk(m[0] * b.x + m[1] * b.y + m[2] * b.z, m[3] * b.x + m[4] * b.y + m[5] * b.z, m[6] * b.x + m[7] * b.y + m[8] * b.z);
}

// ---------------------------------------------------------------------------------------------------
// sinf / cosf, bit-compatible with the libm the reference's sin(float)/cos(float) resolve to on the
// oracle platform (glibc 2.39 x86-64, FMA variant): double-precision minimax polynomial on a quadrant-
// reduced argument, fused multiply-adds where the -mfma build fuses them.  Same published algorithm as
// glibc sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,sincosf.h}; verified over all 2^32 inputs
// (tests/test_gpu_sincos.py against tests/golden/sincos_hash.npz).
// ---------------------------------------------------------------------------------------------------
NP_HD uint32_t f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
NP_HD uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ff; }

// odd n -> cosine polynomial, even n -> sine polynomial; one rounding to float at the end.
// glibc's second table entry negates every cosine coefficient (quadrants 2,3); round-to-nearest is
// sign-symmetric, so that equals negating the result of the positive polynomial (`flip`).
NP_HD float sc_poly(double x, double x2, int flip, int n)
{
    if ((n & 1) == 0) {
        const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
        double x3 = x * x2;
        double s1 = fma(x2, S3, S2);
        double x7 = x3 * x2;
        double s = fma(x3, S1, x);
        return (float)fma(x7, s1, s);
    } else {
        const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
        double x4 = x2 * x2;
        double c2 = fma(x2, C4, C3);
        double c1 = fma(x2, C1, C0);
        double x6 = x4 * x2;
        double c = fma(x4, C2, c1);
        float r = (float)fma(x6, c2, c);
        return flip ? -r : r;
    }
}

// 4/pi to 192 bits in overlapping 32-bit windows (argument reduction for |x| >= 120)
#ifdef __CUDACC__
__device__ __constant__ uint32_t NP_INV_PIO4_DEV[24] = {
    0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041,
};
#endif
static const uint32_t NP_INV_PIO4_HOST[24] = {
    0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041,
};

// quadrant reduction shared by sin and cos: reduced x (|x| <= pi/4), quadrant n, and the input's sign bit
// (only the large-argument path folds the sign in, as glibc does)
NP_HD double sc_reduce(float y, int* quadrant, int* sign)
{
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0, PI63 = 0x1.921FB54442D18p-62;
    double x = (double)y;
    if (abstop12(y) < abstop12(120.0f)) {
        double r = x * HPI_INV;
        int n = ((int32_t)r + 0x800000) >> 24;
        *quadrant = n; *sign = 0;
        return fma(-(double)n, HPI, x);
    }
    uint32_t xi = f2u(y);
    *sign = (int)(xi >> 31);
#ifdef __CUDA_ARCH__
    const uint32_t* arr = &NP_INV_PIO4_DEV[(xi >> 26) & 15];
#else
    const uint32_t* arr = &NP_INV_PIO4_HOST[(xi >> 26) & 15];
#endif
    int shift = (xi >> 23) & 7;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    uint64_t res0 = (uint64_t)(uint32_t)(xi * arr[0]);
    uint64_t res1 = (uint64_t)xi * arr[4];
    uint64_t res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    uint64_t n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    *quadrant = (int)n;
    return (double)(int64_t)res0 * PI63;
}

// one call yields both (the reference evaluates sin and cos of each Euler angle, matrix.h:283-288)
NP_HD void np_sincosf(float y, float* sn, float* cs)
{
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        double x = (double)y, x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) { *sn = y; *cs = 1.0f; return; }
        *sn = sc_poly(x, x2, 0, 0);
        *cs = sc_poly(x, x2, 0, 1);
        return;
    }
    if (abstop12(y) < abstop12(INFINITY)) {
        int n, sign;
        double x = sc_reduce(y, &n, &sign);
        int q = n + sign;
        double s = ((q + 1) & 2) ? -1.0 : 1.0;          // sign[] = { 1, -1, -1, 1 } indexed by q & 3
        int flip = (q & 2) ? 1 : 0;
        double xs = x * s, x2 = x * x;
        *sn = sc_poly(xs, x2, flip, n);
        *cs = sc_poly(xs, x2, flip, n ^ 1);
        return;
    }
    float nan = (y - y) / (y - y);
    *sn = nan; *cs = nan;
}

// ---------------------------------------------------------------------------------------------------
// Physics::GetTransform (physics.h:70-76): identity -> translate (matrix.h:297-303) -> rotate (matrix.h:283-288).
// R = x1 x2 x3 / y1 y2 y3 / z1 z2 z3, t = x4 y4 z4.  Products associate left to right.
// ---------------------------------------------------------------------------------------------------
NP_HD void make_transform(V3 pos, V3 rot, float* R, V3* t)
{
    float sx, cx, sy, cy, sz, cz;
    np_sincosf(rot.x, &sx, &cx); np_sincosf(rot.y, &sy, &cy); np_sincosf(rot.z, &sz, &cz);
    *t = mk(0.0f + pos.x, 0.0f + pos.y, 0.0f + pos.z);
    R[0] = cy * cz;  R[1] = sx * sy * cz - cx * sz;  R[2] = cx * sy * cz + sx * sz;
    R[3] = cy * sz;  R[4] = sx * sy * sz + cx * cz;  R[5] = cx * sy * sz - sx * cz;
    R[6] = -sy;      R[7] = sx * cy;                 R[8] = cx * cy;
}
// matrix4 * vector3 with w = 1 (matrix.h:145-157): x1*px + x2*py + x3*pz + x4*1.0f ; (x4*1.0f == x4 exactly)
NP_HD V3 xf_point(const float* R, V3 t, V3 p)
{
    return mk(R[0] * p.x + R[1] * p.y + R[2] * p.z + t.x, R[3] * p.x + R[4] * p.y + R[5] * p.z + t.y, R[6] * p.x + R[7] * p.y + R[8] * p.z + t.z);
}

// ---------------------------------------------------------------------------------------------------
// util.h closest-point-to-origin solvers.  The query point is the origin in every call on the path
// (collision_detection.cpp:31,62,129,275,309) but the reference still evaluates the p-terms (0*x sums,
// 0 - a); they only affect the sign of zeros, and are reproduced literally through `o` = (0,0,0).
// ---------------------------------------------------------------------------------------------------
enum { LINE_AB = 0, LINE_A, LINE_B };
enum { TRI_ABC = 0, TRI_A, TRI_B, TRI_C, TRI_AB, TRI_AC, TRI_BC };
enum { TET_NONE = -1, TET_ABCD = 0, TET_A, TET_B, TET_C, TET_D, TET_AB, TET_AC, TET_AD, TET_BC, TET_BD, TET_CD, TET_ABC, TET_ABD, TET_ACD, TET_BCD };

NP_HD int closest_line(V3 a, V3 b, float* pu)                                                     // util.h:18-37
{
    const V3 o = mk(0.0f, 0.0f, 0.0f);
    V3 ap = sub(o, a), ab = sub(b, a);
    float u = dot(ap, ab) / dot(ab, ab);
    int r = LINE_AB;
    if (u <= 0.0f) { u = 0.0f; r = LINE_A; }
    else if (u >= 1.0f) { u = 1.0f; r = LINE_B; }
    *pu = u;
    return r;
}
NP_HD int closest_triangle(V3 va, V3 vb, V3 vc, float* pu, float* pv)                             // util.h:144-213
{
    const V3 o = mk(0.0f, 0.0f, 0.0f);
    V3 A = va, B = sub(vb, va), C = sub(vc, va);
    float a = dot(B, B);
    float b = dot(C, C);
    float c = 2 * (dot(B, C));
    float d = 2 * (dot(A, B) - dot(o, B));
    float e = 2 * (dot(A, C) - dot(o, C));
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
// region only (the GJK never uses the ratios of the tetrahedron solve, collision_detection.cpp:121-259)
NP_HD int closest_tetra(V3 va, V3 vb, V3 vc, V3 vd, float* pu, float* pv, float* pw)              // util.h:312-457
{
    const V3 P = mk(0.0f, 0.0f, 0.0f);
    V3 A = va, B = sub(vb, va), C = sub(vc, va), D = sub(vd, va);
    float a = dot(B, B), b = dot(C, C), c = dot(D, D);
    float d = dot(B, C), e = dot(B, D), f = dot(C, D);
    float g = dot(A, B) - dot(P, B);
    float h = dot(A, C) - dot(P, C);
    float i = dot(A, D) - dot(P, D);
    float denom = (a * b * c - a * f * f - b * e * e - c * d * d + 2 * d * e * f);
    if (feq(denom, 0.f)) return TET_NONE;
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

// GetDirectionToOrigin, util.h:461-509
NP_HD V3 dir_fallback(V3 ab)
{
    V3 abt = cross(ab, mk(0.0f, 0.0f, 1.0f));
    if (magsq(abt) == 0.0f) return cross(ab, mk(0.0f, 1.0f, 0.0f));
    return abt;
}
NP_HD V3 dir1(V3 a) { return neg(a); }
NP_HD V3 dir2(V3 a, V3 b)
{
    V3 ab = sub(b, a), ao = neg(b);                                                              // note ao = -b, util.h:469
    V3 t = cross(ab, ao);
    V3 bt = cross(t, ab);
    float sign = dot(bt, ao);
    if (feq(sign, 0.0f)) return dir_fallback(ab);
    return (sign < 0.0f) ? neg(bt) : bt;
}
NP_HD V3 dir3(V3 a, V3 b, V3 c)
{
    V3 ab = sub(b, a), ac = sub(c, a), ao = neg(a);
    V3 n = cross(ab, ac);
    float sign = dot(n, ao);
    if (feq(sign, 0.0f)) return dir_fallback(ab);
    return (sign < 0.0f) ? neg(n) : n;
}

}  // namespace np
