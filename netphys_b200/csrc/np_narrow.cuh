// netphys_b200/csrc/np_narrow.cuh -- GJK and EPA as group-cooperative per-pair state machines.
//
// Replaces DetectCollision / DetectCollisionStep / FindIntersectionPoints[Step] and the simplex store
// (reference engine/collision_detection.cpp:18-732, engine/simplex.h:5-76).
//
// B200 shape of the algorithm
//   * A pair is owned by a GROUP of G lanes (G = 1, 4, 8 or 32; 32/G pairs per warp).  The hot loop of the
//     reference -- the support mapping, 27 flop per vertex (physics_util.cpp:5-35) -- is strided over the
//     group's lanes and finished with a shuffle arg-max whose tie-break (lower vertex index wins) reproduces
//     the reference's strict '>' / first-maximum-wins scan bit for bit.  Small G maximises throughput when
//     there are millions of pairs; G = 32 minimises latency when a step has only thousands (BASELINE
//     configs[1]) and for 162-vertex shapes.
//   * Everything that is not the support loop (simplex solve, search direction, termination tests) is
//     evaluated redundantly by every lane of the group from identical inputs, so the lanes of a group never
//     diverge and no communication is needed: the GJK simplex (<= 4 points x {p,A,B}) lives in registers.
//   * GJK and EPA are separate kernels.  GJK decides hit/miss (and, in a world step, drives the ordered
//     first-hit search); only hits reach EPA, carrying their final simplex (24 floats).  Each kernel is
//     small enough for the instruction cache and GJK needs no local or shared memory at all.
//   * The EPA polytope (<= 24 vertices, <= FMAX faces) is ONE copy per group: shared memory for G >= 4,
//     per-thread local memory for G = 1.  The face scan (closest-point-on-triangle per face) and AddFace are
//     strided over the lanes; the order-dependent horizon extraction is done by the group's lane 0.
//     A polytope that outgrows FMAX is never truncated: the pair is queued for the large-polytope launch
//     (same code, G = 32, global-memory arena).
#pragma once
#include "np_math.cuh"

namespace np {

#define NP_EPA_MAX_VERTS 24        // 4 GJK points + at most one per EPA iteration (20)
#define NP_GJK_MAX_STEPS 20        // collision_detection.cpp:707
#define NP_EPA_MAX_STEPS 20        // collision_detection.cpp:728 (max_steps = maxIterations)
#define NP_EPA_TOL 0.01f           // collision_detection.cpp:423

enum PairFlags { NP_PAIR_HIT = 1, NP_PAIR_SUCCESS = 2, NP_PAIR_OVERFLOW = 4, NP_PAIR_SEED_ALIVE = 8 };
// World mode: hit_j[i] = -1 (no collision) or the partner's index, with CollisionData::success in bit 30 -- the per-body collision
// record is then hit_j + contact (20 bytes) and the flags array never has to leave the process that searched body i.
#define NP_HITJ_SUCCESS 0x40000000
__host__ __device__ __forceinline__ int hitj_body(int v) { return v < 0 ? v : (v & ~NP_HITJ_SUCCESS); }

#define NP_TAB_NONE 0xffffffffu
struct PairGeom {
    float Ra[9], Rb[9];
    V3 ta, tb;
    const float4* va; const float4* vb;   // shape-local vertices (x,y,z,-)
    int na, nb;
    unsigned taba = NP_TAB_NONE, tabb = NP_TAB_NONE;   // support-direction table of each side (tab_arm), NP_TAB_NONE = scan every vertex
    const uint2* cells = nullptr;
};

struct PairStats { unsigned gjk_steps, support_pairs, support_verts, epa_iters, epa_tri_tests; };

// ---------------------------------------------------------------------------------------------------
// group primitives
// ---------------------------------------------------------------------------------------------------
template <int G> NP_D int glane() { return (int)(threadIdx.x & (G - 1)); }
template <int G> NP_D unsigned gmask()
{
    if (G == 32) return 0xffffffffu;
    return ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(G - 1));
}
template <int G> NP_D void gsync() { if (G > 1) __syncwarp(gmask<G>()); }

// PhysUtil_GetPointFurthestInDirection, physics_util.cpp:5-35: world*p per vertex, strict '>' from -inf, first
// maximum wins, result starts at (0,0,0).  Lane l scans vertices l, l+G, ...; the cross-lane reduction prefers the
// larger dot product and, on equal dot products, the lower vertex index.
NP_D int float_order(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }      // monotone float -> int
NP_D float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---------------------------------------------------------------------------------------------------
// Packed binary32 pairs (sm_100 FADD2 / FFMA2).  mul.rn.f32x2 and add.rn.f32x2 round each half exactly like the scalar FMUL /
// FADD (the product is taken as x*y + (-0): the correctly rounded product, signed zeros included), so two vertices go through the reference's transform-and-dot sequence (physics_util.cpp:17-21)
// per instruction with the same bits as one.  Measured on B200: 51 Tflop/s of non-fused FADD/FMUL work against 35 scalar
// (tools/mb/fp32x2_peak.cu) -- and half the issue slots, which is what the thread-per-pair kernels are short of.
// ---------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
NP_D f32x2 pk2(float lo, float hi) { f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
NP_D void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// ptxas CONTRACTS mul.rn.f32x2 + add.rn.f32x2 into a fused FFMA2 whatever -fmad says (checked in SASS; so it does for an fma whose
// addend is a literal -0), which would break parity.  The product is therefore an FFMA2 whose -0 addend comes from constant memory:
// ptxas cannot fold it, and an fma cannot be contracted into the add that follows.
__constant__ unsigned long long c_negz2 = 0x8000000080000000ull;
NP_D f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c_negz2)); return d; }
NP_D f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// transform + direction of one support call, every scalar duplicated into both halves
struct SupPack {
    f32x2 R[9], tx, ty, tz, dx, dy, dz;
    NP_D void set(const float* r, V3 t, V3 d)
    {
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = pk2(r[k], r[k]);
        tx = pk2(t.x, t.x); ty = pk2(t.y, t.y); tz = pk2(t.z, t.z); dx = pk2(d.x, d.x); dy = pk2(d.y, d.y); dz = pk2(d.z, d.z);
    }
    // dot(d, world * p) of two shape-local vertices at once: xf_point then dot, the same operation order per half
    NP_D void dots(float4 q0, float4 q1, float& dp0, float& dp1) const
    {
        const f32x2 X = pk2(q0.x, q1.x), Y = pk2(q0.y, q1.y), Z = pk2(q0.z, q1.z);
        const f32x2 wx = add2(add2(add2(mul2(R[0], X), mul2(R[1], Y)), mul2(R[2], Z)), tx);
        const f32x2 wy = add2(add2(add2(mul2(R[3], X), mul2(R[4], Y)), mul2(R[5], Z)), ty);
        const f32x2 wz = add2(add2(add2(mul2(R[6], X), mul2(R[7], Y)), mul2(R[8], Z)), tz);
        upk2(add2(add2(mul2(dx, wx), mul2(dy, wy)), mul2(dz, wz)), dp0, dp1);
    }
};

#ifndef NP_PACKED
#define NP_PACKED 0      // measured: unfused FFMA2 + FADD2 run at 2 cycles each -- 35.1 Tflop/s, the scalar rate -- and the pack/unpack moves make the 1M world 4 % slower
#endif
template <int G, bool CACHED>
NP_D V3 support(const float4* __restrict__ v, int n, const float* R, V3 t, V3 d)
{
    // CACHED: v holds the body's WORLD-space vertices, written once per step by k_world_verts with the very expression
    // xf_point evaluates here, so both variants see the same bits.
    float bd = -INFINITY;
    int bi = 0x7fffffff;
    if (G == 1 && !CACHED && NP_PACKED) {
        // throughput regime, two vertices per instruction (SupPack); visited in index order, so strict '>' keeps the first maximum
        SupPack sp; sp.set(R, t, d);
#pragma unroll 2
        for (int k = 0; k < n; k += 2) {
            const float4 q0 = v[k], q1 = v[(k + 1 < n) ? k + 1 : k];
            float dp0, dp1; sp.dots(q0, q1, dp0, dp1);
            if (dp0 > bd) { bd = dp0; bi = k; }
            if (k + 1 < n && dp1 > bd) { bd = dp1; bi = k + 1; }
        }
        if (bi == 0x7fffffff) return mk(0.0f, 0.0f, 0.0f);      // no vertex, or every dot product NaN / -inf
        float4 q = v[bi];
        return xf_point(R, t, mk(q.x, q.y, q.z));
    }
    if (G == 1) {
        // throughput regime: track (best dot, index) only and recompute the winner once -- same operands, same bits
#pragma unroll 4
        for (int k = 0; k < n; k++) {
            float4 q = v[k];
            V3 wp = CACHED ? mk(q.x, q.y, q.z) : xf_point(R, t, mk(q.x, q.y, q.z));
            float dp = dot(d, wp);
            if (dp > bd) { bd = dp; bi = k; }
        }
        if (bi == 0x7fffffff) return mk(0.0f, 0.0f, 0.0f);      // no vertex, or every dot product NaN / -inf
        float4 q = v[bi];
        return CACHED ? mk(q.x, q.y, q.z) : xf_point(R, t, mk(q.x, q.y, q.z));
    }
    // latency regime: each lane tracks (best dot, index) only -- the loop is the longest dependent stretch of a trip (half of its cycles,
    // tools/gjk_cycles.py), so it carries as few instructions per vertex as possible.  The cross-lane arg-max is two warp reductions
    // (redux.sync) on an order-preserving integer image of the dot product (-0 folded into +0, as '>' does) and on the vertex index:
    // larger dot wins, lower index on ties == the reference's strict '>' scan.  The winner's position is then re-read (one broadcast
    // load) and, where the vertices are shape-local, transformed again: same operands, same bits.
#pragma unroll 4
    for (int k = glane<G>(); k < n; k += G) {
        float4 q = v[k];
        V3 wp = CACHED ? mk(q.x, q.y, q.z) : xf_point(R, t, mk(q.x, q.y, q.z));
        float dp = dot(d, wp);
        if (dp > bd) { bd = dp; bi = k; }
    }
    const unsigned m = gmask<G>();
    const int key = float_order(bd + 0.0f);
    const int kmax = __reduce_max_sync(m, key);
    const unsigned imin = __reduce_min_sync(m, (key == kmax) ? (unsigned)bi : 0xffffffffu);
    if (imin >= 0x7fffffffu) return mk(0.0f, 0.0f, 0.0f);
    const float4 q = v[imin];
    return CACHED ? mk(q.x, q.y, q.z) : xf_point(R, t, mk(q.x, q.y, q.z));
}

// ---------------------------------------------------------------------------------------------------
// Support-direction tables: the reference's vertex scan (physics_util.cpp:5-35) answered from a handful of candidates, same bits.
//
// The scan returns the FIRST vertex k whose binary32 value f_k = fl(d . fl(R p_k + t)) is the largest.  Write E_k for the same
// expression in real arithmetic.  Every f_k is within delta = 7.001 u |d|_1 (rho r + |t|_inf) of E_k (u = 2^-24: four roundings per
// transformed coordinate, three for the dot product; r = largest |p_k|, rho = largest row norm of R), so the winner -- and every vertex
// that ties with it -- satisfies E_k >= max_j E_j - 2 delta.  E_k - E_j = (R^T d) . (p_k - p_j) does not involve t: which vertices can
// win depends only on the DIRECTION of R^T d in shape space and on a margin mu >= 2 delta / |R^T d| (r/256).  np_suptab.h tabulates, per
// shape, for every cell of a cube map of directions the set { k : u . p_k >= max_j u . p_j - mu for some u in the (dilated) cell },
// at most 7 vertex indices in ascending order.  A support call transforms the direction (any rounding: it only picks the cell, and
// the cells are dilated by more than that error), reads one 8-byte cell and evaluates ITS candidates with the reference's own
// expression in index order under the same strict '>': the winner is the scan's winner, bit for bit, because the scan's winner is
// among the candidates and the candidates are visited in the scan's order.
//
// The margin is only valid while 2 delta / |R^T d| <= mu: tab_arm (once per armed body) requires R^T R = I within 2^-10 and
// |t|_inf + r within the reach stored with the shape, and a call whose direction is not finite or outside [1e-18, 1e18] -- or whose cell
// holds more than 7 candidates -- falls back to the scan.  161 of an icosphere's 162 vertices are never touched.
// ---------------------------------------------------------------------------------------------------
struct TabView { const unsigned* tab; const float* tlim; const uint2* cells; };   // per shape: cell offset | log2(N) << 28; reach; all cells
NP_D unsigned tab_arm(const TabView& tv, int shape, const float* R, V3 t, bool rot_known)
{
    if (!tv.tab) return NP_TAB_NONE;
    const unsigned e = tv.tab[shape];
    if (e == NP_TAB_NONE) return e;
    const float tol = 0.0009765625f;
    const float ax = fabsf(t.x), ay = fabsf(t.y), az = fabsf(t.z);
    bool ok = (ax + ay + az) <= 3.0e38f && fmaxf(ax, fmaxf(ay, az)) <= tv.tlim[shape];             // (the sum catches a NaN, which fmaxf would drop)
    // rot_known: R came out of make_transform (sines and cosines multiplied together): orthonormal to ~1e-6 unless an angle was not
    // finite, and then every entry that angle touches is a NaN -- the sum below finds it
    if (rot_known) return (ok && (R[0] + R[1] + R[2] + R[3] + R[4] + R[5] + R[6] + R[7] + R[8]) <= 9.0f) ? e : NP_TAB_NONE;
    ok = ok && fabsf(R[0] * R[0] + R[3] * R[3] + R[6] * R[6] - 1.0f) <= tol && fabsf(R[1] * R[1] + R[4] * R[4] + R[7] * R[7] - 1.0f) <= tol
            && fabsf(R[2] * R[2] + R[5] * R[5] + R[8] * R[8] - 1.0f) <= tol && fabsf(R[0] * R[1] + R[3] * R[4] + R[6] * R[7]) <= tol
            && fabsf(R[0] * R[2] + R[3] * R[5] + R[6] * R[8]) <= tol && fabsf(R[1] * R[2] + R[4] * R[5] + R[7] * R[8]) <= tol;
    return ok ? e : NP_TAB_NONE;
}
// true: *out is the support point (world space, the reference's bits).  false: the caller scans.
NP_D bool support_tab(const float4* __restrict__ v, const float* R, V3 t, V3 d, unsigned e, const uint2* __restrict__ cells, V3* out)
{
    const float ux = __fmaf_rn(R[6], d.z, __fmaf_rn(R[3], d.y, R[0] * d.x));                       // R^T d
    const float uy = __fmaf_rn(R[7], d.z, __fmaf_rn(R[4], d.y, R[1] * d.x));
    const float uz = __fmaf_rn(R[8], d.z, __fmaf_rn(R[5], d.y, R[2] * d.x));
    const float ax = fabsf(ux), ay = fabsf(uy), az = fabsf(uz);
    const float sum = ax + ay + az;
    if (!(sum >= 1e-18f && sum <= 1e18f)) return false;                                           // zero, denormal, huge, inf or NaN direction
    float major, a, b; int f;
    if (ax >= ay && ax >= az) { major = ux; a = uy; b = uz; f = 0; }
    else if (ay >= az) { major = uy; a = uz; b = ux; f = 2; }
    else { major = uz; a = ux; b = uy; f = 4; }
    f += (major < 0.0f) ? 1 : 0;
    const int lg = (int)(e >> 28), nm1 = (1 << lg) - 1;
    const float half_n = (float)(1 << lg) * 0.5f;
    const float sc = __fdividef(half_n, fabsf(major));
    const int ia = min(max((int)__fmaf_rn(a, sc, half_n), 0), nm1), ib = min(max((int)__fmaf_rn(b, sc, half_n), 0), nm1);
    const uint2 cell = __ldg(cells + (e & 0x0fffffffu) + ((((unsigned)f << lg) + (unsigned)ib) << lg) + (unsigned)ia);
    const int cnt = (int)(cell.y >> 24);
    if (cnt > 7) return false;
    unsigned long long bits = (unsigned long long)cell.x | ((unsigned long long)cell.y << 32);
    float bd = -INFINITY; V3 best = mk(0.0f, 0.0f, 0.0f); bool any = false;
    for (int c = 0; c < cnt; c++) {
        const float4 q = v[(int)(bits & 0xffu)]; bits >>= 8;
        const V3 wp = xf_point(R, t, mk(q.x, q.y, q.z));
        const float dp = dot(d, wp);
        if (dp > bd) { bd = dp; best = wp; any = true; }
    }
    if (!any) return false;
    *out = best;
    return true;
}
// The same lookup for a direction along a world axis, d = sg e_a (the faces of a body's box, k_world_verts): R^T d is sg x row a of R, and
// the score d . (R p + t) is sg x the world coordinate a itself (the other two terms are products with zero), so only that coordinate
// is evaluated -- with xf_point's expression -- and the extreme one among the cell's candidates is returned.  row = R[3a..3a+2], ta = t[a].
// (in two halves, so that a caller with several directions has all its cell loads in flight before it evaluates any)
NP_D const uint2* tab_axis_cell(V3 row, float sg, unsigned e, const uint2* __restrict__ cells)      // null: the caller scans
{
    const float ux = sg * row.x, uy = sg * row.y, uz = sg * row.z;
    const float ax = fabsf(ux), ay = fabsf(uy), az = fabsf(uz);
    const float sum = ax + ay + az;
    if (!(sum >= 1e-18f && sum <= 1e18f)) return nullptr;
    float major, a, b; int f;
    if (ax >= ay && ax >= az) { major = ux; a = uy; b = uz; f = 0; }
    else if (ay >= az) { major = uy; a = uz; b = ux; f = 2; }
    else { major = uz; a = ux; b = uy; f = 4; }
    f += (major < 0.0f) ? 1 : 0;
    const int lg = (int)(e >> 28), nm1 = (1 << lg) - 1;
    const float half_n = (float)(1 << lg) * 0.5f;
    const float sc = __fdividef(half_n, fabsf(major));
    const int ia = min(max((int)__fmaf_rn(a, sc, half_n), 0), nm1), ib = min(max((int)__fmaf_rn(b, sc, half_n), 0), nm1);
    return cells + (e & 0x0fffffffu) + ((((unsigned)f << lg) + (unsigned)ib) << lg) + (unsigned)ia;
}
NP_D bool tab_axis_eval(const float4* __restrict__ v, V3 row, float ta, float sg, uint2 cell, float* out)
{
    const int cnt = (int)(cell.y >> 24);
    if (cnt > 7) return false;
    unsigned long long bits = (unsigned long long)cell.x | ((unsigned long long)cell.y << 32);
    float bd = -INFINITY; bool any = false;
    for (int c = 0; c < cnt; c++) {
        const float4 q = v[(int)(bits & 0xffu)]; bits >>= 8;
        const float dp = sg * (row.x * q.x + row.y * q.y + row.z * q.z + ta);
        if (dp > bd) { bd = dp; any = true; }
    }
    if (!any) return false;
    *out = sg * bd;
    return true;
}
NP_D bool support_tab_axis(const float4* __restrict__ v, V3 row, float ta, float sg, unsigned e, const uint2* __restrict__ cells, float* out)
{
    const uint2* c = tab_axis_cell(row, sg, e, cells);
    return c && tab_axis_eval(v, row, ta, sg, __ldg(c), out);
}
// one side's support of a pair: from the table when the side is armed with one, else the scan
template <int G, bool CACHED>
NP_D V3 support_a(const PairGeom& g, V3 d)
{
    if (!CACHED && g.taba != NP_TAB_NONE && (G == 1 || g.na >= 12 * G)) { V3 r; if (support_tab(g.va, g.Ra, g.ta, d, g.taba, g.cells, &r)) return r; }   // (a group scans a small shape faster than it looks one up)
    return support<G, CACHED>(g.va, g.na, g.Ra, g.ta, d);
}
template <int G, bool CACHED>
NP_D V3 support_b(const PairGeom& g, V3 d)
{
    if (!CACHED && g.tabb != NP_TAB_NONE && (G == 1 || g.nb >= 12 * G)) { V3 r; if (support_tab(g.vb, g.Rb, g.tb, d, g.tabb, g.cells, &r)) return r; }
    return support<G, CACHED>(g.vb, g.nb, g.Rb, g.tb, d);
}

// Thread-per-pair kernels: the lanes of a warp hold different pairs, and a per-lane vertex loop runs as long as the LARGEST shape in
// the warp (a 162-vertex sphere next to 8-vertex boxes: 22 % of the lanes active, measured).  Lanes whose shape has at least
// NP_COOP_MIN vertices are therefore served one after the other by the whole warp -- their arguments broadcast, the vertices strided
// over 32 lanes, the same order-preserving arg-max as support<32> -- unless so many lanes are large that the plain loop is cheaper.
// Small shapes keep the per-lane loop.  Must be called by all 32 lanes; `want` says whether this lane needs a result.
#ifndef NP_COOP_MIN
#define NP_COOP_MIN 48
#endif
// G lanes per pair: `want` is uniform inside a group and the result goes to all of its lanes.  With a plain loop the warp pays
// n/G iterations whenever any group is large; served by the warp a large group costs n/32 iterations plus ~50 instructions of
// broadcast and reduction, so the warp serves them only while at most 20/G groups are large.  (Used with G = 1: in the EPA kernel,
// G = 4, the same change measured 2 % slower -- supports are not what its warps wait for.)
template <int G>
NP_D V3 support_coop(bool want, const float4* __restrict__ v, int n, const float* R, V3 t, V3 d)
{
    const unsigned full = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u);
    V3 res = mk(0.0f, 0.0f, 0.0f);
    bool served = false;
    unsigned big = __ballot_sync(full, want && glane<G>() == 0 && n >= NP_COOP_MIN);
    if (__popc(big) <= 20 / G) {
        while (big) {
            const int src = __ffs(big) - 1; big &= big - 1;
            const float4* sv = (const float4*)__shfl_sync(full, (unsigned long long)v, src);
            const int sn = __shfl_sync(full, n, src);
            float sR[9];
#pragma unroll
            for (int k = 0; k < 9; k++) sR[k] = __shfl_sync(full, R[k], src);
            const V3 st = mk(__shfl_sync(full, t.x, src), __shfl_sync(full, t.y, src), __shfl_sync(full, t.z, src));
            const V3 sd = mk(__shfl_sync(full, d.x, src), __shfl_sync(full, d.y, src), __shfl_sync(full, d.z, src));
            float bd = -INFINITY; int bi = 0x7fffffff;
#pragma unroll 2
            for (int k = lane; k < sn; k += 32) {
                float4 q = sv[k];
                V3 wp = xf_point(sR, st, mk(q.x, q.y, q.z));
                float dp = dot(sd, wp);
                if (dp > bd) { bd = dp; bi = k; }
            }
            const int key = float_order(bd + 0.0f);
            const int kmax = __reduce_max_sync(full, key);
            const unsigned imin = __reduce_min_sync(full, (key == kmax) ? (unsigned)bi : 0xffffffffu);
            if ((lane & ~(G - 1)) == src) {
                if (imin >= 0x7fffffffu) res = mk(0.0f, 0.0f, 0.0f);
                else { const float4 q = sv[imin]; res = xf_point(sR, st, mk(q.x, q.y, q.z)); }
                served = true;
            }
        }
    }
    if (want && !served) res = support<G, false>(v, n, R, t, d);
    return res;
}

// Thread-per-pair kernels, second generation (k_gjk1).  A large shape's support is served by the whole warp as above, with two changes:
//   * CBIG: bodies with at least NP_COOP_MIN vertices have their WORLD-space vertices cached once per step (k_world_verts, big bodies
//     only), so serving a lane costs a pointer, a count and a direction broadcast plus 6 instructions per vertex instead of the
//     transform's 21 flops -- the same world*p bits, computed once per body and step instead of once per support call;
//   * shapes of up to 192 vertices (the icosphere has 162) run the vertex loop fully unrolled with all six loads in flight.
template <bool CBIG>
NP_D V3 support_coop1(bool want, const float4* __restrict__ v, int n, const float* R, V3 t, V3 d)
{
    const unsigned full = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u);
    V3 res = mk(0.0f, 0.0f, 0.0f);
    bool served = false;
    const bool isbig = n >= NP_COOP_MIN;
    unsigned big = __ballot_sync(full, want && isbig);
    if (__popc(big) <= 20) {
        while (big) {
            const int src = __ffs(big) - 1; big &= big - 1;
            const float4* sv = (const float4*)__shfl_sync(full, (unsigned long long)v, src);
            const int sn = __shfl_sync(full, n, src);
            float sR[9]; V3 st = mk(0.0f, 0.0f, 0.0f);
            if (!CBIG) {
#pragma unroll
                for (int k = 0; k < 9; k++) sR[k] = __shfl_sync(full, R[k], src);
                st = mk(__shfl_sync(full, t.x, src), __shfl_sync(full, t.y, src), __shfl_sync(full, t.z, src));
            }
            const V3 sd = mk(__shfl_sync(full, d.x, src), __shfl_sync(full, d.y, src), __shfl_sync(full, d.z, src));
            float bd = -INFINITY; int bi = 0x7fffffff;
            if (!CBIG && NP_PACKED && sn <= 192) {
                SupPack sp; sp.set(sR, st, sd);
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int k0 = lane + 64 * u, k1 = k0 + 32;
                    if (k0 < sn) {                                                   // (warp-uniform for u <= 1 whenever sn >= 96)
                        const float4 q0 = sv[k0], q1 = sv[(k1 < sn) ? k1 : k0];
                        float dp0, dp1; sp.dots(q0, q1, dp0, dp1);
                        if (dp0 > bd) { bd = dp0; bi = k0; }
                        if (k1 < sn && dp1 > bd) { bd = dp1; bi = k1; }
                    }
                }
            } else if (sn <= 192) {
                float4 q[6];
#pragma unroll
                for (int u = 0; u < 6; u++) { const int k = lane + 32 * u; q[u] = (k < sn) ? sv[k] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                for (int u = 0; u < 6; u++) {
                    const int k = lane + 32 * u;
                    const V3 wp = CBIG ? mk(q[u].x, q[u].y, q[u].z) : xf_point(sR, st, mk(q[u].x, q[u].y, q[u].z));
                    const float dp = dot(sd, wp);
                    if (k < sn && dp > bd) { bd = dp; bi = k; }
                }
            } else {
#pragma unroll 2
                for (int k = lane; k < sn; k += 32) {
                    const float4 q = sv[k];
                    const V3 wp = CBIG ? mk(q.x, q.y, q.z) : xf_point(sR, st, mk(q.x, q.y, q.z));
                    const float dp = dot(sd, wp);
                    if (dp > bd) { bd = dp; bi = k; }
                }
            }
            const int key = float_order(bd + 0.0f);
            const int kmax = __reduce_max_sync(full, key);
            const unsigned imin = __reduce_min_sync(full, (key == kmax) ? (unsigned)bi : 0xffffffffu);
            if (lane == src) {
                if (imin >= 0x7fffffffu) res = mk(0.0f, 0.0f, 0.0f);
                else { const float4 q = sv[imin]; res = CBIG ? mk(q.x, q.y, q.z) : xf_point(sR, st, mk(q.x, q.y, q.z)); }
                served = true;
            }
        }
    }
    if (want && !served) res = (CBIG && isbig) ? support<1, true>(v, n, R, t, d) : support<1, false>(v, n, R, t, d);
    return res;
}

NP_D V3 sel4(V3 a, V3 b, V3 c, V3 d, int k) { return k == 0 ? a : (k == 1 ? b : (k == 2 ? c : d)); }

// ---------------------------------------------------------------------------------------------------
// GJK: DetectCollisionStep loop of DetectCollision (collision_detection.cpp:620-713), registers only.
// One trip = one support pair, then the second half of step k (termination tests, push) and the first half
// of step k+1 (solve, containment test, next direction).  The simplex is kept in named scalars (never indexed
// through a pointer) so that it cannot be demoted to local memory.
// ---------------------------------------------------------------------------------------------------
template <int G, bool CACHED>
struct GjkMachine {
    enum { RUN = 0, MISS = 1, HIT = 2 };
    V3 P0, P1, P2, P3, A0, A1, A2, A3, B0, B1, B2, B3;
    V3 d;
    int nv, kn, k0, k1, k2;    // kn,k0..k2: vertices the last solve keeps (applied lazily, after the duplicate test)
    int steps, status, seed_alive;
#ifdef NP_TIMELINE
    long long cyc[3];      // supports | consume | produce (accumulated over the body's whole search, reset by the kernel)
#endif

    NP_D void begin()
    {
        nv = 0; steps = 0; status = RUN; seed_alive = 1;
        d = mk(1.0f, 1.0f, 1.0f);                                                                 // :702-703 seed directions (1,1,1) / (-1,-1,-1)
    }

    NP_D void produce(PairStats& st)
    {
        if (steps >= NP_GJK_MAX_STEPS) { status = MISS; return; }                                 // :710 still CONTINUE => returns false
        steps++; st.gjk_steps++;
        kn = nv; k0 = 0; k1 = 1; k2 = 2;
        if (nv == 2) {                                                                            // SolveLine :18-54
            float u;
            int r = closest_line(P0, P1, &u);
            if (r == LINE_A) { kn = 1; }
            else if (r == LINE_B) { kn = 1; k0 = 1; }
        } else if (nv == 3) {                                                                     // SolveTriangle :56-118
            float u, v;
            int r = closest_triangle(P0, P1, P2, &u, &v);
            switch (r) {
            case TRI_A: kn = 1; break;
            case TRI_B: kn = 1; k0 = 1; break;
            case TRI_C: kn = 1; k0 = 2; break;
            case TRI_AB: kn = 2; break;
            case TRI_AC: kn = 2; k1 = 2; break;
            case TRI_BC: kn = 2; k0 = 1; k1 = 2; break;
            default: break;
            }
        } else if (nv == 4) {                                                                     // SolveTetrahedron :121-259
            float u, v, w;
            int r = closest_tetra(P0, P1, P2, P3, &u, &v, &w);
            switch (r) {
            case TET_A: kn = 1; break;
            case TET_B: kn = 1; k0 = 1; break;
            case TET_C: kn = 1; k0 = 2; break;
            case TET_D: kn = 1; k0 = 3; break;
            case TET_AB: kn = 2; break;
            case TET_AC: kn = 2; k1 = 2; break;
            case TET_AD: kn = 2; k1 = 3; break;
            case TET_BC: kn = 2; k0 = 1; k1 = 2; break;
            case TET_BD: kn = 2; k0 = 1; k1 = 3; break;
            case TET_CD: kn = 2; k0 = 2; k1 = 3; break;
            case TET_ABC: kn = 3; break;
            case TET_ABD: kn = 3; k2 = 3; break;
            case TET_ACD: kn = 3; k1 = 2; k2 = 3; break;
            case TET_BCD: kn = 3; k0 = 1; k1 = 2; k2 = 3; break;
            default: break;                    // ABCD, or None under NDEBUG (:255): four points stay => overlap
            }
        }
        if (kn == 4) { status = HIT; return; }                                                    // :648-652 OVERLAP
        // most solves drop nothing (kn == nv): then the kept points are P0..P(kn-1) as they stand
        const bool ident = (kn == nv);
        V3 q0 = ident ? P0 : sel4(P0, P1, P2, P3, k0);
        if (kn == 1) d = dir1(q0);
        else {
            V3 q1 = ident ? P1 : sel4(P0, P1, P2, P3, k1);
            if (kn == 2) d = dir2(q0, q1);
            else d = dir3(q0, q1, ident ? P2 : sel4(P0, P1, P2, P3, k2));
        }
        if (feq(magsq(d), 0.0f)) status = MISS;                                                   // :655-658 NO_OVERLAP
    }

    NP_D void consume(V3 sa, V3 sb)
    {
        V3 sp = sub(sa, sb);
        if (dot(d, sp) < 0) { status = MISS; return; }                                            // :667
        // :674-680 the new support pair against the PRE-solve points
        bool dup = (veq(sa, A0) && veq(sb, B0));
        if (nv > 1) dup = dup || (veq(sa, A1) && veq(sb, B1));
        if (nv > 2) dup = dup || (veq(sa, A2) && veq(sb, B2));
        if (nv > 3) dup = dup || (veq(sa, A3) && veq(sb, B3));
        if (dup) { status = MISS; return; }
        // the solve's order-preserving compaction (skipped when nothing was dropped), then push (:683-691)
        if (kn != nv) {
            V3 nP0 = sel4(P0, P1, P2, P3, k0), nA0 = sel4(A0, A1, A2, A3, k0), nB0 = sel4(B0, B1, B2, B3, k0);
            V3 nP1 = sel4(P0, P1, P2, P3, k1), nA1 = sel4(A0, A1, A2, A3, k1), nB1 = sel4(B0, B1, B2, B3, k1);
            V3 nP2 = sel4(P0, P1, P2, P3, k2), nA2 = sel4(A0, A1, A2, A3, k2), nB2 = sel4(B0, B1, B2, B3, k2);
            P0 = nP0; A0 = nA0; B0 = nB0; P1 = nP1; A1 = nA1; B1 = nB1; P2 = nP2; A2 = nA2; B2 = nB2;
            if (k0 != 0) seed_alive = 0;         // the seed (p = B - A) can only ever sit at index 0
        }
        if (kn == 1) { P1 = sp; A1 = sa; B1 = sb; }
        else if (kn == 2) { P2 = sp; A2 = sa; B2 = sb; }
        else { P3 = sp; A3 = sa; B3 = sb; }
        nv = kn + 1;
    }

    // a trip whose support pair was evaluated by the caller (support_coop); call while status == RUN
    NP_D void trip_with(const PairGeom& g, V3 sa, V3 sb, PairStats& st)
    {
        st.support_pairs++; st.support_verts += (unsigned)(g.na + g.nb);
        if (nv == 0) { A0 = sa; B0 = sb; P0 = sub(sb, sa); nv = 1; }                              // :704  p = B - A
        else consume(sa, sb);
        if (status == RUN) produce(st);
    }

    // call while status == RUN
    NP_D void trip(const PairGeom& g, PairStats& st)
    {
#ifdef NP_TIMELINE
        const long long c0 = clock64();
#endif
        V3 sa = support_a<G, CACHED>(g, d);
        V3 sb = support_b<G, CACHED>(g, neg(d));
#ifdef NP_TIMELINE
        const long long c1 = clock64(); cyc[0] += c1 - c0;
#endif
        st.support_pairs++; st.support_verts += (unsigned)(g.na + g.nb);
        if (nv == 0) { A0 = sa; B0 = sb; P0 = sub(sb, sa); nv = 1; }                              // :704  p = B - A
        else consume(sa, sb);
#ifdef NP_TIMELINE
        const long long c2 = clock64(); cyc[1] += c2 - c1;
#endif
        if (status == RUN) produce(st);
#ifdef NP_TIMELINE
        cyc[2] += clock64() - c2;
#endif
    }
};

// The solve half of a GJK step (SolveLine / SolveTriangle / SolveTetrahedron, the containment test and GetSearchDirection,
// collision_detection.cpp:18-259, 648-658) on a simplex held in ONE COLUMN of a block's shared simplex array -- any column, not only
// the calling thread's own: k_gjk2 hands the block's pending solves out grouped by simplex size, so that a warp runs one solver.
// in: perm (logical vertex k sits in slot (perm >> 2k) & 3), nv.  out: kn / ksel (the vertices the solve keeps), the next direction, status.
NP_D void gjk_solve_col(const float* col, int stride, int perm, int nv, int& kn, int& ksel, V3& d, int& status)
{
    auto ldP = [&](int k) { const int s3 = 3 * ((perm >> (2 * k)) & 3); return mk(col[s3 * stride], col[(s3 + 1) * stride], col[(s3 + 2) * stride]); };
    kn = nv; int k0 = 0, k1 = 1, k2 = 2;
    if (nv == 2) {                                                                                // SolveLine :18-54
        float u;
        int r = closest_line(ldP(0), ldP(1), &u);
        if (r == LINE_A) { kn = 1; }
        else if (r == LINE_B) { kn = 1; k0 = 1; }
    } else if (nv == 3) {                                                                         // SolveTriangle :56-118
        float u, v;
        int r = closest_triangle(ldP(0), ldP(1), ldP(2), &u, &v);
        switch (r) {
        case TRI_A: kn = 1; break;
        case TRI_B: kn = 1; k0 = 1; break;
        case TRI_C: kn = 1; k0 = 2; break;
        case TRI_AB: kn = 2; break;
        case TRI_AC: kn = 2; k1 = 2; break;
        case TRI_BC: kn = 2; k0 = 1; k1 = 2; break;
        default: break;
        }
    } else if (nv == 4) {                                                                         // SolveTetrahedron :121-259
        float u, v, w;
        int r = closest_tetra(ldP(0), ldP(1), ldP(2), ldP(3), &u, &v, &w);
        switch (r) {
        case TET_A: kn = 1; break;
        case TET_B: kn = 1; k0 = 1; break;
        case TET_C: kn = 1; k0 = 2; break;
        case TET_D: kn = 1; k0 = 3; break;
        case TET_AB: kn = 2; break;
        case TET_AC: kn = 2; k1 = 2; break;
        case TET_AD: kn = 2; k1 = 3; break;
        case TET_BC: kn = 2; k0 = 1; k1 = 2; break;
        case TET_BD: kn = 2; k0 = 1; k1 = 3; break;
        case TET_CD: kn = 2; k0 = 2; k1 = 3; break;
        case TET_ABC: kn = 3; break;
        case TET_ABD: kn = 3; k2 = 3; break;
        case TET_ACD: kn = 3; k1 = 2; k2 = 3; break;
        case TET_BCD: kn = 3; k0 = 1; k1 = 2; k2 = 3; break;
        default: break;                    // ABCD, or None under NDEBUG (:255): four points stay => overlap
        }
    }
    if (kn == 4) { status = 2 /* HIT */; return; }                                                // :648-652
    ksel = k0 | (k1 << 2) | (k2 << 4);
    const V3 q0 = ldP(k0);
    if (kn == 1) d = dir1(q0);
    else {
        const V3 q1 = ldP(k1);
        if (kn == 2) d = dir2(q0, q1);
        else d = dir3(q0, q1, ldP(k2));
    }
    if (feq(magsq(d), 0.0f)) status = 1 /* MISS */;                                               // :655-658
}

// ---------------------------------------------------------------------------------------------------
// GJK, thread per pair, simplex in SHARED memory (k_gjk1).  The register-resident machine above needs 36 floats of
// simplex plus the sel4 chains that compact it; with one pair per lane that is what pushed the kernel to 150 registers
// and 3 resident blocks per SM.  Here the 4 x {p, A, B} slots live in a per-block shared array laid out
// [36 rows][blockDim] (row-major, one column per thread: conflict free) and the solve's order-preserving compaction is a
// permutation of slot numbers: `perm` maps logical vertex k to its physical slot, nothing moves.  Same arithmetic, same
// order of the tests as GjkMachine (collision_detection.cpp:620-713).
// ---------------------------------------------------------------------------------------------------
#define NP_SX_ROWS 36
struct GjkS {
    enum { RUN = 0, MISS = 1, HIT = 2 };
    float* sm;                 // this thread's column of the block's simplex array
    int stride;                // blockDim.x
    V3 d;
    int nv, perm, kn, ksel;    // ksel = k0 | k1 << 2 | k2 << 4: logical vertices the last solve keeps (applied lazily, after the duplicate test)
    int steps, status, seed_alive;

    NP_D int slot(int k) const { return (perm >> (2 * k)) & 3; }
    NP_D V3 ld(int row) const { return mk(sm[row * stride], sm[(row + 1) * stride], sm[(row + 2) * stride]); }
    NP_D V3 ldP(int s) const { return ld(3 * s); }
    NP_D V3 ldA(int s) const { return ld(12 + 3 * s); }
    NP_D V3 ldB(int s) const { return ld(24 + 3 * s); }
    NP_D void stv(int row, V3 v) { sm[row * stride] = v.x; sm[(row + 1) * stride] = v.y; sm[(row + 2) * stride] = v.z; }
    NP_D void put(int s, V3 p, V3 a, V3 b) { stv(3 * s, p); stv(12 + 3 * s, a); stv(24 + 3 * s, b); }

    NP_D void begin()
    {
        nv = 0; steps = 0; status = RUN; seed_alive = 1; perm = 0; kn = 0; ksel = 0 | (1 << 2) | (2 << 4);
        d = mk(1.0f, 1.0f, 1.0f);                                                                 // :702-703
    }
    NP_D void produce(PairStats& st)
    {
        if (steps >= NP_GJK_MAX_STEPS) { status = MISS; return; }                                 // :710
        steps++; st.gjk_steps++;
        gjk_solve_col(sm, stride, perm, nv, kn, ksel, d, status);
    }
    NP_D void consume(V3 sa, V3 sb)
    {
        const V3 sp = sub(sa, sb);
        if (dot(d, sp) < 0) { status = MISS; return; }                                            // :667
        bool dup = false;                                                                         // :674-680 against the PRE-solve points
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < nv) { const int s = slot(k); dup = dup || (veq(sa, ldA(s)) && veq(sb, ldB(s))); }
        if (dup) { status = MISS; return; }
        int np_ = perm;
        if (kn != nv) {                                                                           // the solve's order-preserving compaction
            const int k0 = ksel & 3, k1 = (ksel >> 2) & 3, k2 = (ksel >> 4) & 3;
            np_ = slot(k0) | (slot(k1) << 2) | (slot(k2) << 4);
            if (k0 != 0) seed_alive = 0;         // the seed (p = B - A) can only ever sit at index 0
        }
        int used = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) if (k < kn) used |= 1 << ((np_ >> (2 * k)) & 3);
        const int fr = __ffs(~used & 15) - 1;
        put(fr, sp, sa, sb);                                                                      // :683-691
        perm = (np_ & ((1 << (2 * kn)) - 1)) | (fr << (2 * kn));
        nv = kn + 1;
    }
    // a trip whose support pair was evaluated by the caller; call while status == RUN
    NP_D void trip_with(int nverts, V3 sa, V3 sb, PairStats& st)
    {
        st.support_pairs++; st.support_verts += (unsigned)nverts;
        if (nv == 0) { put(0, sub(sb, sa), sa, sb); perm = 0; nv = 1; kn = 1; }                   // :704  p = B - A
        else consume(sa, sb);
        if (status == RUN) produce(st);
    }
    // the final simplex in logical order, as k_epa reads it: A[4] then B[4], packed in 6 float4
    NP_D void write_hit(float4* o) const
    {
        const V3 a0 = ldA(slot(0)), a1 = ldA(slot(1)), a2 = ldA(slot(2)), a3 = ldA(slot(3));
        const V3 b0 = ldB(slot(0)), b1 = ldB(slot(1)), b2 = ldB(slot(2)), b3 = ldB(slot(3));
        o[0] = make_float4(a0.x, a0.y, a0.z, a1.x); o[1] = make_float4(a1.y, a1.z, a2.x, a2.y); o[2] = make_float4(a2.z, a3.x, a3.y, a3.z);
        o[3] = make_float4(b0.x, b0.y, b0.z, b1.x); o[4] = make_float4(b1.y, b1.z, b2.x, b2.y); o[5] = make_float4(b2.z, b3.x, b3.y, b3.z);
    }
};

// ---------------------------------------------------------------------------------------------------
// polytope storage policies (one polytope per group)
// ---------------------------------------------------------------------------------------------------
#define NP_POLY_ACCESSORS                                                                                              \
    NP_D V3 P(int i) const { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }                                        \
    NP_D V3 gA(int i) const { return mk(A[3 * i], A[3 * i + 1], A[3 * i + 2]); }                                       \
    NP_D V3 gB(int i) const { return mk(B[3 * i], B[3 * i + 1], B[3 * i + 2]); }                                       \
    NP_D void setv(int i, V3 pp, V3 a, V3 b)                                                                           \
    {                                                                                                                  \
        p[3 * i] = pp.x; p[3 * i + 1] = pp.y; p[3 * i + 2] = pp.z;                                                     \
        A[3 * i] = a.x; A[3 * i + 1] = a.y; A[3 * i + 2] = a.z;                                                        \
        B[3 * i] = b.x; B[3 * i + 1] = b.y; B[3 * i + 2] = b.z;                                                        \
    }                                                                                                                  \
    NP_D V3 N(int f) const { return mk(fn[3 * f], fn[3 * f + 1], fn[3 * f + 2]); }                                     \
    NP_D void setf(int f, uint32_t idx, V3 n) { fi[f] = idx; fn[3 * f] = n.x; fn[3 * f + 1] = n.y; fn[3 * f + 2] = n.z; } \
    NP_D uint32_t F(int f) const { return fi[f]; }                                                                     \
    NP_D float D(int f) const { return fd[f]; }                                                                        \
    NP_D void setD(int f, float v) { fd[f] = v; }                                                                      \
    NP_D uint16_t E(int e) const { return edge[e]; }                                                                   \
    NP_D void setE(int e, uint16_t v) { edge[e] = v; }

template <int FMAX>
struct PolyLocal {                      // G = 1: per-thread local memory
    float p[NP_EPA_MAX_VERTS * 3], A[NP_EPA_MAX_VERTS * 3], B[NP_EPA_MAX_VERTS * 3];
    float fn[FMAX * 3];
    float fd[FMAX];                     // squared distance of the face's closest point to the origin (fixed once the face exists)
    uint32_t fi[FMAX];                  // i0 | i1 << 8 | i2 << 16
    uint16_t edge[3 * FMAX];            // va | vb << 8
    int scratch[4];
    static constexpr int fmax = FMAX;
    static constexpr int emax = 3 * FMAX;
    NP_POLY_ACCESSORS
};
struct PolyPtr {                        // shared memory (G >= 4) or the global large-polytope arena
    float* p; float* A; float* B; float* fn; float* fd; uint32_t* fi; uint16_t* edge; int* scratch;
    uint16_t* seq;      // overflow tiers: per horizon token, its sequence number among inserted edges of the same direction
    uint16_t* tab;      // overflow tiers: [2][24*24] inserted / cancelled counters per directed edge
    int fmax, emax;
    NP_POLY_ACCESSORS
};
// words of shared memory one group's polytope needs
__host__ __device__ constexpr int poly_words(int fmax) { return 3 * NP_EPA_MAX_VERTS * 3 + 3 * fmax + fmax + fmax + (3 * fmax + 1) / 2 + 4; }
// the overflow tiers add the token sequence numbers and the directed-edge counter table
__host__ __device__ constexpr int poly_words_fast(int fmax) { return poly_words(fmax) + (3 * fmax + 1) / 2 + NP_EPA_MAX_VERTS * NP_EPA_MAX_VERTS; }
NP_D PolyPtr poly_from(float* base, int fmax, bool fast = false)
{
    PolyPtr q;
    q.p = base; q.A = q.p + 3 * NP_EPA_MAX_VERTS; q.B = q.A + 3 * NP_EPA_MAX_VERTS; q.fn = q.B + 3 * NP_EPA_MAX_VERTS;
    q.fd = q.fn + 3 * fmax; q.fi = (uint32_t*)(q.fd + fmax); q.scratch = (int*)(q.fi + fmax); q.edge = (uint16_t*)(q.scratch + 4);
    q.seq = fast ? (uint16_t*)(base + poly_words(fmax)) : nullptr;
    q.tab = fast ? q.seq + 2 * ((3 * fmax + 1) / 2) : nullptr;
    q.fmax = fmax; q.emax = 3 * fmax;
    return q;
}

// ---------------------------------------------------------------------------------------------------
// EPA: FindIntersectionPoints on an origin-enclosing simplex (collision_detection.cpp:412-566)
// ---------------------------------------------------------------------------------------------------
template <int G, class Poly, bool FAST, bool CACHED>
struct EpaMachine {
    int pnv, pnf, steps;
    V3 d;            // the stored normal of the closest face (the direction the polytope is pushed out along)
    float dist;
    int flags; V3 pen; float depth;
    bool done;
#ifdef NP_TIMELINE
    long long cyc[4];      // produce | supports | consume up to the token replay | survivors + new faces
#endif

    // Simplex::AddFace, simplex.h:37-53 (called by one lane per face).  Also evaluates, once, what every later
    // GetClosestTriangleToOrigin scan (:300-318) would recompute for this face: the squared distance of its closest
    // point to the origin.  The face's vertices never change, so the value is the same bits on every scan.
    NP_D void add_face_at(Poly& poly, int slot, int va, int vb, int vc)
    {
        V3 a = poly.P(va), b = poly.P(vb), c = poly.P(vc);
        V3 n = normalize(cross(sub(b, a), sub(c, a)));
        V3 nn = (dot(n, a) != 0.0f) ? n : neg(n);                                                 // :52 float used as bool
        poly.setf(slot, (uint32_t)va | ((uint32_t)vb << 8) | ((uint32_t)vc << 16), nn);
        float u, v;
        closest_triangle(a, b, c, &u, &v);
        poly.setD(slot, magsq(add(add(a, mul(sub(b, a), u)), mul(sub(c, a), v))));
    }

    // SetupForEPA (simplex.h:29-36) from the GJK result: 4 points {A,B}; p = A - B, except the surviving seed (B - A)
    NP_D void begin(Poly& poly, const V3* A4, const V3* B4, int seed_alive)
    {
        pnv = 4; pnf = 4; steps = 0; flags = NP_PAIR_HIT; pen = mk(0.0f, 0.0f, 0.0f); depth = 0.0f; dist = 0.0f; done = false;
#ifdef NP_TIMELINE
        cyc[0] = cyc[1] = cyc[2] = cyc[3] = 0;
#endif
        d = mk(0.0f, 0.0f, 0.0f);
        if constexpr (FAST) { for (int q = glane<G>(); q < NP_EPA_MAX_VERTS * NP_EPA_MAX_VERTS; q += G) poly.tab[q] = 0; }
        if (glane<G>() == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) poly.setv(k, (k == 0 && seed_alive) ? sub(B4[0], A4[0]) : sub(A4[k], B4[k]), A4[k], B4[k]);
        }
        gsync<G>();
        // faces 012, 023, 310, 132 (simplex.h:29-36), two bits per index, face l in bits [6l, 6l+6): no indexed local array
        const unsigned f0 = (0u | (1u << 2) | (2u << 4)) | ((0u | (2u << 2) | (3u << 4)) << 6) | ((3u | (1u << 2) | (0u << 4)) << 12) | ((1u | (3u << 2) | (2u << 4)) << 18);
        if (G >= 4) { int l = glane<G>(); if (l < 4) { const unsigned f = f0 >> (6 * l); add_face_at(poly, l, f & 3, (f >> 2) & 3, (f >> 4) & 3); } }
        else if (glane<G>() == 0) {
#pragma unroll
            for (int l = 0; l < 4; l++) { const unsigned f = f0 >> (6 * l); add_face_at(poly, l, f & 3, (f >> 2) & 3, (f >> 4) & 3); }
        }
        gsync<G>();
    }

    NP_D void converge()
    {
        pen = neg(normalize(d));                                                                  // :530
        depth = dist;
        flags |= NP_PAIR_SUCCESS;
        done = true;
    }

    // first half of FindIntersectionPointsStep (:412-425): closest face, distance test
    NP_D void produce(Poly& poly, PairStats& st)
    {
        if (steps >= NP_EPA_MAX_STEPS || pnf == 0) { done = true; return; }                       // not converged: success stays false
        steps++; st.epa_iters++;
        float best = 3.402823466e+38f; int bi = 0x7fffffff;                                       // GetClosestTriangleToOrigin :292-325
        for (int f = glane<G>(); f < pnf; f += G) {
            float dsq = poly.D(f);
            if (dsq < best) { best = dsq; bi = f; }                                               // strict '<': first minimum wins
        }
        if (G > 1) {                             // smaller distance wins, lower face index on ties: two warp reductions
            const unsigned m = gmask<G>();
            const int key = float_order(best + 0.0f);
            const int kmin = __reduce_min_sync(m, key);
            bi = (int)__reduce_min_sync(m, (key == kmin) ? (unsigned)bi : 0xffffffffu);
            best = order_float(kmin);
            if ((unsigned)bi >= 0x7fffffffu) bi = 0x7fffffff;
        }
        if (bi == 0x7fffffff) bi = 0;            // every distance NaN: the reference's (zero-pinned) face_index stays 0
        st.epa_tri_tests += pnf;
        dist = sqrtf(best);
        d = poly.N(bi);                                                                           // :408 the STORED normal
        if (!(dist > NP_EPA_TOL)) converge();                                                     // :424, :523-531
    }

    // second half of FindIntersectionPointsStep (:427-520)
    NP_D void consume(Poly& poly, V3 sa, V3 sb)
    {
        const int gl = glane<G>();
        const unsigned m = gmask<G>();
        const int gbase = (int)((threadIdx.x & 31u) & ~(unsigned)(G - 1));     // first lane of this group inside the warp
        V3 sp = sub(sa, sb);
        bool found = false;
        for (int i = gl; i < pnv; i += G)
            if (veq(sa, poly.gA(i), NP_EPA_TOL) && veq(sb, poly.gB(i), NP_EPA_TOL)) found = true; // :435-446
        if (G > 1) found = (__ballot_sync(m, found) & m) != 0u;
        if (found) { converge(); return; }
        // :469-510 drop every face whose stored normal faces the support POINT, keep the rest in order, and collect the
        // horizon: edges in insertion order, an edge cancelling the first earlier copy of its reverse.
        // Faces are handled in rounds of G: visibility per lane, horizon edges by lane 0 (order dependent), then an
        // order-preserving parallel compaction (read, sync, write; destinations never run ahead of the round).
        int ne = 0, keep = 0, overflow = 0;
        for (int base = 0; base < pnf; base += G) {
            const int f = base + gl;
            const bool valid = f < pnf;
            uint32_t idx = 0; V3 fnrm = mk(0.f, 0.f, 0.f); float fdist = 0.f;
            if (valid) { idx = poly.F(f); fnrm = poly.N(f); fdist = poly.D(f); }
            const bool vis = valid && (dot(fnrm, sp) > 0);
            unsigned vmask, kmask;
            if (G > 1) { vmask = (__ballot_sync(m, vis) & m) >> gbase; kmask = (__ballot_sync(m, valid && !vis) & m) >> gbase; }
            else { vmask = vis ? 1u : 0u; kmask = (valid && !vis) ? 1u : 0u; }
            if (gl == 0) {
                unsigned vm = vmask;
                while (vm) {
                    int bit = __ffs(vm) - 1; vm &= vm - 1;
                    uint32_t fidx = poly.F(base + bit);
                    int v0 = fidx & 0xff, v1 = (fidx >> 8) & 0xff, v2 = (fidx >> 16) & 0xff;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        int ea = k == 0 ? v0 : (k == 1 ? v1 : v2);
                        int eb = k == 0 ? v1 : (k == 1 ? v2 : v0);
                        uint16_t rev = (uint16_t)(eb | (ea << 8));
                        int hit = -1;
                        for (int q = 0; q < ne; q++) if (poly.E(q) == rev) { hit = q; break; }
                        if (hit >= 0) { for (int q = hit; q + 1 < ne; q++) poly.setE(q, poly.E(q + 1)); ne--; }
                        else if (ne < poly.emax) { poly.setE(ne, (uint16_t)(ea | (eb << 8))); ne++; }
                        else overflow = 1;
                    }
                }
            }
            gsync<G>();                                                          // lane 0 has read its faces; everyone holds theirs in registers
            if (valid && !vis) {
                int dst = keep + __popc(kmask & ((1u << gl) - 1u));
                if (dst != f) { poly.setf(dst, idx, fnrm); poly.setD(dst, fdist); }
            }
            keep += __popc(kmask);
            gsync<G>();
        }
        if (gl == 0) {
            if (keep + ne > poly.fmax) overflow = 1;
            if (!overflow) poly.setv(pnv, sp, sa, sb);                                            // :512
            poly.scratch[0] = ne; poly.scratch[2] = overflow;
        }
        gsync<G>();
        ne = poly.scratch[0]; overflow = poly.scratch[2];
        if (overflow) { flags |= NP_PAIR_OVERFLOW; done = true; return; }
        for (int e = gl; e < ne; e += G) { uint16_t ed = poly.E(e); add_face_at(poly, keep + e, ed & 0xff, ed >> 8, pnv); }   // :514-517
        pnf = keep + ne; pnv++;
        gsync<G>();
    }

    // Same expansion for polytopes with hundreds of faces (overflow tiers, one warp per pair).  The reference's
    // UniqueEdges (:469-491) makes a new edge (a,b) erase the FIRST remaining copy of (b,a), else appends it.  Per
    // directed edge that is a FIFO: the surviving copies of (a,b) are its inserted copies number [cancelled, added).
    // Edges of different vertex pairs never interact, so every lane owns the pairs that hash to it and replays only
    // their tokens, in order, with O(1) counter updates; survivors keep their insertion order.
    NP_D void consume_fast(Poly& poly, V3 sa, V3 sb)
    {
        static_assert(!FAST || G == 32, "fast horizon needs a full warp per pair");
        const int gl = glane<G>();
        const unsigned m = 0xffffffffu;
        V3 sp = sub(sa, sb);
        bool found = false;
        for (int i = gl; i < pnv; i += G)
            if (veq(sa, poly.gA(i), NP_EPA_TOL) && veq(sb, poly.gB(i), NP_EPA_TOL)) found = true;
        found = __ballot_sync(m, found) != 0u;
        if (found) { converge(); return; }
        // pass 1: visibility, horizon tokens (3 per visible face, in face order), order-preserving compaction of the rest.  The directed
        // edges of the removed faces are also marked in a 24x24 bit table: as long as none occurs twice, UniqueEdges keeps exactly the
        // edges whose reverse is not an edge of a removed face, in token order -- whichever of the two came first, the later one erases
        // the earlier and is not added -- and the replay of pass 2 is not needed.
        uint32_t* bits = (uint32_t*)(poly.tab + NP_EPA_MAX_VERTS * NP_EPA_MAX_VERTS);      // the unused half of the counter-table area
        if (gl < (NP_EPA_MAX_VERTS * NP_EPA_MAX_VERTS + 31) / 32) bits[gl] = 0u;
        __syncwarp(m);
        bool dup = false;
        int keep = 0, ntok = 0;
        for (int base = 0; base < pnf; base += G) {
            const int f = base + gl;
            const bool valid = f < pnf;
            uint32_t idx = 0; V3 fnrm = mk(0.f, 0.f, 0.f); float fdist = 0.f;
            if (valid) { idx = poly.F(f); fnrm = poly.N(f); fdist = poly.D(f); }
            const bool vis = valid && (dot(fnrm, sp) > 0);
            const unsigned vmask = __ballot_sync(m, vis), kmask = __ballot_sync(m, valid && !vis);
            if (vis) {
                int t = ntok + 3 * __popc(vmask & ((1u << gl) - 1u));
                int v0 = idx & 0xff, v1 = (idx >> 8) & 0xff, v2 = (idx >> 16) & 0xff;
                if (t + 2 < poly.emax) { poly.setE(t, (uint16_t)(v0 | (v1 << 8))); poly.setE(t + 1, (uint16_t)(v1 | (v2 << 8))); poly.setE(t + 2, (uint16_t)(v2 | (v0 << 8))); }
                const int e0 = v0 * NP_EPA_MAX_VERTS + v1, e1 = v1 * NP_EPA_MAX_VERTS + v2, e2 = v2 * NP_EPA_MAX_VERTS + v0;
                dup |= (atomicOr(&bits[e0 >> 5], 1u << (e0 & 31)) >> (e0 & 31)) & 1u;
                dup |= (atomicOr(&bits[e1 >> 5], 1u << (e1 & 31)) >> (e1 & 31)) & 1u;
                dup |= (atomicOr(&bits[e2 >> 5], 1u << (e2 & 31)) >> (e2 & 31)) & 1u;
            }
            ntok += 3 * __popc(vmask);
            __syncwarp(m);
            if (valid && !vis) {
                int dst = keep + __popc(kmask & ((1u << gl) - 1u));
                if (dst != f) { poly.setf(dst, idx, fnrm); poly.setD(dst, fdist); }
            }
            keep += __popc(kmask);
            __syncwarp(m);
        }
        // pass 2: replay the tokens in rounds of 32.  tab[a*24+b] = inserted (low byte) | cancelled (high byte) of directed
        // edge (a,b); the table is all zero between expansions.  Tokens of one unordered vertex pair are handled, in order,
        // by the lowest lane holding one of them (__match_any_sync); different pairs touch different entries.
        const bool replay = __ballot_sync(m, dup) != 0u;        // a directed edge occurred twice: the order of the erasures matters
        int bad = 0;
        for (int base = 0; replay && base < ntok; base += G) {
            const int t = base + gl;
            const bool tv = t < ntok;
            const uint16_t ed = tv ? poly.E(t) : (uint16_t)0;
            const int a0 = ed & 0xff, b0 = ed >> 8;
            const unsigned key = tv ? (unsigned)((a0 < b0 ? a0 : b0) * NP_EPA_MAX_VERTS + (a0 < b0 ? b0 : a0)) : (0x10000u + (unsigned)gl);
            unsigned peers = __match_any_sync(m, key);
            if (tv && (__ffs(peers) - 1) == gl) {
                while (peers) {
                    const int bit = __ffs(peers) - 1; peers &= peers - 1;
                    const int tt = base + bit;
                    const uint16_t e2 = (bit == gl) ? ed : poly.E(tt);
                    const int a = e2 & 0xff, b = e2 >> 8;
                    uint16_t fw = poly.tab[a * NP_EPA_MAX_VERTS + b], rv = poly.tab[b * NP_EPA_MAX_VERTS + a];
                    if ((rv & 0xff) > (rv >> 8)) { poly.tab[b * NP_EPA_MAX_VERTS + a] = (uint16_t)(rv + 0x100); poly.seq[tt] = 0xffff; }
                    else { poly.seq[tt] = (uint16_t)(fw & 0xff); if ((fw & 0xff) == 0xff) bad = 1; poly.tab[a * NP_EPA_MAX_VERTS + b] = (uint16_t)(fw + 1); }
                }
            }
            __syncwarp(m);
        }
        bad = __ballot_sync(m, bad != 0) != 0u;
#ifdef NP_TIMELINE
        const long long cp3 = clock64();
#endif
        // pass 3: survivors, in insertion order, compacted into the (already consumed) head of seq[]; then one new face per survivor.
        // AddFace is the long dependent chain of the expansion (normalise + closest point on the triangle), so it runs
        // ceil(survivors / 32) times, not once per round of tokens.
        int ne = 0;
        for (int base = 0; base < ntok; base += G) {
            const int t = base + gl;
            bool alive = false; uint16_t ed = 0;
            if (t < ntok) {
                ed = poly.E(t);
                if (replay) { uint16_t s = poly.seq[t]; alive = (s != 0xffff) && (s >= (poly.tab[(ed & 0xff) * NP_EPA_MAX_VERTS + (ed >> 8)] >> 8)); }
                else { const int r = (ed >> 8) * NP_EPA_MAX_VERTS + (ed & 0xff); alive = ((bits[r >> 5] >> (r & 31)) & 1u) == 0u; }
            }
            const unsigned am = __ballot_sync(m, alive);
            __syncwarp(m);                                                       // every lane has read its seq[t]: positions <= t may be overwritten
            if (alive) poly.seq[ne + __popc(am & ((1u << gl) - 1u))] = ed;
            ne += __popc(am);
            __syncwarp(m);
        }
        if (bad || keep + ne > poly.fmax) { if (replay) clear_table(poly, ntok); flags |= NP_PAIR_OVERFLOW; done = true; return; }
        if (gl == 0) poly.setv(pnv, sp, sa, sb);
        __syncwarp(m);
        for (int e = gl; e < ne; e += G) { const uint16_t ed = poly.seq[e]; add_face_at(poly, keep + e, ed & 0xff, ed >> 8, pnv); }
#ifdef NP_TIMELINE
        cyc[3] += clock64() - cp3;
#endif
        if (replay) clear_table(poly, ntok);
        pnf = keep + ne; pnv++;
        __syncwarp(m);
    }
    // leave the directed-edge table all zero again: only the entries this expansion touched
    NP_D void clear_table(Poly& poly, int ntok)
    {
        __syncwarp(0xffffffffu);
        for (int t = glane<G>(); t < ntok; t += G) {
            uint16_t ed = poly.E(t); int a = ed & 0xff, b = ed >> 8;
            poly.tab[a * NP_EPA_MAX_VERTS + b] = 0; poly.tab[b * NP_EPA_MAX_VERTS + a] = 0;
        }
    }

    // call while !done: [closest face] -> [support pair] -> [convergence test / expansion]
    NP_D void trip(const PairGeom& g, Poly& poly, PairStats& st)
    {
#ifdef NP_TIMELINE
        long long c0 = clock64();
#endif
        produce(poly, st);
#ifdef NP_TIMELINE
        long long c1 = clock64(); cyc[0] += c1 - c0;
#endif
        if (done) return;
        V3 sa = support_a<G, CACHED>(g, d);
        V3 sb = support_b<G, CACHED>(g, neg(d));
#ifdef NP_TIMELINE
        long long c2 = clock64(); cyc[1] += c2 - c1;
#endif
        st.support_pairs++; st.support_verts += (unsigned)(g.na + g.nb);
        if constexpr (FAST) consume_fast(poly, sa, sb); else consume(poly, sa, sb);
#ifdef NP_TIMELINE
        cyc[2] += clock64() - c2;
#endif
    }
};

}  // namespace np
