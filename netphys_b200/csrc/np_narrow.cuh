// netphys_b200/csrc/np_narrow.cuh -- GJK + EPA narrow phase as a per-thread state machine.
//
// Replaces DetectCollision / DetectCollisionStep / FindIntersectionPoints[Step] and the simplex store
// (reference engine/collision_detection.cpp:18-732, engine/simplex.h:5-76).  One thread owns one pair.
//
// B200 shape of the algorithm: the reference alternates "decide a direction" (simplex solve or EPA face
// scan) and "evaluate the support pair".  The support pair is the hot loop (27 flop/vertex in the
// reference) and is identical for GJK and EPA, so the machine is written as
//
//        loop:  [support pair for direction d]  ->  [consume: GJK tests / EPA expansion]  ->  [produce next d]
//
// with exactly ONE support evaluation per trip.  Lanes of a warp that are in different GJK iterations, or
// already in EPA, still execute the support loop together (SIMT-convergent); only the cheaper consume /
// produce phases diverge.  A finished lane is re-armed with new work by the caller (kernels in
// np_kernels.cuh), so lanes do not idle behind the slowest pair of the warp.
//
// Storage: the GJK simplex (<= 4 points x {p,A,B}) lives in registers (every index is static after
// unrolling).  The EPA polytope (<= 24 vertices, <= FMAX faces) lives in per-thread local memory; a pair
// whose polytope outgrows FMAX is flagged NP_PAIR_OVERFLOW and re-run by the large-polytope kernel
// (same code, global-memory arena) -- faces are never truncated.
#pragma once
#include "np_math.cuh"

namespace np {

#define NP_EPA_MAX_VERTS 24        // 4 GJK points + at most one per EPA iteration (20)
#define NP_GJK_MAX_STEPS 20        // collision_detection.cpp:707
#define NP_EPA_MAX_STEPS 20        // collision_detection.cpp:728 (max_steps = maxIterations)
#define NP_EPA_TOL 0.01f           // collision_detection.cpp:423

enum PairFlags { NP_PAIR_HIT = 1, NP_PAIR_SUCCESS = 2, NP_PAIR_OVERFLOW = 4 };

struct PairGeom {
    float Ra[9], Rb[9];
    V3 ta, tb;
    const float4* va; const float4* vb;   // shape-local vertices (x,y,z,-)
    int na, nb;
};

struct PairStats { unsigned gjk_steps, support_pairs, epa_iters, epa_tri_tests; };

// PhysUtil_GetPointFurthestInDirection, physics_util.cpp:5-35: world*p per vertex, strict '>' from -inf,
// first maximum wins, result starts at (0,0,0).
NP_D V3 support(const float4* __restrict__ v, int n, const float* R, V3 t, V3 d)
{
    V3 best = mk(0.0f, 0.0f, 0.0f);
    float bd = -INFINITY;
#pragma unroll 2
    for (int k = 0; k < n; k++) {
        float4 q = v[k];
        V3 wp = xf_point(R, t, mk(q.x, q.y, q.z));
        float dp = dot(d, wp);
        if (dp > bd) { bd = dp; best = wp; }
    }
    return best;
}

NP_D V3 pick4(const V3* P, int k) { return k == 0 ? P[0] : (k == 1 ? P[1] : (k == 2 ? P[2] : P[3])); }

// Per-thread polytope arena.  Storage pointers may be local (fast path) or global (large-polytope path).
template <int FMAX>
struct PolyLocal {
    float p[NP_EPA_MAX_VERTS * 3], A[NP_EPA_MAX_VERTS * 3], B[NP_EPA_MAX_VERTS * 3];
    float fn[FMAX * 3];
    uint32_t fi[FMAX];                 // i0 | i1 << 8 | i2 << 16
    uint16_t edge[3 * FMAX];           // va | vb << 8
    static constexpr int fmax = FMAX;
    static constexpr int emax = 3 * FMAX;
    NP_D V3 P(int i) const { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
    NP_D V3 gA(int i) const { return mk(A[3 * i], A[3 * i + 1], A[3 * i + 2]); }
    NP_D V3 gB(int i) const { return mk(B[3 * i], B[3 * i + 1], B[3 * i + 2]); }
    NP_D void setv(int i, V3 pp, V3 a, V3 b)
    {
        p[3 * i] = pp.x; p[3 * i + 1] = pp.y; p[3 * i + 2] = pp.z;
        A[3 * i] = a.x; A[3 * i + 1] = a.y; A[3 * i + 2] = a.z;
        B[3 * i] = b.x; B[3 * i + 1] = b.y; B[3 * i + 2] = b.z;
    }
    NP_D V3 N(int f) const { return mk(fn[3 * f], fn[3 * f + 1], fn[3 * f + 2]); }
    NP_D void setf(int f, uint32_t idx, V3 n) { fi[f] = idx; fn[3 * f] = n.x; fn[3 * f + 1] = n.y; fn[3 * f + 2] = n.z; }
    NP_D uint32_t F(int f) const { return fi[f]; }
    NP_D void movef(int dst, int src) { fi[dst] = fi[src]; fn[3 * dst] = fn[3 * src]; fn[3 * dst + 1] = fn[3 * src + 1]; fn[3 * dst + 2] = fn[3 * src + 2]; }
    NP_D uint16_t E(int e) const { return edge[e]; }
    NP_D void setE(int e, uint16_t v) { edge[e] = v; }
};

// Same interface over a caller-provided global-memory arena (large-polytope path).
struct PolyGlobal {
    float* p; float* A; float* B; float* fn; uint32_t* fi; uint16_t* edge;
    int fmax, emax;
    NP_D V3 P(int i) const { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
    NP_D V3 gA(int i) const { return mk(A[3 * i], A[3 * i + 1], A[3 * i + 2]); }
    NP_D V3 gB(int i) const { return mk(B[3 * i], B[3 * i + 1], B[3 * i + 2]); }
    NP_D void setv(int i, V3 pp, V3 a, V3 b)
    {
        p[3 * i] = pp.x; p[3 * i + 1] = pp.y; p[3 * i + 2] = pp.z;
        A[3 * i] = a.x; A[3 * i + 1] = a.y; A[3 * i + 2] = a.z;
        B[3 * i] = b.x; B[3 * i + 1] = b.y; B[3 * i + 2] = b.z;
    }
    NP_D V3 N(int f) const { return mk(fn[3 * f], fn[3 * f + 1], fn[3 * f + 2]); }
    NP_D void setf(int f, uint32_t idx, V3 n) { fi[f] = idx; fn[3 * f] = n.x; fn[3 * f + 1] = n.y; fn[3 * f + 2] = n.z; }
    NP_D uint32_t F(int f) const { return fi[f]; }
    NP_D void movef(int dst, int src) { fi[dst] = fi[src]; fn[3 * dst] = fn[3 * src]; fn[3 * dst + 1] = fn[3 * src + 1]; fn[3 * dst + 2] = fn[3 * src + 2]; }
    NP_D uint16_t E(int e) const { return edge[e]; }
    NP_D void setE(int e, uint16_t v) { edge[e] = v; }
};

template <class Poly>
struct PairMachine {
    enum Mode { SEED = 0, GJK = 1, EPA = 2, DONE = 3 };

    // GJK simplex in registers
    V3 P[4], A[4], B[4];
    int nv;
    int kn, k0, k1, k2;        // vertices the last solve keeps (applied lazily, after the duplicate test)
    V3 d;                      // current search direction (EPA: the stored normal of the closest face)
    float dist;                // EPA: distance of the closest face
    int mode;
    int gjk_steps, epa_steps;
    int pnv, pnf;              // polytope sizes
    // result
    int flags; V3 pen; float depth;

    NP_D void begin()
    {
        mode = SEED; nv = 0; gjk_steps = 0; epa_steps = 0; pnv = 0; pnf = 0;
        flags = 0; pen = mk(0.0f, 0.0f, 0.0f); depth = 0.0f; dist = 0.0f;
        d = mk(1.0f, 1.0f, 1.0f);                                                                 // collision_detection.cpp:702-703
    }
    NP_D bool done() const { return mode == DONE; }

    // Simplex::AddFace, simplex.h:37-53
    NP_D void add_face(Poly& poly, int va, int vb, int vc)
    {
        V3 a = poly.P(va), b = poly.P(vb), c = poly.P(vc);
        V3 n = normalize(cross(sub(b, a), sub(c, a)));
        V3 nn = (dot(n, a) != 0.0f) ? n : neg(n);                                                 // :52 float used as bool
        poly.setf(pnf, (uint32_t)va | ((uint32_t)vb << 8) | ((uint32_t)vc << 16), nn);
        pnf++;
    }

    NP_D void finish_epa_converged()
    {
        pen = neg(normalize(d));                                                                  // collision_detection.cpp:530
        depth = dist;
        flags |= NP_PAIR_SUCCESS;
        mode = DONE;
    }

    // first half of DetectCollisionStep (collision_detection.cpp:620-658): solve, containment, direction
    NP_D void produce_gjk(Poly& poly, PairStats& st)
    {
        if (gjk_steps >= NP_GJK_MAX_STEPS) { mode = DONE; return; }                               // :710 still CONTINUE => miss
        gjk_steps++; st.gjk_steps++;
        kn = nv; k0 = 0; k1 = 1; k2 = 2;
        if (nv == 2) {                                                                            // SolveLine :18-54
            float u;
            int r = closest_line(P[0], P[1], &u);
            if (r == LINE_A) { kn = 1; }
            else if (r == LINE_B) { kn = 1; k0 = 1; }
        } else if (nv == 3) {                                                                     // SolveTriangle :56-118
            float u, v;
            int r = closest_triangle(P[0], P[1], P[2], &u, &v);
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
            int r = closest_tetra(P[0], P[1], P[2], P[3], &u, &v, &w);
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
        if (kn == 4) {                                                                            // :648-652 OVERLAP
            flags |= NP_PAIR_HIT;
#pragma unroll
            for (int k = 0; k < 4; k++) poly.setv(k, P[k], A[k], B[k]);
            pnv = 4; pnf = 0;
            add_face(poly, 0, 1, 2); add_face(poly, 0, 2, 3); add_face(poly, 3, 1, 0); add_face(poly, 1, 3, 2);   // simplex.h:29-36
            mode = EPA;
            return;
        }
        V3 q0 = pick4(P, k0);
        if (kn == 1) d = dir1(q0);
        else {
            V3 q1 = pick4(P, k1);
            if (kn == 2) d = dir2(q0, q1);
            else d = dir3(q0, q1, pick4(P, k2));
        }
        if (feq(magsq(d), 0.0f)) mode = DONE;                                                     // :655-658 NO_OVERLAP
    }

    // second half of DetectCollisionStep (:664-691)
    NP_D void consume_gjk(V3 sa, V3 sb)
    {
        V3 sp = sub(sa, sb);
        if (dot(d, sp) < 0) { mode = DONE; return; }                                              // :667
        bool dup = false;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < nv && veq(sa, A[k]) && veq(sb, B[k])) dup = true;                             // :674-680 against the PRE-solve points
        if (dup) { mode = DONE; return; }
        // apply the solve's order-preserving compaction, then push
        V3 nP[3], nA[3], nB[3];
        nP[0] = pick4(P, k0); nA[0] = pick4(A, k0); nB[0] = pick4(B, k0);
        nP[1] = pick4(P, k1); nA[1] = pick4(A, k1); nB[1] = pick4(B, k1);
        nP[2] = pick4(P, k2); nA[2] = pick4(A, k2); nB[2] = pick4(B, k2);
#pragma unroll
        for (int k = 0; k < 3; k++) { P[k] = nP[k]; A[k] = nA[k]; B[k] = nB[k]; }
#pragma unroll
        for (int k = 1; k < 4; k++)
            if (k == kn) { P[k] = sp; A[k] = sa; B[k] = sb; }
        nv = kn + 1;
    }

    // first half of FindIntersectionPointsStep (:412-425): closest face, distance test
    NP_D void produce_epa(Poly& poly, PairStats& st)
    {
        if (epa_steps >= NP_EPA_MAX_STEPS || pnf == 0) { mode = DONE; return; }                   // not converged: success stays false
        epa_steps++; st.epa_iters++;
        float best = 3.402823466e+38f; int bi = 0;                                                // GetClosestTriangleToOrigin :292-325
        for (int f = 0; f < pnf; f++) {
            uint32_t idx = poly.F(f);
            V3 a = poly.P(idx & 0xff), b = poly.P((idx >> 8) & 0xff), c = poly.P((idx >> 16) & 0xff);
            float u, v;
            closest_triangle(a, b, c, &u, &v);
            float dsq = magsq(add(add(a, mul(sub(b, a), u)), mul(sub(c, a), v)));
            if (dsq < best) { best = dsq; bi = f; }
        }
        st.epa_tri_tests += pnf;
        dist = sqrtf(best);
        d = poly.N(bi);                                                                           // :408 the STORED normal
        if (!(dist > NP_EPA_TOL)) finish_epa_converged();                                         // :424, :523-531
    }

    // second half of FindIntersectionPointsStep (:427-520)
    NP_D void consume_epa(Poly& poly, V3 sa, V3 sb)
    {
        V3 sp = sub(sa, sb);
        for (int i = 0; i < pnv; i++)
            if (veq(sa, poly.gA(i), NP_EPA_TOL) && veq(sb, poly.gB(i), NP_EPA_TOL)) { finish_epa_converged(); return; }   // :435-446
        // :469-510 remove faces whose stored normal faces the support POINT; horizon edges in insertion
        // order with reverse-edge cancellation
        int ne = 0, keep = 0;
        bool overflow = false;
        for (int f = 0; f < pnf; f++) {
            uint32_t idx = poly.F(f);
            if (dot(poly.N(f), sp) > 0) {
                int v0 = idx & 0xff, v1 = (idx >> 8) & 0xff, v2 = (idx >> 16) & 0xff;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    int ea = k == 0 ? v0 : (k == 1 ? v1 : v2);
                    int eb = k == 0 ? v1 : (k == 1 ? v2 : v0);
                    uint16_t rev = (uint16_t)(eb | (ea << 8));
                    int found = -1;
                    for (int q = 0; q < ne; q++) if (poly.E(q) == rev) { found = q; break; }
                    if (found >= 0) { for (int q = found; q + 1 < ne; q++) poly.setE(q, poly.E(q + 1)); ne--; }
                    else if (ne < poly.emax) { poly.setE(ne, (uint16_t)(ea | (eb << 8))); ne++; }
                    else overflow = true;
                }
            } else { if (keep != f) poly.movef(keep, f); keep++; }
        }
        if (overflow || keep + ne > poly.fmax) { flags |= NP_PAIR_OVERFLOW; mode = DONE; return; }
        pnf = keep;
        int ni = pnv;
        poly.setv(ni, sp, sa, sb); pnv++;                                                         // :512
        for (int e = 0; e < ne; e++) { uint16_t ed = poly.E(e); add_face(poly, ed & 0xff, ed >> 8, ni); }   // :514-517
    }

    // one trip: support pair, consume, produce.  Call only while !done().
    NP_D void step(const PairGeom& g, Poly& poly, PairStats& st)
    {
        V3 sa = support(g.va, g.na, g.Ra, g.ta, d);
        V3 sb = support(g.vb, g.nb, g.Rb, g.tb, neg(d));
        st.support_pairs++;
        if (mode == SEED) {
            A[0] = sa; B[0] = sb; P[0] = sub(sb, sa); nv = 1;                                     // :704  B - A
            mode = GJK;
        } else if (mode == GJK) consume_gjk(sa, sb);
        else consume_epa(poly, sa, sb);
        if (mode == GJK) produce_gjk(poly, st);
        if (mode == EPA) produce_epa(poly, st);
    }
};

}  // namespace np
