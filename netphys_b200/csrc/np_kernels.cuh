// netphys_b200/csrc/np_kernels.cuh -- CUDA kernels (sm_100a) of the per-step rigid-body path.
//
//   K1  k_integrate          Physics::Update + GetTransform             (physics.cpp:175-194, physics.h:70-76)
//   K3  k_first_hit          GetCollisions: ordered first-hit search    (physics.cpp:116-145) over GJK+EPA
//   K3p k_pairs              DetectCollision over an explicit pair list (collision_detection.cpp:694-732)
//   K4b k_epa_big            large-polytope re-run of flagged pairs
//   K5  k_count_incoming / k_fill_incoming / k_resolve   OnCollision + ApplyImpulseResponse (physics.cpp:27-114)
//   K6  k_snapshot           CommandFrameObject packing                 (netphys_server/world_s.cpp:23-51)
//
// HBM layout: structure-of-arrays by COMPONENT, every array `cap` floats long (cap a multiple of 32), so a
// warp reading component c of 32 consecutive bodies issues one fully coalesced 128-byte request.
#pragma once
#include "np_narrow.cuh"

namespace np {

// component indices of the SoA blocks
enum { S_PX = 0, S_PY, S_PZ, S_RX, S_RY, S_RZ, S_LX, S_LY, S_LZ, S_HX, S_HY, S_HZ, S_COUNT };          // dynamic state
enum { C_GX = 0, C_GY, C_GZ, C_MASS, C_ELAST, C_IINV0, C_COUNT = C_IINV0 + 9 };                       // per-body constants
enum { X_R0 = 0, X_TX = 9, X_TY, X_TZ, X_COUNT };                                                      // transform
enum { I_RESPONSE = 0, I_SHAPE, I_ISLAND_END, I_COUNT };

struct StoreView {
    int n, cap;
    float* state;      // [S_COUNT][cap]
    const float* cst;  // [C_COUNT][cap]
    float* xf;         // [X_COUNT][cap]
    const int* ints;   // [I_COUNT][cap]
};
struct ShapeView {
    const float4* verts;   // all shapes' local vertices, concatenated
    const int* off;        // first vertex of shape s
    const int* cnt;        // vertex count of shape s
    int total_verts;
};
struct DevCounters {       // mirrors np_counters (all uint64)
    unsigned long long steps, pair_tests, hits, gjk_steps, support_pairs, epa_iters, epa_tri_tests, epa_overflow, candidates;
};

// ---------------------------------------------------------------------------------------------------
// K1: integrate + transform.  HBM-bound: per body 12 state floats + 14 constants read, 9 state floats +
// 12 transform floats written.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_integrate(StoreView s, float dt, int write_xf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const int cap = s.cap;
    float* st = s.state; const float* c = s.cst;
    V3 pos = mk(st[S_PX * cap + i], st[S_PY * cap + i], st[S_PZ * cap + i]);
    V3 rot = mk(st[S_RX * cap + i], st[S_RY * cap + i], st[S_RZ * cap + i]);
    V3 L = mk(st[S_LX * cap + i], st[S_LY * cap + i], st[S_LZ * cap + i]);
    V3 H = mk(st[S_HX * cap + i], st[S_HY * cap + i], st[S_HZ * cap + i]);
    V3 g = mk(c[C_GX * cap + i], c[C_GY * cap + i], c[C_GZ * cap + i]);
    float mass = c[C_MASS * cap + i];
    if (!is_none(g)) {                                                                            // physics.cpp:178-181
        L = add(L, mul(mul(g, mass), dt));
        st[S_LX * cap + i] = L.x; st[S_LY * cap + i] = L.y; st[S_LZ * cap + i] = L.z;
    }
    if (!feq(mass, 0.0f)) {                                                                       // :184-193
        float iinv[9];
#pragma unroll
        for (int k = 0; k < 9; k++) iinv[k] = c[(C_IINV0 + k) * cap + i];
        V3 v = mul(L, 1.0f / mass);
        V3 w = m3mul(iinv, H);
        pos = add(pos, mul(v, dt));
        rot = add(rot, mul(w, dt));
        st[S_PX * cap + i] = pos.x; st[S_PY * cap + i] = pos.y; st[S_PZ * cap + i] = pos.z;
        st[S_RX * cap + i] = rot.x; st[S_RY * cap + i] = rot.y; st[S_RZ * cap + i] = rot.z;
    }
    if (write_xf) {
        float R[9]; V3 t;
        make_transform(pos, rot, R, &t);
#pragma unroll
        for (int k = 0; k < 9; k++) s.xf[(X_R0 + k) * cap + i] = R[k];
        s.xf[X_TX * cap + i] = t.x; s.xf[X_TY * cap + i] = t.y; s.xf[X_TZ * cap + i] = t.z;
    }
}

// transforms only (used after np_body_set_state / before a standalone query on world bodies)
__global__ void __launch_bounds__(256) k_transforms(StoreView s)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const int cap = s.cap;
    V3 pos = mk(s.state[S_PX * cap + i], s.state[S_PY * cap + i], s.state[S_PZ * cap + i]);
    V3 rot = mk(s.state[S_RX * cap + i], s.state[S_RY * cap + i], s.state[S_RZ * cap + i]);
    float R[9]; V3 t;
    make_transform(pos, rot, R, &t);
#pragma unroll
    for (int k = 0; k < 9; k++) s.xf[(X_R0 + k) * cap + i] = R[k];
    s.xf[X_TX * cap + i] = t.x; s.xf[X_TY * cap + i] = t.y; s.xf[X_TZ * cap + i] = t.z;
}

// ---------------------------------------------------------------------------------------------------
// narrow-phase helpers
// ---------------------------------------------------------------------------------------------------
#define NP_NARROW_THREADS 128
#define NP_SMEM_VERTS 2048          // shape table staged in shared memory when it has at most this many vertices (32 KB)
#define NP_FMAX_FAST 64             // in-thread polytope faces before a pair is handed to k_epa_big

NP_D void load_xf(const StoreView& s, int i, float* R, V3* t)
{
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = s.xf[(X_R0 + k) * s.cap + i];
    *t = mk(s.xf[X_TX * s.cap + i], s.xf[X_TY * s.cap + i], s.xf[X_TZ * s.cap + i]);
}
NP_D const float4* stage_shapes(const ShapeView& sh, float4* smem, bool use_smem)
{
    if (!use_smem) return sh.verts;
    for (int k = threadIdx.x; k < sh.total_verts; k += blockDim.x) smem[k] = sh.verts[k];
    __syncthreads();
    return smem;
}
NP_D void flush_stats(DevCounters* ctr, const PairStats& st, unsigned pair_tests, unsigned hits, unsigned overflow)
{
    // warp-level reduction, one atomic per warp per counter
    unsigned v[7] = { pair_tests, hits, st.gjk_steps, st.support_pairs, st.epa_iters, st.epa_tri_tests, overflow };
#pragma unroll
    for (int k = 0; k < 7; k++) {
        unsigned x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        v[k] = x;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&ctr->pair_tests, (unsigned long long)v[0]); atomicAdd(&ctr->hits, (unsigned long long)v[1]);
        atomicAdd(&ctr->gjk_steps, (unsigned long long)v[2]); atomicAdd(&ctr->support_pairs, (unsigned long long)v[3]);
        atomicAdd(&ctr->epa_iters, (unsigned long long)v[4]); atomicAdd(&ctr->epa_tri_tests, (unsigned long long)v[5]);
        atomicAdd(&ctr->epa_overflow, (unsigned long long)v[6]);
    }
}
// warp-aggregated work fetch: lanes with `need` get consecutive tickets from *counter
NP_D int fetch_ticket(int* counter, bool need)
{
    unsigned mask = __ballot_sync(0xffffffffu, need);
    if (mask == 0) return -1;
    int lane = threadIdx.x & 31, leader = __ffs(mask) - 1, base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return need ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}

// ---------------------------------------------------------------------------------------------------
// K3: GetCollisions (physics.cpp:116-145).  Work item = body i; the owning lane tests j = i+1, i+2, ...
// inside i's island and stops at the first hit.  Persistent grid, warp-aggregated ticket counter, lanes
// re-armed as soon as their body is resolved.
//   out: hit_j[i] = j or -1 ; contact[i] = (penDir.xyz, depth) ; cflags[i] = PairFlags
// ---------------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(NP_NARROW_THREADS) k_first_hit(StoreView s, ShapeView sh, int* __restrict__ ticket, int* __restrict__ hit_j,
                                                                float4* __restrict__ contact, int* __restrict__ cflags,
                                                                int* __restrict__ big_list, int* __restrict__ big_count, DevCounters* ctr)
{
    extern __shared__ float4 s_verts[];
    const float4* verts = stage_shapes(sh, s_verts, SMEM);
    PolyLocal<NP_FMAX_FAST> poly;
    PairMachine<PolyLocal<NP_FMAX_FAST>> m;
    PairGeom g;
    PairStats st = { 0, 0, 0, 0 };
    unsigned n_tests = 0, n_hits = 0, n_over = 0;
    int i = -1, j = 0, jend = 0;
    bool have = false, exhausted = false;
    const int* shape_of = s.ints + I_SHAPE * s.cap;
    const int* island_end = s.ints + I_ISLAND_END * s.cap;
    m.begin();
    for (;;) {
        int t = fetch_ticket(ticket, !have && !exhausted);
        if (!have && !exhausted) {
            if (t < s.n) {
                i = t; j = i + 1; jend = island_end[i];
                if (j < jend) {
                    load_xf(s, i, g.Ra, &g.ta);
                    int sa = shape_of[i]; g.va = verts + sh.off[sa]; g.na = sh.cnt[sa];
                    load_xf(s, j, g.Rb, &g.tb);
                    int sb = shape_of[j]; g.vb = verts + sh.off[sb]; g.nb = sh.cnt[sb];
                    m.begin(); have = true;
                } else { hit_j[i] = -1; cflags[i] = 0; contact[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
            } else exhausted = true;
        }
        if (!__any_sync(0xffffffffu, have)) { if (__all_sync(0xffffffffu, exhausted)) break; else continue; }
        if (have) {
            m.step(g, poly, st);
            if (m.done()) {
                n_tests++;
                if (m.flags & NP_PAIR_HIT) {
                    n_hits++;
                    hit_j[i] = j; cflags[i] = m.flags; contact[i] = make_float4(m.pen.x, m.pen.y, m.pen.z, m.depth);
                    if (m.flags & NP_PAIR_OVERFLOW) { n_over++; big_list[atomicAdd(big_count, 1)] = i; }
                    have = false;
                } else if (++j < jend) {
                    load_xf(s, j, g.Rb, &g.tb);
                    int sb = shape_of[j]; g.vb = verts + sh.off[sb]; g.nb = sh.cnt[sb];
                    m.begin();
                } else { hit_j[i] = -1; cflags[i] = 0; contact[i] = make_float4(0.f, 0.f, 0.f, 0.f); have = false; }
            }
        }
    }
    if (ctr) flush_stats(ctr, st, n_tests, n_hits, n_over);
}

// ---------------------------------------------------------------------------------------------------
// K3p: DetectCollision over an explicit pair list.  pose = position[3] rotation[3] (GetTransform is
// evaluated in-kernel); or, with xf16 != null, caller-provided row-major 4x4 transforms.
//   out: hit[p] = PairFlags ; contact[p] = (penDir.xyz, depth)
// ---------------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(NP_NARROW_THREADS) k_pairs(int n, ShapeView sh, const int* __restrict__ shape_a, const float* __restrict__ pose_a,
                                                            const int* __restrict__ shape_b, const float* __restrict__ pose_b,
                                                            const float* __restrict__ xf16_a, const float* __restrict__ xf16_b,
                                                            int* __restrict__ ticket, int* __restrict__ hit, float4* __restrict__ contact,
                                                            int* __restrict__ big_list, int* __restrict__ big_count, DevCounters* ctr)
{
    extern __shared__ float4 s_verts[];
    const float4* verts = stage_shapes(sh, s_verts, SMEM);
    PolyLocal<NP_FMAX_FAST> poly;
    PairMachine<PolyLocal<NP_FMAX_FAST>> m;
    PairGeom g;
    PairStats st = { 0, 0, 0, 0 };
    unsigned n_tests = 0, n_hits = 0, n_over = 0;
    int p = -1;
    bool have = false, exhausted = false;
    m.begin();
    for (;;) {
        int t = fetch_ticket(ticket, !have && !exhausted);
        if (!have && !exhausted) {
            if (t < n) {
                p = t;
                if (xf16_a) {
                    const float* a = xf16_a + 16 * (size_t)p; const float* b = xf16_b + 16 * (size_t)p;
#pragma unroll
                    for (int r = 0; r < 3; r++) {
#pragma unroll
                        for (int c = 0; c < 3; c++) { g.Ra[3 * r + c] = a[4 * r + c]; g.Rb[3 * r + c] = b[4 * r + c]; }
                    }
                    g.ta = mk(a[3], a[7], a[11]); g.tb = mk(b[3], b[7], b[11]);
                } else {
                    const float* a = pose_a + 6 * (size_t)p; const float* b = pose_b + 6 * (size_t)p;
                    make_transform(mk(a[0], a[1], a[2]), mk(a[3], a[4], a[5]), g.Ra, &g.ta);
                    make_transform(mk(b[0], b[1], b[2]), mk(b[3], b[4], b[5]), g.Rb, &g.tb);
                }
                int sa = shape_a[p], sb = shape_b[p];
                g.va = verts + sh.off[sa]; g.na = sh.cnt[sa];
                g.vb = verts + sh.off[sb]; g.nb = sh.cnt[sb];
                m.begin(); have = true;
            } else exhausted = true;
        }
        if (!__any_sync(0xffffffffu, have)) { if (__all_sync(0xffffffffu, exhausted)) break; else continue; }
        if (have) {
            m.step(g, poly, st);
            if (m.done()) {
                n_tests++;
                if (m.flags & NP_PAIR_HIT) n_hits++;
                hit[p] = m.flags; contact[p] = make_float4(m.pen.x, m.pen.y, m.pen.z, m.depth);
                if (m.flags & NP_PAIR_OVERFLOW) { n_over++; big_list[atomicAdd(big_count, 1)] = p; }
                have = false;
            }
        }
    }
    if (ctr) flush_stats(ctr, st, n_tests, n_hits, n_over);
}

// ---------------------------------------------------------------------------------------------------
// K4b: pairs whose EPA polytope outgrew the in-thread arena, re-run from the seed with a global-memory
// arena of `fbig` faces per slot.  Rare (p99 of the polytope is ~40 faces); thread t owns slot t.
// In world mode (hit_j != null) entry = body index i, partner = hit_j[i]; in pair mode entry = pair index.
// ---------------------------------------------------------------------------------------------------
struct BigArena { float* p; float* A; float* B; float* fn; uint32_t* fi; uint16_t* edge; int fbig, slots; };

__global__ void __launch_bounds__(64) k_epa_big(StoreView s, ShapeView sh, const int* __restrict__ big_list, const int* __restrict__ big_count,
                                               const int* __restrict__ hit_j, int n_pairs, const int* __restrict__ shape_a, const float* __restrict__ pose_a,
                                               const int* __restrict__ shape_b, const float* __restrict__ pose_b,
                                               const float* __restrict__ xf16_a, const float* __restrict__ xf16_b,
                                               int* __restrict__ out_flags, float4* __restrict__ contact, BigArena ar)
{
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    int count = *big_count;
    if (slot >= ar.slots) return;
    PolyGlobal poly;
    poly.p = ar.p + (size_t)slot * NP_EPA_MAX_VERTS * 3; poly.A = ar.A + (size_t)slot * NP_EPA_MAX_VERTS * 3; poly.B = ar.B + (size_t)slot * NP_EPA_MAX_VERTS * 3;
    poly.fn = ar.fn + (size_t)slot * ar.fbig * 3; poly.fi = ar.fi + (size_t)slot * ar.fbig; poly.edge = ar.edge + (size_t)slot * ar.fbig * 3;
    poly.fmax = ar.fbig; poly.emax = 3 * ar.fbig;
    for (int e = slot; e < count; e += ar.slots) {
        int id = big_list[e];
        PairGeom g;
        if (hit_j) {
            int i = id, j = hit_j[i];
            load_xf(s, i, g.Ra, &g.ta); load_xf(s, j, g.Rb, &g.tb);
            int sa = s.ints[I_SHAPE * s.cap + i], sb = s.ints[I_SHAPE * s.cap + j];
            g.va = sh.verts + sh.off[sa]; g.na = sh.cnt[sa]; g.vb = sh.verts + sh.off[sb]; g.nb = sh.cnt[sb];
        } else {
            if (xf16_a) {
                const float* a = xf16_a + 16 * (size_t)id; const float* b = xf16_b + 16 * (size_t)id;
                for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { g.Ra[3 * r + c] = a[4 * r + c]; g.Rb[3 * r + c] = b[4 * r + c]; }
                g.ta = mk(a[3], a[7], a[11]); g.tb = mk(b[3], b[7], b[11]);
            } else {
                const float* a = pose_a + 6 * (size_t)id; const float* b = pose_b + 6 * (size_t)id;
                make_transform(mk(a[0], a[1], a[2]), mk(a[3], a[4], a[5]), g.Ra, &g.ta);
                make_transform(mk(b[0], b[1], b[2]), mk(b[3], b[4], b[5]), g.Rb, &g.tb);
            }
            int sa = shape_a[id], sb = shape_b[id];
            g.va = sh.verts + sh.off[sa]; g.na = sh.cnt[sa]; g.vb = sh.verts + sh.off[sb]; g.nb = sh.cnt[sb];
        }
        PairMachine<PolyGlobal> m;
        PairStats st = { 0, 0, 0, 0 };
        m.begin();
        while (!m.done()) m.step(g, poly, st);
        out_flags[id] = m.flags;           // NP_PAIR_OVERFLOW survives only if even `fbig` faces were not enough
        contact[id] = make_float4(m.pen.x, m.pen.y, m.pen.z, m.depth);
    }
}

// ---------------------------------------------------------------------------------------------------
// K5: OnCollision (physics.cpp:93-114).  The reference walks the collision list in order (ascending a) and
// applies ApplyImpulseResponse to a, then b.  ApplyImpulseResponse touches only its own body plus the
// frozen CollisionData, so per body the sequence is: every collision (i, k) with hit_j[i] == k in ascending
// i (all of them precede k's own entry in the list because i < k), then k's own collision (k, hit_j[k]).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_incoming(int n, int cap, const int* __restrict__ hit_j, const int* __restrict__ ints, int* __restrict__ cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = hit_j[i];
    if (b >= 0 && ints[I_RESPONSE * cap + b] == 1) atomicAdd(&cnt[b], 1);
}
// exclusive scan of cnt[0..n) in three launches (per-block sums, scan of the sums, per-block scan + offset)
#define NP_SCAN_BLOCK 1024
__global__ void __launch_bounds__(NP_SCAN_BLOCK) k_scan_block_sums(int n, const int* __restrict__ in, int* __restrict__ sums)
{
    __shared__ int warp_sum[32];
    int i = blockIdx.x * NP_SCAN_BLOCK + threadIdx.x;
    int v = i < n ? in[i] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = warp_sum[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = w;
    }
}
__global__ void __launch_bounds__(1024) k_scan_sums(int nb, int* __restrict__ sums)
{
    // single block: serial over chunks of 1024, warp-shuffle scan inside a chunk
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? sums[i] : 0, x = v;
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane], z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
            warp_tot[lane] = z - w;
        }
        __syncthreads();
        int excl = carry + warp_tot[wid] + x - v;
        if (i < nb) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(NP_SCAN_BLOCK) k_scan_apply(int n, const int* __restrict__ in, const int* __restrict__ sums, int* __restrict__ out)
{
    __shared__ int warp_tot[32];
    int i = blockIdx.x * NP_SCAN_BLOCK + threadIdx.x;
    int v = i < n ? in[i] : 0, x = v;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int w = warp_tot[lane], z = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
        warp_tot[lane] = z - w;
    }
    __syncthreads();
    if (i < n) out[i] = sums[blockIdx.x] + warp_tot[wid] + x - v;
}
__global__ void __launch_bounds__(256) k_fill_incoming(int n, int cap, const int* __restrict__ hit_j, const int* __restrict__ ints,
                                                      const int* __restrict__ off, int* __restrict__ cursor, int* __restrict__ inc)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = hit_j[i];
    if (b >= 0 && ints[I_RESPONSE * cap + b] == 1) inc[off[b] + atomicAdd(&cursor[b], 1)] = i;
}

// Physics::ApplyImpulseResponse, physics.cpp:27-46
NP_D void impulse(V3& pos, V3& L, V3& H, float mass, float elast, const float* iinv, V3 r, float cdepth, V3 n)
{
    V3 v = mul(L, 1.0f / mass);
    V3 w = m3mul(iinv, H);
    V3 vp = add(v, cross(w, r));
    float i = -(1.0f + elast) * dot(vp, n);
    float k = (1.0f / mass) + dot(cross(m3mul(iinv, cross(r, n)), r), n);
    float j = i / k;
    L = add(L, mul(n, j));
    H = add(H, mul(cross(r, n), j));
    float adjust = cdepth * 1.01f;
    pos = add(pos, mul(n, adjust));
}
__global__ void __launch_bounds__(256) k_resolve(StoreView s, const int* __restrict__ hit_j, const float4* __restrict__ contact,
                                                const int* __restrict__ inc_off, const int* __restrict__ inc_cnt, int* __restrict__ inc)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.n) return;
    const int cap = s.cap;
    if (s.ints[I_RESPONSE * cap + k] != 1) return;
    int own = hit_j[k], ninc = inc_cnt[k];
    if (own < 0 && ninc == 0) return;
    float* st = s.state; const float* c = s.cst;
    V3 pos = mk(st[S_PX * cap + k], st[S_PY * cap + k], st[S_PZ * cap + k]);
    V3 L = mk(st[S_LX * cap + k], st[S_LY * cap + k], st[S_LZ * cap + k]);
    V3 H = mk(st[S_HX * cap + k], st[S_HY * cap + k], st[S_HZ * cap + k]);
    float mass = c[C_MASS * cap + k], elast = c[C_ELAST * cap + k];
    float iinv[9];
#pragma unroll
    for (int q = 0; q < 9; q++) iinv[q] = c[(C_IINV0 + q) * cap + k];
    // CollisionData::a_normal / b_normal are never written by the reference (collision_detection.cpp:533-543)
    const V3 a_normal = mk(0.0f, 0.0f, 0.0f), b_normal = mk(0.0f, 0.0f, 0.0f);
    int* seg = inc + inc_off[k];
    for (int a = 1; a < ninc; a++) {                     // the fill order is arbitrary: restore ascending i
        int key = seg[a], b = a - 1;
        while (b >= 0 && seg[b] > key) { seg[b + 1] = seg[b]; b--; }
        seg[b + 1] = key;
    }
    for (int a = 0; a < ninc; a++) {                     // k is `b` of collision (i, k): normal = a_normal (physics.cpp:109)
        float4 cd = contact[seg[a]];
        impulse(pos, L, H, mass, elast, iinv, mk(cd.x, cd.y, cd.z), cd.w, a_normal);
    }
    if (own >= 0) {                                      // k is `a` of its own collision: normal = b_normal (physics.cpp:99)
        float4 cd = contact[k];
        impulse(pos, L, H, mass, elast, iinv, mk(cd.x, cd.y, cd.z), cd.w, b_normal);
    }
    st[S_PX * cap + k] = pos.x; st[S_PY * cap + k] = pos.y; st[S_PZ * cap + k] = pos.z;
    st[S_LX * cap + k] = L.x; st[S_LY * cap + k] = L.y; st[S_LZ * cap + k] = L.z;
    st[S_HX * cap + k] = H.x; st[S_HY * cap + k] = H.y; st[S_HZ * cap + k] = H.z;
}

// stand-alone ApplyImpulseResponse / GetTransform on one body (API completeness; not on the step path)
__global__ void k_impulse_one(StoreView s, int k, V3 r, float depth, V3 n)
{
    const int cap = s.cap; float* st = s.state; const float* c = s.cst;
    V3 pos = mk(st[S_PX * cap + k], st[S_PY * cap + k], st[S_PZ * cap + k]);
    V3 L = mk(st[S_LX * cap + k], st[S_LY * cap + k], st[S_LZ * cap + k]);
    V3 H = mk(st[S_HX * cap + k], st[S_HY * cap + k], st[S_HZ * cap + k]);
    float iinv[9];
    for (int q = 0; q < 9; q++) iinv[q] = c[(C_IINV0 + q) * cap + k];
    impulse(pos, L, H, c[C_MASS * cap + k], c[C_ELAST * cap + k], iinv, r, depth, n);
    st[S_PX * cap + k] = pos.x; st[S_PY * cap + k] = pos.y; st[S_PZ * cap + k] = pos.z;
    st[S_LX * cap + k] = L.x; st[S_LY * cap + k] = L.y; st[S_LZ * cap + k] = L.z;
    st[S_HX * cap + k] = H.x; st[S_HY * cap + k] = H.y; st[S_HZ * cap + k] = H.z;
}
__global__ void k_transform_one(StoreView s, int i, float* out16)
{
    const int cap = s.cap;
    float R[9]; V3 t;
    make_transform(mk(s.state[S_PX * cap + i], s.state[S_PY * cap + i], s.state[S_PZ * cap + i]),
                   mk(s.state[S_RX * cap + i], s.state[S_RY * cap + i], s.state[S_RZ * cap + i]), R, &t);
    const float m[16] = { R[0], R[1], R[2], t.x, R[3], R[4], R[5], t.y, R[6], R[7], R[8], t.z, 0.0f, 0.0f, 0.0f, 1.0f };
    for (int q = 0; q < 16; q++) out16[q] = m[q];
}

// ---------------------------------------------------------------------------------------------------
// K6: replication snapshot.  One 36-byte CommandFrameObject per body (netphys_common/common.h:120-127):
// guid, pos[3], rot[4] (quaternion of the Euler XYZ rotation R = Rz*Ry*Rx), isEnabled.
// ---------------------------------------------------------------------------------------------------
struct SnapObj { uint32_t guid; float pos[3]; float rot[4]; uint32_t enabled; };
__global__ void __launch_bounds__(256) k_snapshot(StoreView s, uint32_t guid_base, SnapObj* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const int cap = s.cap;
    float hx = s.state[S_RX * cap + i] * 0.5f, hy = s.state[S_RY * cap + i] * 0.5f, hz = s.state[S_RZ * cap + i] * 0.5f;
    float sx, cx, sy, cy, sz, cz;
    np_sincosf(hx, &sx, &cx); np_sincosf(hy, &sy, &cy); np_sincosf(hz, &sz, &cz);
    SnapObj o;
    o.guid = guid_base + (uint32_t)i;
    o.pos[0] = s.state[S_PX * cap + i]; o.pos[1] = s.state[S_PY * cap + i]; o.pos[2] = s.state[S_PZ * cap + i];
    o.rot[0] = cx * cy * cz + sx * sy * sz;
    o.rot[1] = sx * cy * cz - cx * sy * sz;
    o.rot[2] = cx * sy * cz + sx * cy * sz;
    o.rot[3] = cx * cy * sz - sx * sy * cz;
    o.enabled = 1u;
    out[i] = o;
}

// ---------------------------------------------------------------------------------------------------
// self tests
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sincos_hash(uint32_t start, unsigned long long count, unsigned long long* hs, unsigned long long* hc)
{
    unsigned long long s = 0, c = 0;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < count; k += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t u = start + (uint32_t)k;
        float sn, cs;
        np_sincosf(__uint_as_float(u), &sn, &cs);
        uint32_t us = __float_as_uint(sn), uc = __float_as_uint(cs);
        if ((us & 0x7fffffffu) > 0x7f800000u) us = 0x7fc00000u;
        if ((uc & 0x7fffffffu) > 0x7f800000u) uc = 0x7fc00000u;
        unsigned long long wgt = 2ull * u + 1ull;
        s += wgt * us; c += wgt * uc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(hs, s); atomicAdd(hc, c); }
}
__global__ void k_selftest_closest(int kind, const float* pts, int* region, float* uvw)
{
    V3 a = mk(pts[0], pts[1], pts[2]), b = mk(pts[3], pts[4], pts[5]);
    float u = 0.f, v = 0.f, w = 0.f; int r;
    if (kind == 2) r = closest_line(a, b, &u);
    else if (kind == 3) r = closest_triangle(a, b, mk(pts[6], pts[7], pts[8]), &u, &v);
    else r = closest_tetra(a, b, mk(pts[6], pts[7], pts[8]), mk(pts[9], pts[10], pts[11]), &u, &v, &w);
    *region = r; uvw[0] = u; uvw[1] = v; uvw[2] = w;
}
__global__ void k_selftest_support(ShapeView sh, int shape, const float* dir, const float* xf16, float* out)
{
    float R[9]; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = xf16[4 * r + c];
    V3 t = mk(xf16[3], xf16[7], xf16[11]);
    V3 r = support(sh.verts + sh.off[shape], sh.cnt[shape], R, t, mk(dir[0], dir[1], dir[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

}  // namespace np
