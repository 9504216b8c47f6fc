// netphys_b200/csrc/np_world.cu -- host side of the C ABI declared in include/netphys_b200.h.
//
// A np_world keeps host mirrors of what the caller handed over (shapes, StaticPhysicsData, island ids)
// and a structure-of-arrays body store in HBM (np_kernels.cuh).  Dynamic state lives on the device from
// the first step on; getters copy it back.  Every device operation goes to the world's own stream.
// There is no CPU execution path: without a device np_world_create fails.
#include "../../include/netphys_b200.h"
#include "np_kernels.cuh"
#include "np_suptab.h"

#include <algorithm>
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

using namespace np;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                                      \
    do {                                                                                                              \
        cudaError_t e_ = (call);                                                                                      \
        if (e_ != cudaSuccess) return fail(NP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

struct HostShape { std::vector<float> xyz; bool dead = false; int tab_state = 0 /* 0 not built, 1 built, -1 none */; SupTable tab; };
struct HostBody { np_static_physics_data sd; int shape; int island; int enabled = 1; float pending_force[3] = { 0.f, 0.f, 0.f }; };
struct StepGraph { uint32_t dt_bits; void* snap; uint32_t guid; int kind; cudaGraphExec_t exec; unsigned long long launches; };   // one captured Physics_Update
struct CommandFrame { unsigned id; double time_ms; std::vector<np_snapshot_object> objects; };      // world_s.cpp:12-17

struct np_world {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    // host mirrors
    std::vector<HostShape> shapes;
    std::vector<HostBody> bodies;
    std::vector<np_body_state> host_state;   // authoritative until the first upload, refreshed on demand
    // body ids handed to the caller are HANDLES (creation order, never reused); the device store is dense over the live bodies (slots).
    // Identity until the first np_body_destroy.
    std::vector<int> slot_of_handle, handle_of_slot; bool any_destroyed = false;
    int island = 0;
    std::vector<CommandFrame> frames; unsigned frame_counter = 1;      // world_s.cpp:19-21
    // options
    int broadphase = NP_BROADPHASE_EXACT;
    float grid_cell = 0.0f;
    int counters_on = 0;
    int snapshot_each_step = 0; void* snap_target = nullptr; uint32_t guid_base = 0;
    void* d_flush = nullptr; unsigned long long launches = 0;
    // device store
    bool dirty_bodies = true, dirty_shapes = true, dev_state_valid = false, restore_ext = false;
    int cap = 0, dev_n = 0;
    float *d_state = nullptr, *d_cst = nullptr, *d_xf = nullptr;
    int* d_ints = nullptr;
    float4* d_verts = nullptr; int *d_soff = nullptr, *d_scnt = nullptr; int total_verts = 0;
    unsigned* d_tab = nullptr; float* d_tlim = nullptr; uint2* d_cells = nullptr; bool tables_on = true;   // support-direction tables (np_suptab.h)
    // per-step scratch
    int *d_hit_j = nullptr, *d_cflags = nullptr; float4* d_contact = nullptr;
    // sharded pair search (np_world_set_shard): this process searches bodies [shard_first, shard_first + shard_count); the per-body
    // collision records live in caller-provided buffers that the caller all-gathers between step_begin and step_end
    int shard_first = 0, shard_count = -1; int *x_hit_j = nullptr, *x_cflags = nullptr; float4* x_contact = nullptr;
    bool mid_step = false;
    int* d_zero = nullptr;            // [tickets and counts (32 ints), head[cap] of the incoming-collision lists]  -- one memset per step
    int *d_inc = nullptr /* next[] of the incoming-collision lists */, *d_sums = nullptr, *d_big_list = nullptr, *d_hit_list = nullptr;
    float4* d_simplex = nullptr;      // GJK result of every hit: A[4], B[4]
    float4* d_face0 = nullptr;        // SetupForEPA's four faces of the hits k_epa_first passes on (4 float4 per body)
    float4* d_seed = nullptr;         // per body: the two seed supports of DetectCollision (k_seed_supports), thread-per-body search only
    float4* d_wverts = nullptr; size_t wverts_cap = 0, wv_total = 0, wv_total_big = 0;   // per-step world-vertex cache (K1b), allocated on first use (ensure_cache)
    float* d_ext = nullptr;           // adapter extension, allocated on first use: [4][cap] external force xyz + enabled flag
    // grid broad phase (allocated on first use)
    GridView grid = {}; int grid_cap = 0; int* d_grid_ints = nullptr; float* d_grid_f = nullptr;
    bool swept_valid = false;         // NP_BROADPHASE_SWEPT: grid.aabb_prev holds the previous step's boxes of the current store
    int max_verts = 0;                // largest shape
    int g_gjk_env = 0, g_epa_env = 0; // NP_G_GJK / NP_G_EPA overrides (tuning)
    int order_on = 1; bool order_dirty = true;   // NP_ORDER=0: search bodies in index order (tuning)
    int fuse_on = 1;      // NP_FUSE=0: np_world_step_n runs plain steps (A/B)
    int halo = 0;             // > 0: partitioned sharded step (np_world_set_shard_halo): this rank integrates / responds for its own bodies only and sees `halo` bodies past its range
    int* d_halo_err = nullptr;
    char* d_one = nullptr;     // np_detect_collision's single-pair record
    float step_dt = 0.0f;
    int step_phase = 0;       // partitioned step in progress: 1 after np_world_step_integrate, 2 after np_world_step_search
    int epa2_tail32 = 1;      // NP_EPA2_TAIL=4: k_epa2's leftovers to the 4-lane kernel instead of a warp per pair
    int epa2_on = 1;          // NP_EPA2=0: without k_epa2 (small polytopes, thread per pair + block-wide AddFace)
    int tab_min_verts = 5;    // NP_TAB_MIN: shapes with fewer vertices are scanned (tuning)
    int gjk2_on = 1;          // NP_GJK2=0: k_gjk1 (every lane solves its own simplex) instead of k_gjk2
    int gjk1_old = 0, cbig_on = 0, epa_first = 1, seed_on = 0;   // NP_SEED=1: per-body table of DetectCollision's seed supports (measured slower: 4.01 vs 3.90 ms)   // NP_GJK1_OLD=1: round-1 register-simplex kernel; NP_CBIG=1: large-body vertex cache for the cooperative supports (measured slower: 9 GB of DRAM reads per step); NP_EPA_FIRST=0 (tuning, A/B)
    bool coop_scan = false, coop_tab = false; int coop_env = -1;   // large and small shapes mixed: warp-cooperative supports in the narrow kernels (NP_COOP=0/1 overrides)
    // NCCL (np_comm_*): communicator, its high-priority stream, the double-buffered snapshot gather, the sharded step's record arrays
    ncclComm_t comm = nullptr; int nranks = 1, rank = 0;
    cudaStream_t comm_stream = nullptr;
    char* d_gather[2] = { nullptr, nullptr }; int gather_slot = 0, gather_flip = 0, gather_last = -1; uint32_t gather_guid = 0;
    cudaEvent_t ev_step_done[2] = {}, ev_gather_begin[2] = {}, ev_gather_done[2] = {}; bool gather_pending[2] = { false, false };
    // snapshot gather by copy engines: every peer's two gather buffers mapped here through CUDA IPC (peer_gather[b][rank], null for this rank), see gather_setup_dma
    std::vector<char*> peer_gather[2]; bool gather_dma = false; int* d_bar = nullptr;
    std::vector<cudaStream_t> peer_stream; std::vector<cudaEvent_t> ev_peer; cudaEvent_t ev_fan = nullptr;      // one stream per peer: the copies to different peers run on different copy engines
    int* d_shard_hit = nullptr; int* d_shard_flags = nullptr; float4* d_shard_contact = nullptr; int shard_slot = 0;
    DevCounters* d_ctr = nullptr;
    SnapObj* d_snap = nullptr;
    BigArena arena = {};
    // pinned staging for host round trips
    float* h_stage = nullptr; size_t h_stage_bytes = 0;
    float* d_aos = nullptr; size_t d_aos_bytes = 0;   // device staging of np_body_state[n]
    // timing
    cudaEvent_t ev[6] = {};
    float last_ms[6] = {};
    bool timed = false;
    unsigned long long steps = 0;
    // CUDA graph of one step (np_world_step_n)
    std::vector<StepGraph> graphs; int graph_on = 1; bool capturing = false;
    int friction_on = 0;              // NP_OPT_FRICTION
    int phase_timing = 0;             // NP_OPT_PHASE_TIMING: per-phase events inside the step (they cost ~4 us each)      // np_world_step replays a captured step (NP_GRAPH=0: plain launches)
};

// ------------------------------------------------------------------------------------------------ helpers
static StoreView view(np_world* w)
{
    StoreView s; s.n = (int)w->bodies.size(); s.cap = w->cap; s.state = w->d_state; s.cst = w->d_cst; s.xf = w->d_xf; s.ints = w->d_ints; s.wverts = w->d_wverts; s.ext = w->d_ext; s.friction = w->friction_on;
    s.soff = w->d_soff; s.scnt = w->d_scnt; s.stab = w->tables_on ? w->d_tab : nullptr; s.stlim = w->d_tlim;       // (for the arming records written next to the transforms)
    return s;
}
// tables depend on the vertex list only: worlds of one process that create the same shapes (every test, every rank-local world) share them
static bool cached_support_table(const std::vector<float>& xyz, SupTable& out)
{
    static std::mutex mu; static std::map<std::vector<uint32_t>, std::shared_ptr<SupTable>> cache;     // keyed by the coordinates' bits
    std::lock_guard<std::mutex> lock(mu);
    std::vector<uint32_t> key(xyz.size());
    if (!xyz.empty()) memcpy(key.data(), xyz.data(), sizeof(float) * xyz.size());
    auto it = cache.find(key);
    if (it == cache.end()) {
        auto t = std::make_shared<SupTable>();
        if (!build_support_table(xyz, *t)) t.reset();
        if (cache.size() >= 4096) cache.clear();
        it = cache.emplace(std::move(key), t).first;
    }
    if (!it->second) return false;
    out = *it->second;
    return true;
}
static bool coop_on(const np_world* w) { return w->coop_env >= 0 ? w->coop_env != 0 : (w->tables_on ? w->coop_tab : w->coop_scan); }
static ShapeView sview(np_world* w)
{
    ShapeView s; s.verts = w->d_verts; s.off = w->d_soff; s.cnt = w->d_scnt; s.total_verts = w->total_verts;
    s.tv.tab = w->tables_on ? w->d_tab : nullptr; s.tv.tlim = w->d_tlim; s.tv.cells = w->d_cells;
    return s;
}
enum { Z_GJK_TICKET = 0, Z_EPA_TICKET, Z_MID_COUNT, Z_MID_TICKET, Z_BIG_COUNT, Z_BIG_TICKET, Z_HIT_COUNT = 8 /* NP_HIT_CLASSES counters */, Z_PAIR_CLASS = 12 /* 9 counts + 9 cursors */,
       Z_HIT2_COUNT = 32 /* NP_HIT_CLASSES counters: hits that k_epa_first passes on to k_epa */,
       Z_HIT3_COUNT = 36 /* NP_HIT_CLASSES counters (only [0] used): hits that k_epa2 passes on to k_epa */, Z_EPA2_TICKET = 40, NP_ZERO_INTS = 64 };
static int* hitj(np_world* w) { return w->x_hit_j ? w->x_hit_j : w->d_hit_j; }
static int* cflags(np_world* w) { return w->x_cflags ? w->x_cflags : w->d_cflags; }
static float4* contacts(np_world* w) { return w->x_contact ? w->x_contact : w->d_contact; }
static bool sharded(np_world* w) { return w->shard_count >= 0; }
static int* d_inc_cnt(np_world* w) { return w->d_zero + NP_ZERO_INTS; }

static void free_store(np_world* w)
{
    cudaFree(w->d_state); cudaFree(w->d_cst); cudaFree(w->d_xf); cudaFree(w->d_ints);
    cudaFree(w->d_hit_j); cudaFree(w->d_cflags); cudaFree(w->d_contact); cudaFree(w->d_zero);
    cudaFree(w->d_inc); cudaFree(w->d_sums); cudaFree(w->d_big_list); cudaFree(w->d_snap); cudaFree(w->d_hit_list); cudaFree(w->d_simplex); cudaFree(w->d_face0); w->d_face0 = nullptr; cudaFree(w->d_seed); w->d_seed = nullptr;
    w->d_state = w->d_cst = w->d_xf = nullptr; w->d_ints = nullptr; w->d_hit_j = w->d_cflags = nullptr; w->d_contact = nullptr;
    w->d_zero = w->d_inc = w->d_sums = w->d_big_list = w->d_hit_list = nullptr; w->d_snap = nullptr; w->d_simplex = nullptr;
    w->cap = 0;
}
static int ensure_stage(np_world* w, size_t bytes)
{
    if (bytes <= w->h_stage_bytes) return NP_OK;
    if (w->h_stage) cudaFreeHost(w->h_stage);
    w->h_stage = nullptr; w->h_stage_bytes = 0;
    CU(cudaMallocHost((void**)&w->h_stage, bytes));
    w->h_stage_bytes = bytes;
    return NP_OK;
}
static void drop_graph(np_world* w) { for (auto& g : w->graphs) cudaGraphExecDestroy(g.exec); w->graphs.clear(); }

// state [12][cap] on the device  <->  np_body_state[n] on the host: one contiguous copy through pinned staging + a device transposition
static int ensure_aos(np_world* w, size_t bytes)
{
    if (bytes <= w->d_aos_bytes) return NP_OK;
    cudaFree(w->d_aos); w->d_aos = nullptr; w->d_aos_bytes = 0;
    CU(cudaMalloc((void**)&w->d_aos, bytes));
    w->d_aos_bytes = bytes;
    return NP_OK;
}
// page-locked host memory (np_host_alloc, cudaHostRegister, cudaMallocHost) is copied to / from directly; pageable memory goes
// through the world's pinned staging buffer
static bool host_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
static int download_states(np_world* w, np_body_state* out, int n)
{
    if (n == 0) return NP_OK;
    const size_t bytes = sizeof(float) * S_COUNT * (size_t)n;
    const bool direct = host_pinned(out);
    int r = direct ? NP_OK : ensure_stage(w, bytes); if (r) return r;
    r = ensure_aos(w, bytes); if (r) return r;
    k_state_to_aos<<<(n + 255) / 256, 256, 0, w->stream>>>(n, w->cap, w->d_state, w->d_aos);
    CU(cudaMemcpyAsync(direct ? (void*)out : (void*)w->h_stage, w->d_aos, bytes, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    if (!direct) memcpy(out, w->h_stage, bytes);
    w->launches++;
    return NP_OK;
}
static int upload_states(np_world* w, const np_body_state* in, int n)
{
    if (n == 0) return NP_OK;
    const size_t bytes = sizeof(float) * S_COUNT * (size_t)n;
    const bool direct = host_pinned(in);
    int r = direct ? NP_OK : ensure_stage(w, bytes); if (r) return r;
    r = ensure_aos(w, bytes); if (r) return r;
    if (!direct) {
        CU(cudaStreamSynchronize(w->stream));                  // the staging buffer may still feed an earlier copy
        memcpy(w->h_stage, in, bytes);
    }
    CU(cudaMemcpyAsync(w->d_aos, direct ? (const void*)in : (const void*)w->h_stage, bytes, cudaMemcpyHostToDevice, w->stream));
    k_state_from_aos<<<(n + 255) / 256, 256, 0, w->stream>>>(n, w->cap, w->d_aos, w->d_state);
    CU(cudaGetLastError());
    w->launches++;
    return NP_OK;
}

// device arrays of a store of `c` bodies (w->cap is already set); on failure the caller frees whatever was allocated
static int alloc_store(np_world* w, size_t c)
{
    CU(cudaMalloc((void**)&w->d_state, sizeof(float) * S_COUNT * c)); CU(cudaMalloc((void**)&w->d_cst, sizeof(float) * C_COUNT * c));
    CU(cudaMalloc((void**)&w->d_xf, sizeof(float) * X_COUNT * c)); CU(cudaMalloc((void**)&w->d_ints, sizeof(int) * I_COUNT * c));
    CU(cudaMalloc((void**)&w->d_hit_j, sizeof(int) * c)); CU(cudaMalloc((void**)&w->d_cflags, sizeof(int) * c)); CU(cudaMalloc((void**)&w->d_contact, sizeof(float4) * c));
    CU(cudaMalloc((void**)&w->d_zero, sizeof(int) * (NP_ZERO_INTS + c))); CU(cudaMalloc((void**)&w->d_inc, sizeof(int) * c));
    CU(cudaMalloc((void**)&w->d_sums, sizeof(int) * (4 * c / NP_SCAN_BLOCK + 2))); CU(cudaMalloc((void**)&w->d_big_list, sizeof(int) * c));
    CU(cudaMalloc((void**)&w->d_snap, sizeof(SnapObj) * c));
    CU(cudaMalloc((void**)&w->d_hit_list, sizeof(int) * 2 * NP_HIT_CLASSES * c)); CU(cudaMalloc((void**)&w->d_simplex, sizeof(float4) * 6 * c)); CU(cudaMalloc((void**)&w->d_face0, sizeof(float4) * 4 * c));      // hit lists of k_gjk, then those k_epa_first passes on
    CU(cudaMalloc((void**)&w->d_seed, sizeof(float4) * 2 * c));
    CU(cudaMemsetAsync(w->d_state, 0, sizeof(float) * S_COUNT * c, w->stream));
    CU(cudaMemsetAsync(w->d_hit_j, 0xff, sizeof(int) * c, w->stream)); CU(cudaMemsetAsync(w->d_cflags, 0, sizeof(int) * c, w->stream));
    CU(cudaMemsetAsync(w->d_contact, 0, sizeof(float4) * c, w->stream));
    return NP_OK;
}
static int ensure_grid(np_world* w);
// search order of the range this process searches (the whole world unless sharded): grouped by shape, stable; see NarrowIn::ordered
static int upload_order(np_world* w)
{
    const int n = (int)w->bodies.size();
    w->order_dirty = false;
    if (n == 0 || !w->d_ints) return NP_OK;
    int f = w->shard_count >= 0 ? std::min(w->shard_first, n) : 0, c = w->shard_count >= 0 ? std::min(w->shard_count, n - f) : n;
    std::vector<int> ord(n);
    for (int i = 0; i < n; i++) ord[i] = i;
    // secondary key: the shape of the first partner the exact search will test (body i + 1 of the same island), so that both sides match
    const bool pairkey = w->broadphase == NP_BROADPHASE_EXACT;
    auto key = [&](int a) {
        long long k = (long long)w->bodies[a].shape << 32;
        if (pairkey && a + 1 < n && w->bodies[a + 1].island == w->bodies[a].island) k |= (unsigned)(w->bodies[a + 1].shape + 1);
        return k;
    };
    std::stable_sort(ord.begin() + f, ord.begin() + f + c, [&](int a, int b) { return key(a) < key(b); });
    CU(cudaMemcpyAsync(w->d_ints + (size_t)I_ORDER * w->cap, ord.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return NP_OK;
}
// adapter extension storage: forces start at zero, the enabled row comes from the host mirror
static int ensure_ext(np_world* w, const std::vector<float>* keep_forces = nullptr, int keep_cap = 0)
{
    if (w->d_ext) return NP_OK;
    drop_graph(w);
    const size_t cap = (size_t)w->cap;
    std::vector<float> h(4 * cap, 0.0f);
    for (size_t i = 0; i < cap; i++) h[3 * cap + i] = (i < w->bodies.size() && !w->bodies[i].enabled) ? 0.0f : 1.0f;
    if (keep_forces) for (int r = 0; r < 3; r++) for (int i = 0; i < keep_cap && (size_t)i < cap; i++) h[r * cap + i] = (*keep_forces)[(size_t)r * keep_cap + i];
    CU(cudaMalloc((void**)&w->d_ext, sizeof(float) * 4 * cap));
    CU(cudaMemcpyAsync(w->d_ext, h.data(), sizeof(float) * 4 * cap, cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return NP_OK;
}
static int ensure_device(np_world* w)
{
    CU(cudaSetDevice(w->device));
    const int n = (int)w->bodies.size();
    if (w->dirty_shapes) {
        drop_graph(w);
        cudaFree(w->d_verts); cudaFree(w->d_soff); cudaFree(w->d_scnt); w->d_verts = nullptr; w->d_soff = w->d_scnt = nullptr;
        std::vector<float4> v; std::vector<int> off, cnt;
        for (auto& s : w->shapes) {
            off.push_back((int)v.size()); cnt.push_back((int)s.xyz.size() / 3);
            for (size_t k = 0; k + 2 < s.xyz.size(); k += 3) v.push_back(make_float4(s.xyz[k], s.xyz[k + 1], s.xyz[k + 2], 0.0f));
        }
        w->total_verts = (int)v.size();
        w->max_verts = 0; for (int c_ : cnt) w->max_verts = std::max(w->max_verts, c_);
        size_t nv = std::max<size_t>(v.size(), 1), ns = std::max<size_t>(off.size(), 1);
        CU(cudaMalloc((void**)&w->d_verts, sizeof(float4) * nv)); CU(cudaMalloc((void**)&w->d_soff, sizeof(int) * ns)); CU(cudaMalloc((void**)&w->d_scnt, sizeof(int) * ns));
        if (!v.empty()) CU(cudaMemcpyAsync(w->d_verts, v.data(), sizeof(float4) * v.size(), cudaMemcpyHostToDevice, w->stream));
        if (!off.empty()) {
            CU(cudaMemcpyAsync(w->d_soff, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, w->stream));
            CU(cudaMemcpyAsync(w->d_scnt, cnt.data(), sizeof(int) * cnt.size(), cudaMemcpyHostToDevice, w->stream));
        }
        // support-direction tables: built once per shape, concatenated
        cudaFree(w->d_tab); cudaFree(w->d_tlim); cudaFree(w->d_cells); w->d_tab = nullptr; w->d_tlim = nullptr; w->d_cells = nullptr;
        std::vector<unsigned> tab; std::vector<float> tlim; std::vector<uint64_t> cells;
        for (auto& s : w->shapes) {
            if (s.tab_state == 0) s.tab_state = (!s.dead && cached_support_table(s.xyz, s.tab)) ? 1 : -1;
            if (s.tab_state == 1 && (int)(s.xyz.size() / 3) >= w->tab_min_verts && cells.size() + s.tab.cells.size() < (1u << 28)) {
                tab.push_back((unsigned)cells.size() | ((unsigned)s.tab.log_n << 28)); tlim.push_back(s.tab.tlim);
                cells.insert(cells.end(), s.tab.cells.begin(), s.tab.cells.end());
            } else { tab.push_back(NP_TAB_NONE); tlim.push_back(0.0f); }
        }
        CU(cudaMalloc((void**)&w->d_tab, sizeof(unsigned) * ns)); CU(cudaMalloc((void**)&w->d_tlim, sizeof(float) * ns)); CU(cudaMalloc((void**)&w->d_cells, sizeof(uint64_t) * std::max<size_t>(cells.size(), 1)));
        if (!tab.empty()) {
            CU(cudaMemcpyAsync(w->d_tab, tab.data(), sizeof(unsigned) * tab.size(), cudaMemcpyHostToDevice, w->stream));
            CU(cudaMemcpyAsync(w->d_tlim, tlim.data(), sizeof(float) * tlim.size(), cudaMemcpyHostToDevice, w->stream));
        }
        if (!cells.empty()) CU(cudaMemcpyAsync(w->d_cells, cells.data(), sizeof(uint64_t) * cells.size(), cudaMemcpyHostToDevice, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        w->dirty_shapes = false;
    }
    if (!w->dirty_bodies) return w->order_dirty ? upload_order(w) : NP_OK;
    drop_graph(w);
    if (w->dev_state_valid && w->dev_n > 0) {   // bodies were added after stepping: keep what the device has
        int r = download_states(w, w->host_state.data(), w->dev_n); if (r) return r;
    }
    std::vector<float> old_ext; int old_cap = 0;
    if (n > w->cap && w->d_ext) {               // the store is about to be reallocated: carry pending forces over
        old_cap = w->cap; old_ext.resize(4 * (size_t)old_cap);
        CU(cudaMemcpyAsync(old_ext.data(), w->d_ext, sizeof(float) * 4 * (size_t)old_cap, cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        cudaFree(w->d_ext); w->d_ext = nullptr;
    }
    if (n > w->cap) {
        free_store(w);
        int cap = std::max(1024, (n + 1023) / 1024 * 1024);
        if (n > 4096) cap = std::max(cap, (int)std::min<long long>((long long)n * 5 / 4, 1LL << 30) / 1024 * 1024 + 1024);
        w->cap = cap;
        size_t c = (size_t)cap;
        int ra = alloc_store(w, c);
        if (ra) { free_store(w); w->dev_state_valid = false; w->dev_n = 0; return ra; }      // the bodies stay on the host; a later call retries
    }
    if (old_cap) { int re = ensure_ext(w, &old_ext, old_cap); if (re) return re; }
    else if (w->d_ext) {                        // bodies were added: their enabled flags
        std::vector<float> en((size_t)w->cap, 1.0f);
        for (int i = 0; i < n; i++) en[i] = w->bodies[i].enabled ? 1.0f : 0.0f;
        CU(cudaMemcpyAsync(w->d_ext + 3 * (size_t)w->cap, en.data(), sizeof(float) * (size_t)w->cap, cudaMemcpyHostToDevice, w->stream));
        CU(cudaStreamSynchronize(w->stream));
    }
    const size_t cap = (size_t)w->cap;
    // constants + ints, SoA
    std::vector<float> cst(C_COUNT * cap, 0.0f); std::vector<int> ints(I_COUNT * cap, 0);
    std::vector<int> island_end(n, n);
    for (int i = n - 1, end = n; i >= 0; i--) { if (i + 1 < n && w->bodies[i + 1].island != w->bodies[i].island) end = i + 1; island_end[i] = end; }
    for (int i = 0; i < n; i++) {
        const np_static_physics_data& sd = w->bodies[i].sd;
        cst[C_GX * cap + i] = sd.gravity[0]; cst[C_GY * cap + i] = sd.gravity[1]; cst[C_GZ * cap + i] = sd.gravity[2];
        cst[C_MASS * cap + i] = sd.mass; cst[C_ELAST * cap + i] = sd.elasticity; cst[C_SFRIC * cap + i] = sd.static_friction_coeff;
        for (int k = 0; k < 9; k++) cst[(C_IINV0 + k) * cap + i] = sd.inverse_inertia_tensor[k];
        ints[I_RESPONSE * cap + i] = sd.collision_response; ints[I_SHAPE * cap + i] = w->bodies[i].shape; ints[I_ISLAND_END * cap + i] = island_end[i];
        ints[I_HANDLE * cap + i] = w->handle_of_slot[i];
    }
    {   // large and small shapes side by side (and the large ones a minority): the thread-per-pair kernels serve large supports warp-wide
        // (a large shape with a support-direction table is answered per lane from its table: it does not count while tables are on)
        long long big = 0, big_untabled = 0;
        for (int i = 0; i < n; i++) {
            const HostShape& hs = w->shapes[w->bodies[i].shape];
            if (hs.xyz.size() / 3 >= NP_COOP_MIN) { big++; if (hs.tab_state != 1 || (int)(hs.xyz.size() / 3) < w->tab_min_verts) big_untabled++; }
        }
        w->coop_scan = big > 0 && big * 100 <= (long long)n * 35;
        w->coop_tab = big_untabled > 0 && big_untabled * 100 <= (long long)n * 35;
    }
    size_t wv_total = 0;
    for (int i = 0; i < n; i++) {
        if (wv_total > 0x7fffffffu) return fail(NP_ERR_CAPACITY, "world-vertex cache exceeds 2^31 vertices");
        ints[I_WVOFF * cap + i] = (int)wv_total; wv_total += w->shapes[w->bodies[i].shape].xyz.size() / 3;
    }
    size_t wv_big = 0;
    for (int i = 0; i < n; i++) {
        const size_t nv = w->shapes[w->bodies[i].shape].xyz.size() / 3;
        if (nv >= NP_COOP_MIN) { if (wv_big > 0x7fffffffu) return fail(NP_ERR_CAPACITY, "world-vertex cache exceeds 2^31 vertices"); ints[I_WVOFF_BIG * cap + i] = (int)wv_big; wv_big += nv; }
        else ints[I_WVOFF_BIG * cap + i] = -1;
    }
    w->wv_total = wv_total; w->wv_total_big = wv_big;               // the cache itself is allocated by ensure_cache, only for steps whose kernels read it
    CU(cudaMemcpyAsync(w->d_cst, cst.data(), sizeof(float) * C_COUNT * cap, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(w->d_ints, ints.data(), sizeof(int) * I_COUNT * cap, cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    int r = upload_states(w, w->host_state.data(), n); if (r) return r;
    if (n > 0) { k_transforms<<<(n + 255) / 256, 256, 0, w->stream>>>(view(w)); CU(cudaGetLastError()); }
    w->dev_n = n; w->dev_state_valid = true; w->dirty_bodies = false;
    if (w->restore_ext && !w->d_ext) {                       // np_body_destroy dropped the extension arrays: pending forces and enabled flags come back from the host mirror
        std::vector<float> f(3 * cap, 0.0f);
        for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) f[(size_t)c * cap + i] = w->bodies[i].pending_force[c];
        int re = ensure_ext(w, &f, (int)cap); if (re) return re;
        for (auto& b : w->bodies) b.pending_force[0] = b.pending_force[1] = b.pending_force[2] = 0.0f;
    }
    w->restore_ext = false;
    { int ro = upload_order(w); if (ro) return ro; }
    if (w->broadphase != NP_BROADPHASE_EXACT) { int rg = ensure_grid(w); if (rg) return rg; }
    w->swept_valid = false;                                  // the store was rebuilt: no previous boxes
    return NP_OK;
}
static int ensure_arena(np_world* w)
{
    if (w->arena.base) return NP_OK;
    w->arena.slots = 128;             // one slot (an 8192-face polytope, 263 KB) per warp of the TIER 2 launch: polytopes beyond 1024 faces are pathological
    CU(cudaMalloc((void**)&w->arena.base, sizeof(float) * (size_t)poly_words_fast(NP_FBIG) * w->arena.slots));
    return NP_OK;
}
static bool shapes_in_smem(np_world* w) { return w->total_verts <= NP_SMEM_VERTS; }
enum { CACHE_NONE = 0, CACHE_ALL = 1, CACHE_BIG = 2 };
static int cache_mode(np_world* w, long long items);

// ---- narrow-phase launch plumbing -------------------------------------------------------------------------------
static size_t shapes_smem(np_world* w) { return shapes_in_smem(w) ? sizeof(float4) * (size_t)w->total_verts : 0; }
template <int G> constexpr int epa_fmax() { return G == 4 ? 32 : (G == 32 ? 256 : 64); }
template <int G> static size_t epa_poly_smem() { return G > 1 ? sizeof(float) * (size_t)(G == 32 ? poly_words_fast(epa_fmax<G>()) : poly_words(epa_fmax<G>())) * (NP_NARROW_THREADS / G) : 0; }

template <typename K>
static int persistent_blocks(np_world* w, K kernel, size_t smem, long long items, int G)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NP_NARROW_THREADS, smem) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    long long want = (items * G + NP_NARROW_THREADS - 1) / NP_NARROW_THREADS;
    return (int)std::max<long long>(1, std::min<long long>((long long)std::max(1, per_sm) * w->sm_count, want));
}
// group size: few items -> wide groups (latency, whole-GPU occupancy); many items -> narrow groups (throughput)
static int pick_group(np_world* w, long long items, int env, bool epa)
{
    if (env >= 1 && env <= 32) return env;
    // measured on B200 (tools/sweep2.sh, tools/sweep_world.sh): below a few hundred thousand items the step is bound by the
    // latency of its longest chains -> wide groups; above, by instruction throughput -> narrow groups
    // (round 2, thread-per-body k_gjk1 + k_epa_first: 50/50 box/icosphere worlds, ms per step wide vs narrow groups -- 131 k bodies 1.36 vs 1.62,
    // 262 k bodies 2.57 vs 2.05: the crossover moved from ~300 k to ~180 k items)
    // (with support-direction tables, k_gjk2 and k_epa2 the thread-per-item path took over everything above ~45 k items: 50/50 box/icosphere
    // worlds, narrow phase in ms, wide groups vs thread per item -- 20 k 0.29 / 0.46, 40 k 0.50 / 0.54, 80 k 0.82 / 0.55, 131 k 1.29 / 0.63)
    if (items <= 45000) return 32;                          // a warp per body / per hit: one wave, bound by its longest chains
    return epa ? (w->max_verts <= 16 ? 1 : 4) : 1;          // small shapes: thread-per-hit EPA wins (8.1e8 vs 7.6e8 box pairs/s)
}
struct NarrowBufs { int* zero; int* hit_j; int* flags; float4* contact; float4* simplex; int* hit_list; int* big_list; int* hit_list2; float4* face0; };

static const size_t kSimplexSmem = sizeof(float) * NP_SX_ROWS * NP_NARROW_THREADS;      // k_gjk1: the block's simplices
template <int G, bool WORLD>
static int launch_gjk(np_world* w, const NarrowIn& in, const NarrowBufs& b, DevCounters* ctr)
{
    size_t smem = (WORLD && G >= 8) ? 0 : shapes_smem(w);
    auto kern = k_gjk<G, WORLD>;
    if constexpr (G == 1) {
        if (!w->gjk1_old) {
            smem += kSimplexSmem;
            kern = k_gjk1<WORLD, false, false>;
            if (coop_on(w)) kern = k_gjk1<WORLD, true, false>;
            if constexpr (WORLD) {
                if (coop_on(w) && cache_mode(w, in.n) == CACHE_BIG) kern = k_gjk1<true, true, true>;
                else if (in.seed) kern = coop_on(w) ? k_gjk1<true, true, false, true> : k_gjk1<true, false, false, true>;
            }
            if (w->gjk2_on && !coop_on(w) && !in.seed) {      // solves regrouped by simplex size; its own block size
                smem = shapes_smem(w) + sizeof(float) * NP_SX_ROWS * NP_GJK2_THREADS + sizeof(int) * NP_GJK2_EXTRA_WORDS;
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gjk2<WORLD>, NP_GJK2_THREADS, smem) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
                const int gb = (int)std::max<long long>(1, std::min<long long>((long long)std::max(1, per_sm) * w->sm_count, ((long long)in.n + NP_GJK2_THREADS - 1) / NP_GJK2_THREADS));
                k_gjk2<WORLD><<<gb, NP_GJK2_THREADS, smem, w->stream>>>(in, sview(w), shapes_in_smem(w) ? 1 : 0, b.zero + Z_GJK_TICKET, b.hit_j, b.flags, b.contact, b.simplex,
                                                                          b.hit_list, b.zero + Z_HIT_COUNT, ctr);
                w->launches++;
                return NP_OK;
            }
        } else if constexpr (WORLD) { if (coop_on(w)) kern = k_gjk<G, WORLD, true>; }
    }
    int blocks = persistent_blocks(w, kern, smem, in.n, G);
    kern<<<blocks, NP_NARROW_THREADS, smem, w->stream>>>(in, sview(w), shapes_in_smem(w) ? 1 : 0, b.zero + Z_GJK_TICKET, b.hit_j, b.flags, b.contact, b.simplex,
                                                        b.hit_list, b.zero + Z_HIT_COUNT, ctr);
    w->launches++;
    return NP_OK;
}
// EPA overflow tiers (one warp per pair) over big_list; both exit at once when their list is empty.  hit_list is free again by then.
// TC: they read the world-vertex cache -- world mode, when this step built it (wants_cache)
template <bool WORLD, bool TC>
static int launch_epa_tiers(np_world* w, const NarrowIn& in, const NarrowBufs& b, DevCounters* ctr)
{
    const int sm_shapes = shapes_in_smem(w) ? 1 : 0;
    const size_t sh_bytes = TC ? 0 : shapes_smem(w);
    auto mid = k_epa<32, NP_FMID, 1, NP_MID_THREADS, WORLD, TC>;
    size_t msmem = sh_bytes + sizeof(float) * (size_t)poly_words_fast(NP_FMID) * (NP_MID_THREADS / 32);
    int mblocks = (int)std::min<long long>(w->sm_count * 2, std::max<long long>(1, ((long long)in.n * 32 + NP_MID_THREADS - 1) / NP_MID_THREADS));
    mid<<<mblocks, NP_MID_THREADS, msmem, w->stream>>>(in, sview(w), sm_shapes, b.big_list, b.zero + Z_MID_COUNT, b.zero + Z_MID_TICKET, b.hit_j, b.simplex,
                                                       b.flags, b.contact, b.hit_list, b.zero + Z_BIG_COUNT, w->arena, ctr);
    auto big = k_epa<32, 64, 2, NP_MID_THREADS, WORLD, TC>;
    int bblocks = std::min(w->arena.slots / (NP_MID_THREADS / 32), mblocks);
    big<<<bblocks, NP_MID_THREADS, sh_bytes, w->stream>>>(in, sview(w), sm_shapes, b.hit_list, b.zero + Z_BIG_COUNT, b.zero + Z_BIG_TICKET, b.hit_j, b.simplex,
                                                          b.flags, b.contact, nullptr, nullptr, w->arena, ctr);
    w->launches += 2;
    return NP_OK;
}
template <int G, bool WORLD, bool TC>
static int launch_epa(np_world* w, const NarrowIn& in, const NarrowBufs& b, DevCounters* ctr)
{
    const int sm_shapes = shapes_in_smem(w) ? 1 : 0;
    auto kern = k_epa<G, epa_fmax<G>(), 0, NP_NARROW_THREADS, WORLD>;
    size_t smem = ((WORLD && G >= 8) ? 0 : shapes_smem(w)) + epa_poly_smem<G>();
    int stage = sm_shapes;
    if (smem > epa_poly_smem<G>() && !getenv("NP_EPA_STAGE")) {          // the shape table is staged next to the polytopes only while it costs no resident block
        int with = 0, without = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&with, kern, NP_NARROW_THREADS, smem) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&without, kern, NP_NARROW_THREADS, epa_poly_smem<G>()) == cudaSuccess && without > with) {
            stage = 0; smem = epa_poly_smem<G>();
        } else cudaGetLastError();
    }
    int blocks = persistent_blocks(w, kern, smem, in.n, G);
    const int* list = b.hit_list; const int* counts = b.zero + Z_HIT_COUNT;
    if (G < 32 && w->epa_first) {            // throughput regime: the hits that converge before their first support never reach the group kernel
        const int fb = (int)std::min<long long>((long long)w->sm_count * 8, ((long long)in.n + 255) / 256);
        k_epa_first<WORLD><<<std::max(1, fb), 256, 0, w->stream>>>(in.n, b.hit_list, b.zero + Z_HIT_COUNT, b.simplex, b.hit_j, b.flags, b.contact, b.hit_list2, b.zero + Z_HIT2_COUNT, ctr,
                                                                   w->epa2_on ? b.face0 : nullptr);
        list = b.hit_list2; counts = b.zero + Z_HIT2_COUNT;
        w->launches += 1;
        if (w->epa2_on) {                    // ... and the polytopes that stay small are finished by k_epa2; what it passes on arrives as one plain list (class 0)
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_epa2<WORLD>, NP_E2_THREADS, NP_E2_SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
            const int eb = (int)std::max<long long>(1, std::min<long long>((long long)std::max(1, per_sm) * w->sm_count, ((long long)in.n + NP_E2_SLOTS - 1) / NP_E2_SLOTS));
            k_epa2<WORLD><<<eb, NP_E2_THREADS, NP_E2_SMEM_BYTES, w->stream>>>(in, sview(w), list, counts, b.zero + Z_EPA2_TICKET, b.hit_j, b.simplex, b.flags, b.contact,
                                                                               b.hit_list, b.zero + Z_HIT3_COUNT, ctr, b.face0);
            list = b.hit_list; counts = b.zero + Z_HIT3_COUNT;
            w->launches += 1;
            if (w->epa2_tail32) {
                // what is left is under one percent of the hits (the horizons whose list semantics matter, the few polytopes that grow) and
                // each runs long: a latency problem again, so a warp per pair -- the 256-face tier with the O(1) horizon, on the fly transforms
                auto tail = k_epa<32, epa_fmax<32>(), 0, NP_NARROW_THREADS, WORLD, false>;
                const size_t tsmem = epa_poly_smem<32>();
                const int tb = persistent_blocks(w, tail, tsmem, in.n, 32);
                tail<<<tb, NP_NARROW_THREADS, tsmem, w->stream>>>(in, sview(w), 0, list, counts, b.zero + Z_EPA_TICKET, b.hit_j, b.simplex,
                                                                   b.flags, b.contact, b.big_list, b.zero + Z_MID_COUNT, w->arena, ctr);
                w->launches += 1;
                return launch_epa_tiers<WORLD, TC>(w, in, b, ctr);
            }
        }
    }
    kern<<<blocks, NP_NARROW_THREADS, smem, w->stream>>>(in, sview(w), stage, list, counts, b.zero + Z_EPA_TICKET, b.hit_j, b.simplex,
                                                        b.flags, b.contact, b.big_list, b.zero + Z_MID_COUNT, w->arena, ctr);
    w->launches += 1;
    return launch_epa_tiers<WORLD, TC>(w, in, b, ctr);
}
template <bool WORLD>
static void setup_kernels_t()
{
    // EPA keeps one polytope per group in shared memory next to the staged shape table: up to ~105 KB per block
    const int bytes = 112 * 1024;
    cudaFuncSetAttribute(k_epa<4, epa_fmax<4>(), 0, NP_NARROW_THREADS, WORLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_epa<8, epa_fmax<8>(), 0, NP_NARROW_THREADS, WORLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_epa<32, epa_fmax<32>(), 0, NP_NARROW_THREADS, WORLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_epa<32, epa_fmax<32>(), 0, NP_NARROW_THREADS, WORLD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_epa<32, NP_FMID, 1, NP_MID_THREADS, WORLD, WORLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_epa<32, NP_FMID, 1, NP_MID_THREADS, WORLD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
static void setup_kernels()
{
    setup_kernels_t<true>(); setup_kernels_t<false>();
    const int bytes = 64 * 1024;               // staged shape table (<= 32 KB) + the block's simplices (18 KB)
    cudaFuncSetAttribute(k_gjk1<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_gjk1<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_gjk1<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_gjk1<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_gjk1<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_seed_supports, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_gjk1<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_gjk1<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_epa2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NP_E2_SMEM_BYTES);
    cudaFuncSetAttribute(k_epa2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NP_E2_SMEM_BYTES);
    cudaFuncSetAttribute(k_gjk2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + sizeof(int) * NP_GJK2_EXTRA_WORDS);
    cudaFuncSetAttribute(k_gjk2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + sizeof(int) * NP_GJK2_EXTRA_WORDS);
}
// supported group widths: GJK {1, 8, 32}, EPA {1, 4, 8, 32} (other requests round to the nearest supported one)
template <bool WORLD>
static int launch_narrow_t(np_world* w, const NarrowIn& in, const NarrowBufs& b, DevCounters* ctr)
{
    int r;
    const int gg = pick_group(w, in.n, w->g_gjk_env, false), ge = pick_group(w, in.n, w->g_epa_env, true);
    if (gg <= 1) r = launch_gjk<1, WORLD>(w, in, b, ctr);
    else if (gg <= 8) r = launch_gjk<8, WORLD>(w, in, b, ctr);
    else r = launch_gjk<32, WORLD>(w, in, b, ctr);
    if (r) return r;
    // narrow main kernels (G < 8) transform on the fly; when GJK is narrow too nothing needs the cache and the step does not build it
    if (ge <= 1) r = (WORLD && gg >= 8) ? launch_epa<1, WORLD, WORLD>(w, in, b, ctr) : launch_epa<1, WORLD, false>(w, in, b, ctr);
    else if (ge <= 4) r = (WORLD && gg >= 8) ? launch_epa<4, WORLD, WORLD>(w, in, b, ctr) : launch_epa<4, WORLD, false>(w, in, b, ctr);
    else if (ge <= 8) r = launch_epa<8, WORLD, WORLD>(w, in, b, ctr);
    else r = launch_epa<32, WORLD, WORLD>(w, in, b, ctr);
    return r;
}
// which world-vertex cache does a world-mode search over `items` bodies read?  (same decisions as launch_narrow_t / launch_gjk)
//   ALL: wide groups read every body's vertices;  BIG: the thread-per-body kernel's cooperative supports read the large bodies' only
static int cache_mode(np_world* w, long long items)
{
    if (items <= 0) return CACHE_NONE;
    const int gg = pick_group(w, items, w->g_gjk_env, false), ge = pick_group(w, items, w->g_epa_env, true);
    if (gg >= 8 || ge >= 8) return CACHE_ALL;
    if (gg <= 1 && coop_on(w) && w->cbig_on && !w->gjk1_old) return CACHE_BIG;
    return CACHE_NONE;
}
// the world-vertex cache is allocated the first time a step's kernels read it (never inside a stream capture); on failure nothing changes
static int ensure_cache(np_world* w)
{
    const int n = (int)w->bodies.size();
    const long long count = w->shard_count >= 0 ? w->shard_count : n;
    const int mode = cache_mode(w, count);
    const size_t need = mode == CACHE_ALL ? w->wv_total : (mode == CACHE_BIG ? w->wv_total_big : 0);
    if (need <= w->wverts_cap) return NP_OK;
    drop_graph(w);
    CU(cudaStreamSynchronize(w->stream));
    cudaFree(w->d_wverts); w->d_wverts = nullptr; w->wverts_cap = 0;
    const size_t want = need + need / 4 + 1024;
    CU(cudaMalloc((void**)&w->d_wverts, sizeof(float4) * want));
    w->wverts_cap = want;
    return NP_OK;
}
static int launch_narrow(np_world* w, const NarrowIn& in, const NarrowBufs& b, DevCounters* ctr)
{
    return in.world ? launch_narrow_t<true>(w, in, b, ctr) : launch_narrow_t<false>(w, in, b, ctr);
}

static void enqueue_scan(np_world* w, int n, const int* in, int* out, const int* limit = nullptr)
{
    const int sb = (n + NP_SCAN_ITEMS - 1) / NP_SCAN_ITEMS;
    k_scan_block_sums<<<sb, NP_SCAN_BLOCK, 0, w->stream>>>(n, in, w->d_sums, limit);
    k_scan_sums<<<1, 1024, 0, w->stream>>>(sb, w->d_sums);
    k_scan_apply<<<sb, NP_SCAN_BLOCK, 0, w->stream>>>(n, in, w->d_sums, out, limit);
    w->launches += 3;
}
#define NP_CAND_PER_BODY 64
#ifndef NP_GRID_H_SCALE
#define NP_GRID_H_SCALE 1.25f     // automatic cell edge over the average box extent (measured on the 1M-body world)
#endif
static int ensure_grid(np_world* w)
{
    if (w->grid_cap == w->cap && w->d_grid_ints) return NP_OK;
    cudaFree(w->d_grid_ints); cudaFree(w->d_grid_f); w->d_grid_ints = nullptr; w->d_grid_f = nullptr;
    const size_t c = (size_t)w->cap;
    int nb = 1024; while ((size_t)nb < 2 * c) nb <<= 1;
    const size_t cand_cap = std::min<size_t>((size_t)NP_CAND_PER_BODY * c, (size_t)1 << 30);
    // ints: bucket_cnt[nb + 32] (zeroed every step) | bucket_off[nb + 1 (+31)] | items[c] | large[c] | cand_cnt[c] | cand_off[c] | misc[32] | cand[cand_cap]
    const size_t n_ints = 2 * ((size_t)nb + 32) + 4 * c + 32 + cand_cap;
    CU(cudaMalloc((void**)&w->d_grid_ints, sizeof(int) * n_ints));
    CU(cudaMalloc((void**)&w->d_grid_f, sizeof(float) * (24 * c + 32)));      // aabb[2c] float4 | aabb_prev[2c] float4 | saabb[2c] float4 | params[32]
    GridView& g = w->grid;
    int* p = w->d_grid_ints;
    g.bucket_cnt = p; p += nb + 32; g.bucket_off = p; p += nb + 32; g.items = p; p += c; g.large = p; p += c;
    g.cand_cnt = p; p += c; g.cand_off = p; p += c; g.large_count = p; g.error = p + 1; g.spill_cursor = p + 2; g.active_count = p + 3; p += 32; g.cand = p;
    g.aabb = (float4*)w->d_grid_f; g.aabb_prev = (float4*)w->d_grid_f + 2 * c; g.saabb = (float4*)w->d_grid_f + 4 * c; g.params = w->d_grid_f + 24 * c;
    g.nbuckets = nb; g.cand_cap = (int)std::min<size_t>(cand_cap, 0x7fffffff); g.spill_base = (int)std::min<size_t>(c * NP_CAND_SLOTS, 0x7fffffff);
    if (g.spill_base >= g.cand_cap) return fail(NP_ERR_CAPACITY, "world too large for the grid broad phase's candidate store");
    CU(cudaMemsetAsync(g.large_count, 0, sizeof(int) * 32, w->stream));
    w->grid_cap = w->cap;
    return NP_OK;
}
// K2, enqueued between integrate and the narrow phase
static int enqueue_grid(np_world* w, const StoreView& s, int first, int count, int cache, DevCounters* ctr, int* hit_j, int* flags)
{
    GridView& g = w->grid; g.h_option = w->grid_cell;
    { const char* e = getenv("NP_GRID_H_SCALE"); g.h_scale = e ? (float)atof(e) : NP_GRID_H_SCALE; const char* fh = getenv("NP_GRID_HASH"); g.force_hash = fh ? atoi(fh) : 0; }      // (tools and tests)
    g.swept = w->broadphase == NP_BROADPHASE_SWEPT ? (w->swept_valid ? 2 : 1) : 0;
    if (g.swept && !w->capturing) w->swept_valid = true;
    const int n = s.n, b256 = (n + 255) / 256;
    k_grid_init<<<1, 32, 0, w->stream>>>(g);
    CU(cudaMemsetAsync(g.bucket_cnt, 0, sizeof(int) * ((size_t)g.nbuckets + 32), w->stream));
    const int wvb = (int)std::min<long long>(((long long)n * 8 + 255) / 256, (long long)w->sm_count * 16);      // grid-stride: few blocks, few same-address atomics
    if (cache == CACHE_ALL) k_world_verts<true, 1><<<wvb, 256, 0, w->stream>>>(s, sview(w), g);
    else if (cache == CACHE_BIG) k_world_verts<true, 2><<<wvb, 256, 0, w->stream>>>(s, sview(w), g);
    else if (w->tables_on && w->d_tab) k_body_boxes<<<std::min((n + 255) / 256, w->sm_count * 8), 256, 0, w->stream>>>(s, sview(w), g);
    else k_world_verts<true, 0><<<wvb, 256, 0, w->stream>>>(s, sview(w), g);
    k_grid_params<<<1, 1, 0, w->stream>>>(n, g);
    k_grid_count<<<b256, 256, 0, w->stream>>>(s, g);
    enqueue_scan(w, g.nbuckets + 1, g.bucket_cnt, g.bucket_off, (const int*)g.params + 17);      // one more than the buckets: bucket_cnt[nbuckets] is zero, so bucket_off[nbuckets] = all of them
    k_grid_fill<<<b256, 256, 0, w->stream>>>(s, g);
    if (count > 0) {                 // candidates of the bodies this process searches (the whole world unless sharded)
        k_cand<<<(n + 127) / 128, 128, 0, w->stream>>>(s, g, first, count, ctr, hit_j, flags);
        k_cand_large<<<32, 256, 0, w->stream>>>(s, g, first, count, ctr, hit_j, flags);      // (returns at once in a world without large bodies)
        w->launches += 2;
    }
    w->launches += 5;
    CU(cudaGetLastError());
    return NP_OK;
}

// the kernels of one Physics_Update, enqueued on the world's stream (capturable), in two halves:
//   search : integrate -> [grid] -> GJK/EPA first-hit search of bodies [first, first + count)      (physics.cpp:154-157)
//   respond: OnCollision over every body's collisions -> snapshot                                  (physics.cpp:159-162)
// A sharded world runs them as np_world_step_begin / np_world_step_end with the caller's all-gather of the collision
// records in between; everything but the search itself is replicated and bit-identical on every process.
// part: 0 = Physics::Update of every body, then the search; 1 = Physics::Update only; 2 = the search only (partitioned sharded step)
static int enqueue_search(np_world* w, float dt, bool timed, bool integrate = true, int part = 0)
{
    const int n = (int)w->bodies.size();
    if (n == 0) return NP_OK;
    StoreView s = view(w);
    DevCounters* ctr = w->counters_on ? w->d_ctr : nullptr;
    const int first = sharded(w) ? w->shard_first : 0, count = sharded(w) ? w->shard_count : n;
    const bool part_on = sharded(w) && w->halo > 0;
    if (part != 2) {
        if (timed && !w->capturing) CU(cudaEventRecord(w->ev[0], w->stream));
        if (integrate) {                              // (a step that follows a fused response was integrated by it)
            const int lo = part_on ? first : 0, hi = part_on ? first + count : n;      // a partitioned step integrates its own bodies; the halo's transforms arrive from their owner
            if (hi > lo) {
                const int ib = (hi - lo + 255) / 256;
                if (s.ext) k_integrate<true><<<ib, 256, 0, w->stream>>>(s, dt, 1, lo, hi); else k_integrate<false><<<ib, 256, 0, w->stream>>>(s, dt, 1, lo, hi);
                w->launches += 1;
            }
        }
        if (timed && w->phase_timing) CU(cudaEventRecordWithFlags(w->ev[1], w->stream, w->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
        if (part == 1) { CU(cudaGetLastError()); return NP_OK; }
    }
    const int cache = cache_mode(w, count);
    if (w->broadphase != NP_BROADPHASE_EXACT) { int r = enqueue_grid(w, s, first, count, cache, ctr, hitj(w), cflags(w)); if (r) return r; }
    else if (cache == CACHE_ALL) { k_world_verts<false, 1><<<(n * 8 + 255) / 256, 256, 0, w->stream>>>(s, sview(w), w->grid); w->launches += 1; }
    else if (cache == CACHE_BIG) { k_world_verts<false, 2><<<(n * 8 + 255) / 256, 256, 0, w->stream>>>(s, sview(w), w->grid); w->launches += 1; }
    if (timed && w->phase_timing) CU(cudaEventRecordWithFlags(w->ev[2], w->stream, w->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    CU(cudaMemsetAsync(w->d_zero, 0, sizeof(int) * (NP_ZERO_INTS + (size_t)w->cap), w->stream));        // counters, tickets and the incoming-list heads
    if (count > 0) {
        NarrowIn in; memset(&in, 0, sizeof in); in.world = 1; in.n = count; in.first = first; in.s = s; in.rot_known = 1; in.pf = getenv("NP_PREFETCH") ? atoi(getenv("NP_PREFETCH")) : 1;
        in.ordered = w->order_on;
        if (part_on) { in.halo = w->halo; in.halo_err = w->d_halo_err; }
        if (w->broadphase != NP_BROADPHASE_EXACT) { in.cand = w->grid.cand; in.cand_off = w->grid.cand_off; in.cand_cnt = w->grid.cand_cnt; in.active = w->grid.items; in.n_active = w->grid.active_count; }
        if (w->seed_on && !w->gjk1_old && cache != CACHE_BIG && pick_group(w, count, w->g_gjk_env, false) <= 1) {      // thread-per-body search: tabulate the seed supports once per body
            const int sb_ = std::min(w->sm_count * 8, (n + NP_NARROW_THREADS - 1) / NP_NARROW_THREADS);
            k_seed_supports<<<std::max(1, sb_), NP_NARROW_THREADS, shapes_smem(w), w->stream>>>(s, sview(w), shapes_in_smem(w) ? 1 : 0, w->d_seed);
            in.seed = w->d_seed; w->launches += 1;
        }
        NarrowBufs nb = { w->d_zero, hitj(w), cflags(w), contacts(w), w->d_simplex, w->d_hit_list, w->d_big_list, w->d_hit_list + (size_t)NP_HIT_CLASSES * w->cap, w->d_face0 };
        int r = launch_narrow(w, in, nb, ctr); if (r) return r;
    }
    if (timed && w->phase_timing) CU(cudaEventRecordWithFlags(w->ev[3], w->stream, w->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    CU(cudaGetLastError());
    return NP_OK;
}
static int enqueue_respond(np_world* w, bool timed, const float* fuse_dt = nullptr)
{
    const int n = (int)w->bodies.size();
    if (n == 0) return NP_OK;
    StoreView s = view(w);
    SnapObj* snap = w->snapshot_each_step ? (w->snap_target ? (SnapObj*)w->snap_target : w->d_snap) : nullptr;
    // partitioned sharded step: the responses of the own bodies, fed by the collisions of the own bodies and of the `halo` bodies before them
    const bool part_on = sharded(w) && w->halo > 0;
    const int lo = part_on ? w->shard_first : 0, hi = part_on ? w->shard_first + w->shard_count : n;
    const int llo = part_on ? std::max(0, lo - w->halo) : 0;
    const int b256 = (std::max(hi - lo, 1) + 255) / 256, lb256 = (std::max(hi - llo, 1) + 255) / 256;
    k_link<<<lb256, 256, 0, w->stream>>>(n, w->cap, hitj(w), w->d_ints, d_inc_cnt(w), w->d_inc, llo, hi);
    // + the snapshot record; fuse_dt: also Physics::Update of the NEXT step (np_world_step_n)
    if (fuse_dt) {
        if (s.ext) k_resolve<true, true><<<b256, 256, 0, w->stream>>>(s, hitj(w), contacts(w), d_inc_cnt(w), w->d_inc, snap, w->guid_base, *fuse_dt, lo, hi);
        else k_resolve<true, false><<<b256, 256, 0, w->stream>>>(s, hitj(w), contacts(w), d_inc_cnt(w), w->d_inc, snap, w->guid_base, *fuse_dt, lo, hi);
    } else k_resolve<false, false><<<b256, 256, 0, w->stream>>>(s, hitj(w), contacts(w), d_inc_cnt(w), w->d_inc, snap, w->guid_base, 0.0f, lo, hi);
    if (timed && w->phase_timing) CU(cudaEventRecordWithFlags(w->ev[4], w->stream, w->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    w->launches += 2;
    if (timed && !w->capturing) CU(cudaEventRecord(w->ev[5], w->stream));
    CU(cudaGetLastError());
    return NP_OK;
}
// kind of a step inside np_world_step_n: PLAIN = integrate .. resolve; OPEN = integrate .. resolve fused with the next integrate;
// MID = (integrated by the previous step) .. fused; CLOSE = (integrated by the previous step) .. plain resolve
enum { STEP_PLAIN = 0, STEP_OPEN = 1, STEP_MID = 2, STEP_CLOSE = 3 };
static int enqueue_step(np_world* w, float dt, bool timed, int kind = STEP_PLAIN)
{
    if (w->mid_step) return fail(NP_ERR_INVALID, "a step is open: np_world_step_begin must be closed by np_world_step_end");
    if (sharded(w)) return fail(NP_ERR_INVALID, "sharded world: step with np_world_step_begin / (all-gather) / np_world_step_end");
    int r = enqueue_search(w, dt, timed, kind == STEP_PLAIN || kind == STEP_OPEN); if (r) return r;
    return enqueue_respond(w, timed, (kind == STEP_OPEN || kind == STEP_MID) ? &dt : nullptr);
}

// ------------------------------------------------------------------------------------------------ library
extern "C" {

static int gather_prepare(np_world* w);
static int gather_launch(np_world* w);
static int gather_setup_dma(np_world* w);
int np_abi_version(void) { return NETPHYS_B200_ABI_VERSION; }
const char* np_last_error(void) { return g_err.c_str(); }
int np_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

np_world* np_world_create(int device)
{
    int ndev = np_device_count();
    if (ndev <= 0) { fail(NP_ERR_NO_DEVICE, "no CUDA device: netphys_b200 has no CPU path"); return nullptr; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) { fail(NP_ERR_INVALID, "device index out of range"); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        fail(NP_ERR_NO_DEVICE, "device is not sm_100 (the library is built for sm_100a only)"); return nullptr;
    }
    np_world* w = new np_world();
    w->device = device; w->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) {
        fail(NP_ERR_CUDA, "cannot create stream"); delete w; return nullptr;
    }
    for (auto& e : w->ev) cudaEventCreate(&e);
    { cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); } }   // scratch of pair-list queries stays cached
    setup_kernels();
    if (const char* e = getenv("NP_G_GJK")) w->g_gjk_env = atoi(e);
    if (const char* e = getenv("NP_COOP")) w->coop_env = atoi(e);
    if (const char* e = getenv("NP_GRAPH")) w->graph_on = atoi(e) != 0;
    if (const char* e = getenv("NP_ORDER")) w->order_on = atoi(e) != 0;
    if (const char* e = getenv("NP_G_EPA")) w->g_epa_env = atoi(e);
    if (const char* e = getenv("NP_GJK1_OLD")) w->gjk1_old = atoi(e) != 0;
    if (const char* e = getenv("NP_CBIG")) w->cbig_on = atoi(e) != 0;
    if (const char* e = getenv("NP_EPA_FIRST")) w->epa_first = atoi(e) != 0;
    if (const char* e = getenv("NP_SEED")) w->seed_on = atoi(e) != 0;
    if (const char* e = getenv("NP_FUSE")) w->fuse_on = atoi(e) != 0;
    if (const char* e = getenv("NP_TABLES")) w->tables_on = atoi(e) != 0;
    if (const char* e = getenv("NP_GJK2")) w->gjk2_on = atoi(e) != 0;
    if (const char* e = getenv("NP_EPA2")) w->epa2_on = atoi(e) != 0;
    if (const char* e = getenv("NP_EPA2_TAIL")) w->epa2_tail32 = atoi(e) == 32;
    if (const char* e = getenv("NP_TAB_MIN")) w->tab_min_verts = atoi(e);
    if (cudaMalloc((void**)&w->d_ctr, sizeof(DevCounters)) != cudaSuccess || ensure_arena(w) != NP_OK) { fail(NP_ERR_CUDA, "cannot allocate device scratch"); np_world_destroy(w); return nullptr; }
    cudaMemsetAsync(w->d_ctr, 0, sizeof(DevCounters), w->stream);
    return w;
}
void np_world_destroy(np_world* w)
{
    if (!w) return;
    cudaSetDevice(w->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    np_comm_destroy(w);
    drop_graph(w); free_store(w); cudaFree(w->d_ext);
    cudaFree(w->d_wverts); w->d_wverts = nullptr; w->wverts_cap = 0;
    cudaFree(w->d_aos); w->d_aos = nullptr; w->d_aos_bytes = 0;
    cudaFree(w->d_verts); cudaFree(w->d_soff); cudaFree(w->d_scnt); cudaFree(w->d_tab); cudaFree(w->d_tlim); cudaFree(w->d_cells); cudaFree(w->d_halo_err); cudaFree(w->d_one); cudaFree(w->d_ctr); cudaFree(w->d_flush); cudaFree(w->d_grid_ints); cudaFree(w->d_grid_f);
    cudaFree(w->arena.base);
    if (w->h_stage) cudaFreeHost(w->h_stage);
    for (auto& e : w->ev) if (e) cudaEventDestroy(e);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
}
int np_world_set_option(np_world* w, int option, int64_t value)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    switch (option) {
    case NP_OPT_BROADPHASE:
        if (value != NP_BROADPHASE_EXACT && value != NP_BROADPHASE_GRID && value != NP_BROADPHASE_SWEPT) return fail(NP_ERR_INVALID, "unknown broad-phase");
        w->broadphase = (int)value; drop_graph(w); w->dirty_bodies = true; return NP_OK;
    case NP_OPT_GRID_CELL: { uint32_t u = (uint32_t)value; memcpy(&w->grid_cell, &u, 4); drop_graph(w); return NP_OK; }
    case NP_OPT_COUNTERS: w->counters_on = value ? 1 : 0; drop_graph(w); return NP_OK;
    case NP_OPT_SNAPSHOT_EACH_STEP: w->snapshot_each_step = value ? 1 : 0; drop_graph(w); return NP_OK;
    case NP_OPT_PHASE_TIMING: w->phase_timing = value ? 1 : 0; drop_graph(w); return NP_OK;
    case NP_OPT_SUPPORT_TABLES: w->tables_on = value != 0; drop_graph(w); return NP_OK;
    case NP_OPT_FRICTION: w->friction_on = value ? 1 : 0; drop_graph(w); return NP_OK;
    default: return fail(NP_ERR_INVALID, "unknown option");
    }
}
int np_world_begin_island(np_world* w) { if (!w) return fail(NP_ERR_INVALID, "null world"); if (!w->bodies.empty()) w->island++; return w->island; }
int np_world_body_count(const np_world* w) { return w ? (int)w->bodies.size() : 0; }
static int check_grid_error(np_world* w)
{
    if (w->halo > 0 && w->d_halo_err) {
        int e = 0;
        CU(cudaMemcpyAsync(&e, w->d_halo_err, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        if (e) { CU(cudaMemsetAsync(w->d_halo_err, 0, sizeof(int), w->stream)); return fail(NP_ERR_CAPACITY, "a first-hit search needed a partner beyond the shard halo: raise np_world_set_shard_halo (the step's result is incomplete)"); }
    }
    if (w->broadphase == NP_BROADPHASE_EXACT || !w->d_grid_ints) return NP_OK;
    int e = 0;
    CU(cudaMemcpyAsync(&e, w->grid.error, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    if (e) CU(cudaMemsetAsync(w->grid.error, 0, sizeof(int), w->stream));      // reported once; the flag then starts over
    if (e) return fail(NP_ERR_CAPACITY, "broad phase emitted more than NP_CAND_PER_BODY x capacity candidate pairs; raise the grid cell option or the capacity");
    return NP_OK;
}
int np_world_synchronize(np_world* w) { if (!w) return fail(NP_ERR_INVALID, "null world"); CU(cudaSetDevice(w->device)); CU(cudaStreamSynchronize(w->stream)); return check_grid_error(w); }

// ------------------------------------------------------------------------------------------------ shapes
int np_shape_create_mesh(np_world* w, const float* xyz, int nverts)
{
    if (!w || nverts < 0 || (nverts > 0 && !xyz)) return fail(NP_ERR_INVALID, "bad mesh");
    if (nverts > 255 * 255) return fail(NP_ERR_CAPACITY, "mesh too large");
    HostShape s; s.xyz.assign(xyz, xyz + 3 * (size_t)nverts);
    w->shapes.push_back(std::move(s)); w->dirty_shapes = true;
    return (int)w->shapes.size() - 1;
}
// Mesh::AddVertex, physics_shape.cpp:23-37: tolerant de-duplication, first match wins
static void mesh_add(std::vector<V3>& mesh, V3 p)
{
    for (const V3& q : mesh) if (veq(q, p)) return;
    mesh.push_back(p);
}
static int push_mesh(np_world* w, const std::vector<V3>& mesh)
{
    std::vector<float> xyz; for (const V3& p : mesh) { xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); }
    return np_shape_create_mesh(w, xyz.data(), (int)mesh.size());
}
int np_shape_create_box(np_world* w, float width, float depth, float height)            // CreateBoxMesh, physics_shape.cpp:163-217
{
    (void)depth;                                                                         // :166 half_depth = width / 2.0f
    if (!w) return fail(NP_ERR_INVALID, "null world");
    const float hw = width / 2.0f, hd = width / 2.0f, hh = height / 2.0f;
    const V3 c[8] = { mk(hw, -hh, -hd), mk(hw, -hh, hd), mk(hw, hh, -hd), mk(hw, hh, hd), mk(-hw, -hh, -hd), mk(-hw, -hh, hd), mk(-hw, hh, -hd), mk(-hw, hh, hd) };
    static const int tri[12][3] = { { 0, 1, 3 }, { 0, 3, 2 }, { 1, 5, 7 }, { 1, 7, 3 }, { 5, 4, 6 }, { 5, 6, 7 }, { 4, 0, 2 }, { 4, 2, 6 }, { 2, 3, 7 }, { 2, 7, 6 }, { 4, 5, 1 }, { 4, 1, 0 } };
    std::vector<V3> mesh;
    for (auto& t : tri) for (int k = 0; k < 3; k++) mesh_add(mesh, c[t[k]]);
    return push_mesh(w, mesh);
}
int np_shape_create_sphere(np_world* w, float r)                                         // CreateIcosahadron(r, 3), physics_shape.cpp:42-162
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    const float PIf = 3.1415926535897932384626433832795028841971693993751f;             // lib.h:10
    const float theta = 26.56505117707799f * PIf / 180.0f;
    float st, ct; np_sincosf(theta, &st, &ct);
    std::vector<V3> vl(12, mk(0.f, 0.f, 0.f));
    for (int half = 0; half < 2; half++) {
        float phi = 0;
        for (int i = 1 + 5 * half; i < 6 + 5 * half; i++) {
            float sp, cp; np_sincosf(phi, &sp, &cp);
            vl[i].x = r * ct * cp; vl[i].y = r * ct * sp; vl[i].z = half ? r * st : r * -st;
            phi = (float)((double)phi + (double)(2 * PIf) / 5.0);                        // :56 float += double
        }
    }
    vl[0] = mk(0.0f, 0.0f, -r); vl[11] = mk(0.0f, 0.0f, r);
    std::vector<int> idx = { 0, 2, 1, 0, 3, 2, 0, 4, 3, 0, 5, 4, 0, 1, 5, 1, 2, 7, 2, 3, 8, 3, 4, 9, 4, 5, 10, 5, 1, 6,
                             1, 7, 6, 2, 8, 7, 3, 9, 8, 4, 10, 9, 5, 6, 10, 6, 7, 11, 7, 8, 11, 8, 9, 11, 9, 10, 11, 10, 6, 11 };
    for (int round = 1; round < 3; round++) {                                            // numSubdivisions = 3: the loop body runs twice
        std::vector<V3> nv; std::vector<int> nidx;
        for (size_t j = 0; j + 2 < idx.size(); j += 3) {
            V3 a = vl[idx[j]], b = vl[idx[j + 1]], c = vl[idx[j + 2]];
            V3 x = mul(normalize(add(a, b)), r), y = mul(normalize(add(a, c)), r), z = mul(normalize(add(b, c)), r);
            const V3 seq[12] = { a, x, y, x, b, z, z, y, x, z, c, y };                  // abc -> axy, xbz, zyx, zcy
            for (const V3& p : seq) { nidx.push_back((int)nv.size()); nv.push_back(p); }
        }
        vl.swap(nv); idx.swap(nidx);
    }
    std::vector<V3> mesh;
    for (const V3& p : vl) mesh_add(mesh, p);
    return push_mesh(w, mesh);
}
int np_shape_vertex_count(const np_world* w, int shape) { if (!w || shape < 0 || shape >= (int)w->shapes.size() || w->shapes[shape].dead) return fail(NP_ERR_INVALID, "bad shape"); return (int)w->shapes[shape].xyz.size() / 3; }
int np_shape_get_vertices(const np_world* w, int shape, float* xyz, int cap_verts)
{
    int n = np_shape_vertex_count(w, shape); if (n < 0) return n;
    if (cap_verts < n || !xyz) return fail(NP_ERR_CAPACITY, "vertex buffer too small");
    memcpy(xyz, w->shapes[shape].xyz.data(), sizeof(float) * 3 * (size_t)n);
    return n;
}

// ------------------------------------------------------------------------------------------------ bodies
int np_body_create_batch(np_world* w, int n, const int32_t* shapes, const np_static_physics_data* sd)
{
    if (!w || n < 0 || (n > 0 && (!shapes || !sd))) return fail(NP_ERR_INVALID, "bad arguments");
    for (int i = 0; i < n; i++) if (shapes[i] < 0 || shapes[i] >= (int)w->shapes.size() || w->shapes[shapes[i]].dead) return fail(NP_ERR_INVALID, "bad shape id");
    int first = (int)w->bodies.size();
    if ((long long)first + n > (1LL << 30)) return fail(NP_ERR_CAPACITY, "too many bodies");
    for (int i = 0; i < n; i++) {
        HostBody b; b.sd = sd[i]; b.shape = shapes[i]; b.island = w->island;
        w->bodies.push_back(b);
        np_body_state st; memset(&st, 0, sizeof st);                                     // physics.cpp:11-18: pos/rot from the static data, momenta vector3() == 0
        memcpy(st.position, sd[i].initial_position, 12); memcpy(st.rotation, sd[i].initial_rotation, 12);
        w->host_state.push_back(st);
        w->handle_of_slot.push_back((int)w->slot_of_handle.size()); w->slot_of_handle.push_back((int)w->bodies.size() - 1);
    }
    if (n > 0) w->dirty_bodies = true;
    return n > 0 ? w->handle_of_slot[first] : (int)w->slot_of_handle.size();
}
int np_body_create(np_world* w, int shape, const np_static_physics_data* sd) { int32_t s = shape; return np_body_create_batch(w, 1, &s, sd); }

// validates a caller's body HANDLE and turns it into the slot of the dense store
static int check_body(np_world* w, int& body)
{
    if (!w || body < 0 || body >= (int)w->slot_of_handle.size() || w->slot_of_handle[body] < 0) return fail(NP_ERR_INVALID, "bad body");
    body = w->slot_of_handle[body];
    return NP_OK;
}
int np_body_get_state(np_world* w, int body, np_body_state* out)
{
    int r = check_body(w, body); if (r) return r;
    if (!out) return fail(NP_ERR_INVALID, "null out");
    if (!w->dev_state_valid || body >= w->dev_n) { *out = w->host_state[body]; return NP_OK; }
    CU(cudaSetDevice(w->device));
    CU(cudaMemcpy2DAsync(out, sizeof(float), w->d_state + body, sizeof(float) * w->cap, sizeof(float), S_COUNT, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return NP_OK;
}
static int set_state_slot(np_world* w, int body, const np_body_state* in);
int np_body_set_state(np_world* w, int body, const np_body_state* in)
{
    int r = check_body(w, body); if (r) return r;
    if (!in) return fail(NP_ERR_INVALID, "null in");
    return set_state_slot(w, body, in);
}
static int set_state_slot(np_world* w, int body, const np_body_state* in)
{
    w->host_state[body] = *in;
    if (w->dev_state_valid && body < w->dev_n) {
        CU(cudaSetDevice(w->device));
        CU(cudaMemcpy2DAsync(w->d_state + body, sizeof(float) * w->cap, in, sizeof(float), sizeof(float), S_COUNT, cudaMemcpyHostToDevice, w->stream));
        CU(cudaStreamSynchronize(w->stream));
    }
    return NP_OK;
}
int np_body_reset(np_world* w, int body)                                                 // Physics::Reset, physics.cpp:166-172
{
    int r = check_body(w, body); if (r) return r;
    np_body_state st; memset(&st, 0, sizeof st);
    memcpy(st.position, w->bodies[body].sd.initial_position, 12); memcpy(st.rotation, w->bodies[body].sd.initial_rotation, 12);
    return set_state_slot(w, body, &st);
}
// ---- adapter extension for an ODE-shaped World facade (include/netphys/world.h).  Not reference semantics: the engine
// has gravity only (physics.cpp:178-181).
int np_body_add_force(np_world* w, int body, const float f[3])                            // dBodyAddForce, netphys_common/player.cpp:38
{
    int r = check_body(w, body); if (r) return r;
    if (!f) return fail(NP_ERR_INVALID, "null force");
    r = ensure_device(w); if (r) return r;
    r = ensure_ext(w); if (r) return r;
    k_ext_one<<<1, 1, 0, w->stream>>>(w->d_ext, w->cap, body, 0, f[0], f[1], f[2]);
    CU(cudaGetLastError());
    return NP_OK;
}
int np_body_set_enabled(np_world* w, int body, int enabled)                               // dBodyEnable / dBodyDisable
{
    int r = check_body(w, body); if (r) return r;
    w->bodies[body].enabled = enabled ? 1 : 0;
    r = ensure_device(w); if (r) return r;
    r = ensure_ext(w); if (r) return r;
    k_ext_one<<<1, 1, 0, w->stream>>>(w->d_ext, w->cap, body, 1, enabled ? 1.0f : 0.0f, 0.0f, 0.0f);
    CU(cudaGetLastError());
    return NP_OK;
}
int np_body_is_enabled(np_world* w, int body) { int r = check_body(w, body); if (r) return r; return w->bodies[body].enabled; }   // dBodyIsEnabled, world_s.cpp:48
// Physics::~Physics is a TODO in the reference ("remove from list", physics.cpp:20-23); here a destroyed body leaves the world as if it had
// never been created: the bodies after it move up one slot (the reference's list order is creation order), handles stay valid.
int np_body_destroy(np_world* w, int body)
{
    int r = check_body(w, body); if (r) return r;
    if (w->mid_step) return fail(NP_ERR_INVALID, "np_body_destroy between step_begin and step_end");
    CU(cudaSetDevice(w->device));
    if (w->dev_state_valid && w->dev_n > 0) {                   // the device holds the current state: bring it (and pending forces) home first
        r = download_states(w, w->host_state.data(), w->dev_n); if (r) return r;
        if (w->d_ext) {
            std::vector<float> ext(4 * (size_t)w->cap);
            CU(cudaMemcpyAsync(ext.data(), w->d_ext, sizeof(float) * 4 * (size_t)w->cap, cudaMemcpyDeviceToHost, w->stream));
            CU(cudaStreamSynchronize(w->stream));
            for (int i = 0; i < w->dev_n; i++) for (int c = 0; c < 3; c++) w->bodies[i].pending_force[c] = ext[(size_t)c * w->cap + i];
        }
    }
    const bool had_ext = w->d_ext != nullptr;
    if (w->d_ext) { CU(cudaStreamSynchronize(w->stream)); cudaFree(w->d_ext); w->d_ext = nullptr; }
    const int handle = w->handle_of_slot[body];
    w->bodies.erase(w->bodies.begin() + body); w->host_state.erase(w->host_state.begin() + body); w->handle_of_slot.erase(w->handle_of_slot.begin() + body);
    w->slot_of_handle[handle] = -1;
    for (int i = body; i < (int)w->handle_of_slot.size(); i++) w->slot_of_handle[w->handle_of_slot[i]] = i;
    w->any_destroyed = true; w->restore_ext = w->restore_ext || had_ext;
    drop_graph(w);
    w->dirty_bodies = true; w->dev_state_valid = false; w->dev_n = 0;
    if (sharded(w)) w->order_dirty = true;
    return NP_OK;
}
int np_shape_destroy(np_world* w, int shape)
{
    if (!w || shape < 0 || shape >= (int)w->shapes.size() || w->shapes[shape].dead) return fail(NP_ERR_INVALID, "bad shape");
    for (const HostBody& b : w->bodies) if (b.shape == shape) return fail(NP_ERR_INVALID, "shape is still used by a body");
    w->shapes[shape].xyz.clear(); w->shapes[shape].xyz.shrink_to_fit(); w->shapes[shape].dead = true;      // the id is not reused
    w->dirty_shapes = true;
    return NP_OK;
}
int np_world_clear_bodies(np_world* w)                                                    // World::Reset, netphys_common/world.cpp:129-141
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (w->mid_step) return fail(NP_ERR_INVALID, "np_world_clear_bodies between step_begin and step_end");
    CU(cudaSetDevice(w->device)); CU(cudaStreamSynchronize(w->stream));
    drop_graph(w);
    w->bodies.clear(); w->host_state.clear(); w->island = 0;
    w->slot_of_handle.clear(); w->handle_of_slot.clear(); w->any_destroyed = false;
    w->dirty_bodies = true; w->dev_state_valid = false; w->dev_n = 0;
    if (w->d_ext) { cudaFree(w->d_ext); w->d_ext = nullptr; }
    cudaFree(w->d_wverts); w->d_wverts = nullptr; w->wverts_cap = 0; w->wv_total = 0;
    w->shard_first = 0; w->shard_count = -1; w->x_hit_j = w->x_cflags = nullptr; w->x_contact = nullptr;
    return NP_OK;
}
int np_world_get_states(np_world* w, np_body_state* out, int cap)
{
    if (!w || !out) return fail(NP_ERR_INVALID, "bad arguments");
    int n = (int)w->bodies.size();
    if (cap < n) return fail(NP_ERR_CAPACITY, "state buffer too small");
    if (!w->dev_state_valid) { memcpy(out, w->host_state.data(), sizeof(np_body_state) * (size_t)n); return n; }
    CU(cudaSetDevice(w->device));
    int r = download_states(w, out, w->dev_n); if (r) return r;
    for (int i = w->dev_n; i < n; i++) out[i] = w->host_state[i];
    return n;
}
int np_world_set_states(np_world* w, const np_body_state* in, int n)
{
    if (!w || !in || n != (int)w->bodies.size()) return fail(NP_ERR_INVALID, "bad arguments");
    memcpy(w->host_state.data(), in, sizeof(np_body_state) * (size_t)n);
    if (w->dev_state_valid && !w->dirty_bodies) { CU(cudaSetDevice(w->device)); return upload_states(w, in, n); }
    w->dev_state_valid = false; w->dirty_bodies = true;      // next ensure_device uploads host_state
    return NP_OK;
}
int np_body_get_transform(np_world* w, int body, float out16[16])                         // Physics::GetTransform, physics.h:70-76
{
    int r = check_body(w, body); if (r) return r;
    if (!out16) return fail(NP_ERR_INVALID, "null out");
    r = ensure_device(w); if (r) return r;
    float* d = nullptr;
    CU(cudaMallocAsync((void**)&d, 64, w->stream));
    k_transform_one<<<1, 1, 0, w->stream>>>(view(w), body, d);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out16, d, 64, cudaMemcpyDeviceToHost, w->stream)); CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return NP_OK;
}

// ------------------------------------------------------------------------------------------------ step
// One captured Physics_Update per (dt, snapshot target): with thousands of bodies a step is a dozen kernels of a few microseconds and
// the launch gaps of plain submission cost as much as the kernels (measured 0.31 -> 0.26 ms per step on the 10 k-body world).  The
// phase events are recorded inside the graph, so np_world_last_step_ms keeps its breakdown.
static int get_step_graph(np_world* w, float dt, StepGraph** out, int kind = STEP_PLAIN)
{
    uint32_t bits; memcpy(&bits, &dt, 4);
    for (auto& g : w->graphs) if (g.dt_bits == bits && g.snap == w->snap_target && g.guid == w->guid_base && g.kind == kind) { *out = &g; return NP_OK; }
    if (w->graphs.size() >= 8) { cudaGraphExecDestroy(w->graphs.front().exec); w->graphs.erase(w->graphs.begin()); }
    cudaGraph_t g = nullptr;
    const unsigned long long l0 = w->launches;
    CU(cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal));
    w->capturing = true;                                                  // (phase events become external event-record nodes)
    int r = enqueue_step(w, dt, true, kind);
    w->capturing = false;
    StepGraph sg = { bits, w->snap_target, w->guid_base, kind, nullptr, w->launches - l0 };
    w->launches = l0;                                                     // counted per replay
    cudaError_t e = cudaStreamEndCapture(w->stream, &g);
    if (r) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) return fail(NP_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&sg.exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(NP_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    w->graphs.push_back(sg);
    *out = &w->graphs.back();
    return NP_OK;
}
static int step_once(np_world* w, float dt, int kind = STEP_PLAIN)
{
    if (w->broadphase == NP_BROADPHASE_SWEPT && !w->swept_valid) {      // first swept step after a rebuild: plain launches (its AABB pass has no previous boxes), then the captured steps
        return enqueue_step(w, dt, true, kind);              // (enqueue_grid marks the boxes valid)
    }
    if (!w->graph_on || sharded(w) || w->mid_step) return enqueue_step(w, dt, true, kind);
    StepGraph* g = nullptr;
    int r = get_step_graph(w, dt, &g, kind); if (r) return r;
    CU(cudaEventRecord(w->ev[0], w->stream));
    CU(cudaGraphLaunch(g->exec, w->stream));
    CU(cudaEventRecord(w->ev[5], w->stream));
    w->launches += g->launches;
    return NP_OK;
}
int np_world_step(np_world* w, float dt)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    int r = ensure_device(w); if (r) return r;
    r = ensure_cache(w); if (r) return r;
    if (w->bodies.empty()) return NP_OK;
    r = gather_prepare(w); if (r) return r;
    r = step_once(w, dt); if (r) return r;
    w->timed = true; w->steps++;
    return gather_launch(w);
}
// Sharded pair search of ONE world across processes (BASELINE configs[4]).  Every process holds the whole world;
// process r searches the first hit of bodies [first, first + count) only, and writes hit_j / flags / contact of those bodies into
// the caller's full-length device arrays.  The caller all-gathers the three arrays (the one exchange of the step), then
// np_world_step_end applies the impulse responses of the complete collision list on every process, which keeps the replicated
// state bit-identical everywhere without exchanging it.
int np_world_set_shard(np_world* w, int first, int count, void* d_hit_j, void* d_flags, void* d_contact)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (w->mid_step) return fail(NP_ERR_INVALID, "np_world_set_shard between step_begin and step_end");
    drop_graph(w); w->order_dirty = true;
    if (count < 0) { w->shard_first = 0; w->shard_count = -1; w->x_hit_j = w->x_cflags = nullptr; w->x_contact = nullptr; return NP_OK; }
    if (first < 0 || !d_hit_j || !d_flags || !d_contact) return fail(NP_ERR_INVALID, "bad shard (range or null collision arrays)");
    w->shard_first = first; w->shard_count = count; w->x_hit_j = (int*)d_hit_j; w->x_cflags = (int*)d_flags; w->x_contact = (float4*)d_contact;
    return NP_OK;
}
int np_world_step_begin(np_world* w, float dt)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (w->mid_step) return fail(NP_ERR_INVALID, "np_world_step_begin called twice");
    if (sharded(w) && w->halo > 0) return fail(NP_ERR_INVALID, "partitioned sharded world: np_world_step_integrate / (halo transforms) / np_world_step_search / (halo records) / np_world_step_end, or np_world_step_sharded");
    int r = ensure_device(w); if (r) return r;
    r = ensure_cache(w); if (r) return r;
    const int n = (int)w->bodies.size();
    if (sharded(w) && (long long)w->shard_first + w->shard_count > n) {      // ranks past the end own an empty or clipped range
        if (w->shard_first > n) w->shard_first = n;
        w->shard_count = n - w->shard_first;
    }
    r = enqueue_search(w, dt, true); if (r) return r;
    w->mid_step = true;
    return NP_OK;
}
int np_world_step_end(np_world* w)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (!w->mid_step) return fail(NP_ERR_INVALID, "np_world_step_end without np_world_step_begin");
    CU(cudaSetDevice(w->device));
    w->mid_step = false;
    int r = enqueue_respond(w, true); if (r) return r;
    w->timed = true; w->steps++;
    return NP_OK;
}
// ---- partitioned sharded step (exact search): every rank integrates and responds for its OWN bodies only -------------------------
int np_world_set_shard_halo(np_world* w, int halo_bodies)
{
    if (!w || halo_bodies < 0) return fail(NP_ERR_INVALID, "bad arguments");
    if (w->mid_step || w->step_phase) return fail(NP_ERR_INVALID, "np_world_set_shard_halo inside a step");
    if (halo_bodies > 0 && w->broadphase != NP_BROADPHASE_EXACT) return fail(NP_ERR_INVALID, "partitioned steps support the exact (reference) pair search only");
    CU(cudaSetDevice(w->device));
    if (!w->d_halo_err) { CU(cudaMalloc((void**)&w->d_halo_err, sizeof(int))); CU(cudaMemset(w->d_halo_err, 0, sizeof(int))); }
    drop_graph(w); w->halo = halo_bodies;
    return NP_OK;
}
static void clip_shard(np_world* w)
{
    const int n = (int)w->bodies.size();
    if (sharded(w) && (long long)w->shard_first + w->shard_count > n) {      // ranks past the end own an empty or clipped range
        if (w->shard_first > n) w->shard_first = n;
        w->shard_count = n - w->shard_first;
    }
}
int np_world_step_integrate(np_world* w, float dt)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (!(sharded(w) && w->halo > 0)) return fail(NP_ERR_INVALID, "np_world_step_integrate is the first phase of a PARTITIONED sharded step (np_world_set_shard + np_world_set_shard_halo)");
    if (w->mid_step || w->step_phase) return fail(NP_ERR_INVALID, "a step is already open");
    if (w->broadphase != NP_BROADPHASE_EXACT) return fail(NP_ERR_INVALID, "partitioned steps support the exact (reference) pair search only");
    int r = ensure_device(w); if (r) return r;
    r = ensure_cache(w); if (r) return r;
    clip_shard(w);
    r = enqueue_search(w, dt, true, true, 1); if (r) return r;
    w->step_phase = 1; w->step_dt = dt;
    return NP_OK;
}
int np_world_step_search(np_world* w)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (w->step_phase != 1) return fail(NP_ERR_INVALID, "np_world_step_search without np_world_step_integrate");
    CU(cudaSetDevice(w->device));
    w->step_phase = 0;
    int r = enqueue_search(w, w->step_dt, true, false, 2); if (r) return r;
    w->mid_step = true;                                  // np_world_step_end closes the step
    return NP_OK;
}
int np_world_transforms_device(np_world* w, void** d_xf)
{
    if (!w || !d_xf) return fail(NP_ERR_INVALID, "bad arguments");
    int r = ensure_device(w); if (r) return r;
    *d_xf = w->d_xf;
    return NP_OK;
}
int np_world_step_n(np_world* w, float dt, int steps)
{
    if (!w || steps < 0) return fail(NP_ERR_INVALID, "bad arguments");
    int r = ensure_device(w); if (r) return r;
    r = ensure_cache(w); if (r) return r;
    if (w->bodies.empty() || steps == 0) return NP_OK;
    // consecutive steps: the response of step k and the integration of step k+1 are one kernel (k_resolve<true>)
    for (int i = 0; i < steps; i++) {
        const int kind = (steps == 1 || !w->fuse_on || sharded(w)) ? STEP_PLAIN : (i == 0 ? STEP_OPEN : (i == steps - 1 ? STEP_CLOSE : STEP_MID));
        r = step_once(w, dt, kind); if (r) return r;
    }
    w->timed = true; w->steps += (unsigned long long)steps;                // (the phase events are those of the last step)
    return NP_OK;
}
int np_world_integrate(np_world* w, float dt)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    int r = ensure_device(w); if (r) return r;
    int n = (int)w->bodies.size(); if (n == 0) return NP_OK;
    if (w->d_ext) k_integrate<true><<<(n + 255) / 256, 256, 0, w->stream>>>(view(w), dt, 1, 0, n); else k_integrate<false><<<(n + 255) / 256, 256, 0, w->stream>>>(view(w), dt, 1, 0, n);
    CU(cudaGetLastError());
    return NP_OK;
}
int np_world_get_collisions(np_world* w, np_collision* out, int cap, int* count)
{
    if (!w || !count) return fail(NP_ERR_INVALID, "bad arguments");
    *count = 0;
    int n = w->dev_n; if (!w->dev_state_valid || n == 0 || w->steps == 0) return NP_OK;
    CU(cudaSetDevice(w->device));
    { int rg = check_grid_error(w); if (rg) return rg; }
    std::vector<int> hj(n); std::vector<float4> cd(n);
    CU(cudaMemcpyAsync(hj.data(), hitj(w), sizeof(int) * n, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaMemcpyAsync(cd.data(), contacts(w), sizeof(float4) * n, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    int k = 0;
    const bool part_on = sharded(w) && w->halo > 0;       // partitioned step: this rank knows the collisions of its own bodies (and of the halo before them)
    const int ilo = part_on ? std::min(n, w->shard_first) : 0, ihi = part_on ? std::min(n, w->shard_first + std::max(0, w->shard_count)) : n;
    for (int i = ilo; i < ihi; i++) {
        if (hj[i] < 0) continue;
        if (out && k < cap) {
            np_collision& c = out[k]; memset(&c, 0, sizeof c);
            c.a = w->handle_of_slot[i]; c.b = w->handle_of_slot[hitj_body(hj[i])];
            c.data.penetration_direction[0] = cd[i].x; c.data.penetration_direction[1] = cd[i].y; c.data.penetration_direction[2] = cd[i].z;
            c.data.depth = cd[i].w; c.data.success = (hj[i] & NP_HITJ_SUCCESS) ? 1 : 0;
        }
        k++;
    }
    *count = k;
    if (out && k > cap) return fail(NP_ERR_CAPACITY, "collision buffer too small");
    return NP_OK;
}
int np_world_step_host(np_world* w, float dt, const np_body_state* in, np_body_state* out, int n)
{
    if (!w || !out || n != (int)w->bodies.size()) return fail(NP_ERR_INVALID, "bad arguments");
    int r = ensure_device(w); if (r) return r;
    r = ensure_cache(w); if (r) return r;
    if (in) { r = upload_states(w, in, n); if (r) return r; }
    r = step_once(w, dt); if (r) return r;
    w->timed = true; w->steps++;
    return download_states(w, out, n);
}
int np_world_get_counters(np_world* w, np_counters* out)
{
    if (!w || !out) return fail(NP_ERR_INVALID, "bad arguments");
    CU(cudaSetDevice(w->device));
    DevCounters c;
    CU(cudaMemcpyAsync(&c, w->d_ctr, sizeof c, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    out->steps = w->steps; out->pair_tests = c.pair_tests; out->hits = c.hits; out->gjk_steps = c.gjk_steps; out->support_pairs = c.support_pairs;
    out->epa_iters = c.epa_iters; out->epa_tri_tests = c.epa_tri_tests; out->epa_overflow = c.epa_overflow; out->candidates = c.candidates;
    out->support_verts = c.support_verts; out->kernel_launches = w->launches;
    return NP_OK;
}
int np_world_last_step_ms(np_world* w, float out[6])
{
    if (!w || !out) return fail(NP_ERR_INVALID, "bad arguments");
    CU(cudaSetDevice(w->device));
    CU(cudaStreamSynchronize(w->stream));
    for (int k = 0; k < 6; k++) out[k] = 0.0f;
    if (w->steps == 0) return NP_OK;
    if (w->timed && w->phase_timing) {
        for (int k = 0; k < 5; k++) CU(cudaEventElapsedTime(&out[k], w->ev[k], w->ev[k + 1]));
        CU(cudaEventElapsedTime(&out[5], w->ev[0], w->ev[5]));
    } else CU(cudaEventElapsedTime(&out[5], w->ev[0], w->ev[5]));
    return NP_OK;
}

// ------------------------------------------------------------------------------------------------ narrow phase
static int run_pairs(np_world* w, int n, const int* d_sa, const float* d_pa, const int* d_sb, const float* d_pb, const float* d_xa, const float* d_xb,
                     int* d_hit, float4* d_out, float* kernel_ms, bool is2d = false)
{
    // scratch: counters | xfA | xfB (SoA) | simplex | hit_list | big_list
    const size_t N = (size_t)n;
    const size_t o_zero = 0, o_xa = 4 * NP_ZERO_INTS, o_xb = o_xa + 4 * X_COUNT * N, o_sx = (o_xb + 4 * X_COUNT * N + 15) / 16 * 16, o_hl = o_sx + 96 * N, o_h2 = o_hl + 4 * NP_HIT_CLASSES * N, o_bl = o_h2 + 4 * NP_HIT_CLASSES * N, o_or = o_bl + 4 * N, o_f0 = (o_or + 4 * N + 15) / 16 * 16, total = o_f0 + 64 * N;
    char* d = nullptr;
    CU(cudaMallocAsync((void**)&d, total, w->stream));
    CU(cudaMemsetAsync(d + o_zero, 0, 4 * NP_ZERO_INTS, w->stream));
    if (kernel_ms) CU(cudaEventRecord(w->ev[0], w->stream));
    const int b256 = (n + 255) / 256;
    k_pose_xf<<<b256, 256, 0, w->stream>>>(n, d_pa, d_xa, (float*)(d + o_xa), view(w), d_sa);
    k_pose_xf<<<b256, 256, 0, w->stream>>>(n, d_pb, d_xb, (float*)(d + o_xb), view(w), d_sb);
    w->launches += 2;
    if (is2d) {
        k_detect2d<<<(n + 63) / 64, 64, 0, w->stream>>>(n, sview(w), d_sa, d_sb, (const float*)(d + o_xa), (const float*)(d + o_xb), d_hit, d_out);
        w->launches += 1;
        CU(cudaGetLastError()); CU(cudaFreeAsync(d, w->stream));
        return NP_OK;
    }
    NarrowIn in; memset(&in, 0, sizeof in); in.world = 0; in.n = n; in.s = view(w); in.rot_known = (d_xa == nullptr && d_xb == nullptr) ? 1 : 0;
    in.shape_a = d_sa; in.shape_b = d_sb; in.xf_a = (const float*)(d + o_xa); in.xf_b = (const float*)(d + o_xb);
    if (w->order_on && pick_group(w, n, w->g_gjk_env, false) == 1 && w->shapes.size() > 1) {     // thread-per-pair regime, more than one shape
        int* cnts = (int*)(d + o_zero) + Z_PAIR_CLASS;
        k_pair_order<false><<<b256, 256, 0, w->stream>>>(n, sview(w), d_sa, d_sb, cnts, nullptr);
        k_pair_order<true><<<b256, 256, 0, w->stream>>>(n, sview(w), d_sa, d_sb, cnts, (int*)(d + o_or));
        in.order = (const int*)(d + o_or);
        w->launches += 2;
    }
    NarrowBufs nb = { (int*)(d + o_zero), nullptr, d_hit, d_out, (float4*)(d + o_sx), (int*)(d + o_hl), (int*)(d + o_bl), (int*)(d + o_h2), (float4*)(d + o_f0) };
    int r = launch_narrow(w, in, nb, w->counters_on ? w->d_ctr : nullptr);
    if (kernel_ms) CU(cudaEventRecord(w->ev[5], w->stream));
    CU(cudaGetLastError());
    CU(cudaFreeAsync(d, w->stream));
    if (r) return r;
    if (kernel_ms) { CU(cudaStreamSynchronize(w->stream)); CU(cudaEventElapsedTime(kernel_ms, w->ev[0], w->ev[5])); }
    return NP_OK;
}
static int check_shapes(np_world* w, int n, const int32_t* a, const int32_t* b)
{
    int ns = (int)w->shapes.size();
    for (int i = 0; i < n; i++) if (a[i] < 0 || a[i] >= ns || b[i] < 0 || b[i] >= ns || w->shapes[a[i]].dead || w->shapes[b[i]].dead) return fail(NP_ERR_INVALID, "bad shape id in pair list");
    return NP_OK;
}
static void expand_out(int n, const int* flags, const float4* cd, int32_t* hit, np_collision_data* out)
{
    for (int i = 0; i < n; i++) {
        if (hit) hit[i] = (flags[i] & NP_PAIR_HIT) ? 1 : 0;
        if (out) {
            memset(&out[i], 0, sizeof out[i]);
            out[i].penetration_direction[0] = cd[i].x; out[i].penetration_direction[1] = cd[i].y; out[i].penetration_direction[2] = cd[i].z;
            out[i].depth = cd[i].w; out[i].success = (flags[i] & NP_PAIR_SUCCESS) ? 1 : 0;
        }
    }
}
static int detect_batch_impl(np_world* w, int n, const int32_t* shape_a, const float* pose_a, const int32_t* shape_b, const float* pose_b,
                             int32_t* hit, np_collision_data* out, bool is2d);
int np_detect_collision_batch(np_world* w, int n, const int32_t* shape_a, const float* pose_a, const int32_t* shape_b, const float* pose_b,
                              int32_t* hit, np_collision_data* out)
{ return detect_batch_impl(w, n, shape_a, pose_a, shape_b, pose_b, hit, out, false); }
int np_detect_collision_batch_2d(np_world* w, int n, const int32_t* shape_a, const float* pose_a, const int32_t* shape_b, const float* pose_b,
                                 int32_t* hit, np_collision_data* out)
{ return detect_batch_impl(w, n, shape_a, pose_a, shape_b, pose_b, hit, out, true); }
static int detect_batch_impl(np_world* w, int n, const int32_t* shape_a, const float* pose_a, const int32_t* shape_b, const float* pose_b,
                             int32_t* hit, np_collision_data* out, bool is2d)
{
    if (!w || n < 0 || (n > 0 && (!shape_a || !pose_a || !shape_b || !pose_b))) return fail(NP_ERR_INVALID, "bad arguments");
    if (n == 0) return NP_OK;
    int r = check_shapes(w, n, shape_a, shape_b); if (r) return r;
    r = ensure_device(w); if (r) return r;
    size_t N = (size_t)n;
    char* d = nullptr;
    size_t o_sa = 0, o_sb = o_sa + 4 * N, o_pa = o_sb + 4 * N, o_pb = o_pa + 24 * N, o_hit = o_pb + 24 * N, o_out = (o_hit + 4 * N + 15) / 16 * 16, total = o_out + 16 * N;
    CU(cudaMallocAsync((void**)&d, total, w->stream));
    CU(cudaMemcpyAsync(d + o_sa, shape_a, 4 * N, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemcpyAsync(d + o_sb, shape_b, 4 * N, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(d + o_pa, pose_a, 24 * N, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemcpyAsync(d + o_pb, pose_b, 24 * N, cudaMemcpyHostToDevice, w->stream));
    r = run_pairs(w, n, (int*)(d + o_sa), (float*)(d + o_pa), (int*)(d + o_sb), (float*)(d + o_pb), nullptr, nullptr, (int*)(d + o_hit), (float4*)(d + o_out), nullptr, is2d);
    if (r) { cudaFreeAsync(d, w->stream); return r; }
    std::vector<int> fl(N); std::vector<float4> cd(N);
    CU(cudaMemcpyAsync(fl.data(), d + o_hit, 4 * N, cudaMemcpyDeviceToHost, w->stream)); CU(cudaMemcpyAsync(cd.data(), d + o_out, 16 * N, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    expand_out(n, fl.data(), cd.data(), hit, out);
    return NP_OK;
}
int np_detect_collision(np_world* w, int shape_a, const float a_transform[16], int shape_b, const float b_transform[16], int is3D, np_collision_data* out, int* hit)
{
    if (!w || !a_transform || !b_transform) return fail(NP_ERR_INVALID, "bad arguments");
    int32_t sa = shape_a, sb = shape_b;
    int r = check_shapes(w, 1, &sa, &sb); if (r) return r;
    r = ensure_device(w); if (r) return r;
    // one pair: a persistent 256-byte device record and one copy each way ([0,64) A's matrix, [64,128) B's, [128,136) the shape ids | [144,160) contact, [160,164) flags)
    if (!w->d_one) CU(cudaMalloc((void**)&w->d_one, 256));
    char* d = w->d_one;
    alignas(16) char h_in[136];
    memcpy(h_in, a_transform, 64); memcpy(h_in + 64, b_transform, 64); memcpy(h_in + 128, &sa, 4); memcpy(h_in + 132, &sb, 4);
    CU(cudaMemcpyAsync(d, h_in, sizeof h_in, cudaMemcpyHostToDevice, w->stream));       // (pageable source: staged by the driver before the call returns)
    r = run_pairs(w, 1, (int*)(d + 128), nullptr, (int*)(d + 132), nullptr, (float*)d, (float*)(d + 64), (int*)(d + 160), (float4*)(d + 144), nullptr, !is3D);
    if (r) return r;
    alignas(16) char h_out[32];
    CU(cudaMemcpyAsync(h_out, d + 144, 20, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    float4 cd; int fl; memcpy(&cd, h_out, 16); memcpy(&fl, h_out + 16, 4);
    int32_t h; expand_out(1, &fl, &cd, &h, out);
    if (hit) *hit = h;
    return NP_OK;
}
int np_detect_collision_batch_device(np_world* w, int n, const void* d_shape_a, const void* d_pose_a, const void* d_shape_b, const void* d_pose_b,
                                     void* d_hit, void* d_out, float* kernel_ms)
{
    if (!w || n < 0 || !d_shape_a || !d_pose_a || !d_shape_b || !d_pose_b || !d_hit || !d_out) return fail(NP_ERR_INVALID, "bad arguments");
    int r = ensure_device(w); if (r) return r;
    if (n == 0) return NP_OK;
    return run_pairs(w, n, (const int*)d_shape_a, (const float*)d_pose_a, (const int*)d_shape_b, (const float*)d_pose_b, nullptr, nullptr, (int*)d_hit, (float4*)d_out, kernel_ms);
}
int np_shape_support(np_world* w, int shape, const float dir[3], const float world[16], float out[3])
{
    if (!w || !dir || !world || !out || shape < 0 || shape >= (int)w->shapes.size()) return fail(NP_ERR_INVALID, "bad arguments");
    int r = ensure_device(w); if (r) return r;
    float* d = nullptr;
    CU(cudaMallocAsync((void**)&d, sizeof(float) * 32, w->stream));
    CU(cudaMemcpyAsync(d, dir, 12, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemcpyAsync(d + 4, world, 64, cudaMemcpyHostToDevice, w->stream));
    k_selftest_support<<<1, 1, 0, w->stream>>>(sview(w), shape, d, d + 4, d + 24);
    CU(cudaMemcpyAsync(out, d + 24, 12, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return NP_OK;
}
// ---- the reference's TEST_PROGRAM interface (engine/physics.h:115-127): the algorithm one call at a time on a simplex the caller owns
static_assert(sizeof(np_simplex_point) == 36 && sizeof(np_simplex_face) == 24, "simplex.h:5-20 layouts");
int np_gjk_step(np_world* w, int shape_a, const float a_transform[16], int shape_b, const float b_transform[16], int is3D,
                np_simplex_point* verts, int* nverts, int* contains_origin, int* result)
{
    if (!w || !a_transform || !b_transform || !verts || !nverts || !contains_origin || !result) return fail(NP_ERR_INVALID, "bad arguments");
    if (*nverts < 1 || *nverts > 4 || (!is3D && *nverts > 3)) return fail(NP_ERR_INVALID, "DetectCollisionStep takes a simplex of 1..4 points (1..3 in 2-D)");   // the reference asserts
    int32_t sa = shape_a, sb = shape_b;
    int r = check_shapes(w, 1, &sa, &sb); if (r) return r;
    r = ensure_device(w); if (r) return r;
    char* d = nullptr;
    int io[3] = { *nverts, *contains_origin ? 1 : 0, 0 };
    CU(cudaMallocAsync((void**)&d, 512, w->stream));                         // [0,64) xa  [64,128) xb  [128,272) 4 points  [288,300) io
    CU(cudaMemcpyAsync(d, a_transform, 64, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemcpyAsync(d + 64, b_transform, 64, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(d + 128, verts, 36 * (size_t)*nverts, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemcpyAsync(d + 288, io, 12, cudaMemcpyHostToDevice, w->stream));
    k_gjk_step<<<1, 1, 0, w->stream>>>(sview(w), sa, sb, (const float*)d, (const float*)(d + 64), is3D ? 1 : 0, (float*)(d + 128), (int*)(d + 288));
    np_simplex_point out[4];
    CU(cudaMemcpyAsync(out, d + 128, sizeof out, cudaMemcpyDeviceToHost, w->stream)); CU(cudaMemcpyAsync(io, d + 288, 12, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    memcpy(verts, out, 36 * (size_t)io[0]);
    *nverts = io[0]; *contains_origin = io[1]; *result = io[2];
    return NP_OK;
}
int np_epa_steps(np_world* w, int shape_a, const float a_transform[16], int shape_b, const float b_transform[16], int max_steps, int is3D,
                 np_simplex_point* verts, int* nverts, int verts_cap, np_simplex_face* faces, int* nfaces, int faces_cap,
                 int contains_origin, np_collision_data* out, int* found)
{
    if (!w || !a_transform || !b_transform || !verts || !nverts || !nfaces || !found || *nverts < 0 || *nfaces < 0 || *nverts > verts_cap || *nfaces > faces_cap || (*nfaces > 0 && !faces))
        return fail(NP_ERR_INVALID, "bad arguments");
    if (*nverts > 255 || *nfaces > NP_FBIG) return fail(NP_ERR_CAPACITY, "polytope too large for the step interface");
    for (int f = 0; f < *nfaces; f++) for (int k = 0; k < 3; k++)
        if (faces[f].point_index[k] < 0 || faces[f].point_index[k] >= *nverts) return fail(NP_ERR_INVALID, "face refers to a vertex the simplex does not have");
    int32_t sa = shape_a, sb = shape_b;
    int r = check_shapes(w, 1, &sa, &sb); if (r) return r;
    r = ensure_device(w); if (r) return r;
    const size_t vcap = (size_t)std::max(verts_cap, 1), fcap = (size_t)std::max(faces_cap, 1);
    const size_t o_v = 128, o_fi = o_v + 36 * vcap, o_fn = o_fi + 12 * fcap, o_io = o_fn + 12 * fcap, o_out = o_io + 32, total = o_out + 16;
    std::vector<int32_t> fi(3 * fcap, 0); std::vector<float> fn(3 * fcap, 0.0f);
    for (int f = 0; f < *nfaces; f++) for (int k = 0; k < 3; k++) { fi[3 * f + k] = faces[f].point_index[k]; fn[3 * f + k] = faces[f].normal[k]; }
    int io[8] = { *nverts, *nfaces, contains_origin ? 1 : 0, 0, 0, 0, 0, 0 };
    char* d = nullptr;
    CU(cudaMallocAsync((void**)&d, total, w->stream));
    CU(cudaMemcpyAsync(d, a_transform, 64, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemcpyAsync(d + 64, b_transform, 64, cudaMemcpyHostToDevice, w->stream));
    if (*nverts) CU(cudaMemcpyAsync(d + o_v, verts, 36 * (size_t)*nverts, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(d + o_fi, fi.data(), 12 * fcap, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemcpyAsync(d + o_fn, fn.data(), 12 * fcap, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(d + o_io, io, 32, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemsetAsync(d + o_out, 0, 16, w->stream));
    k_epa_steps<<<1, 1, 0, w->stream>>>(sview(w), sa, sb, (const float*)d, (const float*)(d + 64), max_steps, is3D ? 1 : 0, (float*)(d + o_v), verts_cap,
                                        (int*)(d + o_fi), (float*)(d + o_fn), faces_cap, (int*)(d + o_io), (float*)(d + o_out), w->arena);
    float o4[4];
    CU(cudaMemcpyAsync(io, d + o_io, 32, cudaMemcpyDeviceToHost, w->stream)); CU(cudaMemcpyAsync(o4, d + o_out, 16, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    if (io[4]) { cudaFreeAsync(d, w->stream); return fail(NP_ERR_CAPACITY, "polytope outgrew verts_cap / faces_cap (or the 24-vertex step limit)"); }
    if (io[0]) CU(cudaMemcpyAsync(verts, d + o_v, 36 * (size_t)io[0], cudaMemcpyDeviceToHost, w->stream));
    if (io[1]) { CU(cudaMemcpyAsync(fi.data(), d + o_fi, 12 * (size_t)io[1], cudaMemcpyDeviceToHost, w->stream)); CU(cudaMemcpyAsync(fn.data(), d + o_fn, 12 * (size_t)io[1], cudaMemcpyDeviceToHost, w->stream)); }
    CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    for (int f = 0; f < io[1]; f++) for (int k = 0; k < 3; k++) { faces[f].point_index[k] = fi[3 * f + k]; faces[f].normal[k] = fn[3 * f + k]; }
    *nverts = io[0]; *nfaces = io[1]; *found = io[3];
    if (out && io[3]) { memset(out, 0, sizeof *out); memcpy(out->penetration_direction, o4, 12); out->depth = o4[3]; }   // FindIntersectionPoints never writes `success` (:523-531)
    return NP_OK;
}
int np_search_direction(np_world* w, const np_simplex_point* verts, int nverts, float out[3])
{
    if (!w || !verts || !out || nverts < 1 || nverts > 3) return fail(NP_ERR_INVALID, "GetSearchDirection takes 1..3 points");   // the reference returns an uninitialised vector otherwise
    int r = ensure_device(w); if (r) return r;
    float* d = nullptr;
    CU(cudaMallocAsync((void**)&d, 128, w->stream));
    CU(cudaMemcpyAsync(d, verts, 36 * (size_t)nverts, cudaMemcpyHostToDevice, w->stream));
    k_search_direction<<<1, 1, 0, w->stream>>>(d, nverts, d + 28);
    CU(cudaMemcpyAsync(out, d + 28, 12, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return NP_OK;
}
void np_simplex_face_normal(const np_simplex_point* verts, int va, int vb, int vc, float out[3])
{
    // Simplex::AddFace (simplex.h:37-53) on the host: same binary32 operations as EpaMachine::add_face_at
    V3 a = mk(verts[va].p[0], verts[va].p[1], verts[va].p[2]), b = mk(verts[vb].p[0], verts[vb].p[1], verts[vb].p[2]), c = mk(verts[vc].p[0], verts[vc].p[1], verts[vc].p[2]);
    V3 n = normalize(cross(sub(b, a), sub(c, a)));
    V3 nn = (dot(n, a) != 0.0f) ? n : neg(n);
    out[0] = nn.x; out[1] = nn.y; out[2] = nn.z;
}
int np_selftest_support_table(np_world* w, int shape, int n, const float* dirs, const float* xf16, float* out_scan, float* out_table, int32_t* used)
{
    if (!w || n < 0 || !dirs || !xf16 || !out_scan || !out_table || !used || shape < 0 || shape >= (int)w->shapes.size()) return fail(NP_ERR_INVALID, "bad arguments");
    int r = ensure_device(w); if (r) return r;
    if (n == 0) return NP_OK;
    const size_t N = (size_t)n;
    char* d = nullptr;
    const size_t o_d = 0, o_x = o_d + 12 * N, o_a = o_x + 64 * N, o_b = o_a + 12 * N, o_u = o_b + 12 * N, total = o_u + 4 * N;
    CU(cudaMallocAsync((void**)&d, total, w->stream));
    CU(cudaMemcpyAsync(d + o_d, dirs, 12 * N, cudaMemcpyHostToDevice, w->stream)); CU(cudaMemcpyAsync(d + o_x, xf16, 64 * N, cudaMemcpyHostToDevice, w->stream));
    k_selftest_support_table<<<(n + 127) / 128, 128, 0, w->stream>>>(sview(w), shape, n, (const float*)(d + o_d), (const float*)(d + o_x), (float*)(d + o_a), (float*)(d + o_b), (int*)(d + o_u));
    CU(cudaMemcpyAsync(out_scan, d + o_a, 12 * N, cudaMemcpyDeviceToHost, w->stream)); CU(cudaMemcpyAsync(out_table, d + o_b, 12 * N, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaMemcpyAsync(used, d + o_u, 4 * N, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return NP_OK;
}
// table statistics of a shape (host side): out = { N, cells that fall back to the scan, mean candidates per cell, |t|_1 reach }
int np_shape_support_table_info(np_world* w, int shape, float out[4])
{
    if (!w || !out || shape < 0 || shape >= (int)w->shapes.size()) return fail(NP_ERR_INVALID, "bad arguments");
    HostShape& s = w->shapes[(size_t)shape];
    if (s.tab_state == 0) s.tab_state = (!s.dead && cached_support_table(s.xyz, s.tab)) ? 1 : -1;
    if (s.tab_state != 1) { out[0] = out[1] = out[2] = out[3] = 0.0f; return 0; }
    out[0] = (float)(1 << s.tab.log_n); out[1] = (float)s.tab.scan_cells; out[2] = (float)s.tab.mean_candidates; out[3] = s.tab.tlim;
    return 1;
}
int np_body_apply_impulse_response(np_world* w, int body, const np_collision_data* data, const float normal[3])
{
    // Physics::ApplyImpulseResponse (physics.cpp:27-46) as a stand-alone call on one body
    int r = check_body(w, body); if (r) return r;
    if (!data || !normal) return fail(NP_ERR_INVALID, "bad arguments");
    r = ensure_device(w); if (r) return r;
    k_impulse_one<<<1, 1, 0, w->stream>>>(view(w), body, mk(data->penetration_direction[0], data->penetration_direction[1], data->penetration_direction[2]),
                                          data->depth, mk(normal[0], normal[1], normal[2]));
    CU(cudaGetLastError());
    return NP_OK;
}

// ------------------------------------------------------------------------------------------------ snapshot
int np_world_snapshot_device(np_world* w, void** d_ptr, size_t* bytes)
{
    if (!w || !d_ptr || !bytes) return fail(NP_ERR_INVALID, "bad arguments");
    int r = ensure_device(w); if (r) return r;
    int n = (int)w->bodies.size();
    if (n > 0) { k_snapshot<<<(n + 255) / 256, 256, 0, w->stream>>>(view(w), w->guid_base, w->d_snap); CU(cudaGetLastError()); w->launches++; }
    *d_ptr = w->d_snap; *bytes = sizeof(SnapObj) * (size_t)n;
    return NP_OK;
}
int np_world_set_snapshot_target(np_world* w, void* d_dst, uint32_t guid_base)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    w->snap_target = d_dst; w->guid_base = guid_base;          // (captured steps are keyed by the target)
    return NP_OK;
}
int np_world_snapshot(np_world* w, np_snapshot_object* dst, int cap)
{
    if (!w || !dst) return fail(NP_ERR_INVALID, "bad arguments");
    int n = (int)w->bodies.size();
    if (cap < n) return fail(NP_ERR_CAPACITY, "snapshot buffer too small");
    void* d; size_t bytes;
    int r = np_world_snapshot_device(w, &d, &bytes); if (r) return r;
    if (bytes) CU(cudaMemcpyAsync(dst, d, bytes, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return n;
}

// ------------------------------------------------------------------------------------------------ command frames (world_s.cpp)
int np_world_record_frame(np_world* w, double time_ms)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    CommandFrame f; f.id = w->frame_counter++; f.time_ms = time_ms;
    f.objects.resize(w->bodies.size());
    if (!f.objects.empty()) { int r = np_world_snapshot(w, f.objects.data(), (int)f.objects.size()); if (r < 0) return r; }
    w->frames.push_back(std::move(f));
    size_t drop = 0;                                          // erase old command frames, world_s.cpp:57-68
    while (drop < w->frames.size() && !(time_ms - w->frames[drop].time_ms < 5000.0)) drop++;
    if (drop) w->frames.erase(w->frames.begin(), w->frames.begin() + (long)drop);
    return (int)(w->frame_counter - 1);
}
int np_world_frame_count(const np_world* w) { return w ? (int)w->frames.size() : 0; }
int np_world_fill_packet(np_world* w, int packet_id, void* buffer, int cap)
{
    if (!w || !buffer) return fail(NP_ERR_INVALID, "bad arguments");
    if (w->frames.empty()) return 0;                          // "Tried to send world state before any command frames were generated"
    const CommandFrame& f = w->frames.back();
    static_assert(sizeof(np_snapshot_object) == 36, "CommandFrameObject is 36 bytes");
    const int count = (int)f.objects.size();
    const long long data = 4 + 8 + 4 + 36LL * count, total = 8 + data;
    if (total > cap) return 0;                                // Packet::Serialize returns 0 when the buffer is too small
    char* p = (char*)buffer;
    const int id = packet_id, size = (int)data, fid = (int)f.id;
    memcpy(p, &id, 4); memcpy(p + 4, &size, 4); memcpy(p + 8, &fid, 4); memcpy(p + 12, &f.time_ms, 8); memcpy(p + 20, &count, 4);
    if (count) memcpy(p + 24, f.objects.data(), 36 * (size_t)count);
    return (int)total;
}

// ------------------------------------------------------------------------------------------------ NCCL behind the C ABI
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
int load_nccl()
{
    if (g_nccl.lib) return NP_OK;
    const char* names[] = { getenv("NP_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void* lib = nullptr;
    for (const char* n : names) if (n && *n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) return fail(NP_ERR_INVALID, std::string("cannot load NCCL (libnccl.so.2; set NP_NCCL_LIB): ") + dlerror());
    NcclApi a; a.lib = lib;
    *(void**)&a.GetUniqueId = dlsym(lib, "ncclGetUniqueId"); *(void**)&a.CommInitRank = dlsym(lib, "ncclCommInitRank");
    *(void**)&a.CommDestroy = dlsym(lib, "ncclCommDestroy"); *(void**)&a.AllGather = dlsym(lib, "ncclAllGather");
    *(void**)&a.GroupStart = dlsym(lib, "ncclGroupStart"); *(void**)&a.GroupEnd = dlsym(lib, "ncclGroupEnd");
    *(void**)&a.GetErrorString = dlsym(lib, "ncclGetErrorString");
    *(void**)&a.Send = dlsym(lib, "ncclSend"); *(void**)&a.Recv = dlsym(lib, "ncclRecv");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GroupStart || !a.GroupEnd) return fail(NP_ERR_INVALID, "libnccl lacks a required symbol");
    g_nccl = a;
    return NP_OK;
}
}  // namespace
#define NC(call)                                                                                                      \
    do {                                                                                                              \
        ncclResult_t e_ = (call);                                                                                     \
        if (e_ != ncclSuccess) return fail(NP_ERR_CUDA, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e_) : "nccl error")); \
    } while (0)

int np_comm_get_unique_id(void* id128)
{
    if (!id128) return fail(NP_ERR_INVALID, "null id");
    static_assert(sizeof(ncclUniqueId) == NP_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    int r = load_nccl(); if (r) return r;
    ncclUniqueId id; NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return NP_OK;
}
static void free_gather(np_world* w)
{
    for (int b = 0; b < 2; b++) { for (char* p : w->peer_gather[b]) if (p) cudaIpcCloseMemHandle(p); w->peer_gather[b].clear(); }
    if (w->gather_dma && w->comm && w->d_bar) {      // nobody frees a buffer a peer still has mapped: every rank has closed its mappings once this (collective) exchange is through
        if (g_nccl.AllGather(w->d_bar + w->rank, w->d_bar, sizeof(int), ncclChar, w->comm, w->comm_stream) == ncclSuccess) cudaStreamSynchronize(w->comm_stream);
    }
    w->gather_dma = false; cudaFree(w->d_bar); w->d_bar = nullptr; cudaGetLastError();
    for (cudaStream_t st : w->peer_stream) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (cudaEvent_t e : w->ev_peer) if (e) cudaEventDestroy(e);
    if (w->ev_fan) cudaEventDestroy(w->ev_fan);
    w->peer_stream.clear(); w->ev_peer.clear(); w->ev_fan = nullptr;
    for (int b = 0; b < 2; b++) { cudaFree(w->d_gather[b]); w->d_gather[b] = nullptr; w->gather_pending[b] = false; }
    w->gather_slot = 0; w->gather_last = -1;
}
int np_comm_init(np_world* w, int nranks, int rank, const void* id128)
{
    if (!w || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(NP_ERR_INVALID, "bad arguments");
    if (w->comm) return fail(NP_ERR_INVALID, "the world already has a communicator");
    int r = load_nccl(); if (r) return r;
    CU(cudaSetDevice(w->device));
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    NC(g_nccl.CommInitRank(&w->comm, nranks, id, rank));
    w->nranks = nranks; w->rank = rank;
    int lo = 0, hi = 0; CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));        // (hi = numerically lowest = highest priority)
    CU(cudaStreamCreateWithPriority(&w->comm_stream, cudaStreamNonBlocking, hi));
    for (int b = 0; b < 2; b++) { CU(cudaEventCreate(&w->ev_step_done[b])); CU(cudaEventCreate(&w->ev_gather_begin[b])); CU(cudaEventCreate(&w->ev_gather_done[b])); }
    return NP_OK;
}
int np_comm_destroy(np_world* w)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (!w->comm) return NP_OK;
    CU(cudaSetDevice(w->device));
    CU(cudaStreamSynchronize(w->stream)); CU(cudaStreamSynchronize(w->comm_stream));
    drop_graph(w);
    if (w->gather_slot) { w->snap_target = nullptr; w->snapshot_each_step = 0; }
    free_gather(w);
    if (w->d_shard_hit) { w->shard_first = 0; w->shard_count = -1; w->x_hit_j = w->x_cflags = nullptr; w->x_contact = nullptr; w->order_dirty = true; }
    cudaFree(w->d_shard_hit); cudaFree(w->d_shard_flags); cudaFree(w->d_shard_contact); w->d_shard_hit = w->d_shard_flags = nullptr; w->d_shard_contact = nullptr; w->shard_slot = 0;
    g_nccl.CommDestroy(w->comm); w->comm = nullptr; w->nranks = 1; w->rank = 0;
    for (int b = 0; b < 2; b++) { cudaEventDestroy(w->ev_step_done[b]); cudaEventDestroy(w->ev_gather_begin[b]); cudaEventDestroy(w->ev_gather_done[b]); w->ev_step_done[b] = w->ev_gather_begin[b] = w->ev_gather_done[b] = nullptr; }
    cudaStreamDestroy(w->comm_stream); w->comm_stream = nullptr;
    return NP_OK;
}
// The snapshot all-gather WITHOUT SMs.  ncclAllGather runs as a kernel: on a GPU whose SMs are all held by the persistent narrow-phase
// grids it takes blocks away from the step it is meant to hide behind (8 ranks: every step 0.15 ms longer than on a lone GPU).  The
// payload is one contiguous slot per rank, so each rank can PUSH its slot into every peer's buffer with plain device-to-device copies
// -- copy engines over NVLink / NVSwitch, no kernel -- once the peers' buffers are mapped here (CUDA IPC; one process per GPU).  Two
// 4-byte NCCL all-gathers per step frame the copies: one before them (every rank has CALLED this step, so nobody still reads the
// buffer about to be overwritten: the gathered snapshot of step k stays valid until step k + 2 is called, as with ncclAllGather) and
// one after (every rank's copies have landed).  Set up collectively by np_world_set_snapshot_gather; if any rank cannot export or map
// a buffer (ranks that share a process, a platform without peer access), or NP_GATHER_NCCL=1, every rank keeps ncclAllGather.
static int gather_setup_dma(np_world* w)
{
    w->gather_dma = false;
    if (w->nranks < 2) return NP_OK;
    struct Rec { cudaIpcMemHandle_t h[2]; int ok; int pad[3]; };
    static_assert(sizeof(Rec) == 2 * sizeof(cudaIpcMemHandle_t) + 16, "record layout");
    const int n = w->nranks;
    char* d_x = nullptr; CU(cudaMalloc((void**)&d_x, sizeof(Rec) * (size_t)n));
    CU(cudaMalloc((void**)&w->d_bar, sizeof(int) * (size_t)(n + 1))); CU(cudaMemset(w->d_bar, 0, sizeof(int) * (size_t)(n + 1)));
    std::vector<Rec> all(n);
    auto exchange = [&](const Rec& mine) -> int {
        CU(cudaMemcpy(d_x + sizeof(Rec) * (size_t)w->rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice));
        NC(g_nccl.AllGather(d_x + sizeof(Rec) * (size_t)w->rank, d_x, sizeof(Rec), ncclChar, w->comm, w->comm_stream));
        CU(cudaStreamSynchronize(w->comm_stream));
        CU(cudaMemcpy(all.data(), d_x, sizeof(Rec) * (size_t)n, cudaMemcpyDeviceToHost));
        return NP_OK;
    };
    Rec me; memset(&me, 0, sizeof me);
    me.ok = !getenv("NP_GATHER_NCCL") || atoi(getenv("NP_GATHER_NCCL")) == 0;
    for (int b = 0; b < 2 && me.ok; b++) if (cudaIpcGetMemHandle(&me.h[b], w->d_gather[b]) != cudaSuccess) { cudaGetLastError(); me.ok = 0; }
    int r = exchange(me); if (r) { cudaFree(d_x); return r; }
    bool every = true; for (int p = 0; p < n; p++) every = every && all[p].ok;
    Rec second; memset(&second, 0, sizeof second); second.ok = every ? 1 : 0;
    if (every) {
        for (int b = 0; b < 2; b++) w->peer_gather[b].assign(n, nullptr);
        for (int p = 0; p < n && second.ok; p++) {
            if (p == w->rank) continue;
            for (int b = 0; b < 2 && second.ok; b++) {
                void* q = nullptr;
                if (cudaIpcOpenMemHandle(&q, all[p].h[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); second.ok = 0; }
                else w->peer_gather[b][p] = (char*)q;
            }
        }
    }
    r = exchange(second); cudaFree(d_x); if (r) return r;      // (every rank takes part in both exchanges, whatever happened to it)
    bool use = every; for (int p = 0; p < n; p++) use = use && all[p].ok;
    if (!use) { for (int b = 0; b < 2; b++) { for (char* q : w->peer_gather[b]) if (q) cudaIpcCloseMemHandle(q); w->peer_gather[b].clear(); } cudaGetLastError(); }
    if (use) {
        int lo = 0, hi = 0; CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        w->peer_stream.assign(n, nullptr); w->ev_peer.assign(n, nullptr);
        for (int p = 0; p < n; p++) if (p != w->rank) { CU(cudaStreamCreateWithPriority(&w->peer_stream[p], cudaStreamNonBlocking, hi)); CU(cudaEventCreateWithFlags(&w->ev_peer[p], cudaEventDisableTiming)); }
        CU(cudaEventCreateWithFlags(&w->ev_fan, cudaEventDisableTiming));
    }
    w->gather_dma = use;
    return NP_OK;
}
int np_world_set_snapshot_gather(np_world* w, int slot_bodies, uint32_t guid_base)
{
    if (!w || slot_bodies < 0) return fail(NP_ERR_INVALID, "bad arguments");
    if (!w->comm) return fail(NP_ERR_INVALID, "np_comm_init first");
    CU(cudaSetDevice(w->device));
    CU(cudaStreamSynchronize(w->stream)); CU(cudaStreamSynchronize(w->comm_stream));
    drop_graph(w); free_gather(w);
    w->snap_target = nullptr;
    if (slot_bodies == 0) return NP_OK;
    if (slot_bodies < (int)w->bodies.size()) return fail(NP_ERR_CAPACITY, "snapshot slot smaller than the world");
    const size_t bytes = sizeof(SnapObj) * (size_t)slot_bodies * (size_t)w->nranks;
    for (int b = 0; b < 2; b++) { CU(cudaMalloc((void**)&w->d_gather[b], bytes)); CU(cudaMemsetAsync(w->d_gather[b], 0, bytes, w->stream)); }
    w->gather_slot = slot_bodies; w->gather_guid = guid_base; w->gather_flip = 0; w->snapshot_each_step = 1;
    return gather_setup_dma(w);
}
// before a step: aim the snapshot at this rank's slot of the next gather buffer, once that buffer's previous gather has drained
static int gather_prepare(np_world* w)
{
    if (!w->gather_slot) return NP_OK;
    const int b = w->gather_flip;
    if (w->gather_pending[b]) CU(cudaStreamWaitEvent(w->stream, w->ev_gather_done[b], 0));
    w->snap_target = w->d_gather[b] + sizeof(SnapObj) * (size_t)w->gather_slot * (size_t)w->rank;
    w->guid_base = w->gather_guid;
    return NP_OK;
}
// after a step's kernels: the all-gather, in place, on the communicator's own high-priority stream
static int gather_launch(np_world* w)
{
    if (!w->gather_slot) return NP_OK;
    const int b = w->gather_flip;
    CU(cudaEventRecord(w->ev_step_done[b], w->stream));
    CU(cudaStreamWaitEvent(w->comm_stream, w->ev_step_done[b], 0));
    CU(cudaEventRecord(w->ev_gather_begin[b], w->comm_stream));
    const size_t slot_bytes = sizeof(SnapObj) * (size_t)w->gather_slot;
    if (w->gather_dma) {
        char* mine = w->d_gather[b] + slot_bytes * (size_t)w->rank;
        NC(g_nccl.AllGather(w->d_bar + w->rank, w->d_bar, sizeof(int), ncclChar, w->comm, w->comm_stream));      // every rank is here: buffer b of every rank may be overwritten
        CU(cudaEventRecord(w->ev_fan, w->comm_stream));
        for (int k = 1; k < w->nranks; k++) {                                                                     // one stream per peer (a single stream moved 36 MB slots at ~250 GB/s, one copy engine's worth)
            const int p = (w->rank + k) % w->nranks;
            CU(cudaStreamWaitEvent(w->peer_stream[p], w->ev_fan, 0));
            CU(cudaMemcpyAsync(w->peer_gather[b][p] + slot_bytes * (size_t)w->rank, mine, slot_bytes, cudaMemcpyDefault, w->peer_stream[p]));
            CU(cudaEventRecord(w->ev_peer[p], w->peer_stream[p]));
            CU(cudaStreamWaitEvent(w->comm_stream, w->ev_peer[p], 0));
        }
        NC(g_nccl.AllGather(w->d_bar + w->rank, w->d_bar, sizeof(int), ncclChar, w->comm, w->comm_stream));      // every rank's copies have landed
    } else NC(g_nccl.AllGather(w->d_gather[b] + slot_bytes * (size_t)w->rank, w->d_gather[b], slot_bytes, ncclChar, w->comm, w->comm_stream));
    CU(cudaEventRecord(w->ev_gather_done[b], w->comm_stream));
    w->gather_pending[b] = true; w->gather_last = b; w->gather_flip = b ^ 1;
    return NP_OK;
}
int np_world_gathered_snapshot(np_world* w, void** d_ptr, size_t* bytes)
{
    if (!w || !d_ptr || !bytes) return fail(NP_ERR_INVALID, "bad arguments");
    if (!w->gather_slot || w->gather_last < 0) return fail(NP_ERR_INVALID, "no snapshot gather has run");
    CU(cudaSetDevice(w->device));
    CU(cudaEventSynchronize(w->ev_gather_done[w->gather_last]));
    *d_ptr = w->d_gather[w->gather_last]; *bytes = sizeof(SnapObj) * (size_t)w->gather_slot * (size_t)w->nranks;
    return NP_OK;
}
int np_world_gather_mode(const np_world* w) { return !w || !w->gather_slot ? 0 : (w->gather_dma ? 2 : 1); }
int np_world_last_gather_ms(np_world* w, int back, float out[2])
{
    if (!w || !out || back < 0 || back > 1) return fail(NP_ERR_INVALID, "bad arguments");
    out[0] = out[1] = 0.0f;
    if (!w->gather_slot || w->gather_last < 0) return NP_OK;
    const int b = back ? (w->gather_last ^ 1) : w->gather_last;
    if (!w->gather_pending[b]) return NP_OK;
    CU(cudaSetDevice(w->device));
    CU(cudaEventSynchronize(w->ev_gather_done[b]));
    CU(cudaEventElapsedTime(&out[0], w->ev_gather_begin[b], w->ev_gather_done[b]));
    CU(cudaEventElapsedTime(&out[1], w->ev_step_done[b], w->ev_gather_done[b]));
    return NP_OK;
}
int np_world_shard_over_comm(np_world* w)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (!w->comm) return fail(NP_ERR_INVALID, "np_comm_init first");
    int r = ensure_device(w); if (r) return r;
    const int n = (int)w->bodies.size();
    const int slot = std::max(1, (n + w->nranks - 1) / w->nranks);
    const size_t total = (size_t)slot * (size_t)w->nranks;
    CU(cudaStreamSynchronize(w->stream));
    cudaFree(w->d_shard_hit); cudaFree(w->d_shard_flags); cudaFree(w->d_shard_contact); w->d_shard_hit = w->d_shard_flags = nullptr; w->d_shard_contact = nullptr;
    CU(cudaMalloc((void**)&w->d_shard_hit, sizeof(int) * total)); CU(cudaMalloc((void**)&w->d_shard_flags, sizeof(int) * total)); CU(cudaMalloc((void**)&w->d_shard_contact, sizeof(float4) * total));
    CU(cudaMemsetAsync(w->d_shard_hit, 0xff, sizeof(int) * total, w->stream)); CU(cudaMemsetAsync(w->d_shard_flags, 0, sizeof(int) * total, w->stream));
    CU(cudaMemsetAsync(w->d_shard_contact, 0, sizeof(float4) * total, w->stream));
    w->shard_slot = slot;
    const int first = std::min(n, w->rank * slot), count = std::min(n, first + slot) - first;
    return np_world_set_shard(w, first, count, w->d_shard_hit, w->d_shard_flags, w->d_shard_contact);
}
int np_world_step_sharded(np_world* w, float dt)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    if (!w->comm || !w->d_shard_hit) return fail(NP_ERR_INVALID, "np_comm_init and np_world_shard_over_comm first");
    if (w->halo > 0) {
        // Partitioned: integrate own -> the next rank's first `halo` transforms (ncclSend/Recv between neighbours) -> search own -> the
        // previous rank's last `halo` collision records (their partners may be bodies of mine) -> respond own.  No all-gather.
        if (!g_nccl.Send || !g_nccl.Recv) return fail(NP_ERR_INVALID, "libnccl lacks ncclSend / ncclRecv");
        const int n = (int)w->bodies.size(), H = w->halo, slot = w->shard_slot;
        if (H > slot) return fail(NP_ERR_INVALID, "shard halo larger than a rank's slot of bodies");
        auto range = [&](int rk, int* f, int* c) { *f = (int)std::min<long long>(n, (long long)rk * slot); *c = (int)std::min<long long>(n, (long long)*f + slot) - *f; };
        int first, count, nf = n, nc = 0, pf = 0, pc = 0;
        range(w->rank, &first, &count);
        if (w->rank + 1 < w->nranks) range(w->rank + 1, &nf, &nc);
        if (w->rank > 0) range(w->rank - 1, &pf, &pc);
        int r = np_world_step_integrate(w, dt); if (r) return r;
        const int xs = (w->rank > 0 && pc > 0) ? std::min(H, count) : 0, xr = std::min(H, nc);       // transforms: mine to the previous rank, the next rank's to me
        if (xs > 0 || xr > 0) {
            NC(g_nccl.GroupStart());
            ncclResult_t e1 = xs > 0 ? g_nccl.Send(w->d_xf + (size_t)X_COUNT * first, (size_t)X_COUNT * xs, ncclFloat32, w->rank - 1, w->comm, w->stream) : ncclSuccess;
            ncclResult_t e2 = xr > 0 ? g_nccl.Recv(w->d_xf + (size_t)X_COUNT * nf, (size_t)X_COUNT * xr, ncclFloat32, w->rank + 1, w->comm, w->stream) : ncclSuccess;
            NC(g_nccl.GroupEnd());
            if (e1 != ncclSuccess || e2 != ncclSuccess) { w->step_phase = 0; return fail(NP_ERR_CUDA, "ncclSend/Recv of the halo transforms failed"); }
        }
        r = np_world_step_search(w); if (r) return r;
        const int rs = (nc > 0) ? std::min(H, count) : 0, rr = (w->rank > 0 && count > 0) ? std::min(H, pc) : 0;       // records: my last ones to the next rank, the previous rank's last ones to me
        if (rs > 0 || rr > 0) {
            NC(g_nccl.GroupStart());
            ncclResult_t e[4] = { ncclSuccess, ncclSuccess, ncclSuccess, ncclSuccess };
            if (rs > 0) {
                e[0] = g_nccl.Send(w->d_shard_hit + (first + count - rs), (size_t)rs, ncclInt32, w->rank + 1, w->comm, w->stream);
                e[1] = g_nccl.Send(w->d_shard_contact + (first + count - rs), (size_t)rs * 4, ncclFloat32, w->rank + 1, w->comm, w->stream);
            }
            if (rr > 0) {
                e[2] = g_nccl.Recv(w->d_shard_hit + (first - rr), (size_t)rr, ncclInt32, w->rank - 1, w->comm, w->stream);
                e[3] = g_nccl.Recv(w->d_shard_contact + (first - rr), (size_t)rr * 4, ncclFloat32, w->rank - 1, w->comm, w->stream);
            }
            NC(g_nccl.GroupEnd());
            for (ncclResult_t x : e) if (x != ncclSuccess) { w->mid_step = false; return fail(NP_ERR_CUDA, "ncclSend/Recv of the halo collision records failed"); }
        }
        return np_world_step_end(w);
    }
    int r = np_world_step_begin(w, dt); if (r) return r;
    const size_t slot = (size_t)w->shard_slot, off = slot * (size_t)w->rank;
    NC(g_nccl.GroupStart());
    ncclResult_t e1 = g_nccl.AllGather(w->d_shard_hit + off, w->d_shard_hit, slot, ncclInt32, w->comm, w->stream);
    ncclResult_t e2 = g_nccl.AllGather(w->d_shard_contact + off, w->d_shard_contact, slot * 4, ncclFloat32, w->comm, w->stream);
    NC(g_nccl.GroupEnd());
    if (e1 != ncclSuccess || e2 != ncclSuccess) { w->mid_step = false; return fail(NP_ERR_CUDA, "ncclAllGather of the collision records failed"); }
    return np_world_step_end(w);
}

// ------------------------------------------------------------------------------------------------ device memory helpers
void* np_device_alloc(np_world* w, size_t bytes) { if (!w) return nullptr; void* p = nullptr; if (cudaSetDevice(w->device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { fail(NP_ERR_CUDA, "cudaMalloc failed"); return nullptr; } return p; }
int np_device_free(np_world* w, void* p) { if (!w) return fail(NP_ERR_INVALID, "null world"); CU(cudaSetDevice(w->device)); CU(cudaStreamSynchronize(w->stream)); CU(cudaFree(p)); return NP_OK; }
int np_device_upload(np_world* w, void* d_dst, const void* h_src, size_t bytes) { if (!w) return fail(NP_ERR_INVALID, "null world"); CU(cudaSetDevice(w->device)); CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, w->stream)); CU(cudaStreamSynchronize(w->stream)); return NP_OK; }
int np_device_download(np_world* w, void* h_dst, const void* d_src, size_t bytes) { if (!w) return fail(NP_ERR_INVALID, "null world"); CU(cudaSetDevice(w->device)); CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, w->stream)); CU(cudaStreamSynchronize(w->stream)); return NP_OK; }
void* np_host_alloc(np_world* w, size_t bytes) { if (!w) return nullptr; void* p = nullptr; if (cudaSetDevice(w->device) != cudaSuccess || cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { fail(NP_ERR_CUDA, "cudaMallocHost failed"); return nullptr; } return p; }
int np_host_free(np_world* w, void* p) { if (!w) return fail(NP_ERR_INVALID, "null world"); CU(cudaSetDevice(w->device)); CU(cudaStreamSynchronize(w->stream)); CU(cudaFreeHost(p)); return NP_OK; }
void* np_world_stream(np_world* w) { return w ? (void*)w->stream : nullptr; }
int np_device_flush_l2(np_world* w)
{
    if (!w) return fail(NP_ERR_INVALID, "null world");
    CU(cudaSetDevice(w->device));
    const size_t bytes = 256u << 20;
    if (!w->d_flush) CU(cudaMalloc(&w->d_flush, bytes));
    CU(cudaMemsetAsync(w->d_flush, 0x5a, bytes, w->stream));
    return NP_OK;
}

// ------------------------------------------------------------------------------------------------ self tests
int np_selftest_sincos_hash(np_world* w, uint32_t start, uint64_t count, uint64_t* hash_sin, uint64_t* hash_cos)
{
    if (!w || !hash_sin || !hash_cos) return fail(NP_ERR_INVALID, "bad arguments");
    CU(cudaSetDevice(w->device));
    unsigned long long* d = nullptr;
    CU(cudaMallocAsync((void**)&d, 16, w->stream)); CU(cudaMemsetAsync(d, 0, 16, w->stream));
    k_sincos_hash<<<w->sm_count * 8, 256, 0, w->stream>>>(start, (unsigned long long)count, d, d + 1);
    CU(cudaGetLastError());
    unsigned long long h[2];
    CU(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, w->stream)); CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    *hash_sin = h[0]; *hash_cos = h[1];
    return NP_OK;
}
int np_selftest_fp32_peak(np_world* w, float* tflops)
{
    if (!w || !tflops) return fail(NP_ERR_INVALID, "bad arguments");
    CU(cudaSetDevice(w->device));
    float* d = nullptr; CU(cudaMallocAsync((void**)&d, 4, w->stream));
    const int iters = 20000, blocks = w->sm_count * 8;
    k_fp32_peak<<<blocks, 256, 0, w->stream>>>(100, d);                      // warm-up
    float best = 0.0f;
    for (int rep = 0; rep < 3; rep++) {
        CU(cudaEventRecord(w->ev[0], w->stream));
        k_fp32_peak<<<blocks, 256, 0, w->stream>>>(iters, d);
        CU(cudaEventRecord(w->ev[5], w->stream));
        CU(cudaStreamSynchronize(w->stream));
        float ms = 0.0f; CU(cudaEventElapsedTime(&ms, w->ev[0], w->ev[5]));
        double flop = (double)blocks * 256.0 * iters * 16.0;
        best = std::max(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    CU(cudaFreeAsync(d, w->stream)); CU(cudaGetLastError());
    *tflops = best;
    return NP_OK;
}
int np_selftest_fp32_peaks(np_world* w, float out[2])
{
    if (!w || !out) return fail(NP_ERR_INVALID, "bad arguments");
    int r = np_selftest_fp32_peak(w, &out[0]); if (r) return r;
    float* d = nullptr; CU(cudaMallocAsync((void**)&d, 4, w->stream));
    const int iters = 20000, blocks = w->sm_count * 8;
    k_fp32x2_peak<<<blocks, 256, 0, w->stream>>>(100, d);
    float best = 0.0f;
    for (int rep = 0; rep < 3; rep++) {
        CU(cudaEventRecord(w->ev[0], w->stream));
        k_fp32x2_peak<<<blocks, 256, 0, w->stream>>>(iters, d);
        CU(cudaEventRecord(w->ev[5], w->stream));
        CU(cudaStreamSynchronize(w->stream));
        float ms = 0.0f; CU(cudaEventElapsedTime(&ms, w->ev[0], w->ev[5]));
        best = std::max(best, (float)((double)blocks * 256.0 * iters * 32.0 / (ms * 1e-3) / 1e12));
    }
    CU(cudaFreeAsync(d, w->stream)); CU(cudaGetLastError());
    out[1] = best;
    return NP_OK;
}
#ifdef NP_TIMELINE
int np_debug_dump(np_world* w, unsigned long long* out, int n)
{
    CU(cudaSetDevice(w->device)); CU(cudaStreamSynchronize(w->stream));
    CU(cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * 8 * (size_t)std::min(n, 65536)));
    return NP_OK;
}
#endif
int np_selftest_closest(np_world* w, int kind, const float* pts, int32_t* region, float uvw[3])
{
    if (!w || kind < 2 || kind > 4 || !pts || !region || !uvw) return fail(NP_ERR_INVALID, "bad arguments");
    CU(cudaSetDevice(w->device));
    float* d = nullptr;
    CU(cudaMallocAsync((void**)&d, sizeof(float) * 20, w->stream));
    CU(cudaMemcpyAsync(d, pts, sizeof(float) * 3 * kind, cudaMemcpyHostToDevice, w->stream));
    k_selftest_closest<<<1, 1, 0, w->stream>>>(kind, d, (int*)(d + 12), d + 13);
    CU(cudaGetLastError());
    float h[4];
    CU(cudaMemcpyAsync(h, d + 12, 16, cudaMemcpyDeviceToHost, w->stream)); CU(cudaFreeAsync(d, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    memcpy(region, &h[0], 4); uvw[0] = h[1]; uvw[1] = h[2]; uvw[2] = h[3];
    return NP_OK;
}

#ifdef NP_E2_DEBUG
int np_debug_e2(unsigned long long out[8]) { return cudaMemcpyFromSymbol(out, g_e2dbg, 64) == cudaSuccess ? 0 : -1; }
#endif
}  // extern "C"
