// tests/cpp/comm_two_ranks.cpp -- pure C++ (no Python, no torch) driver of the multi-GPU entry points of include/netphys_b200.h,
// the way the reference's server loop (netphys_server/server.cpp:38-56) would call them: one process per GPU, forked here.
//
//   ./comm_two_ranks <libnetphys_b200.so> [nranks=2] [bodies=4000] [steps=6]
//
// Every rank builds the same world; (1) ONE world sharded over the ranks with np_world_step_sharded (collision records all-gathered by
// the library over NCCL) must stay bit-identical to an unsharded world stepped by rank 0 alone and on every rank; (2) independent
// worlds with np_world_set_snapshot_gather: after a step every rank holds every rank's 36-byte records; (3) the world of (1) once more
// with PARTITIONED steps (np_world_set_shard_halo: own bodies only, ncclSend/Recv of the halos): every rank's own range == (1).
// Exit code 0 = all checks passed.
#include <dlfcn.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../include/netphys_b200.h"

#define SYM(name) decltype(&name) p_##name = (decltype(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }
#define CK(call) do { int r_ = (call); if (r_ < 0) { fprintf(stderr, "rank %d: %s failed: %s\n", rank, #call, p_np_last_error()); return 3; } } while (0)

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
static float urand(uint32_t& s, float lo, float hi) { return lo + (hi - lo) * (float)lcg(s) / 16777216.0f; }

static int run_rank(void* lib, int rank, int nranks, const unsigned char* id, int bodies, int steps, int out_fd)
{
    SYM(np_last_error) SYM(np_world_create) SYM(np_world_destroy) SYM(np_shape_create_box) SYM(np_shape_create_sphere) SYM(np_body_create_batch)
    SYM(np_comm_init) SYM(np_world_shard_over_comm) SYM(np_world_step_sharded) SYM(np_world_step) SYM(np_world_get_states) SYM(np_world_synchronize)
    SYM(np_world_set_snapshot_gather) SYM(np_world_gathered_snapshot) SYM(np_device_download) SYM(np_world_snapshot) SYM(np_comm_destroy)
    SYM(np_world_set_shard_halo) SYM(np_world_gather_mode)
    np_world* w = p_np_world_create(rank);
    if (!w) { fprintf(stderr, "rank %d: np_world_create: %s\n", rank, p_np_last_error()); return 3; }
    auto build = [&](np_world* x, uint32_t seed) {
        int32_t box = p_np_shape_create_box(x, 1.0f, 1.0f, 1.0f), sph = p_np_shape_create_sphere(x, 0.5f);
        std::vector<int32_t> shp(bodies); std::vector<np_static_physics_data> sd(bodies);
        int side = 1; while (side * side * side < bodies) side++;
        uint32_t s = seed;
        for (int i = 0; i < bodies; i++) {
            np_static_physics_data& d = sd[i]; memset(&d, 0, sizeof d);
            d.gravity[1] = -9.8f;
            d.initial_position[0] = 1.3f * (i % side) + urand(s, -0.3f, 0.3f); d.initial_position[1] = 1.3f * ((i / side) % side) + urand(s, -0.3f, 0.3f);
            d.initial_position[2] = 1.3f * (i / (side * side)) + urand(s, -0.3f, 0.3f);
            for (int k = 0; k < 3; k++) d.initial_rotation[k] = urand(s, -1.0f, 1.0f);
            d.mass = 1.0f; d.elasticity = 0.4f; d.collision_response = NP_COLLISION_RESPONSE_IMPULSE; d.moment_of_inertia = 1.0f / 6.0f;
            for (int k = 0; k < 3; k++) { d.inertia_tensor[4 * k] = 1.0f / 6.0f; d.inverse_inertia_tensor[4 * k] = 6.0f; }
            shp[i] = (lcg(s) % 5 == 0) ? sph : box;
        }
        return p_np_body_create_batch(x, bodies, shp.data(), sd.data());
    };
    const float dt = (1000.f / 30.f) / 1000.f;
    // (1) one world, search sharded over the ranks
    CK(build(w, 12345u));
    CK(p_np_comm_init(w, nranks, rank, id));
    CK(p_np_world_shard_over_comm(w));
    for (int k = 0; k < steps; k++) CK(p_np_world_step_sharded(w, dt));
    std::vector<np_body_state> st(bodies); CK(p_np_world_get_states(w, st.data(), bodies));
    int ok = 1;
    if (rank == 0) {                       // the same world, unsharded, on this GPU
        np_world* u = p_np_world_create(0); if (!u) return 3;
        CK(build(u, 12345u));
        for (int k = 0; k < steps; k++) CK(p_np_world_step(u, dt));
        std::vector<np_body_state> su(bodies); CK(p_np_world_get_states(u, su.data(), bodies));
        ok = memcmp(st.data(), su.data(), sizeof(np_body_state) * (size_t)bodies) == 0;
        if (!ok) fprintf(stderr, "sharded state differs from the unsharded world\n");
        p_np_world_destroy(u);
    }
    uint64_t h = 1469598103934665603ull; for (size_t i = 0; i < sizeof(np_body_state) * (size_t)bodies; i++) h = (h ^ ((const unsigned char*)st.data())[i]) * 1099511628211ull;
    CK(p_np_comm_destroy(w)); p_np_world_destroy(w);
    // (2) independent worlds, snapshot gather
    np_world* v = p_np_world_create(rank); if (!v) return 3;
    CK(build(v, 777u + rank));
    CK(p_np_comm_init(v, nranks, rank, id + NP_COMM_ID_BYTES));
    CK(p_np_world_set_snapshot_gather(v, bodies, (uint32_t)(rank * bodies)));
    const int gmode = p_np_world_gather_mode(v);      // 2 = copy engines over CUDA IPC (one process per GPU: expected here), 1 = ncclAllGather
    for (int k = 0; k < 5; k++) CK(p_np_world_step(v, dt));
    void* dptr = nullptr; size_t nbytes = 0; CK(p_np_world_gathered_snapshot(v, &dptr, &nbytes));
    std::vector<np_snapshot_object> all((size_t)bodies * nranks), mine(bodies);
    if (nbytes != all.size() * sizeof(np_snapshot_object)) ok = 0;
    CK(p_np_device_download(v, all.data(), dptr, nbytes));
    CK(p_np_world_snapshot(v, mine.data(), bodies));
    if (memcmp(all.data() + (size_t)rank * bodies, mine.data(), sizeof(np_snapshot_object) * (size_t)bodies) != 0) { ok = 0; fprintf(stderr, "rank %d: own slot of the gather differs from np_world_snapshot\n", rank); }
    for (int r = 0; r < nranks; r++) for (int i = 0; i < bodies; i += 97) if (all[(size_t)r * bodies + i].guid != (uint32_t)(r * bodies + i)) { ok = 0; }
    uint64_t hs = 1469598103934665603ull; for (size_t i = 0; i < nbytes; i++) hs = (hs ^ ((const unsigned char*)all.data())[i]) * 1099511628211ull;
    CK(p_np_comm_destroy(v)); p_np_world_destroy(v);
    // (3) the sharded world of (1) again with PARTITIONED steps: own bodies only, halos between neighbours; the own range must match (1)
    {
        np_world* q = p_np_world_create(rank); if (!q) return 3;
        CK(build(q, 12345u));
        CK(p_np_comm_init(q, nranks, rank, id + 2 * NP_COMM_ID_BYTES));
        CK(p_np_world_shard_over_comm(q));
        const int slot = (bodies + nranks - 1) / nranks, first = std::min(bodies, rank * slot), count = std::min(bodies, first + slot) - first;
        CK(p_np_world_set_shard_halo(q, std::min(256, slot)));
        for (int k = 0; k < steps; k++) CK(p_np_world_step_sharded(q, dt));
        CK(p_np_world_synchronize(q));                     // (reports a halo that was too small)
        std::vector<np_body_state> sq(bodies); CK(p_np_world_get_states(q, sq.data(), bodies));
        if (count > 0 && memcmp(sq.data() + first, st.data() + first, sizeof(np_body_state) * (size_t)count) != 0) { ok = 0; fprintf(stderr, "rank %d: partitioned steps changed the own bodies' state\n", rank); }
        CK(p_np_comm_destroy(q)); p_np_world_destroy(q);
    }
    uint64_t msg[4] = { (uint64_t)ok, h, hs, (uint64_t)gmode };
    if (write(out_fd, msg, sizeof msg) != (ssize_t)sizeof msg) return 4;
    return ok ? 0 : 5;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <libnetphys_b200.so> [nranks] [bodies] [steps]\n", argv[0]); return 2; }
    const int nranks = argc > 2 ? atoi(argv[2]) : 2, bodies = argc > 3 ? atoi(argv[3]) : 4000, steps = argc > 4 ? atoi(argv[4]) : 6;
    void* lib = dlopen(argv[1], RTLD_NOW);                 // (no CUDA call before the fork: the children initialise their own contexts)
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    auto get_id = (decltype(&np_comm_get_unique_id))dlsym(lib, "np_comm_get_unique_id");
    auto dev_count = (decltype(&np_device_count))dlsym(lib, "np_device_count");
    auto last_error = (decltype(&np_last_error))dlsym(lib, "np_last_error");
    if (!get_id || !dev_count) return 2;
    unsigned char id[3 * NP_COMM_ID_BYTES];
    if (get_id(id) < 0 || get_id(id + NP_COMM_ID_BYTES) < 0 || get_id(id + 2 * NP_COMM_ID_BYTES) < 0) { fprintf(stderr, "np_comm_get_unique_id: %s\n", last_error()); return 2; }
    std::vector<pid_t> kids; std::vector<int> fds;
    for (int r = 0; r < nranks; r++) {
        int pfd[2]; if (pipe(pfd)) return 2;
        pid_t pid = fork();
        if (pid == 0) { close(pfd[0]); int rc = run_rank(lib, r, nranks, id, bodies, steps, pfd[1]); _exit(rc); }
        close(pfd[1]); kids.push_back(pid); fds.push_back(pfd[0]);
    }
    int bad = 0; std::vector<uint64_t> hstate, hsnap, gm;
    for (int r = 0; r < nranks; r++) {
        uint64_t msg[4] = { 0, 0, 0, 0 };
        if (read(fds[r], msg, sizeof msg) != (ssize_t)sizeof msg) bad = 1;
        int status = 0; waitpid(kids[r], &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0 || !msg[0]) { fprintf(stderr, "rank %d failed (status %d)\n", r, status); bad = 1; }
        hstate.push_back(msg[1]); hsnap.push_back(msg[2]); gm.push_back(msg[3]);
    }
    for (int r = 1; r < nranks; r++) if (gm[r] != gm[0]) { fprintf(stderr, "rank %d chose another gather mode than rank 0\n", r); bad = 1; }
    for (int r = 1; r < nranks; r++) if (hstate[r] != hstate[0] || hsnap[r] != hsnap[0]) { fprintf(stderr, "rank %d holds different state / gathered snapshot than rank 0\n", r); bad = 1; }
    printf("%s: %d ranks, %d bodies, %d sharded steps: state hash %016llx on every rank == unsharded (and own ranges == after partitioned steps); gathered snapshot hash %016llx on every rank, gather mode %d\n",
           bad ? "FAILED" : "ok", nranks, bodies, steps, (unsigned long long)hstate[0], (unsigned long long)hsnap[0], (int)gm[0]);
    return bad;
}
