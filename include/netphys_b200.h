/* include/netphys_b200.h -- the drop-in boundary: C ABI of the B200 rigid-body hot path.
 *
 * Plain C, plain pointers and sizes, no CUDA or torch types.  Each entry point names the reference
 * interface (dougfrazer/netphys, engine/) it replaces.  The C++ shim include/netphys/physics.h re-declares
 * the reference's Physics / PhysicsShape / Physics_Update / DetectCollision on top of these calls, so
 * code written against engine/physics.h + engine/physics_shape.h recompiles unchanged.
 *
 * Conventions
 *   - every call returns NP_OK (0) or a negative np_status; np_last_error() gives the text.  Nothing aborts
 *     (the reference asserts, engine/collision_detection.cpp:255,321).
 *   - a np_world owns one CUDA device, one stream and all device memory of its bodies and shapes.  One host
 *     thread per world.  np_world_step is asynchronous; every getter synchronises the world's stream.
 *   - vectors are float[3]; matrices are row-major (matrix3: x1 x2 x3 / y1.. ; matrix4: x1..x4 / y1..,
 *     engine/matrix.h:36-42,109-116), i.e. exactly the reference's member order.
 *   - all arithmetic on the path is IEEE binary32 in the reference's operation order (no FMA contraction);
 *     results are bit-identical to the reference built as described in oracle/Makefile.
 *   - there is NO CPU fallback: without a CUDA device np_world_create fails with NP_ERR_NO_DEVICE.
 */
#ifndef NETPHYS_B200_H
#define NETPHYS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NETPHYS_B200_ABI_VERSION 1

typedef enum np_status {
    NP_OK = 0,
    NP_ERR_INVALID = -1,     /* bad handle / index / argument */
    NP_ERR_NO_DEVICE = -2,   /* no CUDA device, or device is not sm_100 */
    NP_ERR_CUDA = -3,        /* a CUDA call failed; see np_last_error() */
    NP_ERR_CAPACITY = -4,    /* output buffer too small, or an internal hard cap was hit */
    NP_ERR_UNSUPPORTED = -5  /* e.g. a custom virtual PhysicsShape */
} np_status;

/* COLLISION_RESPONSE, engine/physics.h:13-17 */
typedef enum np_collision_response { NP_COLLISION_RESPONSE_NONE = 0, NP_COLLISION_RESPONSE_IMPULSE = 1 } np_collision_response;

/* StaticPhysicsData, engine/physics.h:36-51 (same fields, same order) */
typedef struct np_static_physics_data {
    float gravity[3];
    float initial_position[3];
    float initial_rotation[3];          /* Euler XYZ radians, engine/matrix.h:283-288 */
    float mass;
    float elasticity;
    float static_friction_coeff;        /* read only under NP_OPT_FRICTION (the reference's friction block is commented out, physics.cpp:49-72) */
    float dynamic_friction_coeff;
    int32_t collision_response;         /* np_collision_response */
    float moment_of_inertia;
    float inertia_tensor[9];
    float inverse_inertia_tensor[9];
} np_static_physics_data;

/* CollisionData, engine/physics.h:18-25.  depth is 0 (not garbage) when success == 0. */
typedef struct np_collision_data {
    float a_normal[3];
    float b_normal[3];
    float penetration_direction[3];
    float depth;
    int32_t success;
} np_collision_data;

/* Collision, engine/physics.h:26-31, with body indices (creation order) instead of Physics* */
typedef struct np_collision {
    int32_t a, b;
    np_collision_data data;
} np_collision;

/* dynamic state of one body: Physics::m_position, m_rotation, m_linearMomentum, m_angularMomentum (physics.h:87-94) */
typedef struct np_body_state {
    float position[3], rotation[3], linear_momentum[3], angular_momentum[3];
} np_body_state;

/* the replication record of netphys_common/common.h:120-127 (CommandFrameObject, 36 bytes) */
typedef struct np_snapshot_object {
    uint32_t guid;
    float pos[3];
    float rot[4];        /* quaternion (w,x,y,z) of the body's Euler rotation */
    uint32_t is_enabled;
} np_snapshot_object;

typedef struct np_world np_world;

/* pair-search policy of np_world_step */
typedef enum np_broadphase {
    NP_BROADPHASE_EXACT = 0,  /* reference semantics: for each body i ascending, j = i+1.. until the first hit, then stop
                                 (GetCollisions, engine/physics.cpp:116-145).  Collision list identical to the reference. */
    NP_BROADPHASE_GRID = 1,   /* uniform-grid AABB culling first (the TODO at collision_detection.cpp:698 / todo.txt:7);
                                 candidates are then searched in the same ascending-j, first-hit order */
    NP_BROADPHASE_SWEPT = 2   /* the same grid over SWEPT boxes: each body's AABB united with its AABB of the previous step
                                 ("broad-phase should do swept AABB from previous state to avoid tunnelling", todo.txt:8); a superset of
                                 the GRID candidates, tested at the current poses.  The first step after bodies change has no previous box */
} np_broadphase;

typedef enum np_option {
    NP_OPT_BROADPHASE = 1,        /* np_broadphase */
    NP_OPT_GRID_CELL = 2,         /* float bits: grid cell edge; 0 = auto (largest AABB extent) */
    NP_OPT_COUNTERS = 3,          /* 1 = accumulate pair-test / hit / iteration counters (np_world_get_counters) */
    NP_OPT_SNAPSHOT_EACH_STEP = 4,/* 1 = np_world_step also packs the replication snapshot (world_s.cpp:23-51) */
    NP_OPT_PHASE_TIMING = 5,      /* 1 = time the phases of every step with CUDA events (np_world_last_step_ms [0..4]); the events cost a few
                                     microseconds each, so a step runs faster without them.  The total [5] is always measured. */
    NP_OPT_SUPPORT_TABLES = 6,    /* default 1: PhysUtil_GetPointFurthestInDirection (physics_util.cpp:5-35) evaluates only the vertices a per-shape
                                     direction table proves can win (same result bits); 0 = scan every vertex as the reference does */
    NP_OPT_FRICTION = 7           /* default 0 (the reference's behaviour).  1 = ApplyImpulseResponse also runs the friction block the reference
                                     keeps commented out at physics.cpp:49-72, exactly as written there (static_friction_coeff of the body).  In
                                     a world step the collision normals are zero (collision_detection.cpp:533-543 never writes them), so the
                                     block only acts through np_body_apply_impulse_response with a caller's normal */
} np_option;

typedef struct np_counters {
    uint64_t steps;
    uint64_t pair_tests;     /* DetectCollision evaluations */
    uint64_t hits;
    uint64_t gjk_steps;      /* DetectCollisionStep evaluations */
    uint64_t support_pairs;  /* support(A,d) / support(B,-d) pairs, GJK + EPA, incl. the seed */
    uint64_t epa_iters;      /* FindIntersectionPointsStep evaluations */
    uint64_t epa_tri_tests;  /* closest-point-on-triangle evaluations of the EPA face scans */
    uint64_t epa_overflow;   /* pairs re-run on the large-polytope path */
    uint64_t candidates;     /* NP_BROADPHASE_GRID: AABB-overlapping pairs emitted */
    uint64_t support_verts;  /* vertices visited by all support evaluations (27 reference flops each) */
    uint64_t kernel_launches;/* CUDA kernels this world has launched (host-side count) */
} np_counters;

/* ---- library ---------------------------------------------------------------------------------- */
int np_abi_version(void);
const char* np_last_error(void);
int np_device_count(void);                         /* CUDA devices visible; 0 when there is no driver/GPU */

/* ---- world (the process-global body list of engine/physics.cpp:9 made explicit) ------------------ */
np_world* np_world_create(int device);             /* device < 0: current device.  NULL on failure */
void np_world_destroy(np_world* w);
int np_world_set_option(np_world* w, int option, int64_t value);
/* Bodies created after this call form a new ISLAND: an independent world that shares the store.  Pairs are
 * only searched inside an island (RL-style batches of worlds, BASELINE configs[3]).  Returns the island id. */
int np_world_begin_island(np_world* w);
int np_world_body_count(const np_world* w);
int np_world_synchronize(np_world* w);

/* ---- shapes: MeshPhysicsShape, engine/physics_shape.h:62-74 --------------------------------------- */
int np_shape_create_mesh(np_world* w, const float* xyz, int nverts);       /* m_mesh.m_vertexPos, physics_shape.h:28; returns shape id */
int np_shape_create_sphere(np_world* w, float radius);                     /* CreateSphere, physics_shape.cpp:313-316 (162-vertex icosphere) */
int np_shape_create_box(np_world* w, float width, float depth, float height); /* CreateBox, physics_shape.cpp:317-320 (depth ignored, :166) */
int np_shape_vertex_count(const np_world* w, int shape);
int np_shape_get_vertices(const np_world* w, int shape, float* xyz, int cap_verts);
/* PhysicsShape::GetPointFurthestInDirection, physics_shape.h:59 (world = row-major 4x4) */
int np_shape_support(np_world* w, int shape, const float dir[3], const float world[16], float out[3]);

/* ---- bodies: class Physics, engine/physics.h:58-95 ------------------------------------------------ */
int np_body_create(np_world* w, int shape, const np_static_physics_data* sd);            /* Physics::Physics, physics.cpp:11-18; returns body id */
int np_body_create_batch(np_world* w, int n, const int32_t* shapes, const np_static_physics_data* sd); /* n ctor calls; returns first id */
/* Physics::~Physics is a TODO in the reference ("remove from list", physics.cpp:20-23): the body leaves the world as if it had never
 * been created (later bodies keep their relative order); every other body id stays valid, the destroyed id is never reused.
 * np_world_get_states / set_states / snapshot then cover the live bodies in creation order; snapshot guids and np_collision a/b are ids. */
int np_body_destroy(np_world* w, int body);
int np_shape_destroy(np_world* w, int shape);                                            /* fails while a body still uses the shape */
int np_body_reset(np_world* w, int body);                                                /* Physics::Reset, physics.cpp:166-172 */
int np_body_get_state(np_world* w, int body, np_body_state* out);                         /* GetPosition/GetRotation (+ momenta) */
int np_body_set_state(np_world* w, int body, const np_body_state* in);                    /* SetPosition/SetRotation, physics.h:82-85 */
int np_body_get_transform(np_world* w, int body, float out16[16]);                        /* GetTransform, physics.h:70-76 */
int np_body_apply_impulse_response(np_world* w, int body, const np_collision_data* data, const float normal[3]); /* physics.cpp:27-46 */
int np_world_get_states(np_world* w, np_body_state* out, int cap);                        /* all bodies, creation order */
int np_world_set_states(np_world* w, const np_body_state* in, int n);

/* ---- adapter extension: what an ODE-shaped World facade (netphys_common/world.h:8-28; include/netphys/ode_shim/ode/ode.h) needs beyond the
 * engine's API.  NOT reference semantics -- engine/ has gravity only (physics.cpp:178-181) -- defined here and mirrored by the oracle:
 * a force acts during the next step only, as L += F*dt right after the gravity term of Physics::Update (calls accumulate, F = sum);
 * a disabled body is not integrated (it still collides) and reports isEnabled = 0 in the snapshot (world_s.cpp:48). */
int np_body_add_force(np_world* w, int body, const float f[3]);                          /* dBodyAddForce, netphys_common/player.cpp:38 */
int np_body_set_enabled(np_world* w, int body, int enabled);                             /* dBodyEnable / dBodyDisable */
int np_body_is_enabled(np_world* w, int body);                                           /* dBodyIsEnabled, netphys_server/world_s.cpp:48 */
int np_world_clear_bodies(np_world* w);                                                  /* World::Reset, netphys_common/world.cpp:129-141: drops every body, keeps the shapes */

/* ---- the step: Physics_Update, engine/physics.h:99-100 / physics.cpp:148-163 ----------------------- */
int np_world_step(np_world* w, float dt);
int np_world_step_n(np_world* w, float dt, int steps);    /* `steps` consecutive Physics_Update(dt) calls, one submission */
/* One world, pair search sharded over processes (one per GPU).  Every process holds the whole world; this one runs GetCollisions'
 * outer loop (physics.cpp:119) for bodies [first, first + count) only and writes their records into the caller's full-length device
 * arrays: d_hit_j int32[n] (partner b with CollisionData::success in bit 30, or -1), d_contact float[4][n] (penetrationDirection.xyz,
 * depth), and d_flags int32[n], scratch of the searching process that never travels.  Per step:  np_world_step_begin (Physics::Update
 * of every body + the owned part of GetCollisions)  ->  all-gather of d_hit_j and d_contact (20 B/body, the step's only exchange; by the
 * caller, or by np_world_step_sharded over the world's own NCCL communicator)  ->  np_world_step_end (OnCollision over the complete
 * list, snapshot).  State stays bit-identical on every process without being exchanged.  count < 0 = not sharded. */
int np_world_set_shard(np_world* w, int first, int count, void* d_hit_j, void* d_flags, void* d_contact);
int np_world_step_begin(np_world* w, float dt);
int np_world_step_end(np_world* w);
/* ---- multi-GPU behind the C ABI: one process per GPU, NCCL (loaded with dlopen("libnccl.so.2") on first use, no link-time
 * dependency).  The caller distributes the 128-byte id from rank 0 to every rank as it would an ncclUniqueId; what the server's C++
 * loop (netphys_server/server.cpp:38-56) needs to run sharded steps and replicate snapshots without Python. */
#define NP_COMM_ID_BYTES 128
int np_comm_get_unique_id(void* id128);
int np_comm_init(np_world* w, int nranks, int rank, const void* id128);
int np_comm_destroy(np_world* w);
/* Replication snapshot of independent worlds (one per rank): from now on every np_world_step packs this world's 36-byte records
 * (guid = guid_base + body) into this rank's slot of a gather buffer [nranks][slot_bodies] and all-gathers it in place on a
 * high-priority stream of its own, overlapped with the next step (double buffered).  np_world_gathered_snapshot waits for the most
 * recent gather and returns the buffer every rank now holds (valid until the next-but-one step is called).  slot_bodies >= this
 * world's body count; 0 disables.  Collective.  With one process per GPU the gather runs on the COPY ENGINES: every rank maps its peers'
 * buffers (CUDA IPC) and pushes its slot into each with device-to-device copies over NVLink, framed by two 4-byte NCCL exchanges, so it
 * takes no SM from the step it overlaps; ranks that cannot map each other (or NP_GATHER_NCCL=1) use ncclAllGather.
 * np_world_gather_mode: 0 = no gather, 1 = ncclAllGather, 2 = copy engines. */
int np_world_set_snapshot_gather(np_world* w, int slot_bodies, uint32_t guid_base);
int np_world_gather_mode(const np_world* w);
int np_world_gathered_snapshot(np_world* w, void** d_ptr, size_t* bytes);
/* device times (ms) of the all-gather launched by the most recent step (back = 0; waits for it) or by the step before (back = 1):
 * [0] its duration on its own stream, [1] how long after its step's last kernel it finished (what a following step must cover) */
int np_world_last_gather_ms(np_world* w, int back, float out[2]);
/* ONE world sharded over the communicator's ranks: allocates the collision-record arrays (ceil(n / nranks) bodies per rank, padded)
 * and sets this rank's shard; then np_world_step_sharded = step_begin -> ncclAllGather of hit_j and contact (one grouped call on the
 * world's stream) -> step_end. */
int np_world_shard_over_comm(np_world* w);
int np_world_step_sharded(np_world* w, float dt);
/* PARTITIONED sharded step (exact search; SURVEY 8(e)-2's slabs, cut by body index because that is where the reference's partners are):
 * with a halo of H > 0 bodies a rank runs Physics::Update, OnCollision and the snapshot for ITS OWN bodies only -- the per-step work and
 * traffic of everything but the search divide by the rank count -- and exchanges with its two neighbours instead of all-gathering:
 *   np_world_step_integrate   Physics::Update of the own bodies
 *   (the next rank's first H transform records -> this rank's array, right after its own range: np_world_transforms_device, [n][16] floats:
 *    rotation 9, translation 3, then 4 ints of arming data the narrow phase reads next to them -- copy all 64 bytes)
 *   np_world_step_search      GetCollisions of the own bodies; a search may look at most H bodies ahead
 *   (the previous rank's last H collision records -> this rank's d_hit_j / d_contact, right before its own range)
 *   np_world_step_end         OnCollision of the own bodies
 * np_world_step_sharded does all five over the communicator (ncclSend / ncclRecv between neighbours).  A first-hit search that would
 * have needed a partner more than H bodies ahead sets an error that np_world_synchronize / np_world_get_collisions report
 * (NP_ERR_CAPACITY): results are the reference's or the step fails loudly.  State, transforms and records of bodies outside
 * [first - H, first + count + H) are stale on this rank; np_world_get_collisions returns the own bodies' collisions. */
int np_world_set_shard_halo(np_world* w, int halo_bodies);            /* 0 = off: replicated steps, all-gather of the records */
int np_world_step_integrate(np_world* w, float dt);
int np_world_step_search(np_world* w);
int np_world_transforms_device(np_world* w, void** d_xf);
/* Physics::Update for every body only (physics.cpp:151-155), no collision phase */
int np_world_integrate(np_world* w, float dt);
/* collisions of the most recent step, in the reference's list order (ascending a) */
int np_world_get_collisions(np_world* w, np_collision* out, int cap, int* count);
/* host-buffer round trip of one step (the e2e path): upload `in` (may be NULL = keep device state),
 * step, download all states into `out`.  Copies are inside the call. */
int np_world_step_host(np_world* w, float dt, const np_body_state* in, np_body_state* out, int n);
int np_world_get_counters(np_world* w, np_counters* out);
/* device time (ms, CUDA events on the world's stream) of the kernels of the most recent step: [5] total; with NP_OPT_PHASE_TIMING also
 * [0] integrate+transform, [1] broad-phase / world-vertex cache, [2] narrow-phase (GJK+EPA), [3] response (+ snapshot record), [4] - */
int np_world_last_step_ms(np_world* w, float out[6]);

/* ---- narrow phase: DetectCollision, engine/physics.h:103-110 --------------------------------------- */
int np_detect_collision(np_world* w, int shape_a, const float a_transform[16], int shape_b, const float b_transform[16],
                        int is3D, np_collision_data* out /* may be NULL */, int* hit);
/* n independent queries, host buffers; pose = position[3] rotation[3] -> GetTransform.  hit / out may be NULL. */
int np_detect_collision_batch(np_world* w, int n, const int32_t* shape_a, const float* pose_a, const int32_t* shape_b,
                              const float* pose_b, int32_t* hit, np_collision_data* out);
/* the same with is3D == false (engine/test_2d.cpp's mode: 3-point simplex, edge-based EPA, collision_detection.cpp:364-384,590-605) */
int np_detect_collision_batch_2d(np_world* w, int n, const int32_t* shape_a, const float* pose_a, const int32_t* shape_b,
                                 const float* pose_b, int32_t* hit, np_collision_data* out);
/* same over device-resident buffers (SoA on the world's device); kernel_ms = device time of the GJK/EPA kernel */
int np_detect_collision_batch_device(np_world* w, int n, const void* d_shape_a, const void* d_pose_a, const void* d_shape_b,
                                     const void* d_pose_b, void* d_hit, void* d_out, float* kernel_ms);

/* ---- the algorithm one call at a time: the TEST_PROGRAM interface, engine/physics.h:115-127 ------------------------------------
 * engine/test_3d.cpp:57-85 and test_2d.cpp:66,73 drive GJK and EPA step by step on a Simplex THEY hold between calls (engine/simplex.h).
 * The simplex crosses the boundary as plain arrays; one device thread executes each call (an inspection interface, not a hot path). */
typedef struct np_simplex_point { float p[3]; float A[3]; float B[3]; } np_simplex_point;          /* SimplexPoint, simplex.h:5-13 */
typedef struct np_simplex_face { int32_t point_index[3]; float normal[3]; } np_simplex_face;       /* SimplexFace,  simplex.h:16-20 */
enum np_collision_result {                                                                         /* COLLISION_RESULT, physics.h:116-123 */
    NP_COLLISION_RESULT_NONE = 0, NP_COLLISION_RESULT_OVERLAP = 1, NP_COLLISION_RESULT_NO_OVERLAP = 2, NP_COLLISION_RESULT_CONTINUE = 3
};
/* DetectCollisionStep (collision_detection.cpp:620-692): solve the simplex, then either report OVERLAP (and set *contains_origin),
 * NO_OVERLAP, or append one support point and CONTINUE.  verts has room for 4 points; *nverts in 1..4 (the reference asserts). */
int np_gjk_step(np_world* w, int shape_a, const float a_transform[16], int shape_b, const float b_transform[16], int is3D,
                np_simplex_point* verts, int* nverts, int* contains_origin, int* result);
/* FindIntersectionPoints (collision_detection.cpp:549-608): up to max_steps EPA iterations on the caller's polytope (faces as built by
 * Simplex::SetupForEPA/AddFace, stored normals kept), or the separated-shapes branches when !contains_origin.  *found = its return
 * value; `out` (may be NULL) receives penetrationDirection and depth when found.  NP_ERR_CAPACITY when the polytope would outgrow
 * verts_cap / faces_cap or 24 vertices (the reference's std::vector is unbounded; its own loop stops at 20 iterations, :716). */
int np_epa_steps(np_world* w, int shape_a, const float a_transform[16], int shape_b, const float b_transform[16], int max_steps, int is3D,
                 np_simplex_point* verts, int* nverts, int verts_cap, np_simplex_face* faces, int* nfaces, int faces_cap,
                 int contains_origin, np_collision_data* out, int* found);
/* GetSearchDirection (collision_detection.cpp:327-337) of a 1..3 point simplex */
int np_search_direction(np_world* w, const np_simplex_point* verts, int nverts, float out[3]);
/* the normal Simplex::AddFace stores for face (va, vb, vc) (simplex.h:37-53); host arithmetic, same binary32 operations as the device */
void np_simplex_face_normal(const np_simplex_point* verts, int va, int vb, int vc, float out[3]);

/* ---- replication snapshot (netphys_server/world_s.cpp:23-51 -> CommandFrameObject) ----------------- */
int np_world_snapshot(np_world* w, np_snapshot_object* dst, int cap);      /* device -> host */
/* device buffer holding the packed snapshot of the last step (for an NCCL all-gather by the host layer) */
int np_world_snapshot_device(np_world* w, void** d_ptr, size_t* bytes);
/* pack into a caller-owned DEVICE buffer instead (e.g. a torch tensor handed to ncclAllGather); NULL = internal buffer.
 * guid_base is added to the body index (rank offset when worlds are sharded over GPUs). */
int np_world_set_snapshot_target(np_world* w, void* d_dst, uint32_t guid_base);

/* ---- command frames: netphys_server/world_s.cpp:12-108 (the step AFTER the path: what the server sends to clients) ---------- */
/* World_S_Update(now): snapshot every body into a new command frame (ids count from 1) and drop frames older than 5000 ms
 * (MAX_COMMAND_FRAME_TIME, world_s.cpp:21,57-68).  Returns the new frame id. */
int np_world_record_frame(np_world* w, double time_ms);
int np_world_frame_count(const np_world* w);
/* World_S_FillWorldUpdateMessage / FillNewConnectionMessage + Packet::Serialize (common.h:77-93,130-163): writes the newest frame as
 *   int32 packet_id | int32 data_size | int32 frame_id | float64 time_ms | int32 count | count x CommandFrameObject (36 B)
 * packet_id: NP_PACKET_WORLD_STATE_UPDATE or NP_PACKET_NEW_CONNECTION.  Returns the bytes written, 0 when `cap` is too small or no frame exists. */
#define NP_PACKET_WORLD_STATE_UPDATE 2342144
#define NP_PACKET_NEW_CONNECTION 2342341
int np_world_fill_packet(np_world* w, int packet_id, void* buffer, int cap);

/* ---- device memory helpers for callers that keep buffers resident (bench, multi-GPU plumbing) ------- */
void* np_device_alloc(np_world* w, size_t bytes);
int np_device_free(np_world* w, void* p);
int np_device_upload(np_world* w, void* d_dst, const void* h_src, size_t bytes);
int np_device_download(np_world* w, void* h_dst, const void* d_src, size_t bytes);
/* page-locked host buffers: np_world_step_host / get_states / set_states copy straight to and from them (any other host pointer is
 * staged through an internal pinned buffer with one extra memcpy each way) */
void* np_host_alloc(np_world* w, size_t bytes);
int np_host_free(np_world* w, void* p);
void* np_world_stream(np_world* w);                                        /* cudaStream_t */
/* benchmarking hygiene: overwrite a 256 MiB scratch on the world's stream so the next step starts with a cold L2 */
int np_device_flush_l2(np_world* w);

/* ---- self tests of device primitives (used by tests/; cheap) ---------------------------------------- */
/* order-independent checksum of sinf/cosf over float bit patterns [start, start+count) computed on the device */
int np_selftest_sincos_hash(np_world* w, uint32_t start, uint64_t count, uint64_t* hash_sin, uint64_t* hash_cos);
/* n support queries (dirs[n][3], row-major 4x4 world transforms xf16[n][16]) on one shape, each answered by the reference's scan over every
 * vertex and through the shape's direction table (NP_OPT_SUPPORT_TABLES); used[q] = 1 when the table answered.  The two must agree bit for bit. */
int np_selftest_support_table(np_world* w, int shape, int n, const float* dirs, const float* xf16, float* out_scan, float* out_table, int32_t* used);
/* host-side statistics of a shape's direction table: out = { cells per face edge, cells that fall back to the scan, mean candidates per cell,
 * reach (largest |t|_1 the table is valid for) }.  Returns 1 when the shape has a table, 0 when it has none (fewer than 5 or more than 255 vertices). */
int np_shape_support_table_info(np_world* w, int shape, float out[4]);
/* measured non-FMA FP32 issue rate of this device (independent FADD/FMUL chains), in Tflop/s: the roofline
 * denominator of the narrow phase, whose arithmetic may not be contracted into FMAs */
int np_selftest_fp32_peak(np_world* w, float* tflops);
/* [0] the same scalar figure, [1] the chains on packed pairs (sm_100 FADD2 / FFMA2-as-multiply, never fused): the non-FMA FP32 roofline of the support loops */
int np_selftest_fp32_peaks(np_world* w, float out[2]);
/* util.h closest-point solvers on the device: kind 2/3/4 = line/triangle/tetrahedron; pts = kind points, origin query */
int np_selftest_closest(np_world* w, int kind, const float* pts, int32_t* region, float uvw[3]);

#ifdef __cplusplus
}
#endif
#endif /* NETPHYS_B200_H */
