// oracle/ref/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A headless C-ABI driver around the UNMODIFIED reference engine (dougfrazer/netphys,
// engine/*.cpp compiled from where they lie under /root/reference -- see oracle/Makefile).
// It exists so that tests, golden-vector generation and bench.py's cpu_baseline /
// --impl reference legs can execute the reference's own Physics_Update / DetectCollision.
// Nothing in netphys_b200/ (the product) may link or load this.
//
// The reference keeps its body list in a file-static (engine/physics.cpp:9) and its
// dynamic state private (engine/physics.h:87-94); this file is a unity build of
// physics.cpp / physics_util.cpp / physics_shape.cpp so it can clear that list between
// scenes and read the momenta.  collision_detection.cpp is a separate translation unit
// (it needs the one-line guard described in oracle/Makefile).
#define private public
#ifdef REF_PHYSICS_CPP          // the friction variant: physics.cpp with the comment markers of lines 49-72 stripped on the fly (oracle/Makefile)
#include REF_PHYSICS_CPP
#else
#include "physics.cpp"
#endif
#include "physics_util.cpp"
#include "physics_shape.cpp"
#undef private
#include "util.h"
#include "simplex.h"

#include <chrono>
#include <thread>

namespace {
std::vector<MeshPhysicsShape*> g_shapes;
std::vector<Physics*>& bodies() { return s_physicsList; }

matrix4 load16(const float* m) { float t[16]; memcpy(t, m, sizeof t); return matrix4(t); }
void store3(float* o, const vector3& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
vector3 load3(const float* p) { return vector3(p[0], p[1], p[2]); }
}  // namespace

extern "C" {

// ---- shapes (engine/physics_shape.cpp:313-320) ------------------------------------
int ref_shape_sphere(float r) { auto* s = new MeshPhysicsShape(); s->CreateSphere(r); g_shapes.push_back(s); return (int)g_shapes.size() - 1; }
int ref_shape_box(float w, float d, float h) { auto* s = new MeshPhysicsShape(); s->CreateBox(w, d, h); g_shapes.push_back(s); return (int)g_shapes.size() - 1; }
int ref_shape_mesh(const float* xyz, int n)
{
    auto* s = new MeshPhysicsShape();
    for (int i = 0; i < n; i++) s->m_mesh.m_vertexPos.push_back(load3(xyz + 3 * i));  // public member, physics_shape.h:28
    g_shapes.push_back(s);
    return (int)g_shapes.size() - 1;
}
int ref_shape_nverts(int s) { return (int)g_shapes[s]->m_mesh.m_vertexPos.size(); }
void ref_shape_verts(int s, float* out)
{
    const auto& v = g_shapes[s]->m_mesh.m_vertexPos;
    for (size_t i = 0; i < v.size(); i++) store3(out + 3 * i, v[i]);
}
void ref_shapes_clear() { for (auto* s : g_shapes) delete s; g_shapes.clear(); }

// ---- bodies (engine/physics.h:36-95) -----------------------------------------------
void ref_world_clear() { for (auto* b : bodies()) delete b; bodies().clear(); }
int ref_world_count() { return (int)bodies().size(); }
// sd: gravity[3] pos[3] rot[3] mass elasticity staticFriction dynamicFriction momentOfInertia
//     inertia[9] invInertia[9]  = 32 floats ; response = COLLISION_RESPONSE enum value
int ref_body_create(int shape, const float* sd, int response)
{
    StaticPhysicsData d;
    d.m_gravity = load3(sd); d.m_initialPosition = load3(sd + 3); d.m_initialRotation = load3(sd + 6);
    d.m_mass = sd[9]; d.m_elasticity = sd[10]; d.m_staticFrictionCoeff = sd[11]; d.m_dynamicFrictionCoeff = sd[12];
    d.m_momentOfInertia = sd[13];
    const float* a = sd + 14; const float* b = sd + 23;
    d.m_inertiaTensor = matrix3(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
    d.m_inverseInertiaTensor = matrix3(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8]);
    d.m_collisionResponseType = (COLLISION_RESPONSE)response;
    new Physics(g_shapes[shape], d);  // self-registers, physics.cpp:17
    return (int)bodies().size() - 1;
}
// state = pos[3] rot[3] linearMomentum[3] angularMomentum[3]
void ref_body_get_state(int b, float* out)
{
    Physics* p = bodies()[b];
    store3(out, p->m_position); store3(out + 3, p->m_rotation);
    store3(out + 6, p->m_linearMomentum); store3(out + 9, p->m_angularMomentum);
}
void ref_body_set_state(int b, const float* in)
{
    Physics* p = bodies()[b];
    p->m_position = load3(in); p->m_rotation = load3(in + 3);
    p->m_linearMomentum = load3(in + 6); p->m_angularMomentum = load3(in + 9);
}
// Physics::ApplyImpulseResponse (physics.cpp:27-46) with a caller's contact and normal
void ref_body_apply_impulse(int b, const float* pen_dir, float depth, const float* n)
{
    CollisionData d; d.penetrationDirection = load3(pen_dir); d.depth = depth;
    bodies()[b]->ApplyImpulseResponse(d, load3(n));
}
#ifdef REF_PHYSICS_CPP
int ref_has_friction() { return 1; }
#else
int ref_has_friction() { return 0; }
#endif
void ref_world_get_state(float* out) { for (int i = 0; i < ref_world_count(); i++) ref_body_get_state(i, out + 12 * i); }
void ref_body_reset(int b) { bodies()[b]->Reset(); }
void ref_body_transform(int b, float* out16) { matrix4 m = bodies()[b]->GetTransform(); memcpy(out16, &m, 64); }
// matrix3::inv as used by the test scenes (engine/matrix.h:88-103, test_physics.cpp:88)
void ref_matrix3_inv(const float* in9, float* out9)
{
    matrix3 m(in9[0], in9[1], in9[2], in9[3], in9[4], in9[5], in9[6], in9[7], in9[8]);
    matrix3 r = m.inv(); memcpy(out9, &r, 36);
}

// ---- the step (engine/physics.cpp:148-163) ----------------------------------------
void ref_update(float dt) { Physics_Update(dt); }
void ref_update_n(float dt, int steps) { for (int i = 0; i < steps; i++) Physics_Update(dt); }

// Same three phases as Physics_Update, calling the reference's own functions, but keeping
// the collision list so it can be compared.  a/b = body indices; data = 11 floats per
// collision (a_normal, b_normal, penetrationDirection, depth, success).
int ref_update_traced(float dt, int* a, int* b, float* data, int cap)
{
    for (auto* p : bodies()) p->Update(dt);
    std::vector<Collision> c = GetCollisions(bodies());
    int n = 0;
    for (size_t i = 0; i < c.size(); i++)
    {
        if (n < cap)
        {
            a[n] = (int)(std::find(bodies().begin(), bodies().end(), c[i].a) - bodies().begin());
            b[n] = (int)(std::find(bodies().begin(), bodies().end(), c[i].b) - bodies().begin());
            float* o = data + 11 * n;
            store3(o, c[i].data.a_normal); store3(o + 3, c[i].data.b_normal); store3(o + 6, c[i].data.penetrationDirection);
            o[9] = c[i].data.depth; o[10] = c[i].data.success ? 1.0f : 0.0f;
            n++;
        }
        OnCollision(c[i], true);
    }
    return (int)c.size();
}

// ---- narrow phase (engine/collision_detection.cpp:694-732) ------------------------
void ref_transform(const float* pos, const float* rot, float* out16)
{
    matrix4 m; m.translate(load3(pos)); m.rotate(load3(rot));  // physics.h:70-76
    memcpy(out16, &m, 64);
}
// out = penetrationDirection[3] depth success ; returns DetectCollision's bool
int ref_detect(int sa, const float* xa, int sb, const float* xb, int is3D, float* out)
{
    CollisionParams p; p.a = g_shapes[sa]; p.b = g_shapes[sb]; p.aTransform = load16(xa); p.bTransform = load16(xb);
    CollisionData d; d.depth = 0.0f;
    bool hit = DetectCollision(p, is3D != 0, &d);
    store3(out, d.penetrationDirection); out[3] = d.depth; out[4] = d.success ? 1.0f : 0.0f;
    return hit ? 1 : 0;
}
// batch over (shape, pos, rot) per side; poses = 6 floats; out = 5 floats per pair; hit = int per pair
void ref_detect_batch(int n, const int* sa, const float* posea, const int* sb, const float* poseb, int* hit, float* out)
{
    for (int i = 0; i < n; i++)
    {
        float xa[16], xb[16];
        ref_transform(posea + 6 * i, posea + 6 * i + 3, xa);
        ref_transform(poseb + 6 * i, poseb + 6 * i + 3, xb);
        hit[i] = ref_detect(sa[i], xa, sb[i], xb, 1, out + 5 * i);
    }
}
void ref_support(int s, const float* dir, const float* xf, float* out)
{
    vector3 r = g_shapes[s]->GetPointFurthestInDirection(load3(dir), load16(xf), true);
    store3(out, r);
}

// ---- util.h solvers, for the test_util.cpp known-answer vectors --------------------
int ref_closest_line(const float* p, const float* a, const float* b, float* u)
{ return (int)ClosestPoint_LinePointRatio(load3(p), load3(a), load3(b), *u); }
int ref_closest_triangle(const float* p, const float* a, const float* b, const float* c, float* uv)
{ return (int)ClosestPoint_TrianglePointRatio(load3(p), load3(a), load3(b), load3(c), uv[0], uv[1]); }
int ref_closest_tetra(const float* p, const float* a, const float* b, const float* c, const float* d, float* uvw)
{ uvw[0] = uvw[1] = uvw[2] = 0.0f; return (int)ClosestPoint_TetrahedronPointRatio(load3(p), load3(a), load3(b), load3(c), load3(d), uvw[0], uvw[1], uvw[2]); }
void ref_closest_line_point(const float* p, const float* a, const float* b, float* out) { store3(out, ClosestPoint_LinePoint(load3(p), load3(a), load3(b))); }
void ref_closest_triangle_point(const float* p, const float* a, const float* b, const float* c, float* out) { store3(out, ClosestPoint_TrianglePoint(load3(p), load3(a), load3(b), load3(c))); }
float ref_line_point_distance(const float* p, const float* a, const float* b) { return LinePointDistance(load3(p), load3(a), load3(b)); }
void ref_dir_to_origin(int n, const float* pts, float* out)
{
    vector3 r;
    if (n == 1) r = GetDirectionToOrigin(load3(pts));
    else if (n == 2) r = GetDirectionToOrigin(load3(pts), load3(pts + 3));
    else r = GetDirectionToOrigin(load3(pts), load3(pts + 3), load3(pts + 6));
    store3(out, r);
}
int ref_float_equals(float a, float b, float tol) { return FloatEquals(a, b, tol) ? 1 : 0; }
// the libm entry points matrix4::rotate resolves to under `-include math.h`
float ref_sinf(float x) { return sin(x); }
float ref_cosf(float x) { return cos(x); }
void ref_sincos_batch(const float* x, long n, float* s, float* c) { for (long i = 0; i < n; i++) { s[i] = sin(x[i]); c[i] = cos(x[i]); } }

// ---- timing helpers for bench.py (--impl reference, cpu_baseline) -----------------
double ref_time_update(float dt, int steps)
{
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; i++) Physics_Update(dt);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// DetectCollision over a pair list with nthreads std::threads (the reference itself is
// single-threaded; DetectCollision touches no global state, so this is embarrassingly parallel)
double ref_time_detect(int n, const int* sa, const float* posea, const int* sb, const float* poseb, int* hit, float* out, int nthreads)
{
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads <= 1) ref_detect_batch(n, sa, posea, sb, poseb, hit, out);
    else
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
        {
            int lo = (int)((long)n * t / nthreads), hi = (int)((long)n * (t + 1) / nthreads);
            th.emplace_back([=]() { ref_detect_batch(hi - lo, sa + lo, posea + 6 * lo, sb + lo, poseb + 6 * lo, hit + lo, out + 5 * lo); });
        }
        for (auto& x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"

// ---- the TEST_PROGRAM step interface (engine/physics.h:115-127), driven the way engine/test_3d.cpp:36-85 drives it ------------
// The client code is tests/cpp/step_trace_body.h, compiled here against the reference's own headers and sources.
#include "../../tests/cpp/step_trace_body.h"
extern "C" int ref_step_trace(const char* in_path, const char* out_path, int max_calls) { return step_trace::run(in_path, out_path, max_calls); }
