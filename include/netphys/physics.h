// include/netphys/physics.h -- source-compatible C++ shim of the reference engine's body-and-shape API
// (dougfrazer/netphys engine/physics.h + engine/physics_shape.h) on top of the C ABI in netphys_b200.h.
//
// Code written against the reference -- engine/test_physics.cpp:71-106 is the model -- recompiles against this
// header and links libnetphys_b200.so instead of engine/physics.cpp + collision_detection.cpp:
//
//     MeshPhysicsShape* shape = new MeshPhysicsShape();  shape->CreateSphere(5.0f);
//     StaticPhysicsData d;  d.m_gravity = { 0.0f, -9.8f, 0.0f };  ...  d.m_inverseInertiaTensor = d.m_inertiaTensor.inv();
//     Physics* body = new Physics(shape, d);
//     Physics_Update(dt);                      // runs on the B200
//     vector3 p = body->GetPosition();
//
// Same class names, member names, signatures and ownership rules (the caller news and keeps shapes and bodies; the
// engine keeps a process-global body list, engine/physics.cpp:9 -- here the library's default world).  Differences:
// errors are reported through netphys::last_status() instead of assert(); only vertex-list shapes are supported
// (MeshPhysicsShape is the only PhysicsShape the reference implements, physics_shape.h:76-91).
// The small vector3 / matrix3 / matrix4 value types below carry just what callers of this API touch.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "../netphys_b200.h"

namespace netphys {
inline np_world*& default_world_slot() { static np_world* w = nullptr; return w; }
inline int& last_status() { static int s = NP_OK; return s; }
inline np_world* default_world()
{
    np_world*& w = default_world_slot();
    if (!w) { w = np_world_create(-1); if (!w) last_status() = NP_ERR_NO_DEVICE; }
    return w;
}
inline void destroy_default_world() { np_world*& w = default_world_slot(); if (w) { np_world_destroy(w); w = nullptr; } }
inline int check(int r) { if (r < 0) last_status() = r; return r; }
}  // namespace netphys

// ---- engine/vector.h, engine/matrix.h: the subset the physics API exposes ----------------------------------------
class vector3 {
public:
    float x, y, z;
    vector3() : x(0.0f), y(0.0f), z(0.0f) {}
    vector3(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
    explicit vector3(float v) : x(v), y(v), z(v) {}
    vector3 operator-() const { return vector3(-x, -y, -z); }
    vector3 operator-(const vector3 b) const { return vector3(x - b.x, y - b.y, z - b.z); }
    vector3 operator+(const vector3 b) const { return vector3(x + b.x, y + b.y, z + b.z); }
    vector3 operator*(float s) const { return vector3(x * s, y * s, z * s); }
    float dot(const vector3& b) const { return x * b.x + y * b.y + z * b.z; }
    float magnitude() const { return std::sqrt(x * x + y * y + z * z); }
};
namespace Coordinates {                           // vector.h:248-256
inline vector3 GetUp() { return vector3(0.0f, 1.0f, 0.0f); }
inline vector3 GetDown() { return vector3(0.0f, -1.0f, 0.0f); }
inline vector3 GetLeft() { return vector3(0.0f, 0.0f, -1.0f); }
inline vector3 GetRight() { return vector3(0.0f, 0.0f, 1.0f); }
inline vector3 GetIn() { return vector3(1.0f, 0.0f, 0.0f); }
inline vector3 GetOut() { return vector3(-1.0f, 0.0f, 0.0f); }
}  // namespace Coordinates
class matrix3 {
public:
    float x1, x2, x3, y1, y2, y3, z1, z2, z3;
    matrix3() : x1(1), x2(0), x3(0), y1(0), y2(1), y3(0), z1(0), z2(0), z3(1) {}
    matrix3(float a1, float a2, float a3, float b1, float b2, float b3, float c1, float c2, float c3)
        : x1(a1), x2(a2), x3(a3), y1(b1), y2(b2), y3(b3), z1(c1), z2(c2), z3(c3) {}
    float det() const { return x1 * y2 * z3 - x1 * y3 * z2 + x2 * y3 * z1 - x2 * y1 * z3 + x3 * y1 * z2 - x3 * y2 * z1; }
    // adjugate / determinant, in the operation order of engine/matrix.h:88-103 so inertia tensors get the same bits
    matrix3 inv() const
    {
        const float d = det();
        auto m2 = [](float a, float b, float c, float e) { return a * e - b * c; };
        return matrix3(m2(y2, y3, z2, z3) / d, m2(x3, x2, z3, z2) / d, m2(x2, x3, y2, y3) / d,
                       m2(y3, y1, z3, z1) / d, m2(x1, x3, z1, z3) / d, m2(x3, x1, y3, y1) / d,
                       m2(y1, y2, z1, z2) / d, m2(x2, x1, z2, z1) / d, m2(x1, x2, y1, y2) / d);
    }
};
class matrix4 {
public:
    float x1, x2, x3, x4, y1, y2, y3, y4, z1, z2, z3, z4, w1, w2, w3, w4;   // rows, as engine/matrix.h:109-116
    matrix4() : x1(1), x2(0), x3(0), x4(0), y1(0), y2(1), y3(0), y4(0), z1(0), z2(0), z3(1), z4(0), w1(0), w2(0), w3(0), w4(1) {}
    explicit matrix4(const float in[16]) { std::memcpy(&x1, in, 64); }
    // matrix.h:145-157: the point is promoted to (x, y, z, 1) and every row summed left to right; the w row is dropped by vector3(vector4)
    vector3 operator*(const vector3& B) const
    {
        return vector3(x1 * B.x + x2 * B.y + x3 * B.z + x4 * 1.0f, y1 * B.x + y2 * B.y + y3 * B.z + y4 * 1.0f, z1 * B.x + z2 * B.y + z3 * B.z + z4 * 1.0f);
    }
    void transpose()
    {
        matrix4 t = *this;
        x2 = t.y1; x3 = t.z1; x4 = t.w1; y1 = t.x2; y3 = t.z2; y4 = t.w2; z1 = t.x3; z2 = t.y3; z4 = t.w3; w1 = t.x4; w2 = t.y4; w3 = t.z4;
    }
};

// ---- engine/physics_shape.h ---------------------------------------------------------------------------------------
class Mesh {
public:
    std::vector<vector3> m_vertexPos;           // physics_shape.h:28 -- the support mapping scans exactly this list
};
class PhysicsShape {
public:
    virtual ~PhysicsShape() {}
    virtual vector3 GetPointFurthestInDirection(const vector3& dir, const matrix4& world, bool is3D) const = 0;   // physics_shape.h:59
    virtual int DeviceShape() const = 0;        // shim: id of the vertex list on the device
};
class MeshPhysicsShape : public PhysicsShape {
public:
    void CreateSphere(float radius) { adopt(np_shape_create_sphere(netphys::default_world(), radius)); }            // physics_shape.cpp:313-316
    void CreateBox(float width, float depth, float height) { adopt(np_shape_create_box(netphys::default_world(), width, depth, height)); }   // :317-320
    vector3 GetPointFurthestInDirection(const vector3& dir, const matrix4& world, bool) const override
    {
        vector3 out;
        netphys::check(np_shape_support(netphys::default_world(), DeviceShape(), &dir.x, &world.x1, &out.x));
        return out;
    }
    int DeviceShape() const override
    {
        if (m_shape < 0 || m_uploaded != m_mesh.m_vertexPos.size())           // vertices pushed by hand (physics_shape.h:73 is public)
            const_cast<MeshPhysicsShape*>(this)->upload();
        return m_shape;
    }
    Mesh m_mesh;

private:
    void adopt(int id)
    {
        if (netphys::check(id) < 0) return;
        m_shape = id;
        int n = np_shape_vertex_count(netphys::default_world(), id);
        m_mesh.m_vertexPos.resize(n > 0 ? n : 0);
        if (n > 0) np_shape_get_vertices(netphys::default_world(), id, &m_mesh.m_vertexPos[0].x, n);
        m_uploaded = m_mesh.m_vertexPos.size();
    }
    void upload()
    {
        static_assert(sizeof(vector3) == 12, "vector3 must be three packed floats");
        m_shape = netphys::check(np_shape_create_mesh(netphys::default_world(), m_mesh.m_vertexPos.empty() ? nullptr : &m_mesh.m_vertexPos[0].x, (int)m_mesh.m_vertexPos.size()));
        m_uploaded = m_mesh.m_vertexPos.size();
    }
    int m_shape = -1;
    size_t m_uploaded = 0;
};

// ---- engine/physics.h ---------------------------------------------------------------------------------------------
enum COLLISION_RESPONSE { COLLISION_RESPONSE_NONE, COLLISION_RESPONSE_IMPULSE };
struct CollisionData {
    vector3 a_normal, b_normal, penetrationDirection;
    float depth = 0.0f;
    bool success = false;
};
struct StaticPhysicsData {
    vector3 m_gravity, m_initialPosition, m_initialRotation;
    float m_mass = 0.0f, m_elasticity = 0.0f, m_staticFrictionCoeff = 0.0f, m_dynamicFrictionCoeff = 0.0f;
    COLLISION_RESPONSE m_collisionResponseType = COLLISION_RESPONSE_NONE;
    float m_momentOfInertia = 0.0f;
    matrix3 m_inertiaTensor, m_inverseInertiaTensor;
};
class Physics {
public:
    Physics(PhysicsShape* shape, const StaticPhysicsData& d) : m_physicsShape(shape), m_static(d)                   // physics.cpp:11-18
    {
        np_static_physics_data s;
        std::memcpy(s.gravity, &d.m_gravity.x, 12); std::memcpy(s.initial_position, &d.m_initialPosition.x, 12); std::memcpy(s.initial_rotation, &d.m_initialRotation.x, 12);
        s.mass = d.m_mass; s.elasticity = d.m_elasticity; s.static_friction_coeff = d.m_staticFrictionCoeff; s.dynamic_friction_coeff = d.m_dynamicFrictionCoeff;
        s.collision_response = (int)d.m_collisionResponseType; s.moment_of_inertia = d.m_momentOfInertia;
        std::memcpy(s.inertia_tensor, &d.m_inertiaTensor.x1, 36); std::memcpy(s.inverse_inertia_tensor, &d.m_inverseInertiaTensor.x1, 36);
        m_body = netphys::check(np_body_create(netphys::default_world(), shape->DeviceShape(), &s));
    }
    ~Physics() {}                                                           // the reference never unregisters either (physics.cpp:20-23)
    void Reset() { netphys::check(np_body_reset(netphys::default_world(), m_body)); }
    vector3 GetPosition() const { np_body_state s = state(); return vector3(s.position[0], s.position[1], s.position[2]); }
    vector3 GetRotation() const { np_body_state s = state(); return vector3(s.rotation[0], s.rotation[1], s.rotation[2]); }
    matrix4 GetTransform() const { float m[16] = { 0 }; netphys::check(np_body_get_transform(netphys::default_world(), m_body, m)); return matrix4(m); }
    COLLISION_RESPONSE GetCollisionResponse() const { return m_static.m_collisionResponseType; }
    const PhysicsShape* GetPhysicsShape() const { return m_physicsShape; }
    void ApplyImpulseResponse(CollisionData& data, vector3 normal)                                                  // physics.cpp:27-46
    {
        np_collision_data c; std::memset(&c, 0, sizeof c);
        std::memcpy(c.penetration_direction, &data.penetrationDirection.x, 12); c.depth = data.depth; c.success = data.success;
        netphys::check(np_body_apply_impulse_response(netphys::default_world(), m_body, &c, &normal.x));
    }
    void SetPosition(const vector3& v) { np_body_state s = state(); std::memcpy(s.position, &v.x, 12); netphys::check(np_body_set_state(netphys::default_world(), m_body, &s)); }
    void SetRotation(const vector3& v) { np_body_state s = state(); std::memcpy(s.rotation, &v.x, 12); netphys::check(np_body_set_state(netphys::default_world(), m_body, &s)); }
    int DeviceBody() const { return m_body; }

private:
    np_body_state state() const { np_body_state s; std::memset(&s, 0, sizeof s); netphys::check(np_body_get_state(netphys::default_world(), m_body, &s)); return s; }
    const PhysicsShape* m_physicsShape;
    const StaticPhysicsData m_static;
    int m_body = -1;
};

// C-style interface, engine/physics.h:99-110
inline void Physics_Update(float dt) { netphys::check(np_world_step(netphys::default_world(), dt)); }
struct CollisionParams {
    const PhysicsShape* a;
    const PhysicsShape* b;
    matrix4 aTransform, bTransform;
};
inline bool DetectCollision(const CollisionParams& p, bool is3D, CollisionData* out)
{
    np_collision_data c; int hit = 0;
    if (netphys::check(np_detect_collision(netphys::default_world(), p.a->DeviceShape(), &p.aTransform.x1, p.b->DeviceShape(), &p.bTransform.x1, is3D ? 1 : 0, out ? &c : nullptr, &hit)) < 0) return false;
    if (out) {
        out->a_normal = vector3(); out->b_normal = vector3();
        out->penetrationDirection = vector3(c.penetration_direction[0], c.penetration_direction[1], c.penetration_direction[2]);
        out->depth = c.depth; out->success = c.success != 0;
    }
    return hit != 0;
}


// ---- engine/simplex.h + the TEST_PROGRAM interface, engine/physics.h:115-127 -----------------------------------------------------
// engine/test_3d.cpp:36-85 and test_2d.cpp:52-75 single-step GJK and EPA on a Simplex they keep between calls.  The container lives on
// the host exactly as in the reference (two std::vectors the caller may resize, clear and index); every call ships it to the device,
// where one thread runs the step, and back.  Always declared: the reference hides these behind #ifdef TEST_PROGRAM only to keep them
// out of its game build.
struct SimplexPoint {
    vector3 p, A, B;
};
struct SimplexFace {
    int point_index[3];
    vector3 normal;
};
struct Simplex {
    bool m_containsOrigin = false;
    std::vector<SimplexPoint> verts;
    std::vector<SimplexFace> faces;
    int size() const { return (int)verts.size(); }
    void AddFace(int va, int vb, int vc)                                   // simplex.h:37-53; the normal is computed by the library
    {
        static_assert(sizeof(SimplexPoint) == sizeof(np_simplex_point) && sizeof(SimplexFace) == sizeof(np_simplex_face), "simplex layouts");
        SimplexFace f; f.point_index[0] = va; f.point_index[1] = vb; f.point_index[2] = vc;
        np_simplex_face_normal(reinterpret_cast<const np_simplex_point*>(verts.data()), va, vb, vc, &f.normal.x);
        faces.push_back(f);
    }
    void SetupForEPA() { AddFace(0, 1, 2); AddFace(0, 2, 3); AddFace(3, 1, 0); AddFace(1, 3, 2); }   // simplex.h:29-36
    int insert(const SimplexPoint& point, int index = -1)                  // simplex.h:59-75
    {
        if (index < 0) { verts.push_back(point); return (int)verts.size() - 1; }
        verts.insert(verts.begin() + index, point);
        return index;
    }
};
enum COLLISION_RESULT { COLLISION_RESULT_NONE, COLLISION_RESULT_OVERLAP, COLLISION_RESULT_NO_OVERLAP, COLLISION_RESULT_CONTINUE };
inline COLLISION_RESULT DetectCollisionStep(const CollisionParams& p, bool is3D, Simplex& simplex)               // collision_detection.cpp:620-692
{
    np_simplex_point v[4]; int n = simplex.size(), contains = simplex.m_containsOrigin ? 1 : 0, result = COLLISION_RESULT_NONE;
    if (n < 1 || n > 4) { netphys::last_status() = NP_ERR_INVALID; return COLLISION_RESULT_NONE; }
    std::memcpy(v, simplex.verts.data(), sizeof(np_simplex_point) * (size_t)n);
    if (netphys::check(np_gjk_step(netphys::default_world(), p.a->DeviceShape(), &p.aTransform.x1, p.b->DeviceShape(), &p.bTransform.x1, is3D ? 1 : 0,
                                   v, &n, &contains, &result)) < 0) return COLLISION_RESULT_NONE;
    simplex.verts.resize((size_t)n);
    std::memcpy(simplex.verts.data(), v, sizeof(np_simplex_point) * (size_t)n);
    simplex.m_containsOrigin = contains != 0;
    return (COLLISION_RESULT)result;
}
inline bool FindIntersectionPoints(const CollisionParams& p, Simplex& simplex, int max_steps, bool is3D, CollisionData* out)   // :549-608
{
    // room for one new vertex per iteration; an expansion creates one face per horizon edge, at most three per face it removes
    const int nv0 = simplex.size(), nf0 = (int)simplex.faces.size();
    const int vcap = nv0 + (max_steps > 0 ? max_steps : 0);
    int fcap = nf0 + 4;
    for (int k = 0; k < max_steps && fcap < 8192; k++) fcap = 4 * fcap;
    if (fcap > 8192) fcap = 8192;
    int nv = nv0, nf = nf0, found = 0;
    simplex.verts.resize((size_t)vcap); simplex.faces.resize((size_t)fcap);
    np_collision_data c; std::memset(&c, 0, sizeof c);
    int r = netphys::check(np_epa_steps(netphys::default_world(), p.a->DeviceShape(), &p.aTransform.x1, p.b->DeviceShape(), &p.bTransform.x1, max_steps, is3D ? 1 : 0,
                                        reinterpret_cast<np_simplex_point*>(simplex.verts.data()), &nv, vcap,
                                        reinterpret_cast<np_simplex_face*>(simplex.faces.data()), &nf, fcap, simplex.m_containsOrigin ? 1 : 0, &c, &found));
    if (r < 0) { nv = nv0; nf = nf0; found = 0; }
    simplex.verts.resize((size_t)nv); simplex.faces.resize((size_t)nf);
    if (found && out) {
        out->penetrationDirection = vector3(c.penetration_direction[0], c.penetration_direction[1], c.penetration_direction[2]);
        out->depth = c.depth;
    }
    return found != 0;
}
inline vector3 GetSearchDirection(const Simplex& simplex)                                                          // :327-337
{
    vector3 d;
    netphys::check(np_search_direction(netphys::default_world(), reinterpret_cast<const np_simplex_point*>(simplex.verts.data()), simplex.size(), &d.x));
    return d;
}
