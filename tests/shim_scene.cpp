// tests/shim_scene.cpp -- reference-style client code (the scene of engine/test_physics.cpp:71-106, written the way that
// file writes it) compiled against include/netphys/physics.h.  Prints per-step state bits; tests/test_gpu_shim.py compares
// them with the oracle.  Build: see tests/test_abi.py::test_cpp_shim_compiles_against_reference_style_code.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "netphys/physics.h"
#include "netphys/physics_shape.h"

static struct { MeshPhysicsShape* m_shape = nullptr; Physics* m_phys = nullptr; } s_circle, s_board;

static void CreateCircle()
{
    s_circle.m_shape = new MeshPhysicsShape();
    constexpr float CIRCLE_SIZE = 5.0f;
    s_circle.m_shape->CreateSphere(CIRCLE_SIZE);
    StaticPhysicsData physData;
    physData.m_gravity = { 0.0f, -9.8f, 0.0f };
    physData.m_initialPosition = { 0.0f, 20.0f, 0.0f };
    physData.m_initialRotation = { 0.0f, 20.0f, 0.0f };
    physData.m_mass = 10.0f;
    physData.m_elasticity = 0.4f;
    physData.m_staticFrictionCoeff = 0.4f;
    physData.m_momentOfInertia = 0.4f * physData.m_mass * CIRCLE_SIZE * CIRCLE_SIZE;
    physData.m_inertiaTensor = matrix3(physData.m_momentOfInertia, 0.0f, 0.0f, 0.0f, physData.m_momentOfInertia, 0.0f, 0.0f, 0.0f, physData.m_momentOfInertia);
    physData.m_inverseInertiaTensor = physData.m_inertiaTensor.inv();
    physData.m_collisionResponseType = COLLISION_RESPONSE_IMPULSE;
    s_circle.m_phys = new Physics(s_circle.m_shape, physData);
}
static void CreateBoard()
{
    s_board.m_shape = new MeshPhysicsShape();
    s_board.m_shape->CreateBox(50.f, 50.f, 2.f);
    StaticPhysicsData physData;
    physData.m_elasticity = 0.0f;
    physData.m_gravity = { 0.0f, 0.0f, 0.0f };
    physData.m_collisionResponseType = COLLISION_RESPONSE_NONE;
    physData.m_initialPosition = { 0.0f, -1.0f, 0.0f };
    physData.m_initialRotation = vector3(0.0f);
    s_board.m_phys = new Physics(s_board.m_shape, physData);
}
static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv)
{
    int steps = argc > 1 ? atoi(argv[1]) : 100;
    CreateCircle();
    CreateBoard();
    if (netphys::last_status() != NP_OK) { fprintf(stderr, "netphys: %s\n", np_last_error()); return 2; }
    const float dt = (1000.f / 30.f) / 1000.f;                       // platform.h:19, windows.cpp:390
    for (int k = 0; k < steps; k++) {
        Physics_Update(dt);
        vector3 p = s_circle.m_phys->GetPosition(), r = s_circle.m_phys->GetRotation();
        printf("%d %08x %08x %08x %08x %08x %08x\n", k, bits(p.x), bits(p.y), bits(p.z), bits(r.x), bits(r.y), bits(r.z));
    }
    // a stand-alone DetectCollision on the final poses, as test_3d.cpp drives it
    CollisionParams cp; cp.a = s_circle.m_shape; cp.aTransform = s_circle.m_phys->GetTransform(); cp.b = s_board.m_shape; cp.bTransform = s_board.m_phys->GetTransform();
    CollisionData cd;
    bool hit = DetectCollision(cp, true, &cd);
    printf("detect %d %d %08x %08x %08x %08x\n", hit ? 1 : 0, cd.success ? 1 : 0, bits(cd.penetrationDirection.x), bits(cd.penetrationDirection.y), bits(cd.penetrationDirection.z), bits(cd.depth));
    if (netphys::last_status() != NP_OK) { fprintf(stderr, "netphys: %s\n", np_last_error()); return 3; }
    netphys::destroy_default_world();
    return 0;
}
