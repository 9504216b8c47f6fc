/* include/netphys/ode_shim/ode/ode.h -- the part of the ODE C API that dougfrazer/netphys calls, implemented on the netphys-b200 C ABI.
 *
 * the server's world step does not call engine/ today; it calls ODE through World (netphys_common/world.h:8-28,
 * world.cpp, worldobject.cpp:7-20, player.cpp:9-38, netphys_server/world_s.cpp:23-51, player_s.cpp:33-45).  Compiling those files with
 * -I<repo>/include/netphys/ode_shim -I<repo>/include (instead of ODE's include directory) and linking libnetphys_b200.so makes that
 * step run on this library, source unchanged.  What changes is the PHYSICS, deliberately: the world is stepped by the engine's path
 * (Physics_Update: semi-implicit Euler, ordered first-hit GJK/EPA, the reference's impulse response), not by ODE's LCP solver -- ODE
 * is an absent third-party dependency with no pinned results.  dSpaceCollide is therefore a no-op (the pair search
 * is inside the step), contact joints are never created, damping / CFM / auto-disable settings are accepted and ignored.
 *
 * Mapping.  dBodyID = one Physics body, created in the library when it has both a geom and a mass (dGeomSetBody / dBodySetMass, in
 * either order) or at the next step; dGeomID = one vertex-list shape (box: its 8 corners; sphere: the engine's 162-vertex icosphere;
 * plane: a static 200 x 200 x 1 slab under the plane).  dBodyAddForce -> np_body_add_force (one-step force, L += F*dt).
 * dBodyGetPosition / dBodyGetQuaternion / dBodyIsEnabled read one bulk np_world_snapshot per step, not one device access per body.
 * dBodyDestroy disables the body; when every body of the world has been destroyed (World::Reset, world.cpp:129-141) the store is
 * cleared.  Rotations are the engine's Euler XYZ angles; quaternions (w,x,y,z) come from the snapshot kernel.
 */
#ifndef NETPHYS_ODE_SHIM_H
#define NETPHYS_ODE_SHIM_H

#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "../../../netphys_b200.h"

typedef double dReal;                                    /* netphys_common/common.h:3-5 builds ODE with dDOUBLE */
typedef dReal dVector3[4];
typedef dReal dVector4[4];
typedef dReal dMatrix3[12];
typedef dReal dQuaternion[4];
#define dInfinity ((dReal)INFINITY)

struct dMass { dReal mass; dVector3 c; dMatrix3 I; };
struct dxBody; struct dxGeom; struct dxWorld; struct dxSpace; struct dxJoint; struct dxJointGroup;
struct dxThreadingImplementation; struct dxThreadingThreadPool; struct dxThreadingFunctionsInfo;
typedef dxBody* dBodyID; typedef dxGeom* dGeomID; typedef dxWorld* dWorldID; typedef dxSpace* dSpaceID; typedef dxJoint* dJointID;
typedef dxJointGroup* dJointGroupID; typedef dxThreadingImplementation* dThreadingImplementationID; typedef dxThreadingThreadPool* dThreadingThreadPoolID;

enum { dContactMu2 = 0x001, dContactFDir1 = 0x002, dContactBounce = 0x004, dContactSoftERP = 0x008, dContactSoftCFM = 0x010, dContactMotion1 = 0x020, dContactMotion2 = 0x040 };
enum { dJointTypeContact = 4 };
enum { dAllocateFlagBasicData = 0, dAllocateMaskAll = ~0 };
struct dSurfaceParameters { int mode; dReal mu, mu2, rho, rho2, rhoN, bounce, bounce_vel, soft_erp, soft_cfm, motion1, motion2, motionN, slip1, slip2; };
struct dContactGeom { dVector3 pos, normal; dReal depth; dGeomID g1, g2; int side1, side2; };
struct dContact { dSurfaceParameters surface; dContactGeom geom; dVector3 fdir1; };
typedef void dNearCallback(void* data, dGeomID o1, dGeomID o2);

struct dxGeom { int shape = -1; dxBody* body = nullptr; dxWorld* world = nullptr; int static_body = -1; };
struct dxBody {
    dxWorld* world = nullptr; int id = -1; bool dead = false; int enabled = 1;
    float pos[3] = { 0, 0, 0 }; float pending_force[3] = { 0, 0, 0 };
    dMass m; bool has_mass = false; dxGeom* geom = nullptr;
    dReal cpos[4] = { 0, 0, 0, 0 }, cquat[4] = { 1, 0, 0, 0 };
};
struct dxSpace { dxWorld* world = nullptr; };
struct dxWorld {
    np_world* w = nullptr; float gravity[3] = { 0, 0, 0 };
    std::vector<dxBody*> bodies; std::vector<dxGeom*> statics;          /* statics: geoms without a body (the ground plane) */
    std::map<std::vector<float>, int> shapes;                          /* one library shape per distinct geometry */
    std::vector<np_snapshot_object> snap; bool snap_valid = false;
};

namespace netphys_ode {
inline dxWorld*& current() { static dxWorld* w = nullptr; return w; }
inline int shape_for(dxWorld* W, const std::vector<float>& key, const std::vector<float>& xyz, float sphere_r)
{
    auto it = W->shapes.find(key);
    if (it != W->shapes.end()) return it->second;
    int s = sphere_r > 0 ? np_shape_create_sphere(W->w, sphere_r) : np_shape_create_mesh(W->w, xyz.data(), (int)xyz.size() / 3);
    W->shapes[key] = s;
    return s;
}
inline np_static_physics_data static_data(const float g[3], const float p[3], const dMass* m)
{
    np_static_physics_data d; memset(&d, 0, sizeof d);
    for (int k = 0; k < 3; k++) { d.initial_position[k] = p[k]; d.gravity[k] = m ? g[k] : 0.0f; }
    d.mass = m ? (float)m->mass : 0.0f;
    d.collision_response = m ? NP_COLLISION_RESPONSE_IMPULSE : NP_COLLISION_RESPONSE_NONE;
    if (m) {
        float I[9], inv[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) I[3 * r + c] = (float)m->I[4 * r + c];
        d.moment_of_inertia = I[0];
        float det = I[0] * (I[4] * I[8] - I[5] * I[7]) - I[1] * (I[3] * I[8] - I[5] * I[6]) + I[2] * (I[3] * I[7] - I[4] * I[6]);
        float id = det != 0.0f ? 1.0f / det : 0.0f;
        inv[0] = (I[4] * I[8] - I[5] * I[7]) * id; inv[1] = (I[2] * I[7] - I[1] * I[8]) * id; inv[2] = (I[1] * I[5] - I[2] * I[4]) * id;
        inv[3] = (I[5] * I[6] - I[3] * I[8]) * id; inv[4] = (I[0] * I[8] - I[2] * I[6]) * id; inv[5] = (I[2] * I[3] - I[0] * I[5]) * id;
        inv[6] = (I[3] * I[7] - I[4] * I[6]) * id; inv[7] = (I[1] * I[6] - I[0] * I[7]) * id; inv[8] = (I[0] * I[4] - I[1] * I[3]) * id;
        memcpy(d.inertia_tensor, I, sizeof I); memcpy(d.inverse_inertia_tensor, inv, sizeof inv);
    }
    return d;
}
/* create, in dBodyCreate order, every body that has its geom by now (and the static geoms) */
inline void realise(dxWorld* W)
{
    for (dxGeom* g : W->statics)
        if (g->static_body < 0) { float p[3] = { 0, 0, 0 }; np_static_physics_data d = static_data(W->gravity, p, nullptr); g->static_body = np_body_create(W->w, g->shape, &d); W->snap_valid = false; }
    for (dxBody* b : W->bodies) {
        if (b->id >= 0 || b->dead || !b->geom) continue;
        if (!b->has_mass) { memset(&b->m, 0, sizeof b->m); b->m.mass = 1; b->m.I[0] = b->m.I[5] = b->m.I[10] = 1; b->has_mass = true; }
        np_static_physics_data d = static_data(W->gravity, b->pos, &b->m);
        b->id = np_body_create(W->w, b->geom->shape, &d);
        if (b->id < 0) continue;
        if (!b->enabled) np_body_set_enabled(W->w, b->id, 0);
        if (b->pending_force[0] != 0 || b->pending_force[1] != 0 || b->pending_force[2] != 0) { np_body_add_force(W->w, b->id, b->pending_force); b->pending_force[0] = b->pending_force[1] = b->pending_force[2] = 0; }
        W->snap_valid = false;
    }
}
inline void refresh(dxWorld* W)
{
    if (W->snap_valid) return;
    realise(W);
    int n = np_world_body_count(W->w);
    W->snap.resize((size_t)(n > 0 ? n : 1));
    if (n > 0) np_world_snapshot(W->w, W->snap.data(), n);
    for (dxBody* b : W->bodies) {
        if (b->id < 0 || b->id >= n) continue;
        const np_snapshot_object& o = W->snap[(size_t)b->id];
        for (int k = 0; k < 3; k++) b->cpos[k] = o.pos[k];
        for (int k = 0; k < 4; k++) b->cquat[k] = o.rot[k];
    }
    W->snap_valid = true;
}
}  // namespace netphys_ode

/* ---- library / world / space ---------------------------------------------------------------------------------------------- */
inline int dInitODE2(unsigned) { return 1; }
inline void dCloseODE() {}
inline int dAllocateODEDataForThread(unsigned) { return 1; }
inline dWorldID dWorldCreate() { dxWorld* W = new dxWorld; W->w = np_world_create(-1); netphys_ode::current() = W; return W; }
inline void dWorldDestroy(dWorldID W) { if (!W) return; for (dxBody* b : W->bodies) delete b; for (dxGeom* g : W->statics) delete g; np_world_destroy(W->w); if (netphys_ode::current() == W) netphys_ode::current() = nullptr; delete W; }
inline dSpaceID dHashSpaceCreate(dSpaceID) { dxSpace* s = new dxSpace; s->world = netphys_ode::current(); return s; }
inline void dSpaceDestroy(dSpaceID s) { delete s; }
inline void dWorldSetGravity(dWorldID W, dReal x, dReal y, dReal z) { W->gravity[0] = (float)x; W->gravity[1] = (float)y; W->gravity[2] = (float)z; }
inline void dWorldSetCFM(dWorldID, dReal) {}
inline void dWorldSetAutoDisableFlag(dWorldID, int) {}
inline void dWorldSetAutoDisableAverageSamplesCount(dWorldID, unsigned) {}
inline void dWorldSetLinearDamping(dWorldID, dReal) {}
inline void dWorldSetAngularDamping(dWorldID, dReal) {}
inline void dWorldSetMaxAngularSpeed(dWorldID, dReal) {}
inline void dWorldSetContactMaxCorrectingVel(dWorldID, dReal) {}
inline void dWorldSetContactSurfaceLayer(dWorldID, dReal) {}
inline void dSpaceCollide(dSpaceID, void*, dNearCallback*) {}           /* GetCollisions runs inside the step (physics.cpp:157) */
inline int dWorldQuickStep(dWorldID W, dReal dt) { netphys_ode::realise(W); W->snap_valid = false; return np_world_step(W->w, (float)dt) == NP_OK; }   /* Physics_Update */
inline int dWorldStep(dWorldID W, dReal dt) { return dWorldQuickStep(W, dt); }
/* threading: the step runs on the GPU */
inline dThreadingImplementationID dThreadingAllocateMultiThreadedImplementation() { return nullptr; }
inline dThreadingThreadPoolID dThreadingAllocateThreadPool(unsigned, size_t, unsigned, void*) { return nullptr; }
inline void dThreadingThreadPoolServeMultiThreadedImplementation(dThreadingThreadPoolID, dThreadingImplementationID) {}
inline void dWorldSetStepIslandsProcessingMaxThreadCount(dWorldID, unsigned) {}
inline const dxThreadingFunctionsInfo* dThreadingImplementationGetFunctions(dThreadingImplementationID) { return nullptr; }
inline void dWorldSetStepThreadingImplementation(dWorldID, const dxThreadingFunctionsInfo*, dThreadingImplementationID) {}
inline void dThreadingImplementationShutdownProcessing(dThreadingImplementationID) {}
inline void dThreadingFreeThreadPool(dThreadingThreadPoolID) {}
inline void dThreadingFreeImplementation(dThreadingImplementationID) {}
/* joints: contacts are resolved by the engine's response, not by joints */
inline dJointGroupID dJointGroupCreate(int) { return nullptr; }
inline void dJointGroupDestroy(dJointGroupID) {}
inline void dJointGroupEmpty(dJointGroupID) {}
inline dJointID dJointCreateContact(dWorldID, dJointGroupID, const dContact*) { return nullptr; }
inline void dJointAttach(dJointID, dBodyID, dBodyID) {}
inline int dAreConnectedExcluding(dBodyID, dBodyID, int) { return 0; }
inline int dCollide(dGeomID, dGeomID, int, dContactGeom*, int) { return 0; }
inline void dPlaneSpace(const dVector3 n, dVector3 p, dVector3 q)
{
    if (std::fabs(n[2]) > 0.7071067811865476) { dReal a = n[1] * n[1] + n[2] * n[2], k = 1 / std::sqrt(a); p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k; q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1]; }
    else { dReal a = n[0] * n[0] + n[1] * n[1], k = 1 / std::sqrt(a); p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0; q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k; }
}
inline dReal dCalcVectorDot3(const dReal* a, const dReal* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* ---- mass ----------------------------------------------------------------------------------------------------------------- */
inline void dMassSetZero(dMass* m) { memset(m, 0, sizeof *m); }
inline void dMassSetBox(dMass* m, dReal density, dReal lx, dReal ly, dReal lz)
{
    dMassSetZero(m); m->mass = density * lx * ly * lz;
    m->I[0] = m->mass / 12 * (ly * ly + lz * lz); m->I[5] = m->mass / 12 * (lx * lx + lz * lz); m->I[10] = m->mass / 12 * (lx * lx + ly * ly);
}
inline void dMassSetSphere(dMass* m, dReal density, dReal r)
{
    dMassSetZero(m); m->mass = (4.0 / 3.0) * 3.14159265358979323846 * r * r * r * density;
    m->I[0] = m->I[5] = m->I[10] = 0.4 * m->mass * r * r;
}

/* ---- geoms ---------------------------------------------------------------------------------------------------------------- */
inline dGeomID dCreateBox(dSpaceID s, dReal lx, dReal ly, dReal lz)
{
    dxWorld* W = s ? s->world : netphys_ode::current();
    const float hx = (float)lx * 0.5f, hy = (float)ly * 0.5f, hz = (float)lz * 0.5f;
    std::vector<float> xyz;
    for (int k = 0; k < 8; k++) { xyz.push_back((k & 1) ? hx : -hx); xyz.push_back((k & 2) ? hy : -hy); xyz.push_back((k & 4) ? hz : -hz); }
    dxGeom* g = new dxGeom; g->world = W; g->shape = netphys_ode::shape_for(W, { 1.0f, hx, hy, hz }, xyz, 0.0f);
    return g;
}
inline dGeomID dCreateSphere(dSpaceID s, dReal r)
{
    dxWorld* W = s ? s->world : netphys_ode::current();
    dxGeom* g = new dxGeom; g->world = W; g->shape = netphys_ode::shape_for(W, { 2.0f, (float)r }, {}, (float)r);
    return g;
}
inline dGeomID dCreatePlane(dSpaceID s, dReal a, dReal b, dReal c, dReal d)
{
    /* a static slab of thickness 1 whose top face lies in the plane a x + b y + c z = d (axis-aligned normals only) */
    dxWorld* W = s ? s->world : netphys_ode::current();
    const int axis = std::fabs(a) > 0.5 ? 0 : (std::fabs(b) > 0.5 ? 1 : 2);
    const float sign = ((axis == 0 ? a : axis == 1 ? b : c) < 0) ? -1.0f : 1.0f, top = (float)d * sign, ext = 100.0f;
    std::vector<float> xyz;
    for (int k = 0; k < 8; k++) {
        float p[3]; int bit = 0;
        for (int ax = 0; ax < 3; ax++) p[ax] = (ax == axis) ? (((k >> 2) & 1) ? top : top - sign) : (((k >> bit++) & 1) ? ext : -ext);
        xyz.insert(xyz.end(), p, p + 3);
    }
    dxGeom* g = new dxGeom; g->world = W; g->shape = netphys_ode::shape_for(W, { 3.0f, (float)axis, sign, top }, xyz, 0.0f);
    W->statics.push_back(g);
    return g;
}
inline void dGeomSetBody(dGeomID g, dBodyID b) { g->body = b; if (b) b->geom = g; }
inline dBodyID dGeomGetBody(dGeomID g) { return g ? g->body : nullptr; }
inline void dGeomDestroy(dGeomID g)
{
    if (!g) return;
    if (g->body && g->body->geom == g) g->body->geom = nullptr;
    if (g->world) for (size_t k = 0; k < g->world->statics.size(); k++) if (g->world->statics[k] == g) return;   /* static geoms live as long as the world */
    delete g;
}

/* ---- bodies --------------------------------------------------------------------------------------------------------------- */
inline dBodyID dBodyCreate(dWorldID W) { dxBody* b = new dxBody; b->world = W; W->bodies.push_back(b); return b; }
inline void dBodySetMass(dBodyID b, const dMass* m) { b->m = *m; b->has_mass = true; }
inline void dBodyGetMass(dBodyID b, dMass* m) { *m = b->m; }
inline void dBodySetAutoDisableFlag(dBodyID, int) {}
inline void dBodySetPosition(dBodyID b, dReal x, dReal y, dReal z)
{
    b->pos[0] = (float)x; b->pos[1] = (float)y; b->pos[2] = (float)z;
    if (b->id >= 0) { np_body_state st; if (np_body_get_state(b->world->w, b->id, &st) == NP_OK) { memcpy(st.position, b->pos, sizeof b->pos); np_body_set_state(b->world->w, b->id, &st); } b->world->snap_valid = false; }
    else { b->cpos[0] = x; b->cpos[1] = y; b->cpos[2] = z; }
}
inline const dReal* dBodyGetPosition(dBodyID b) { if (b->id >= 0 || b->geom) netphys_ode::refresh(b->world); return b->cpos; }
inline const dReal* dBodyGetQuaternion(dBodyID b) { if (b->id >= 0 || b->geom) netphys_ode::refresh(b->world); return b->cquat; }
inline int dBodyIsEnabled(dBodyID b) { return b->enabled; }
inline void dBodyEnable(dBodyID b) { b->enabled = 1; if (b->id >= 0) { np_body_set_enabled(b->world->w, b->id, 1); b->world->snap_valid = false; } }
inline void dBodyDisable(dBodyID b) { b->enabled = 0; if (b->id >= 0) { np_body_set_enabled(b->world->w, b->id, 0); b->world->snap_valid = false; } }
inline void dBodyAddForce(dBodyID b, dReal fx, dReal fy, dReal fz)
{
    const float f[3] = { (float)fx, (float)fy, (float)fz };
    if (b->id < 0 && b->geom) netphys_ode::realise(b->world);
    if (b->id >= 0) np_body_add_force(b->world->w, b->id, f);
    else for (int k = 0; k < 3; k++) b->pending_force[k] += f[k];
}
inline void dBodyDestroy(dBodyID b)
{
    if (!b || b->dead) return;
    dxWorld* W = b->world;
    b->dead = true; if (b->geom && b->geom->body == b) b->geom->body = nullptr;
    if (b->id >= 0) np_body_set_enabled(W->w, b->id, 0);
    bool all_dead = true;
    for (dxBody* o : W->bodies) all_dead = all_dead && o->dead;
    if (all_dead) {                                                    /* World::Reset: start from an empty store */
        np_world_clear_bodies(W->w);
        for (dxBody* o : W->bodies) delete o;
        W->bodies.clear();
        for (dxGeom* g : W->statics) g->static_body = -1;
    }
    W->snap_valid = false;
}

#endif /* NETPHYS_ODE_SHIM_H */
