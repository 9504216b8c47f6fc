// The server's world code paths, written against <ode/ode.h> exactly as netphys_common/world.cpp (Init / Create / Reset / Update),
// worldobject.cpp:7-20 (WorldObject::Init), player.cpp:9-38 + netphys_server/player_s.cpp:33-45 (player body, input forces) and
// netphys_server/world_s.cpp:23-51 (per-object state readout) use it -- compiled with include/netphys/ode_shim on the include path, so
// every d* call below lands on the netphys-b200 C ABI.  Prints one line per body per frame: frame body pos[3] quat[4] enabled (hex).
#include <ode/ode.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

static const int NUM_ROWS_COLS = 5;                 // common.h:58 uses 10
static const float GRAVITY = 9.8f, BOX_SIZE = 0.5f, DENSITY = 5.0f, PLAYER_ACCEL = 20.0f;   // common.h:61-63, player.h:6
static const dReal PLAYER_SIZE = 2.f;               // player.h:7

struct Obj { dBodyID body; dGeomID geom; };
static dWorldID s_world; static dSpaceID s_space;

static Obj world_object(float x, float y, float z)  // WorldObject::Init
{
    Obj o; o.body = dBodyCreate(s_world); o.geom = dCreateBox(s_space, BOX_SIZE, BOX_SIZE, BOX_SIZE);
    dBodySetPosition(o.body, x, y, z);
    dMass mass; dMassSetBox(&mass, DENSITY, BOX_SIZE, BOX_SIZE, BOX_SIZE);
    dGeomSetBody(o.geom, o.body); dBodySetMass(o.body, &mass);
    return o;
}
static void reset(std::vector<Obj>& objs)           // World::Reset
{
    for (Obj& o : objs) { dBodyDestroy(o.body); dGeomDestroy(o.geom); }
    objs.clear();
    for (int i = 0; i < NUM_ROWS_COLS; i++) for (int j = 0; j < NUM_ROWS_COLS; j++)
        objs.push_back(world_object(float(i - NUM_ROWS_COLS / 2), float(j - NUM_ROWS_COLS / 2), BOX_SIZE));
}
static void print_frame(int frame, const std::vector<Obj>& objs)   // World_S_Update
{
    for (size_t k = 0; k < objs.size(); k++) {
        const dReal* pos = dBodyGetPosition(objs[k].body); const dReal* rot = dBodyGetQuaternion(objs[k].body);
        float v[7] = { (float)pos[0], (float)pos[1], (float)pos[2], (float)rot[0], (float)rot[1], (float)rot[2], (float)rot[3] };
        printf("%d %zu", frame, k);
        for (int c = 0; c < 7; c++) { uint32_t u; memcpy(&u, &v[c], 4); printf(" %08x", u); }
        printf(" %d\n", dBodyIsEnabled(objs[k].body));
    }
}
int main(int argc, char** argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 30;
    const dReal dt = 0.01;                          // server tick
    dInitODE2(0);
    s_world = dWorldCreate(); s_space = dHashSpaceCreate(0);
    dWorldSetGravity(s_world, 0, 0, -GRAVITY); dWorldSetCFM(s_world, 1e-5); dWorldSetAutoDisableFlag(s_world, 1);
    dCreatePlane(s_space, 0, 0, 1, 0);              // World::Create
    std::vector<Obj> objs; reset(objs);             // World::Start
    Obj player = { nullptr, nullptr };
    for (int f = 0; f < frames; f++) {
        if (f == 4) {                               // Player_S: INPUT_SPACE with no body yet
            player.body = dBodyCreate(s_world); dBodySetAutoDisableFlag(player.body, 0); dBodySetPosition(player.body, 0, 0, 5.f);
            dMass mass; dMassSetSphere(&mass, DENSITY, PLAYER_SIZE); dBodySetMass(player.body, &mass);
            player.geom = dCreateSphere(s_space, PLAYER_SIZE); dGeomSetBody(player.geom, player.body);
            objs.push_back(player);
        }
        if (player.body && f % 3 == 0) {            // Player::HandleInputsInternal: forward + left + jump
            float a[3] = { PLAYER_ACCEL, PLAYER_ACCEL, f % 2 ? PLAYER_ACCEL : 0.0f };
            dMass m; dBodyGetMass(player.body, &m);
            dReal force[3] = { a[0] * m.mass, a[1] * m.mass, a[2] * m.mass };
            dBodyAddForce(player.body, force[0], force[1], force[2]);
        }
        if (f == 10) dBodyDisable(objs[3].body);
        if (f == 16) dBodyEnable(objs[3].body);
        if (f == 20) { reset(objs); player.body = nullptr; }                      // INPUT_RESET_WORLD
        dSpaceCollide(s_space, 0, nullptr); dWorldQuickStep(s_world, dt);         // World::Update
        print_frame(f, objs);
    }
    dSpaceDestroy(s_space); dWorldDestroy(s_world); dCloseODE();
    return 0;
}
