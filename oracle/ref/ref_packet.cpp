// oracle/ref/ref_packet.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's snapshot producer and wire serialiser, compiled UNCHANGED from where they lie: netphys_server/world_s.cpp
// (World_S_Update's command-frame ring :23-69, World_S_FillNewConnectionMessage / World_S_FillWorldUpdateMessage :73-108) with
// netphys_common/common.h (Packet::Serialize :77-93, CommandFrameObject :120-127, the packet classes :130-163) and lib.h's DataStore.
// The ODE getters world_s.cpp calls are the table-driven stub in stub_ode/ode/ode.h; ObjectManager_S and the logger are the few
// lines below.  Output: the exact bytes the reference server would put on the wire for given body records, which
// tests compare with np_world_record_frame + np_world_fill_packet.  Built by oracle/Makefile into oracle/_ref/ (never shipped, never
// loaded by the product); -ftrivial-auto-var-init=zero pins the 3 padding bytes of CommandFrameObject (bool isEnabled), which the
// reference leaves uninitialised.
#include <cstdarg>
#include <cstdio>
#include <vector>
#include <ode/ode.h>
// MSVC drops the comma in front of an empty __VA_ARGS__, g++ does not (world_s.cpp:77,91 call LOG_ERROR with the format only): take
// log.h first (#pragma once), then give LOG_ERROR a portable spelling.  No reference source is edited or stored.
#include "../netphys_common/log.h"
#undef LOG_ERROR
#define LOG_ERROR(...) Log_Write(__FILE__, __LINE__, 0, true, true, __VA_ARGS__)
#include "world_s.cpp"          // -I$(REF)/../netphys_server

// netphys_common/log.h
void Log_Write(const char*, int, int, bool, bool, const char*, ...) {}

namespace {
struct StubObject : public Object {
    StubObject(const NPGUID& g, dBodyID b) : Object(g), body(b) {}
    dBodyID GetBodyID() const override { return body; }
    dBodyID body;
};
std::vector<StubObject*> g_objects;
std::vector<dxBody> g_bodies;
}  // namespace
// netphys_server/objectmanager_s.h
Object* ObjectManager_S_GetFirst() { return g_objects.empty() ? nullptr : g_objects.front(); }
Object* ObjectManager_S_GetNext(Object* obj) { return obj->m_next; }

extern "C" {
int refpkt_object_size() { return (int)sizeof(CommandFrameObject); }
// the objects the server would iterate: unique ids (NPGUID(uid, ObjectType_WorldObject)), positions, quaternions (w,x,y,z), enabled flags
void refpkt_set_objects(int n, const unsigned* uid, const float* pos3, const float* quat4, const int* enabled)
{
    for (auto* o : g_objects) delete o;
    g_objects.clear(); g_bodies.assign(n, dxBody());
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) g_bodies[i].pos[k] = pos3[3 * i + k];
        for (int k = 0; k < 4; k++) g_bodies[i].quat[k] = quat4[4 * i + k];
        g_bodies[i].enabled = enabled[i];
        g_objects.push_back(new StubObject(NPGUID(uid[i], ObjectType_WorldObject), &g_bodies[i]));
    }
    for (int i = 0; i < n; i++) { g_objects[i]->m_prev = i ? g_objects[i - 1] : nullptr; g_objects[i]->m_next = i + 1 < n ? g_objects[i + 1] : nullptr; }
}
void refpkt_reset_frames() { s_commandFrames.clear(); s_frameCounter = 1; }
void refpkt_update(double now) { World_S_Update(now); }                       // world_s.cpp:23
int refpkt_frame_count() { return (int)s_commandFrames.size(); }
// kind 0: ClientWorldStateUpdatePacket, 1: ClientNewConnection; returns Packet::Serialize's result
int refpkt_fill(int kind, char* buffer, unsigned cap)
{
    if (kind == 0) { ClientWorldStateUpdatePacket p; World_S_FillWorldUpdateMessage(&p, 0); if (!s_commandFrames.size()) { p.data.Free(); return 0; } int r = p.Serialize(buffer, cap); if (!r) p.data.Free(); return r; }
    ClientNewConnection p; World_S_FillNewConnectionMessage(&p); if (!s_commandFrames.size()) { p.data.Free(); return 0; } int r = p.Serialize(buffer, cap); if (!r) p.data.Free(); return r;
}
}
