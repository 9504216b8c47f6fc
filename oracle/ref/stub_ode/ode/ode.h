/* oracle/ref/stub_ode/ode/ode.h -- TEST INFRASTRUCTURE.  A table-driven stand-in for the three ODE getters that
 * netphys_server/world_s.cpp:37-48 calls, so that the reference's own World_S_Update / World_S_Fill*Message / Packet::Serialize can
 * be compiled UNCHANGED on the CPU and fed known body records (oracle/ref/ref_packet.cpp).  Not ODE, no physics. */
#pragma once
typedef double dReal;
struct dxBody { dReal pos[4]; dReal quat[4]; int enabled; };
typedef dxBody* dBodyID;
struct dxGeom; typedef dxGeom* dGeomID;
inline const dReal* dBodyGetPosition(dBodyID b) { return b->pos; }
inline const dReal* dBodyGetQuaternion(dBodyID b) { return b->quat; }
inline int dBodyIsEnabled(dBodyID b) { return b->enabled; }
