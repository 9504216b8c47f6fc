// tests/cpp/step_trace_body.h -- TEST INFRASTRUCTURE.  Client code written ONLY against the reference's names (engine/physics.h:103-127,
// engine/simplex.h, engine/physics_shape.h): it single-steps GJK and EPA the way engine/test_3d.cpp:36-85 and test_2d.cpp:34-75 do on
// their 'T' key and writes every intermediate simplex as hex words.  The same text is compiled twice:
//   * against the reference's own engine sources (oracle/ref/ref_driver.cpp -> ref_step_trace; fixture tests/golden/step_trace_ref.txt.gz)
//   * against include/netphys/physics.h (tests/shim_step.cpp), running on the B200,
// and the two traces must be identical bit for bit.
#pragma once
#include <cstdio>
#include <cstring>
#include <vector>

namespace step_trace {
inline void put(FILE* f, float v) { unsigned u; std::memcpy(&u, &v, 4); if ((u & 0x7fffffffu) > 0x7f800000u) u = 0x7fc00000u; std::fprintf(f, " %08x", u); }
inline void put(FILE* f, const vector3& v) { put(f, v.x); put(f, v.y); put(f, v.z); }

inline void dump(FILE* f, int pair, int call, const char* what, int result, const Simplex& s, const vector3& dir, bool found, const CollisionData& cd)
{
    std::fprintf(f, "%d %d %s r=%d nv=%d nf=%d o=%d found=%d |", pair, call, what, result, s.size(), (int)s.faces.size(), s.m_containsOrigin ? 1 : 0, found ? 1 : 0);
    for (int i = 0; i < s.size(); i++) { put(f, s.verts[i].p); put(f, s.verts[i].A); put(f, s.verts[i].B); }
    std::fprintf(f, " |");
    for (size_t i = 0; i < s.faces.size(); i++) { std::fprintf(f, " %d %d %d", s.faces[i].point_index[0], s.faces[i].point_index[1], s.faces[i].point_index[2]); put(f, s.faces[i].normal); }
    std::fprintf(f, " |"); put(f, dir);
    std::fprintf(f, " |"); put(f, cd.penetrationDirection); put(f, cd.depth);
    std::fprintf(f, "\n");
}

// one press of 'R' followed by up to max_calls presses of 'T'
inline void trace_pair(FILE* f, int pair, const CollisionParams& params, bool is3D, int max_calls)
{
    Simplex simplex;
    CollisionData cd; cd.depth = 0.0f; cd.penetrationDirection = vector3();
    vector3 dir;
    simplex.verts.resize(1);
    if (is3D) {                                               // test_3d.cpp:36-51: support in shape space, then moved to the world
        vector3 a_local = params.a->GetPointFurthestInDirection({ 1, 0, 0 }, matrix4(), true);
        vector3 b_local = params.b->GetPointFurthestInDirection({ -1, 0, 0 }, matrix4(), true);
        simplex.verts[0].A = params.aTransform * a_local;
        simplex.verts[0].B = params.bTransform * b_local;
    } else {                                                  // test_2d.cpp:34-49: support in the world
        simplex.verts[0].A = params.a->GetPointFurthestInDirection({ 1, 0, 0 }, params.aTransform, false);
        simplex.verts[0].B = params.b->GetPointFurthestInDirection({ -1, 0, 0 }, params.bTransform, false);
    }
    simplex.verts[0].p = simplex.verts[0].A - simplex.verts[0].B;
    simplex.faces.clear();
    simplex.m_containsOrigin = false;
    dir = GetSearchDirection(simplex);
    COLLISION_RESULT result = COLLISION_RESULT_NONE;
    bool found = false;
    dump(f, pair, -1, "reset", (int)result, simplex, dir, found, cd);
    for (int call = 0; call < max_calls && !found; call++) {
        const char* what;
        if (result == COLLISION_RESULT_NONE || result == COLLISION_RESULT_CONTINUE) {
            what = "gjk";
            result = DetectCollisionStep(params, is3D, simplex);
            if (simplex.size() < 4) dir = GetSearchDirection(simplex);
            if (is3D && simplex.m_containsOrigin) simplex.SetupForEPA();
        } else {
            what = "epa";
            found = FindIntersectionPoints(params, simplex, 1, is3D, &cd);
            if (!found && !simplex.m_containsOrigin) { dump(f, pair, call, what, (int)result, simplex, dir, found, cd); break; }   // a miss with nothing to report
        }
        dump(f, pair, call, what, (int)result, simplex, dir, found, cd);
    }
}

// input: one pair per record -- "is3D  kindA nA  kindB nB" then nA / nB parameter floats (kind 0 box: w d h; 1 sphere: r; 2 vertex list:
// 3 per vertex), then the two row-major 4x4 transforms, all floats as 8-digit hex words
inline bool read_float(FILE* f, float* v) { unsigned u; if (std::fscanf(f, "%x", &u) != 1) return false; std::memcpy(v, &u, 4); return true; }
inline MeshPhysicsShape* read_shape(FILE* f, int kind, int n)
{
    std::vector<float> p((size_t)n);
    for (int i = 0; i < n; i++) if (!read_float(f, &p[(size_t)i])) return nullptr;
    MeshPhysicsShape* s = new MeshPhysicsShape();
    if (kind == 0) s->CreateBox(p[0], p[1], p[2]);
    else if (kind == 1) s->CreateSphere(p[0]);
    else for (int i = 0; i + 2 < n; i += 3) s->m_mesh.m_vertexPos.push_back(vector3(p[(size_t)i], p[(size_t)i + 1], p[(size_t)i + 2]));
    return s;
}
inline int run(const char* in_path, const char* out_path, int max_calls)
{
    FILE* in = std::fopen(in_path, "r"); if (!in) return -1;
    FILE* out = std::fopen(out_path, "w"); if (!out) { std::fclose(in); return -1; }
    int pairs = 0, is3D, ka, na, kb, nb;
    while (std::fscanf(in, "%d %d %d %d %d", &is3D, &ka, &na, &kb, &nb) == 5) {
        MeshPhysicsShape* a = read_shape(in, ka, na); MeshPhysicsShape* b = read_shape(in, kb, nb);
        float m[32]; bool ok = a && b;
        for (int i = 0; i < 32 && ok; i++) ok = read_float(in, &m[i]);
        if (!ok) break;
        CollisionParams p; p.a = a; p.b = b;
        std::memcpy(&p.aTransform.x1, m, 64); std::memcpy(&p.bTransform.x1, m + 16, 64);
        trace_pair(out, pairs, p, is3D != 0, max_calls);
        pairs++;
    }
    std::fclose(in); std::fclose(out);
    return pairs;
}
}  // namespace step_trace
