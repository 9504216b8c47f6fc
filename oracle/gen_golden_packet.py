"""oracle/gen_golden_packet.py -- TEST INFRASTRUCTURE.  tests/golden/packet_few_boxes.npz: the bytes the REFERENCE's own server code
(netphys_server/world_s.cpp World_S_Update / World_S_Fill*Message + netphys_common/common.h Packet::Serialize, compiled unchanged into
oracle/_ref/librefpacket.so) puts on the wire for the few-boxes scene, frame by frame, including the 5-second frame ring.

Body records fed to it: position = the compiled reference engine's state after each step; rotation = the quaternion (w,x,y,z) of the
engine's Euler XYZ angles, R = Rz*Ry*Rx, restated here in float32 with the reference libm's sinf/cosf (the engine has no quaternions:
this conversion is the ODE facade's definition); guid = NPGUID(uid, ObjectType_WorldObject).

    python oracle/gen_golden_packet.py          (dev container only: needs oracle/_ref)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from netphys_b200 import scenes  # noqa: E402
from oracle.bind import Ref, RefPacket  # noqa: E402

F = np.float32
TIMES_MS = [0.0, 1000.0, 2500.0, 4000.0, 5200.0, 6100.0, 9000.0, 9050.0, 9100.0]     # frame 8 is taken after body 2 was disabled, without a step
OBJECT_TYPE_WORLDOBJECT = 2                                                          # netphys_common/common.h:13-18


def euler_to_quat(ref, rot):
    """float32, operation order of np::snap_of (netphys_b200/csrc/np_kernels.cuh): half angles, sincos, (w, x, y, z)."""
    h = (rot.astype(F) * F(0.5)).astype(F)
    s, c = ref.sincos(h.ravel()); s = s.reshape(h.shape); c = c.reshape(h.shape)
    sx, sy, sz = s[:, 0], s[:, 1], s[:, 2]; cx, cy, cz = c[:, 0], c[:, 1], c[:, 2]
    m = lambda a, b: (a.astype(F) * b.astype(F)).astype(F)
    w = (m(m(cx, cy), cz) + m(m(sx, sy), sz)).astype(F)
    x = (m(m(sx, cy), cz) - m(m(cx, sy), sz)).astype(F)
    y = (m(m(cx, sy), cz) + m(m(sx, cy), sz)).astype(F)
    z = (m(m(cx, cy), sz) - m(m(sx, sy), cz)).astype(F)
    return np.stack([w, x, y, z], axis=1)


def main():
    r = Ref(); pk = RefPacket(); pk.reset_frames()

    class Fac:
        def make_sphere(self, rad): return r.shape_verts(r.sphere(rad))
        def make_box(self, w, d, h): return r.shape_verts(r.box(w, d, h))
    sc = scenes.few_boxes_scene()
    r.load_scene(scenes.resolve(sc, Fac()))
    n = len(sc["sd"])
    uid = np.arange(1, n + 1, dtype=np.uint32)
    out = {"times_ms": np.array(TIMES_MS), "guid_base": np.uint32((OBJECT_TYPE_WORLDOBJECT << 24) + 1), "n": np.int32(n)}
    enabled = np.ones(n, np.int32)
    for k, t in enumerate(TIMES_MS):
        if k < 8: r.step(sc["dt"])
        else: enabled[2] = 0
        st = r.state()
        pk.set_objects(uid, st[:, 0:3], euler_to_quat(r, st[:, 3:6]), enabled)
        pk.update(t)
        out["state_%d" % k] = st
        out["frames_kept_%d" % k] = np.int32(pk.frame_count())
        out["update_%d" % k] = np.frombuffer(pk.fill(0), np.uint8)
        out["newconn_%d" % k] = np.frombuffer(pk.fill(1), np.uint8)
        out["too_small_%d" % k] = np.int32(len(pk.fill(0, cap=24 + 36 * n - 1)))       # Packet::Serialize returns 0 when the buffer is too small
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "packet_few_boxes.npz"), **out)
    print("frames kept:", [int(out["frames_kept_%d" % k]) for k in range(len(TIMES_MS))], "packet bytes:", len(out["update_0"]))


if __name__ == "__main__":
    main()
