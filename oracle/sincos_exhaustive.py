"""oracle/sincos_exhaustive.py -- TEST INFRASTRUCTURE.

Checks the oracle's restated sinf/cosf against this box's libm over ALL 2^32 float bit patterns
(threads; ~1-2 min on 8 cores) and (re)writes tests/golden/sincos_hash.npz: per-2^24-block checksums of
the oracle's sin/cos that the GPU exhaustive test compares its own checksums with.

    python oracle/sincos_exhaustive.py [--hash-only]
"""
import ctypes as C
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bind import Oracle  # noqa: E402

BLOCK = 1 << 24


def main():
    L = Oracle().L
    L.npo_sincos_vs_libm.restype = C.c_longlong
    L.npo_sincos_vs_libm.argtypes = [C.c_uint, C.c_ulonglong, C.POINTER(C.c_uint)]
    L.npo_sincos_hash.argtypes = [C.c_uint, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    hash_only = "--hash-only" in sys.argv

    def work(b):
        bad = 0; first = C.c_uint(0)
        if not hash_only:
            bad = L.npo_sincos_vs_libm(b * BLOCK, BLOCK, C.byref(first))
        hs = C.c_ulonglong(); hc = C.c_ulonglong()
        L.npo_sincos_hash(b * BLOCK, BLOCK, C.byref(hs), C.byref(hc))
        return bad, first.value, hs.value, hc.value

    with ThreadPoolExecutor(os.cpu_count()) as ex:
        res = list(ex.map(work, range(256)))
    bad = sum(r[0] for r in res)
    print("inputs checked: 2^32   mismatches vs libm:", "skipped" if hash_only else bad)
    for b, r in enumerate(res):
        if r[0]:
            print("  block %d: %d mismatches, first at bits 0x%08x" % (b, r[0], r[1]))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sincos_hash.npz")
    np.savez(out, block=np.uint64(BLOCK), sin=np.array([r[2] for r in res], np.uint64), cos=np.array([r[3] for r in res], np.uint64),
             mismatches_vs_libm=np.int64(-1 if hash_only else bad))
    print("wrote", out)
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
