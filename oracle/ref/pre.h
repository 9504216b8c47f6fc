/* oracle/ref/pre.h -- TEST INFRASTRUCTURE (never shipped, never on the product path).
 * Force-included ahead of the untouched reference sources so they build with g++:
 *  - pulls in every std header the engine uses BEFORE engine/lib.h defines its
 *    min/max/clamp macros (reference engine/lib.h:13-23), which otherwise break libstdc++;
 *  - <math.h> (the C++ wrapper) brings the float overloads of sqrt/sin/cos into the
 *    global namespace, so the unqualified calls in engine/vector.h (magnitude),
 *    engine/matrix.h:283-288 (rotate) and engine/collision_detection.cpp:288,322
 *    resolve to sqrtf/sinf/cosf exactly as they do under MSVC;
 *  - MSVC accepts `vector3 v = { 0.0f }` through an explicit ctor
 *    (engine/physics.cpp:170-171); g++ does not, so `explicit` is dropped AFTER
 *    the std headers have been parsed.
 */
#pragma once
#include <math.h>
#include <cmath>
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cassert>
#include <algorithm>
#include <vector>
#include <list>
#include <limits>
#include <chrono>
#include <thread>
#define explicit
