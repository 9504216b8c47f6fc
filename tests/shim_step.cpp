// tests/shim_step.cpp -- the reference's single-step GJK/EPA client code (tests/cpp/step_trace_body.h, written against engine/physics.h
// names only) compiled against the drop-in shim and run on the device.  usage: shim_step.bin <pairs.txt> <trace.txt> <max_calls>
#define TEST_PROGRAM
#include <netphys/physics.h>
#include <netphys/physics_shape.h>

#include <cstdlib>

#include "cpp/step_trace_body.h"

int main(int argc, char** argv)
{
    if (argc < 4) return 2;
    if (!netphys::default_world()) { std::fprintf(stderr, "no device: %s\n", np_last_error()); return 3; }
    int pairs = step_trace::run(argv[1], argv[2], std::atoi(argv[3]));
    if (netphys::last_status() != NP_OK) { std::fprintf(stderr, "library error %d: %s\n", netphys::last_status(), np_last_error()); return 4; }
    std::printf("%d\n", pairs);
    netphys::destroy_default_world();
    return pairs > 0 ? 0 : 1;
}
