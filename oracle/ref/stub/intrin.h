/* MSVC <intrin.h> stand-in for compiling the reference's netphys_common sources with g++ (test infrastructure only). */
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cstdlib>
static inline unsigned char _BitScanReverse(unsigned long* index, unsigned long mask) { if (!mask) return 0; *index = 63u - (unsigned long)__builtin_clzl(mask); return 1; }
static inline unsigned char _BitScanForward(unsigned long* index, unsigned long mask) { if (!mask) return 0; *index = (unsigned long)__builtin_ctzl(mask); return 1; }
