/* oracle/ref/pre_server.h -- TEST INFRASTRUCTURE: force-included when a netphys_server source of the reference is compiled with g++.
 * MSVC drops the comma in front of an empty __VA_ARGS__, g++ does not: take the reference's log.h first (#pragma once), then give
 * the LOG macros a portable spelling.  Needs -I<reference>/netphys_server. */
#pragma once
#include <cstdint>
#include "../netphys_common/log.h"
#undef LOG_ERROR
#undef LOG_WARNING
#undef LOG_CONSOLE
#undef LOG
#define LOG(...) Log_Write(__FILE__, __LINE__, 0, false, false, __VA_ARGS__)
#define LOG_CONSOLE(...) Log_Write(__FILE__, __LINE__, 0, true, false, __VA_ARGS__)
#define LOG_WARNING(...) Log_Write(__FILE__, __LINE__, 0, true, false, __VA_ARGS__)
#define LOG_ERROR(...) Log_Write(__FILE__, __LINE__, 0, true, true, __VA_ARGS__)
