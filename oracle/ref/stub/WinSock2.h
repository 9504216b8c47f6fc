/* oracle/ref/stub/WinSock2.h -- TEST INFRASTRUCTURE: just enough for netphys_server/network_s.h to parse on Linux */
#pragma once
#include <netinet/in.h>
typedef unsigned long DWORD;
