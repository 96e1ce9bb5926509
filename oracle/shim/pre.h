// Force-included before every reference TU by oracle/build_ref.sh (ours, not reference code).
#pragma once
#define _VCRT_ALIGN(x) __attribute__((aligned(x)))
#define _MM_ALIGN16 __attribute__((aligned(16)))
#include <string>
#include <cstring>
#include <functional>
#include <Windows.h>
