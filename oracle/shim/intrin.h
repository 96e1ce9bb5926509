#pragma once
#include <immintrin.h>
