// Test-infrastructure shim (ours): the handful of Win32 names the reference's tracer and
// loader touch, mapped onto the C++ standard library.
#pragma once
#include <chrono>
#include <thread>
#include <string>
#include <cstdio>
#include <cstdint>
#include <functional>
union LARGE_INTEGER { long long QuadPart; };
inline int QueryPerformanceFrequency(LARGE_INTEGER *f) { f->QuadPart = 1000000000LL; return 1; }
inline int QueryPerformanceCounter(LARGE_INTEGER *c)
{
	c->QuadPart = std::chrono::duration_cast<std::chrono::nanoseconds>(
		std::chrono::steady_clock::now().time_since_epoch()).count();
	return 1;
}
inline void Sleep(unsigned ms) { std::this_thread::sleep_for(std::chrono::milliseconds(ms)); }
inline FILE *_wfopen(const wchar_t *name, const wchar_t *mode)
{
	std::string n, m;
	for (; *name; ++name) n.push_back((char)*name);
	for (; *mode; ++mode) m.push_back((char)*mode);
	return fopen(n.c_str(), m.c_str());
}
#pragma pack(push, 2)
struct BITMAPFILEHEADER { uint16_t bfType; uint32_t bfSize; uint16_t bfReserved1, bfReserved2; uint32_t bfOffBits; };
#pragma pack(pop)
struct BITMAPINFOHEADER
{
	uint32_t biSize; int32_t biWidth, biHeight; uint16_t biPlanes, biBitCount;
	uint32_t biCompression, biSizeImage; int32_t biXPelsPerMeter, biYPelsPerMeter;
	uint32_t biClrUsed, biClrImportant;
};
namespace std { namespace placeholders {} }
