// Bit-exact device restatements of the reference's per-primitive intersect operators
// (the DrawObject::intersect plugin surface, /root/reference/3DElement.h:201), plus the two box
// tests the Model path uses as culling predicates.  Compiled with -fmad=false.
#pragma once
#include "rt_device.cuh"

struct RayD
{
	F3 o, d;
	float mtlrfr;
	uint32_t skip;      // HitRes::obj of the surface the ray leaves (RT_ID_NONE for primary rays)
	uint8_t type;       // MY_RAY_*
	uint8_t isInside;   // 0x00 / 0xFF
};

// ---- Sphere::intersect, Basic3DObject.cpp:135-190 ----------------------------------------------
// returns t (1e20f = keep hr).  `self` = hr.obj == this.  The caller compares t with hr.distance.
__device__ __forceinline__ float sphere_t(const RayD &ray, const F3 &c, float r2, bool self)
{
	const F3 s2r = ray.o - c;
	const float b = dot(ray.d, s2r);
	if (self)
	{
		if (!ray.isInside)
			return 1e20f;
		// (shadow rays never carry isInside, so the reference's `return HitRes(radius)` is unreachable)
		const float dis = b * b - dot(s2r, s2r) + r2;
		const float t = -b + sqrtf(dis);
		return gt_1em6(t) ? t : 1e20f;
	}
	if (b > 0)
		return 1e20f;
	const float dis = b * b - dot(s2r, s2r) + r2;
	if (dis < 0)
		return 1e20f;
	const float t = -(b + sqrtf(dis));
	return gt_1em6(t) ? t : 1e20f;
}

// ---- BorderTest, Basic3DObject.cpp:44-81 -------------------------------------------------------
// `rrd` = 1.0f / direction, the reference's _mm_div_ps(1, direction); callers that already hold the
// ray's reciprocal direction pass it in (same IEEE division, same bits).
__device__ __forceinline__ float border_test(const F3 &o, const F3 &d, const F3 &rrd, const F3 &Min, const F3 &Max)
{
	const float rx = rrd.x, ry = rrd.y, rz = rrd.z;
	const float ax = (Min.x - o.x) * rx, bx = (Max.x - o.x) * rx;
	const float ay = (Min.y - o.y) * ry, by = (Max.y - o.y) * ry;
	const float az = (Min.z - o.z) * rz, bz = (Max.z - o.z) * rz;
	float minx = sse_min(ax, bx), maxx = sse_max(ax, bx);
	float miny = sse_min(ay, by), maxy = sse_max(ay, by);
	float minz = sse_min(az, bz), maxz = sse_max(az, bz);
	if (lt_1em6(fabsf(d.y)))
	{
		if (o.y > Max.y || o.y < Min.y) return 1e20f;
		miny = -1, maxy = 1e10f;
	}
	if (lt_1em6(fabsf(d.x)))
	{
		if (o.x > Max.x || o.x < Min.x) return 1e20f;
		minx = -1, maxx = 1e10f;
	}
	if (lt_1em6(fabsf(d.z)))
	{
		if (o.z > Max.z || o.z < Min.z) return 1e20f;
		minz = -1, maxz = 1e10f;
	}
	const float dmin = std_max(std_max(minx, miny), std_max(minz, 0.0f));
	const float dmax = std_min(std_min(maxx, maxy), maxz);
	if (dmax < dmin)
		return 1e20f;
	return dmin;
}

// ---- Box::intersect, Basic3DObject.cpp:274-300 -------------------------------------------------
__device__ __forceinline__ float box_t(const RayD &ray, const F3 &wmin, const F3 &wmax)
{
	const float res = border_test(ray.o, ray.d, f3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z), wmin, wmax);
	return gt_1em6(res) ? res : 1e20f;
}
__device__ __forceinline__ F3 box_normal(const F3 &P, const F3 &pos, const F3 &lmax)
{
	const F3 b2p = P - pos;
	F3 point = f3(0, 0, 0);
	if (lt_1em6(fabsf(fabsf(b2p.z) - lmax.z))) point.z = b2p.z > 0 ? 1.0f : -1.0f;
	if (lt_1em6(fabsf(fabsf(b2p.y) - lmax.y))) point.y = b2p.y > 0 ? 1.0f : -1.0f;
	if (lt_1em6(fabsf(fabsf(b2p.x) - lmax.x))) point.x = b2p.x > 0 ? 1.0f : -1.0f;
	return normalize(point);   // may be 0/0 = NaN, like the reference
}

// ---- Plane::intersect, Basic3DObject.cpp:376-410 -----------------------------------------------
__device__ __forceinline__ float plane_t(const RayD &ray, const F3 &pos, const F3 &n)
{
	const float a = dot(ray.d, n);
	if (lt_1em6(fabsf(a)))
		return 1e20f;
	const F3 p2r = ray.o - pos;
	const float b = dot(p2r, n);
	const float dis = -b / a;
	if (dis < 0)
		return 1e20f;
	return dis;
}
__device__ __forceinline__ float2 plane_tcoord(const F3 &o, const F3 &d, const F3 &pos, const F3 &axisx, const F3 &axisy)
{
	const F3 p2r = o - pos;
	const F3 tmp1 = cross(d, axisy);
	const float f = 1.0f / dot(axisx, tmp1) / 5;
	const float u = dot(p2r, tmp1) * f;
	const F3 tmp2 = cross(p2r, axisx);
	const float v = dot(d, tmp2) * f;
	return make_float2(u, v);
}

// ---- TriangleTest, Model.cpp:720-744 -----------------------------------------------------------
// returns t or 1e20f; bary = (1-u-v, u, v)
__device__ __forceinline__ float triangle_t(const F3 &o, const F3 &d, const F3 &e1, const F3 &e2, const F3 &p0, F3 *bary)
{
	const F3 tmp1 = cross(d, e2);
	const F3 t2r = o - p0;
	const float f = 1.0f / dot(e1, tmp1);
	const float u = dot(t2r, tmp1) * f;
	if (u < 0.0f || u > 1.0f)
		return 1e20f;
	const F3 tmp2 = cross(t2r, e1);
	const float v = dot(d, tmp2) * f, duv = 1 - u - v;
	if (v < 0.0f || duv < 0.0f)
		return 1e20f;
	const float t = dot(e2, tmp2) * f;
	if (t > 1e-5f)
	{
		if (bary) *bary = f3(duv, u, v);
		return t;
	}
	return 1e20f;
}

// ---- BorderTestEx, Model.cpp:482-664 -----------------------------------------------------------
// The 8 octants of a part's box in one pass.  Returns the reference's `minist`; *mask gets bit a
// set iff octant a (x half = a&4, y half = a&1, z half = a&2) passes ansmin <= ansmax.
static __device__ __noinline__ float border_test_ex(const F3 &o, const F3 &d, const F3 &rrd, const F3 &Min, const F3 &Max, uint32_t *mask)
{
	const F3 Mid = (Min + Max) * 0.5f;
	const float rx = rrd.x, ry = rrd.y, rz = rrd.z;
	const float x0 = (Min.x - o.x) * rx, x1 = (Mid.x - o.x) * rx, x2 = (Max.x - o.x) * rx;
	const float y0 = (Min.y - o.y) * ry, y1 = (Mid.y - o.y) * ry, y2 = (Max.y - o.y) * ry;
	const float z0 = (Min.z - o.z) * rz, z1 = (Mid.z - o.z) * rz, z2 = (Max.z - o.z) * rz;
	// per axis: [lo half min, lo half max, hi half min, hi half max]
	float xmin[2] = { sse_min(x0, x1), sse_min(x1, x2) }, xmax[2] = { sse_max(x0, x1), sse_max(x1, x2) };
	float ymin[2] = { sse_min(y0, y1), sse_min(y1, y2) }, ymax[2] = { sse_max(y0, y1), sse_max(y1, y2) };
	float zmin[2] = { sse_min(z0, z1), sse_min(z1, z2) }, zmax[2] = { sse_max(z0, z1), sse_max(z1, z2) };
	*mask = 0;
	if (lt_1em6(fabsf(d.y)))
	{
		if (o.y > Max.y || o.y < Min.y) return 1e20f;
		const bool hi = o.y > Mid.y, lo = o.y < Mid.y;   // neither: both halves open
		ymin[0] = hi ? 1e20f : 0.0f, ymax[0] = hi ? 0.0f : 1e20f;
		ymin[1] = lo ? 1e20f : 0.0f, ymax[1] = lo ? 0.0f : 1e20f;
	}
	if (lt_1em6(fabsf(d.x)))
	{
		if (o.x > Max.x || o.x < Min.x) return 1e20f;
		const bool hi = o.x > Mid.x, lo = o.x < Mid.x;
		xmin[0] = hi ? 1e20f : 0.0f, xmax[0] = hi ? 0.0f : 1e20f;
		xmin[1] = lo ? 1e20f : 0.0f, xmax[1] = lo ? 0.0f : 1e20f;
	}
	if (lt_1em6(fabsf(d.z)))
	{
		if (o.z > Max.z || o.z < Min.z) return 1e20f;
		const bool hi = o.z > Mid.z, lo = o.z < Mid.z;
		zmin[0] = hi ? 1e20f : 0.0f, zmax[0] = hi ? 0.0f : 1e20f;
		zmin[1] = lo ? 1e20f : 0.0f, zmax[1] = lo ? 0.0f : 1e20f;
	}
	float minist = 1e20f;
	uint32_t m = 0;
#pragma unroll
	for (int a = 0; a < 8; ++a)
	{
		const int hx = (a >> 2) & 1, hy = a & 1, hz = (a >> 1) & 1;
		float ansmin = sse_max(xmin[hx], ymin[hy]), ansmax = sse_min(xmax[hx], ymax[hy]);
		const float ttansmin = sse_max(zmin[hz], 0.0f);
		ansmin = sse_max(ansmin, ttansmin);
		ansmax = sse_min(ansmax, zmax[hz]);
		if (ansmin <= ansmax)
		{
			m |= 1u << a;
			minist = std_min(minist, ansmin);
		}
	}
	*mask = m;
	return minist;
}
