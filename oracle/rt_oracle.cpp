// TEST INFRASTRUCTURE -- CPU restatement of the reference's per-pixel trace-and-shade path.
//
// This file is the parity ORACLE for the CUDA path.  It is never linked into, imported by or
// executed from the product (raytrace_b200/): only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load liboracle.so.
//
// It restates, in scalar C++ with one IEEE-754 binary32 rounding per source-level operation
// (build with -ffp-contract=off, no -ffast-math), the algorithm of XZiar/RayTrace:
//   RayTracer::parallelRT   /root/reference/RayTracer.cpp:5-45
//   RayTracer::RTfrac       /root/reference/RayTracer.cpp:450-596   (and RTflec :332-448)
//   Sphere::intersect       /root/reference/Basic3DObject.cpp:135-190
//   BorderTest, Box::intersect      Basic3DObject.cpp:44-81, :274-300
//   Plane::intersect        Basic3DObject.cpp:376-410
//   Model::RTPrepare / BorderTestEx / TriangleTest / Model::intersect
//                           /root/reference/Model.cpp:402-480, :482-664, :720-744, :748-811
//   Color(tex,coord), Color::put    /root/reference/3DElement.cpp:430-450, :463-468
// operating on the flattened scene of include/rt_b200.h (the same bytes the GPU receives).
//
// PARITY PINNED: validated bit-for-bit (image hash, primary hit ids, ray counts) against the
// reference itself compiled here (oracle/_ref/ref_render, see oracle/build_ref.sh); the committed
// fixtures in tests/golden/ were produced by that reference binary, and tests/test_oracle.py
// checks this restatement against them.  The reference ships no golden vectors of its own
// (SURVEY.md section 4).
#include "../include/rt_b200.h"

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace
{

struct V4
{
	float x, y, z, w;
};
inline V4 mk(float x, float y, float z, float w = 0) { return V4{ x, y, z, w }; }
inline V4 from(const rt_vec4 &v) { return V4{ v.x, v.y, v.z, v.w }; }
inline V4 add(const V4 &a, const V4 &b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline V4 sub(const V4 &a, const V4 &b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline V4 mul(const V4 &a, float s) { return mk(a.x * s, a.y * s, a.z * s, a.w * s); }
inline V4 mixmul(const V4 &a, const V4 &b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
// dpps mask 0x71: (x0y0 + x1y1) + (x2y2 + 0)   3DElement.cpp:206-214
inline float dot(const V4 &a, const V4 &b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + 0.0f); }
// 3DElement.cpp:190-197
inline V4 cross(const V4 &a, const V4 &b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, a.w * b.w - a.w * b.w); }
// 3DElement.cpp:218-238: true divide by sqrt
inline V4 normalize(const V4 &v)
{
	const float len = std::sqrt(dot(v, v));
	return mk(v.x / len, v.y / len, v.z / len, v.w / len);
}
// SSE min/max: second operand wins on NaN / equality
inline float sse_min(float a, float b) { return a < b ? a : b; }
inline float sse_max(float a, float b) { return a > b ? a : b; }
// std::min / std::max as the reference calls them
inline float std_min(float a, float b) { return b < a ? b : a; }
inline float std_max(float a, float b) { return a < b ? b : a; }

constexpr int64_t OBJ_NONE = -1;
inline int64_t prim_id(uint32_t flat) { return (int64_t)flat; }
inline int64_t tri_id(uint32_t tri, uint32_t oct) { return ((int64_t)1 << 40) | ((int64_t)oct << 32) | tri; }

struct RayO
{
	V4 origin, direction;
	float mtlrfr = 1.0f;
	uint8_t type = 0, isInside = 0;
};

struct Hit   // HitRes, 3DElement.h:166-183
{
	V4 position = mk(0, 0, 0), normal = mk(0, 0, 0);
	float tu = 0, tv = 0;
	int mtl = -1, tex = -1;
	int64_t obj = OBJ_NONE;
	float distance = 1e20f, rfr = 1.0f;
	uint8_t isInside = 0;
	explicit Hit(float d = 1e20f) : distance(d) {}
};

struct OctTri   // clTri, 3DElement.h:122-127 (+ which triangle it copies)
{
	V4 axisu, axisv, p0;
	int16_t numa, numb;
	uint32_t tri;
};

struct ModelPrep   // what Model::RTPrepare leaves behind, Model.cpp:402-480
{
	V4 borderMin, borderMax;
	std::vector<V4> bboxs;                       // 2 per part
	std::vector<std::vector<OctTri>> oct;        // 8 per part
};

struct Item { int kind; uint32_t index; };       // scene order: kind 0 = prim, 1 = model

struct Prepared
{
	const rt_scene_desc *s;
	std::vector<ModelPrep> models;
	std::vector<Item> items;
};

// Model::RTPrepare, Model.cpp:402-480
void prepare(const rt_scene_desc &s, Prepared &out)
{
	out.s = &s;
	out.models.resize(s.n_models);
	for (uint32_t m = 0; m < s.n_models; ++m)
	{
		const rt_model &rm = s.models[m];
		ModelPrep &mp = out.models[m];
		const V4 position = from(rm.position);
		mp.borderMin = add(from(rm.ver_min), position), mp.borderMax = add(from(rm.ver_max), position);
		mp.oct.resize((size_t)rm.part_count * 8);
		for (uint32_t cnta = 0; cnta < rm.part_count; ++cnta)
		{
			const rt_part &part = s.parts[rm.part_begin + cnta];
			V4 va = from(part.border_min), vb = from(part.border_max);
			mp.bboxs.push_back(add(va, position)), mp.bboxs.push_back(add(vb, position));
			va = mul(add(va, vb), 0.5f);
			for (uint32_t cntb = 0; cntb < part.tri_count; ++cntb)
			{
				const uint32_t t = part.tri_begin + cntb;
				const V4 p0 = from(s.tri_points[3 * t]), p1 = from(s.tri_points[3 * t + 1]), p2 = from(s.tri_points[3 * t + 2]);
				OctTri clt{ sub(p1, p0), sub(p2, p0), add(p0, position), (int16_t)cnta, (int16_t)cntb, t };
				const float tminx = sse_min(p0.x, sse_min(p1.x, p2.x)), tminy = sse_min(p0.y, sse_min(p1.y, p2.y)), tminz = sse_min(p0.z, sse_min(p1.z, p2.z));
				const float tmaxx = sse_max(p0.x, sse_max(p1.x, p2.x)), tmaxy = sse_max(p0.y, sse_max(p1.y, p2.y)), tmaxz = sse_max(p0.z, sse_max(p1.z, p2.z));
				std::vector<OctTri> *o = &mp.oct[(size_t)cnta * 8];
				if (tminx <= va.x)
				{
					if (tminz <= va.z)
					{
						if (tminy <= va.y) o[0].push_back(clt);
						if (tmaxy >= va.y) o[1].push_back(clt);
					}
					if (tmaxz >= va.z)
					{
						if (tminy <= va.y) o[2].push_back(clt);
						if (tmaxy >= va.y) o[3].push_back(clt);
					}
				}
				if (tmaxx >= va.x)
				{
					if (tminz <= va.z)
					{
						if (tminy <= va.y) o[4].push_back(clt);
						if (tmaxy >= va.y) o[5].push_back(clt);
					}
					if (tmaxz >= va.z)
					{
						if (tminy <= va.y) o[6].push_back(clt);
						if (tmaxy >= va.y) o[7].push_back(clt);
					}
				}
			}
		}
	}
	// scene order = Objects order (RayTracer.cpp:458): merge prims and models by object index
	uint32_t pi = 0, mi = 0;
	while (pi < s.n_prims || mi < s.n_models)
	{
		const bool takePrim = mi >= s.n_models || (pi < s.n_prims && s.prims[pi].object < s.models[mi].object);
		if (takePrim) out.items.push_back(Item{ 0, pi++ });
		else out.items.push_back(Item{ 1, mi++ });
	}
}

// BorderTest, Basic3DObject.cpp:44-81
float BorderTest(const RayO &ray, const V4 &Min, const V4 &Max, float *getMax)
{
	V4 tdismin = sub(Min, ray.origin), tdismax = sub(Max, ray.origin);
	const V4 rrd = mk(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z, 1.0f / ray.direction.w);
	tdismin = mixmul(tdismin, rrd);
	tdismax = mixmul(tdismax, rrd);
	V4 dismin = mk(sse_min(tdismin.x, tdismax.x), sse_min(tdismin.y, tdismax.y), sse_min(tdismin.z, tdismax.z)),
		dismax = mk(sse_max(tdismin.x, tdismax.x), sse_max(tdismin.y, tdismax.y), sse_max(tdismin.z, tdismax.z));
	if (std::fabs(ray.direction.y) < 1e-6)
	{
		if (ray.origin.y > Max.y || ray.origin.y < Min.y)
			return 1e20f;
		dismin.y = -1, dismax.y = 1e10f;
	}
	if (std::fabs(ray.direction.x) < 1e-6)
	{
		if (ray.origin.x > Max.x || ray.origin.x < Min.x)
			return 1e20f;
		dismin.x = -1, dismax.x = 1e10f;
	}
	if (std::fabs(ray.direction.z) < 1e-6)
	{
		if (ray.origin.z > Max.z || ray.origin.z < Min.z)
			return 1e20f;
		dismin.z = -1, dismax.z = 1e10f;
	}
	const float dmin = std_max(std_max(dismin.x, dismin.y), std_max(dismin.z, 0.0f)),
		dmax = std_min(std_min(dismax.x, dismax.y), dismax.z);
	if (dmax < dmin)
		return 1e20;
	*getMax = dmax;
	return dmin;
}

// BorderTestEx, Model.cpp:482-664: the 8 octants of [Min,Max] split at Mid in one pass.
// Lane a: x half = a&4, y half = a&1, z half = a&2.  Writes mask only when it does not bail out.
float BorderTestEx(const RayO &ray, const V4 &Min, const V4 &Max, bool mask[8])
{
	const V4 Mid = mul(add(Min, Max), 0.5f);
	V4 tmin = sub(Min, ray.origin), tmax = sub(Max, ray.origin), tmid = sub(Mid, ray.origin);
	const V4 rrd = mk(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z, 1.0f / ray.direction.w);
	tmin = mixmul(tmin, rrd), tmax = mixmul(tmax, rrd), tmid = mixmul(tmid, rrd);
	float txmin[8], txmax[8], tymin[8], tymax[8], tzmin[8], tzmax[8];
	for (int a = 0; a < 8; ++a)
	{
		const float xa = (a & 4) ? tmid.x : tmin.x, xb = (a & 4) ? tmax.x : tmid.x;
		const float ya = (a & 1) ? tmid.y : tmin.y, yb = (a & 1) ? tmax.y : tmid.y;
		const float za = (a & 2) ? tmid.z : tmin.z, zb = (a & 2) ? tmax.z : tmid.z;
		txmin[a] = sse_min(xa, xb), txmax[a] = sse_max(xa, xb);
		tymin[a] = sse_min(ya, yb), tymax[a] = sse_max(ya, yb);
		tzmin[a] = sse_min(za, zb), tzmax[a] = sse_max(za, zb);
	}
	auto flat = [](float omin[8], float omax[8], int bit, float o, float mid)
	{
		for (int a = 0; a < 8; ++a)
		{
			const bool high = (a & bit) != 0;
			bool open;
			if (o > mid) open = high;
			else if (o < mid) open = !high;
			else { omin[a] = 0, omax[a] = 1e20f; continue; }
			omin[a] = open ? 0 : 1e20f, omax[a] = open ? 1e20f : 0;
		}
	};
	if (std::fabs(ray.direction.y) < 1e-6)
	{
		if (ray.origin.y > Max.y || ray.origin.y < Min.y)
			return 1e20f;
		flat(tymin, tymax, 1, ray.origin.y, Mid.y);
	}
	if (std::fabs(ray.direction.x) < 1e-6)
	{
		if (ray.origin.x > Max.x || ray.origin.x < Min.x)
			return 1e20f;
		flat(txmin, txmax, 4, ray.origin.x, Mid.x);
	}
	if (std::fabs(ray.direction.z) < 1e-6)
	{
		if (ray.origin.z > Max.z || ray.origin.z < Min.z)
			return 1e20f;
		flat(tzmin, tzmax, 2, ray.origin.z, Mid.z);
	}
	float minist = 1e20f;
	for (int a = 0; a < 8; ++a)
	{
		float ansmin = sse_max(txmin[a], tymin[a]), ansmax = sse_min(txmax[a], tymax[a]);
		const float ttansmin = sse_max(tzmin[a], 0.0f);
		ansmin = sse_max(ansmin, ttansmin);
		ansmax = sse_min(ansmax, tzmax[a]);
		mask[a] = ansmin <= ansmax;
		if (mask[a])
			minist = std_min(minist, ansmin);
	}
	return minist;
}

// TriangleTest, Model.cpp:720-744
inline float TriangleTest(const RayO &ray, const OctTri &tri, V4 &coord)
{
	const V4 tmp1 = cross(ray.direction, tri.axisv);
	const V4 t2r = sub(ray.origin, tri.p0);
	float f = dot(tri.axisu, tmp1);
	f = 1.0f / f;
	const float u = dot(t2r, tmp1) * f;
	if (u < 0.0f || u > 1.0f)
		return 1e20f;
	const V4 tmp2 = cross(t2r, tri.axisu);
	const float v = dot(ray.direction, tmp2) * f, duv = 1 - u - v;
	if (v < 0.0f || duv < 0.0f)
		return 1e20f;
	const float t = dot(tri.axisv, tmp2) * f;
	if (t > 1e-5f)
	{
		coord = mk(duv, u, v);
		return t;
	}
	return 1e20f;
}

struct Tracer
{
	const Prepared &P;
	const rt_scene_desc &s;
	uint32_t maxLevel;
	uint64_t nPrimary = 0, nShadow = 0, nReflect = 0, nRefract = 0;

	Tracer(const Prepared &p, uint32_t lvl) : P(p), s(*p.s), maxLevel(lvl) {}

	void count(const RayO &r)
	{
		switch (r.type)
		{
		case 1: ++nPrimary; break;
		case 3: ++nReflect; break;
		case 4: ++nRefract; break;
		default: ++nShadow; break;
		}
	}

	// Sphere::intersect, Basic3DObject.cpp:135-190 (also one lattice slot of BallPlane::intersect :486-552)
	Hit sphere(uint32_t flat, const rt_prim &sp, const RayO &ray, const Hit &hr) const
	{
		const V4 position = from(sp.position);
		if (hr.obj == prim_id(flat))
		{
			if (!ray.isInside)
				return hr;
			if (ray.type == 2)
				return Hit(sp.radius);
			const V4 s2r = sub(ray.origin, position);
			const float rdDOTr2s = dot(ray.direction, s2r);
			const float dis = rdDOTr2s * rdDOTr2s - dot(s2r, s2r) + sp.radius_sqr;
			const float t = -dot(ray.direction, s2r) + std::sqrt(dis);
			if (t < hr.distance && t > 1e-6)
			{
				Hit nh(t);
				nh.position = add(ray.origin, mul(ray.direction, t));
				nh.normal = normalize(sub(position, nh.position));
				nh.mtl = (int)sp.material;
				nh.obj = prim_id(flat);
				nh.isInside = (uint8_t)~ray.isInside;
				nh.rfr = ray.type == 4 ? 1.0f : s.materials[sp.material].rfr;
				return nh;
			}
			return hr;
		}
		const V4 s2r = sub(ray.origin, position);
		const float rdDOTr2s = dot(ray.direction, s2r);
		if (rdDOTr2s > 0)
			return hr;
		const float dis = rdDOTr2s * rdDOTr2s - dot(s2r, s2r) + sp.radius_sqr;
		if (dis < 0)
			return hr;
		const float t = -(dot(ray.direction, s2r) + std::sqrt(dis));
		if (t < hr.distance && t > 1e-6)
		{
			Hit nh(t);
			nh.position = add(ray.origin, mul(ray.direction, t));
			nh.normal = normalize(sub(nh.position, position));
			nh.mtl = (int)sp.material;
			nh.obj = prim_id(flat);
			nh.isInside = (uint8_t)~ray.isInside;
			nh.rfr = s.materials[sp.material].rfr;
			return nh;
		}
		return hr;
	}

	// Box::intersect, Basic3DObject.cpp:274-300
	Hit box(uint32_t flat, const rt_prim &b, const RayO &ray, const Hit &hr) const
	{
		if (hr.obj == prim_id(flat))
			return hr;
		const V4 position = from(b.position), bmin = from(b.a), bmax = from(b.b);
		float empty;
		const float res = BorderTest(ray, add(bmin, position), add(bmax, position), &empty);
		if (res < hr.distance && res > 1e-6)
		{
			Hit nh(res);
			nh.position = add(ray.origin, mul(ray.direction, res));
			const V4 b2p = sub(nh.position, position);
			V4 point = mk(0, 0, 0);
			if (std::fabs(std::fabs(b2p.z) - bmax.z) < 1e-6)
				point.z = b2p.z > 0 ? 1 : -1;
			if (std::fabs(std::fabs(b2p.y) - bmax.y) < 1e-6)
				point.y = b2p.y > 0 ? 1 : -1;
			if (std::fabs(std::fabs(b2p.x) - bmax.x) < 1e-6)
				point.x = b2p.x > 0 ? 1 : -1;
			nh.normal = normalize(point);
			nh.mtl = (int)b.material;
			nh.obj = prim_id(flat);
			return nh;
		}
		return hr;
	}

	// Plane::intersect, Basic3DObject.cpp:376-410
	Hit plane(uint32_t flat, const rt_prim &p, const RayO &ray, const Hit &hr) const
	{
		if (hr.obj == prim_id(flat))
			return hr;
		const V4 position = from(p.position), normal = from(p.a), axisx = from(p.b), axisy = from(p.c);
		const float a = dot(ray.direction, normal);
		if (std::fabs(a) < 1e-6)
			return hr;
		const V4 p2r = sub(ray.origin, position);
		const float b = dot(p2r, normal);
		const float dis = -b / a;
		if (dis < 0)
			return hr;
		if (dis < hr.distance)
		{
			const V4 tmp1 = cross(ray.direction, axisy);
			const float f = 1.0f / dot(axisx, tmp1) / 5;
			const float u = dot(p2r, tmp1) * f;
			const V4 tmp2 = cross(p2r, axisx);
			const float v = dot(ray.direction, tmp2) * f;
			Hit nh(dis);
			nh.normal = normal;
			nh.mtl = (int)p.material;
			nh.tex = p.texture;
			nh.position = add(ray.origin, mul(ray.direction, dis));
			nh.tu = u, nh.tv = v;
			nh.obj = prim_id(flat);
			return nh;
		}
		return hr;
	}

	// Model::intersect, Model.cpp:748-811
	Hit model(uint32_t m, const RayO &ray, const Hit &hr, const float min) const
	{
		const rt_model &rm = s.models[m];
		const ModelPrep &mp = P.models[m];
		float empty;
		bool mask[8] = { false, false, false, false, false, false, false, false };
		float ans = BorderTest(ray, mp.borderMin, mp.borderMax, &empty);
		if (ans < hr.distance)
		{
			ans = hr.distance;
			int objpart = -1;
			const OctTri *objclt = nullptr;
			int objoct = -1;
			V4 coord = mk(0, 0, 0), tmpc = mk(0, 0, 0);
			const int pcnt = (int)rm.part_count;
			for (int a = 0; a < pcnt; ++a)
				if (BorderTestEx(ray, mp.bboxs[a * 2], mp.bboxs[a * 2 + 1], mask) < hr.distance)
					for (int b = 0, pcur = a * 8; b < 8; ++b, ++pcur)
						if (mask[b])
						{
							const std::vector<OctTri> &list = mp.oct[pcur];
							for (size_t c = 0; c < list.size(); ++c)
							{
								const OctTri &t = list[c];
								if (hr.obj == tri_id(t.tri, (uint32_t)b))
									continue;
								const float newans = TriangleTest(ray, t, tmpc);
								if (newans < ans)
								{
									objpart = a, objclt = &t, objoct = b;
									ans = newans;
									coord = tmpc;
									if (newans < min)
										goto ____EOS;
								}
							}
						}
		____EOS:
			if (ans < hr.distance)
			{
				const uint32_t t = objclt->tri;
				Hit nh(ans);
				nh.position = add(ray.origin, mul(ray.direction, ans));
				const V4 n0 = from(s.tri_norms[3 * t]), n1 = from(s.tri_norms[3 * t + 1]), n2 = from(s.tri_norms[3 * t + 2]);
				nh.normal = normalize(add(add(mul(n0, coord.x), mul(n1, coord.y)), mul(n2, coord.z)));
				const float *tc = s.tri_tcoords + 6 * (size_t)t;
				nh.tu = (tc[0] * coord.x + tc[2] * coord.y) + tc[4] * coord.z;
				nh.tv = (tc[1] * coord.x + tc[3] * coord.y) + tc[5] * coord.z;
				const rt_part &part = s.parts[rm.part_begin + objpart];
				nh.mtl = (int)part.material;
				nh.rfr = s.materials[part.material].rfr;
				nh.tex = part.texture;
				nh.obj = tri_id(t, (uint32_t)objoct);
				return nh;
			}
		}
		return hr;
	}

	Hit intersect(const Item &it, const RayO &ray, const Hit &hr, const float min = 0) const
	{
		if (it.kind == 1)
			return model(it.index, ray, hr, min);
		const rt_prim &p = s.prims[it.index];
		switch (p.kind)
		{
		case RT_OBJ_SPHERE: return sphere(it.index, p, ray, hr);
		case RT_OBJ_CUBE: return box(it.index, p, ray, hr);
		default: return plane(it.index, p, ray, hr);
		}
	}

	// Color(const Texture*, Coord2D), 3DElement.cpp:430-450
	V4 texel(int tex, float cu, float cv) const
	{
		if (tex < 0)
			return mk(1.0f, 1.0f, 1.0f);
		const rt_texture &t = s.textures[tex];
		float whole;
		float nu = std::modf(cu, &whole), nv = std::modf(cv, &whole);
		if (nu < 0) nu += 1;
		if (nv < 0) nv += 1;
		const int16_t x = (int16_t)(nu * t.w), y = (int16_t)(nv * t.h);
		const uint8_t *px = s.texels + t.offset + (y * t.w + x) * 3;
		return mk(px[2] / 255.0f, px[1] / 255.0f, px[0] / 255.0f);
	}

	// closest hit over the scene in object order, RayTracer.cpp:455-465
	Hit closest(const RayO &ray, const Hit &basehr, int64_t &newobj)
	{
		count(ray);
		Hit hr = basehr;
		newobj = basehr.obj;
		for (const Item &it : P.items)
		{
			hr.obj = basehr.obj;
			hr = intersect(it, ray, hr);
			if (hr.obj != basehr.obj)
				newobj = hr.obj;
		}
		return hr;
	}

	// The staged debug shaders RTdepth :81, RTnorm :94, RTtex :109, RTmtl :127, RTshd :222
	// (RayTracer.cpp).  They run the object loop WITHOUT resetting hr.obj, which cannot change the
	// result for a primary ray (every object is visited once), and shade one level only.
	V4 debug(uint32_t type, float zNear, float zFar, const RayO &baseray)
	{
		int64_t newobj;
		const Hit hr = closest(baseray, Hit(), newobj);
		if (type == RT_TYPE_DEPTH)
		{
			// Color::set, 3DElement.cpp:451-462
			if (hr.distance <= zNear) return mk(1.0f, 0.0f, 0.0f);
			if (hr.distance >= zFar) return mk(0.0f, 0.0f, 0.0f);
			const float after = std::log(hr.distance), mx = std::log(zFar);
			const float g = (mx - after) / mx;
			return mk(g, g, g);
		}
		if (hr.distance > zFar) return mk(0.0f, 0.0f, 0.0f);
		if (hr.distance < zNear) return mk(1.0f, 1.0f, 1.0f);
		if (type == RT_TYPE_NORMAL)   // Color(const Normal&), 3DElement.cpp:423-428
			return mk((float)(0.5 * (hr.normal.x + 1)), (float)(0.5 * (hr.normal.y + 1)), (float)(0.5 * (hr.normal.z + 1)));
		if (type == RT_TYPE_TEXTURE)
			return hr.tex >= 0 ? texel(hr.tex, hr.tu, hr.tv) : mk(0.588f, 0.588f, 0.588f);
		const rt_material &mtl = s.materials[hr.mtl];
		const V4 vc = texel(hr.tex, hr.tu, hr.tv);
		V4 mix_vd = mk(0, 0, 0), mix_va = mk(0, 0, 0), mix_vsc = mk(0, 0, 0);
		for (uint32_t li = 0; li < s.n_lights; ++li)
		{
			const rt_light &lit = s.lights[li];
			if (!lit.enabled)
				continue;
			float light_lum, dis = 1e10;
			V4 p2l;
			const bool point = type == RT_TYPE_MATERIAL ? lit.position.w > 1e-6 : lit.type == RT_LIGHT_POINT;
			if (point)
			{
				const V4 p2l_v = sub(from(lit.position), hr.position);
				dis = dot(p2l_v, p2l_v);
				float step;
				if (type == RT_TYPE_MATERIAL)   // RTmtl sums in a different order, RayTracer.cpp:153-155
					step = lit.attenuation.x + lit.attenuation.y * std::sqrt(dis) + lit.attenuation.z * dis;
				else
				{
					step = lit.attenuation.x + lit.attenuation.z * dis;
					dis = std::sqrt(dis);
					step += lit.attenuation.y * dis;
				}
				light_lum = 1 / step;
				p2l = normalize(p2l_v);
			}
			else
			{
				light_lum = 1.0f;
				p2l = normalize(from(lit.position));
			}
			const V4 light_a = mul(from(lit.ambient), light_lum), light_d = mul(from(lit.diffuse), light_lum), light_s = mul(from(lit.specular), light_lum);
			mix_va = add(mix_va, mixmul(from(mtl.ambient), light_a));
			if (type == RT_TYPE_SHADOW)
			{
				RayO shadowray;
				shadowray.origin = hr.position, shadowray.direction = p2l, shadowray.type = 0;
				count(shadowray);
				Hit shr(dis);
				shr.obj = newobj;
				bool blocked = false;
				for (const Item &it : P.items)
				{
					shr = intersect(it, shadowray, shr, dis);
					if (shr.distance < dis) { blocked = true; break; }
				}
				if (blocked)
					continue;
			}
			float n_n = dot(hr.normal, p2l);
			if (n_n > 0)
				mix_vd = add(mix_vd, mul(mixmul(from(mtl.diffuse), light_d), n_n));
			const V4 h = normalize(sub(p2l, baseray.direction));
			n_n = dot(hr.normal, h);
			if (n_n > 0)
				mix_vsc = add(mix_vsc, mul(mixmul(from(mtl.specular), light_s), std::pow(n_n, mtl.shiness)));
		}
		mix_va = add(mix_va, mixmul(from(mtl.ambient), from(s.env_light)));   // environment term LAST here
		return add(mixmul(vc, add(mix_vd, mix_va)), mix_vsc);
	}

	// RTfrac (refraction = true) / RTflec (false); returns rgb + alpha = hit distance
	V4 shade(float zNear, float zFar, const RayO &baseray, uint32_t level, float bwc, const Hit &basehr, bool refraction)
	{
		if (level > maxLevel || bwc < 1e-5f)
			return mk(0, 0, 0, 1e20f);
		int64_t newobj;
		const Hit hr = closest(baseray, basehr, newobj);
		if (hr.distance > zFar || hr.distance < zNear)
			return mk(0, 0, 0, 1e20f);
		const rt_material &mtl = s.materials[hr.mtl];
		const V4 vc = texel(hr.tex, hr.tu, hr.tv);
		V4 mix_vd = mk(0, 0, 0), mix_vsc = mk(0, 0, 0);
		V4 mix_va = mixmul(from(mtl.ambient), from(s.env_light));
		for (uint32_t li = 0; li < s.n_lights; ++li)
		{
			const rt_light &lit = s.lights[li];
			if (!lit.enabled)
				continue;
			V4 light_a, light_d, light_s, p2l;
			float dis;
			if (lit.type == RT_LIGHT_POINT)
			{
				const V4 p2l_v = sub(from(lit.position), hr.position);
				dis = dot(p2l_v, p2l_v);
				float step = lit.attenuation.x + lit.attenuation.z * dis;
				dis = std::sqrt(dis);
				step += lit.attenuation.y * dis;
				const float light_lum = 1 / step;
				light_a = mul(from(lit.ambient), light_lum);
				light_d = mul(from(lit.diffuse), light_lum);
				light_s = mul(from(lit.specular), light_lum);
				p2l = normalize(p2l_v);
			}
			else
			{
				dis = 1e10;
				light_a = from(lit.ambient), light_d = from(lit.diffuse), light_s = from(lit.specular);
				p2l = normalize(from(lit.position));
			}
			mix_va = add(mix_va, mixmul(from(mtl.ambient), light_a));
			// shadow any-hit, RayTracer.cpp:510-520
			RayO shadowray;
			shadowray.origin = hr.position, shadowray.direction = p2l, shadowray.type = refraction ? 2 : 0;
			count(shadowray);
			Hit shr(dis);
			shr.obj = newobj;
			bool blocked = false;
			for (const Item &it : P.items)
			{
				shr = intersect(it, shadowray, shr, dis);
				if (shr.distance < dis)
				{
					blocked = true;
					break;
				}
			}
			if (blocked)
				continue;
			float n_n = dot(hr.normal, p2l);
			if (n_n > 0)
				mix_vd = add(mix_vd, mul(mixmul(from(mtl.diffuse), light_d), n_n));
			const V4 h = normalize(sub(p2l, baseray.direction));
			n_n = dot(hr.normal, h);
			if (n_n > 0)
			{
				const V4 vs = mul(mixmul(from(mtl.specular), light_s), std::pow(n_n, mtl.shiness));
				mix_vsc = add(mix_vsc, vs);
			}
		}
		V4 c_all = add(mixmul(vc, add(mix_vd, mix_va)), mix_vsc);
		if (mtl.reflect > 0.01f)
		{
			const float flecrate = mtl.reflect;
			c_all = mul(c_all, 1 - flecrate);
			const float n_n = 2 * dot(baseray.direction, hr.normal);
			RayO flecray;
			flecray.origin = hr.position;
			flecray.direction = normalize(sub(baseray.direction, mul(hr.normal, n_n)));
			flecray.type = refraction ? 3 : 0;
			Hit flechr;
			flechr.obj = newobj;
			const V4 c_flec = shade(0.0f, zFar, flecray, level + 1, bwc * flecrate, flechr, refraction);
			c_all = add(c_all, mul(c_flec, flecrate));
		}
		if (refraction && mtl.refract > 0.01f)
		{
			const float fracrate = mtl.refract;
			c_all = mul(c_all, 1 - fracrate);
			const float n = baseray.mtlrfr / hr.rfr;
			const float cosIn = -dot(baseray.direction, hr.normal);
			const float cosOut2 = 1.0f - (n * n) * (1.0f - cosIn * cosIn);
			if (!(cosOut2 < 0.0f))
			{
				const V4 l2 = mul(baseray.direction, n), l1 = mul(hr.normal, n * cosIn - std::sqrt(cosOut2));
				RayO fracray;
				fracray.origin = hr.position;
				fracray.direction = normalize(add(l1, l2));
				fracray.type = 4;
				fracray.mtlrfr = hr.rfr;
				fracray.isInside = hr.isInside;
				Hit frachr;
				frachr.obj = newobj;
				const V4 c_frac = shade(0.0f, zFar, fracray, level + 1, bwc * fracrate, frachr, refraction);
				V4 vc_frac = mk(1, 1, 1);
				if (hr.isInside)
				{
					const V4 e = mul(mul(from(mtl.diffuse), 0.15f), -c_frac.w);
					vc_frac = mk(std::exp(e.x), std::exp(e.y), std::exp(e.z));
				}
				c_all = add(c_all, mul(mixmul(c_frac, vc_frac), fracrate));
			}
		}
		c_all.w = hr.distance;
		return c_all;
	}
};

// Color::put, 3DElement.cpp:463-468 (NaN -> 0 like cvttss2si's low byte)
inline uint8_t put1(float c)
{
	if (c > 1.0f) return 255;
	if (c < 0.0f) return 0;
	const float v = c * 255;
	if (!(v == v)) return 0;
	return (uint8_t)v;
}

}  // namespace

extern "C" {

// Renders the frame exactly as RayTracer::start + parallelRT do (RayTracer.cpp:5-45,614-696).
// rgb: width*height*3 bytes, filled with 127 first; ids/counters may be NULL.
int rto_render(const rt_scene_desc *scene, const rt_render_params *params, uint8_t *rgb, rt_hit_id *ids,
	rt_counters *counters, int threads)
{
	if (!scene || !params || !rgb)
		return RT_E_INVALID;
	const uint32_t type = params->type;
	if (type != RT_TYPE_RAYTRACE && (type < RT_TYPE_CHECK || type > RT_TYPE_REFRACT))
		return RT_E_INVALID;
	Prepared P;
	prepare(*scene, P);
	const rt_camera &cam = scene->camera;
	const int width = cam.width, height = cam.height;
	memset(rgb, 127, (size_t)width * height * 3);
	const int blk_h = height / 64, blk_w = width / 64;
	const double dp = tan(cam.fovy * 3.1415926535897932384 / 360) / (height / 2);
	const float zNear = cam.zNear, zFar = sqrt(2) * cam.zFar;
	const uint32_t world = params->world > 1 ? params->world : 1, rank = params->world > 1 ? params->rank : 0;
	const uint32_t tileRows = params->tile_rows ? params->tile_rows : 64u;
	// which rank renders image row y (include/rt_b200.h: tile t -> rank t % world; RT_FLAG_SERPENTINE deals the
	// odd groups of `world` tiles in reverse order)
	const bool serpentine = (params->flags & RT_FLAG_SERPENTINE) && world > 1;
	// tile window (rt_render_params::tile_first / tile_count): only the shard's own tiles first .. first + count - 1;
	// a row outside it has no owner here
	const uint32_t tileFirst = params->tile_first, tileCount = params->tile_count;
	auto ownerOfRow = [=](uint32_t y) -> uint32_t
	{
		const uint32_t t = y / tileRows, g = t / world, i = t % world;
		if (g < tileFirst || (tileCount && g >= tileFirst + tileCount))
			return 0xFFFFFFFFu;   // tile g*world + .. is the g-th tile of whichever rank owns it
		return (serpentine && (g & 1u)) ? world - 1u - i : i;
	};
	if (threads < 1) threads = 1;
	std::atomic<int> nextRow(0);
	std::vector<uint64_t> cnt((size_t)threads * 4, 0);
	std::vector<std::thread> pool;
	for (int tid = 0; tid < threads; ++tid)
		pool.emplace_back([&, tid]
		{
			Tracer tr(P, params->max_level);
			const V4 cn = from(cam.n), cu = from(cam.u), cv = from(cam.v), cpos = from(cam.position);
			for (int y = nextRow.fetch_add(1); y < blk_h * 64; y = nextRow.fetch_add(1))
			{
				if (ownerOfRow((uint32_t)y) != rank)
					continue;
				for (int x = 0; x < blk_w * 64; ++x)
				{
					const int xcur = x - width / 2, ycur = y - height / 2;
					const V4 dir = add(add(cn, mul(cu, (float)(xcur * dp))), mul(cv, (float)(ycur * dp)));
					RayO baseray;
					baseray.origin = cpos, baseray.direction = normalize(dir), baseray.type = 1;
					Hit base;
					V4 c;
					if (type == RT_TYPE_CHECK)
					{
						// RTcheck, RayTracer.cpp:48-79: 64x64 checkerboard, no rays
						const float v = ((y / 64) & 1) == ((x / 64) & 1) ? 1.0f : 0.0f;
						c = mk(v, v, v);
					}
					else if (type < RT_TYPE_REFLECT)
						c = tr.debug(type, zNear, zFar, baseray);
					else
						c = tr.shade(zNear, zFar, baseray, 0, 1.0f, base, type != RT_TYPE_REFLECT);
					uint8_t *o = rgb + ((size_t)y * width + x) * 3;
					o[0] = put1(c.x), o[1] = put1(c.y), o[2] = put1(c.z);
					if (ids && type != RT_TYPE_CHECK)
					{
						// primary closest hit again, without counting it
						Tracer probe(P, 0);
						int64_t newobj;
						const Hit hr = probe.closest(baseray, base, newobj);
						rt_hit_id id = { -1, -1, -1, -1, hr.distance };
						if (newobj != OBJ_NONE)
						{
							if (newobj >> 40)
							{
								const uint32_t t = (uint32_t)(newobj & 0xffffffff);
								for (uint32_t m = 0; m < scene->n_models && id.object < 0; ++m)
									for (uint32_t p = 0; p < scene->models[m].part_count; ++p)
									{
										const rt_part &part = scene->parts[scene->models[m].part_begin + p];
										if (t >= part.tri_begin && t < part.tri_begin + part.tri_count)
										{
											id.object = (int32_t)scene->models[m].object, id.sub = (int32_t)p;
											id.index = (int32_t)(t - part.tri_begin), id.octant = (int32_t)((newobj >> 32) & 7);
											break;
										}
									}
							}
							else
							{
								const rt_prim &p = scene->prims[newobj];
								id.object = (int32_t)p.object, id.sub = (int32_t)p.sub;
							}
						}
						ids[(size_t)y * width + x] = id;
					}
				}
			}
			cnt[tid * 4 + 0] = tr.nPrimary, cnt[tid * 4 + 1] = tr.nShadow, cnt[tid * 4 + 2] = tr.nReflect, cnt[tid * 4 + 3] = tr.nRefract;
		});
	for (auto &t : pool) t.join();
	if (ids)
		for (int y = 0; y < height; ++y)
			for (int x = 0; x < width; ++x)
				if (y >= blk_h * 64 || x >= blk_w * 64 || ownerOfRow((uint32_t)y) != rank)
					ids[(size_t)y * width + x] = rt_hit_id{ -1, -1, -1, -1, 1e20f };
	if (counters)
	{
		memset(counters, 0, sizeof *counters);
		for (int t = 0; t < threads; ++t)
			counters->primary += cnt[t * 4], counters->shadow += cnt[t * 4 + 1], counters->reflect += cnt[t * 4 + 2], counters->refract += cnt[t * 4 + 3];
	}
	return RT_OK;
}

// DrawObject::intersect(ray, hr, min) of ONE object of the scene (3DElement.h:201), for n rays.
// Mirrors rt_intersect_object of the product ABI; used by tests/test_gpu_parity.py.
int rto_intersect_object(const rt_scene_desc *scene, uint32_t object, const rt_ray *rays, const rt_hit *in, float min, rt_hit *out, uint32_t n)
{
	if (!scene || !rays || !in || !out)
		return RT_E_INVALID;
	Prepared P;
	prepare(*scene, P);
	Tracer tr(P, 0);
	auto encode = [&](const rt_hit_id &id) -> int64_t
	{
		if (id.object < 0) return OBJ_NONE;
		for (uint32_t m = 0; m < scene->n_models; ++m)
			if ((int32_t)scene->models[m].object == id.object)
			{
				if (id.sub < 0 || (uint32_t)id.sub >= scene->models[m].part_count || id.index < 0) return OBJ_NONE;
				return tri_id(scene->parts[scene->models[m].part_begin + id.sub].tri_begin + (uint32_t)id.index, (uint32_t)(id.octant & 7));
			}
		for (uint32_t p = 0; p < scene->n_prims; ++p)
			if ((int32_t)scene->prims[p].object == id.object && (int32_t)scene->prims[p].sub == id.sub) return prim_id(p);
		return OBJ_NONE;
	};
	for (uint32_t i = 0; i < n; ++i)
	{
		RayO ray;
		ray.origin = from(rays[i].origin), ray.direction = from(rays[i].direction), ray.mtlrfr = rays[i].mtlrfr;
		ray.type = (uint8_t)rays[i].type, ray.isInside = (uint8_t)rays[i].is_inside;
		const int64_t skip = encode(in[i].id);
		Hit hr(in[i].id.distance);
		hr.obj = skip;
		int64_t newobj = OBJ_NONE;
		bool hit = false;
		for (const Item &it : P.items)
		{
			const uint32_t obj = it.kind == 1 ? scene->models[it.index].object : scene->prims[it.index].object;
			if (obj != object)
				continue;
			hr.obj = skip;   // a BallPlane compares every lattice slot with the caller's hr.obj
			const float before = hr.distance;
			hr = tr.intersect(it, ray, hr, min);
			if (hr.distance < before) newobj = hr.obj, hit = true;
		}
		out[i] = in[i];
		if (hit)
		{
			rt_hit &o = out[i];
			o.position = rt_vec4{ hr.position.x, hr.position.y, hr.position.z, 0 };
			o.normal = rt_vec4{ hr.normal.x, hr.normal.y, hr.normal.z, 0 };
			o.tu = hr.tu, o.tv = hr.tv, o.material = hr.mtl, o.texture = hr.tex;
			o.rfr = hr.rfr, o.is_inside = hr.isInside;
			rt_hit_id id = { -1, -1, -1, -1, hr.distance };
			if (newobj >> 40)
			{
				const uint32_t t = (uint32_t)(newobj & 0xffffffff);
				for (uint32_t m = 0; m < scene->n_models && id.object < 0; ++m)
					for (uint32_t p = 0; p < scene->models[m].part_count; ++p)
					{
						const rt_part &part = scene->parts[scene->models[m].part_begin + p];
						if (t >= part.tri_begin && t < part.tri_begin + part.tri_count)
						{
							id.object = (int32_t)scene->models[m].object, id.sub = (int32_t)p;
							id.index = (int32_t)(t - part.tri_begin), id.octant = (int32_t)((newobj >> 32) & 7);
							break;
						}
					}
			}
			else if (newobj >= 0)
				id.object = (int32_t)scene->prims[newobj].object, id.sub = (int32_t)scene->prims[newobj].sub;
			o.id = id;
		}
	}
	return RT_OK;
}

int rto_abi_version(void) { return RT_ABI_VERSION; }

}  // extern "C"
