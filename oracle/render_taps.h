// TEST INFRASTRUCTURE (oracle side) -- reference arm of tools/render_main.cpp.
//
// Compiled only by oracle/build_ref.sh, against the reference's own headers under
// /root/reference with -fno-access-control, into oracle/_ref/ref_render.  Nothing in the product
// (raytrace_b200/) includes this file.  It taps the UNMODIFIED reference for
//   * primary closest-hit identities  (the loop of RayTracer.cpp:455-465 through the public
//     virtual DrawObject::intersect, 3DElement.h:201),
//   * ray counts (a pass-through counting DrawObject at Objects[0]; one count per query because
//     both loops RayTracer.cpp:458 and :513 visit Objects[0] first),
//   * bounded tile samples for CPU baselines through the private per-pixel entry
//     RayTracer::RTfrac (RayTracer.cpp:450), reachable because of -fno-access-control.
#pragma once
#include <atomic>
#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

namespace rt_taps
{

struct HitId { int32_t obj, sub, idx, oct; float t; };
struct Counts { unsigned long long primary = 0, shadow = 0, reflect = 0, refract = 0; };

inline const char *arm() { return "reference"; }
inline void set_gpus(int) {}
inline std::string extra_json(RayTracer &) { return ""; }

inline void ensure_output(RayTracer &rt, size_t bytes)
{
	// the reference owns `output` with new[]/delete[] (RayTracer.cpp:603,611): hand it a bigger one
	delete[] rt.output;
	rt.output = new uint8_t[bytes];
	memset(rt.output, 127, bytes);
}

class CountingProxy : public DrawObject
{
public:
	std::atomic<unsigned long long> n[8];
	CountingProxy() : DrawObject(0) { type = 0; for (auto &c : n) c = 0; }
	void GLPrepare() override {}
	HitRes intersect(const Ray &ray, const HitRes &hr, const float = 0) override
	{
		n[ray.type & 7].fetch_add(1, std::memory_order_relaxed);
		return hr;
	}
};

inline Counts count_rays(Scene &scene, RayTracer &rt, int type, int threads)
{
	CountingProxy *proxy = new CountingProxy();
	scene.Objects.insert(scene.Objects.begin(), proxy);
	rt.start((uint8_t)type, (int8_t)threads);
	while (!rt.isFinish) std::this_thread::sleep_for(std::chrono::milliseconds(1));
	scene.Objects.erase(scene.Objects.begin());
	Counts c;
	c.primary = proxy->n[MY_RAY_BASERAY];
	c.shadow = proxy->n[MY_RAY_SHADOWRAY] + proxy->n[0];   // RTshd/RTflec use untyped rays
	c.reflect = proxy->n[MY_RAY_REFLECTRAY];
	c.refract = proxy->n[MY_RAY_REFRACTRAY];
	delete proxy;
	return c;
}

struct ModelRange { intptr_t lo, hi; int obj, pcur; };

// Map HitRes::obj (an address) back to (object index, sub id, triangle index, octant).
inline void primary_ids(Scene &scene, RayTracer &rt, int width, int height, HitId *out)
{
	for (auto dobj : scene.Objects)
		if (dobj->bShow) dobj->RTPrepare();
	std::vector<ModelRange> ranges;
	for (size_t i = 0; i < scene.Objects.size(); ++i)
		if (scene.Objects[i]->type == MY_OBJECT_MODEL)
		{
			Model &m = dynamic_cast<Model &>(*scene.Objects[i]);
			for (size_t p = 0; p < m.octclparts.size(); ++p)
				if (!m.octclparts[p].empty())
					ranges.push_back({ (intptr_t)&m.octclparts[p].front(), (intptr_t)(&m.octclparts[p].back() + 1), (int)i, (int)p });
		}
	std::sort(ranges.begin(), ranges.end(), [](const ModelRange &a, const ModelRange &b) { return a.lo < b.lo; });

	const Camera &cam = scene.cam;
	const int blk_h = height / 64, blk_w = width / 64;
	const double dp = tan(cam.fovy * PI / 360) / (height / 2);
	for (size_t i = 0; i < (size_t)width * height; ++i) out[i] = { -1, -1, -1, -1, 1e20f };
	const int nthr = 8;
	std::vector<std::thread> pool;
	for (int tid = 0; tid < nthr; ++tid)
		pool.emplace_back([&, tid]
		{
			for (int y = tid; y < blk_h * 64; y += nthr)
				for (int x = 0; x < blk_w * 64; ++x)
				{
					const int xcur = x - width / 2, ycur = y - height / 2;
					Vertex dir = cam.n + cam.u*(xcur*dp) + cam.v*(ycur*dp);
					Ray baseray(cam.position, dir, MY_RAY_BASERAY);
					HitRes basehr;
					HitRes hr = basehr;
					intptr_t newobj = basehr.obj;
					for (auto dobj : scene.Objects)
						if (dobj->bShow)
						{
							hr.obj = basehr.obj;
							hr = dobj->intersect(baseray, hr);
							if (hr.obj != basehr.obj)
								newobj = hr.obj;
						}
					HitId id = { -1, -1, -1, -1, hr.distance };
					hr.obj = newobj;
					if (hr.obj != basehr.obj)
					{
						for (size_t o = 0; o < scene.Objects.size() && id.obj < 0; ++o)
						{
							DrawObject *d = scene.Objects[o];
							if (hr.obj == (intptr_t)d) id = { (int)o, 0, -1, -1, hr.distance };
							else if (d->type == MY_OBJECT_BALLPLANE && hr.obj > (intptr_t)d && hr.obj <= (intptr_t)d + 16)
								id = { (int)o, (int)(hr.obj - (intptr_t)d), -1, -1, hr.distance };
						}
						if (id.obj < 0)
						{
							auto it = std::upper_bound(ranges.begin(), ranges.end(), hr.obj,
								[](intptr_t v, const ModelRange &r) { return v < r.lo; });
							if (it != ranges.begin())
							{
								--it;
								if (hr.obj >= it->lo && hr.obj < it->hi)
								{
									const clTri *t = (const clTri *)hr.obj;
									id = { it->obj, t->numa, t->numb, it->pcur % 8, hr.distance };
								}
							}
						}
					}
					out[(size_t)y * width + x] = id;
				}
		});
	for (auto &t : pool) t.join();
}

// Bounded CPU sample: `tiles` 64x64 tiles traced pixel by pixel through RTfrac/RTflec on `threads` threads.
//   seed >= 0: seeded random tiles;  seed < 0: STRATIFIED -- tile k of the sample is tile floor((k + 0.5) * nblk / tiles)
//   of the frame in row-major tile order, so every sample covers all tile rows evenly (sky rows are cheap, mesh
//   rows expensive: a handful of random tiles is not ray-count representative).
// Work is handed out per 64-pixel ROW of a tile (one atomic per row), so min(threads, 64 * tiles) threads work --
// the reference's own parallelRT hands out whole tiles (RayTracer.cpp:28-33), which is fine for a full frame of
// hundreds of tiles but leaves most threads idle on a small sample.
// RTPrepare (single-threaded in the reference, once per start(), RayTracer.cpp:622-627) runs first and is timed
// separately: *prepare_s.  A caller that extrapolates to a frame charges it ONCE per frame, i.e.
// frame_s ~= prepare_s + trace_s * (frame tiles / sample tiles).
inline long render_tiles(Scene &scene, RayTracer &rt, int width, int height, int tiles, int seed, int threads, int type, Counts *counts,
	double *prepare_s = nullptr, double *trace_s = nullptr)
{
	CountingProxy *proxy = nullptr;
	if (counts)
	{
		proxy = new CountingProxy();
		scene.Objects.insert(scene.Objects.begin(), proxy);
	}
	rt.width = width, rt.height = height;
	const auto t0 = std::chrono::steady_clock::now();
	for (auto dobj : scene.Objects)
		if (dobj->bShow) dobj->RTPrepare();
	const auto t1 = std::chrono::steady_clock::now();
	const Camera &cam = scene.cam;
	const int blk_h = height / 64, blk_w = width / 64, nblk = blk_h * blk_w;
	if (tiles > nblk) tiles = nblk;
	std::vector<int> order(nblk);
	for (int i = 0; i < nblk; ++i) order[i] = i;
	if (seed >= 0)
	{
		unsigned long long s = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)seed;
		for (int i = nblk - 1; i > 0; --i)
		{
			s = s * 6364136223846793005ULL + 1442695040888963407ULL;
			std::swap(order[i], order[(int)((s >> 33) % (unsigned)(i + 1))]);
		}
	}
	else
		for (int k = 0; k < tiles; ++k) order[k] = (int)(((long long)(2 * k + 1) * nblk) / (2 * tiles));
	const double dp = tan(cam.fovy * PI / 360) / (height / 2);
	const float zNear = cam.zNear, zFar = sqrt(2)*cam.zFar;
	std::atomic<int> next(0);
	std::vector<std::thread> pool;
	for (int tid = 0; tid < threads; ++tid)
		pool.emplace_back([&]
		{
			HitRes base;
			for (int k = next.fetch_add(1); k < tiles * 64; k = next.fetch_add(1))
			{
				const int tile = order[k >> 6], bx = tile % blk_w, by = tile / blk_w;
				const int y = by * 64 + (k & 63);
				for (int x = bx * 64; x < bx * 64 + 64; ++x)
				{
					const int xcur = x - width / 2, ycur = y - height / 2;
					Vertex dir = cam.n + cam.u*(xcur*dp) + cam.v*(ycur*dp);
					Ray baseray(cam.position, dir, MY_RAY_BASERAY);
					Color c = (type == MY_MODEL_REFLECTTEST) ? rt.RTflec(zNear, zFar, baseray, 0, 1.0f, base)
						: rt.RTfrac(zNear, zFar, baseray, 0, 1.0f, base);
					c.put(rt.output + ((size_t)y * width + x) * 3);
				}
			}
		});
	for (auto &t : pool) t.join();
	const auto t2 = std::chrono::steady_clock::now();
	if (prepare_s) *prepare_s = std::chrono::duration<double>(t1 - t0).count();
	if (trace_s) *trace_s = std::chrono::duration<double>(t2 - t1).count();
	if (proxy)
	{
		scene.Objects.erase(scene.Objects.begin());
		counts->primary = proxy->n[MY_RAY_BASERAY], counts->shadow = proxy->n[MY_RAY_SHADOWRAY] + proxy->n[0];
		counts->reflect = proxy->n[MY_RAY_REFLECTRAY], counts->refract = proxy->n[MY_RAY_REFRACTRAY];
		delete proxy;
	}
	return (long)tiles * 64 * 64;
}

}  // namespace rt_taps
