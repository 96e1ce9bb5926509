// Product arm of tools/render_main.cpp: hit ids and ray counters come from the CUDA library.
#pragma once
#include "SceneUpload.h"
#include <string>

namespace rt_taps
{
struct HitId { int32_t obj, sub, idx, oct; float t; };
struct Counts { unsigned long long primary = 0, shadow = 0, reflect = 0, refract = 0; };
static int g_gpus = 1;

inline const char *arm() { return "b200"; }
inline void set_gpus(int n) { g_gpus = n; }
inline void ensure_output(RayTracer &rt, size_t bytes) { rt.reserveOutput(bytes); }

inline void primary_ids(Scene &, RayTracer &rt, int width, int height, HitId *out)
{
	static_assert(sizeof(HitId) == sizeof(rt_hit_id), "layout");
	rt.renderFlags |= RT_FLAG_HIT_IDS;
	rt.start(MY_MODEL_RAYTRACE, 1);
	rt.wait();
	rt.renderFlags &= ~RT_FLAG_HIT_IDS;
	if (!rt.readHitIds((rt_hit_id *)out)) { fprintf(stderr, "rt_read_hit_ids: %s\n", rt_last_error()); exit(3); }
	(void)width, (void)height;
}

inline Counts count_rays(Scene &, RayTracer &rt, int, int)
{
	rt_counters c;
	if (!rt.readCounters(&c)) { fprintf(stderr, "rt_read_counters: %s\n", rt_last_error()); exit(3); }
	Counts o;
	o.primary = c.primary, o.shadow = c.shadow, o.reflect = c.reflect, o.refract = c.refract;
	return o;
}

inline long render_tiles(Scene &, RayTracer &, int, int, int, int, int, int, Counts *, double * = nullptr, double * = nullptr)
{
	fprintf(stderr, "--tiles is a CPU-baseline sampling mode of the reference arm only\n");
	exit(2);
}

inline std::string extra_json(RayTracer &rt)
{
	rt_counters c;
	if (!rt.readCounters(&c)) return "";
	char buf[512];
	snprintf(buf, sizeof buf, ",\"render_ms\":%.4f,\"upload_ms\":%.3f,\"build_ms\":%.3f,\"launches\":%u,\"bvh_nodes\":%u,\"bvh_depth\":%u",
		c.render_ms, c.upload_ms, c.build_ms, c.launches, c.bvh_nodes, c.bvh_depth);
	return buf;
}
}  // namespace rt_taps
