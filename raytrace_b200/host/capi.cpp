// C shim over the C++ object model so that Python (ctypes) tests and bench.py can drive the
// SAME classes a C++ user of the reference API would: Scene, the scene builders and RayTracer.
// Prefix rth_ ("ray tracer host").  No per-ray work happens here.
#include "RayTracer.h"
#include "SceneUpload.h"
#include "../scenes/scenes.h"
#include <cstdio>
#include <stdexcept>
#include <string>

namespace
{
struct HostScene
{
	Scene scene;
	SceneFlattener flattener;
	rt_scene_desc desc;
};
thread_local std::string g_err;
template<class F> int guarded(F &&f)
{
	try { f(); return 0; }
	catch (const std::exception &e) { g_err = e.what(); return -1; }
}
}

extern "C" {

const char *rth_last_error(void) { return g_err.c_str(); }

// FNV-1a-64, the golden-hash convention of BASELINE.md (offset 1469598103934665603, prime 1099511628211)
unsigned long long rth_fnv1a64(const uint8_t *p, size_t n)
{
	unsigned long long h = 1469598103934665603ULL;
	for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
	return h;
}

void *rth_scene_new(void) { return new HostScene(); }
void rth_scene_free(void *h) { delete (HostScene *)h; }

// named synthetic scenes of raytrace_b200/scenes/scenes.h ("c1".."c5", "t_*")
int rth_scene_build(void *h, const char *name, int n, int parts, int width, int height, const char *tmpdir)
{
	return guarded([&]
	{
		HostScene *hs = (HostScene *)h;
		rtscenes::SceneArgs a;
		a.name = name, a.n = n, a.parts = parts;
		if (tmpdir && *tmpdir) a.tmpdir = tmpdir;
		hs->scene.cam.resize(width, height);
		if (!rtscenes::build(hs->scene, a))
			throw std::runtime_error(std::string("unknown scene ") + name);
	});
}

int rth_scene_resize(void *h, int width, int height) { ((HostScene *)h)->scene.cam.resize(width, height); return 0; }
int rth_scene_object_count(void *h) { return (int)((HostScene *)h)->scene.Objects.size(); }
int rth_scene_light_count(void *h) { return (int)((HostScene *)h)->scene.Lights.size(); }

// Scene::MovePos / Switch / ChgMtl(library material) / camera edits, for incremental-upload tests
int rth_scene_move(void *h, int type, int num, float x, float y, float z)
{
	return ((HostScene *)h)->scene.MovePos((uint8_t)type, (uint8_t)num, Vertex(x, y, z)) ? 0 : -1;
}
int rth_scene_switch(void *h, int type, int num, int show)
{
	((HostScene *)h)->scene.Switch((uint8_t)type, (uint8_t)num, show != 0);
	return 0;
}
int rth_scene_chgmtl(void *h, int num, int libIndex)
{
	Scene &s = ((HostScene *)h)->scene;
	if (libIndex < 0 || libIndex >= (int)s.MtlLiby.size()) return -1;
	return s.ChgMtl((uint8_t)num, s.MtlLiby[libIndex]) ? 0 : -1;
}
int rth_scene_set_object_position(void *h, int num, float x, float y, float z)
{
	Scene &s = ((HostScene *)h)->scene;
	if (num < 0 || num >= (int)s.Objects.size()) return -1;
	s.Objects[num]->position = Vertex(x, y, z);
	return 0;
}
int rth_scene_set_light_position(void *h, int num, float x, float y, float z, float w)
{
	Scene &s = ((HostScene *)h)->scene;
	if (num < 0 || num >= (int)s.Lights.size()) return -1;
	s.Lights[num].position = Vertex(x, y, z, w);
	return 0;
}
int rth_scene_camera_move(void *h, float x, float y, float z) { ((HostScene *)h)->scene.cam.move(x, y, z); return 0; }
int rth_scene_camera_yaw(void *h, float a) { ((HostScene *)h)->scene.cam.yaw(a); return 0; }
int rth_scene_camera_pitch(void *h, float a) { ((HostScene *)h)->scene.cam.pitch(a); return 0; }
// sub-pixel jitter as the harness of SURVEY.md 8c does it: n' = n + u*(dx*dp) + v*(dy*dp)
int rth_scene_camera_jitter(void *h, float dx, float dy)
{
	Camera &c = ((HostScene *)h)->scene.cam;
	const double dp = tan(c.fovy * PI / 360) / (c.height / 2);
	const Vertex n = c.n + c.u * (float)(dx * dp) + c.v * (float)(dy * dp);
	c.n.x = n.x, c.n.y = n.y, c.n.z = n.z, c.n.w = n.w;   // deliberately NOT re-normalised
	return 0;
}

// camera k of the K-camera orbit around the scene's CURRENT camera (scenes.h orbit_camera), as the C-ABI
// record rt_render_batch_async takes; the Scene itself is not changed
int rth_scene_orbit_camera(void *h, int k, int K, rt_camera *out)
{
	const Camera c = rtscenes::orbit_camera(((HostScene *)h)->scene.cam, k, K);
	SceneFlattener::cameraRecord(c, *out);
	return 0;
}
// ... and the same camera put INTO the scene (start() then renders through it); `base` is restored by k = 0 of
// a second call only if the caller kept it: callers pass the base camera they read with rth_scene_camera_get
int rth_scene_camera_get(void *h, rt_camera *out) { SceneFlattener::cameraRecord(((HostScene *)h)->scene.cam, *out); return 0; }
int rth_scene_camera_set_position(void *h, float x, float y, float z)
{
	Camera &c = ((HostScene *)h)->scene.cam;
	c.position.x = x, c.position.y = y, c.position.z = z;
	return 0;
}

// the jittered sample camera of (dx, dy) as a C-ABI record; the Scene is not changed
int rth_scene_jittered_camera(void *h, float dx, float dy, rt_camera *out)
{
	SceneFlattener::cameraRecord(rtscenes::jittered_camera(((HostScene *)h)->scene.cam, dx, dy), *out);
	return 0;
}
// side*side (dx, dy) pairs of the fixed stratified table (scenes.h stratified_table)
int rth_stratified_table(int side, int seed, float *out)
{
	std::vector<float> t;
	rtscenes::stratified_table(side, seed, t);
	for (size_t i = 0; i < t.size(); ++i) out[i] = t[i];
	return (int)t.size() / 2;
}

int rth_scene_camera_get_n(void *h, float *xyzw)
{
	const Camera &c = ((HostScene *)h)->scene.cam;
	xyzw[0] = c.n.x, xyzw[1] = c.n.y, xyzw[2] = c.n.z, xyzw[3] = c.n.w;
	return 0;
}
int rth_scene_camera_set_n(void *h, const float *xyzw)
{
	Camera &c = ((HostScene *)h)->scene.cam;
	c.n.x = xyzw[0], c.n.y = xyzw[1], c.n.z = xyzw[2], c.n.w = xyzw[3];
	return 0;
}

// flatten to the C-ABI description; the pointer stays valid until the next flatten/free
const rt_scene_desc *rth_scene_flatten(void *h)
{
	HostScene *hs = (HostScene *)h;
	for (DrawObject *o : hs->scene.Objects)
		if (o->bShow) o->RTPrepare();
	if (guarded([&] { hs->flattener.flatten(hs->scene, hs->desc); }) != 0)
		return nullptr;
	return &hs->desc;
}

// RayTracer (the drop-in surface)
void *rth_tracer_new(void *scene, int device)
{
	RayTracer *t = new RayTracer(((HostScene *)scene)->scene);
	t->device = device;
	return t;
}
void rth_tracer_free(void *t) { delete (RayTracer *)t; }
int rth_tracer_start(void *t, int type, int tnum) { return guarded([&] { ((RayTracer *)t)->start((uint8_t)type, (int8_t)tnum); }); }
void rth_tracer_stop(void *t) { ((RayTracer *)t)->stop(); }
int rth_tracer_failed(void *t) { return ((RayTracer *)t)->failed ? 1 : 0; }
const char *rth_tracer_last_error(void *t) { return ((RayTracer *)t)->lastError.c_str(); }
int rth_tracer_is_finished(void *t) { return ((RayTracer *)t)->isFinish ? 1 : 0; }
void rth_tracer_wait(void *t) { ((RayTracer *)t)->wait(); }
double rth_tracer_use_time(void *t) { return ((RayTracer *)t)->useTime; }
const uint8_t *rth_tracer_output(void *t) { return ((RayTracer *)t)->output; }
int rth_tracer_width(void *t) { return ((RayTracer *)t)->width; }
int rth_tracer_height(void *t) { return ((RayTracer *)t)->height; }
void rth_tracer_set_max_level(void *t, int level) { ((RayTracer *)t)->maxLevel = (uint8_t)level; }
void rth_tracer_set_shard(void *t, int rank, int world, int tileRows)
{
	RayTracer *r = (RayTracer *)t;
	r->shardRank = rank, r->shardWorld = world, r->shardTileRows = tileRows > 0 ? tileRows : 64;
}
void rth_tracer_set_flags(void *t, unsigned flags) { ((RayTracer *)t)->renderFlags = flags; }
// RayTracer::samples: n (dx, dy) sub-pixel offsets; n = 0 or 1 sample at (0, 0) = the plain frame
int rth_tracer_set_samples(void *t, int n, const float *dxdy)
{
	RayTracer *r = (RayTracer *)t;
	r->samples.clear();
	for (int i = 0; i < n; ++i) r->samples.push_back(std::make_pair(dxdy[2 * i], dxdy[2 * i + 1]));
	return 0;
}
void rth_tracer_set_progressive(void *t, int bands) { ((RayTracer *)t)->progressiveBands = bands; }
int rth_tracer_bands_done(void *t) { return ((RayTracer *)t)->bandsDone; }
void rth_tracer_set_coalesce(void *t, int on) { ((RayTracer *)t)->coalesce = on != 0; }
void rth_tracer_set_sm_share(void *t, int ctasPerSm) { ((RayTracer *)t)->smShare = ctasPerSm; }
int rth_tracer_read_hit_ids(void *t, rt_hit_id *ids) { return ((RayTracer *)t)->readHitIds(ids) ? 0 : -1; }
int rth_tracer_read_counters(void *t, rt_counters *c) { return ((RayTracer *)t)->readCounters(c) ? 0 : -1; }
// B2 through the C++ classes: scene.Objects[index]->intersect(ray, hr, min).  `hit` is in/out: on
// entry hit->id / hit->id.distance describe hr (obj + distance), on exit the returned HitRes.
int rth_object_intersect(void *scene, int index, const rt_ray *ray, rt_hit *hit, float min)
{
	return guarded([&]
	{
		Scene &s = ((HostScene *)scene)->scene;
		if (index < 0 || index >= (int)s.Objects.size()) throw std::runtime_error("object index out of range");
		Ray r(Vertex(ray->origin.x, ray->origin.y, ray->origin.z, ray->origin.w),
			Normal(ray->direction.x, ray->direction.y, ray->direction.z, ray->direction.w), (uint8_t)ray->type);
		r.mtlrfr = ray->mtlrfr, r.isInside = (uint8_t)ray->is_inside;
		HitRes hr(hit->id.distance);
		// hr.obj: rebuild the pointer-style identity the classes use
		if (hit->id.object >= 0 && hit->id.object < (int)s.Objects.size())
		{
			DrawObject *o = s.Objects[hit->id.object];
			hr.obj = o->type == MY_OBJECT_MODEL
				? ((intptr_t)1 << 62) | ((intptr_t)hit->id.object << 33) | ((intptr_t)(hit->id.octant & 7) << 30) | ((intptr_t)(hit->id.sub & 0x7FFF) << 15) | (intptr_t)(hit->id.index & 0x7FFF)
				: (intptr_t)o + hit->id.sub;
		}
		const HitRes res = s.Objects[index]->intersect(r, hr, min);
		hit->id.distance = res.distance;
		if (res.distance < hr.distance)
		{
			hit->position = rt_vec4{ res.position.x, res.position.y, res.position.z, 0 };
			hit->normal = rt_vec4{ res.normal.x, res.normal.y, res.normal.z, 0 };
			hit->tu = res.tcoord.u, hit->tv = res.tcoord.v, hit->rfr = res.rfr, hit->is_inside = res.isInside;
			hit->material = res.mtl ? 1 : -1, hit->texture = res.tex ? 1 : -1;   // presence only: pointers do not cross the shim
			hit->id.object = -2;   // identity is pointer-valued on this path; tests compare geometry
		}
	});
}

void *rth_tracer_context(void *t)
{
	void *c = nullptr;
	guarded([&] { c = ((RayTracer *)t)->context(); });
	return c;
}

}  // extern "C"
