#include "RayTracer.h"
#include "SceneUpload.h"
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

static void fail(const char *what, int code)
{
	// The B200 path has no CPU fallback: any library error is fatal and loud.
	char msg[1024];
	snprintf(msg, sizeof msg, "raytrace_b200: %s failed (%d): %s", what, code, rt_last_error());
	fprintf(stderr, "%s\n", msg);
	throw std::runtime_error(msg);
}

// `output` is page-locked when a CUDA device is present (the D2H copy of a frame then runs at PCIe
// speed); on a box without a GPU it is plain memory and start() will fail loudly anyway.
static uint8_t *alloc_output(size_t bytes, bool &pinned)
{
	void *p = nullptr;
	pinned = rt_host_alloc(&p, bytes) == RT_OK && p != nullptr;
	if (!pinned) p = new uint8_t[bytes];
	memset(p, 127, bytes);
	return (uint8_t *)p;
}

static void free_output(uint8_t *p, bool pinned)
{
	if (pinned) rt_host_free(p);
	else delete[] p;
}

RayTracer::RayTracer(Scene &scene) : scene(&scene)
{
	outputBytes = (size_t)2048 * 2048 * 3;   // the reference's fixed framebuffer, RayTracer.cpp:603
	output = new uint8_t[outputBytes];       // plain memory here (no CUDA call during static init);
	memset(output, 127, outputBytes);        // swapped for page-locked memory by the first start()
}

RayTracer::~RayTracer()
{
	if (monitor.joinable())
		monitor.join();
	if (ctx)
		rt_destroy(ctx);
	delete flattener;
	if (output) free_output(output, outputPinned);
}

void RayTracer::ensureContext()
{
	if (ctx)
		return;
	const int rc = rt_create(device, &ctx);
	if (rc != RT_OK)
		fail("rt_create", rc);
	flattener = new SceneFlattener();
}

void RayTracer::reserveOutput(size_t bytes)
{
	if (output && bytes <= outputBytes && (outputPinned || !ctx))
		return;
	if (monitor.joinable())
		monitor.join();
	if (output) free_output(output, outputPinned);
	if (bytes > outputBytes) outputBytes = bytes;
	output = alloc_output(outputBytes, outputPinned);
}

void RayTracer::start(const uint8_t type, const int8_t)
{
	if (monitor.joinable())
		monitor.join();
	isFinish = false;
	width = scene->cam.width;
	height = scene->cam.height;
	ensureContext();
	reserveOutput((size_t)width * height * 3);
	for (DrawObject *o : scene->Objects)
		if (o->bShow)
			o->RTPrepare();

	rt_scene_desc desc;
	flattener->flatten(*scene, desc);
	int rc = rt_upload_scene(ctx, &desc);
	if (rc != RT_OK)
		fail("rt_upload_scene", rc);
	rt_render_params rp;
	memset(&rp, 0, sizeof rp);
	rp.type = type, rp.max_level = maxLevel;
	rp.rank = shardRank, rp.world = shardWorld, rp.flags = renderFlags, rp.tile_rows = shardTileRows;
	rc = rt_render_async(ctx, &rp);
	if (rc != RT_OK)
		fail("rt_render_async", rc);

	// the monitor thread of RayTracer.cpp:674-695: waits for the frame, publishes it
	monitor = std::thread([this]
	{
		double seconds = 0.0;
		int rc = rt_wait(ctx, &seconds);
		if (rc == RT_OK)
			rc = rt_read_output(ctx, output, (size_t)width * 3);
		if (rc != RT_OK)
		{
			fprintf(stderr, "raytrace_b200: frame failed (%d): %s\n", rc, rt_last_error());
			abort();
		}
		useTime = seconds;
		isFinish = true;
	});
}

void RayTracer::stop()
{
	if (ctx)
		rt_stop(ctx);
}

void RayTracer::wait()
{
	if (monitor.joinable())
		monitor.join();
}

bool RayTracer::readHitIds(rt_hit_id *ids)
{
	wait();
	return ctx && rt_read_hit_ids(ctx, ids) == RT_OK;
}

bool RayTracer::readCounters(rt_counters *out)
{
	wait();
	return ctx && rt_read_counters(ctx, out) == RT_OK;
}

// DrawObject::intersect on the host: the operator lives on the GPU; a host-side caller gets a
// clear failure instead of a silent CPU path.
HitRes DrawObject::intersect(const Ray &, const HitRes &, const float)
{
	throw std::runtime_error("raytrace_b200: DrawObject::intersect runs on the GPU only (see include/rt_b200.h); "
		"there is no CPU implementation in this build");
}
