// Render driver of the object model -- the drop-in for the reference's RayTracer
// (/root/reference/RayTracer.h:16-56).  Same public surface (start/stop, output, width, height,
// maxLevel, isFinish, useTime); underneath, start() flattens the Scene, hands it to the CUDA
// library through the C ABI of include/rt_b200.h and returns at once, exactly like the
// reference returns after spawning its worker threads (RayTracer.cpp:664-695).
#pragma once
#include "Scene.h"
#include "../../include/rt_b200.h"
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <thread>

#define MY_MODEL_CHECK 0x1
#define MY_MODEL_DEPTHTEST 0x2
#define MY_MODEL_NORMALTEST 0x3
#define MY_MODEL_TEXTURETEST 0x4
#define MY_MODEL_MATERIALTEST 0x5
#define MY_MODEL_SHADOWTEST 0x6
#define MY_MODEL_REFLECTTEST 0x7
#define MY_MODEL_REFRACTTEST 0x8
#define MY_MODEL_RAYTRACE 0x80

class SceneFlattener;
struct SceneResidency;

class RayTracer
{
	Scene *scene;
	// The Scene's device residency (flattened tables, triangles, BVHs) is shared by every RayTracer over
	// the same Scene on the same GPU; each RayTracer owns one frame pipeline on it (rt_create_shared),
	// so several RayTracers can have frames in flight at once without a second copy of the scene.
	std::shared_ptr<SceneResidency> residency;
	rt_ctx *ctx = nullptr;
	SceneFlattener *flattener = nullptr;
	std::thread monitor;
	rt_ctx *lastCtx = nullptr;      // the pipeline that rendered the last frame (own, or a batch worker's)
	size_t outputBytes;
	bool outputPinned = false;
	std::mutex doneMutex;           // coalesced frames: completion is signalled, not joined
	std::condition_variable doneCv;
	bool queued = false;            // a coalesced frame of this tracer is waiting or rendering
	bool countedAsCoalescer = false;
	void ensureContext();
public:
	GLuint texID = 0;
	uint8_t *output;
	volatile bool isFinish = true;
	volatile double useTime = 0.0;
	uint8_t maxLevel = 1;
	int width = 0, height = 0;
	RayTracer(Scene &scene);
	~RayTracer();

	// `tnum` was the CPU worker-thread count; the GPU path ignores it.
	void start(const uint8_t type, const int8_t tnum = 1);
	void stop();

	// ---- additions of this build (not in the reference) ----
	// The reference's start() cannot fail.  Here a frame can (device error, out of memory while a ray level is
	// regrown): the frame then ends like any other -- isFinish becomes true, `output` keeps its previous content --
	// with `failed` set and the library's message in `lastError`; the next start() clears both.
	volatile bool failed = false;
	std::string lastError;
	int device = 0;                 // CUDA device ordinal used when the context is created
	uint32_t shardRank = 0, shardWorld = 1;   // image-space shard rendered by this tracer
	uint32_t shardTileRows = 64;              // rows per shard tile (8, 16, 32, 64)
	uint32_t renderFlags = 0;       // RT_FLAG_* passed to the next start()
	uint64_t outputShardKey = 0;    // shard layout of the frame `output` holds (0 = whole frame / nothing): same layout again -> only its rows are read back
	int smShare = 0;                // resident traversal CTAs per SM of this tracer's pipeline (0 = all 8), for tracers that run concurrently
	// Throughput mode.  With `coalesce` set, start() does not launch this tracer's own pipeline: the frame is
	// queued at the Scene's device residency, whose three batch workers render whatever frames of the Scene's
	// tracers are waiting -- each through the camera its start() saw -- in ONE launch (rt_render_batch_async:
	// the frames share the ray queues) and hand every tracer its frame.  Same start()/isFinish/output
	// protocol; useTime is the time of the launch the frame was part of, readCounters() that launch's totals.
	// RAYTRACE frames without RT_FLAG_HIT_IDS only; anything else takes the tracer's own pipeline.
	bool coalesce = false;
	// Jittered supersampling (BASELINE configs[4]; not in the reference, whose frames are one sample per pixel): with
	// `samples` holding n > 1 sub-pixel offsets (dx, dy) in [0,1)^2, start() renders n frames through the camera with
	// its forward vector offset by (dx, dy) pixels -- each quantised by Color::put like a frame of its own -- and
	// `output` receives their integer mean per channel, accumulated on the GPU (rt_render_supersampled).  Identical to
	// averaging n start() calls of the reference with the same offsets.
	std::vector<std::pair<float, float>> samples;
	// Progressive display (main.cpp:207-208, 246-254: the reference's UI blits `output` every 50 ms while !isFinish and sees
	// the 64x64 tiles fill in).  With progressiveBands = k > 1 the frame is rendered as k bands of its row tiles, one after
	// the other, and every band is copied into `output` as soon as it is complete; bandsDone counts them.  Same pixels as
	// one launch; smaller launches, so it is a display mode, not a throughput mode (own pipeline, one sample per pixel).
	int progressiveBands = 0;
	volatile int bandsDone = 0;
	void completeFrame(double seconds, rt_ctx *renderedBy, const char *error = nullptr);   // called by a batch worker when this tracer's frame is in `output` (or failed / was cancelled)
	void reserveOutput(size_t bytes);          // frames beyond 2048x2048
	void wait();                               // block until isFinish
	bool readHitIds(rt_hit_id *ids);           // primary closest-hit identities of the last frame
	bool readCounters(rt_counters *out);
	rt_ctx *context() { ensureContext(); return ctx; }
	// DrawObject::intersect of Objects[index], evaluated on the device (rt_intersect_object)
	HitRes intersectObject(uint32_t index, const Ray &ray, const HitRes &hr, const float min);
	Scene *attachedScene() const { return scene; }
};
