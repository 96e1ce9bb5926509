#include "RayTracer.h"
#include "SceneUpload.h"
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <vector>
#include <stdexcept>

// sample camera of a sub-pixel offset: n' = n + u*(dx*dp) + v*(dy*dp), not re-normalised (scenes.h jittered_camera restates
// it for both arms; RayTracer.cpp:20-27 is where the reference adds the same multiples of u and v per pixel)
static Camera rtscenes_jittered(const Camera &base, float dx, float dy)
{
	Camera c = base;
	const double dp = tan(c.fovy * PI / 360) / (c.height / 2);
	const Vertex n = c.n + c.u * (float)(dx * dp) + c.v * (float)(dy * dp);
	c.n.x = n.x, c.n.y = n.y, c.n.z = n.z, c.n.w = n.w;
	return c;
}

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

// A frame a coalescing tracer asked for: rendered by a batch worker together with whatever else is waiting.
struct FrameRequest
{
	RayTracer *tracer;
	rt_camera camera;        // the Scene's camera as start() saw it
	rt_render_params rp;
	bool rowsOnly;
};

// One per (Scene, device): the context that holds the uploaded scene, and the flattener that feeds it.
struct SceneResidency
{
	rt_ctx *parent = nullptr;
	SceneFlattener flattener;
	std::mutex mutex;   // flatten + upload of concurrent start() calls

	// ---- batch workers (RayTracer::coalesce): two pipelines on the resident scene, each renders the frames
	// that are waiting when it becomes free in one launch, so two launches are in flight ----
	static constexpr int kWorkers = 8;   // at most; nWorkers are started
	static constexpr size_t kMaxBatch = 64;
	std::mutex qMutex;
	std::condition_variable qCv;
	std::deque<FrameRequest> pending;
	std::thread workers[kWorkers];
	rt_ctx *workerCtx[kWorkers] = {};
	// RT_B200_COALESCE_WORKERS: batch pipelines (launches that can be in flight or delivering), default 3;
	// RT_B200_COALESCE_DIV: a launch takes up to (coalescing tracers / div) frames, default workers + 1, so that a
	// quarter of the tracers is always waiting: a worker that has just delivered finds its next launch ready
	// instead of waiting for its own tracers to be restarted.  With div = workers every tracer is in flight at
	// once, the workers fall into step and all read back at the same time with nothing rendering (C3, one
	// GPU, 4 frames per launch: 3 520 Mrays/s against 4 191 with 16 tracers, 3 workers, div 4).
	int nWorkers = envInt("RT_B200_COALESCE_WORKERS", 3, 1, kWorkers);
	int wantDiv = envInt("RT_B200_COALESCE_DIV", 0, 0, 64);
	static int envInt(const char *name, int dflt, int lo, int hi)
	{
		const char *e = getenv(name);
		const int v = e ? atoi(e) : dflt;
		return v < lo ? lo : v > hi ? hi : v;
	}
	bool workersUp = false, quit = false;
	int coalescers = 0;       // tracers of this Scene in throughput mode (sizes the coalescing window)
	size_t reserved[8] = {};  // per worker: the batch size its pipeline's buffers were announced for (rt_reserve_batch)
	std::atomic<int> launchesInFlight{0};   // batch launches between rt_render_batch_async and the end of rt_wait

	static bool sameLaunch(const FrameRequest &a, const FrameRequest &b)
	{
		return a.rp.type == b.rp.type && a.rp.max_level == b.rp.max_level && a.rp.rank == b.rp.rank && a.rp.world == b.rp.world
			&& a.rp.flags == b.rp.flags && a.rp.tile_rows == b.rp.tile_rows && a.camera.width == b.camera.width && a.camera.height == b.camera.height
			&& a.camera.fovy == b.camera.fovy && a.camera.zNear == b.camera.zNear && a.camera.zFar == b.camera.zFar;
	}

	static double nowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

	void workerLoop(int w)
	{
		std::vector<FrameRequest> batch;
		std::vector<rt_camera> cams;
		const bool trace = getenv("RT_B200_COALESCE_TRACE") != nullptr;
		double tFirst = 0, tLaunch = 0, tRendered = 0;
		size_t wantLast = 0;
		while (true)
		{
			batch.clear();
			{
				std::unique_lock<std::mutex> lock(qMutex);
				qCv.wait(lock, [this] { return quit || !pending.empty(); });
				if (pending.empty())
					return;   // quit
				// Coalescing window: a caller that keeps T tracers in flight restarts them one after the other,
				// a few tens of microseconds apart; a worker that took the first request at once would launch
				// batches of one.  Wait while requests keep arriving, up to this worker's share of the tracers.
				const int div = wantDiv > 0 ? wantDiv : nWorkers + 1;
				const size_t want = std::max<size_t>(1, std::min(kMaxBatch, (size_t)(coalescers + div - 1) / div));
				wantLast = want, tFirst = nowS();
				// While another worker's launch keeps the GPU busy there is no hurry: a worker that took what was there
				// after one quiet window launched batches of one or two frames, every such launch pays the fixed tail of
				// a launch, its tracers come back early, are restarted alone, ... -- a convoy of small batches that the
				// bench fell into on every other run at two GPUs (2 063 against 8 120 Mrays/s).  The quiet-window rule only
				// applies when the GPU would otherwise idle; the wait for a full batch is bounded (a caller may be blocked
				// in wait() on a tracer whose request sits here).
				const double tCollect = nowS();
				while (!quit && pending.size() < want)
				{
					const size_t before = pending.size();
					qCv.wait_for(lock, std::chrono::microseconds(250));
					if (pending.size() == before && (launchesInFlight.load() == 0 || nowS() - tCollect > 4e-3))
						break;   // nothing new within the window and nothing rendering (or waited long enough)
				}
				if (pending.empty())
					continue;   // the other worker took them
				batch.push_back(pending.front());
				pending.pop_front();
				// everything behind it that can share the launch, in arrival order
				for (auto it = pending.begin(); it != pending.end() && batch.size() < want;)
					if (sameLaunch(batch[0], *it)) { batch.push_back(*it); it = pending.erase(it); }
					else ++it;
			}
			cams.clear();
			for (const FrameRequest &r : batch) cams.push_back(r.camera);
			rt_ctx *c = workerCtx[w];
			double seconds = 0.0;
			int rc;
			tLaunch = nowS();
			if (wantLast > reserved[w])
			{
				// the queues of this worker's pipeline are sized for its full share once, not re-allocated as the batches grow
				rt_reserve_batch(c, (uint32_t)std::min(wantLast, kMaxBatch));
				reserved[w] = wantLast;
			}
			++launchesInFlight;
			{
				// the launch adopts the parent's scene tables: not while a start() is uploading into them
				std::lock_guard<std::mutex> lock(mutex);
				rc = rt_render_batch_async(c, &batch[0].rp, (uint32_t)batch.size(), cams.data(), nullptr);
			}
			if (rc == RT_OK) rc = rt_wait(c, &seconds);
			--launchesInFlight;
			qCv.notify_all();   // a worker that is collecting re-evaluates "is anything rendering"
			tRendered = nowS();
			// all copies enqueued back to back, one wait (the last call completes them all)
			for (size_t f = 0; f < batch.size() && rc == RT_OK; ++f)
				rc = rt_read_batch_output(c, (uint32_t)f, batch[f].tracer->output, (size_t)batch[f].camera.width * 3,
					(batch[f].rowsOnly ? 1 : 0) | (f + 1 < batch.size() ? 2 : 0));
			if (trace)
				fprintf(stderr, "raytrace_b200 coalesce[w%d]: %zu frames (want %zu), waited %.3f ms for the batch, render+wait %.3f ms (device %.3f ms), read-back %.3f ms\n",
					w, batch.size(), wantLast, (tLaunch - tFirst) * 1e3, (tRendered - tLaunch) * 1e3, seconds * 1e3, (nowS() - tRendered) * 1e3);
			if (rc != RT_OK)
			{
				// no abort: every tracer of the batch gets its frame back as failed (isFinish, failed, lastError)
				const std::string err = rt_last_error();
				fprintf(stderr, "raytrace_b200: batch of %zu frames failed (%d): %s\n", batch.size(), rc, err.c_str());
				for (const FrameRequest &r : batch)
					r.tracer->completeFrame(seconds, c, err.c_str());
				continue;
			}
			for (const FrameRequest &r : batch)
				r.tracer->completeFrame(seconds, c);
		}
	}

	// -> the worker pipelines exist (created on first use)
	int ensureWorkers()
	{
		std::lock_guard<std::mutex> lock(qMutex);
		if (workersUp) return RT_OK;
		for (int w = 0; w < nWorkers; ++w)
		{
			const int rc = rt_create_shared(parent, &workerCtx[w]);
			if (rc != RT_OK) return rc;
		}
		for (int w = 0; w < nWorkers; ++w)
			workers[w] = std::thread([this, w] { workerLoop(w); });
		workersUp = true;
		return RT_OK;
	}

	void enqueue(const FrameRequest &r)
	{
		{
			std::lock_guard<std::mutex> lock(qMutex);
			pending.push_back(r);
		}
		qCv.notify_all();
	}
	// RayTracer::stop of a coalescing tracer: a request that is still waiting is taken out of the queue (-> true, the
	// caller completes it as cancelled); one that is already part of a launch is left alone -- the launch holds other
	// tracers' frames too, so it is never stopped, the frame is simply delivered (a cooperative cancel may finish)
	bool cancelPending(RayTracer *t)
	{
		std::lock_guard<std::mutex> lock(qMutex);
		for (auto it = pending.begin(); it != pending.end(); ++it)
			if (it->tracer == t) { pending.erase(it); return true; }
		return false;
	}
	void addCoalescer(int d)
	{
		std::lock_guard<std::mutex> lock(qMutex);
		coalescers += d;
	}

	rt_ctx *workerContext(int w) const { return workerCtx[w]; }

	~SceneResidency()
	{
		{
			std::lock_guard<std::mutex> lock(qMutex);
			quit = true;
		}
		qCv.notify_all();
		for (auto &t : workers) if (t.joinable()) t.join();
		for (rt_ctx *c : workerCtx) if (c) rt_destroy(c);
		if (parent) rt_destroy(parent);
	}
};
static std::mutex g_residencyMutex;
static std::map<std::pair<const Scene *, int>, std::weak_ptr<SceneResidency>> g_residency;

// Every live RayTracer, so that DrawObject::intersect (which has no back pointer) can find the
// device context its Scene is attached to.
static std::mutex g_tracersMutex;
static std::vector<RayTracer *> g_tracers;

RayTracer::RayTracer(Scene &scene) : scene(&scene)
{
	{
		std::lock_guard<std::mutex> lock(g_tracersMutex);
		g_tracers.push_back(this);
	}
	outputBytes = (size_t)2048 * 2048 * 3;   // the reference's fixed framebuffer, RayTracer.cpp:603
	output = new uint8_t[outputBytes];       // plain memory here (no CUDA call during static init);
	memset(output, 127, outputBytes);        // swapped for page-locked memory by the first start()
}

RayTracer::~RayTracer()
{
	{
		std::lock_guard<std::mutex> lock(g_tracersMutex);
		g_tracers.erase(std::remove(g_tracers.begin(), g_tracers.end(), this), g_tracers.end());
	}
	wait();
	if (countedAsCoalescer && residency) residency->addCoalescer(-1);
	if (ctx)
		rt_destroy(ctx);       // the pipeline first, then (with the last tracer of the Scene) the residency
	residency.reset();
	if (output) free_output(output, outputPinned);
}

void RayTracer::ensureContext()
{
	if (ctx)
		return;
	{
		std::lock_guard<std::mutex> lock(g_residencyMutex);
		std::weak_ptr<SceneResidency> &slot = g_residency[std::make_pair((const Scene *)scene, device)];
		residency = slot.lock();
		if (!residency)
		{
			residency = std::make_shared<SceneResidency>();
			const int rc = rt_create(device, &residency->parent);
			if (rc != RT_OK)
			{
				residency.reset();
				fail("rt_create", rc);
			}
			slot = residency;
		}
	}
	const int rc = rt_create_shared(residency->parent, &ctx);
	if (rc != RT_OK)
		fail("rt_create_shared", rc);
	flattener = &residency->flattener;
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

void RayTracer::completeFrame(double seconds, rt_ctx *renderedBy, const char *error)
{
	{
		std::lock_guard<std::mutex> lock(doneMutex);
		if (error) lastError = error, failed = true;
		useTime = seconds;
		lastCtx = renderedBy;
		queued = false;
		isFinish = true;
	}
	doneCv.notify_all();
}

void RayTracer::start(const uint8_t type, const int8_t)
{
	wait();
	failed = false, lastError.clear();
	isFinish = false;
	width = scene->cam.width;
	height = scene->cam.height;
	ensureContext();
	reserveOutput((size_t)width * height * 3);
	for (DrawObject *o : scene->Objects)
		if (o->bShow)
			o->RTPrepare();

	int rc;
	rt_camera frameCamera;
	{
		// unchanged tables are not re-sent; while another tracer of this Scene has a frame in flight the
		// scene must not change (the reference's rule, main.cpp:258,344), so the upload is then a no-op
		std::lock_guard<std::mutex> lock(residency->mutex);
		rt_scene_desc desc;
		flattener->flatten(*scene, desc);
		frameCamera = desc.camera;
		rc = rt_upload_scene(residency->parent, &desc);
	}
	if (rc != RT_OK)
		fail("rt_upload_scene", rc);
	rt_render_params rp;
	memset(&rp, 0, sizeof rp);
	rp.type = type, rp.max_level = maxLevel;
	rp.rank = shardRank, rp.world = shardWorld, rp.flags = renderFlags, rp.tile_rows = shardTileRows;
	const uint64_t key = ((uint64_t)shardRank << 48) ^ ((uint64_t)shardWorld << 32) ^ ((uint64_t)shardTileRows << 24) ^ ((uint64_t)(renderFlags & RT_FLAG_SERPENTINE) << 16)
		^ ((uint64_t)width << 40) ^ (uint64_t)height ^ ((uint64_t)(uintptr_t)output << 1);
	const bool rowsOnly = shardWorld > 1 && key == outputShardKey;
	outputShardKey = shardWorld > 1 ? key : 0;

	if (coalesce && samples.size() <= 1 && progressiveBands <= 1 && type == MY_MODEL_RAYTRACE && !(renderFlags & (RT_FLAG_HIT_IDS | RT_FLAG_STATS | RT_FLAG_BRUTE)))
	{
		// throughput mode: the Scene's batch workers render this frame together with whatever else is waiting
		rc = residency->ensureWorkers();
		if (rc != RT_OK)
			fail("rt_create_shared (batch worker)", rc);
		if (!countedAsCoalescer) residency->addCoalescer(1), countedAsCoalescer = true;
		FrameRequest r;
		r.tracer = this, r.camera = frameCamera, r.rp = rp, r.rowsOnly = rowsOnly;
		{
			std::lock_guard<std::mutex> lock(doneMutex);
			queued = true;
		}
		lastCtx = nullptr;
		residency->enqueue(r);
		return;
	}

	rc = rt_set_sm_share(ctx, smShare);
	if (rc != RT_OK)
		fail("rt_set_sm_share", rc);
	lastCtx = ctx;
	if (samples.size() > 1 && (type == MY_MODEL_RAYTRACE || type == MY_MODEL_REFLECTTEST || type == MY_MODEL_REFRACTTEST))
	{
		// n samples per pixel: the band loop of rt_render_supersampled is synchronous, so it runs in the monitor thread
		std::vector<rt_camera> cams;
		for (const auto &dxy : samples)
		{
			rt_camera c;
			SceneFlattener::cameraRecord(rtscenes_jittered(scene->cam, dxy.first, dxy.second), c);
			cams.push_back(c);
		}
		monitor = std::thread([this, rowsOnly, rp, cams]
		{
			const auto t0 = std::chrono::steady_clock::now();
			int rc = rt_render_supersampled(ctx, &rp, (uint32_t)cams.size(), cams.data());
			if (rc == RT_OK)
				rc = rowsOnly ? rt_read_output_rows(ctx, output, (size_t)width * 3) : rt_read_output(ctx, output, (size_t)width * 3);
			if (rc != RT_OK)
			{
				lastError = rt_last_error();
				fprintf(stderr, "raytrace_b200: supersampled frame failed (%d): %s\n", rc, lastError.c_str());
				failed = true;
			}
			useTime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			isFinish = true;
		});
		return;
	}
	if (progressiveBands > 1)
	{
		// the shard's own tiles, k-th tile = k * world + (rank, or world - 1 - rank in the odd groups of a boustrophedon shard)
		const uint32_t world = shardWorld > 1 ? shardWorld : 1, rank = shardWorld > 1 ? shardRank : 0;
		const bool serp = (renderFlags & RT_FLAG_SERPENTINE) && world > 1;
		const uint32_t tilesInFrame = (uint32_t)(height / 64) * 64u / shardTileRows;
		uint32_t mine = 0;
		while (mine * world + ((serp && (mine & 1u)) ? world - 1u - rank : rank) < tilesInFrame) ++mine;
		const uint32_t per = std::max<uint32_t>(1, (mine + (uint32_t)progressiveBands - 1) / (uint32_t)progressiveBands);
		bandsDone = 0;
		memset(output, 127, (size_t)width * height * 3);      // RayTracer.cpp:620: every start() greys the frame, the bands then fill it in
		monitor = std::thread([this, rp, mine, per]
		{
			double total = 0.0;
			int rc = RT_OK;
			for (uint32_t k0 = 0; k0 < mine && rc == RT_OK; k0 += per)
			{
				rt_render_params band = rp;
				band.tile_first = k0, band.tile_count = std::min(per, mine - k0);
				double seconds = 0.0;
				rc = rt_render_async(ctx, &band);
				if (rc == RT_OK) rc = rt_wait(ctx, &seconds);
				if (rc == RT_OK) rc = rt_read_output_rows(ctx, output, (size_t)width * 3);   // only the band's rows
				total += seconds;
				if (rc == RT_OK) bandsDone = bandsDone + 1;
			}
			if (rc != RT_OK)
			{
				lastError = rt_last_error();
				fprintf(stderr, "raytrace_b200: progressive frame failed (%d): %s\n", rc, lastError.c_str());
				failed = true;
			}
			useTime = total;
			isFinish = true;
		});
		return;
	}
	rc = rt_render_async(ctx, &rp);
	if (rc != RT_OK)
		fail("rt_render_async", rc);

	// (rowsOnly, above: a shard reads back only the rows it rendered once `output` holds a frame of the same
	// shard layout -- the other rows are the 127 fill of RayTracer.cpp:620 and do not change; the first frame of
	// a layout is read back whole.  Decided in start(), not in the monitor thread, from the values it latched.)

	// the monitor thread of RayTracer.cpp:674-695: waits for the frame, publishes it
	monitor = std::thread([this, rowsOnly]
	{
		double seconds = 0.0;
		int rc = rt_wait(ctx, &seconds);
		if (rc == RT_OK)
			rc = rowsOnly ? rt_read_output_rows(ctx, output, (size_t)width * 3) : rt_read_output(ctx, output, (size_t)width * 3);
		if (rc != RT_OK)
		{
			// no abort: the frame ends, flagged (the reference's start() has no error channel; `failed` is ours)
			lastError = rt_last_error();
			fprintf(stderr, "raytrace_b200: frame failed (%d): %s\n", rc, lastError.c_str());
			failed = true;
		}
		useTime = seconds;
		isFinish = true;
	});
}

void RayTracer::stop()
{
	if (residency && queued)
	{
		// a coalesced frame: cancel THIS request only.  Still waiting -> taken out of the queue and completed as
		// cancelled (output untouched); already in a launch -> that launch also holds other tracers' frames and is
		// never stopped, the frame is delivered normally.
		if (residency->cancelPending(this))
			completeFrame(0.0, nullptr);
		return;
	}
	if (ctx)
		rt_stop(ctx);   // own pipeline: cooperative cancel of the running frame (names its epoch; a late stop cancels nothing)
}

void RayTracer::wait()
{
	if (monitor.joinable())
		monitor.join();
	std::unique_lock<std::mutex> lock(doneMutex);
	doneCv.wait(lock, [this] { return !queued; });
}

bool RayTracer::readHitIds(rt_hit_id *ids)
{
	wait();
	return ctx && rt_read_hit_ids(ctx, ids) == RT_OK;
}

bool RayTracer::readCounters(rt_counters *out)
{
	wait();
	// a coalesced frame: the totals of the launch it was part of, valid until that batch worker starts its next launch
	rt_ctx *c = lastCtx ? lastCtx : ctx;
	return c && rt_read_counters(c, out) == RT_OK;
}

// ---- B2: the per-primitive operator, evaluated on the device ----------------------------------------
// HitRes::obj of a Model hit cannot be a clTri address here (the octant lists live on the GPU); it is
// an opaque tagged value that only has to survive a round trip through hr.obj.
static const intptr_t kTriTag = (intptr_t)1 << 62;

static intptr_t encodeObj(const Scene &scene, const rt_hit_id &id)
{
	if (id.object < 0 || id.object >= (int)scene.Objects.size()) return 0;
	DrawObject *o = scene.Objects[id.object];
	if (o->type == MY_OBJECT_MODEL)
		return kTriTag | ((intptr_t)id.object << 33) | ((intptr_t)(id.octant & 7) << 30) | ((intptr_t)(id.sub & 0x7FFF) << 15) | (intptr_t)(id.index & 0x7FFF);
	return (intptr_t)o + id.sub;   // Sphere/Box/Plane: this; BallPlane: this + cnt (Basic3DObject.cpp:538)
}

static rt_hit_id decodeObj(const Scene &scene, intptr_t obj, float distance)
{
	rt_hit_id id = { -1, -1, -1, -1, distance };
	if (obj & kTriTag)
	{
		id.object = (int32_t)((obj >> 33) & 0xFFFF), id.octant = (int32_t)((obj >> 30) & 7);
		id.sub = (int32_t)((obj >> 15) & 0x7FFF), id.index = (int32_t)(obj & 0x7FFF);
		return id;
	}
	for (size_t i = 0; i < scene.Objects.size(); ++i)
	{
		const intptr_t base = (intptr_t)scene.Objects[i];
		const intptr_t span = scene.Objects[i]->type == MY_OBJECT_BALLPLANE ? 16 : 0;
		if (obj >= base && obj <= base + span)
		{
			id.object = (int32_t)i, id.sub = (int32_t)(obj - base);
			return id;
		}
	}
	return id;
}

HitRes RayTracer::intersectObject(uint32_t index, const Ray &ray, const HitRes &hr, const float min)
{
	wait();
	ensureContext();
	for (DrawObject *o : scene->Objects)
		if (o->bShow)
			o->RTPrepare();
	int rc;
	{
		std::lock_guard<std::mutex> lock(residency->mutex);
		rt_scene_desc desc;
		flattener->flatten(*scene, desc);
		rc = rt_upload_scene(residency->parent, &desc);   // no-op when nothing changed since the last frame
	}
	if (rc != RT_OK)
		fail("rt_upload_scene", rc);
	rt_ray r;
	memset(&r, 0, sizeof r);
	r.origin = rt_vec4{ ray.origin.x, ray.origin.y, ray.origin.z, ray.origin.w };
	r.direction = rt_vec4{ ray.direction.x, ray.direction.y, ray.direction.z, ray.direction.w };
	r.mtlrfr = ray.mtlrfr, r.type = ray.type, r.is_inside = ray.isInside;
	rt_hit in, out;
	memset(&in, 0, sizeof in);
	in.material = in.texture = -1;
	in.id = decodeObj(*scene, hr.obj, hr.distance);
	rc = rt_intersect_object(ctx, index, &r, &in, min, &out, 1);
	if (rc != RT_OK)
		fail("rt_intersect_object", rc);
	if (!(out.id.distance < hr.distance))
		return hr;   // `return hr` of every reference operator
	HitRes nh(out.id.distance);
	nh.position = Vertex(out.position.x, out.position.y, out.position.z);
	nh.normal.x = out.normal.x, nh.normal.y = out.normal.y, nh.normal.z = out.normal.z;
	nh.tcoord = Coord2D(out.tu, out.tv);
	nh.mtl = out.material >= 0 ? const_cast<Material *>(flattener->materialPtrs[out.material]) : nullptr;
	nh.tex = out.texture >= 0 ? const_cast<Texture *>(flattener->texturePtrs[out.texture]) : nullptr;
	nh.obj = encodeObj(*scene, out.id);
	nh.rfr = out.rfr, nh.isInside = (uint8_t)out.is_inside;
	return nh;
}

// DrawObject::intersect (3DElement.h:201).  The operator of every shipped primitive lives on the GPU:
// the call is forwarded, for this one ray, to the device context of the RayTracer whose Scene holds
// the object.  An object that belongs to no such Scene cannot be evaluated: loud failure, no CPU path.
HitRes DrawObject::intersect(const Ray &ray, const HitRes &hr, const float min)
{
	RayTracer *owner = nullptr;
	uint32_t index = 0;
	{
		std::lock_guard<std::mutex> lock(g_tracersMutex);
		for (RayTracer *t : g_tracers)
		{
			const std::vector<DrawObject *> &objs = t->attachedScene()->Objects;
			for (size_t i = 0; i < objs.size() && !owner; ++i)
				if (objs[i] == this) owner = t, index = (uint32_t)i;
			if (owner) break;
		}
	}
	if (!owner)
		throw std::runtime_error("raytrace_b200: DrawObject::intersect needs the object to be in a Scene that has a RayTracer "
			"(the operator runs on the GPU, include/rt_b200.h rt_intersect_object); there is no CPU implementation");
	return owner->intersectObject(index, ray, hr, min);
}
