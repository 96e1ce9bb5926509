// C ABI of librt_b200.so (include/rt_b200.h): context, scene upload (SoA flattening on the device
// side + LBVH builds) and the per-frame wavefront schedule.  Host-side glue only; every per-ray
// operation runs in the kernels of rt_kernels.cu.  There is deliberately no CPU path here.
#include "rt_kernels.h"
#include <atomic>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_err;
static std::atomic<uint64_t> g_h2dTotal{0}, g_d2hTotal{0};   // rt_transfer_totals
// landing buffers this process owns: rt_landing_create leaves them filled with 127, and their rows belong to the
// peers as much as to the owner -- a pipeline that renders straight into one must NOT grey it again (that fill
// would not be ordered against the peers' one-sided row copies and could wipe rows that already landed)
static std::mutex g_landingMutex;
static std::vector<const uint8_t *> g_landingBases;
static bool is_landing_base(const uint8_t *p)
{
	std::lock_guard<std::mutex> lock(g_landingMutex);
	for (const uint8_t *b : g_landingBases) if (b == p) return true;
	return false;
}

static int fail(int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_err = buf;
	return code;
}

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(RT_E_CUDA, "%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

template<class T> struct DevBuf
{
	T *p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t n)
	{
		if (n <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr, cap = 0;
		const size_t want = n + n / 16 + 16;
		cudaError_t e = cudaMalloc(&p, want * sizeof(T));
		if (e == cudaSuccess) cap = want;
		return e;
	}
	cudaError_t upload(const T *src, size_t n, cudaStream_t st, uint64_t *bytes = nullptr)
	{
		if (bytes) *bytes += n * sizeof(T);
		g_h2dTotal += n * sizeof(T);
		cudaError_t e = reserve(n);
		if (e != cudaSuccess || n == 0) return e;
		return cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
	}
	void release() { if (p) cudaFree(p); p = nullptr, cap = 0; }
};

struct LevelStore
{
	DevBuf<float4> ray_o, ray_d, hit_p, color, hit_n, hit_uv;
	DevBuf<uint2> ray_meta, hit_id;
	DevBuf<int4> aux;
	DevBuf<uint8_t> shadow;
	DevBuf<uint32_t> hit_list, order, sort_key;
	uint32_t capacity = 0, lights = 0;
	void release() { ray_o.release(), ray_d.release(), hit_p.release(), color.release(), ray_meta.release(), hit_id.release(), aux.release(), shadow.release(), hit_list.release(), hit_n.release(), hit_uv.release(), order.release(), sort_key.release(); capacity = 0; }
};

struct rt_ctx
{
	int device = 0, sms = 148;
	cudaStream_t stream = nullptr, stopStream = nullptr;
	bool ownStream = false;
	uint32_t *hStopWord = nullptr;   // pinned ring of epochs: rt_stop copies the running frame's epoch into WaveState::stop_epoch
	uint32_t stopSlot = 0;
	cudaEvent_t evStart = nullptr, evStop = nullptr, evA = nullptr, evB = nullptr, evRead = nullptr;
	cudaEvent_t evStage[4 * (RT_MAX_LEVELS + 1) + 2];   // per level: before trace, after trace, after shadow, after shade; then combine begin/end
	bool stageTiming = true;
	double traceMs = 0, shadowMs = 0, shadeMs = 0, otherMs = 0;
	uint8_t *extOut = nullptr;      // caller-provided device framebuffer (rt_set_output)
	size_t extOutBytes = 0;

	// resident scene
	bool hasScene = false;
	uint64_t geometryEpoch = 0;
	std::vector<rt_prim> prims;
	std::vector<rt_model> models;
	std::vector<rt_part> parts;
	std::vector<rt_light> lights;
	std::vector<rt_material> matCache;      // last uploaded tables: unchanged tables are not re-sent
	std::vector<rt_texture> texCache;
	std::vector<uint8_t> texelCache;
	rt_camera camera;
	rt_vec4 envLight;
	bool anyRefract = false;
	uint32_t nTris = 0;
	DevBuf<float4> primGeom, materials, triPoints, triNorms, triGeomOrig, triGeom, boxLo, boxHi, partMid, partPos;
	DevBuf<int4> primMeta, textures;
	DevBuf<uint8_t> texels;
	DevBuf<float2> triTcoords;
	DevBuf<uint32_t> bvhPrims, triSlot, triPart, leafOrder, leafOrderAll;
	std::vector<std::vector<uint32_t>> modelLevels;   // per model: 4-wide nodes per tree level (rtb_refit4)
	std::vector<uint32_t> modelNodeBase;
	bool lastUploadRefit = false;
	DevBuf<DevModel> dModels;
	DevBuf<DevPart> dParts;
	DevBuf<BvhNode> nodes;      // binary LBVH (build-time only)
	DevBuf<BvhNode4> nodes4;    // 4-wide collapse used by the traversal
	DevBuf<BvhNode8> nodes8;    // 8-wide quantised collapse of the Model BVHs (RT_BVH8)
	uint32_t bvhStackNeed = 0;  // deepest traversal stack any path of the 8-wide trees can need
	// Coherence binning of the secondary rays (wave scheduler, rtk_bin_rays).  RT_B200_BIN = cell bits per axis, 0 = off;
	// unset = auto: 3 bits when the scene refracts -- binary ray trees interleave the reflect and refract children of every
	// parent warp in the queues (c4: +32 %) -- and off otherwise, where the queues inherit the screen-space order of the
	// primary rays and binning only costs its three launches per level (c3: -8 %, c2: -7 %; profiles/r2_binning.md)
	int binMode = -1;
	uint32_t binBits = 0;
	BinGrid binGrid = { { 0, 0, 0 }, { 1, 1, 1 }, 1, 7 };
	DevBuf<uint32_t> binHist;
	DevBuf<SceneItem> items;
	SceneDev S;
	BuildScratch *scratch = nullptr;
	uint32_t bvhNodes = 0, bvhDepth = 0, leafSize = 2;

	// per frame
	FrameParams *hFrame = nullptr, *dFrame = nullptr;
	WaveState *hWave = nullptr, *hWaveInit = nullptr, *dWave = nullptr;   // hWave: D2H results, hWaveInit: H2D initial state
	LevelStore levels[RT_MAX_LEVELS + 2];
	DevBuf<uint8_t> out;
	int outW = 0, outH = 0;
	uint8_t *fb = nullptr;          // framebuffer of the last frame (c->out or extOut)
	rt_render_params lastParams;
	uint32_t lastPixels = 0, lastLaunches = 0, lastMaxLevel = 0, lastTileFirst = 0;
	rt_counters ssTotals;           // rt_render_supersampled: ray counts and times summed over the bands of the last frame
	bool ssValid = false;
	bool frameInFlight = false, frameValid = false;
	double uploadMs = 0, buildMs = 0, renderMs = 0;
	uint64_t uploadBytes = 0, frameH2D = 0, frameD2H = 0;
	float levelFactor = 2.0f;
	uint32_t reserveFrames = 0;   // rt_reserve_batch: size the ray queues for launches of up to this many frames at once
	uint32_t minCap[RT_MAX_LEVELS + 2] = {};   // per-level queue capacities learnt from overflowing frames (finish_frame regrows and re-renders)
	uint32_t regrowTries = 0;
	std::vector<rt_camera> lastCams;           // the last launch's arguments, kept for that re-render
	bool lastCamsGiven = false, lastOutsGiven = false;
	int schedMode = 0;              // 0 auto, 1 k_frame (one persistent launch per frame), 2 per-level waves (RT_B200_SCHED=auto|frame|waves)
	bool frameSched = false;        // what the last frame used
	int travWalk = 0;               // RT_B200_TRAV: how the wave kernels walk the scene -- voted (default: one batch of 32 rays at a time), split (two-stage: rays that enter the last Model's BVH are collected per warp; bit-exact, -5 % instructions, measured 2 % slower)
	int waveGen = 1;                // k_wave(0) makes the primary rays itself (RT_B200_WAVE_GENPRIMARY=0: k_raygen writes them first)
	uint32_t frameEpoch = 0;
	// several frames in flight on one GPU: a pipeline created by rt_create_shared renders its parent's
	// resident scene (device tables and BVH are NOT copied), on its own stream with its own ray queues
	rt_ctx *sceneFrom = nullptr;
	uint64_t sceneVersion = 0, adoptedVersion = 0;     // bumped by every rt_upload_scene of the parent
	uint64_t tablesVersion = 0, adoptedTables = 0;     // bumped when prims / models / parts changed
	const uint8_t *ssFillPtr = nullptr; // rt_render_supersampled: the target it last greyed
	int ssFillW = 0, ssFillH = 0;
	uint64_t ssFillShard = ~0ull;
	bool regrewLastFrame = false;       // finish_frame regrew a ray level and rendered the launch again
	const uint8_t *fillPtr = nullptr;   // what the framebuffer was last filled with 127 for
	int fillW = 0, fillH = 0;
	uint32_t fillRank = 0, fillWorld = 0, fillTile = 0, fillSerp = 0;
	std::vector<DevBuf<uint8_t>> batchOut;   // library-owned framebuffers of batch frames (rt_render_batch_async without outputs)
	std::vector<uint8_t *> batchFill;        // which buffer each batch slot was last greyed for
	uint32_t lastBatch = 1;                  // frames of the last launch
	uint8_t *lastOuts[RT_MAX_BATCH] = {};    // their framebuffers
	uint32_t keepSalt = 0;      // rotates which CTAs of k_frame are pinned, per pipeline (RT_B200_KEEP_DIV)
	unsigned ctasPerSm = 0;         // resident traversal CTAs per SM this pipeline may use (0 = all 8), rt_set_sm_share
};

#ifndef RT_WAVE_GENPRIMARY_DEFAULT
#define RT_WAVE_GENPRIMARY_DEFAULT 1   // k_wave(0) makes the primary rays itself (RT_B200_WAVE_GENPRIMARY=0: k_raygen writes them first); C3 +1.5 %, C2 +1.9 %
#endif

extern "C" const char *rt_last_error(void) { return g_err.c_str(); }
extern "C" int rt_abi_version(void) { return RT_ABI_VERSION; }

extern "C" int rt_create(int device, rt_ctx **out)
{
	if (!out) return fail(RT_E_INVALID, "rt_create: out is NULL");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return fail(RT_E_NODEVICE, "rt_create: no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
	if (device < 0 || device >= n) return fail(RT_E_INVALID, "rt_create: device %d out of range (%d devices)", device, n);
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
		return fail(RT_E_NODEVICE, "rt_create: device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major, prop.minor);
	CU(cudaSetDevice(device));
	rt_ctx *c = new rt_ctx();
	c->device = device, c->sms = prop.multiProcessorCount;
	CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	c->ownStream = true;
	CU(cudaStreamCreateWithFlags(&c->stopStream, cudaStreamNonBlocking));
	CU(cudaMallocHost(&c->hStopWord, 64 * sizeof(uint32_t)));
	CU(cudaEventCreate(&c->evStart)); CU(cudaEventCreate(&c->evStop)); CU(cudaEventCreate(&c->evA)); CU(cudaEventCreateWithFlags(&c->evB, cudaEventBlockingSync)); CU(cudaEventCreateWithFlags(&c->evRead, cudaEventDisableTiming | cudaEventBlockingSync));
	for (auto &e : c->evStage) CU(cudaEventCreate(&e));
	if (const char *v = getenv("RT_B200_STAGE_TIMING")) c->stageTiming = atoi(v) != 0;
	CU(cudaMallocHost(&c->hFrame, sizeof(FrameParams)));
	CU(cudaMalloc(&c->dFrame, sizeof(FrameParams)));
	CU(cudaMallocHost(&c->hWave, sizeof(WaveState)));
	CU(cudaMallocHost(&c->hWaveInit, sizeof(WaveState)));
	CU(cudaMalloc(&c->dWave, sizeof(WaveState)));
	memset(&c->S, 0, sizeof c->S);
	if (const char *v = getenv("RT_B200_LEAF_SIZE")) c->leafSize = (uint32_t)atoi(v);
	if (const char *v = getenv("RT_B200_LEVEL_FACTOR")) c->levelFactor = (float)atof(v);
	if (const char *v = getenv("RT_B200_SCHED")) c->schedMode = !strcmp(v, "frame") ? 1 : (!strcmp(v, "waves") ? 2 : 0);
	if (const char *v = getenv("RT_B200_TRAV")) c->travWalk = !strcmp(v, "split") ? 2 : (!strcmp(v, "defer") ? 3 : (!strcmp(v, "steal") ? 4 : 0));
	if (const char *v = getenv("RT_B200_BIN")) { int b = atoi(v); c->binMode = b < 0 ? 0 : (b > 6 ? 6 : b); }
	c->waveGen = RT_WAVE_GENPRIMARY_DEFAULT;
	if (const char *v = getenv("RT_B200_WAVE_GENPRIMARY")) c->waveGen = atoi(v);
	{ static std::atomic<uint32_t> created{0}; c->keepSalt = created.fetch_add(1u); }
	*out = c;
	return RT_OK;
}

extern "C" int rt_create_shared(rt_ctx *parent, rt_ctx **out)
{
	if (!parent || !out) return fail(RT_E_INVALID, "rt_create_shared: NULL argument");
	if (parent->sceneFrom) parent = parent->sceneFrom;   // a sibling of a shared pipeline: same scene owner
	int rc = rt_create(parent->device, out);
	if (rc != RT_OK) return rc;
	(*out)->sceneFrom = parent;
	return RT_OK;
}

extern "C" int rt_set_sm_share(rt_ctx *c, int ctas_per_sm)
{
	if (!c) return fail(RT_E_INVALID, "rt_set_sm_share: ctx is NULL");
	if (ctas_per_sm < 0 || ctas_per_sm > 8) return fail(RT_E_INVALID, "rt_set_sm_share: %d CTAs per SM (0 = default, 1..8)", ctas_per_sm);
	c->ctasPerSm = (unsigned)ctas_per_sm;
	return RT_OK;
}

// A shared pipeline looks at its parent's scene: host-side tables are copied (small), device tables are
// aliased.  rt_upload_scene leaves the parent's stream synchronised whenever it touched device memory,
// so the aliases are valid for kernels on any stream as soon as the upload call has returned.
static void adopt_scene(rt_ctx *c)
{
	const rt_ctx *p = c->sceneFrom;
	if (!p || c->adoptedVersion == p->sceneVersion) return;
	c->hasScene = p->hasScene, c->geometryEpoch = p->geometryEpoch;
	if (c->adoptedTables != p->tablesVersion || c->adoptedVersion == 0)
		c->prims = p->prims, c->models = p->models, c->parts = p->parts, c->adoptedTables = p->tablesVersion;
	c->lights = p->lights;
	c->camera = p->camera, c->envLight = p->envLight, c->anyRefract = p->anyRefract, c->nTris = p->nTris;
	c->binGrid = p->binGrid, c->binBits = p->binBits;
	c->S = p->S, c->bvhNodes = p->bvhNodes, c->bvhDepth = p->bvhDepth, c->leafSize = p->leafSize;
	c->uploadMs = p->uploadMs, c->buildMs = p->buildMs, c->uploadBytes = p->uploadBytes;
	c->adoptedVersion = p->sceneVersion;
	c->frameValid = false;
}

extern "C" void rt_destroy(rt_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	for (auto &l : c->levels) l.release();
	c->primGeom.release(), c->materials.release(), c->triPoints.release(), c->triNorms.release(), c->triGeomOrig.release(), c->triGeom.release();
	c->boxLo.release(), c->boxHi.release(), c->partMid.release(), c->partPos.release(), c->primMeta.release(), c->textures.release(), c->texels.release();
	c->triTcoords.release(), c->bvhPrims.release(), c->triSlot.release(), c->triPart.release(), c->leafOrder.release(), c->leafOrderAll.release();
	c->dModels.release(), c->dParts.release(), c->nodes.release(), c->nodes4.release(), c->nodes8.release(), c->items.release(), c->out.release();
	for (auto &b : c->batchOut) b.release();
	rtb_free_scratch(c->scratch);
	cudaFreeHost(c->hFrame), cudaFree(c->dFrame), cudaFreeHost(c->hWave), cudaFreeHost(c->hWaveInit), cudaFree(c->dWave);
	cudaEventDestroy(c->evStart), cudaEventDestroy(c->evStop), cudaEventDestroy(c->evA), cudaEventDestroy(c->evB), cudaEventDestroy(c->evRead);
	for (auto &e : c->evStage) cudaEventDestroy(e);
	if (c->ownStream) cudaStreamDestroy(c->stream);
	cudaStreamDestroy(c->stopStream), cudaFreeHost(c->hStopWord);
	delete c;
}

extern "C" int rt_set_stream(rt_ctx *c, void *s)
{
	if (!c) return fail(RT_E_INVALID, "rt_set_stream: ctx is NULL");
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	if (c->ownStream) cudaStreamDestroy(c->stream);
	c->stream = (cudaStream_t)s, c->ownStream = false;
	return RT_OK;
}

static inline float4 f4(const rt_vec4 &v) { return make_float4(v.x, v.y, v.z, v.w); }
static inline rt_vec4 addv(const rt_vec4 &a, const rt_vec4 &b) { return rt_vec4{ a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }

template<class T> static bool same(const std::vector<T> &a, const T *b, size_t n)
{
	return a.size() == n && (n == 0 || memcmp(a.data(), b, n * sizeof(T)) == 0);
}

static int upload_scene_tables(rt_ctx *c, const rt_scene_desc *s, bool hadScene);

extern "C" int rt_upload_scene(rt_ctx *c, const rt_scene_desc *s)
{
	if (c && c->sceneFrom) return rt_upload_scene(c->sceneFrom, s);   // a shared pipeline has no scene of its own
	if (!c || !s) return fail(RT_E_INVALID, "rt_upload_scene: NULL argument");
	CU(cudaSetDevice(c->device));
	if (c->frameInFlight) { CU(cudaEventSynchronize(c->evB)); c->frameInFlight = false; }
	if (s->n_lights > RT_MAX_LIGHTS) return fail(RT_E_LIMIT, "rt_upload_scene: %u lights (max %d, Scene.cpp:85)", s->n_lights, RT_MAX_LIGHTS);
	if (s->n_tris >= 0x0FFFFFFFu) return fail(RT_E_LIMIT, "rt_upload_scene: too many triangles");
	for (uint32_t i = 1; i < s->n_prims; ++i)
		if (s->prims[i].object < s->prims[i - 1].object || (s->prims[i].object == s->prims[i - 1].object && s->prims[i].sub <= s->prims[i - 1].sub))
			return fail(RT_E_INVALID, "rt_upload_scene: prims must be sorted by (object, sub)");
	for (uint32_t i = 0; i < s->n_prims; ++i)
	{
		const rt_prim &p = s->prims[i];
		if (p.kind != RT_OBJ_SPHERE && p.kind != RT_OBJ_CUBE && p.kind != RT_OBJ_PLANE) return fail(RT_E_INVALID, "rt_upload_scene: prim %u has kind %u", i, p.kind);
		if (p.material >= s->n_materials || p.texture >= (int32_t)s->n_textures) return fail(RT_E_INVALID, "rt_upload_scene: prim %u material/texture out of range", i);
	}
	uint32_t triCheck = 0;
	for (uint32_t m = 0; m < s->n_models; ++m)
	{
		const rt_model &M = s->models[m];
		if (M.part_begin + M.part_count > s->n_parts) return fail(RT_E_INVALID, "rt_upload_scene: model %u part range", m);
		if (m && M.object <= s->models[m - 1].object) return fail(RT_E_INVALID, "rt_upload_scene: models must be sorted by object");
		for (uint32_t p = 0; p < M.part_count; ++p)
		{
			const rt_part &P = s->parts[M.part_begin + p];
			if (P.tri_begin != triCheck) return fail(RT_E_INVALID, "rt_upload_scene: part triangle ranges must be contiguous in part order");
			if (P.tri_count > 32767) return fail(RT_E_LIMIT, "rt_upload_scene: part with %u triangles (clTri::numb is int16)", P.tri_count);
			if (P.material >= s->n_materials || P.texture >= (int32_t)s->n_textures) return fail(RT_E_INVALID, "rt_upload_scene: part material/texture out of range");
			triCheck += P.tri_count;
		}
	}
	if (triCheck != s->n_tris) return fail(RT_E_INVALID, "rt_upload_scene: parts cover %u triangles, n_tris = %u", triCheck, s->n_tris);

	// From here on device tables are touched: a failure leaves them half-written (buffers may have been freed by a
	// reserve), so after one the context has NO scene until a later upload succeeds; shared pipelines see the new
	// version, re-adopt and refuse to render.
	const bool hadScene = c->hasScene;
	const int rc = upload_scene_tables(c, s, hadScene);
	if (rc != RT_OK)
	{
		c->hasScene = false;
		++c->sceneVersion;
		memset(&c->S, 0, sizeof c->S);
		c->geometryEpoch = 0, c->nTris = 0;   // force a full re-upload next time
		c->matCache.clear(), c->texCache.clear(), c->texelCache.clear(), c->prims.clear(), c->models.clear(), c->parts.clear();
		return rc;
	}
	c->hasScene = true;
	c->frameValid = false;
	++c->sceneVersion;
	return RT_OK;
}

static int upload_scene_tables(rt_ctx *c, const rt_scene_desc *s, bool hadScene)
{
	cudaStream_t st = c->stream;
	CU(cudaEventRecord(c->evA, st));
	c->uploadBytes = 0;
	uint64_t *ub = &c->uploadBytes;

	// ---- cheap tables: always refreshed -------------------------------------------------------
	c->camera = s->camera, c->envLight = s->env_light;
	c->lights.assign(s->lights, s->lights + s->n_lights);
	if (!same(c->matCache, s->materials, s->n_materials))
	{
		std::vector<float4> m(4 * (size_t)s->n_materials);
		c->anyRefract = false;
		for (uint32_t i = 0; i < s->n_materials; ++i)
		{
			const rt_material &r = s->materials[i];
			m[4 * i] = f4(r.ambient), m[4 * i + 1] = f4(r.diffuse), m[4 * i + 2] = f4(r.specular);
			m[4 * i + 3] = make_float4(r.shiness, r.reflect, r.refract, r.rfr);
			if (r.refract > 0.01f) c->anyRefract = true;
		}
		CU(c->materials.upload(m.data(), m.size(), st, ub));
		CU(cudaStreamSynchronize(st));   // staging vector goes out of scope
		c->matCache.assign(s->materials, s->materials + s->n_materials);
	}
	if (!same(c->texCache, s->textures, s->n_textures) || !same(c->texelCache, s->texels, s->texel_bytes))
	{
		std::vector<int4> t(s->n_textures);
		for (uint32_t i = 0; i < s->n_textures; ++i) t[i] = make_int4(s->textures[i].w, s->textures[i].h, (int)s->textures[i].offset, 0);
		CU(c->textures.upload(t.data(), t.size(), st, ub));
		CU(c->texels.upload(s->texels, s->texel_bytes, st, ub));
		CU(cudaStreamSynchronize(st));
		c->texCache.assign(s->textures, s->textures + s->n_textures);
		c->texelCache.assign(s->texels, s->texels + s->texel_bytes);
	}

	// ---- what changed? ------------------------------------------------------------------------
	const bool trisChanged = !hadScene || s->geometry_epoch == 0 || s->geometry_epoch != c->geometryEpoch || s->n_tris != c->nTris;
	const bool modelsChanged = trisChanged || !same(c->models, s->models, s->n_models) || !same(c->parts, s->parts, s->n_parts);
	const bool primsChanged = !hadScene || !same(c->prims, s->prims, s->n_prims);
	if (trisChanged && s->n_tris && (!s->tri_points || !s->tri_norms || !s->tri_tcoords))
		return fail(RT_E_INVALID, "rt_upload_scene: geometry changed but triangle arrays are NULL");
	// Position-only edit (Scene::MovePos of a Model, Scene.cpp:248): same triangles, same parts, same objects -- only
	// rt_model::position differs.  The reference's RTPrepare then merely re-translates its bounds (Model.cpp:404,418-419);
	// here the BVHs keep their topology and are REFITTED (rtb_refit4) instead of rebuilt.
	std::vector<uint32_t> movedModels;
	bool refitOnly = false;
	{
		static const int allow = []{ const char *e = getenv("RT_B200_REFIT"); return (e && !atoi(e)) ? 0 : 1; }();
		if (allow && hadScene && !trisChanged && !primsChanged && modelsChanged && c->models.size() == s->n_models
			&& same(c->parts, s->parts, s->n_parts) && c->modelLevels.size() == s->n_models)
		{
			refitOnly = true;
			for (uint32_t m = 0; m < s->n_models && refitOnly; ++m)
			{
				rt_model a = c->models[m], b = s->models[m];
				const bool moved = memcmp(&a.position, &b.position, sizeof a.position) != 0;
				a.position = b.position;
				if (memcmp(&a, &b, sizeof a) != 0) refitOnly = false;                     // something besides the position changed
				else if (moved && c->modelLevels[m].empty() && b.part_count) refitOnly = false;   // no level table (tiny or legacy-collapsed tree)
				else if (moved) movedModels.push_back(m);
			}
		}
	}
	c->lastUploadRefit = false;
	c->prims.assign(s->prims, s->prims + s->n_prims);
	c->models.assign(s->models, s->models + s->n_models);
	c->parts.assign(s->parts, s->parts + s->n_parts);
	c->nTris = s->n_tris;
	c->geometryEpoch = s->geometry_epoch;

	if (trisChanged && s->n_tris)
	{
		CU(c->triPoints.upload((const float4 *)s->tri_points, 3 * (size_t)s->n_tris, st, ub));
		CU(c->triNorms.upload((const float4 *)s->tri_norms, 3 * (size_t)s->n_tris, st, ub));
		CU(c->triTcoords.upload((const float2 *)s->tri_tcoords, 3 * (size_t)s->n_tris, st, ub));
		std::vector<uint32_t> tp(s->n_tris);
		for (uint32_t p = 0; p < s->n_parts; ++p)
			for (uint32_t k = 0; k < s->parts[p].tri_count; ++k) tp[s->parts[p].tri_begin + k] = p;
		CU(c->triPart.upload(tp.data(), tp.size(), st, ub));
		CU(cudaStreamSynchronize(st));
	}

	if (primsChanged || modelsChanged)
		++c->tablesVersion;
	if (refitOnly)
	{
		// ---- refit: new translated bounds, new clTri records (p0 + position, octant membership), same trees ----
		cudaEvent_t b0 = c->evB;
		CU(cudaEventRecord(b0, st));
		std::vector<DevModel> dm(s->n_models);
		std::vector<DevPart> dp(s->n_parts);
		std::vector<float4> pos(s->n_parts);
		for (uint32_t m = 0; m < s->n_models; ++m)
		{
			const rt_model &M = s->models[m];
			memset(&dm[m], 0, sizeof(DevModel));
			dm[m].border_min = f4(addv(M.ver_min, M.position)), dm[m].border_max = f4(addv(M.ver_max, M.position));
			dm[m].part_begin = M.part_begin, dm[m].part_count = M.part_count, dm[m].object = M.object;
			uint32_t tb = 0xFFFFFFFFu, tc = 0;
			for (uint32_t q = 0; q < M.part_count; ++q)
			{
				const rt_part &P = s->parts[M.part_begin + q];
				if (tb == 0xFFFFFFFFu) tb = P.tri_begin;
				tc += P.tri_count;
				DevPart &D = dp[M.part_begin + q];
				D.box_min = f4(addv(P.border_min, M.position)), D.box_max = f4(addv(P.border_max, M.position));
				D.tri_begin = P.tri_begin, D.tri_count = P.tri_count, D.material = P.material, D.texture = P.texture;
				pos[M.part_begin + q] = f4(M.position);
			}
			dm[m].tri_begin = tb == 0xFFFFFFFFu ? 0 : tb, dm[m].tri_count = tc;
		}
		CU(c->dModels.upload(dm.data(), dm.size(), st, ub));
		CU(c->dParts.upload(dp.data(), dp.size(), st, ub));
		CU(c->partPos.upload(pos.data(), pos.size(), st, ub));
		for (uint32_t m : movedModels)
		{
			const DevModel &M = dm[m];
			if (M.tri_count == 0) continue;
			TriPrepArgs a;
			a.points = c->triPoints.p + 3 * (size_t)M.tri_begin, a.models = c->dModels.p, a.parts = c->dParts.p;
			a.tri_part = c->triPart.p + M.tri_begin, a.part_mid_pos = c->partMid.p, a.part_position = c->partPos.p;
			a.tri_geom_orig = c->triGeomOrig.p + 3 * (size_t)M.tri_begin, a.box_lo = c->boxLo.p, a.box_hi = c->boxHi.p, a.n = M.tri_count;
			a.id_base = M.tri_begin;
			rtb_prepare_tris(st, a);
			rtb_scatter_tris(st, c->triGeomOrig.p, c->leafOrderAll.p + M.tri_begin, M.tri_begin, M.tri_begin, M.tri_count, c->triGeom.p, c->triSlot.p);
			const std::vector<uint32_t> &lv = c->modelLevels[m];
#if RT_BVH8
			rtb_refit8(st, c->nodes8.p, c->modelNodeBase[m], lv.data(), (uint32_t)lv.size(), c->boxLo.p, c->boxHi.p, c->leafOrderAll.p);
#else
			rtb_refit4(st, c->nodes4.p, c->modelNodeBase[m], lv.data(), (uint32_t)lv.size(), c->boxLo.p, c->boxHi.p, c->leafOrderAll.p);
#endif
		}
		CU(cudaGetLastError());
		CU(cudaEventRecord(c->evStop, st));
		CU(cudaStreamSynchronize(st));
		float ms = 0;
		cudaEventElapsedTime(&ms, b0, c->evStop);
		c->buildMs = ms;
		c->lastUploadRefit = true;
	}
	else if (primsChanged || modelsChanged)
	{
		// ---- analytic primitives ------------------------------------------------------------------
		std::vector<float4> pg(4 * (size_t)s->n_prims);
		std::vector<int4> pm(s->n_prims);
		for (uint32_t i = 0; i < s->n_prims; ++i)
		{
			const rt_prim &p = s->prims[i];
			pg[4 * i] = make_float4(p.position.x, p.position.y, p.position.z, p.radius);
			if (p.kind == RT_OBJ_SPHERE)
				pg[4 * i + 1] = make_float4(p.radius_sqr, 0, 0, 0), pg[4 * i + 2] = pg[4 * i + 3] = make_float4(0, 0, 0, 0);
			else if (p.kind == RT_OBJ_CUBE)
			{
				// Box::intersect tests (min + position, max + position), Basic3DObject.cpp:280
				pg[4 * i + 1] = f4(addv(p.a, p.position)), pg[4 * i + 2] = f4(addv(p.b, p.position)), pg[4 * i + 3] = f4(p.b);
			}
			else
				pg[4 * i + 1] = f4(p.a), pg[4 * i + 2] = f4(p.b), pg[4 * i + 3] = f4(p.c);
			pm[i] = make_int4((int)p.kind, (int)p.material, p.texture, (int)((p.object << 8) | (p.sub & 0xFF)));
		}
		CU(c->primGeom.upload(pg.data(), pg.size(), st, ub));
		CU(c->primMeta.upload(pm.data(), pm.size(), st, ub));

		// ---- models / parts -----------------------------------------------------------------------
		std::vector<DevModel> dm(s->n_models);
		std::vector<DevPart> dp(s->n_parts);
		std::vector<float4> mid(s->n_parts), pos(s->n_parts);
		for (uint32_t m = 0; m < s->n_models; ++m)
		{
			const rt_model &M = s->models[m];
			memset(&dm[m], 0, sizeof(DevModel));
			dm[m].border_min = f4(addv(M.ver_min, M.position)), dm[m].border_max = f4(addv(M.ver_max, M.position));
			dm[m].part_begin = M.part_begin, dm[m].part_count = M.part_count, dm[m].object = M.object;
			uint32_t tb = 0xFFFFFFFFu, tc = 0;
			for (uint32_t p = 0; p < M.part_count; ++p)
			{
				const rt_part &P = s->parts[M.part_begin + p];
				if (tb == 0xFFFFFFFFu) tb = P.tri_begin;
				tc += P.tri_count;
				DevPart &D = dp[M.part_begin + p];
				D.box_min = f4(addv(P.border_min, M.position)), D.box_max = f4(addv(P.border_max, M.position));
				D.tri_begin = P.tri_begin, D.tri_count = P.tri_count, D.material = P.material, D.texture = P.texture;
				// va = (va + vb) * 0.5, Model.cpp:421 (untranslated)
				const rt_vec4 sum = addv(P.border_min, P.border_max);
				mid[M.part_begin + p] = make_float4(sum.x * 0.5f, sum.y * 0.5f, sum.z * 0.5f, 0);
				pos[M.part_begin + p] = f4(M.position);
			}
			dm[m].tri_begin = tb == 0xFFFFFFFFu ? 0 : tb, dm[m].tri_count = tc;
		}
		CU(c->dModels.upload(dm.data(), dm.size(), st, ub));
		CU(c->dParts.upload(dp.data(), dp.size(), st, ub));
		CU(c->partMid.upload(mid.data(), mid.size(), st, ub));
		CU(c->partPos.upload(pos.data(), pos.size(), st, ub));
		CU(cudaStreamSynchronize(st));

		// ---- scene-order item list + BVH budget ---------------------------------------------------
		const uint32_t runMin = 8;   // runs of >= runMin bounded primitives get their own BVH
		std::vector<SceneItem> items;
		uint32_t pi = 0, mi = 0, nodeBudget = 0, primLeafSlots = 0;
		while (pi < s->n_prims || mi < s->n_models)
		{
			const bool takePrim = mi >= s->n_models || (pi < s->n_prims && s->prims[pi].object < s->models[mi].object);
			if (!takePrim)
			{
				items.push_back(SceneItem{ RT_ITEM_MODEL, mi, 0, 0 });
				nodeBudget += dm[mi].tri_count;
				++mi;
				continue;
			}
			if (s->prims[pi].kind == RT_OBJ_PLANE) { items.push_back(SceneItem{ RT_ITEM_PRIM, pi, 1, 0 }); ++pi; continue; }
			uint32_t end = pi;
			const uint32_t limitObj = mi < s->n_models ? s->models[mi].object : 0xFFFFFFFFu;
			while (end < s->n_prims && s->prims[end].kind != RT_OBJ_PLANE && s->prims[end].object < limitObj) ++end;
			if (end - pi >= runMin)
			{
				items.push_back(SceneItem{ RT_ITEM_PRIMBVH, pi, end - pi, 0 });
				nodeBudget += end - pi, primLeafSlots += end - pi;
			}
			else
				for (uint32_t k = pi; k < end; ++k) items.push_back(SceneItem{ RT_ITEM_PRIM, k, 1, 0 });
			pi = end;
		}
		CU(c->nodes.reserve(nodeBudget + 1));
		CU(c->nodes4.reserve(nodeBudget + 1));
#if RT_BVH8
		CU(c->nodes8.reserve(nodeBudget + 1));
#endif
		CU(c->bvhPrims.reserve(primLeafSlots + 1));
		CU(c->triGeomOrig.reserve(3 * (size_t)s->n_tris + 1));
		CU(c->triGeom.reserve(RT_TRI_F4 * (size_t)s->n_tris + 1));
		CU(c->triSlot.reserve(s->n_tris + 1));
		CU(c->leafOrderAll.reserve(s->n_tris + 1));
		c->modelLevels.assign(s->n_models, std::vector<uint32_t>());
		c->modelNodeBase.assign(s->n_models, 0u);
		uint32_t maxBoxes = 1;
		for (const SceneItem &it : items)
			maxBoxes = std::max(maxBoxes, it.kind == RT_ITEM_MODEL ? dm[it.first].tri_count : it.count);
		CU(c->boxLo.reserve(maxBoxes)); CU(c->boxHi.reserve(maxBoxes)); CU(c->leafOrder.reserve(maxBoxes));

		// ---- builds ---------------------------------------------------------------------------------
		cudaEvent_t b0 = c->evB;
		CU(cudaEventRecord(b0, st));
		uint32_t nodeCursor = 0, primLeafCursor = 0;
		c->bvhDepth = 0, c->bvhStackNeed = 0;
		for (SceneItem &it : items)
		{
			BvhBuildResult res;
			if (it.kind == RT_ITEM_PRIMBVH)
			{
				PrimBoxArgs a{ c->primGeom.p, c->primMeta.p, c->boxLo.p, c->boxHi.p, it.first, it.count };
				rtb_prim_boxes(st, a);
				int rc = rtb_build(st, &c->scratch, c->boxLo.p, c->boxHi.p, it.count, 2, c->nodes.p, c->nodes4.p, nodeCursor, primLeafCursor, c->bvhPrims.p + primLeafCursor, &res);
				if (rc) return fail(RT_E_CUDA, "LBVH build (primitives) failed: %s", cudaGetErrorString((cudaError_t)rc));
				rtb_offset_order(st, c->bvhPrims.p + primLeafCursor, it.count, it.first);
				it.root = res.root;
				nodeCursor += res.nodesUsed, primLeafCursor += it.count;
				c->bvhDepth = std::max(c->bvhDepth, res.depth);
			}
			else if (it.kind == RT_ITEM_MODEL)
			{
				const DevModel &M = dm[it.first];
				if (M.tri_count == 0) { it.root = (int)0x80000000u; it.count = 0; continue; }
				TriPrepArgs a;
				a.points = c->triPoints.p + 3 * (size_t)M.tri_begin, a.models = c->dModels.p, a.parts = c->dParts.p;
				a.tri_part = c->triPart.p + M.tri_begin, a.part_mid_pos = c->partMid.p, a.part_position = c->partPos.p;
				a.tri_geom_orig = c->triGeomOrig.p + 3 * (size_t)M.tri_begin, a.box_lo = c->boxLo.p, a.box_hi = c->boxHi.p, a.n = M.tri_count;
				a.id_base = M.tri_begin;
				rtb_prepare_tris(st, a);
				int rc = rtb_build(st, &c->scratch, c->boxLo.p, c->boxHi.p, M.tri_count, c->leafSize, c->nodes.p, c->nodes4.p, nodeCursor, M.tri_begin, c->leafOrder.p, &res, RT_BVH8 ? c->nodes8.p : nullptr);
				c->bvhStackNeed = std::max(c->bvhStackNeed, res.maxStack);
				if (rc) return fail(RT_E_CUDA, "LBVH build (model %u) failed: %s", it.first, cudaGetErrorString((cudaError_t)rc));
				rtb_scatter_tris(st, c->triGeomOrig.p, c->leafOrder.p, M.tri_begin, M.tri_begin, M.tri_count, c->triGeom.p, c->triSlot.p);
				// kept for refits after position-only edits: the leaf order (global slot -> model-local triangle) and the level table
				CU(cudaMemcpyAsync(c->leafOrderAll.p + M.tri_begin, c->leafOrder.p, sizeof(uint32_t) * M.tri_count, cudaMemcpyDeviceToDevice, st));
				c->modelLevels[it.first].assign(res.levelNodes, res.levelNodes + res.nLevels);
				c->modelNodeBase[it.first] = nodeCursor;
				it.root = res.root, it.count = M.tri_count;
				nodeCursor += res.nodesUsed;
				c->bvhDepth = std::max(c->bvhDepth, res.depth);
			}
		}
		c->bvhNodes = nodeCursor;
		// a 4-wide step pushes up to three siblings and descends two binary levels
		if (3 * ((c->bvhDepth + 1) / 2) + 1 > RT_STACK)
			return fail(RT_E_LIMIT, "LBVH depth %u exceeds the traversal stack (%d)", c->bvhDepth, RT_STACK);
		// an 8-wide step pushes up to seven siblings: the collapse reports the deepest need over all root-to-leaf paths
		if (c->bvhStackNeed + 1 > RT_STACK)
			return fail(RT_E_LIMIT, "8-wide BVH needs %u traversal stack slots (%d)", c->bvhStackNeed + 1, RT_STACK);
		CU(c->items.upload(items.data(), items.size(), st, ub));
		CU(cudaEventRecord(c->evStop, st));
		CU(cudaStreamSynchronize(st));
		float ms = 0;
		cudaEventElapsedTime(&ms, b0, c->evStop);
		c->buildMs = ms;

		SceneDev &S = c->S;
		S.prim_geom = c->primGeom.p, S.prim_meta = c->primMeta.p, S.bvh_prims = c->bvhPrims.p;
		S.models = c->dModels.p, S.parts = c->dParts.p;
		S.tri_geom = c->triGeom.p, S.tri_norms = c->triNorms.p, S.tri_tcoords = c->triTcoords.p;
		S.tri_slot = c->triSlot.p, S.tri_part = c->triPart.p, S.nodes4 = c->nodes4.p, S.nodes8 = c->nodes8.p, S.items = c->items.p;
		S.n_items = (uint32_t)items.size(), S.n_prims = s->n_prims, S.n_tris = s->n_tris, S.n_parts = s->n_parts;
	}
	{
		// grid of the coherence binning: the bounded objects of the scene (planes are infinite; origins beyond the grid clamp)
		float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
		auto grow = [&](float x, float y, float z) { const float v[3] = { x, y, z }; for (int a = 0; a < 3; ++a) lo[a] = std::min(lo[a], v[a]), hi[a] = std::max(hi[a], v[a]); };
		for (const rt_model &M : c->models) { grow(M.ver_min.x + M.position.x, M.ver_min.y + M.position.y, M.ver_min.z + M.position.z); grow(M.ver_max.x + M.position.x, M.ver_max.y + M.position.y, M.ver_max.z + M.position.z); }
		for (const rt_prim &P : c->prims)
		{
			if (P.kind == RT_OBJ_SPHERE) { grow(P.position.x - P.radius, P.position.y - P.radius, P.position.z - P.radius); grow(P.position.x + P.radius, P.position.y + P.radius, P.position.z + P.radius); }
			else if (P.kind == RT_OBJ_CUBE) { grow(P.a.x + P.position.x, P.a.y + P.position.y, P.a.z + P.position.z); grow(P.b.x + P.position.x, P.b.y + P.position.y, P.b.z + P.position.z); }
		}
		c->binBits = c->binMode >= 0 ? (uint32_t)c->binMode : (c->anyRefract ? 3u : 0u);
		const uint32_t bits = c->binBits ? c->binBits : 1u;
		c->binGrid.bits = bits;
		{ const char *e = getenv("RT_B200_BIN_KEY"); c->binGrid.parts = e ? (uint32_t)atoi(e) : 7u; }
		for (int a = 0; a < 3; ++a)
		{
			const float ext = hi[a] > lo[a] ? hi[a] - lo[a] : 1.0f;
			c->binGrid.lo[a] = hi[a] >= lo[a] ? lo[a] : 0.0f;
			c->binGrid.scale[a] = (float)(1u << bits) / ext;
		}
	}
	c->S.materials = c->materials.p, c->S.textures = c->textures.p, c->S.texels = c->texels.p;
	if (c->uploadBytes)
	{
		CU(cudaEventRecord(c->evStop, st));
		CU(cudaStreamSynchronize(st));
		float ms = 0;
		cudaEventElapsedTime(&ms, c->evA, c->evStop);
		c->uploadMs = ms;
	}
	return RT_OK;
}

static int ensure_level(rt_ctx *c, uint32_t l, uint32_t cap, uint32_t lights)
{
	LevelStore &L = c->levels[l];
	if (c->binBits && l >= 1 && L.order.cap < std::max<size_t>(cap, L.capacity)) { CU(L.order.reserve(std::max<size_t>(cap, L.capacity))); CU(L.sort_key.reserve(std::max<size_t>(cap, L.capacity))); }
	if (cap <= L.capacity && lights <= L.lights) return RT_OK;
	L.capacity = 0;
	CU(L.ray_o.reserve(cap)); CU(L.ray_d.reserve(cap)); CU(L.ray_meta.reserve(cap)); CU(L.hit_p.reserve(cap));
	CU(L.hit_id.reserve(cap)); CU(L.color.reserve(cap)); CU(L.aux.reserve(cap)); CU(L.shadow.reserve((size_t)cap * (lights ? lights : 1))); CU(L.hit_list.reserve(cap)); CU(L.hit_n.reserve(cap)); CU(L.hit_uv.reserve(cap));
	if (c->binBits && l >= 1) { CU(L.order.reserve(cap)); CU(L.sort_key.reserve(cap)); }   // (reserve keeps a large enough buffer)
	// k_frame recognises a written slot by the epoch in ray_meta: fresh memory must not look written
	CU(cudaMemsetAsync(L.ray_meta.p, 0, sizeof(uint2) * L.ray_meta.cap, c->stream));
	CU(cudaMemsetAsync(L.hit_list.p, 0, sizeof(uint32_t) * L.hit_list.cap, c->stream));
	L.capacity = cap, L.lights = lights;
	return RT_OK;
}

static LevelBuf level_buf(const LevelStore &L)
{
	LevelBuf b;
	b.ray_o = L.ray_o.p, b.ray_d = L.ray_d.p, b.ray_meta = L.ray_meta.p, b.hit_p = L.hit_p.p, b.hit_id = L.hit_id.p;
	b.color = L.color.p, b.aux = L.aux.p, b.shadow = L.shadow.p, b.hit_list = L.hit_list.p, b.hit_n = L.hit_n.p, b.hit_uv = L.hit_uv.p, b.capacity = L.capacity;
	b.order = nullptr, b.sort_key = L.sort_key.p;   // order is switched on per frame by the wave scheduler
	return b;
}

static int finish_frame(rt_ctx *c);

static int render_frames(rt_ctx *c, const rt_render_params *p, uint32_t nFrames, const rt_camera *cams, void *const *outs);

extern "C" int rt_render_async(rt_ctx *c, const rt_render_params *p)
{
	return render_frames(c, p, 1, nullptr, nullptr);
}

extern "C" int rt_reserve_batch(rt_ctx *c, uint32_t n_frames)
{
	if (!c) return fail(RT_E_INVALID, "rt_reserve_batch: ctx is NULL");
	if (n_frames > RT_MAX_BATCH) return fail(RT_E_LIMIT, "rt_reserve_batch: %u frames (0..%d)", n_frames, RT_MAX_BATCH);
	c->reserveFrames = n_frames;
	return RT_OK;
}

extern "C" int rt_render_batch_async(rt_ctx *c, const rt_render_params *p, uint32_t n_frames, const rt_camera *cameras, void *const *device_outputs)
{
	if (n_frames < 1 || n_frames > RT_MAX_BATCH) return fail(RT_E_LIMIT, "rt_render_batch_async: %u frames (1..%d)", n_frames, RT_MAX_BATCH);
	return render_frames(c, p, n_frames, cameras, device_outputs);
}

static int render_frames(rt_ctx *c, const rt_render_params *p, uint32_t nFrames, const rt_camera *cams, void *const *outs)
{
	if (!c || !p) return fail(RT_E_INVALID, "rt_render_async: NULL argument");
	adopt_scene(c);
	if (!c->hasScene) return fail(RT_E_STATE, "rt_render_async: no scene uploaded");
	if (p->type != RT_TYPE_RAYTRACE && (p->type < RT_TYPE_CHECK || p->type > RT_TYPE_REFRACT))
		return fail(RT_E_INVALID, "rt_render_async: unknown render type 0x%x (RayTracer.h:5-13)", p->type);
	const bool debugStage = p->type >= RT_TYPE_CHECK && p->type <= RT_TYPE_SHADOW;
	if (nFrames > 1 && (debugStage || (p->flags & RT_FLAG_HIT_IDS)))
		return fail(RT_E_INVALID, "rt_render_batch_async: the staged debug shaders and RT_FLAG_HIT_IDS render one frame at a time");
	const uint32_t maxLevel = debugStage ? 0u : p->max_level;   // the staged shaders shade one level only
	if (maxLevel >= RT_MAX_LEVELS) return fail(RT_E_LIMIT, "rt_render_async: max_level %u (limit %d)", maxLevel, RT_MAX_LEVELS - 1);
	CU(cudaSetDevice(c->device));
	// Back-to-back frames: the next frame may be enqueued while the previous one is still running.
	// All that has to be over is the previous frame's H2D copies out of the pinned staging buffers
	// (evStart is recorded right after them).  Status and counters are those of the LAST frame only;
	// callers that want every frame checked call rt_wait between frames (RayTracer::start does).
	if (c->frameInFlight) CU(cudaEventSynchronize(c->evStart));
	cudaStream_t st = c->stream;
	const rt_camera &cam = c->camera;
	const int W = cam.width, H = cam.height;
	if (W <= 0 || H <= 0) return fail(RT_E_INVALID, "rt_render_async: camera is %dx%d", W, H);
	for (uint32_t f = 0; cams && f < nFrames; ++f)
		if (cams[f].width != W || cams[f].height != H || cams[f].fovy != cam.fovy || cams[f].zNear != cam.zNear || cams[f].zFar != cam.zFar)
			return fail(RT_E_INVALID, "rt_render_batch_async: camera %u differs from the uploaded camera in size, fovy or depth range (only position and orientation may vary inside a batch)", f);
	const uint32_t world = p->world > 1 ? p->world : 1, rank = p->world > 1 ? p->rank : 0;
	if (rank >= world) return fail(RT_E_INVALID, "rt_render_async: rank %u of world %u", rank, world);

	// frame constants (RayTracer.cpp:13-15)
	FrameParams &F = *c->hFrame;
	memset(&F, 0, sizeof F);
	F.cam_u = f4(cam.u), F.cam_v = f4(cam.v), F.cam_n = f4(cam.n), F.cam_pos = f4(cam.position);
	F.dp = tan(cam.fovy * 3.1415926535897932384 / 360) / (H / 2);
	F.zNear = cam.zNear, F.zFar = (float)(sqrt(2) * cam.zFar);
	F.width = W, F.height = H, F.blk_w = W / 64, F.blk_h = H / 64, F.half_w = W / 2, F.half_h = H / 2;
	F.max_level = maxLevel, F.type = p->type, F.rank = rank, F.world = world;
	const uint32_t tileRows = p->tile_rows ? p->tile_rows : 64u;
	if (tileRows != 8u && tileRows != 16u && tileRows != 32u && tileRows != 64u)
		return fail(RT_E_INVALID, "rt_render_async: tile_rows %u (must be 8, 16, 32 or 64)", tileRows);
	F.tile_rows = tileRows;
	{
		static const int dfs = []{ const char *e = getenv("RT_B200_CLAIM"); return (e && !strcmp(e, "dfs")) ? 1 : 0; }();
		static const int retire = []{ const char *e = getenv("RT_B200_RETIRE"); return (e && !atoi(e)) ? 0 : 1; }();
		F.sched_flags = (uint32_t)dfs | ((uint32_t)retire << 1);   // bit 2 (primary rays made inside k_frame) is set below once the scheduler is chosen
		F.sms = (uint32_t)c->sms;
		static const int retireRays = []{ const char *e = getenv("RT_B200_RETIRE_RAYS"); const int v = e ? atoi(e) : 16; return v > 0 ? v : 16; }();
		F.retire_rays = (uint32_t)retireRays;
		// RT_B200_KEEP_DIV=d: with d > 1 only every d-th CTA of k_frame is pinned until the frame is complete, the
		// others leave when the frame runs thin -- also when a pipeline's share is one CTA per SM (small shards,
		// many frames in flight), where "the first `sms` CTAs stay" pins the whole grid through the frame's tail
		static const int keepDiv = []{ const char *e = getenv("RT_B200_KEEP_DIV"); const int v = e ? atoi(e) : 1; return v > 1 ? v : 1; }();
		F.keep_div = (uint32_t)keepDiv, F.keep_salt = c->keepSalt;
	}
	F.serpentine = (p->flags & RT_FLAG_SERPENTINE) && world > 1 ? 1u : 0u;
	uint32_t bands = 0;
	while (shard_tile(bands, rank, world, F.serpentine) < (uint32_t)F.blk_h * 64u / tileRows) ++bands;
	// tile window (rt_render_params::tile_first / tile_count): a band of the shard's own tiles
	F.tile_first = p->tile_first < bands ? p->tile_first : bands;
	if (p->tile_count) bands = std::min(bands - F.tile_first, p->tile_count);
	else bands -= F.tile_first;
	F.n_rows = bands * tileRows;
	F.n_lights = (uint32_t)c->lights.size();
	F.env_light = f4(c->envLight);
	uint32_t enabledLights = 0;
	for (uint32_t k = 0; k < F.n_lights; ++k)
	{
		const rt_light &l = c->lights[k];
		DevLight &d = F.lights[k];
		d.position = f4(l.position), d.ambient = f4(l.ambient), d.diffuse = f4(l.diffuse), d.specular = f4(l.specular), d.attenuation = f4(l.attenuation);
		d.type = l.type, d.enabled = l.enabled;
		if (l.enabled) F.enabled_index[enabledLights++] = k;
	}
	F.n_enabled = enabledLights;
	const uint32_t nPixFrame = (uint32_t)F.blk_w * 64u * F.n_rows;
	if ((uint64_t)nPixFrame * nFrames > 0x7FFFFFFFull) return fail(RT_E_LIMIT, "rt_render_batch_async: %u frames of %u pixels", nFrames, nPixFrame);
	const uint32_t nPix = nPixFrame * nFrames;
	F.batch = nFrames, F.pix_per_frame = nPixFrame;

	// framebuffer: margins stay 127 (RayTracer.cpp:620)
	uint8_t *fb;
	if (nFrames > 1 || outs)
	{
		// a batch: one framebuffer per frame, the caller's or the library's (rt_read_batch_output)
		if (c->batchOut.size() < nFrames) c->batchOut.resize(nFrames);
		if (c->batchFill.size() < nFrames) c->batchFill.resize(nFrames, nullptr);
		if (!outs && c->reserveFrames > c->batchOut.size())
		{
			// the library-owned framebuffers of an announced batch size, all at once
			c->batchOut.resize(std::min<uint32_t>(c->reserveFrames, RT_MAX_BATCH)), c->batchFill.resize(c->batchOut.size(), nullptr);
			for (auto &b : c->batchOut) CU(b.reserve((size_t)W * H * 3));
		}
		const bool shardChanged = W != c->fillW || H != c->fillH || rank != c->fillRank || world != c->fillWorld || tileRows != c->fillTile || F.serpentine != c->fillSerp;
		for (uint32_t f = 0; f < nFrames; ++f)
		{
			uint8_t *o = outs ? (uint8_t *)outs[f] : nullptr;
			if (outs && !o) return fail(RT_E_INVALID, "rt_render_batch_async: device_outputs[%u] is NULL", f);
			if (!o) { CU(c->batchOut[f].reserve((size_t)W * H * 3)); o = c->batchOut[f].p; }
			if (shardChanged || c->batchFill[f] != o)
			{
				if (!is_landing_base(o)) CU(cudaMemsetAsync(o, 127, (size_t)W * H * 3, st));   // a landing buffer is born grey and shared with the peers' row copies
				c->batchFill[f] = o;
			}
			F.frames[f].out = o;
		}
		c->fillW = W, c->fillH = H, c->fillRank = rank, c->fillWorld = world, c->fillTile = tileRows, c->fillSerp = F.serpentine, c->fillPtr = nullptr;
		fb = F.frames[0].out;
		c->outW = W, c->outH = H, c->fb = fb;
	}
	else
	{
	if (c->extOut)
	{
		if (c->extOutBytes < (size_t)W * H * 3) return fail(RT_E_INVALID, "rt_render_async: external framebuffer holds %zu bytes, frame needs %zu", c->extOutBytes, (size_t)W * H * 3);
		fb = c->extOut;
	}
	else
	{
		CU(c->out.reserve((size_t)W * H * 3));
		fb = c->out.p;
	}
	c->outW = W, c->outH = H, c->fb = fb;
	// RayTracer.cpp:620 greys the whole buffer at every start(); the rendered region is overwritten by
	// every frame, so the fill is only repeated when the buffer, the frame size or the shard changes
	if (fb != c->fillPtr || W != c->fillW || H != c->fillH || rank != c->fillRank || world != c->fillWorld || tileRows != c->fillTile || F.serpentine != c->fillSerp)
	{
		c->fillSerp = F.serpentine;
		if (!is_landing_base(fb)) CU(cudaMemsetAsync(fb, 127, (size_t)W * H * 3, st));   // a landing buffer is born grey; its other rows belong to the peers' one-sided copies
		c->fillPtr = fb, c->fillW = W, c->fillH = H, c->fillRank = rank, c->fillWorld = world, c->fillTile = tileRows;
		for (auto &b : c->batchFill) b = nullptr;   // the shard the batch buffers were greyed for is no longer current
	}
	F.frames[0].out = fb;
	}
	for (uint32_t f = 0; f < nFrames; ++f)
	{
		const rt_camera &cf = cams ? cams[f] : cam;
		F.frames[f].cam_u = f4(cf.u), F.frames[f].cam_v = f4(cf.v), F.frames[f].cam_n = f4(cf.n), F.frames[f].cam_pos = f4(cf.position);
	}

	const bool refr = c->anyRefract && p->type != RT_TYPE_REFLECT;
	// (rt_reserve_batch: queues sized for the largest launch the caller announced, so that launches that grow from one frame
	// to a full batch do not re-allocate -- cudaFree + cudaMalloc of GBs, tens of ms -- at every new size)
	const uint32_t capPix = (uint32_t)std::min<uint64_t>((uint64_t)nPixFrame * std::max(nFrames, std::min<uint32_t>(c->reserveFrames, RT_MAX_BATCH)), 0x7FFFFFFFull);
	for (uint32_t l = 0; l <= maxLevel; ++l)
	{
		uint32_t cap = l == 0 || !refr ? capPix : (uint32_t)std::min<double>((double)capPix * c->levelFactor, 4.0e9);
		if (l > 0 && c->minCap[l] > cap) cap = c->minCap[l];
		int rc = ensure_level(c, l, cap ? cap : 1, F.n_lights);
		if (rc != RT_OK) return rc;
	}
	{ int rc = ensure_level(c, maxLevel + 1, 1, 1); if (rc != RT_OK) return rc; }

	// Scheduler choice (measured): the whole-frame persistent kernel wins when the
	// per-level queues are short next to the chip and rays are long and divergent (triangle meshes at
	// <= ~3 M pixels per GPU: -13 % on c3, and it is what lets a 1080p frame scale over GPUs); the
	// per-level waves win on big frames and on cheap analytic-primitive rays, where the scheduler's
	// atomics and polling cost more than the tails they hide.
	c->frameSched = c->schedMode == 1 || (c->schedMode == 0 && !c->models.empty() && c->nTris > 0 && nPix <= 3000000u);
	if (++c->frameEpoch > 65535u)
	{
		// epoch wrap: forget every stamp so a slot written 65535 frames ago cannot look fresh
		for (LevelStore &L : c->levels)   // every allocated level, also those deeper than this frame's max_level
			if (L.ray_meta.p) CU(cudaMemsetAsync(L.ray_meta.p, 0, sizeof(uint2) * L.ray_meta.cap, st));
		c->frameEpoch = 1;
	}
	F.epoch = c->frameEpoch;
	// level-0 rays are made inside the traversal kernels (k_frame always; k_wave(0) for the ray-traced types --
	// the staged debug shaders keep k_raygen, k_debug reads the stored rays)
	const bool genPrimary = (c->frameSched || (c->waveGen && !debugStage)) && p->type != RT_TYPE_CHECK;
	if (genPrimary) F.sched_flags |= 4u;
	WaveState &Wv = *c->hWaveInit;
	memset(&Wv, 0, sizeof Wv);
	Wv.count[0] = nPix;
	Wv.outstanding = (int)nPix;
	CU(cudaMemcpyAsync(c->dFrame, c->hFrame, sizeof(FrameParams), cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(c->dWave, c->hWaveInit, sizeof(WaveState), cudaMemcpyHostToDevice, st));
	c->frameH2D = sizeof(FrameParams) + sizeof(WaveState), c->frameD2H = sizeof(WaveState);
	g_h2dTotal += c->frameH2D, g_d2hTotal += c->frameD2H;
	CU(cudaEventRecord(c->evStart, st));

	const bool stats = (p->flags & RT_FLAG_STATS) != 0;
	c->S.brute = (p->flags & RT_FLAG_BRUTE) ? 1u : 0u;
	{
		static const uint32_t trig = []{ const char *e = getenv("RT_B200_DQ"); int v = e ? atoi(e) : 4; return (uint32_t)(v < 1 ? 1 : (v > 64 ? 64 : v)); }();
		c->S.dq_trigger = trig;
	}
	uint32_t launches = 0;
	if (nPix)
	{
		if (!genPrimary) { rtk_raygen(st, c->dFrame, level_buf(c->levels[0]), nPix, c->sms); ++launches; }
		// wave l = closest hit of level l (+ spawn of level l+1) fused with the shadow rays of level l-1
		if (c->stageTiming) CU(cudaEventRecord(c->evStage[0], st));
		LevelSet LS;
		for (uint32_t l = 0; l <= maxLevel + 1; ++l) LS.l[l] = level_buf(c->levels[l]);
		if (c->frameSched && p->type != RT_TYPE_CHECK)
		{
			rtk_frame(st, c->S, c->dFrame, LS, c->dWave, nPix, c->sms, stats, c->ctasPerSm); ++launches;
		}
		for (uint32_t l = 0; l <= maxLevel + 1 && p->type != RT_TYPE_CHECK && !c->frameSched; ++l)
		{
			// staged shaders: only RTshd (type 6) shoots shadow rays
			const bool traceOn = l <= maxLevel, shadowOn = l >= 1 && enabledLights > 0 && (!debugStage || p->type == RT_TYPE_SHADOW);
			if (!traceOn && !shadowOn) continue;
			const float zNear = l == 0 ? F.zNear : 0.0f;
			uint32_t items = traceOn ? c->levels[l].capacity : 0;
			if (shadowOn) items = std::max(items, (uint32_t)std::min<uint64_t>((uint64_t)c->levels[l - 1].capacity * enabledLights, 0x7FFFFFFFu));
			const int walk = c->travWalk;
			if (c->binBits && traceOn && l >= 1 && walk != 2 && !debugStage && c->levels[l].order.p)
			{
				// coherence binning of this level's rays; k_wave(l) then fetches them through `order`
				if (!c->binHist.p) CU(c->binHist.reserve((size_t)1u << (3u * 6u + 4u)));
				LS.l[l].order = c->levels[l].order.p;
				rtk_bin_rays(st, LS.l[l], c->dWave, l, c->binGrid, c->binHist.p, c->sms);
				launches += 3;
			}
			rtk_wave(st, c->S, c->dFrame, LS.l[l], LS.l[traceOn ? l + 1 : l], LS.l[l ? l - 1 : 0], c->dWave, l, traceOn, shadowOn, zNear, items, c->sms, stats, c->ctasPerSm, walk);
			++launches;
		}
		if (c->stageTiming) CU(cudaEventRecord(c->evStage[1], st));
		if (debugStage)
		{
			rtk_debug(st, c->S, c->dFrame, LS.l[0], nPix, fb, c->sms); ++launches;
		}
		else
		{
			// k_shade reads every hit_list entry exactly once and zeroes it for the next frame
			rtk_shade(st, c->S, c->dFrame, LS, c->dWave, maxLevel + 1, c->levels[0].capacity, c->sms, true); ++launches;
		}
		if (c->stageTiming) CU(cudaEventRecord(c->evStage[2], st));
		// One tree-walk launch when every ray tree is a chain (no refraction: one child per node); with
		// refraction the trees are binary and up to max_level deep, and one pass per level is faster
		// (C4: 0.5 ms against 2.7 ms for the walk).
		const bool combineByLevel = (p->flags & RT_FLAG_COMBINE_LEVELS) != 0 || refr;
		if (!debugStage && !combineByLevel)
		{
			rtk_resolve(st, c->S, c->dFrame, LS, c->dWave, fb, nPix, c->sms); ++launches;
		}
		for (int l = (int)maxLevel; l >= 0 && !debugStage && combineByLevel; --l)
		{
			rtk_combine(st, c->S, c->dFrame, level_buf(c->levels[l]), level_buf(c->levels[l + 1]), c->dWave, (uint32_t)l, fb, c->levels[l].capacity, c->sms);
			++launches;
		}
		if (p->type != RT_TYPE_CHECK && debugStage)
		{
			// leave every hit_list zeroed for the next frame (k_frame reads 0 as "not published yet")
			rtk_reset_hits(st, LS, c->dWave, maxLevel + 1, c->levels[0].capacity, c->sms); ++launches;
		}
	}
	CU(cudaGetLastError());
	CU(cudaEventRecord(c->evStop, st));
	CU(cudaMemcpyAsync(c->hWave, c->dWave, sizeof(WaveState), cudaMemcpyDeviceToHost, st));
	CU(cudaEventRecord(c->evB, st));
	if (cams && cams != c->lastCams.data()) c->lastCams.assign(cams, cams + nFrames);
	c->lastCamsGiven = cams != nullptr, c->lastOutsGiven = outs != nullptr;
	c->lastTileFirst = F.tile_first, c->ssValid = false;
	c->lastParams = *p, c->lastPixels = nPix, c->lastLaunches = launches, c->lastMaxLevel = maxLevel, c->lastBatch = nFrames;
	for (uint32_t f = 0; f < nFrames; ++f) c->lastOuts[f] = F.frames[f].out;
	c->frameInFlight = true, c->frameValid = false;
	return RT_OK;
}

// Wait for an event without monopolising a core: a few polls, then poll + yield.  RayTracer keeps one monitor
// thread per frame in flight (RayTracer.cpp:674-695 has one too); with 8 ranks x 8 frames on a 32-core
// host cudaEventSynchronize's busy wait starved the threads that enqueue the next frames.
static cudaError_t wait_event(cudaEvent_t ev)
{
	// a short spin for frames that are about to finish, then sleep in the driver: evB / evRead are created with
	// cudaEventBlockingSync, so cudaEventSynchronize parks the thread instead of burning a core per frame in flight
	// (8 ranks x (3 batch workers + monitor threads) on a 16-32 core host starved the threads that enqueue frames)
	for (unsigned spins = 0; spins < 64u; ++spins)
	{
		const cudaError_t e = cudaEventQuery(ev);
		if (e != cudaErrorNotReady) return e;
	}
	return cudaEventSynchronize(ev);
}

static int finish_frame(rt_ctx *c)
{
	CU(cudaSetDevice(c->device));
	if (c->regrowTries == 0) c->regrewLastFrame = false;
	CU(wait_event(c->evB));
	float ms = 0;
	CU(cudaEventElapsedTime(&ms, c->evStart, c->evStop));
	c->renderMs = ms;
	c->frameInFlight = false;
	c->traceMs = c->shadowMs = c->shadeMs = c->otherMs = 0;
	if (c->stageTiming && c->lastPixels)
	{
		// the closest-hit and shadow queries share the fused wave kernels: reported together as trace_ms
		float a = 0, d = 0;
		cudaEventElapsedTime(&a, c->evStage[0], c->evStage[1]);
		cudaEventElapsedTime(&d, c->evStage[1], c->evStage[2]);
		c->traceMs = a, c->shadeMs = d;
		c->otherMs = c->renderMs - c->traceMs - c->shadeMs;
	}
	if (c->hWave->stop_epoch == c->frameEpoch)
	{
		c->regrowTries = 0;
		return RT_OK;   // stopped by rt_stop: incomplete frame, not an error (frameValid stays false)
	}
	if (c->hWave->overflow == 2u)
		return fail(RT_E_STATE, "the frame scheduler stopped making progress (k_frame gave up waiting); frame is incomplete");
	if (c->hWave->overflow)
	{
		// A ray level ran out of slots: with refraction a ray tree can hold 2^l rays at level l, the queues are sized
		// for levelFactor x pixels.  WaveState::count[l] kept counting past the capacity, so it says what level l needs;
		// the levels below an overflowing one were starved, they are grown by the same ratio.  Then the frame is
		// rendered again, transparently (the caller sees one longer frame).
		if (c->regrowTries >= RT_MAX_LEVELS + 2)
			return fail(RT_E_LIMIT, "a ray level still overflows its queue after %u regrows", c->regrowTries);
		double ratio = 1.0;
		for (uint32_t l = 1; l <= c->lastMaxLevel; ++l)
		{
			const uint64_t cap = c->levels[l].capacity, need = c->hWave->count[l];
			if (need > cap) ratio = std::max(ratio, (double)need / (double)std::max<uint64_t>(cap, 1));
			const uint64_t want = std::max<uint64_t>(need + need / 8 + 1024, (uint64_t)((double)cap * ratio) + 1024);
			if (need > cap || ratio > 1.0) c->minCap[l] = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(c->minCap[l], want), 0xFFFFFFF0ull);
		}
		++c->regrowTries;
		c->regrewLastFrame = true;
		const rt_render_params p = c->lastParams;
		const uint32_t nb = c->lastBatch;
		int rc = render_frames(c, &p, nb, c->lastCamsGiven ? c->lastCams.data() : nullptr, c->lastOutsGiven ? (void *const *)c->lastOuts : nullptr);
		if (rc != RT_OK) return rc;
		return finish_frame(c);
	}
	c->regrowTries = 0;
	c->frameValid = true;
	return RT_OK;
}

// Pixels (x samples) one band launch may hold: the wave kernels run best with millions of rays per queue, and the
// ray-level buffers of a launch are ~140 bytes per slot and level.
static uint32_t supersample_budget()
{
	const char *e = getenv("RT_B200_SS_PIXELS");   // read per frame: tests shrink it to force several bands
	const long n = e ? atol(e) : 8000000L;
	return (uint32_t)(n > 4096 ? n : 4096);
}

extern "C" int rt_render_supersampled(rt_ctx *c, const rt_render_params *p, uint32_t n_samples, const rt_camera *cams)
{
	if (!c || !p || !cams) return fail(RT_E_INVALID, "rt_render_supersampled: NULL argument");
	if (n_samples < 1 || n_samples > RT_MAX_BATCH) return fail(RT_E_LIMIT, "rt_render_supersampled: %u samples per pixel (1..%d)", n_samples, RT_MAX_BATCH);
	if (p->type != RT_TYPE_RAYTRACE && p->type != RT_TYPE_REFLECT && p->type != RT_TYPE_REFRACT)
		return fail(RT_E_INVALID, "rt_render_supersampled: only the ray-traced types are supersampled");
	if (p->flags & RT_FLAG_HIT_IDS) return fail(RT_E_INVALID, "rt_render_supersampled: RT_FLAG_HIT_IDS needs a single sample");
	if (p->tile_first || p->tile_count) return fail(RT_E_INVALID, "rt_render_supersampled: the tile window is used internally; pass 0");
	adopt_scene(c);
	if (!c->hasScene) return fail(RT_E_STATE, "rt_render_supersampled: no scene uploaded");
	CU(cudaSetDevice(c->device));
	if (c->frameInFlight) { int rc = finish_frame(c); if (rc != RT_OK) return rc; }
	const int W = c->camera.width, H = c->camera.height;
	if (W <= 0 || H <= 0) return fail(RT_E_INVALID, "rt_render_supersampled: camera is %dx%d", W, H);
	const uint32_t world = p->world > 1 ? p->world : 1, rank = p->world > 1 ? p->rank : 0;
	if (rank >= world) return fail(RT_E_INVALID, "rt_render_supersampled: rank %u of world %u", rank, world);
	const uint32_t tileRows = p->tile_rows ? p->tile_rows : 64u;
	if (tileRows != 8u && tileRows != 16u && tileRows != 32u && tileRows != 64u)
		return fail(RT_E_INVALID, "rt_render_supersampled: tile_rows %u (must be 8, 16, 32 or 64)", tileRows);
	const uint32_t serp = (p->flags & RT_FLAG_SERPENTINE) && world > 1 ? 1u : 0u;
	uint32_t mine = 0;
	while (shard_tile(mine, rank, world, serp) < (uint32_t)(H / 64) * 64u / tileRows) ++mine;
	// target framebuffer: the caller's (rt_set_output) or the library's, greyed once like a frame's
	uint8_t *fb;
	if (c->extOut)
	{
		if (c->extOutBytes < (size_t)W * H * 3) return fail(RT_E_INVALID, "rt_render_supersampled: external framebuffer holds %zu bytes, frame needs %zu", c->extOutBytes, (size_t)W * H * 3);
		fb = c->extOut;
	}
	else
	{
		CU(c->out.reserve((size_t)W * H * 3));
		fb = c->out.p;
	}
	const uint64_t shardKey = ((uint64_t)rank << 40) | ((uint64_t)world << 16) | ((uint64_t)tileRows << 4) | serp;
	if (fb != c->ssFillPtr || W != c->ssFillW || H != c->ssFillH || shardKey != c->ssFillShard)
	{
		// RayTracer.cpp:620 greys the whole buffer at every start(); the shard's rows are overwritten by every frame, so
		// the fill is only repeated when the buffer, the frame size or the shard changes
		if (!is_landing_base(fb)) CU(cudaMemsetAsync(fb, 127, (size_t)W * H * 3, c->stream));
		c->ssFillPtr = fb, c->ssFillW = W, c->ssFillH = H, c->ssFillShard = shardKey;
	}
	c->fillPtr = nullptr;   // a later single frame into the same buffer greys it again (other rows may hold averaged pixels)
	// band = as many of the shard's tiles as fit the budget with n_samples frames in the batch
	const uint64_t tilePix = (uint64_t)(W / 64) * 64u * tileRows;
	uint32_t tilesPerBand = (uint32_t)std::max<uint64_t>(1, supersample_budget() / std::max<uint64_t>(1, tilePix * n_samples));
	rt_counters tot;
	memset(&tot, 0, sizeof tot);
	rt_render_params bp = *p;
	for (uint32_t k0 = 0; k0 < mine; k0 += tilesPerBand)
	{
		const uint32_t kc = std::min(tilesPerBand, mine - k0);
		bp.tile_first = k0, bp.tile_count = kc;
		int rc = render_frames(c, &bp, n_samples, cams, nullptr);      // the samples of the band: one batch, library-owned sample frames
		if (rc != RT_OK) return rc;
		rtk_average(c->stream, c->lastOuts, n_samples, fb, W, tileRows, k0, kc, rank, world, serp, c->sms);
		CU(cudaGetLastError());
		rc = finish_frame(c);                                         // counters of this band (and a regrow + re-render if a level overflowed)
		if (rc != RT_OK) return rc;
		if (c->regrewLastFrame)
		{
			// the band was rendered again after the averaging kernel had been enqueued: average the final sample frames
			rtk_average(c->stream, c->lastOuts, n_samples, fb, W, tileRows, k0, kc, rank, world, serp, c->sms);
			CU(cudaGetLastError());
		}
		if (!c->frameValid) break;                                    // rt_stop
		rt_counters bc;
		c->ssValid = false;
		rc = rt_read_counters(c, &bc);
		if (rc != RT_OK) return rc;
		tot.primary += bc.primary, tot.shadow += bc.shadow, tot.reflect += bc.reflect, tot.refract += bc.refract;
		tot.nodes_visited += bc.nodes_visited, tot.tri_tests += bc.tri_tests, tot.prim_tests += bc.prim_tests;
		tot.render_ms += bc.render_ms, tot.trace_ms += bc.trace_ms, tot.shade_ms += bc.shade_ms, tot.other_ms += bc.other_ms;
		tot.launches += bc.launches + 1;
	}
	CU(cudaStreamSynchronize(c->stream));
	c->ssTotals = tot, c->ssValid = c->frameValid;
	c->lastParams = *p, c->lastTileFirst = 0, c->lastBatch = 1;   // read-backs and row pushes see the whole shard in `fb`
	c->lastOuts[0] = fb, c->fb = fb, c->outW = W, c->outH = H;
	c->renderMs = tot.render_ms;
	return RT_OK;
}

extern "C" int rt_poll(rt_ctx *c, int *done, double *seconds)
{
	if (!c || !done) return fail(RT_E_INVALID, "rt_poll: NULL argument");
	if (!c->frameInFlight) { *done = 1; if (seconds) *seconds = c->renderMs * 1e-3; return RT_OK; }
	cudaSetDevice(c->device);
	cudaError_t e = cudaEventQuery(c->evB);
	if (e == cudaErrorNotReady) { *done = 0; return RT_OK; }
	if (e != cudaSuccess) return fail(RT_E_CUDA, "rt_poll: %s", cudaGetErrorString(e));
	int rc = finish_frame(c);
	*done = 1;
	if (seconds) *seconds = c->renderMs * 1e-3;
	return rc;
}

extern "C" int rt_wait(rt_ctx *c, double *seconds)
{
	if (!c) return fail(RT_E_INVALID, "rt_wait: ctx is NULL");
	int rc = RT_OK;
	if (c->frameInFlight) rc = finish_frame(c);
	if (seconds) *seconds = c->renderMs * 1e-3;
	return rc;
}

// Cooperative cancel, like RayTracer::stop clearing isRun (RayTracer.cpp:698-701): a side stream
// writes the running frame's epoch into WaveState::stop_epoch; the traversal warps look at that word every
// time they fetch work and leave when it names THEIR frame, so the frame ends within one batch.  A stop that
// lands late -- after the next frame's state was initialised -- names an epoch that is no longer running and
// cancels nothing.  The cancelled frame is incomplete (rt_wait still returns RT_OK, frameValid stays false; the
// pixels not reached keep whatever k_combine wrote from partial data).
extern "C" int rt_stop(rt_ctx *c)
{
	if (!c) return fail(RT_E_INVALID, "rt_stop: ctx is NULL");
	if (!c->frameInFlight) return RT_OK;
	CU(cudaSetDevice(c->device));
	uint32_t *src = &c->hStopWord[c->stopSlot++ & 63u];
	*src = c->frameEpoch;
	CU(cudaMemcpyAsync(&c->dWave->stop_epoch, src, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stopStream));
	return RT_OK;
}

extern "C" int rt_read_output(rt_ctx *c, uint8_t *rgb, size_t stride)
{
	if (!c || !rgb) return fail(RT_E_INVALID, "rt_read_output: NULL argument");
	if (c->frameInFlight) { int rc = finish_frame(c); if (rc != RT_OK) return rc; }
	if (!c->fb) return fail(RT_E_STATE, "rt_read_output: nothing rendered yet");
	CU(cudaSetDevice(c->device));
	const size_t row = (size_t)c->outW * 3;
	if (stride < row) return fail(RT_E_INVALID, "rt_read_output: stride %zu < %zu", stride, row);
	CU(cudaMemcpy2DAsync(rgb, stride, c->fb, row, row, (size_t)c->outH, cudaMemcpyDeviceToHost, c->stream));
	c->frameD2H += row * (size_t)c->outH;
	g_d2hTotal += row * (size_t)c->outH;
	CU(cudaEventRecord(c->evRead, c->stream));
	CU(wait_event(c->evRead));
	return RT_OK;
}

// The rows the last frame's shard rendered, device framebuffer -> `dst` (a full-frame buffer with `stride`
// bytes per row), as strided 2-D copies on the copy engines: the tiles of a rank are `world` tiles apart
// (one copy), with RT_FLAG_SERPENTINE the even and the odd tile groups are each 2*world tiles apart (two).
// A "row" of such a copy is one tile (tile_rows image rows) when dst is as dense as the framebuffer.
static int copy_shard_rows(rt_ctx *c, const uint8_t *src, uint8_t *dst, size_t stride, cudaMemcpyKind kind, cudaStream_t st, size_t *bytesOut)
{
	const rt_render_params &p = c->lastParams;
	const uint32_t world = p.world > 1 ? p.world : 1, rank = p.world > 1 ? p.rank : 0;
	const uint32_t serp = (p.flags & RT_FLAG_SERPENTINE) && world > 1 ? 1u : 0u;
	const uint32_t tileRows = p.tile_rows ? p.tile_rows : 64u;
	const size_t row = (size_t)c->outW * 3;
	const uint32_t tiles = (uint32_t)(c->outH / 64) * 64u / tileRows;
	uint32_t mine = 0;
	while (shard_tile(mine, rank, world, serp) < tiles) ++mine;
	// the tile window of the last frame (rt_render_params::tile_first / tile_count): only those tiles were rendered
	const uint32_t k0 = p.tile_first < mine ? p.tile_first : mine;
	const uint32_t k1 = p.tile_count ? std::min(mine, k0 + p.tile_count) : mine;
	size_t bytes = 0;
	// family f: tiles k = f, f + step, f + 2*step ... of this shard are equally spaced in the frame
	const uint32_t step = serp ? 2u : 1u;
	for (uint32_t f = 0; f < step; ++f)
	{
		const uint32_t kf = k0 + ((f + step - k0 % step) % step);      // first tile >= k0 of this family
		if (kf >= k1) continue;
		const uint32_t count = (k1 - kf + step - 1u) / step;
		const size_t first = (size_t)shard_tile(kf, rank, world, serp) * tileRows;       // first image row of the family
		const size_t pitchRows = (size_t)step * world * tileRows;                        // image rows between two of its tiles
		if (stride == row)
			CU(cudaMemcpy2DAsync(dst + first * row, pitchRows * row, src + first * row, pitchRows * row, tileRows * row, count, kind, st));
		else
			for (uint32_t t = 0; t < count; ++t)
				CU(cudaMemcpy2DAsync(dst + (first + t * pitchRows) * stride, stride, src + (first + t * pitchRows) * row, row, row, tileRows, kind, st));
		bytes += (size_t)count * tileRows * row;
	}
	if (bytesOut) *bytesOut = bytes;
	return RT_OK;
}

extern "C" int rt_read_output_rows(rt_ctx *c, uint8_t *rgb, size_t stride)
{
	if (!c || !rgb) return fail(RT_E_INVALID, "rt_read_output_rows: NULL argument");
	if (c->lastParams.world <= 1 && !c->lastParams.tile_first && !c->lastParams.tile_count) return rt_read_output(c, rgb, stride);
	if (c->frameInFlight) { int rc = finish_frame(c); if (rc != RT_OK) return rc; }
	if (!c->fb) return fail(RT_E_STATE, "rt_read_output_rows: nothing rendered yet");
	CU(cudaSetDevice(c->device));
	if (stride < (size_t)c->outW * 3) return fail(RT_E_INVALID, "rt_read_output_rows: stride %zu < %zu", stride, (size_t)c->outW * 3);
	size_t bytes = 0;
	int rc = copy_shard_rows(c, c->fb, rgb, stride, cudaMemcpyDeviceToHost, c->stream, &bytes);
	if (rc != RT_OK) return rc;
	c->frameD2H += bytes;
	g_d2hTotal += bytes;
	CU(cudaEventRecord(c->evRead, c->stream));
	CU(wait_event(c->evRead));
	return RT_OK;
}

extern "C" int rt_read_batch_output(rt_ctx *c, uint32_t frame, uint8_t *rgb, size_t stride, int rows_only)
{
	if (!c || !rgb) return fail(RT_E_INVALID, "rt_read_batch_output: NULL argument");
	if (c->frameInFlight) { int rc = finish_frame(c); if (rc != RT_OK) return rc; }
	if (!c->fb) return fail(RT_E_STATE, "rt_read_batch_output: nothing rendered yet");
	if (frame >= c->lastBatch) return fail(RT_E_INVALID, "rt_read_batch_output: frame %u of a batch of %u", frame, c->lastBatch);
	CU(cudaSetDevice(c->device));
	const size_t row = (size_t)c->outW * 3;
	if (stride < row) return fail(RT_E_INVALID, "rt_read_batch_output: stride %zu < %zu", stride, row);
	size_t bytes = row * (size_t)c->outH;
	if ((rows_only & 1) && c->lastParams.world > 1)
	{
		int rc = copy_shard_rows(c, c->lastOuts[frame], rgb, stride, cudaMemcpyDeviceToHost, c->stream, &bytes);
		if (rc != RT_OK) return rc;
	}
	else
		CU(cudaMemcpy2DAsync(rgb, stride, c->lastOuts[frame], row, row, (size_t)c->outH, cudaMemcpyDeviceToHost, c->stream));
	c->frameD2H += bytes;
	g_d2hTotal += bytes;
	if (rows_only & 2)
		return RT_OK;   // enqueued only: a later call without this bit (same stream, in order) completes them all
	CU(cudaEventRecord(c->evRead, c->stream));
	CU(wait_event(c->evRead));
	return RT_OK;
}

extern "C" int rt_output_device(rt_ctx *c, void **ptr, size_t *bytes)
{
	if (!c || !ptr) return fail(RT_E_INVALID, "rt_output_device: NULL argument");
	if (!c->fb) return fail(RT_E_STATE, "rt_output_device: nothing rendered yet");
	*ptr = c->fb;
	if (bytes) *bytes = (size_t)c->outW * c->outH * 3;
	return RT_OK;
}

// ---- NVLink frame gather without SM work: one-sided put with signal ----------------------------------
// The traversal kernels are persistent and fill every SM, so a gather that needs kernels of its own
// (pack, NCCL send/recv, unpack) queues behind them.  Here the destination rank exposes a frame-sized
// "landing" buffer over CUDA IPC; every rank copies its row tiles straight to their final offsets in
// it with ONE strided peer-to-peer copy on the copy engines (tiles of one rank are `world` tiles
// apart in both frames), then writes the frame's sequence number into its flag word behind the data;
// the destination waits for the flags with a stream wait-value.  No pack, no unpack, no SM.
struct rt_landing
{
	uint8_t *base = nullptr;      // frame bytes, then 64 flag words (one per rank)
	size_t frameBytes = 0;
	int width = 0, height = 0, device = 0;
	bool owner = false;
	uint64_t *hSeq = nullptr;     // pinned ring: source of the flag copies when stream mem-ops are unavailable
	uint64_t pushedSeq = 0;       // the frame this handle pushed last: the next push waits until it has been released
	bool acks = false;            // owner: rt_landing_release has been called at least once (the ack word is live)
};
#define RT_LANDING_FLAGS 64
#define RT_LANDING_ACK (RT_LANDING_FLAGS - 1)   // last flag word: highest frame the consumer has released (RT_LANDING_UNARMED until armed)
#define RT_LANDING_UNARMED 0x3FFFFFFFFFFFFFFFull
#define RT_LANDING_RING 4096

typedef int (*StreamValue64Fn)(cudaStream_t, unsigned long long, unsigned long long, unsigned int);
static StreamValue64Fn driver_fn(const char *name)
{
	void *fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
	{
		cudaGetLastError();
		return nullptr;
	}
	return (StreamValue64Fn)fn;
}

extern "C" int rt_landing_create(rt_ctx *c, int width, int height, rt_landing **out, void *ipc_handle64)
{
	if (!c || !out || !ipc_handle64 || width <= 0 || height <= 0) return fail(RT_E_INVALID, "rt_landing_create: bad argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	CU(cudaSetDevice(c->device));
	rt_landing *L = new rt_landing();
	L->frameBytes = (size_t)width * height * 3, L->width = width, L->height = height, L->device = c->device, L->owner = true;
	const size_t total = ((L->frameBytes + 255) & ~(size_t)255) + RT_LANDING_FLAGS * sizeof(uint64_t);
	if (cudaMalloc(&L->base, total) != cudaSuccess) { delete L; return fail(RT_E_CUDA, "rt_landing_create: cudaMalloc of %zu bytes failed", total); }
	CU(cudaMemset(L->base, 127, L->frameBytes));
	CU(cudaMemset(L->base + ((L->frameBytes + 255) & ~(size_t)255), 0, RT_LANDING_FLAGS * sizeof(uint64_t)));
	{
		// "everything released" for a buffer that is never armed.  cuStreamWaitValue64(GEQ) compares CYCLICALLY,
		// (int64_t)(*addr - value) >= 0, so the marker is the largest value that is >= every sequence number in that sense
		const uint64_t unarmed = RT_LANDING_UNARMED;
		CU(cudaMemcpy(L->base + ((L->frameBytes + 255) & ~(size_t)255) + RT_LANDING_ACK * sizeof(uint64_t), &unarmed, sizeof unarmed, cudaMemcpyHostToDevice));
	}
	CU(cudaHostAlloc(&L->hSeq, RT_LANDING_RING * sizeof(uint64_t), cudaHostAllocDefault));
	CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)ipc_handle64, L->base));
	{ std::lock_guard<std::mutex> lock(g_landingMutex); g_landingBases.push_back(L->base); }
	*out = L;
	return RT_OK;
}

extern "C" int rt_landing_open(rt_ctx *c, int width, int height, const void *ipc_handle64, rt_landing **out)
{
	if (!c || !out || !ipc_handle64 || width <= 0 || height <= 0) return fail(RT_E_INVALID, "rt_landing_open: bad argument");
	CU(cudaSetDevice(c->device));
	rt_landing *L = new rt_landing();
	L->frameBytes = (size_t)width * height * 3, L->width = width, L->height = height, L->device = c->device, L->owner = false;
	cudaIpcMemHandle_t h;
	memcpy(&h, ipc_handle64, sizeof h);
	void *p = nullptr;
	if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
	{
		const cudaError_t e = cudaGetLastError();
		delete L;
		return fail(RT_E_CUDA, "rt_landing_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
	}
	L->base = (uint8_t *)p;
	CU(cudaHostAlloc(&L->hSeq, RT_LANDING_RING * sizeof(uint64_t), cudaHostAllocDefault));
	*out = L;
	return RT_OK;
}

extern "C" void rt_landing_close(rt_landing *L)
{
	if (!L) return;
	cudaSetDevice(L->device);
	if (L->owner)
	{
		{ std::lock_guard<std::mutex> lock(g_landingMutex); for (auto &b : g_landingBases) if (b == L->base) { b = g_landingBases.back(); g_landingBases.pop_back(); break; } }
		cudaFree(L->base);
	}
	else cudaIpcCloseMemHandle(L->base);
	cudaFreeHost(L->hSeq);
	delete L;
}

extern "C" int rt_landing_ptr(rt_landing *L, void **device_ptr, size_t *bytes)
{
	if (!L || !device_ptr) return fail(RT_E_INVALID, "rt_landing_ptr: NULL argument");
	*device_ptr = L->base;
	if (bytes) *bytes = L->frameBytes;
	return RT_OK;
}

static uint64_t *landing_flags(rt_landing *L) { return (uint64_t *)(L->base + ((L->frameBytes + 255) & ~(size_t)255)); }

static int push_rows(rt_ctx *c, rt_landing *L, uint64_t seq, uint32_t frame);

extern "C" int rt_push_rows(rt_ctx *c, rt_landing *L, uint64_t seq) { return push_rows(c, L, seq, 0); }

extern "C" int rt_push_batch_rows(rt_ctx *c, uint32_t frame, rt_landing *L, uint64_t seq)
{
	if (c && frame >= c->lastBatch) return fail(RT_E_INVALID, "rt_push_batch_rows: frame %u of a batch of %u", frame, c->lastBatch);
	return push_rows(c, L, seq, frame);
}

static int push_rows(rt_ctx *c, rt_landing *L, uint64_t seq, uint32_t frame)
{
	if (!c || !L) return fail(RT_E_INVALID, "rt_push_rows: NULL argument");
	if (!c->fb) return fail(RT_E_STATE, "rt_push_rows: nothing rendered yet");
	if (c->outW != L->width || c->outH != L->height) return fail(RT_E_INVALID, "rt_push_rows: frame is %dx%d, landing buffer %dx%d", c->outW, c->outH, L->width, L->height);
	CU(cudaSetDevice(c->device));
	const rt_render_params &p = c->lastParams;
	const uint32_t world = p.world > 1 ? p.world : 1, rank = p.world > 1 ? p.rank : 0;
	if (rank >= RT_LANDING_ACK) return fail(RT_E_LIMIT, "rt_push_rows: rank %u (landing buffers hold %d flags)", rank, RT_LANDING_ACK);
	cudaStream_t st = c->stream;
	const uint8_t *src = c->lastBatch > 1 || frame ? c->lastOuts[frame] : c->fb;
	if (src != L->base)
	{
		// back-pressure: the frame this rank pushed into the buffer before must have been released by its consumer
		// (cyclic >= against the ack word; RT_LANDING_UNARMED = the consumer does not release, nothing to wait for)
		if (L->pushedSeq)
		{
			static const StreamValue64Fn waitAck = driver_fn("cuStreamWaitValue64");
			uint64_t *ack = landing_flags(L) + RT_LANDING_ACK;
			if (!waitAck || waitAck(st, (unsigned long long)(uintptr_t)ack, L->pushedSeq, 0u /* GEQ */) != 0)
			{
				CU(cudaStreamSynchronize(st));
				uint64_t v = 0;
				do CU(cudaMemcpy(&v, ack, sizeof v, cudaMemcpyDeviceToHost)); while ((int64_t)(v - L->pushedSeq) < 0);
			}
		}
		int rc = copy_shard_rows(c, src, L->base, (size_t)c->outW * 3, cudaMemcpyDeviceToDevice, st, nullptr);
		if (rc != RT_OK) return rc;
	}
	// the flag goes out behind the data on the same stream
	static const StreamValue64Fn writeValue = driver_fn("cuStreamWriteValue64");
	uint64_t *flag = landing_flags(L) + rank;
	if (!writeValue || writeValue(st, (unsigned long long)(uintptr_t)flag, seq, 0u) != 0)
	{
		{ static bool told = false; if (!told) { told = true; fprintf(stderr, "raytrace_b200: cuStreamWriteValue64 unavailable (%s), rt_push_rows signals with a copy\n", writeValue ? "call failed" : "no entry point"); } }
		uint64_t *src = &L->hSeq[seq % RT_LANDING_RING];
		*src = seq;
		CU(cudaMemcpyAsync(flag, src, sizeof(uint64_t), cudaMemcpyDefault, st));
	}
	L->pushedSeq = seq;
	return RT_OK;
}

extern "C" int rt_landing_release(rt_ctx *c, rt_landing *L, uint64_t seq, void *consumer_stream)
{
	if (!c || !L) return fail(RT_E_INVALID, "rt_landing_release: NULL argument");
	if (!L->owner) return fail(RT_E_INVALID, "rt_landing_release: only the rank that created the landing buffer releases it");
	CU(cudaSetDevice(c->device));
	cudaStream_t st = consumer_stream ? (cudaStream_t)consumer_stream : c->stream;
	static const StreamValue64Fn writeValue = driver_fn("cuStreamWriteValue64");
	uint64_t *ack = landing_flags(L) + RT_LANDING_ACK;
	if (seq == 0)
	{
		// arming (before the handle is shipped): from now on a push of frame k waits for the release of frame k - 1
		const uint64_t zero = 0;
		CU(cudaMemcpy(ack, &zero, sizeof zero, cudaMemcpyHostToDevice));
		L->acks = true;
		return RT_OK;
	}
	if (!writeValue || writeValue(st, (unsigned long long)(uintptr_t)ack, seq, 0u) != 0)
	{
		uint64_t *src = &L->hSeq[(seq + RT_LANDING_RING / 2) % RT_LANDING_RING];
		*src = seq;
		CU(cudaMemcpyAsync(ack, src, sizeof(uint64_t), cudaMemcpyDefault, st));
	}
	L->acks = true;
	return RT_OK;
}

extern "C" int rt_landing_wait(rt_ctx *c, rt_landing *L, uint64_t seq, uint32_t world, void *consumer_stream)
{
	if (!c || !L) return fail(RT_E_INVALID, "rt_landing_wait: NULL argument");
	if (world > RT_LANDING_FLAGS) return fail(RT_E_LIMIT, "rt_landing_wait: world %u", world);
	CU(cudaSetDevice(c->device));
	// the peers' rows never overlap the rows this rank renders, so the pipeline itself need not wait:
	// a consumer that reads the assembled frame on its own stream passes that stream here
	cudaStream_t st = consumer_stream ? (cudaStream_t)consumer_stream : c->stream;
	static const StreamValue64Fn waitValue = driver_fn("cuStreamWaitValue64");
	for (uint32_t r = 0; r < world; ++r)
	{
		uint64_t *flag = landing_flags(L) + r;
		// CU_STREAM_WAIT_VALUE_GEQ (0) | CU_STREAM_WAIT_VALUE_FLUSH (1 << 30): the peers' row data is visible once the flag is
		// (a peer's flag is written by a stream operation that starts only after its row copy has completed)
		if (waitValue && waitValue(st, (unsigned long long)(uintptr_t)flag, seq, 0u | (1u << 30)) == 0)
			continue;
		if (waitValue && waitValue(st, (unsigned long long)(uintptr_t)flag, seq, 0u) == 0)   // device cannot flush remote writes: plain >= wait
			continue;
		// no stream mem-ops: wait on the host (still no kernel)
		{ static bool told = false; if (!told) { told = true; fprintf(stderr, "raytrace_b200: cuStreamWaitValue64 unavailable (%s), rt_landing_wait falls back to a host wait\n", waitValue ? "call failed" : "no entry point"); } }
		CU(cudaStreamSynchronize(st));
		uint64_t v = 0;
		do CU(cudaMemcpy(&v, flag, sizeof v, cudaMemcpyDeviceToHost)); while (v < seq);
	}
	return RT_OK;
}

extern "C" int rt_host_alloc(void **ptr, size_t bytes)
{
	if (!ptr) return fail(RT_E_INVALID, "rt_host_alloc: ptr is NULL");
	CU(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
	return RT_OK;
}

extern "C" int rt_host_free(void *ptr)
{
	if (ptr) CU(cudaFreeHost(ptr));
	return RT_OK;
}

extern "C" int rt_set_output(rt_ctx *c, void *device_ptr, size_t bytes)
{
	if (c) c->fillPtr = nullptr;
	if (!c) return fail(RT_E_INVALID, "rt_set_output: ctx is NULL");
	if (c->frameInFlight) { int rc = finish_frame(c); if (rc != RT_OK) return rc; }
	c->extOut = (uint8_t *)device_ptr, c->extOutBytes = device_ptr ? bytes : 0;
	return RT_OK;
}

// identity <-> device id (RT_ID_NONE | prim flat index | RT_ID_TRI | oct<<28 | triangle)
static uint32_t encode_id(const rt_ctx *c, const rt_hit_id &id)
{
	if (id.object < 0) return RT_ID_NONE;
	for (const rt_model &M : c->models)
		if ((int32_t)M.object == id.object)
		{
			if (id.sub < 0 || (uint32_t)id.sub >= M.part_count || id.index < 0) return RT_ID_NONE;
			const rt_part &P = c->parts[M.part_begin + id.sub];
			return RT_ID_TRI | ((uint32_t)(id.octant & 7) << 28) | (P.tri_begin + (uint32_t)id.index);
		}
	for (size_t p = 0; p < c->prims.size(); ++p)
		if ((int32_t)c->prims[p].object == id.object && (int32_t)c->prims[p].sub == id.sub) return (uint32_t)p;
	return RT_ID_NONE;
}

static rt_hit_id decode_id(const rt_ctx *c, uint32_t h, float distance)
{
	rt_hit_id id = { -1, -1, -1, -1, distance };
	if (h == RT_ID_NONE) return id;
	if (h & RT_ID_TRI)
	{
		const uint32_t t = h & 0x0FFFFFFFu;
		for (const rt_model &M : c->models)
			for (uint32_t q = 0; q < M.part_count; ++q)
			{
				const rt_part &P = c->parts[M.part_begin + q];
				if (t >= P.tri_begin && t < P.tri_begin + P.tri_count)
				{
					id.object = (int32_t)M.object, id.sub = (int32_t)q, id.index = (int32_t)(t - P.tri_begin), id.octant = (int32_t)((h >> 28) & 7u);
					return id;
				}
			}
	}
	else if (h < c->prims.size())
		id.object = (int32_t)c->prims[h].object, id.sub = (int32_t)c->prims[h].sub;
	return id;
}

extern "C" int rt_intersect_object(rt_ctx *c, uint32_t object, const rt_ray *rays, const rt_hit *in, float min, rt_hit *out, uint32_t n)
{
	if (c) adopt_scene(c);
	if (!c || !rays || !in || !out) return fail(RT_E_INVALID, "rt_intersect_object: NULL argument");
	if (!c->hasScene) return fail(RT_E_STATE, "rt_intersect_object: no scene uploaded");
	CU(cudaSetDevice(c->device));
	if (c->frameInFlight) { int rc = finish_frame(c); if (rc != RT_OK) return rc; }
	uint32_t p0 = 0, p1 = 0;
	int model = -1;
	for (size_t p = 0; p < c->prims.size(); ++p)
		if (c->prims[p].object == object) { if (p1 == 0) p0 = (uint32_t)p; p1 = (uint32_t)p + 1; }
	for (size_t m = 0; m < c->models.size(); ++m)
		if (c->models[m].object == object) model = (int)m;
	if (p1 == 0 && model < 0) return fail(RT_E_INVALID, "rt_intersect_object: object %u is not part of the uploaded scene (hidden or out of range)", object);
	std::vector<uint32_t> skip(n), ids(n);
	for (uint32_t i = 0; i < n; ++i) skip[i] = encode_id(c, in[i].id);
	DevBuf<uint8_t> dRays, dIn, dOut;
	DevBuf<uint32_t> dSkip, dIds;
	cudaStream_t st = c->stream;
	static_assert(sizeof(rt_ray) == 48 && sizeof(rt_hit) == 80, "rt_ray / rt_hit layout");
	CU(dRays.upload((const uint8_t *)rays, sizeof(rt_ray) * (size_t)n, st));
	CU(dIn.upload((const uint8_t *)in, sizeof(rt_hit) * (size_t)n, st));
	CU(dSkip.upload(skip.data(), n, st));
	CU(dOut.reserve(sizeof(rt_hit) * (size_t)n));
	CU(dIds.reserve(n));
	rtk_intersect_object(st, c->S, p0, p1, model, dRays.p, dIn.p, dSkip.p, min, dIds.p, dOut.p, n);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out, dOut.p, sizeof(rt_hit) * (size_t)n, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(ids.data(), dIds.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	for (uint32_t i = 0; i < n; ++i)
		if (ids[i] != RT_ID_NONE && out[i].id.distance < in[i].id.distance)
			out[i].id = decode_id(c, ids[i], out[i].id.distance);
		else
			out[i] = in[i];
	dRays.release(), dIn.release(), dOut.release(), dSkip.release(), dIds.release();
	return RT_OK;
}

extern "C" int rt_read_hit_ids(rt_ctx *c, rt_hit_id *ids)
{
	if (c) adopt_scene(c);
	if (!c || !ids) return fail(RT_E_INVALID, "rt_read_hit_ids: NULL argument");
	if (c->frameInFlight) { int rc = finish_frame(c); if (rc != RT_OK) return rc; }
	if (!c->frameValid) return fail(RT_E_STATE, "rt_read_hit_ids: no finished frame");
	CU(cudaSetDevice(c->device));
	const uint32_t n = c->lastPixels;
	std::vector<uint2> hid(n);
	std::vector<float4> hp(n);
	CU(cudaMemcpy(hid.data(), c->levels[0].hit_id.p, n * sizeof(uint2), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(hp.data(), c->levels[0].hit_p.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
	const int W = c->outW, H = c->outH;
	for (size_t i = 0; i < (size_t)W * H; ++i) ids[i] = rt_hit_id{ -1, -1, -1, -1, 1e20f };
	const uint32_t world = c->lastParams.world > 1 ? c->lastParams.world : 1, rank = c->lastParams.world > 1 ? c->lastParams.rank : 0;
	const uint32_t w64 = (uint32_t)(W / 64) * 64u, tilesX = w64 >> 3;
	for (uint32_t i = 0; i < n; ++i)
	{
		const uint32_t tile = i >> 6, in = i & 63u, tx = tile % tilesX, ty = tile / tilesX;
		const uint32_t tileRows = c->lastParams.tile_rows ? c->lastParams.tile_rows : 64u;
		const uint32_t x = tx * 8u + (in & 7u), row = ty * 8u + (in >> 3), band = row / tileRows + c->lastTileFirst;
		const uint32_t y = shard_tile(band, rank, world, (c->lastParams.flags & RT_FLAG_SERPENTINE) && world > 1) * tileRows + row % tileRows;
		rt_hit_id id = { -1, -1, -1, -1, hp[i].w };
		const uint32_t h = hid[i].y;
		if (h != RT_ID_NONE)
		{
			if (h & RT_ID_TRI)
			{
				const uint32_t t = h & 0x0FFFFFFFu;
				for (const rt_model &M : c->models)
				{
					bool found = false;
					for (uint32_t q = 0; q < M.part_count && !found; ++q)
					{
						const rt_part &P = c->parts[M.part_begin + q];
						if (t >= P.tri_begin && t < P.tri_begin + P.tri_count)
						{
							id.object = (int32_t)M.object, id.sub = (int32_t)q, id.index = (int32_t)(t - P.tri_begin), id.octant = (int32_t)((h >> 28) & 7u);
							found = true;
						}
					}
					if (found) break;
				}
			}
			else if (h < c->prims.size())
				id.object = (int32_t)c->prims[h].object, id.sub = (int32_t)c->prims[h].sub;
		}
		ids[(size_t)y * W + x] = id;
	}
	return RT_OK;
}

extern "C" int rt_transfer_totals(uint64_t *h2d, uint64_t *d2h)
{
	if (h2d) *h2d = g_h2dTotal.load();
	if (d2h) *d2h = g_d2hTotal.load();
	return RT_OK;
}

extern "C" int rt_read_counters(rt_ctx *c, rt_counters *out)
{
	if (c) adopt_scene(c);
	if (!c || !out) return fail(RT_E_INVALID, "rt_read_counters: NULL argument");
	if (c->frameInFlight) { int rc = finish_frame(c); if (rc != RT_OK) return rc; }
	memset(out, 0, sizeof *out);
	out->upload_ms = c->uploadMs, out->build_ms = c->buildMs, out->bvh_nodes = c->bvhNodes, out->bvh_depth = c->bvhDepth;
	out->bvh_refit = (c->sceneFrom ? c->sceneFrom->lastUploadRefit : c->lastUploadRefit) ? 1u : 0u;
	out->h2d_bytes = c->uploadBytes + c->frameH2D, out->d2h_bytes = c->frameD2H;
	if (!c->frameValid) return RT_OK;
	const WaveState &W = *c->hWave;
	uint32_t enabled = 0;
	for (const rt_light &l : c->lights) if (l.enabled) ++enabled;
	out->primary = c->lastParams.type == RT_TYPE_CHECK ? 0 : W.count[0];
	out->reflect = W.n_reflect, out->refract = W.n_refract;
	unsigned long long hits = 0;
	for (uint32_t l = 0; l <= c->lastMaxLevel; ++l) hits += W.n_hit[l];
	const uint32_t ty = c->lastParams.type;
	out->shadow = (ty >= RT_TYPE_CHECK && ty <= RT_TYPE_MATERIAL) ? 0 : hits * enabled;
	out->nodes_visited = W.nodes_visited, out->tri_tests = W.tri_tests, out->prim_tests = W.prim_tests;
	if (getenv("RT_B200_PRINT_HIST") && W.nodes_visited)
	{
		fprintf(stderr, "nodes-per-ray histogram (log2 buckets) closest:");
		for (int b = 0; b < 12; ++b) fprintf(stderr, " %u", W.node_hist[b]);
		fprintf(stderr, "\n                                          shadow:");
		for (int b = 12; b < 24; ++b) fprintf(stderr, " %u", W.node_hist[b]);
		fprintf(stderr, "\n");
		if (W.t0)
		{
			fprintf(stderr, "k_frame timeline: rays started per 16.4 us bin; columns: closest L0..L7 | shadow L0..L7\n");
			int last = 127;
			while (last > 0) { bool any = false; for (int k = 0; k < 16; ++k) any |= W.timeline[last][k] != 0; if (any) break; --last; }
			for (int b = 0; b <= last; ++b)
			{
				fprintf(stderr, "%5.0f us:", b * 16.384);
				for (int k = 0; k < 16; ++k) fprintf(stderr, k == 8 ? " | %6u" : " %6u", W.timeline[b][k]);
				fprintf(stderr, "\n");
			}
		}
		if (W.lane_cap[0] || W.lane_cap[1])
			fprintf(stderr, "lane utilisation bound (sum nodes / 32 x longest lane per batch): closest %.3f shadow %.3f\n",
				W.lane_cap[0] ? (double)W.lane_sum[0] / (double)W.lane_cap[0] : 0.0, W.lane_cap[1] ? (double)W.lane_sum[1] / (double)W.lane_cap[1] : 0.0);
	}
	if (c->ssValid)
	{
		out->primary = c->ssTotals.primary, out->shadow = c->ssTotals.shadow, out->reflect = c->ssTotals.reflect, out->refract = c->ssTotals.refract;
		out->nodes_visited = c->ssTotals.nodes_visited, out->tri_tests = c->ssTotals.tri_tests, out->prim_tests = c->ssTotals.prim_tests;
		out->render_ms = c->ssTotals.render_ms, out->trace_ms = c->ssTotals.trace_ms, out->shade_ms = c->ssTotals.shade_ms, out->other_ms = c->ssTotals.other_ms;
		out->launches = c->ssTotals.launches, out->frame_sched = c->frameSched ? 1u : 0u;
		return RT_OK;
	}
	out->render_ms = c->renderMs;
	out->trace_ms = c->traceMs, out->shadow_ms = c->shadowMs, out->shade_ms = c->shadeMs, out->other_ms = c->otherMs;
	out->launches = c->lastLaunches;
	out->frame_sched = c->frameSched ? 1u : 0u;
	return RT_OK;
}
