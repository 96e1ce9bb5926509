/*
 * rt_b200.h -- C ABI of the B200-native trace-and-shade path (librt_b200.so).
 *
 * The reference (XZiar/RayTrace) has no FFI; its render path is entered through two C++ class
 * surfaces:
 *   B1  RayTracer::start/stop + isFinish/useTime/output   /root/reference/RayTracer.h:16-56
 *   B2  DrawObject::intersect (per-primitive operator)     /root/reference/3DElement.h:185-202
 * This header is the boundary a maintainer binds instead of RayTracer.cpp's CPU workers: every
 * entry point below names the reference code it replaces.  Plain C types only; caller-owned host
 * buffers; library-owned device memory; every call returns 0 or a negative RT_E_* code and
 * leaves a message in rt_last_error().  There is NO CPU fallback: without a CUDA device (or
 * without the sm_100a kernels) rt_create() fails.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 2   /* 2: rt_render_params gained the tile window, rt_render_supersampled, rt_transfer_totals */

/* error codes */
#define RT_OK            0
#define RT_E_INVALID    -1   /* bad argument / inconsistent scene description */
#define RT_E_CUDA       -2   /* CUDA runtime error (text in rt_last_error) */
#define RT_E_NODEVICE   -3   /* no usable sm_100 device: the product path refuses to run */
#define RT_E_STATE      -4   /* call order violated (e.g. render before upload) */
#define RT_E_LIMIT      -5   /* a structural limit was exceeded (stack depth, counts) */

/* object kinds == MY_OBJECT_* (3DElement.h:4-8) */
#define RT_OBJ_SPHERE    1
#define RT_OBJ_CUBE      2
#define RT_OBJ_MODEL     3
#define RT_OBJ_PLANE     4
#define RT_OBJ_BALLPLANE 5
/* light kinds == MY_LIGHT_* (3DElement.h:13-15) */
#define RT_LIGHT_PARALLEL 1
#define RT_LIGHT_POINT    2
#define RT_LIGHT_SPOT     3
/* render types == MY_MODEL_* (RayTracer.h:5-13), the `type` argument of RayTracer::start; all nine run on the device */
#define RT_TYPE_CHECK     0x01
#define RT_TYPE_DEPTH     0x02
#define RT_TYPE_NORMAL    0x03
#define RT_TYPE_TEXTURE   0x04
#define RT_TYPE_MATERIAL  0x05
#define RT_TYPE_SHADOW    0x06
#define RT_TYPE_REFLECT   0x07
#define RT_TYPE_REFRACT   0x08
#define RT_TYPE_RAYTRACE  0x80

typedef struct rt_ctx rt_ctx;

typedef struct { float x, y, z, w; } rt_vec4;

/* Material, 3DElement.h:106-120 */
typedef struct
{
	rt_vec4 ambient, diffuse, specular, emission;
	float shiness, reflect, refract, rfr;
} rt_material;

/* Light, 3DElement.h:204-223: the fields RTfrac reads (RayTracer.cpp:474-503) */
typedef struct
{
	rt_vec4 position, ambient, diffuse, specular, attenuation;
	uint32_t type;      /* RT_LIGHT_* */
	uint32_t enabled;   /* Light::bLight */
	uint32_t pad0, pad1;
} rt_light;

/* Camera, 3DElement.h:225-237 */
typedef struct
{
	rt_vec4 u, v, n, position;
	int32_t width, height;
	float fovy, zNear, zFar;
	uint32_t pad0, pad1, pad2;
} rt_camera;

/* Texture, 3DElement.h:92-104: BGR8 rows, `offset` bytes into rt_scene_desc::texels */
typedef struct
{
	int32_t w, h;
	uint32_t offset;
	uint32_t pad0;
} rt_texture;

/*
 * One analytic primitive.  Sphere (Basic3DObject.h:5-16): position=centre, radius.
 * Cube (:18-33): position, a=min, b=max (object-local).  Plane (:35-50): position, a=normal,
 * b=axisx, c=axisy.  A BallPlane (:52-68) is passed as its 16 lattice spheres, sub = 1..16 in
 * the reference's loop order (Basic3DObject.cpp:492-497), position = the lattice centre.
 * Records must be sorted by (object, sub); object order is the closest-hit tie-break order
 * (RayTracer.cpp:458-465).
 */
typedef struct
{
	uint32_t kind;       /* RT_OBJ_SPHERE / RT_OBJ_CUBE / RT_OBJ_PLANE */
	uint32_t object;     /* index in Scene::Objects */
	uint32_t sub;
	uint32_t material;   /* index into rt_scene_desc::materials */
	int32_t texture;     /* index into textures, -1 = none */
	float radius, radius_sqr;
	uint32_t pad0;
	rt_vec4 position;
	rt_vec4 a, b, c;
} rt_prim;

/* One Model (Model.h:6-52): placement + bounds; its parts are parts[part_begin .. +part_count) */
typedef struct
{
	uint32_t object;             /* index in Scene::Objects */
	uint32_t part_begin, part_count;
	uint32_t pad0;
	rt_vec4 position;            /* DrawObject::position */
	rt_vec4 ver_min, ver_max;    /* Model::VerMin/VerMax, untranslated (Model.cpp:404) */
} rt_model;

/* One `usemtl` part (Model::parts[p], borders[2p], borders[2p+1], part_mtl, mtl_tex) */
typedef struct
{
	rt_vec4 border_min, border_max;  /* untranslated */
	uint32_t tri_begin, tri_count;   /* range in the triangle arrays; count <= 32767 */
	uint32_t material;
	int32_t texture;
} rt_part;

/*
 * Flattened Scene (Scene.h:24-29 + everything hanging off Objects).  Hidden objects (bShow ==
 * false) are simply not listed.  Triangle arrays are SoA copies of `Triangle` (3DElement.h:
 * 128-139): 3 points, 3 normals, 3 texture coordinates per triangle, parts in model order.
 * `geometry_epoch` lets the caller skip the heavy arrays: if it equals the epoch of the previous
 * upload on this context the tri_* pointers may be NULL and the resident triangles are reused.
 */
typedef struct
{
	rt_camera camera;
	rt_vec4 env_light;
	uint32_t n_lights;     const rt_light *lights;
	uint32_t n_materials;  const rt_material *materials;
	uint32_t n_textures;   const rt_texture *textures;
	size_t texel_bytes;    const uint8_t *texels;
	uint32_t n_prims;      const rt_prim *prims;
	uint32_t n_models;     const rt_model *models;
	uint32_t n_parts;      const rt_part *parts;
	uint32_t n_tris;
	const rt_vec4 *tri_points;    /* 3 * n_tris */
	const rt_vec4 *tri_norms;     /* 3 * n_tris */
	const float *tri_tcoords;     /* 6 * n_tris */
	uint64_t geometry_epoch;      /* 0 = always upload */
} rt_scene_desc;

/* Per-render parameters: the arguments and latched state of RayTracer::start (RayTracer.cpp:614-626) */
typedef struct
{
	uint32_t type;        /* RT_TYPE_* */
	uint32_t max_level;   /* RayTracer::maxLevel; levels 0..max_level are traced (RayTracer.cpp:453) */
	uint32_t rank, world; /* image-space shard: row tile t is rendered iff t % world == rank (RT_FLAG_SERPENTINE: see below); world=0 or 1 = whole frame */
	uint32_t flags;       /* RT_FLAG_* */
	uint32_t tile_rows;   /* height of a shard tile: 8, 16, 32 or 64 rows (0 = 64); finer tiles balance the ranks better */
	uint32_t tile_first;  /* window into the shard's own tiles: only its tiles tile_first .. tile_first + tile_count - 1 (counted in */
	uint32_t tile_count;  /* the order it owns them, top row of the frame last) are rendered; tile_count = 0: all of them */
} rt_render_params;

#define RT_FLAG_HIT_IDS   0x1   /* keep primary closest-hit identities for rt_read_hit_ids */
#define RT_FLAG_STATS     0x2   /* count BVH node visits / primitive tests (slower kernels) */
#define RT_FLAG_BRUTE     0x4   /* diagnostic: ignore the BVHs, test every primitive (small scenes) */
#define RT_FLAG_COMBINE_LEVELS 0x8 /* diagnostic: colour combine as one pass per ray level instead of the one-launch tree walk */
/* shard order: the tiles of group g = t / world go to ranks 0..world-1 for even g and world-1..0 for odd g
 * (boustrophedon), so a ray-cost gradient down the image (sky rows cheap, floor rows expensive) averages
 * out per rank instead of giving rank 0 the expensive tile of every group */
#define RT_FLAG_SERPENTINE 0x10

/* primary closest-hit identity, the GPU-side meaning of HitRes::obj (3DElement.h:174) */
typedef struct
{
	int32_t object;   /* index in Scene::Objects, -1 = miss */
	int32_t sub;      /* BallPlane slot (1..16) / Model part (clTri::numa) / 0 */
	int32_t index;    /* triangle index in its part (clTri::numb), else -1 */
	int32_t octant;   /* which octant copy of the triangle won (Model.cpp:765-785), else -1 */
	float distance;   /* HitRes::distance, 1e20 = miss */
} rt_hit_id;

/* Ray (3DElement.h:155-164) and HitRes (3DElement.h:166-183) as plain records, for rt_intersect_object */
typedef struct
{
	rt_vec4 origin, direction;   /* direction must be unit length (the Ray constructor normalises) */
	float mtlrfr;
	uint32_t type;               /* MY_RAY_* */
	uint32_t is_inside;          /* 0x00 / 0xFF */
	uint32_t pad0;
} rt_ray;

typedef struct
{
	rt_vec4 position, normal;
	float tu, tv;                /* HitRes::tcoord */
	int32_t material, texture;   /* indices into the uploaded tables, -1 = none (HitRes::mtl / tex) */
	rt_hit_id id;                /* HitRes::obj as an identity + HitRes::distance */
	float rfr;
	uint32_t is_inside;
	uint32_t pad0;
} rt_hit;

typedef struct
{
	uint64_t primary, shadow, reflect, refract;      /* rays = closest-hit or any-hit queries */
	uint64_t nodes_visited, tri_tests, prim_tests;   /* filled only with RT_FLAG_STATS */
	double render_ms;      /* device time of the last frame (CUDA events) */
	double trace_ms, shadow_ms, shade_ms, other_ms;   /* per-stage split (RT_FLAG_STATS) */
	double upload_ms, build_ms;                      /* last scene upload / LBVH build */
	uint32_t launches;                               /* kernels launched for the last frame */
	uint32_t bvh_nodes, bvh_depth;
	uint32_t frame_sched;                            /* 1: the last frame used the whole-frame persistent kernel, 0: per-level waves */
	uint64_t h2d_bytes;    /* host->device bytes of the last rt_upload_scene + rt_render_async */
	uint64_t d2h_bytes;    /* device->host bytes of the last frame (counters + rt_read_output) */
	uint32_t bvh_refit;    /* 1: the last rt_upload_scene refitted the Model BVHs (position-only edit) instead of rebuilding them; build_ms is then the refit */
	uint32_t pad0;
} rt_counters;

const char *rt_last_error(void);
int rt_abi_version(void);

/* replaces RayTracer::RayTracer (RayTracer.cpp:600-607): one context per GPU */
int rt_create(int device, rt_ctx **out);
void rt_destroy(rt_ctx *ctx);
/* optional: run on a caller-provided cudaStream_t (e.g. torch's current stream) */
int rt_set_stream(rt_ctx *ctx, void *cuda_stream);
/* Several frames in flight on one GPU (the reference renders one frame at a time, RayTracer.cpp:614-696;
 * its idiom for more is one RayTracer per view over the same Scene, RayTracer.cpp:600).  A shared
 * pipeline renders its PARENT's resident scene -- device tables and BVHs are aliased, not copied --
 * on its own stream with its own ray queues and framebuffer, so frame k+1 fills the SMs that the long
 * ray chains at the end of frame k leave idle.  The parent must outlive it; rt_upload_scene on a shared
 * pipeline uploads to the parent (and so changes what every pipeline of that parent renders next). */
int rt_create_shared(rt_ctx *parent, rt_ctx **out);
/* resident traversal CTAs per SM this pipeline may occupy (1..8, 0 = all): pipelines that run
 * concurrently split the 8 slots between them */
int rt_set_sm_share(rt_ctx *ctx, int ctas_per_sm);
/* (new) announce the largest rt_render_batch_async launch this pipeline will see: ray queues and library-owned framebuffers are
 * then sized for n_frames frames by the next launch, so batches that grow (the coalesced RayTracer::start() calls: one frame,
 * then three, then eight) do not free and re-allocate gigabytes at every new size.  0 = size by the launch (default). */
int rt_reserve_batch(rt_ctx *ctx, uint32_t n_frames);

/* replaces the per-start scene walk + Model::RTPrepare (RayTracer.cpp:621-625, Model.cpp:402-480):
 * copies the description to SoA device buffers and (re)builds the LBVHs when geometry changed */
int rt_upload_scene(rt_ctx *ctx, const rt_scene_desc *scene);

/* replaces RayTracer::start's thread fan-out (RayTracer.cpp:626-695): enqueues one frame, returns at once */
int rt_render_async(rt_ctx *ctx, const rt_render_params *params);
/* A batch of frames of the uploaded scene in ONE launch: frame f is seen through cameras[f] (NULL: the uploaded
 * camera for every frame; only position and orientation may differ from it) and lands in device_outputs[f]
 * (caller-owned device buffers of >= 3*width*height bytes each; NULL: library-owned, read with
 * rt_read_batch_output).  The reference renders one frame per start() with a thread pool that shares one tile
 * counter (RayTracer.cpp:626-672); here the frames of a batch share the ray queues the same way, so the thin
 * tail of a frame's ray trees is paid once per batch -- what makes 1/8-frame shards on 8 GPUs efficient.
 * n_frames <= 64.  rt_wait / rt_poll / rt_read_counters then refer to the whole batch. */
int rt_render_batch_async(rt_ctx *ctx, const rt_render_params *params, uint32_t n_frames, const rt_camera *cameras, void *const *device_outputs);
/* frame `frame` of the last batch -> host.  rows_only bit 0: copy only the shard's rows (see rt_read_output_rows);
 * bit 1: enqueue the copy and return -- a later call without bit 1 waits for all of them (they run in order) */
int rt_read_batch_output(rt_ctx *ctx, uint32_t frame, uint8_t *rgb, size_t stride, int rows_only);
/* Jittered supersampling (BASELINE configs[4]: 16 samples per pixel): sample s of every pixel is seen through
 * sample_cameras[s] -- the caller's camera with its forward vector offset by a sub-pixel step, n' = n + u*(dx*dp) +
 * v*(dy*dp), the way primary rays are made in RayTracer.cpp:20-27 -- every sample is quantised by Color::put
 * (3DElement.cpp:463-468) exactly like a frame of its own, and the frame is the INTEGER mean of the n_samples bytes per
 * channel (floor), i.e. bit-identical to averaging n_samples reference renders.  On the device the samples of a band of
 * row tiles are ONE frame batch (they share the ray queues) followed by an averaging kernel; a frame too large for one
 * launch (8K x 16 spp = 527 M primary rays) is walked band by band with the tile window of rt_render_params, so the
 * sample accumulation never leaves the GPU.  The result lands in the context's framebuffer (rt_read_output,
 * rt_set_output, rt_push_rows work as after rt_render_async); counters are summed over the bands.  Synchronous.
 * n_samples <= 64. */
int rt_render_supersampled(rt_ctx *ctx, const rt_render_params *params, uint32_t n_samples, const rt_camera *sample_cameras);
/* replaces the isFinish / useTime polling protocol (RayTracer.h:47-48) */
int rt_poll(rt_ctx *ctx, int *done, double *seconds);
int rt_wait(rt_ctx *ctx, double *seconds);
/* replaces RayTracer::stop (RayTracer.cpp:698-701) */
int rt_stop(rt_ctx *ctx);

/* RayTracer::output (RayTracer.h:46): RGB8, row 0 = bottom, `stride` bytes per row (>= 3*width).
 * Only floor(W/64)*64 x floor(H/64)*64 pixels are rendered, the rest is 127 (RayTracer.cpp:13,620).
 * With world > 1 only this rank's rows are valid. */
int rt_read_output(rt_ctx *ctx, uint8_t *rgb, size_t stride);
/* same destination layout, but copies only the rows the last frame's shard rendered (world > 1: 1/world of the
 * D2H bytes); every other row of `rgb` is left as it is (RayTracer pre-fills `output` with 127 like the
 * reference constructor, RayTracer.cpp:603-606).  world <= 1: identical to rt_read_output. */
int rt_read_output_rows(rt_ctx *ctx, uint8_t *rgb, size_t stride);
/* device-resident framebuffer of the last frame (for NCCL gathers / zero-copy consumers) */
int rt_output_device(rt_ctx *ctx, void **device_ptr, size_t *bytes);
/* render into a caller-owned device buffer of >= 3*width*height bytes (e.g. a torch tensor that is
 * then gathered over NCCL); NULL switches back to the library-owned framebuffer */
int rt_set_output(rt_ctx *ctx, void *device_ptr, size_t bytes);

/* Multi-GPU frame gather over NVLink with no SM work (one-sided put with signal).  The reference has
 * one framebuffer that all worker threads write (RayTracer.cpp:28-30); with image-space shards the
 * destination GPU exposes a frame-sized landing buffer over CUDA IPC (rt_landing_create -> 64-byte
 * handle, shipped to the other processes by the caller, e.g. torch.distributed), every rank maps it
 * (rt_landing_open) and, after rt_render_async, rt_push_rows copies its row tiles to their final
 * offsets with one strided peer copy on the copy engines and then writes `seq` to its flag word;
 * rt_landing_wait makes a stream of the destination (the consumer's, or the pipeline's own) wait until
 * ranks 0..world-1 delivered `seq`.
 * The landing buffer of the destination rank may also be its own rt_set_output target (then its own
 * rows need no copy).
 * Back-pressure: the consumer hands a buffer back with rt_landing_release(seq) -- a stream write of `seq` into the
 * buffer's ack word, behind whatever it read on that stream -- and a rank's NEXT push into the same buffer first makes
 * its stream wait (cuStreamWaitValue64 on the peer-mapped ack word, no SM work, no host round trip) until the frame it
 * pushed before has been released.  The owner arms the protocol with rt_landing_release(seq = 0) before it ships the
 * handle; a buffer that was never armed is not waited for (its ack word reads "all released"): the caller's own pacing
 * then decides, as before. */
typedef struct rt_landing rt_landing;
int rt_landing_create(rt_ctx *ctx, int width, int height, rt_landing **out, void *ipc_handle64);
int rt_landing_open(rt_ctx *ctx, int width, int height, const void *ipc_handle64, rt_landing **out);
void rt_landing_close(rt_landing *landing);
int rt_landing_ptr(rt_landing *landing, void **device_ptr, size_t *bytes);
int rt_push_rows(rt_ctx *ctx, rt_landing *landing, uint64_t seq);
int rt_push_batch_rows(rt_ctx *ctx, uint32_t frame, rt_landing *landing, uint64_t seq);   /* the same for frame `frame` of the last batch */
int rt_landing_wait(rt_ctx *ctx, rt_landing *landing, uint64_t seq, uint32_t world, void *consumer_stream /* NULL: ctx's stream */);
int rt_landing_release(rt_ctx *ctx, rt_landing *landing, uint64_t seq, void *consumer_stream /* NULL: ctx's stream */);

/* page-locked host memory for RayTracer::output: rt_read_output into it runs at full PCIe speed */
int rt_host_alloc(void **ptr, size_t bytes);
int rt_host_free(void *ptr);

/* B2: the per-primitive operator `HitRes DrawObject::intersect(const Ray&, const HitRes &hr, float min)`
 * (3DElement.h:201) of object `object` of the uploaded scene, evaluated on the device for n rays:
 * out[i] = in[i] on a miss, else the new hit with a strictly smaller distance.  in[i].id is hr.obj
 * (the primitive to skip) and in[i].id.distance is hr.distance; `min` is the any-hit threshold of
 * Model::intersect (Model.cpp:786).  Walks the object's primitives in the reference's own order
 * (no BVH), so even the order-dependent early exit is reproduced.  Synchronous. */
int rt_intersect_object(rt_ctx *ctx, uint32_t object, const rt_ray *rays, const rt_hit *in, float min, rt_hit *out, uint32_t n);

/* diagnostics / parity taps */
int rt_read_hit_ids(rt_ctx *ctx, rt_hit_id *ids /* width*height */);
int rt_read_counters(rt_ctx *ctx, rt_counters *out);
/* bytes this library has copied host->device / device->host in this process so far (every context: scene uploads,
 * per-launch tables and counters, frame read-backs).  A caller that brackets a region with two calls gets the bytes
 * that crossed the bus inside it (bench.py's e2e leg). */
int rt_transfer_totals(uint64_t *h2d_bytes, uint64_t *d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
