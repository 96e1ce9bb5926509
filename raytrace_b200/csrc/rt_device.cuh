// Device-side scene layout and the bit-exact FP32 element math.
//
// Everything in the "parity" section reproduces the reference's operation order, one IEEE
// rounding per source-level operation (this translation unit is compiled with -fmad=false, so
// nvcc never contracts a*b+c; the conservative BVH slab code asks for FMA explicitly):
//   dot   = (x0*y0 + x1*y1) + x2*y2          dpps 0x71, /root/reference/3DElement.cpp:206-214
//           (the reference adds +0.0f to the z product; that only changes the sign of an
//            exact zero and is dropped on the device)
//   cross = (a.y*b.z - a.z*b.y, ...)          3DElement.cpp:190-197
//   normalise = v / sqrt(dot)  IEEE div/sqrt  3DElement.cpp:218-238
//   v / s  = v * (1/s)                        3DElement.cpp:137-147
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rt_b200.h"

#define RT_ID_NONE 0xFFFFFFFFu
#define RT_ID_TRI  0x80000000u
#define RT_MAX_LIGHTS 8
#define RT_MAX_LEVELS 16
#define RT_STACK 64

struct F3 { float x, y, z; };

__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{ x, y, z }; }
__device__ __forceinline__ F3 f3(const float4 &v) { return F3{ v.x, v.y, v.z }; }
__device__ __forceinline__ F3 operator+(const F3 &a, const F3 &b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(const F3 &a, const F3 &b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(const F3 &a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 mixmul(const F3 &a, const F3 &b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(const F3 &a, const F3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ F3 cross(const F3 &a, const F3 &b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ F3 normalize(const F3 &v)
{
	const float len = sqrtf(dot(v, v));
	return f3(v.x / len, v.y / len, v.z / len);
}
// SSE min/max (second operand on NaN/equal) and std::min/std::max as the reference calls them
__device__ __forceinline__ float sse_min(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float sse_max(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float std_min(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return a < b ? b : a; }

// float images of the reference's double-literal comparisons (t > 1e-6, |a| < 1e-6):
// 0x358637BD = 9.99999997e-7f is the largest float below 1e-6, so for float x
//   x > 1e-6  <=>  x > 9.99999997e-7f      and      x < 1e-6  <=>  x <= 9.99999997e-7f
#define RT_EPS6_BELOW 9.99999997e-7f
__device__ __forceinline__ bool gt_1em6(float x) { return x > RT_EPS6_BELOW; }
__device__ __forceinline__ bool lt_1em6(float x) { return x <= RT_EPS6_BELOW; }

// ---- flattened scene, device pointers --------------------------------------------------------

struct DevLight
{
	float4 position, ambient, diffuse, specular, attenuation;
	uint32_t type, enabled, pad0, pad1;
};

// A launch traces a BATCH of frames of one scene (rt_render_batch_async; rt_render_async = a batch of one): same
// size, lights, depth and shard, one camera and one framebuffer per frame.  The frames share the ray queues --
// level-0 slot s belongs to frame s / pix_per_frame -- so every warp serves every frame and the thin tail of
// the ray trees is paid once per batch instead of once per frame.
#define RT_MAX_BATCH 64
struct BatchFrame
{
	float4 cam_u, cam_v, cam_n, cam_pos;
	uint8_t *out;              // RGB8 framebuffer of this frame (device)
	uint64_t pad_;
};

// per-launch constants (device memory, rewritten by every rt_render_async)
struct FrameParams
{
	float4 cam_u, cam_v, cam_n, cam_pos;
	double dp;                 // tan(fovy*pi/360)/(height/2), RayTracer.cpp:14
	float zNear, zFar;         // zFar = float(sqrt(2)*cam.zFar), RayTracer.cpp:15
	int width, height;         // latched frame size
	int blk_w, blk_h;          // 64x64 tiles actually rendered
	int half_w, half_h;        // width/2, height/2 (integer division)
	uint32_t max_level, type;
	uint32_t rank, world;
	uint32_t n_rows;           // rows rendered by this shard
	uint32_t n_lights;
	uint32_t n_enabled;        // enabled lights, listed in light order
	uint32_t enabled_index[RT_MAX_LIGHTS];
	uint32_t epoch;            // 1..65535, stamps ray_meta so k_frame consumers can tell a written slot
	uint32_t tile_rows;        // shard tile height (8..64 rows)
	uint32_t sched_flags;      // k_frame: bit 0 = ray queues are claimed deepest level first, bit 1 = idle CTAs retire when the frame runs thin, bit 2 = level-0 rays are generated inside k_frame
	uint32_t sms;              // SM count (k_frame: CTAs below this index never retire)
	uint32_t retire_rays;      // k_frame: a CTA that may retire does so when idle and outstanding < (i + 1) * retire_rays  (keep_div == 1: (i - sms + 1))
	uint32_t serpentine;       // RT_FLAG_SERPENTINE: odd tile groups are dealt to the ranks in reverse order
	uint32_t keep_div;         // k_frame retire policy: 1 = CTAs below `sms` never retire; d > 1 = only every d-th CTA is kept
	uint32_t keep_salt;        //   (CTA i is kept iff (i + i / sms + keep_salt) % d == 0: spread over the SMs, rotated per pipeline)
	uint32_t batch;            // frames in this launch (1..RT_MAX_BATCH)
	uint32_t pix_per_frame;    // level-0 slots of one frame
	uint32_t tile_first;       // first of the shard's own tiles that this launch renders (rt_render_params::tile_first)
	float4 env_light;
	DevLight lights[RT_MAX_LIGHTS];
	BatchFrame frames[RT_MAX_BATCH];
};

// level-0 slot -> frame of the batch; `i` becomes the slot inside that frame
__device__ __forceinline__ uint32_t frame_of(const FrameParams &F, uint32_t &i)
{
	if (F.batch <= 1u) return 0u;
	const uint32_t f = i / F.pix_per_frame;
	i -= f * F.pix_per_frame;
	return f;
}

struct SceneItem   // one entry of the scene-order walk (RayTracer.cpp:458)
{
	uint32_t kind;      // 0 = single analytic primitive, 1 = BVH over a run of primitives, 2 = Model
	uint32_t first;     // prim flat index / first prim of the run / model index
	uint32_t count;     // run length
	int32_t root;       // BVH root node (kinds 1, 2)
};
#define RT_ITEM_PRIM 0u
#define RT_ITEM_PRIMBVH 1u
#define RT_ITEM_MODEL 2u

struct DevModel
{
	float4 border_min, border_max;   // VerMin/VerMax + position (Model.cpp:404)
	uint32_t part_begin, part_count, tri_begin, tri_count;
	uint32_t object, pad0, pad1, pad2;
};

// global row tile of the k-th tile of a shard (include/rt_b200.h rt_render_params, RT_FLAG_SERPENTINE)
__host__ __device__ __forceinline__ uint32_t shard_tile(uint32_t k, uint32_t rank, uint32_t world, uint32_t serpentine)
{
	return k * world + ((serpentine && (k & 1u)) ? world - 1u - rank : rank);
}

struct DevPart
{
	float4 box_min, box_max;         // borders + position (Model.cpp:418-419)
	uint32_t tri_begin, tri_count, material;
	int32_t texture;
};

struct BvhNode   // 64 bytes: both children's boxes + links, fetched as 4 x 128-bit loads
{
	float4 a;    // c0.lo.xyz, c0.hi.x
	float4 b;    // c0.hi.yz, c1.lo.xy
	float4 c;    // c1.lo.z, c1.hi.xyz
	int4 link;   // child0, child1 (>=0 node, <0 leaf: 0x80000000 | first<<3 | (count-1)), unused
};

// 4-wide node used by the traversal kernels: the two children of each child of a binary LBVH node
// (collapsed on the GPU, rt_build.cu k_collapse4).  SoA so that one float4 holds the same plane of
// all four child boxes; 7 x 128-bit loads, 128-byte aligned.  Unused slots carry a degenerate far-away box that no ray hits.
// RT_NODE_FETCH selects how a lane fetches its node (all variants read the same planes, only the data
// movement differs -- measured A/B in DESIGN.md "L1 wavefronts"):
//   0  seven LDG.128, near/far planes addressed by the ray's direction signs (lo* then hi*)
//   1  three LDG.256 (sm_100 256-bit loads: lo|hi of one axis per load) + one LDG.128, near/far by select
#ifndef RT_NODE_FETCH
#define RT_NODE_FETCH 0
#endif
// RT_TRI_FETCH: 0 = 48-byte triangle records in leaf order, three LDG.128; 1 = records padded to 64 bytes
// (32-byte aligned), two LDG.256
#ifndef RT_TRI_FETCH
#define RT_TRI_FETCH 0
#endif
#define RT_TRI_F4 (RT_TRI_FETCH == 1 ? 4 : 3)   // float4 per leaf-order triangle record
struct BvhNode4
{
#if RT_NODE_FETCH == 1
	float4 lox, hix, loy, hiy, loz, hiz;
#else
	float4 lox, loy, loz, hix, hiy, hiz;
#endif
	int4 link;     // per child: >=0 node index, <0 leaf code 0x80000000 | first<<3 | (count-1)
	int4 pad;
};

// 8-wide node with quantised child boxes, used for the Model BVHs (RT_BVH8): 96 bytes for 8 children (12 bytes per
// child against 32 in BvhNode4), six 128-bit loads.  Child k's box is [p + qlo_k * s, p + qhi_k * s] per axis with
// s = 2^e (the byte e is the biased float exponent), rounded OUTWARD and padded by one quantum at build time, so it
// contains the exact box: the BVH boxes are only a conservative cull, every accepted distance still comes from the exact
// triangle operator.  Fewer, fatter steps per ray (measured: DESIGN.md) and a third of the node bytes.
// Measured (profiles/r2k_bvh8_ab.md): 5.9 instead of 8.1 node visits per ray and 42 % instead of 67 % L1 wavefront load, but an
// 8-wide step costs 265 instructions against 145 and the wave kernels sit at their ~70 % issue ceiling with either tree:
// +17 % warp instructions, 11 % slower.  Off by default.
#ifndef RT_BVH8
#define RT_BVH8 0
#endif
struct BvhNode8
{
	float px, py, pz;          // origin = low corner of the union of the child boxes
	uint32_t exyz;             // ex | ey << 8 | ez << 16 | valid-child mask << 24
	uint32_t qlox[2], qloy[2]; // child k: byte k & 3 of word k >> 2
	uint32_t qloz[2], qhix[2];
	uint32_t qhiy[2], qhiz[2];
	int link[8];               // >= 0 node index, < 0 leaf code 0x80000000 | first << 3 | (count - 1), 0x7FFFFFFF unused
};

struct SceneDev
{
	// analytic primitives: 4 float4 + 1 int4 each
	const float4 *prim_geom;     // [4*i+0] pos.xyz,radius  [1] sphere: r2 | box: wmin | plane: normal  [2] box: wmax | plane: axisx  [3] box: lmax | plane: axisy
	const int4 *prim_meta;       // kind, material, texture, object<<8|sub
	const uint32_t *bvh_prims;   // leaf order -> prim flat index (prim-run BVHs)
	// materials: 4 float4 each (ambient, diffuse, specular, {shiness, reflect, refract, rfr})
	const float4 *materials;
	const int4 *textures;        // w, h, offset, -
	const uint8_t *texels;
	// models
	const DevModel *models;
	const DevPart *parts;
	// triangles in BVH leaf order: e1|id, e2|part<<8|octmask, p0|-   (clTri, 3DElement.h:122-127)
	const float4 *tri_geom;
	// shading data in original (part, index) order
	const float4 *tri_norms;     // 3 per triangle
	const float2 *tri_tcoords;   // 3 per triangle
	const uint32_t *tri_slot;    // original index -> leaf-order slot (to re-run the hit test when shading)
	const uint32_t *tri_part;    // original index -> global part index
	const BvhNode4 *nodes4;      // prim-run BVHs (and Model BVHs when RT_BVH8 == 0)
	const BvhNode8 *nodes8;      // Model BVHs (RT_BVH8)
	const SceneItem *items;
	uint32_t n_items, n_prims, n_tris, n_parts;
	uint32_t brute;              // RT_FLAG_BRUTE: ignore the BVHs, test every primitive (diagnostic cross-check)
	uint32_t dq_trigger;         // deferred triangle tests (rt_defer.cuh): a test round starts when fewer than 1/dq_trigger of the walking lanes are not waiting for results
};

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }
// one leaf-order triangle record: e1|id, e2|part<<8|octants, p0|-
__device__ __forceinline__ void load_tri(const float4 *tri_geom, uint32_t slot, float4 &g0, float4 &g1, float4 &g2);
// 256-bit read-only load (LDG.E.256, sm_100+): two float4 from a 32-byte aligned address, ONE L1 wavefront
__device__ __forceinline__ void ldg8(const void *p, float4 &a, float4 &b)
{
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

__device__ __forceinline__ void load_tri(const float4 *tri_geom, uint32_t slot, float4 &g0, float4 &g1, float4 &g2)
{
#if RT_TRI_FETCH == 1
	float4 pad;
	ldg8(&tri_geom[4 * (size_t)slot], g0, g1);
	ldg8(&tri_geom[4 * (size_t)slot + 2], g2, pad);
#else
	g0 = ldg4(&tri_geom[3 * (size_t)slot]), g1 = ldg4(&tri_geom[3 * (size_t)slot + 1]), g2 = ldg4(&tri_geom[3 * (size_t)slot + 2]);
#endif
}
