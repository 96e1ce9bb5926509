// Host-visible declarations of the wavefront buffers and kernel launchers (rt_kernels.cu) and
// of the LBVH builder (rt_build.cu).
#pragma once
#include "rt_device.cuh"

#define MY_RAY_BASERAY_    0x1
#define MY_RAY_SHADOWRAY_  0x2
#define MY_RAY_REFLECTRAY_ 0x3
#define MY_RAY_REFRACTRAY_ 0x4

// One recursion level of the ray tree: slot i is both ray i and ray-tree node i of that level.
struct LevelBuf
{
	float4 *ray_o;       // origin.xyz, Ray::mtlrfr
	float4 *ray_d;       // direction.xyz (unit), bwc ("benefit weight", RayTracer.cpp:451)
	uint2 *ray_meta;     // x: HitRes::obj to skip, y: Ray::type | Ray::isInside << 8
	float4 *hit_p;       // HitRes::position.xyz, HitRes::distance (1e20 = miss)
	uint2 *hit_id;       // x: closest primitive, y: `newobj` (RayTracer.cpp:456-465) = identity for child rays
	float4 *color;       // node-local colour, then combined colour; w = Color::alpha (distance)
	int4 *aux;           // x: reflect child slot, y: refract child slot (-1 none), z: material (-1 = no surface), w: bit0 reflect, bit1 refract, bit2 Beer
	uint8_t *shadow;     // [light][capacity]: 1 = occluded
	float4 *hit_n;       // HitRes::normal.xyz, bits of the material index
	float4 *hit_uv;      // HitRes::tcoord.u, .v, bits of the texture index (-1 none), unused
	uint32_t *hit_list;  // compacted (slot + 1) of the rays that found a surface inside [zNear, zFar]; 0 = not written yet
	uint32_t *order;     // coherence binning (rtk_bin_rays): k-th ray to trace -> slot; NULL = slot order
	uint32_t *sort_key;  // scratch of the binning pass: bin of every slot
	uint32_t capacity;
};

// Coherence binning of one level's rays before they are traced (wave scheduler, levels >= 1): a counting sort of the ray
// slots by (starts inside a sphere, direction octant, Morton cell of the origin in a grid over the bounded scene objects).
// Only the ORDER in which k_wave fetches the rays changes -- slot = ray-tree node stays as it is, so nothing downstream
// (children links, hit lists, combine) notices.
struct BinGrid
{
	float lo[3], scale[3];   // cell = clamp((origin - lo) * scale, 0, 2^bits - 1) per axis
	uint32_t bits;           // per axis: 3 * bits + 4 key bits, 2^(3 * bits + 4) bins
	uint32_t parts;          // which components make the key (experiments): 1 starts inside a sphere, 2 direction octant, 4 origin cell, 8 ray type
};
struct WaveState;
void rtk_bin_rays(cudaStream_t st, const LevelBuf &L, const WaveState *ws, uint32_t level, const BinGrid &G, uint32_t *hist, unsigned sms);

struct LevelSet { LevelBuf l[RT_MAX_LEVELS + 2]; };

struct WaveState
{
	uint32_t count[RT_MAX_LEVELS + 2];   // rays queued per level
	uint32_t n_hit[RT_MAX_LEVELS + 2];   // surfaces found per level (length of hit_list)
	uint32_t head_trace[RT_MAX_LEVELS + 2], head_shadow[RT_MAX_LEVELS + 2];   // work-fetch cursors of the persistent warps
	uint32_t head_light[RT_MAX_LEVELS + 2][RT_MAX_LIGHTS];                    // k_frame: shadow cursor per (level, enabled light)
	uint32_t overflow;                   // 1: a level ran out of slots, 2: the frame scheduler gave up waiting
	uint32_t stop_epoch;                 // rt_stop: the epoch of the frame to cancel (a late-landing stop of an older frame matches nothing)
	int outstanding;                     // k_frame: rays reserved and not yet finished (0 = frame complete)
	unsigned long long n_reflect, n_refract;
	unsigned long long nodes_visited, tri_tests, prim_tests;
	unsigned long long lane_sum[2], lane_cap[2];   // RT_FLAG_STATS (k_frame): per batch, sum of node visits / 32 x longest lane; [0] closest, [1] shadow
	unsigned long long t0;               // RT_FLAG_STATS (k_frame): %globaltimer of the first batch
	unsigned int timeline[128][16];      // RT_FLAG_STATS (k_frame): rays started per 16.4 us bin; column = level (closest) / 8 + level (shadow)
	unsigned int node_hist[24];          // RT_FLAG_STATS: rays by floor(log2(nodes visited + 1)), closest-hit [0..11], shadow [12..23]
};

void rtk_raygen(cudaStream_t st, const FrameParams *F, const LevelBuf &L, uint32_t n, unsigned sms);
void rtk_wave(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelBuf &L, const LevelBuf &N, const LevelBuf &Lprev, WaveState *ws,
	uint32_t level, bool traceOn, bool shadowOn, float zNear, uint32_t maxItems, unsigned sms, bool stats, unsigned ctasPerSm = 0, int walk = 0 /* 0: batch-synchronous voted walk, 2: two-stage (k_wave_split) */);
// whole-frame persistent scheduler (all ray levels in one launch)
void rtk_frame(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelSet &LS, WaveState *ws, uint32_t nPix, unsigned sms, bool stats, unsigned ctasPerSm = 0);
void rtk_shade(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelSet &LS, const WaveState *ws, uint32_t levels, uint32_t maxRays, unsigned sms, bool resetHits = false);
// one-launch post-order combine of every pixel's ray tree + Color::put (same arithmetic as rtk_combine level by level)
void rtk_resolve(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelSet &LS, const WaveState *ws, uint8_t *out, uint32_t nPix, unsigned sms);
void rtk_debug(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelBuf &L, uint32_t n, uint8_t *out, unsigned sms);
void rtk_intersect_object(cudaStream_t st, const SceneDev &S, uint32_t primBegin, uint32_t primEnd, int modelIndex, const void *rays, const void *in,
	const uint32_t *skipIds, float minT, uint32_t *outIds, void *out, uint32_t n);
void rtk_reset_hits(cudaStream_t st, const LevelSet &LS, const WaveState *ws, uint32_t levels, uint32_t maxRays, unsigned sms);
// integer mean of n sample frames over `tiles` row tiles of a shard, starting at its tile `tileFirst` (rt_render_supersampled)
void rtk_average(cudaStream_t st, const uint8_t *const *frames, uint32_t n, uint8_t *out, int width, uint32_t tileRows, uint32_t tileFirst, uint32_t tiles,
	uint32_t rank, uint32_t world, uint32_t serpentine, unsigned sms);
void rtk_combine(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelBuf &L, const LevelBuf &N, const WaveState *ws,
	uint32_t level, uint8_t *out, uint32_t maxRays, unsigned sms);

// ---- LBVH build (rt_build.cu) --------------------------------------------------------------------

struct BuildScratch;   // opaque, owned by the context

// Triangle preparation = the GPU restatement of Model::RTPrepare (Model.cpp:402-480): clTri edge
// form, world-space p0, octant membership mask per triangle.
struct TriPrepArgs
{
	const float4 *points;      // 3 per triangle, original order (untranslated)
	const DevModel *models;
	const DevPart *parts;
	const uint32_t *tri_part;  // original index -> global part
	const float4 *part_mid_pos;   // per part: (borders[2p]+borders[2p+1])*0.5 (untranslated), w unused
	const float4 *part_position;  // per part: model position
	float4 *tri_geom_orig;     // out: 3 float4 per triangle, original order: e1|id, e2|part<<8|octs, p0w|0
	float4 *box_lo, *box_hi;   // out: padded world AABB per triangle
	uint32_t n;
	uint32_t id_base;          // global original index of local triangle 0
};
void rtb_prepare_tris(cudaStream_t st, const TriPrepArgs &a);

struct PrimBoxArgs
{
	const float4 *prim_geom;
	const int4 *prim_meta;
	float4 *box_lo, *box_hi;
	uint32_t first, n;
};
void rtb_prim_boxes(cudaStream_t st, const PrimBoxArgs &a);

// Builds one LBVH over boxes [0,n): Morton codes of box centres -> radix sort -> Karras
// hierarchy -> bottom-up refit, emitting 64-byte BvhNodes at nodes[nodeBase ...) and the leaf
// order.  Returns the root link (node index, or a leaf code when n <= leafSize) and tree depth.
struct BvhBuildResult
{
	int root; uint32_t nodesUsed; uint32_t depth;
	// 4-wide nodes per level of the collapsed tree (breadth-first slots: level 0 = the root at nodeBase, level k >= 1 the
	// next levelNodes[k] slots); nLevels = 0 when the tree was not collapsed level by level.  What rtb_refit4 walks.
	uint32_t nLevels; uint32_t levelNodes[128];
	uint32_t maxStack;   // 8-wide collapse: the deepest traversal stack any root-to-leaf path can need (sum of children - 1)
};
int rtb_build(cudaStream_t st, BuildScratch **scratch, const float4 *box_lo, const float4 *box_hi, uint32_t n,
	uint32_t leafSize, BvhNode *nodes, BvhNode4 *nodes4, uint32_t nodeBase, uint32_t leafBase, uint32_t *leafOrder /* out: leaf slot -> input index */,
	BvhBuildResult *res, BvhNode8 *nodes8 = nullptr /* non-NULL: collapse to 8-wide quantised nodes instead of nodes4 */);
void rtb_refit8(cudaStream_t st, BvhNode8 *nodes8, uint32_t nodeBase, const uint32_t *levelNodes, uint32_t nLevels,
	const float4 *box_lo, const float4 *box_hi, const uint32_t *leafOrder);
void rtb_free_scratch(BuildScratch *s);
// Refit instead of rebuild (Model::RTPrepare after a MovePos only re-translates bounds, Model.cpp:404,418-419): the
// 4-wide tree of a model keeps its topology and leaf order, every node's child boxes are recomputed bottom-up, level by
// level, from the new per-triangle boxes.  leafOrder: global leaf slot -> index into box_lo / box_hi.
void rtb_refit4(cudaStream_t st, BvhNode4 *nodes4, uint32_t nodeBase, const uint32_t *levelNodes, uint32_t nLevels,
	const float4 *box_lo, const float4 *box_hi, const uint32_t *leafOrder);

// tri_geom (leaf order) = tri_geom_orig[leafOrder]; tri_slot[orig] = leaf slot
void rtb_scatter_tris(cudaStream_t st, const float4 *geomOrig, const uint32_t *leafOrder, uint32_t leafBase, uint32_t origBase, uint32_t n,
	float4 *geomLeaf, uint32_t *triSlot);
void rtb_offset_order(cudaStream_t st, uint32_t *order, uint32_t n, uint32_t add);
