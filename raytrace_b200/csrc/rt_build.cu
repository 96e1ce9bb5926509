// GPU LBVH build: Morton codes -> radix sort (hand-written LSD, 8 bits per pass) -> Karras (2012)
// hierarchy -> bottom-up refit that emits 64-byte two-child nodes -> collapse to 4-wide nodes.  Replaces the reference's per-frame, single-threaded octant
// binning (Model::RTPrepare, /root/reference/Model.cpp:402-480); the octant membership itself is
// still computed (bit-exactly) per triangle in k_prepare_tris because the traversal replays the
// reference's culling predicate with it.
#include "rt_kernels.h"
#include "rt_intersect.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

// radix sort tiling (see radix_sort_pairs)
#define RS_THREADS 256
#define RS_ROUNDS 16
#define RS_TILE (RS_THREADS * RS_ROUNDS)

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

struct BuildScratch
{
	uint32_t cap = 0;
	unsigned long long *keysIn = nullptr, *keysOut = nullptr;
	uint32_t *valsIn = nullptr, *valsOut = nullptr;
	uint32_t *sortHist = nullptr;   // radix sort: [256 digits][blocks] counts, then exclusive offsets
	int *bounds = nullptr;          // 6 ordered-int floats: min xyz, max xyz
	int *parentOfInternal = nullptr, *parentOfLeaf = nullptr;
	int2 *children = nullptr;       // per internal: left, right (>=0 internal, <0: ~leaf)
	int2 *range = nullptr;          // per internal: first, last leaf (inclusive)
	uint32_t *flags = nullptr;
	float4 *ilo = nullptr, *ihi = nullptr;   // internal node boxes
	uint32_t *height = nullptr;
};

void rtb_free_scratch(BuildScratch *s)
{
	if (!s) return;
	cudaFree(s->keysIn), cudaFree(s->keysOut), cudaFree(s->valsIn), cudaFree(s->valsOut), cudaFree(s->sortHist);
	cudaFree(s->bounds), cudaFree(s->parentOfInternal), cudaFree(s->parentOfLeaf), cudaFree(s->children), cudaFree(s->range);
	cudaFree(s->flags), cudaFree(s->ilo), cudaFree(s->ihi), cudaFree(s->height);
	delete s;
}

static int ensure_scratch(BuildScratch **ps, uint32_t n)
{
	if (!*ps) *ps = new BuildScratch();
	BuildScratch *s = *ps;
	if (n <= s->cap) return 0;
	BuildScratch fresh;
	cudaFree(s->keysIn), cudaFree(s->keysOut), cudaFree(s->valsIn), cudaFree(s->valsOut), cudaFree(s->sortHist);
	cudaFree(s->bounds), cudaFree(s->parentOfInternal), cudaFree(s->parentOfLeaf), cudaFree(s->children), cudaFree(s->range);
	cudaFree(s->flags), cudaFree(s->ilo), cudaFree(s->ihi), cudaFree(s->height);
	*s = fresh;
	const uint32_t cap = n + n / 8 + 1024;
	CK(cudaMalloc(&s->keysIn, sizeof(unsigned long long) * cap));
	CK(cudaMalloc(&s->keysOut, sizeof(unsigned long long) * cap));
	CK(cudaMalloc(&s->valsIn, sizeof(uint32_t) * cap));
	CK(cudaMalloc(&s->valsOut, sizeof(uint32_t) * cap));
	CK(cudaMalloc(&s->sortHist, sizeof(uint32_t) * 256 * ((cap + RS_TILE - 1) / RS_TILE + 1)));
	CK(cudaMalloc(&s->bounds, sizeof(int) * 8));
	CK(cudaMalloc(&s->parentOfInternal, sizeof(int) * cap));
	CK(cudaMalloc(&s->parentOfLeaf, sizeof(int) * cap));
	CK(cudaMalloc(&s->children, sizeof(int4) * cap));   // int2 per internal node; twice the room: the 8-wide collapse keeps int4 frontier records here
	CK(cudaMalloc(&s->range, sizeof(int4) * cap));
	CK(cudaMalloc(&s->flags, sizeof(uint32_t) * cap));
	CK(cudaMalloc(&s->ilo, sizeof(float4) * cap));
	CK(cudaMalloc(&s->ihi, sizeof(float4) * cap));
	CK(cudaMalloc(&s->height, sizeof(uint32_t) * cap));
	s->cap = cap;
	return 0;
}

// order-preserving float <-> int so atomicMin/atomicMax work on floats
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// ---- triangle preparation (Model::RTPrepare on the GPU) ------------------------------------------

__global__ void k_prepare_tris(TriPrepArgs a)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= a.n) return;
	const uint32_t part = a.tri_part[t];
	const F3 pos = f3(a.part_position[part]);
	const F3 p0 = f3(a.points[3 * t]), p1 = f3(a.points[3 * t + 1]), p2 = f3(a.points[3 * t + 2]);
	// clTri(t.points[1]-t.points[0], t.points[2]-t.points[0], t.points[0]+position), Model.cpp:426
	const F3 e1 = p1 - p0, e2 = p2 - p0, p0w = p0 + pos;
	// octant membership against the part's box centre, untranslated coordinates, Model.cpp:421,430-465
	const F3 va = f3(a.part_mid_pos[part]);
	const float tminx = sse_min(p0.x, sse_min(p1.x, p2.x)), tminy = sse_min(p0.y, sse_min(p1.y, p2.y)), tminz = sse_min(p0.z, sse_min(p1.z, p2.z));
	const float tmaxx = sse_max(p0.x, sse_max(p1.x, p2.x)), tmaxy = sse_max(p0.y, sse_max(p1.y, p2.y)), tmaxz = sse_max(p0.z, sse_max(p1.z, p2.z));
	uint32_t octs = 0;
	const bool ylo = tminy <= va.y, yhi = tmaxy >= va.y;
	if (tminx <= va.x)
	{
		if (tminz <= va.z) octs |= (ylo ? 1u : 0u) | (yhi ? 2u : 0u);
		if (tmaxz >= va.z) octs |= (ylo ? 4u : 0u) | (yhi ? 8u : 0u);
	}
	if (tmaxx >= va.x)
	{
		if (tminz <= va.z) octs |= (ylo ? 16u : 0u) | (yhi ? 32u : 0u);
		if (tmaxz >= va.z) octs |= (ylo ? 64u : 0u) | (yhi ? 128u : 0u);
	}
	a.tri_geom_orig[3 * t] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(t + a.id_base));
	a.tri_geom_orig[3 * t + 1] = make_float4(e2.x, e2.y, e2.z, __uint_as_float((part << 8) | octs));
	a.tri_geom_orig[3 * t + 2] = make_float4(p0w.x, p0w.y, p0w.z, 0.0f);
	// conservative world box of the triangle the hit test actually sees: p0w, p0w+e1, p0w+e2
	const F3 q1 = p0w + e1, q2 = p0w + e2;
	F3 lo = f3(fminf(p0w.x, fminf(q1.x, q2.x)), fminf(p0w.y, fminf(q1.y, q2.y)), fminf(p0w.z, fminf(q1.z, q2.z)));
	F3 hi = f3(fmaxf(p0w.x, fmaxf(q1.x, q2.x)), fmaxf(p0w.y, fmaxf(q1.y, q2.y)), fmaxf(p0w.z, fmaxf(q1.z, q2.z)));
	const float mag = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
	const float pad = mag * 4e-6f + 1e-7f;
	a.box_lo[t] = make_float4(lo.x - pad, lo.y - pad, lo.z - pad, 0);
	a.box_hi[t] = make_float4(hi.x + pad, hi.y + pad, hi.z + pad, 0);
}

void rtb_prepare_tris(cudaStream_t st, const TriPrepArgs &a)
{
	if (a.n) k_prepare_tris<<<(a.n + 255) / 256, 256, 0, st>>>(a);
}

__global__ void k_prim_boxes(PrimBoxArgs a)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n) return;
	const uint32_t p = a.first + i;
	const int kind = a.prim_meta[p].x;
	F3 lo, hi;
	if (kind == RT_OBJ_SPHERE)
	{
		const float4 g = a.prim_geom[4 * p];
		lo = f3(g.x - g.w, g.y - g.w, g.z - g.w), hi = f3(g.x + g.w, g.y + g.w, g.z + g.w);
	}
	else
		lo = f3(a.prim_geom[4 * p + 1]), hi = f3(a.prim_geom[4 * p + 2]);
	const float mag = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
	const float pad = mag * 8e-6f + 1e-6f;
	a.box_lo[i] = make_float4(lo.x - pad, lo.y - pad, lo.z - pad, 0);
	a.box_hi[i] = make_float4(hi.x + pad, hi.y + pad, hi.z + pad, 0);
}

void rtb_prim_boxes(cudaStream_t st, const PrimBoxArgs &a)
{
	if (a.n) k_prim_boxes<<<(a.n + 255) / 256, 256, 0, st>>>(a);
}

// ---- Morton codes --------------------------------------------------------------------------------

__global__ void k_init_bounds(int *b)
{
	if (threadIdx.x < 3) b[threadIdx.x] = 0x7F7FFFFF;              // +FLT_MAX ordered
	else if (threadIdx.x < 6) b[threadIdx.x] = f2ord(-3.4e38f);
}

__global__ void k_bounds(const float4 *lo, const float4 *hi, uint32_t n, int *b)
{
	float mn[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, mx[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const float4 l = lo[i], h = hi[i];
		const float cx = 0.5f * (l.x + h.x), cy = 0.5f * (l.y + h.y), cz = 0.5f * (l.z + h.z);
		mn[0] = fminf(mn[0], cx), mn[1] = fminf(mn[1], cy), mn[2] = fminf(mn[2], cz);
		mx[0] = fmaxf(mx[0], cx), mx[1] = fmaxf(mx[1], cy), mx[2] = fmaxf(mx[2], cz);
	}
	for (int k = 0; k < 3; ++k)
	{
		for (int o = 16; o > 0; o >>= 1)
		{
			mn[k] = fminf(mn[k], __shfl_down_sync(0xffffffffu, mn[k], o));
			mx[k] = fmaxf(mx[k], __shfl_down_sync(0xffffffffu, mx[k], o));
		}
		if ((threadIdx.x & 31) == 0)
		{
			atomicMin(&b[k], f2ord(mn[k]));
			atomicMax(&b[3 + k], f2ord(mx[k]));
		}
	}
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long v)
{
	v &= 0x1FFFFFull;
	v = (v | v << 32) & 0x1F00000000FFFFull;
	v = (v | v << 16) & 0x1F0000FF0000FFull;
	v = (v | v << 8) & 0x100F00F00F00F00Full;
	v = (v | v << 4) & 0x10C30C30C30C30C3ull;
	v = (v | v << 2) & 0x1249249249249249ull;
	return v;
}

// cube != 0: all three axes are quantised with the LARGEST extent, so Morton cells are cubes in world
// space.  Per-axis normalisation (cube == 0) splits a flat mesh (a height field) by height as often as
// by x and z, which yields sibling boxes that overlap almost completely in the ground plane.
__global__ void k_morton(const float4 *lo, const float4 *hi, uint32_t n, const int *b, unsigned long long *keys, uint32_t *vals, int cube)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float bx = ord2f(b[0]), by = ord2f(b[1]), bz = ord2f(b[2]);
	float ex = fmaxf(ord2f(b[3]) - bx, 1e-30f), ey = fmaxf(ord2f(b[4]) - by, 1e-30f), ez = fmaxf(ord2f(b[5]) - bz, 1e-30f);
	if (cube)
		ex = ey = ez = fmaxf(ex, fmaxf(ey, ez));
	const float4 l = lo[i], h = hi[i];
	const float scale = 2097151.0f;   // 2^21 - 1
	const float fx = (0.5f * (l.x + h.x) - bx) / ex, fy = (0.5f * (l.y + h.y) - by) / ey, fz = (0.5f * (l.z + h.z) - bz) / ez;
	const unsigned long long qx = (unsigned long long)fminf(fmaxf(fx * scale, 0.0f), scale);
	const unsigned long long qy = (unsigned long long)fminf(fmaxf(fy * scale, 0.0f), scale);
	const unsigned long long qz = (unsigned long long)fminf(fmaxf(fz * scale, 0.0f), scale);
	keys[i] = spread21(qx) << 2 | spread21(qy) << 1 | spread21(qz);
	vals[i] = i;
}

// ---- radix sort of (64-bit Morton key, 32-bit index) pairs ------------------------------------------
// Least-significant-digit first, 8 bits per pass, 8 passes, stable.  Per pass: (1) every block
// histograms its tile of RS_TILE keys, (2) one block turns the [digit][block] counts into exclusive
// offsets, (3) every block re-reads its tile in order and scatters: within a round of 256 keys a
// key's rank is (keys with its digit in earlier warps of the round) + (earlier lanes of its warp
// with that digit, __match_any_sync), on top of the block's running offset for the digit.

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned long long *keys, uint32_t n, int shift, uint32_t *hist, uint32_t nBlocks)
{
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * RS_TILE;
	for (uint32_t r = 0; r < RS_ROUNDS; ++r)
	{
		const uint32_t i = base + r * RS_THREADS + threadIdx.x;
		if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	hist[threadIdx.x * nBlocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(1024) k_rs_scan(uint32_t *hist, uint32_t total)
{
	// exclusive prefix sum over `total` = 256 * nBlocks counters (digit-major), one block
	__shared__ uint32_t part[1024];
	const uint32_t per = (total + 1023u) / 1024u;
	const uint32_t lo = threadIdx.x * per, hi = min(lo + per, total);
	uint32_t sum = 0;
	for (uint32_t i = lo; i < hi; ++i) sum += hist[i];
	part[threadIdx.x] = sum;
	__syncthreads();
	for (uint32_t off = 1; off < 1024; off <<= 1)
	{
		const uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	uint32_t run = part[threadIdx.x] - sum;
	for (uint32_t i = lo; i < hi; ++i)
	{
		const uint32_t c = hist[i];
		hist[i] = run;
		run += c;
	}
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned long long *keysIn, const uint32_t *valsIn, unsigned long long *keysOut,
	uint32_t *valsOut, uint32_t n, int shift, const uint32_t *hist, uint32_t nBlocks)
{
	__shared__ uint32_t offset[256];           // next free output slot per digit for this block
	__shared__ uint32_t warpCount[8][256];     // per round: keys per (warp, digit)
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	offset[threadIdx.x] = hist[threadIdx.x * nBlocks + blockIdx.x];
	const uint32_t base = blockIdx.x * RS_TILE;
	for (uint32_t r = 0; r < RS_ROUNDS; ++r)
	{
		for (uint32_t w = 0; w < 8; ++w) warpCount[w][threadIdx.x] = 0;
		__syncthreads();
		const uint32_t i = base + r * RS_THREADS + threadIdx.x;
		const bool valid = i < n;
		unsigned long long key = 0;
		uint32_t val = 0, digit = 0xFFFFFFFFu;
		if (valid) key = keysIn[i], val = valsIn[i], digit = (uint32_t)(key >> shift) & 255u;
		const uint32_t peers = __match_any_sync(0xffffffffu, digit);
		const uint32_t rankInWarp = __popc(peers & ((1u << lane) - 1u));
		if (valid && rankInWarp == 0) warpCount[warp][digit] = __popc(peers);
		__syncthreads();
		uint32_t before = 0;
		if (valid)
		{
			for (uint32_t w = 0; w < warp; ++w) before += warpCount[w][digit];
			const uint32_t dst = offset[digit] + before + rankInWarp;
			keysOut[dst] = key, valsOut[dst] = val;
		}
		__syncthreads();
		uint32_t add = 0;
		for (uint32_t w = 0; w < 8; ++w) add += warpCount[w][threadIdx.x];
		offset[threadIdx.x] += add;
		__syncthreads();
	}
}

static void radix_sort_pairs(cudaStream_t st, BuildScratch *s, uint32_t n)
{
	const uint32_t nBlocks = (n + RS_TILE - 1) / RS_TILE;
	unsigned long long *kin = s->keysIn, *kout = s->keysOut;
	uint32_t *vin = s->valsIn, *vout = s->valsOut;
	for (int pass = 0; pass < 8; ++pass)
	{
		const int shift = 8 * pass;
		k_rs_hist<<<nBlocks, RS_THREADS, 0, st>>>(kin, n, shift, s->sortHist, nBlocks);
		k_rs_scan<<<1, 1024, 0, st>>>(s->sortHist, 256u * nBlocks);
		k_rs_scatter<<<nBlocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, s->sortHist, nBlocks);
		unsigned long long *tk = kin; kin = kout; kout = tk;
		uint32_t *tv = vin; vin = vout; vout = tv;
	}
	// 8 passes: the sorted data are back in the buffers the sort started from; the callers read
	// keysOut / valsOut, so swap the names
	unsigned long long *tk = s->keysIn; s->keysIn = s->keysOut; s->keysOut = tk;
	uint32_t *tv = s->valsIn; s->valsIn = s->valsOut; s->valsOut = tv;
}

// ---- Karras hierarchy ----------------------------------------------------------------------------

__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j)
{
	if (j < 0 || j >= n) return -1;
	const unsigned long long a = keys[i], b = keys[j];
	if (a == b) return 64 + __clz(i ^ j);
	return __clzll((long long)(a ^ b));
}

__global__ void k_karras(const unsigned long long *keys, int n, int2 *children, int2 *range, int *parentOfInternal, int *parentOfLeaf)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
	const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	const int dmin = delta(keys, n, i, i - d);
	int lmax = 2;
	while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
	int l = 0;
	for (int t = lmax >> 1; t >= 1; t >>= 1)
		if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
	const int j = i + l * d;
	const int dnode = delta(keys, n, i, j);
	int s = 0, t = l;
	do
	{
		t = (t + 1) >> 1;
		if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
	} while (t > 1);
	const int gamma = i + s * d + min(d, 0);
	const int lo = min(i, j), hi = max(i, j);
	const int left = (lo == gamma) ? ~gamma : gamma;
	const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
	children[i] = make_int2(left, right);
	range[i] = make_int2(lo, hi);
	if (left >= 0) parentOfInternal[left] = i; else parentOfLeaf[~left] = i;
	if (right >= 0) parentOfInternal[right] = i; else parentOfLeaf[~right] = i;
	if (i == 0) parentOfInternal[0] = -1;
}

// ---- bottom-up refit + node emission -------------------------------------------------------------

struct RefitArgs
{
	const float4 *box_lo, *box_hi;   // input boxes (unsorted)
	const uint32_t *sorted;          // leaf slot -> input index
	const int2 *children, *range;
	const int *parentOfInternal, *parentOfLeaf;
	uint32_t *flags;
	float4 *ilo, *ihi;
	uint32_t *height;
	BvhNode *nodes;
	uint32_t nodeBase, leafBase, leafSize;
	int n;
};

__device__ __forceinline__ int child_link(const RefitArgs &a, int child)
{
	if (child < 0)
		return (int)(0x80000000u | ((a.leafBase + (uint32_t)(~child)) << 3));
	const int2 r = a.range[child];
	const uint32_t cnt = (uint32_t)(r.y - r.x + 1);
	if (cnt <= a.leafSize)
		return (int)(0x80000000u | ((a.leafBase + (uint32_t)r.x) << 3) | (cnt - 1u));
	return (int)(a.nodeBase + (uint32_t)child);
}

__global__ void k_refit(RefitArgs a)
{
	const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
	if (leaf >= a.n) return;
	int node = a.parentOfLeaf[leaf];
	while (node >= 0)
	{
		// the second thread to arrive owns the node; its sibling subtree is complete and visible
		if (atomicAdd(&a.flags[node], 1u) == 0u)
			return;
		__threadfence();
		const int2 ch = a.children[node];
		float4 l0, h0, l1, h1;
		uint32_t hgt0 = 0, hgt1 = 0;
		if (ch.x < 0) { const uint32_t s = a.sorted[~ch.x]; l0 = a.box_lo[s], h0 = a.box_hi[s]; }
		else { l0 = __ldcg(&a.ilo[ch.x]), h0 = __ldcg(&a.ihi[ch.x]), hgt0 = __ldcg(&a.height[ch.x]); }   // L2 reads: written by another SM
		if (ch.y < 0) { const uint32_t s = a.sorted[~ch.y]; l1 = a.box_lo[s], h1 = a.box_hi[s]; }
		else { l1 = __ldcg(&a.ilo[ch.y]), h1 = __ldcg(&a.ihi[ch.y]), hgt1 = __ldcg(&a.height[ch.y]); }
		const int link0 = child_link(a, ch.x), link1 = child_link(a, ch.y);
		if (link0 < 0) hgt0 = 0;
		if (link1 < 0) hgt1 = 0;
		BvhNode out;
		out.a = make_float4(l0.x, l0.y, l0.z, h0.x);
		out.b = make_float4(h0.y, h0.z, l1.x, l1.y);
		out.c = make_float4(l1.z, h1.x, h1.y, h1.z);
		out.link = make_int4(link0, link1, 0, 0);
		a.nodes[a.nodeBase + node] = out;
		a.ilo[node] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0);
		a.ihi[node] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0);
		a.height[node] = 1u + max(hgt0, hgt1);
		__threadfence();
		node = a.parentOfInternal[node];
	}
}

// ---- collapse to 4-wide nodes ----------------------------------------------------------------------
// Every binary node at even depth becomes a 4-wide node whose children are its grandchildren
// (a child that is already a leaf stays a single child).  Only even-depth nodes are reachable from
// the root through 4-wide links, so odd-depth slots of nodes4 stay unused.
__global__ void k_collapse4(const BvhNode *nodes, BvhNode4 *nodes4, const int *parentOfInternal, uint32_t nodeBase, int nInternal)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nInternal) return;
	int depth = 0;
	for (int p = parentOfInternal[i]; p >= 0; p = parentOfInternal[p]) ++depth;
	if (depth & 1) return;
	const BvhNode n = nodes[nodeBase + i];
	// unused slot: a degenerate box far away -- the min/max slab test is order-agnostic, so an
	// "inverted" box would be hit; a point at 3e38 gives |t| ~ 3e38/|d| on every axis and always misses
	const float far = 3.0e38f;
	float lo[4][3], hi[4][3];
	int link[4];
	for (int k = 0; k < 4; ++k)
	{
		lo[k][0] = lo[k][1] = lo[k][2] = far, hi[k][0] = hi[k][1] = hi[k][2] = far;
		link[k] = 0x7FFFFFFF;
	}
	int m = 0;
	for (int side = 0; side < 2; ++side)
	{
		const int l = side ? n.link.y : n.link.x;
		if (l < 0)
		{
			// leaf child: keep it with the box this node stores for it
			if (side == 0) lo[m][0] = n.a.x, lo[m][1] = n.a.y, lo[m][2] = n.a.z, hi[m][0] = n.a.w, hi[m][1] = n.b.x, hi[m][2] = n.b.y;
			else lo[m][0] = n.b.z, lo[m][1] = n.b.w, lo[m][2] = n.c.x, hi[m][0] = n.c.y, hi[m][1] = n.c.z, hi[m][2] = n.c.w;
			link[m++] = l;
		}
		else
		{
			const BvhNode c = nodes[l];
			lo[m][0] = c.a.x, lo[m][1] = c.a.y, lo[m][2] = c.a.z, hi[m][0] = c.a.w, hi[m][1] = c.b.x, hi[m][2] = c.b.y;
			link[m++] = c.link.x;
			lo[m][0] = c.b.z, lo[m][1] = c.b.w, lo[m][2] = c.c.x, hi[m][0] = c.c.y, hi[m][1] = c.c.z, hi[m][2] = c.c.w;
			link[m++] = c.link.y;
		}
	}
	BvhNode4 o;
	o.lox = make_float4(lo[0][0], lo[1][0], lo[2][0], lo[3][0]);
	o.loy = make_float4(lo[0][1], lo[1][1], lo[2][1], lo[3][1]);
	o.loz = make_float4(lo[0][2], lo[1][2], lo[2][2], lo[3][2]);
	o.hix = make_float4(hi[0][0], hi[1][0], hi[2][0], hi[3][0]);
	o.hiy = make_float4(hi[0][1], hi[1][1], hi[2][1], hi[3][1]);
	o.hiz = make_float4(hi[0][2], hi[1][2], hi[2][2], hi[3][2]);
	o.link = make_int4(link[0], link[1], link[2], link[3]);
	o.pad = make_int4(0, 0, 0, 0);
	nodes4[nodeBase + i] = o;
}

__global__ void k_copy_order(const uint32_t *sorted, uint32_t n, uint32_t *out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = sorted[i];
}

// ---- collapse to 4-wide nodes, surface-area guided ---------------------------------------------------
// Top-down, one launch per level of the 4-wide tree.  A 4-wide node starts with the two children of
// its binary node and keeps opening the inner child with the LARGEST surface area (the one a ray
// is most likely to enter anyway) until it has four children or only leaves are left.  Compared with
// "every even-depth node adopts its grandchildren" this fills the four slots where the fixed pattern
// leaves one empty next to every leaf child, and it cuts across the binary tree where the LBVH split
// was lopsided.  Nodes are allocated in breadth-first order, so the top of the tree is contiguous.
struct Collapse4Args
{
	const BvhNode *nodes;      // binary nodes, links are global indices (leaves < 0)
	BvhNode4 *nodes4;
	const int2 *frontier;      // (binary node, 4-wide slot), both global indices
	int2 *next;
	uint32_t *counters;        // [level] = frontier length of level + 1, [127] = slots handed out
	uint32_t level, nodeBase;
};

__device__ __forceinline__ float box_area(const float *lo, const float *hi)
{
	const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
	return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ void node_child(const BvhNode &n, int side, float *lo, float *hi, int &link)
{
	if (side == 0) lo[0] = n.a.x, lo[1] = n.a.y, lo[2] = n.a.z, hi[0] = n.a.w, hi[1] = n.b.x, hi[2] = n.b.y, link = n.link.x;
	else lo[0] = n.b.z, lo[1] = n.b.w, lo[2] = n.c.x, hi[0] = n.c.y, hi[1] = n.c.z, hi[2] = n.c.w, link = n.link.y;
}

__global__ void k_collapse4_sah(Collapse4Args a)
{
	const uint32_t nFrontier = a.level == 0 ? 1u : a.counters[a.level - 1];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nFrontier; i += gridDim.x * blockDim.x)
	{
		const int2 item = a.level == 0 ? make_int2((int)a.nodeBase, (int)a.nodeBase) : a.frontier[i];
		float lo[4][3], hi[4][3];
		int link[4];
		const BvhNode root = a.nodes[item.x];
		node_child(root, 0, lo[0], hi[0], link[0]);
		node_child(root, 1, lo[1], hi[1], link[1]);
		int m = 2;
		while (m < 4)
		{
			int best = -1;
			float bestArea = -1.0f;
			for (int k = 0; k < m; ++k)
				if (link[k] >= 0)
				{
					const float ar = box_area(lo[k], hi[k]);
					if (ar > bestArea) bestArea = ar, best = k;
				}
			if (best < 0) break;
			const BvhNode c = a.nodes[link[best]];
			node_child(c, 0, lo[best], hi[best], link[best]);
			node_child(c, 1, lo[m], hi[m], link[m]);
			++m;
		}
		int inner = 0;
		for (int k = 0; k < m; ++k) inner += link[k] >= 0;
		if (inner)
		{
			const uint32_t slot0 = atomicAdd(&a.counters[127], (uint32_t)inner);
			const uint32_t q0 = atomicAdd(&a.counters[a.level], (uint32_t)inner);
			int j = 0;
			for (int k = 0; k < m; ++k)
				if (link[k] >= 0)
				{
					const int slot = (int)(a.nodeBase + 1u + slot0 + (uint32_t)j);
					a.next[q0 + j] = make_int2(link[k], slot);
					link[k] = slot;
					++j;
				}
		}
		// unused slot: a point far away -- |t| ~ 3e38/|d| on every axis always misses
		const float far = 3.0e38f;
		for (int k = m; k < 4; ++k)
			lo[k][0] = lo[k][1] = lo[k][2] = hi[k][0] = hi[k][1] = hi[k][2] = far, link[k] = 0x7FFFFFFF;
		BvhNode4 o;
		o.lox = make_float4(lo[0][0], lo[1][0], lo[2][0], lo[3][0]);
		o.loy = make_float4(lo[0][1], lo[1][1], lo[2][1], lo[3][1]);
		o.loz = make_float4(lo[0][2], lo[1][2], lo[2][2], lo[3][2]);
		o.hix = make_float4(hi[0][0], hi[1][0], hi[2][0], hi[3][0]);
		o.hiy = make_float4(hi[0][1], hi[1][1], hi[2][1], hi[3][1]);
		o.hiz = make_float4(hi[0][2], hi[1][2], hi[2][2], hi[3][2]);
		o.link = make_int4(link[0], link[1], link[2], link[3]);
		o.pad = make_int4(0, 0, 0, 0);
		a.nodes4[item.y] = o;
	}
}


// ---- collapse to 8-wide nodes with quantised child boxes (BvhNode8) ---------------------------------------
// Same top-down scheme as k_collapse4_sah -- a node keeps opening its inner child with the largest surface area -- up to
// eight children, then the child boxes are quantised to 8 bits per plane relative to the node's own box.

// s = 2^e with 250 * s >= extent (five quanta of head-room for the outward rounding and the padding); returns the biased
// exponent byte
__device__ __forceinline__ uint32_t quant_exponent(float extent)
{
	const float want = fmaxf(extent, 1e-30f) * (1.0f / 250.0f);
	int e = (int)((__float_as_uint(want) >> 23) & 0xFFu) + 1;   // 2^(floor(log2 want) + 1) >= want
	e = e < 1 ? 1 : (e > 254 ? 254 : e);
	return (uint32_t)e;
}

// the node record for m <= 8 child boxes (exact floats) and links; boxes are rounded outward and padded by one quantum
__device__ __forceinline__ BvhNode8 quantise_node8(const float (*lo)[3], const float (*hi)[3], const int *link, int m)
{
	const float inf = __int_as_float(0x7f800000);
	float nlo[3] = { inf, inf, inf }, nhi[3] = { -inf, -inf, -inf };
	for (int k = 0; k < m; ++k)
		for (int a = 0; a < 3; ++a)
			nlo[a] = fminf(nlo[a], lo[k][a]), nhi[a] = fmaxf(nhi[a], hi[k][a]);
	BvhNode8 o;
	o.px = nlo[0], o.py = nlo[1], o.pz = nlo[2];
	uint32_t eb[3];
	float inv[3];
	for (int a = 0; a < 3; ++a)
	{
		eb[a] = quant_exponent(nhi[a] - nlo[a]);
		inv[a] = __uint_as_float((254u - eb[a]) << 23);   // 1 / 2^(e - 127), exact
	}
	uint32_t q[6][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
	uint32_t valid = 0;
	for (int k = 0; k < 8; ++k)
	{
		uint32_t ql[3] = { 255u, 255u, 255u }, qh[3] = { 0u, 0u, 0u };   // unused child: an inverted box no ray hits
		if (k < m)
		{
			valid |= 1u << k;
			for (int a = 0; a < 3; ++a)
			{
				// outward rounding with a margin (the quotient itself carries a rounding error), then one quantum of padding:
				// the traversal reconstructs plane distances with an absolute error far below a quantum (rt_traverse.cuh)
				const float xl = (lo[k][a] - nlo[a]) * inv[a], xh = (hi[k][a] - nlo[a]) * inv[a];
				const int il = (int)floorf(xl - 1e-3f) - 1, ih = (int)ceilf(xh + 1e-3f) + 1;
				ql[a] = (uint32_t)(il < 0 ? 0 : (il > 255 ? 255 : il));
				qh[a] = (uint32_t)(ih < 0 ? 0 : (ih > 255 ? 255 : ih));
			}
		}
		const int w = k >> 2, sh = (k & 3) * 8;
		q[0][w] |= ql[0] << sh, q[1][w] |= ql[1] << sh, q[2][w] |= ql[2] << sh;
		q[3][w] |= qh[0] << sh, q[4][w] |= qh[1] << sh, q[5][w] |= qh[2] << sh;
		o.link[k] = k < m ? link[k] : 0x7FFFFFFF;
	}
	o.exyz = eb[0] | (eb[1] << 8) | (eb[2] << 16) | (valid << 24);
	o.qlox[0] = q[0][0], o.qlox[1] = q[0][1], o.qloy[0] = q[1][0], o.qloy[1] = q[1][1], o.qloz[0] = q[2][0], o.qloz[1] = q[2][1];
	o.qhix[0] = q[3][0], o.qhix[1] = q[3][1], o.qhiy[0] = q[4][0], o.qhiy[1] = q[4][1], o.qhiz[0] = q[5][0], o.qhiz[1] = q[5][1];
	return o;
}

// the (padded) box the traversal sees for child k of a node: [p + qlo * s, p + qhi * s]
__device__ __forceinline__ void dequant_child8(const BvhNode8 &n, int k, float *lo, float *hi)
{
	const float s[3] = { __uint_as_float((n.exyz & 0xFFu) << 23), __uint_as_float(((n.exyz >> 8) & 0xFFu) << 23), __uint_as_float(((n.exyz >> 16) & 0xFFu) << 23) };
	const float p[3] = { n.px, n.py, n.pz };
	const uint32_t *ql[3] = { n.qlox, n.qloy, n.qloz }, *qh[3] = { n.qhix, n.qhiy, n.qhiz };
	for (int a = 0; a < 3; ++a)
	{
		const float l = (float)((ql[a][k >> 2] >> ((k & 3) * 8)) & 0xFFu), h = (float)((qh[a][k >> 2] >> ((k & 3) * 8)) & 0xFFu);
		lo[a] = p[a] + l * s[a], hi[a] = p[a] + h * s[a];
		// the sums round to nearest: step outward so that the float box contains the real-valued one
		lo[a] = nextafterf(lo[a], -3.0e38f), hi[a] = nextafterf(hi[a], 3.0e38f);
	}
}

struct Collapse8Args
{
	const BvhNode *nodes;      // binary nodes, links are global indices (leaves < 0)
	BvhNode8 *nodes8;
	const int4 *frontier;      // (binary node, 8-wide slot, stack need of the path so far, -)
	int4 *next;
	uint32_t *counters;        // [level] = frontier length of level + 1, [126] = deepest stack need, [127] = slots handed out
	uint32_t level, nodeBase;
};

__global__ void k_collapse8(Collapse8Args a)
{
	const uint32_t nFrontier = a.level == 0 ? 1u : a.counters[a.level - 1];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nFrontier; i += gridDim.x * blockDim.x)
	{
		const int4 item = a.level == 0 ? make_int4((int)a.nodeBase, (int)a.nodeBase, 0, 0) : a.frontier[i];
		float lo[8][3], hi[8][3];
		int link[8];
		const BvhNode root = a.nodes[item.x];
		node_child(root, 0, lo[0], hi[0], link[0]);
		node_child(root, 1, lo[1], hi[1], link[1]);
		int m = 2;
		while (m < 8)
		{
			int best = -1;
			float bestArea = -1.0f;
			for (int k = 0; k < m; ++k)
				if (link[k] >= 0)
				{
					const float ar = box_area(lo[k], hi[k]);
					if (ar > bestArea) bestArea = ar, best = k;
				}
			if (best < 0) break;
			const BvhNode c = a.nodes[link[best]];
			node_child(c, 0, lo[best], hi[best], link[best]);
			node_child(c, 1, lo[m], hi[m], link[m]);
			++m;
		}
		int inner = 0;
		for (int k = 0; k < m; ++k) inner += link[k] >= 0;
		const int need = item.z + m - 1;   // a step pushes up to m - 1 siblings
		atomicMax(&a.counters[126], (uint32_t)need);
		if (inner)
		{
			const uint32_t slot0 = atomicAdd(&a.counters[127], (uint32_t)inner);
			const uint32_t q0 = atomicAdd(&a.counters[a.level], (uint32_t)inner);
			int j = 0;
			for (int k = 0; k < m; ++k)
				if (link[k] >= 0)
				{
					const int slot = (int)(a.nodeBase + 1u + slot0 + (uint32_t)j);
					a.next[q0 + j] = make_int4(link[k], slot, need, 0);
					link[k] = slot;
					++j;
				}
		}
		a.nodes8[item.y] = quantise_node8(lo, hi, link, m);
	}
}

// refit of one level of the 8-wide tree: child boxes again from the triangle boxes (leaf children) or from the child
// node's own -- already refitted, dequantised -- child boxes; then the node is quantised anew
__global__ void k_refit8_level(BvhNode8 *nodes8, uint32_t begin, uint32_t count, const float4 *box_lo, const float4 *box_hi, const uint32_t *leafOrder)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const BvhNode8 n = nodes8[begin + i];
	const float inf = __int_as_float(0x7f800000);
	float lo[8][3], hi[8][3];
	int link[8], m = 0;
	for (int k = 0; k < 8; ++k)
	{
		if (n.link[k] == 0x7FFFFFFF) continue;   // (valid children are a prefix: m counts them)
		float l[3] = { inf, inf, inf }, h[3] = { -inf, -inf, -inf };
		if (n.link[k] < 0)
		{
			const uint32_t first = ((uint32_t)n.link[k] & 0x7FFFFFFFu) >> 3, cnt = ((uint32_t)n.link[k] & 7u) + 1u;
			for (uint32_t s = first; s < first + cnt; ++s)
			{
				const uint32_t t = leafOrder[s];
				const float4 a = box_lo[t], b = box_hi[t];
				l[0] = fminf(l[0], a.x), l[1] = fminf(l[1], a.y), l[2] = fminf(l[2], a.z), h[0] = fmaxf(h[0], b.x), h[1] = fmaxf(h[1], b.y), h[2] = fmaxf(h[2], b.z);
			}
		}
		else
		{
			const BvhNode8 c = nodes8[n.link[k]];
			for (int j = 0; j < 8; ++j)
				if (c.link[j] != 0x7FFFFFFF)
				{
					float cl[3], ch[3];
					dequant_child8(c, j, cl, ch);
					for (int a = 0; a < 3; ++a) l[a] = fminf(l[a], cl[a]), h[a] = fmaxf(h[a], ch[a]);
				}
		}
		for (int a = 0; a < 3; ++a) lo[m][a] = l[a], hi[m][a] = h[a];
		link[m++] = n.link[k];
	}
	nodes8[begin + i] = quantise_node8(lo, hi, link, m);
}

void rtb_refit8(cudaStream_t st, BvhNode8 *nodes8, uint32_t nodeBase, const uint32_t *levelNodes, uint32_t nLevels,
	const float4 *box_lo, const float4 *box_hi, const uint32_t *leafOrder)
{
	uint32_t begin[129];
	begin[0] = nodeBase;
	for (uint32_t k = 0; k < nLevels; ++k) begin[k + 1] = begin[k] + levelNodes[k];
	for (int k = (int)nLevels - 1; k >= 0; --k)
		if (levelNodes[k])
			k_refit8_level<<<(levelNodes[k] + 63) / 64, 64, 0, st>>>(nodes8, begin[k], levelNodes[k], box_lo, box_hi, leafOrder);
}

int rtb_build(cudaStream_t st, BuildScratch **scratch, const float4 *box_lo, const float4 *box_hi, uint32_t n,
	uint32_t leafSize, BvhNode *nodes, BvhNode4 *nodes4, uint32_t nodeBase, uint32_t leafBase, uint32_t *leafOrder, BvhBuildResult *res, BvhNode8 *nodes8)
{
	res->nLevels = 0, res->maxStack = 0;
	if (n == 0) { res->root = 0, res->nodesUsed = 0, res->depth = 0; return 0; }
	if (leafSize < 1) leafSize = 1;
	if (leafSize > 8) leafSize = 8;
	int rc = ensure_scratch(scratch, n);
	if (rc) return rc;
	BuildScratch *s = *scratch;
	const unsigned blocks = (n + 255) / 256;
	k_init_bounds<<<1, 32, 0, st>>>(s->bounds);
	k_bounds<<<blocks < 1024 ? blocks : 1024, 256, 0, st>>>(box_lo, box_hi, n, s->bounds);
	static const int mortonCube = []{ const char *e = getenv("RT_B200_MORTON"); return (e && !strcmp(e, "axis")) ? 0 : 1; }();
	k_morton<<<blocks, 256, 0, st>>>(box_lo, box_hi, n, s->bounds, s->keysIn, s->valsIn, mortonCube);
	radix_sort_pairs(st, s, n);   // result in keysOut / valsOut
	k_copy_order<<<blocks, 256, 0, st>>>(s->valsOut, n, leafOrder);
	if (n <= leafSize)
	{
		res->root = (int)(0x80000000u | (leafBase << 3) | (n - 1u));
		res->nodesUsed = 0, res->depth = 0;
		return (int)cudaGetLastError();
	}
	k_karras<<<blocks, 256, 0, st>>>(s->keysOut, (int)n, s->children, s->range, s->parentOfInternal, s->parentOfLeaf);
	CK(cudaMemsetAsync(s->flags, 0, sizeof(uint32_t) * n, st));
	RefitArgs a;
	a.box_lo = box_lo, a.box_hi = box_hi, a.sorted = s->valsOut, a.children = s->children, a.range = s->range;
	a.parentOfInternal = s->parentOfInternal, a.parentOfLeaf = s->parentOfLeaf, a.flags = s->flags;
	a.ilo = s->ilo, a.ihi = s->ihi, a.height = s->height, a.nodes = nodes;
	a.nodeBase = nodeBase, a.leafBase = leafBase, a.leafSize = leafSize, a.n = (int)n;
	k_refit<<<blocks, 256, 0, st>>>(a);
	uint32_t depth = 0;
	CK(cudaMemcpyAsync(&depth, s->height, sizeof depth, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	static const int collapseSah = []{ const char *e = getenv("RT_B200_COLLAPSE"); return (e && !strcmp(e, "even")) ? 0 : 1; }();
	if (nodes8)
	{
		// 8-wide quantised nodes (Model BVHs): frontier ping-pong in children / range (int2 pairs reused as int4: n - 1 >= 2 * frontier)
		CK(cudaMemsetAsync(s->flags, 0, sizeof(uint32_t) * 128, st));
		Collapse8Args ca;
		ca.nodes = nodes, ca.nodes8 = nodes8, ca.counters = s->flags, ca.nodeBase = nodeBase;
		const unsigned cblocks = blocks < 2048 ? blocks : 2048;
		for (uint32_t level = 0; level < depth && level < 120; ++level)
		{
			ca.level = level;
			ca.frontier = (const int4 *)((level & 1) ? s->range : s->children);
			ca.next = (int4 *)((level & 1) ? s->children : s->range);
			k_collapse8<<<level < 3 ? 1 : cblocks, 128, 0, st>>>(ca);
		}
		uint32_t counters[128];
		CK(cudaMemcpyAsync(counters, s->flags, sizeof counters, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		res->nLevels = 1, res->levelNodes[0] = 1;
		for (uint32_t level = 0; level < depth && level < 120 && counters[level]; ++level)
			res->levelNodes[res->nLevels++] = counters[level];
		res->maxStack = counters[126];
	}
	else if (!collapseSah)
		k_collapse4<<<blocks, 256, 0, st>>>(nodes, nodes4, s->parentOfInternal, nodeBase, (int)n - 1);
	else
	{
		// children / range / flags are free again after the refit: frontier ping-pong and counters
		CK(cudaMemsetAsync(s->flags, 0, sizeof(uint32_t) * 128, st));
		Collapse4Args ca;
		ca.nodes = nodes, ca.nodes4 = nodes4, ca.counters = s->flags, ca.nodeBase = nodeBase;
		const unsigned cblocks = blocks < 2048 ? blocks : 2048;
		for (uint32_t level = 0; level < depth && level < 120; ++level)   // a 4-wide level consumes at least one binary level
		{
			ca.level = level;
			ca.frontier = (level & 1) ? s->range : s->children;
			ca.next = (level & 1) ? s->children : s->range;
			k_collapse4_sah<<<level < 4 ? 1 : cblocks, 256, 0, st>>>(ca);
		}
		// nodes per 4-wide level, for later refits: level 0 is the root, level k + 1 holds counters[k] nodes
		uint32_t counters[128];
		CK(cudaMemcpyAsync(counters, s->flags, sizeof counters, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		res->nLevels = 1, res->levelNodes[0] = 1;
		for (uint32_t level = 0; level < depth && level < 120 && counters[level]; ++level)
			res->levelNodes[res->nLevels++] = counters[level];
	}
	res->root = (int)nodeBase;
	res->nodesUsed = n - 1;
	res->depth = depth;
	return (int)cudaGetLastError();
}

// ---- refit of the 4-wide tree (position-only edits) --------------------------------------------------
__global__ void k_refit4_level(BvhNode4 *nodes4, uint32_t begin, uint32_t count, const float4 *box_lo, const float4 *box_hi, const uint32_t *leafOrder)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	BvhNode4 n = nodes4[begin + i];
	const int link[4] = { n.link.x, n.link.y, n.link.z, n.link.w };
	float lo[4][3], hi[4][3];
	for (int k = 0; k < 4; ++k)
	{
		const float inf = __int_as_float(0x7f800000);
		float l0 = inf, l1 = inf, l2 = inf, h0 = -inf, h1 = -inf, h2 = -inf;
		if (link[k] == 0x7FFFFFFF)
			l0 = l1 = l2 = h0 = h1 = h2 = 3.0e38f;   // unused slot: stays the far-away point
		else if (link[k] < 0)
		{
			const uint32_t first = ((uint32_t)link[k] & 0x7FFFFFFFu) >> 3, cnt = ((uint32_t)link[k] & 7u) + 1u;
			for (uint32_t s = first; s < first + cnt; ++s)
			{
				const uint32_t t = leafOrder[s];
				const float4 a = box_lo[t], b = box_hi[t];
				l0 = fminf(l0, a.x), l1 = fminf(l1, a.y), l2 = fminf(l2, a.z), h0 = fmaxf(h0, b.x), h1 = fmaxf(h1, b.y), h2 = fmaxf(h2, b.z);
			}
		}
		else
		{
			const BvhNode4 c = nodes4[link[k]];   // a deeper level: already refitted
			const int cl[4] = { c.link.x, c.link.y, c.link.z, c.link.w };
			const float clo[3][4] = { { c.lox.x, c.lox.y, c.lox.z, c.lox.w }, { c.loy.x, c.loy.y, c.loy.z, c.loy.w }, { c.loz.x, c.loz.y, c.loz.z, c.loz.w } };
			const float chi[3][4] = { { c.hix.x, c.hix.y, c.hix.z, c.hix.w }, { c.hiy.x, c.hiy.y, c.hiy.z, c.hiy.w }, { c.hiz.x, c.hiz.y, c.hiz.z, c.hiz.w } };
			for (int j = 0; j < 4; ++j)
				if (cl[j] != 0x7FFFFFFF)
					l0 = fminf(l0, clo[0][j]), l1 = fminf(l1, clo[1][j]), l2 = fminf(l2, clo[2][j]), h0 = fmaxf(h0, chi[0][j]), h1 = fmaxf(h1, chi[1][j]), h2 = fmaxf(h2, chi[2][j]);
		}
		lo[k][0] = l0, lo[k][1] = l1, lo[k][2] = l2, hi[k][0] = h0, hi[k][1] = h1, hi[k][2] = h2;
	}
	n.lox = make_float4(lo[0][0], lo[1][0], lo[2][0], lo[3][0]);
	n.loy = make_float4(lo[0][1], lo[1][1], lo[2][1], lo[3][1]);
	n.loz = make_float4(lo[0][2], lo[1][2], lo[2][2], lo[3][2]);
	n.hix = make_float4(hi[0][0], hi[1][0], hi[2][0], hi[3][0]);
	n.hiy = make_float4(hi[0][1], hi[1][1], hi[2][1], hi[3][1]);
	n.hiz = make_float4(hi[0][2], hi[1][2], hi[2][2], hi[3][2]);
	nodes4[begin + i] = n;
}

void rtb_refit4(cudaStream_t st, BvhNode4 *nodes4, uint32_t nodeBase, const uint32_t *levelNodes, uint32_t nLevels,
	const float4 *box_lo, const float4 *box_hi, const uint32_t *leafOrder)
{
	// slots: level 0 at nodeBase, level k >= 1 behind the levels before it; deepest level first
	uint32_t begin[129];
	begin[0] = nodeBase;
	for (uint32_t k = 0; k < nLevels; ++k) begin[k + 1] = begin[k] + levelNodes[k];
	for (int k = (int)nLevels - 1; k >= 0; --k)
		if (levelNodes[k])
			k_refit4_level<<<(levelNodes[k] + 127) / 128, 128, 0, st>>>(nodes4, begin[k], levelNodes[k], box_lo, box_hi, leafOrder);
}

// ---- leaf-order scatter --------------------------------------------------------------------------

__global__ void k_scatter_tris(const float4 *geomOrig, const uint32_t *leafOrder, uint32_t leafBase, uint32_t origBase, uint32_t n,
	float4 *geomLeaf, uint32_t *triSlot)
{
	const uint32_t sIdx = blockIdx.x * blockDim.x + threadIdx.x;
	if (sIdx >= n) return;
	const uint32_t orig = origBase + leafOrder[sIdx];   // leafOrder holds indices local to this model
	const uint32_t slot = leafBase + sIdx;
	geomLeaf[RT_TRI_F4 * (size_t)slot] = geomOrig[3 * orig];
	geomLeaf[RT_TRI_F4 * (size_t)slot + 1] = geomOrig[3 * orig + 1];
	geomLeaf[RT_TRI_F4 * (size_t)slot + 2] = geomOrig[3 * orig + 2];
	triSlot[orig] = slot;
}

void rtb_scatter_tris(cudaStream_t st, const float4 *geomOrig, const uint32_t *leafOrder, uint32_t leafBase, uint32_t origBase, uint32_t n,
	float4 *geomLeaf, uint32_t *triSlot)
{
	if (n) k_scatter_tris<<<(n + 255) / 256, 256, 0, st>>>(geomOrig, leafOrder, leafBase, origBase, n, geomLeaf, triSlot);
}

__global__ void k_offset_order(uint32_t *order, uint32_t n, uint32_t add)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) order[i] += add;
}

void rtb_offset_order(cudaStream_t st, uint32_t *order, uint32_t n, uint32_t add)
{
	if (n) k_offset_order<<<(n + 255) / 256, 256, 0, st>>>(order, n, add);
}
