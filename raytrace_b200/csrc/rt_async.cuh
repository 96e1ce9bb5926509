// Lane-asynchronous scene walk: every lane of a warp owns ONE ray at a time and walks the scene's item list
// (RayTracer.cpp:458-465 / :513-520) and its BVHs at its own pace; a lane whose ray is through takes the next ray
// out of a per-warp pool in shared memory a few instructions later, instead of idling until the slowest of its 32
// batch mates is done.
//
// Why: ncu on the batch-synchronous walk (profiles/r1i_ncu_full_k_wave_batch_c3.md) showed 8-20 of 32 lanes active
// per instruction.  Two thirds of the loss is batch synchrony: rays that miss the Model's box or leave the BVH after
// two nodes wait for the ray that skims the height field for 100 nodes (sum of node visits / 32 x longest lane = 0.59).
// Replacement inside the walk failed twice in round 1 because the swap-in (claim a ray with an atomic, load or
// generate it, publish the old one) ran with the one or two lanes that needed it.  Here both ends are batched:
//   * rays are claimed and set up 32 at a time by the whole warp (refill) into the pool, a swap-in is three LDS.128;
//   * a finished ray's result goes back into its pool entry (one STS.128) and the long, divergent epilogue -- surface
//     attributes, child rays, queue appends -- runs later for 32 finished rays at once (service).
// The three per-lane phases -- inner node, leaf, item walk -- are voted per step like in rt_traverse.cuh (majority runs,
// the minority keeps its state for a later step).
//
// Exactness: the per-ray arithmetic is the one of rt_traverse.cuh / rt_intersect.cuh (same operators, same
// culling replay, same tie rules); only the order in which rays are processed changes, and no result depends on it.
#pragma once
#include "rt_traverse.cuh"

#ifndef RT_POOL
#define RT_POOL 64          // pool entries per warp: <= 32 in flight + <= 32 ready or finished
#endif
#ifndef RT_REFILL_IDLE
#define RT_REFILL_IDLE 4    // service (epilogue + refill) when the pool is empty and this many lanes have nothing to do
#endif

struct PoolEntry   // 64 bytes
{
	float4 o;                    // origin.xyz | Ray::mtlrfr          (shadow rays: light distance)
	float4 d;                    // direction.xyz | bwc
	uint32_t skip, meta, slot;   // HitRes::obj to skip | type + isInside << 8 | ray slot in its level (shadow rays: destination index)
	float t;                     // result: HitRes::distance
	uint32_t id, newobj;         // result: closest primitive, identity for child rays
	uint32_t pad0, pad1;
};

struct WarpPool
{
	PoolEntry e[RT_POOL];
	uint8_t readyQ[RT_POOL], doneQ[RT_POOL], freeQ[RT_POOL];   // stacks of entry indices
	uint32_t nReady, nDone, nFree, pad;
};

// per-lane flags
#define LF_HAS    0x01u   // the lane holds a ray
#define LF_INBVH  0x02u   // ... and is inside (or has just left) the BVH of item `item`
#define LF_TRI    0x04u   // that BVH is a Model (triangle leaves)
#define LF_FAST   0x08u   // closest-hit Model walk with the postponed culling replay (rt_traverse.cuh FAST)
#define LF_SLOW   0x10u   // a candidate the postponed check cannot decide was seen
#define LF_PH0    0x20u   // prim run walked in order-respecting phases (ray starts inside one of its spheres): before / after that sphere
#define LF_PH2    0x40u
#define LF_OCCL   0x80u   // any-hit: occluded

struct Lane
{
	F3 o, d, idir;
	uint32_t skip, flags;        // flags: LF_* | Ray::type << 8 | Ray::isInside << 16
	float bt;                    // Best
	uint32_t bid, bnew;
	int cur, sp;
	uint32_t item;
	float beforeT;               // best.t when the current Model was entered (hr.distance of Model::intersect)
	uint32_t idBefore;
	uint32_t auxA, auxB;         // Model: PartCache (part, mask) | prim run: window [lo, hi)
	uint32_t entry;              // pool entry of the ray (closest hit) / destination index (shadow)
};

__device__ __forceinline__ RayD lane_ray(const Lane &L)
{
	RayD r;
	r.o = L.o, r.d = L.d, r.mtlrfr = 1.0f, r.skip = L.skip;
	r.type = (uint8_t)(L.flags >> 8), r.isInside = (uint8_t)(L.flags >> 16);
	return r;
}

// closest hit keeps the entry distance beside the link (culled again on pop), any-hit only the link
template<bool ANY> struct StackOf { typedef uint2 T; };
template<> struct StackOf<true> { typedef int T; };

template<bool ANY> __device__ __forceinline__ void stack_pop(Lane &L, const typename StackOf<ANY>::T *stack);
template<> __device__ __forceinline__ void stack_pop<true>(Lane &L, const int *stack)
{
	L.cur = L.sp ? stack[--L.sp] : RT_TRAV_DONE;
}
template<> __device__ __forceinline__ void stack_pop<false>(Lane &L, const uint2 *stack)
{
	L.cur = RT_TRAV_DONE;
	while (L.sp)
	{
		const uint2 e = stack[--L.sp];
		if (__uint_as_float(e.y) <= L.bt) { L.cur = (int)e.x; break; }
	}
}
__device__ __forceinline__ void stack_push(Lane &L, int *stack, int link, float) { stack[L.sp++] = link; }
__device__ __forceinline__ void stack_push(Lane &L, uint2 *stack, int link, float t) { stack[L.sp++] = make_uint2((uint32_t)link, __float_as_uint(t)); }

// ---- one inner-node step (same box arithmetic as traverse() in rt_traverse.cuh) ----------------------------------
template<bool ANY, bool STATS>
__device__ __forceinline__ void node_step(const SceneDev &S, Lane &L, typename StackOf<ANY>::T *stack, TravStats &st)
{
	const uint32_t sx = __float_as_uint(L.d.x) >> 31, sy = __float_as_uint(L.d.y) >> 31, sz = __float_as_uint(L.d.z) >> 31;
	const uint32_t onx = sx ? 48u : 0u, ony = sy ? 64u : 16u, onz = sz ? 80u : 32u;
	const uint32_t ofx = sx ? 0u : 48u, ofy = sy ? 16u : 64u, ofz = sz ? 32u : 80u;
	const char *n = (const char *)&S.nodes4[L.cur];
	const float4 nx = ldg4((const float4 *)(n + onx)), ny = ldg4((const float4 *)(n + ony)), nz = ldg4((const float4 *)(n + onz));
	const float4 fx = ldg4((const float4 *)(n + ofx)), fy = ldg4((const float4 *)(n + ofy)), fz = ldg4((const float4 *)(n + ofz));
	const int4 link = __ldg((const int4 *)(n + 96));
	if (STATS) ++st.nodes;
	float t0, t1, t2, t3;
	const bool h0 = slab_hit_nf(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, L.o, L.idir, L.bt, t0);
	const bool h1 = slab_hit_nf(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, L.o, L.idir, L.bt, t1);
	const bool h2 = slab_hit_nf(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, L.o, L.idir, L.bt, t2);
	const bool h3 = slab_hit_nf(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, L.o, L.idir, L.bt, t3);
	if (!(h0 | h1 | h2 | h3))
	{
		stack_pop<ANY>(L, stack);
		return;
	}
	// nearest hit child first, the other hit children go on the stack
	const float inf = __int_as_float(0x7f800000);
	float bt = h0 ? t0 : inf;
	int bi = 0;
	if (h1 && t1 < bt) bt = t1, bi = 1;
	if (h2 && t2 < bt) bt = t2, bi = 2;
	if (h3 && t3 < bt) bt = t3, bi = 3;
	if (h0 && bi != 0) stack_push(L, stack, link.x, t0);
	if (h1 && bi != 1) stack_push(L, stack, link.y, t1);
	if (h2 && bi != 2) stack_push(L, stack, link.z, t2);
	if (h3 && bi != 3) stack_push(L, stack, link.w, t3);
	L.cur = bi == 0 ? link.x : bi == 1 ? link.y : bi == 2 ? link.z : link.w;
}

// ---- one leaf step -----------------------------------------------------------------------------------------------
template<bool ANY, bool STATS>
__device__ __forceinline__ void leaf_step(const SceneDev &S, Lane &L, typename StackOf<ANY>::T *stack, TravStats &st)
{
	const uint32_t first = ((uint32_t)L.cur & 0x7FFFFFFFu) >> 3, count = ((uint32_t)L.cur & 7u) + 1u;
	const RayD ray = lane_ray(L);
	Best best = { L.bt, L.bid, L.bnew };
	bool done = false;
	if (L.flags & LF_TRI)
	{
		PartCache pc;
		pc.part = L.auxA, pc.mask = L.auxB;
		bool slow = false;
		if (ANY)
			leaf_tris<true, false, STATS>(S, ray, L.idir, first, count, L.bt, 0u, 0u, pc, best, done, slow, st);
		else if (L.flags & LF_FAST)
			leaf_tris<false, true, STATS>(S, ray, L.idir, first, count, L.beforeT, 0u, 0u, pc, best, done, slow, st);
		else
		{
			// immediate culling replay (rare: only after the postponed check gave up on this ray and Model)
			const DevModel &M = S.models[S.items[L.item].first];
			const uint32_t tb = __ldg(&M.tri_begin), te = tb + __ldg(&M.tri_count);
			leaf_tris<false, false, STATS>(S, ray, L.idir, first, count, L.beforeT, tb, te, pc, best, done, slow, st);
		}
		L.auxA = pc.part, L.auxB = pc.mask;
		if (slow) L.flags |= LF_SLOW;
	}
	else
	{
		const uint32_t rangeBegin = __ldg(&S.items[L.item].first);
		for (uint32_t k = 0; k < count; ++k)
		{
			const uint32_t p = __ldg(&S.bvh_prims[first + k]);
			if (p < L.auxA || p >= L.auxB)
				continue;
			if (STATS) ++st.prims;
			test_prim<ANY>(S, ray, p, !(best.id & RT_ID_TRI) && best.id >= rangeBegin, best, done);
		}
	}
	L.bt = best.t, L.bid = best.id, L.bnew = best.newobj;
	if (ANY && done)
	{
		L.flags |= LF_OCCL;
		L.sp = 0;
		L.cur = RT_TRAV_DONE;
		return;
	}
	stack_pop<ANY>(L, stack);
}

// ---- the item walk: everything between two BVH walks of one ray ------------------------------------------------------
// Called for a lane whose `cur` is RT_TRAV_DONE: a ray that has just been taken (item = 0, LF_INBVH clear) or that has
// just left the BVH of item `item`.  Returns true when the ray is finished (closest: L.bt/bid/bnew final; any-hit:
// LF_OCCL says occluded); otherwise the lane is at the root of the next BVH.
template<bool ANY, bool STATS>
__device__ __forceinline__ bool advance_items(const SceneDev &S, Lane &L, TravStats &st)
{
	const RayD ray = lane_ray(L);
	Best best = { L.bt, L.bid, L.bnew };
	bool done = (L.flags & LF_OCCL) != 0;
	bool finished = false;
	if (L.flags & LF_INBVH)
	{
		L.flags &= ~LF_INBVH;
		if (ANY && done)
			return true;
		const SceneItem it = S.items[L.item];
		if (L.flags & LF_TRI)
		{
			if (!ANY && (L.flags & LF_FAST))
			{
				// the one postponed replay of the reference's culling predicate (verify step of traverse<FAST>)
				bool slow = (L.flags & LF_SLOW) != 0;
				if (!slow && best.id != L.idBefore)
				{
					const uint32_t pinfo = L.auxA, tri = best.id & 0x0FFFFFFFu;
					PartCache one;
					one.part = 0xFFFFFFFFu, one.mask = 0;
					const uint32_t mask = part_mask(S, ray, L.idir, pinfo >> 8, L.beforeT, one);
					const int oct = tested_octant(pinfo & 0xFFu, mask, tri, ray.skip);
					if (oct < 0) slow = true;
					else best.id = best.newobj = RT_ID_TRI | ((uint32_t)oct << 28) | tri;
				}
				L.flags &= ~(LF_FAST | LF_SLOW);
				if (slow)
				{
					// redo this Model with the immediate per-candidate replay
					L.bt = L.beforeT, L.bid = L.idBefore;   // (the FAST walk never touches newobj)
					L.auxA = 0xFFFFFFFFu, L.auxB = 0;
					L.cur = it.root, L.sp = 0;
					L.flags |= LF_INBVH;
					return false;
				}
			}
		}
		else if (L.flags & LF_PH0)
		{
			// ray starts inside sphere `skip` of this run: the sphere itself, then the primitives after it
			L.flags &= ~LF_PH0;
			if (STATS) ++st.prims;
			test_prim<ANY>(S, ray, ray.skip, false, best, done);
			L.bt = best.t, L.bid = best.id, L.bnew = best.newobj;
			L.auxA = ray.skip + 1u, L.auxB = it.first + it.count;
			L.cur = it.root, L.sp = 0;
			L.flags |= LF_INBVH | LF_PH2;
			return false;
		}
		else
			L.flags &= ~LF_PH2;
		++L.item;
	}
	while (true)
	{
		if (L.item >= S.n_items) { finished = true; break; }
		const SceneItem it = S.items[L.item];
		if (it.kind == RT_ITEM_PRIM)
		{
			if (STATS) ++st.prims;
			test_prim<ANY>(S, ray, it.first, false, best, done);
			if (ANY && done) { finished = true; break; }
			++L.item;
			continue;
		}
		if (it.kind == RT_ITEM_PRIMBVH)
		{
			const uint32_t end = it.first + it.count;
			L.flags &= ~LF_TRI;
			if (!ANY && ray.isInside && !(ray.skip & RT_ID_TRI) && ray.skip >= it.first && ray.skip < end)
				L.auxA = it.first, L.auxB = ray.skip, L.flags |= LF_PH0;
			else
				L.auxA = 0u, L.auxB = 0xFFFFFFFFu;
			L.cur = it.root, L.sp = 0;
			L.flags |= LF_INBVH;
			break;
		}
		// Model.cpp:752: `if (BorderTest(ray, BorderMin, BorderMax) < hr.distance)`
		const DevModel &M = S.models[it.first];
		const float4 mn = __ldg(&M.border_min), mx = __ldg(&M.border_max);
		if (it.count == 0u || !(border_test(ray.o, ray.d, L.idir, f3(mn), f3(mx)) < best.t))
		{
			++L.item;
			continue;
		}
		L.beforeT = best.t, L.idBefore = best.id;
		L.auxA = 0xFFFFFFFFu, L.auxB = 0;
		L.cur = it.root, L.sp = 0;
		L.flags |= LF_INBVH | LF_TRI | (ANY ? 0u : LF_FAST);
		break;
	}
	L.bt = best.t, L.bid = best.id, L.bnew = best.newobj;
	if (ANY && done) L.flags |= LF_OCCL;
	return finished;
}

// pool bookkeeping helpers (warp-converged callers)
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ void pool_init(WarpPool &P)
{
	const uint32_t lane = threadIdx.x & 31u;
	for (uint32_t i = lane; i < RT_POOL; i += 32u) P.freeQ[i] = (uint8_t)i;
	if (lane == 0) P.nReady = 0, P.nDone = 0, P.nFree = RT_POOL;
	__syncwarp();
}

__device__ __forceinline__ void lane_take(Lane &L, const PoolEntry &E, uint32_t entry)
{
	const float4 o = E.o, d = E.d;
	L.o = f3(o), L.d = f3(d);
	L.idir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	L.skip = E.skip;
	L.flags = LF_HAS | ((E.meta & 0xFFFFu) << 8);
	L.cur = RT_TRAV_DONE, L.sp = 0, L.item = 0;
	L.entry = entry;
}
