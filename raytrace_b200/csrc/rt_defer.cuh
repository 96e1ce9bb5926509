// Model BVH walk with DEFERRED triangle tests (wave kernels).
//
// What the profile of the voted walk (rt_traverse.cuh) shows on the 1 M-triangle mesh: a batch of 32 rays runs its inner-node
// steps with 11.8 and its leaf steps with 6.5 of 32 lanes -- a lane that holds a leaf waits for the vote, a lane that is
// through waits for the batch, and the triangle test (an IEEE division and three cross/dot pairs, ~110 instructions) is the
// most expensive code the lanes run one by one.  The wave kernels are bound by issue slots, so what counts is warp instructions.
//
// Here a lane that reaches a leaf does not test it.  It appends (its lane, triangle slot) to a small per-warp queue in shared
// memory, pops its stack and keeps walking.  Once 32 tests are pending (or nobody has anything left to walk) the WHOLE warp --
// lanes whose ray is through or never entered the Model included -- runs ONE test round: lane j takes entry j, fetches the
// owner's ray with shuffles, runs the bit-exact TriangleTest and hands the result back through shared memory (closest hit:
// atomicMin per owner on the distance bits + a winner word; any-hit: a mask of occluded owners).  Triangle tests then run
// with up to 32 of 32 lanes, and leaf lanes no longer stall the node steps.
//
// A ray walks on with a stale `best.t` until its results arrive: it may visit nodes the voted walk would have culled.  Those
// visits change no result -- boxes are only a conservative cull, every accepted distance comes from the exact operator, and
// the FAST contract of rt_traverse.cuh (tentative nearest hit, one replay per ray, fall back to the immediate replay on a tie
// or on the triangle the ray left from) does not depend on the order in which candidates are seen:
//   * a candidate is "t <= best.t of its owner WHEN THE ROUND RUNS", a superset of what the voted walk accepts;
//   * an exact tie -- with the owner's standing best or between two candidates of one round -- raises `slow`;
//   * the final winner goes through verify (part predicate + octant list) exactly as before.
#pragma once
#include "rt_traverse.cuh"

#define RT_DQ_CAP 128u   // entries per warp: a round leaves < 32 behind, one enqueue step adds <= 64

struct WarpDefer
{
	uint32_t slot[RT_DQ_CAP];   // leaf-order triangle slot of a pending test
	uint8_t owner[RT_DQ_CAP];   // lane whose ray wants it
	uint32_t bestT[32];         // closest hit: smallest candidate distance per owner in the running round (float bits), ~0 = none
	uint32_t win[32];           // lane + 1 of the entry that delivered it
	uint32_t slow;              // owners that met a tie / the triangle they left from
};

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// Called by ALL 32 lanes of a converged warp; `enter` = this lane's ray has to walk the tree.
template<bool ANY>
__device__ __forceinline__ void traverse_deferred(const SceneDev &S, const RayD &ray, const F3 &idir, bool enter, int root,
	float hr_distance, Best &best, bool &done, WarpDefer &W)
{
	const uint32_t full = 0xffffffffu;
	const uint32_t lane = threadIdx.x & 31u, lt = lanemask_lt();
	bool slow = false;
	const uint32_t idBefore = best.id;
	uint32_t pinfoWin = 0xFFFFFFFFu;   // closest: part << 8 | octants of the tentative winner
	typename StackSlot<ANY>::T stack[RT_STACK];
	int sp = 0;
	int cur = enter ? root : RT_TRAV_DONE;
	uint32_t qhead = 0, qcount = 0;    // warp-uniform
	bool pend = false;                 // this lane has tests in the queue: whatever it walks now, it walks with a stale best.t
	if (!ANY)
	{
		W.bestT[lane] = 0xFFFFFFFFu, W.win[lane] = 0u;
		if (lane == 0) W.slow = 0u;
	}
	__syncwarp();
	const uint32_t sx = __float_as_uint(ray.d.x) >> 31, sy = __float_as_uint(ray.d.y) >> 31, sz = __float_as_uint(ray.d.z) >> 31;
	const uint32_t onx = sx ? 48u : 0u, ony = sy ? 64u : 16u, onz = sz ? 80u : 32u;
	const uint32_t ofx = sx ? 0u : 48u, ofy = sy ? 16u : 64u, ofz = sz ? 32u : 80u;
	const uint32_t trigger = S.dq_trigger;
	while (true)
	{
		const bool atLeaf = cur < 0 && cur != RT_TRAV_DONE;
		const uint32_t mL = __ballot_sync(full, atLeaf);
		if (mL)
		{
			// ---- enqueue: up to two triangles per leaf lane and step, then on to the next subtree -----------------------
			const uint32_t first = ((uint32_t)cur & 0x7FFFFFFFu) >> 3, count = ((uint32_t)cur & 7u) + 1u;
			const bool two = atLeaf && count >= 2u;
			const uint32_t m2 = __ballot_sync(full, two);
			if (atLeaf)
			{
				const uint32_t pos = qhead + qcount + __popc(mL & lt) + __popc(m2 & lt);
				W.slot[pos & (RT_DQ_CAP - 1u)] = first, W.owner[pos & (RT_DQ_CAP - 1u)] = (uint8_t)lane;
				pend = true;
				if (two) W.slot[(pos + 1u) & (RT_DQ_CAP - 1u)] = first + 1u, W.owner[(pos + 1u) & (RT_DQ_CAP - 1u)] = (uint8_t)lane;
				if (count > 2u)
					cur = (int)(0x80000000u | ((first + 2u) << 3) | (count - 3u));
				else
				{
					cur = RT_TRAV_DONE;
					while (sp)
					{
						const typename StackSlot<ANY>::T e = stack[--sp];
						if (ANY || slot_t(e) <= best.t) { cur = slot_link(e); break; }
					}
				}
			}
			qcount += __popc(mL) + __popc(m2);
			__syncwarp();
		}
		// ---- test rounds: when enough tests are pending, or when nobody has anything left to walk --------------------------
		// A round also starts when (almost) every lane that still walks is waiting for results: its steps would be speculative.
		const uint32_t mWork = __ballot_sync(full, cur != RT_TRAV_DONE);
		if (mWork == 0u && qcount == 0u)
			break;
		const uint32_t mFresh = __ballot_sync(full, cur != RT_TRAV_DONE && !pend);
		while (qcount >= 32u || (qcount != 0u && (uint32_t)__popc(mFresh) * trigger <= (uint32_t)__popc(mWork)))
		{
			const uint32_t n = qcount < 32u ? qcount : 32u;
			const bool have = lane < n;
			const uint32_t idx = (qhead + lane) & (RT_DQ_CAP - 1u);
			const uint32_t slot = have ? W.slot[idx] : 0u;
			const uint32_t own = have ? (uint32_t)W.owner[idx] : lane;
			qhead += n, qcount -= n;
			if (qcount == 0u) pend = false;
			const F3 oo = f3(__shfl_sync(full, ray.o.x, own), __shfl_sync(full, ray.o.y, own), __shfl_sync(full, ray.o.z, own));
			const F3 od = f3(__shfl_sync(full, ray.d.x, own), __shfl_sync(full, ray.d.y, own), __shfl_sync(full, ray.d.z, own));
			const float obt = __shfl_sync(full, best.t, own);
			const uint32_t oskip = __shfl_sync(full, ray.skip, own);
			bool cand = false;
			float t = 1e20f;
			uint32_t tri = 0, pinfo = 0;
			if (have)
			{
				float4 g0, g1, g2;
				load_tri(S.tri_geom, slot, g0, g1, g2);
				t = triangle_t(oo, od, f3(g0), f3(g1), f3(g2), nullptr);
				tri = __float_as_uint(g0.w), pinfo = __float_as_uint(g1.w);
				cand = ANY ? (t < obt) : (t <= obt && t < 1e20f);
			}
			if (ANY)
			{
				// the reference's culling predicate for the owner's ray (Model.cpp:752-768, :775), evaluated by the testing lane
				bool acc = false;
				if (cand)
				{
					const F3 oid = f3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
					const DevPart &P = S.parts[pinfo >> 8];
					const float4 bmin = __ldg(&P.box_min), bmax = __ldg(&P.box_max);
					uint32_t m;
					const float minist = border_test_ex(oo, od, oid, f3(bmin), f3(bmax), &m);
					// hr.distance of an any-hit walk is the light distance the owner started with (= its best.t: never lowered)
					acc = tested_octant(pinfo & 0xFFu, (minist < obt) ? m : 0u, tri, oskip) >= 0;
				}
				const uint32_t occl = __reduce_or_sync(full, acc ? (1u << own) : 0u);
				if ((occl >> lane) & 1u)
					done = true, sp = 0, cur = RT_TRAV_DONE;   // occluded: nothing else to look at
			}
			else
			{
				const bool tie = cand && (t == obt || (is_tri(oskip) && (oskip & 0x0FFFFFFFu) == tri));
				if (tie) atomicOr(&W.slow, 1u << own);
				const bool c2 = cand && !tie;
				if (c2) atomicMin(&W.bestT[own], __float_as_uint(t));
				__syncwarp();
				if (c2 && W.bestT[own] == __float_as_uint(t))
					if (atomicExch(&W.win[own], lane + 1u) != 0u)
						atomicOr(&W.slow, 1u << own);   // two candidates at the same distance
				__syncwarp();
				const uint32_t bt = W.bestT[lane], w = W.win[lane];
				const uint32_t src = w ? w - 1u : lane;
				const uint32_t wtri = __shfl_sync(full, tri, src), wpinfo = __shfl_sync(full, pinfo, src);
				if (bt != 0xFFFFFFFFu)
				{
					best.t = __uint_as_float(bt);
					best.id = RT_ID_TRI | wtri;   // octant filled in below
					pinfoWin = wpinfo;
					W.bestT[lane] = 0xFFFFFFFFu, W.win[lane] = 0u;
				}
				__syncwarp();
			}
		}
		// ---- one inner-node step for every lane that holds a node ------------------------------------------------------------
		if (cur >= 0)
		{
			const char *n = (const char *)&S.nodes4[cur];
			const float4 nx = ldg4((const float4 *)(n + onx)), ny = ldg4((const float4 *)(n + ony)), nz = ldg4((const float4 *)(n + onz));
			const float4 fx = ldg4((const float4 *)(n + ofx)), fy = ldg4((const float4 *)(n + ofy)), fz = ldg4((const float4 *)(n + ofz));
			const int4 link = __ldg((const int4 *)(n + 96));
			float t0, t1, t2, t3;
			const bool h0 = slab_hit_nf(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, ray.o, idir, best.t, t0);
			const bool h1 = slab_hit_nf(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, ray.o, idir, best.t, t1);
			const bool h2 = slab_hit_nf(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, ray.o, idir, best.t, t2);
			const bool h3 = slab_hit_nf(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, ray.o, idir, best.t, t3);
			const float inf = __int_as_float(0x7f800000);
			float bt = h0 ? t0 : inf;
			int bi = 0;
			if (h1 && t1 < bt) bt = t1, bi = 1;
			if (h2 && t2 < bt) bt = t2, bi = 2;
			if (h3 && t3 < bt) bt = t3, bi = 3;
			if (!(h0 | h1 | h2 | h3))
			{
				cur = RT_TRAV_DONE;
				while (sp)
				{
					const typename StackSlot<ANY>::T e = stack[--sp];
					if (ANY || slot_t(e) <= best.t) { cur = slot_link(e); break; }
				}
			}
			else
			{
				if (h0 && bi != 0) slot_put(stack[sp++], link.x, t0);
				if (h1 && bi != 1) slot_put(stack[sp++], link.y, t1);
				if (h2 && bi != 2) slot_put(stack[sp++], link.z, t2);
				if (h3 && bi != 3) slot_put(stack[sp++], link.w, t3);
				cur = bi == 0 ? link.x : bi == 1 ? link.y : bi == 2 ? link.z : link.w;
			}
		}
	}
	if (ANY)
		return;
	if ((W.slow >> lane) & 1u)
		slow = true;
	if (!slow && best.id != idBefore)
	{
		// one replay per ray (verify of the FAST contract)
		const uint32_t tri = best.id & 0x0FFFFFFFu;
		PartCache one;
		one.part = 0xFFFFFFFFu, one.mask = 0;
		const uint32_t mask = part_mask(S, ray, idir, pinfoWin >> 8, hr_distance, one);
		const int oct = tested_octant(pinfoWin & 0xFFu, mask, tri, ray.skip);
		if (oct < 0) slow = true;
		else best.id = best.newobj = RT_ID_TRI | ((uint32_t)oct << 28) | tri;
	}
	if (slow) best.t = -1.0f;   // tells the caller to redo the item with the immediate replay
	__syncwarp();
}

#include "rt_steal.cuh"

// the Model walk of a whole converged warp, by the kind of per-warp workspace it is given
template<bool ANY>
__device__ __forceinline__ void traverse_warp(const SceneDev &S, const RayD &ray, const F3 &idir, bool enter, int root, float hr_distance, Best &best, bool &done, WarpDefer &W)
{
	traverse_deferred<ANY>(S, ray, idir, enter, root, hr_distance, best, done, W);
}
template<bool ANY>
__device__ __forceinline__ void traverse_warp(const SceneDev &S, const RayD &ray, const F3 &idir, bool enter, int root, float hr_distance, Best &best, bool &done, WarpSteal &W)
{
	traverse_steal<ANY>(S, ray, idir, enter, root, hr_distance, best, done, W);
}

// trace_scene for a whole converged warp: lanes without a ray pass valid = false and only help (test rounds / stolen subtrees).
template<bool ANY, class WS>
__device__ __forceinline__ void trace_scene_warp(const SceneDev &S, const RayD &ray, bool valid, Best &best, bool &done, WS &W)
{
	const F3 idir = f3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
	TravStats st = { 0, 0, 0 };
	for (uint32_t i = 0; i < S.n_items; ++i)
	{
		const SceneItem it = S.items[i];
		const bool live = valid && !(ANY && done);
		if (it.kind == RT_ITEM_MODEL && !S.brute)
		{
			const DevModel &M = S.models[it.first];
			bool enter = live;
			if (enter)
			{
				const float4 mn = __ldg(&M.border_min), mx = __ldg(&M.border_max);
				// Model.cpp:752: `if (BorderTest(ray, BorderMin, BorderMax) < hr.distance)`
				enter = border_test(ray.o, ray.d, idir, f3(mn), f3(mx)) < best.t;
			}
			if (__ballot_sync(0xffffffffu, enter) == 0u)
				continue;
			const Best before = best;
			traverse_warp<ANY>(S, ray, idir, enter, it.root, before.t, best, done, W);
			if (!ANY && enter && best.t < 0.0f)
			{
				const uint32_t tb = __ldg(&M.tri_begin);
				best = before;
				traverse<false, true, false, false>(S, ray, idir, it.root, before.t, tb, tb + __ldg(&M.tri_count), best, done, st);
			}
		}
		else if (!live)
			continue;
		else if (it.kind == RT_ITEM_PRIM)
			test_prim<ANY>(S, ray, it.first, false, best, done);
		else if (it.kind == RT_ITEM_PRIMBVH && S.brute)
		{
			for (uint32_t p = it.first; p < it.first + it.count && !(ANY && done); ++p)
				test_prim<ANY>(S, ray, p, false, best, done);
		}
		else if (it.kind == RT_ITEM_PRIMBVH)
		{
			const uint32_t end = it.first + it.count;
			if (!ANY && ray.isInside && !(ray.skip & RT_ID_TRI) && ray.skip >= it.first && ray.skip < end)
			{
				traverse<ANY, false, false, false>(S, ray, idir, it.root, best.t, it.first, end, best, done, st, it.first, ray.skip);
				test_prim<ANY>(S, ray, ray.skip, false, best, done);
				traverse<ANY, false, false, false>(S, ray, idir, it.root, best.t, it.first, end, best, done, st, ray.skip + 1u, end);
			}
			else
				traverse<ANY, false, false, false>(S, ray, idir, it.root, best.t, it.first, end, best, done, st);
		}
		else
		{
			// RT_FLAG_BRUTE: every triangle of the model with the immediate culling replay, no BVH
			const DevModel &M = S.models[it.first];
			const float4 mn = __ldg(&M.border_min), mx = __ldg(&M.border_max);
			if (!(border_test(ray.o, ray.d, idir, f3(mn), f3(mx)) < best.t))
				continue;
			const uint32_t tb = __ldg(&M.tri_begin), te = tb + __ldg(&M.tri_count);
			PartCache pc;
			pc.part = 0xFFFFFFFFu, pc.mask = 0;
			bool slow = false;
			leaf_tris<ANY, false, false>(S, ray, idir, tb, te - tb, best.t, tb, te, pc, best, done, slow, st);
		}
	}
}
