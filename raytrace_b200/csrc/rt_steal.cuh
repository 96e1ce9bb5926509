// Model BVH walk with WORK STEALING inside the warp (wave kernels, RT_B200_TRAV=steal).
//
// The voted walk (rt_traverse.cuh) runs a batch of 32 rays until its longest ray is through: on the 1 M-triangle mesh 22
// lanes enter the walk and its node steps run with 11.8 of them, on the refractive 4 M-triangle scene 24 and 9.6 -- most rays
// need 4..15 node visits, a few 30..120 (rays that graze the surface, unoccluded shadow rays), and almost every batch holds one
// of those.  Here a lane whose own work is done takes the top entry of the deepest traversal stack in the warp and walks that
// subtree FOR THE OTHER LANE'S RAY.  Every ray's state that more than one lane may need lives in shared memory: its origin,
// direction, reciprocal direction and skip id (read once by a lane that takes over), and its result --
//   closest hit: a 64-bit key (distance bits << 32 | triangle index) lowered with atomicMin, so that every worker of the ray
//                culls with the nearest distance found so far by any of them;
//   any-hit:     one bit per ray in a mask of occluded rays.
// The FAST contract of rt_traverse.cuh holds as it is: a candidate is `t <= the ray's best distance at that moment`, an exact
// tie (the atomicMin returns a key with the same distance bits and another triangle, or the ray's standing best from an earlier
// scene object) or a hit on the triangle the ray left from raises `slow`, and the final winner goes through verify.  Which lane
// walks which subtree changes no result: boxes are only a conservative cull, distances come from the exact operator.
#pragma once
#include "rt_traverse.cuh"

struct WarpSteal
{
	float ox[32], oy[32], oz[32], dx[32], dy[32], dz[32], ix[32], iy[32], iz[32];   // the 32 rays of the batch
	uint32_t skip[32];
	unsigned long long key[32];   // closest: best distance bits << 32 | triangle (0xFFFFFFFF: none of this Model); any-hit: light distance << 32
	uint32_t slow, occluded;      // bit r: ray r met a tie / its own triangle; ray r is occluded
};

// Called by ALL 32 lanes of a converged warp; `enter` = this lane's ray has to walk the tree.
template<bool ANY>
__device__ __forceinline__ void traverse_steal(const SceneDev &S, RayD ray, F3 idir, bool enter, int root,
	float hr_distance, Best &best, bool &done, WarpSteal &W)
{
	const uint32_t full = 0xffffffffu;
	const uint32_t lane = threadIdx.x & 31u;
	W.ox[lane] = ray.o.x, W.oy[lane] = ray.o.y, W.oz[lane] = ray.o.z;
	W.dx[lane] = ray.d.x, W.dy[lane] = ray.d.y, W.dz[lane] = ray.d.z;
	W.ix[lane] = idir.x, W.iy[lane] = idir.y, W.iz[lane] = idir.z;
	W.skip[lane] = ray.skip;
	W.key[lane] = ((unsigned long long)__float_as_uint(best.t) << 32) | 0xFFFFFFFFull;
	if (lane == 0) W.slow = 0u, W.occluded = 0u;
	__syncwarp();
	const uint32_t *keyHi = (const uint32_t *)W.key + 1;   // little endian: word 2r + 1 = distance bits of ray r
	typename StackSlot<ANY>::T stack[RT_STACK];
	int sp = 0;
	int cur = enter ? root : RT_TRAV_DONE;
	uint32_t own = lane;               // whose ray this lane is walking
	while (true)
	{
		// ---- stealing: one hand-over per step, from the deepest stack to the first idle lane -----------------------------------
		const bool idle = cur == RT_TRAV_DONE && sp == 0;
		const uint32_t mIdle = __ballot_sync(full, idle);
		if (mIdle == full)
			break;
		if (mIdle)
		{
			const int deepest = __reduce_max_sync(full, sp);
			if (deepest > 0)
			{
				const uint32_t donor = __ffs((int)__ballot_sync(full, sp == deepest)) - 1u, taker = __ffs((int)mIdle) - 1u;
				typename StackSlot<ANY>::T e = stack[sp > 0 ? sp - 1 : 0];
				if (lane == donor) --sp;
				const int link = __shfl_sync(full, slot_link(e), donor);
				const float tEntry = __shfl_sync(full, slot_t(e), donor);
				const uint32_t whose = __shfl_sync(full, own, donor);
				if (lane == taker)
				{
					own = whose;
					ray.o = f3(W.ox[own], W.oy[own], W.oz[own]), ray.d = f3(W.dx[own], W.dy[own], W.dz[own]);
					idir = f3(W.ix[own], W.iy[own], W.iz[own]);
					ray.skip = W.skip[own];
					cur = (ANY || tEntry <= __uint_as_float(keyHi[2u * own])) ? link : RT_TRAV_DONE;
				}
			}
		}
		if (ANY && ((*(volatile uint32_t *)&W.occluded >> own) & 1u))
			cur = RT_TRAV_DONE, sp = 0;   // somebody found an occluder of this ray
		const float bestT = __uint_as_float(*(volatile const uint32_t *)&keyHi[2u * own]);   // the ray's nearest distance so far (any-hit: its light distance)
		const bool atNode = cur >= 0, atLeaf = cur < 0 && cur != RT_TRAV_DONE;
		const uint32_t mN = __ballot_sync(full, atNode), mL = __ballot_sync(full, atLeaf);
		if ((mN | mL) == 0u)
		{
			// nobody holds a node or a leaf: whoever still has stack entries pops (guarantees progress), the others wait to take over
			while (sp)
			{
				const typename StackSlot<ANY>::T e = stack[--sp];
				if (ANY || slot_t(e) <= bestT) { cur = slot_link(e); break; }
			}
			continue;
		}
		if (__popc(mN) >= __popc(mL))
		{
			if (atNode)
			{
				const uint32_t sx = __float_as_uint(ray.d.x) >> 31, sy = __float_as_uint(ray.d.y) >> 31, sz = __float_as_uint(ray.d.z) >> 31;
				const uint32_t onx = sx ? 48u : 0u, ony = sy ? 64u : 16u, onz = sz ? 80u : 32u;
				const uint32_t ofx = sx ? 0u : 48u, ofy = sy ? 16u : 64u, ofz = sz ? 32u : 80u;
				const char *n = (const char *)&S.nodes4[cur];
				const float4 nx = ldg4((const float4 *)(n + onx)), ny = ldg4((const float4 *)(n + ony)), nz = ldg4((const float4 *)(n + onz));
				const float4 fx = ldg4((const float4 *)(n + ofx)), fy = ldg4((const float4 *)(n + ofy)), fz = ldg4((const float4 *)(n + ofz));
				const int4 link = __ldg((const int4 *)(n + 96));
				float t0, t1, t2, t3;
				const bool h0 = slab_hit_nf(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, ray.o, idir, bestT, t0);
				const bool h1 = slab_hit_nf(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, ray.o, idir, bestT, t1);
				const bool h2 = slab_hit_nf(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, ray.o, idir, bestT, t2);
				const bool h3 = slab_hit_nf(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, ray.o, idir, bestT, t3);
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
						if (ANY || slot_t(e) <= bestT) { cur = slot_link(e); break; }
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
		else if (atLeaf)
		{
			const uint32_t first = ((uint32_t)cur & 0x7FFFFFFFu) >> 3, count = ((uint32_t)cur & 7u) + 1u;
			float tb = bestT;
			for (uint32_t k = 0; k < count; ++k)
			{
				float4 g0, g1, g2;
				load_tri(S.tri_geom, first + k, g0, g1, g2);
				const float t = triangle_t(ray.o, ray.d, f3(g0), f3(g1), f3(g2), nullptr);
				if (ANY ? !(t < tb) : !(t <= tb && t < 1e20f))
					continue;
				const uint32_t tri = __float_as_uint(g0.w), pinfo = __float_as_uint(g1.w);
				if (ANY)
				{
					// the reference's culling predicate for this ray (Model.cpp:752-768, :775); hr.distance of an any-hit walk = the light distance
					const DevPart &P = S.parts[pinfo >> 8];
					const float4 bmin = __ldg(&P.box_min), bmax = __ldg(&P.box_max);
					uint32_t m;
					const float minist = border_test_ex(ray.o, ray.d, idir, f3(bmin), f3(bmax), &m);
					if (tested_octant(pinfo & 0xFFu, (minist < tb) ? m : 0u, tri, ray.skip) >= 0)
					{
						atomicOr(&W.occluded, 1u << own);
						sp = 0;
						break;
					}
					continue;
				}
				if (is_tri(ray.skip) && (ray.skip & 0x0FFFFFFFu) == tri)
				{
					atomicOr(&W.slow, 1u << own);
					continue;
				}
				const unsigned long long mine = ((unsigned long long)__float_as_uint(t) << 32) | tri;
				const unsigned long long old = atomicMin(&W.key[own], mine);
				if ((uint32_t)(old >> 32) == __float_as_uint(t) && (uint32_t)old != tri)
					atomicOr(&W.slow, 1u << own);   // exact tie: with another triangle, or with the ray's best from an earlier object
				tb = fminf(tb, t);
			}
			cur = RT_TRAV_DONE;
			while (sp)
			{
				const typename StackSlot<ANY>::T e = stack[--sp];
				if (ANY || slot_t(e) <= tb) { cur = slot_link(e); break; }
			}
		}
	}
	__syncwarp();
	if (ANY)
	{
		if ((W.occluded >> lane) & 1u) done = true;
		__syncwarp();
		return;
	}
	bool slow = ((W.slow >> lane) & 1u) != 0u;
	const unsigned long long key = W.key[lane];
	if (!slow && (uint32_t)key != 0xFFFFFFFFu)
	{
		// one replay per ray (verify of the FAST contract), by the ray's own lane with its own ray (read back: this lane may have
		// walked for others since)
		RayD myRay;
		myRay.o = f3(W.ox[lane], W.oy[lane], W.oz[lane]), myRay.d = f3(W.dx[lane], W.dy[lane], W.dz[lane]);
		myRay.skip = W.skip[lane], myRay.mtlrfr = ray.mtlrfr, myRay.type = ray.type, myRay.isInside = ray.isInside;
		const F3 myIdir = f3(W.ix[lane], W.iy[lane], W.iz[lane]);
		const uint32_t tri = (uint32_t)key;
		const uint32_t pinfo = __float_as_uint(__ldg(&S.tri_geom[RT_TRI_F4 * (size_t)__ldg(&S.tri_slot[tri]) + 1]).w);
		PartCache one;
		one.part = 0xFFFFFFFFu, one.mask = 0;
		const uint32_t mask = part_mask(S, myRay, myIdir, pinfo >> 8, hr_distance, one);
		const int oct = tested_octant(pinfo & 0xFFu, mask, tri, myRay.skip);
		if (oct < 0) slow = true;
		else
		{
			best.t = __uint_as_float((uint32_t)(key >> 32));
			best.id = best.newobj = RT_ID_TRI | ((uint32_t)oct << 28) | tri;
		}
	}
	if (slow) best.t = -1.0f;   // tells the caller to redo the item with the immediate replay
	__syncwarp();               // W is re-used by the next walk
}
