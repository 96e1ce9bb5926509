// Software BVH traversal on the SM (B200 exposes no RT cores to CUDA) and the scene-order walk
// that replaces the reference's per-ray virtual dispatch loop (RayTracer.cpp:455-465 closest hit,
// :510-520 shadow any-hit) and Model::intersect's brute-force part/octant/triangle loops
// (Model.cpp:748-811).
//
// Exactness contract:
//  * every accepted hit distance comes from the bit-exact operators of rt_intersect.cuh;
//  * BVH boxes are only a conservative cull (padded boxes, widened slab interval);
//  * a triangle hit is accepted only if the reference would have tested that triangle for this
//    ray: its Model passes BorderTest (< hr.distance), its part passes BorderTestEx
//    (< hr.distance), and one of the octant lists that hold a copy of it is enabled in the ray's
//    octant mask -- excluding the single copy the ray left from (the reference's self-skip
//    compares clTri addresses, Model.cpp:775);
//  * equal distances resolve to the first candidate in reference iteration order
//    (object, part, first enabled octant, triangle index), independent of traversal order.
#pragma once
#include "rt_intersect.cuh"

#define RT_BLOCK 128
#ifndef RT_CULL_ON_POP
#define RT_CULL_ON_POP 1
#endif
#ifndef RT_CTAS_PER_SM
#define RT_CTAS_PER_SM 8
#endif

struct TravStats { uint32_t nodes, tris, prims; };

__device__ __forceinline__ bool is_tri(uint32_t id) { return id != RT_ID_NONE && (id & RT_ID_TRI); }

struct Best
{
	float t;        // HitRes::distance so far
	uint32_t id;    // RT_ID_NONE | prim flat index | RT_ID_TRI | oct<<28 | tri original index
	// `newobj` of RayTracer.cpp:456-465: identity handed to shadow/secondary rays.  It follows the
	// accepted hits EXCEPT a sphere's own inside-exit hit (Basic3DObject.cpp:153 returns obj == hr.obj,
	// so `if (hr.obj != basehr.obj) newobj = hr.obj` leaves it at the previous winner).
	uint32_t newobj;
};

// Conservative ray/box interval test: (plane - o) * (1/d) has bounded relative error, the far
// bound is widened by a few ulp (Ize 2013) and the boxes themselves are padded at build time.
__device__ __forceinline__ bool slab_hit(float lx, float ly, float lz, float hx, float hy, float hz,
	const F3 &o, const F3 &id, float tbest, float &tnear)
{
	const float ax = (lx - o.x) * id.x, bx = (hx - o.x) * id.x;
	const float ay = (ly - o.y) * id.y, by = (hy - o.y) * id.y;
	const float az = (lz - o.z) * id.z, bz = (hz - o.z) * id.z;
	const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
	const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tbest));
	tnear = t0;
	return t0 <= t1 * 1.0000005f;
}

// 4-wide variant: the near/far plane of every axis is picked by the SIGN of the ray direction when
// the node is loaded (per-ray byte offsets into the SoA node), so no per-axis min/max is needed --
// that moves ~24 instructions per node off the ALU pipe, the busiest pipe of this kernel.
// d = +-0 gives +-inf reciprocals: inside the slab the products are -inf/+inf (no constraint),
// outside they are +inf/-inf (miss), and 0*inf = NaN drops out of fmaxf/fminf (no constraint).
__device__ __forceinline__ bool slab_hit_nf(float nx, float ny, float nz, float fx, float fy, float fz,
	const F3 &o, const F3 &id, float tbest, float &tnear)
{
	const float ax = (nx - o.x) * id.x, ay = (ny - o.y) * id.y, az = (nz - o.z) * id.z;
	const float bx = (fx - o.x) * id.x, by = (fy - o.y) * id.y, bz = (fz - o.z) * id.z;
	const float t0 = fmaxf(fmaxf(ax, ay), fmaxf(az, 0.0f));
	const float t1 = fminf(fminf(bx, by), fminf(bz, tbest));
	tnear = t0;
	return t0 <= t1 * 1.0000005f;
}

// FMA form of the same test: (plane - o) * id  ==  plane * id + (-(o * id)), one FFMA per plane instead of FADD + FMUL
// (24 instead of 48 FP instructions per 4-wide node; the box test is the busiest code of the traversal kernels).
// The product o * id is rounded once per ray, so every plane distance carries an ABSOLUTE error of up to
// |o * id| * 2^-24 on top of the FFMA's own rounding; `eps` (2^-22 * max over the axes of |o * id|, computed once per
// ray) widens the interval by more than both.  The test stays a conservative cull: it may keep a box the exact test
// would drop, never the other way round.  Rays for which eps is not small (a direction component that is zero or
// nearly zero: id = inf or huge, e.g. the centre column / centre row of every frame) use the exact form above; the
// choice is made per warp step (RT_SLAB_FMA_OK), so the branch is uniform.
__device__ __forceinline__ bool slab_hit_fma(float nx, float ny, float nz, float fx, float fy, float fz,
	const F3 &oi, const F3 &id, float tbest, float eps, float &tnear)
{
	const float ax = __fmaf_rn(nx, id.x, oi.x), ay = __fmaf_rn(ny, id.y, oi.y), az = __fmaf_rn(nz, id.z, oi.z);
	const float bx = __fmaf_rn(fx, id.x, oi.x), by = __fmaf_rn(fy, id.y, oi.y), bz = __fmaf_rn(fz, id.z, oi.z);
	const float t0 = fmaxf(fmaxf(ax, ay), fmaxf(az, 0.0f));
	const float t1 = fminf(fminf(bx, by), fminf(bz, tbest));
	tnear = t0;
	return t0 <= __fmaf_rn(t1, 1.0000005f, eps);
}
// Measured (profiles/r2_fma_slab_ab.txt): -5 % warp instructions in the wave kernels, +1 % rays/s, k_frame 4 % slower
// (four more live registers at the 64-register cap) -- the walk waits on node loads, not on issue slots.  Off by default.
#ifndef RT_SLAB_FMA
#define RT_SLAB_FMA 0
#endif
#define RT_SLAB_EPS_MAX 0.01f   // rays whose eps is not below this (or NaN) take the exact form

// Per-ray cache of the reference's part-level predicate (BorderTestEx), so the replay costs one
// evaluation per (ray, part) that produces a candidate.
struct PartCache
{
	uint32_t part;   // global part index, 0xFFFFFFFF = empty
	uint32_t mask;   // enabled octants, 0 if the part fails `minist < hr.distance`
};

__device__ __forceinline__ uint32_t part_mask(const SceneDev &S, const RayD &ray, const F3 &idir, uint32_t part, float hr_distance, PartCache &pc)
{
	if (pc.part != part)
	{
		const DevPart &P = S.parts[part];
		const float4 bmin = __ldg(&P.box_min), bmax = __ldg(&P.box_max);
		uint32_t m;
		const float minist = border_test_ex(ray.o, ray.d, idir, f3(bmin), f3(bmax), &m);
		pc.part = part;
		pc.mask = (minist < hr_distance) ? m : 0u;
	}
	return pc.mask;
}

// `first octant list that would test this triangle`, or -1 if the reference never tests it
__device__ __forceinline__ int tested_octant(uint32_t tri_octs, uint32_t mask, uint32_t tri, uint32_t skip)
{
	uint32_t s = tri_octs & mask;
	if (is_tri(skip) && (skip & 0x0FFFFFFFu) == tri)
		s &= ~(1u << ((skip >> 28) & 7u));
	return s ? (__ffs((int)s) - 1) : -1;
}

// reference iteration rank of a triangle hit inside one Model: (part, octant, index in part)
__device__ __forceinline__ bool tri_rank_less(const SceneDev &S, uint32_t partA, int octA, uint32_t triA, uint32_t idB)
{
	const uint32_t triB = idB & 0x0FFFFFFFu;
	const uint32_t partB = __ldg(&S.tri_part[triB]);
	const int octB = (int)((idB >> 28) & 7u);
	if (partA != partB) return partA < partB;
	if (octA != octB) return octA < octB;
	return triA < triB;   // same part: original index order == index-in-part order
}

// FAST (closest hit only): accept the nearest exact hit tentatively and postpone the replay of the
// reference's culling predicate to ONE check per ray after the walk (verify_fast_hit); anything
// the postponed check cannot decide -- an exact distance tie, a hit on the triangle the ray left
// from -- raises `slow`, and the caller redoes this Model with the immediate per-candidate replay.
// The result is identical either way: if the final nearest hit passes the replay, every candidate
// that was tentatively accepted before it was farther and irrelevant; if it fails, we fall back.
template<bool ANY, bool FAST, bool STATS>
__device__ __forceinline__ void leaf_tris(const SceneDev &S, const RayD &ray, const F3 &idir, uint32_t first, uint32_t count,
	float hr_distance, uint32_t model_tri_begin, uint32_t model_tri_end, PartCache &pc, Best &best, bool &done, bool &slow, TravStats &st)
{
	for (uint32_t k = 0; k < count; ++k)
	{
		float4 g0, g1, g2;
		load_tri(S.tri_geom, first + k, g0, g1, g2);
		if (STATS) ++st.tris;
		const float t = triangle_t(ray.o, ray.d, f3(g0), f3(g1), f3(g2), nullptr);
		if (ANY ? !(t < best.t) : !(t <= best.t && t < 1e20f))
			continue;
		const uint32_t tri = __float_as_uint(g0.w);
		const uint32_t pinfo = __float_as_uint(g1.w);
		if (FAST)
		{
			if (t == best.t || (is_tri(ray.skip) && (ray.skip & 0x0FFFFFFFu) == tri))
			{
				slow = true;
				continue;
			}
			best.t = t;
			best.id = RT_ID_TRI | tri;   // octant filled in by verify_fast_hit
			pc.part = pinfo;             // FAST reuses the cache slot to remember the winner's part | octants
			continue;
		}
		const uint32_t part = pinfo >> 8, octs = pinfo & 0xFFu;
		const uint32_t mask = part_mask(S, ray, idir, part, hr_distance, pc);
		const int oct = tested_octant(octs, mask, tri, ray.skip);
		if (oct < 0)
			continue;
		if (ANY)
		{
			best.t = t;
			done = true;
			return;
		}
		if (t == best.t)
		{
			// exact tie: an earlier object keeps the hit; inside this model the reference order decides
			const bool bestHere = is_tri(best.id) && (best.id & 0x0FFFFFFFu) >= model_tri_begin && (best.id & 0x0FFFFFFFu) < model_tri_end;
			if (!bestHere || !tri_rank_less(S, part, oct, tri, best.id))
				continue;
		}
		best.t = t;
		best.id = best.newobj = RT_ID_TRI | ((uint32_t)oct << 28) | tri;
	}
}

// one analytic primitive against the running best (closest) or the light distance (any-hit)
template<bool ANY>
__device__ __forceinline__ void test_prim(const SceneDev &S, const RayD &ray, uint32_t p, bool tieLower, Best &best, bool &done)
{
	const int4 meta = __ldg(&S.prim_meta[p]);
	const float4 g0 = ldg4(&S.prim_geom[4 * p]);
	const bool self = ray.skip == p;
	float t;
	if (meta.x == RT_OBJ_SPHERE)
	{
		const float4 g1 = ldg4(&S.prim_geom[4 * p + 1]);
		t = sphere_t(ray, f3(g0), g1.x, self);
	}
	else if (self)
		return;   // Box / Plane: `if (hr.obj == this) return hr`
	else if (meta.x == RT_OBJ_PLANE)
	{
		const float4 g1 = ldg4(&S.prim_geom[4 * p + 1]);
		t = plane_t(ray, f3(g0), f3(g1));
	}
	else
	{
		const float4 g1 = ldg4(&S.prim_geom[4 * p + 1]), g2 = ldg4(&S.prim_geom[4 * p + 2]);
		t = box_t(ray, f3(g1), f3(g2));
	}
	if (t < best.t || (!ANY && tieLower && t == best.t && t < 1e20f && p < best.id))
	{
		best.t = t;
		best.id = p;
		if (!self) best.newobj = p;
		if (ANY) done = true;
	}
}

// BVH walk shared by Models (triangle leaves) and runs of analytic primitives (prim leaves).
//
// Inner-node steps and leaf steps are separate warp-wide phases (the long, divergent leaf code never
// interleaves with box tests); which phase runs next is voted per step, see below.  The stack lives
// in local memory (L1-resident, one 128-byte line per depth and warp) so no shared memory is
// reserved and occupancy is bound by registers only.
#if RT_BVH8
// ---- 8-wide quantised node step (Model BVHs) ---------------------------------------------------------------------------
// Plane distance of quantised coordinate q along one axis: t = (p + q*s - o) * id.  The byte is planted into the mantissa of
// 1.0f with one PRMT (f = 1 + q * 2^-15), so t = f * A + B with A = 2^15 * s * id (exact scaling of id) and
// B = (p - o) * id - A, both per node and axis: one PRMT and one FFMA per plane, no integer-to-float conversion.  B carries a
// rounding error of about 2^-17 of the node's parametric length, far below the one-quantum padding of the boxes
// (rt_build.cu quantise_node8), so the test stays conservative without an epsilon of its own.
// `one` holds 0x3F800000 in a register the compiler cannot fold (otherwise ptxas makes the constant the immediate and
// re-materialises the byte selector into a register for every plane).
template<int K>
__device__ __forceinline__ float qplane(uint32_t word, uint32_t one, float A, float B)
{
	uint32_t f;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(word), "r"(one), "n"(0x7604 + (K << 4)));
	return __fmaf_rn(__uint_as_float(f), A, B);
}
#endif

// RT_NODE_FETCH == 2: the near / far plane vectors of an axis are fetched by PREDICATED loads at immediate offsets -- the sign
// of the ray direction picks which of two loads runs -- instead of loads at per-ray byte offsets, whose six 64-bit address
// sums (and the offsets themselves, re-made every step for want of registers) were 24 of the ~150 instructions of a node step.
template<int LO, int HI>
__device__ __forceinline__ void ldg_near_far(const char *n, uint32_t s, float4 &nr, float4 &fr)
{
	asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %8, 0;\n\t"
		"@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%9+%11];\n\t@!p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%9+%10];\n\t"
		"@p ld.global.nc.v4.f32 {%4,%5,%6,%7}, [%9+%10];\n\t@!p ld.global.nc.v4.f32 {%4,%5,%6,%7}, [%9+%11];\n\t}"
		: "=f"(nr.x), "=f"(nr.y), "=f"(nr.z), "=f"(nr.w), "=f"(fr.x), "=f"(fr.y), "=f"(fr.z), "=f"(fr.w)
		: "r"(s), "l"(n), "n"(LO), "n"(HI));
}

template<bool ANY> struct StackSlot { typedef uint2 T; };
template<> struct StackSlot<true> { typedef int T; };
__device__ __forceinline__ void slot_put(int &s, int link, float) { s = link; }
__device__ __forceinline__ void slot_put(uint2 &s, int link, float t) { s = make_uint2((uint32_t)link, __float_as_uint(t)); }
__device__ __forceinline__ int slot_link(const int &s) { return s; }
__device__ __forceinline__ int slot_link(const uint2 &s) { return (int)s.x; }
__device__ __forceinline__ float slot_t(const int &) { return 0.0f; }
__device__ __forceinline__ float slot_t(const uint2 &s) { return __uint_as_float(s.y); }
#define RT_TRAV_DONE (-1)   // never a valid leaf code: that would be first = 2^28-1, count = 8

#ifndef RT_EARLY_HOLD
#define RT_EARLY_HOLD 4u   // EARLY: steps a finished lane waits for others to finish before the group leaves
#endif

template<bool ANY, bool TRIS, bool FAST, bool STATS, bool EARLY = false>
__device__ __forceinline__ void traverse(const SceneDev &S, const RayD &ray, const F3 &idir, int root,
	float hr_distance, uint32_t rangeBegin, uint32_t rangeEnd, Best &best, bool &done, TravStats &st,
	uint32_t winLo = 0u, uint32_t winHi = 0xFFFFFFFFu)
{
	PartCache pc;
	pc.part = 0xFFFFFFFFu, pc.mask = 0;
	bool slow = false;
	const uint32_t idBefore = best.id;
	// closest hit: link + entry distance of each stacked subtree in ONE 64-bit slot (one STL.64 per push, one LDL.64
	// per pop; the subtree is culled again on pop); any-hit: links only
	typename StackSlot<ANY>::T stack[RT_STACK];
	int sp = 0;
	int cur = root;
	// byte offsets of the near / far plane vectors inside a BvhNode4 for this ray's direction signs
	const uint32_t sx = __float_as_uint(ray.d.x) >> 31, sy = __float_as_uint(ray.d.y) >> 31, sz = __float_as_uint(ray.d.z) >> 31;
#if RT_NODE_FETCH == 0
	const uint32_t onx = sx ? 48u : 0u, ony = sy ? 64u : 16u, onz = sz ? 80u : 32u;
	const uint32_t ofx = sx ? 0u : 48u, ofy = sy ? 16u : 64u, ofz = sz ? 32u : 80u;
#endif
	// Phase voting.  The lanes that entered together decide every step, by majority, whether the warp
	// runs ONE inner-node step or ONE leaf step; a lane in the minority keeps its node / leaf for a
	// later step.  The classic while-while shape (every lane descends until it holds a leaf, then all
	// leaves are tested) makes each round as long as the slowest descent of 32 lanes: measured on C3,
	// 7 of 32 lanes were active in the node code although the lanes' total lengths alone allow 19.
	//
	// EARLY (whole-frame scheduler, last scene item of a closest-hit ray): lanes that are through do not
	// wait for the longest ray of the batch; they leave together once one of them has idled for
	// RT_EARLY_HOLD steps, and the caller publishes their results while the rest keeps walking.
	uint32_t wmask = __activemask();
	uint32_t waited = 0;
#if RT_BVH8
	// The quantised step works with a reciprocal direction clamped to +-1e12 (+-inf for a zero component included): t = f * A + B
	// must not meet inf - inf.  The clamp keeps the test conservative -- along such an axis the ray moves less than 1e-9 over
	// the whole scene, so "origin inside the padded slab" decides, and (plane - o) * 1e12 still lies far beyond every distance
	// the other axes allow unless the origin sits exactly on the padded plane, one quantum outside the real box.
	const uint32_t one = 0x3F800000u + (S.n_items >> 31);   // = 0x3F800000: see qplane
	const F3 qid = f3(copysignf(fminf(fabsf(idir.x), 1e12f), idir.x), copysignf(fminf(fabsf(idir.y), 1e12f), idir.y), copysignf(fminf(fabsf(idir.z), 1e12f), idir.z));
#endif
#if RT_SLAB_FMA
	const F3 oi = f3(-(ray.o.x * idir.x), -(ray.o.y * idir.y), -(ray.o.z * idir.z));
	float eps = fmaxf(fmaxf(fabsf(oi.x), fabsf(oi.y)), fabsf(oi.z)) * 2.4e-7f;
	// every axis is checked on its own: fmaxf drops a NaN (0 * inf: an origin coordinate of exactly 0 with a direction
	// component of exactly 0 -- the centre column of a frame whose camera sits at x = 0), and such an axis would then
	// constrain nothing in the FMA form
	const float lim = RT_SLAB_EPS_MAX / 2.4e-7f;
	const bool fmaOk = fabsf(oi.x) < lim && fabsf(oi.y) < lim && fabsf(oi.z) < lim;
	if (!fmaOk) eps = 0.0f;   // such a lane only ever runs the exact form (the lanes that entered together all take it)
	const bool useFma = __ballot_sync(wmask, !fmaOk) == 0u;
#else
	const float eps = 0.0f;
#endif
	while (true)
	{
		const bool atNode = cur >= 0, atLeaf = cur < 0 && cur != RT_TRAV_DONE;
		const uint32_t mN = __ballot_sync(wmask, atNode), mL = __ballot_sync(wmask, atLeaf);
		if ((mN | mL) == 0u)
			break;
		if (EARLY && (mN | mL) != wmask)
		{
			const bool idle = !(atNode || atLeaf);
			if (idle) ++waited;
			if (__ballot_sync(wmask, idle && waited >= RT_EARLY_HOLD))
			{
				wmask = mN | mL;
				if (idle) break;
			}
		}
		if (__popc(mN) >= __popc(mL))
		{
#if RT_BVH8
			if (atNode && TRIS)
			{
				const uint4 *nd = (const uint4 *)&S.nodes8[cur];
				const uint4 hd = __ldg(nd), q0 = __ldg(nd + 1), q1 = __ldg(nd + 2), q2 = __ldg(nd + 3);
				const uint4 la = __ldg(nd + 4), lb = __ldg(nd + 5);
				if (STATS) ++st.nodes;
				// A = 2^15 * s * id, B = (p - o) * id - A per axis
				const float Ax = __uint_as_float(((hd.w & 0xFFu) + 15u) << 23) * qid.x;
				const float Ay = __uint_as_float((((hd.w >> 8) & 0xFFu) + 15u) << 23) * qid.y;
				const float Az = __uint_as_float((((hd.w >> 16) & 0xFFu) + 15u) << 23) * qid.z;
				const float Bx = (__uint_as_float(hd.x) - ray.o.x) * qid.x - Ax;
				const float By = (__uint_as_float(hd.y) - ray.o.y) * qid.y - Ay;
				const float Bz = (__uint_as_float(hd.z) - ray.o.z) * qid.z - Az;
				// near / far plane words by the ray's direction signs (qlo* = q0.xy, q0.zw, q1.xy; qhi* = q1.zw, q2.xy, q2.zw)
				const uint32_t nx0 = sx ? q1.z : q0.x, nx1 = sx ? q1.w : q0.y, fx0 = sx ? q0.x : q1.z, fx1 = sx ? q0.y : q1.w;
				const uint32_t ny0 = sy ? q2.x : q0.z, ny1 = sy ? q2.y : q0.w, fy0 = sy ? q0.z : q2.x, fy1 = sy ? q0.w : q2.y;
				const uint32_t nz0 = sz ? q2.z : q1.x, nz1 = sz ? q2.w : q1.y, fz0 = sz ? q1.x : q2.z, fz1 = sz ? q1.y : q2.w;
				// nearest hit child is walked next, the other hit children go on the stack as they are found (no arrays: the
				// running nearest lives in two registers and is swapped out to the stack by a nearer sibling)
				int bl = RT_TRAV_DONE;
				float bt = __int_as_float(0x7f800000);
#define RT_CHILD8(k, NX, NY, NZ, FX, FY, FZ, LINK) \
				{ \
					const float t0 = fmaxf(fmaxf(qplane<(k) & 3>(NX, one, Ax, Bx), qplane<(k) & 3>(NY, one, Ay, By)), fmaxf(qplane<(k) & 3>(NZ, one, Az, Bz), 0.0f)); \
					const float t1 = fminf(fminf(qplane<(k) & 3>(FX, one, Ax, Bx), qplane<(k) & 3>(FY, one, Ay, By)), fminf(qplane<(k) & 3>(FZ, one, Az, Bz), best.t)); \
					const bool hit = t0 <= t1;   /* an unused child is an inverted box: near plane behind far plane on every axis */ \
					const bool nearer = hit && t0 < bt; \
					const int pl = nearer ? bl : (int)(LINK); \
					const float pt = nearer ? bt : t0; \
					bl = nearer ? (int)(LINK) : bl, bt = nearer ? t0 : bt; \
					if (hit && pl != RT_TRAV_DONE) { slot_put(stack[sp], pl, pt); ++sp; } \
				}
				RT_CHILD8(0, nx0, ny0, nz0, fx0, fy0, fz0, la.x)
				RT_CHILD8(1, nx0, ny0, nz0, fx0, fy0, fz0, la.y)
				RT_CHILD8(2, nx0, ny0, nz0, fx0, fy0, fz0, la.z)
				RT_CHILD8(3, nx0, ny0, nz0, fx0, fy0, fz0, la.w)
				RT_CHILD8(4, nx1, ny1, nz1, fx1, fy1, fz1, lb.x)
				RT_CHILD8(5, nx1, ny1, nz1, fx1, fy1, fz1, lb.y)
				RT_CHILD8(6, nx1, ny1, nz1, fx1, fy1, fz1, lb.z)
				RT_CHILD8(7, nx1, ny1, nz1, fx1, fy1, fz1, lb.w)
#undef RT_CHILD8
				cur = bl;
				if (bl == RT_TRAV_DONE)
					while (sp)
					{
						const typename StackSlot<ANY>::T e = stack[--sp];
						if (ANY || slot_t(e) <= best.t) { cur = slot_link(e); break; }
					}
			}
			else
#endif
			if (atNode)
			{
				const char *n = (const char *)&S.nodes4[cur];
#if RT_NODE_FETCH == 1
				float4 lx, hx, ly, hy, lz, hz;
				ldg8(n, lx, hx), ldg8(n + 32, ly, hy), ldg8(n + 64, lz, hz);
				const float4 nx = sx ? hx : lx, ny = sy ? hy : ly, nz = sz ? hz : lz;
				const float4 fx = sx ? lx : hx, fy = sy ? ly : hy, fz = sz ? lz : hz;
#elif RT_NODE_FETCH == 2
				float4 nx, ny, nz, fx, fy, fz;
				ldg_near_far<0, 48>(n, sx, nx, fx), ldg_near_far<16, 64>(n, sy, ny, fy), ldg_near_far<32, 80>(n, sz, nz, fz);
#else
				const float4 nx = ldg4((const float4 *)(n + onx)), ny = ldg4((const float4 *)(n + ony)), nz = ldg4((const float4 *)(n + onz));
				const float4 fx = ldg4((const float4 *)(n + ofx)), fy = ldg4((const float4 *)(n + ofy)), fz = ldg4((const float4 *)(n + ofz));
#endif
				const int4 link = __ldg((const int4 *)(n + 96));
				if (STATS) ++st.nodes;
				float t0, t1, t2, t3;
				bool h0, h1, h2, h3;
#if RT_SLAB_FMA
				if (useFma)
				{
					h0 = slab_hit_fma(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, oi, idir, best.t, eps, t0);
					h1 = slab_hit_fma(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, oi, idir, best.t, eps, t1);
					h2 = slab_hit_fma(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, oi, idir, best.t, eps, t2);
					h3 = slab_hit_fma(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, oi, idir, best.t, eps, t3);
				}
				else
#endif
				{
					h0 = slab_hit_nf(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, ray.o, idir, best.t, t0);
					h1 = slab_hit_nf(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, ray.o, idir, best.t, t1);
					h2 = slab_hit_nf(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, ray.o, idir, best.t, t2);
					h3 = slab_hit_nf(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, ray.o, idir, best.t, t3);
				}
				// nearest hit child first, the other hit children go on the stack
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
						if (ANY || slot_t(e) <= best.t + eps) { cur = slot_link(e); break; }
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
			if (TRIS)
				leaf_tris<ANY, FAST, STATS>(S, ray, idir, first, count, hr_distance, rangeBegin, rangeEnd, pc, best, done, slow, st);
			else
				for (uint32_t k = 0; k < count; ++k)
				{
					const uint32_t p = __ldg(&S.bvh_prims[first + k]);
					if (p < winLo || p >= winHi)
						continue;
					if (STATS) ++st.prims;
					test_prim<ANY>(S, ray, p, !(best.id & RT_ID_TRI) && best.id >= rangeBegin, best, done);
				}
			cur = RT_TRAV_DONE;
			if (ANY && done)
				sp = 0;   // occluded: nothing else to look at, but stay in the vote until the other lanes are through
			while (sp)
			{
				const typename StackSlot<ANY>::T e = stack[--sp];
				if (ANY || slot_t(e) <= best.t + eps) { cur = slot_link(e); break; }
			}
		}
	}
	if (ANY && done)
		return;
	if (FAST)
	{
		// one replay per ray, executed by the whole warp together
		if (!slow && best.id != idBefore)
		{
			const uint32_t pinfo = pc.part, tri = best.id & 0x0FFFFFFFu;
			PartCache one;
			one.part = 0xFFFFFFFFu, one.mask = 0;
			const uint32_t mask = part_mask(S, ray, idir, pinfo >> 8, hr_distance, one);
			const int oct = tested_octant(pinfo & 0xFFu, mask, tri, ray.skip);
			if (oct < 0) slow = true;
			else best.id = best.newobj = RT_ID_TRI | ((uint32_t)oct << 28) | tri;
		}
		if (slow) best.t = -1.0f;   // tells the caller to redo the item with the immediate replay
	}
}

// The scene walk in Objects order.  Closest hit: best starts at (1e20, NONE).  Any-hit: best.t
// starts at the light distance and `done` reports occlusion.
template<bool ANY, bool STATS, bool EARLY = false>
__device__ __forceinline__ void trace_scene(const SceneDev &S, const RayD &ray, Best &best, bool &done, TravStats &st)
{
	const F3 idir = f3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
	for (uint32_t i = 0; i < S.n_items; ++i)
	{
		const SceneItem it = S.items[i];
		if (it.kind == RT_ITEM_PRIM)
		{
			if (STATS) ++st.prims;
			test_prim<ANY>(S, ray, it.first, false, best, done);
		}
		else if (S.brute && it.kind == RT_ITEM_PRIMBVH)
		{
			// diagnostic: the run in object order, no BVH
			for (uint32_t p = it.first; p < it.first + it.count; ++p)
			{
				if (STATS) ++st.prims;
				test_prim<ANY>(S, ray, p, false, best, done);
				if (ANY && done)
					return;
			}
		}
		else if (it.kind == RT_ITEM_PRIMBVH)
		{
			const uint32_t end = it.first + it.count;
			if (!ANY && ray.isInside && !(ray.skip & RT_ID_TRI) && ray.skip >= it.first && ray.skip < end)
			{
				// the ray starts inside a sphere of this run: `newobj` depends on which hits were
				// accepted before / after that sphere in object order, so walk the run in three
				// order-respecting phases (primitives before it, the sphere itself, primitives after)
				traverse<ANY, false, false, STATS>(S, ray, idir, it.root, best.t, it.first, end, best, done, st, it.first, ray.skip);
				if (STATS) ++st.prims;
				test_prim<ANY>(S, ray, ray.skip, false, best, done);
				traverse<ANY, false, false, STATS>(S, ray, idir, it.root, best.t, it.first, end, best, done, st, ray.skip + 1u, end);
			}
			else if (EARLY && i + 1u == S.n_items)
				traverse<ANY, false, false, STATS, EARLY>(S, ray, idir, it.root, best.t, it.first, end, best, done, st);
			else
				traverse<ANY, false, false, STATS>(S, ray, idir, it.root, best.t, it.first, end, best, done, st);
		}
		else
		{
			const DevModel &M = S.models[it.first];
			const float4 mn = __ldg(&M.border_min), mx = __ldg(&M.border_max);
			// Model.cpp:752: `if (BorderTest(ray, BorderMin, BorderMax) < hr.distance)`
			if (!(border_test(ray.o, ray.d, idir, f3(mn), f3(mx)) < best.t))
				continue;
			const uint32_t tb = __ldg(&M.tri_begin);
			if (S.brute)
			{
				// diagnostic: every triangle of the model with the immediate culling replay, no BVH
				const uint32_t te = tb + __ldg(&M.tri_count);
				PartCache pc;
				pc.part = 0xFFFFFFFFu, pc.mask = 0;
				bool slow = false;
				leaf_tris<ANY, false, STATS>(S, ray, idir, tb, te - tb, best.t, tb, te, pc, best, done, slow, st);
				if (ANY && done)
					return;
				continue;
			}
			if (ANY)
				traverse<true, true, false, STATS>(S, ray, idir, it.root, best.t, tb, tb + __ldg(&M.tri_count), best, done, st);
			else
			{
				const Best before = best;
				if (EARLY && i + 1u == S.n_items)
					traverse<false, true, true, STATS, EARLY>(S, ray, idir, it.root, before.t, tb, tb + __ldg(&M.tri_count), best, done, st);
				else
					traverse<false, true, true, STATS>(S, ray, idir, it.root, before.t, tb, tb + __ldg(&M.tri_count), best, done, st);
				if (best.t < 0.0f)
				{
					best = before;
					traverse<false, true, false, STATS>(S, ray, idir, it.root, before.t, tb, tb + __ldg(&M.tri_count), best, done, st);
				}
			}
		}
		if (ANY && done)
			return;
	}
}

// ---- the scene walk in two stages ------------------------------------------------------------------------------------
// Measured on C3 (RT_FLAG_STATS histogram): 60 % of the closest-hit rays and 45 % of the shadow rays never enter the mesh's
// BVH -- they miss the Model's box (sky, rays leaving the height field) -- while the others visit 4..15 nodes.  In a batch of
// 32 rays the lanes of the former idle through the whole walk of the latter.  The wave kernels therefore run the walk in
// two stages: `head` = everything up to and including the box test of the LAST scene item when that item is a Model
// (RayTracer.cpp:458-465 visits the objects in order; Model.cpp:752 is the box test), `tail` = the walk of that Model's
// BVH.  Rays that pass the box test are collected per warp until 32 of them are there, so the tail always starts with full
// lanes.  The per-ray arithmetic and the order of the objects are those of trace_scene.
__device__ __forceinline__ bool scene_has_tail(const SceneDev &S)
{
	return S.n_items > 0u && !S.brute && S.items[S.n_items - 1u].kind == RT_ITEM_MODEL && S.items[S.n_items - 1u].count > 0u;
}

// -> true: the ray passed the last Model's box test and its BVH has to be walked (trace_scene_tail); false: the ray is
// through (any-hit: `done` says occluded)
template<bool ANY, bool STATS>
__device__ __forceinline__ bool trace_scene_head(const SceneDev &S, const RayD &ray, Best &best, bool &done, TravStats &st, bool hasTail)
{
	const F3 idir = f3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
	const uint32_t nHead = hasTail ? S.n_items - 1u : S.n_items;
	for (uint32_t i = 0; i < nHead; ++i)
	{
		const SceneItem it = S.items[i];
		if (it.kind == RT_ITEM_PRIM)
		{
			if (STATS) ++st.prims;
			test_prim<ANY>(S, ray, it.first, false, best, done);
		}
		else if (S.brute && it.kind == RT_ITEM_PRIMBVH)
		{
			for (uint32_t p = it.first; p < it.first + it.count; ++p)
			{
				if (STATS) ++st.prims;
				test_prim<ANY>(S, ray, p, false, best, done);
				if (ANY && done)
					return false;
			}
		}
		else if (it.kind == RT_ITEM_PRIMBVH)
		{
			const uint32_t end = it.first + it.count;
			if (!ANY && ray.isInside && !(ray.skip & RT_ID_TRI) && ray.skip >= it.first && ray.skip < end)
			{
				traverse<ANY, false, false, STATS>(S, ray, idir, it.root, best.t, it.first, end, best, done, st, it.first, ray.skip);
				if (STATS) ++st.prims;
				test_prim<ANY>(S, ray, ray.skip, false, best, done);
				traverse<ANY, false, false, STATS>(S, ray, idir, it.root, best.t, it.first, end, best, done, st, ray.skip + 1u, end);
			}
			else
				traverse<ANY, false, false, STATS>(S, ray, idir, it.root, best.t, it.first, end, best, done, st);
		}
		else
		{
			const DevModel &M = S.models[it.first];
			const float4 mn = __ldg(&M.border_min), mx = __ldg(&M.border_max);
			if (!(border_test(ray.o, ray.d, idir, f3(mn), f3(mx)) < best.t))
				continue;
			const uint32_t tb = __ldg(&M.tri_begin);
			if (S.brute)
			{
				const uint32_t te = tb + __ldg(&M.tri_count);
				PartCache pc;
				pc.part = 0xFFFFFFFFu, pc.mask = 0;
				bool slow = false;
				leaf_tris<ANY, false, STATS>(S, ray, idir, tb, te - tb, best.t, tb, te, pc, best, done, slow, st);
			}
			else if (ANY)
				traverse<true, true, false, STATS>(S, ray, idir, it.root, best.t, tb, tb + __ldg(&M.tri_count), best, done, st);
			else
			{
				const Best before = best;
				traverse<false, true, true, STATS>(S, ray, idir, it.root, before.t, tb, tb + __ldg(&M.tri_count), best, done, st);
				if (best.t < 0.0f)
				{
					best = before;
					traverse<false, true, false, STATS>(S, ray, idir, it.root, before.t, tb, tb + __ldg(&M.tri_count), best, done, st);
				}
			}
		}
		if (ANY && done)
			return false;
	}
	if (!hasTail)
		return false;
	// Model.cpp:752: `if (BorderTest(ray, BorderMin, BorderMax) < hr.distance)`
	const DevModel &M = S.models[S.items[S.n_items - 1u].first];
	const float4 mn = __ldg(&M.border_min), mx = __ldg(&M.border_max);
	return border_test(ray.o, ray.d, idir, f3(mn), f3(mx)) < best.t;
}

template<bool ANY, bool STATS>
__device__ __forceinline__ void trace_scene_tail(const SceneDev &S, const RayD &ray, Best &best, bool &done, TravStats &st)
{
	const F3 idir = f3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
	const SceneItem it = S.items[S.n_items - 1u];
	const DevModel &M = S.models[it.first];
	const uint32_t tb = __ldg(&M.tri_begin), te = tb + __ldg(&M.tri_count);
	if (ANY)
		traverse<true, true, false, STATS>(S, ray, idir, it.root, best.t, tb, te, best, done, st);
	else
	{
		const Best before = best;
		traverse<false, true, true, STATS>(S, ray, idir, it.root, before.t, tb, te, best, done, st);
		if (best.t < 0.0f)
		{
			best = before;
			traverse<false, true, false, STATS>(S, ray, idir, it.root, before.t, tb, te, best, done, st);
		}
	}
}

