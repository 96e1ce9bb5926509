// Wavefront kernels of the trace-and-shade path (sm_100a, compiled with -fmad=false):
//
//   k_raygen        RayTracer::parallelRT pixel loop          /root/reference/RayTracer.cpp:20-27
//   k_wave(l)       closest hit of ray level l  +  surface attributes  +  secondary-ray queue (level l+1)
//                   fused with the shadow any-hit queries of level l-1
//                                                             RayTracer.cpp:455-468, 548-581 and :482-520
//   k_shade         Blinn-Phong of every surface of every level (one launch)   RayTracer.cpp:470-546
//   k_combine       post-order colour combine + Color::put    RayTracer.cpp:552-594, 3DElement.cpp:463-468
//
// Why trace(l) and shadow(l-1) share a launch: a few rays skim the height field and visit
// thousands of BVH nodes; with one kernel per stage every launch ended in a ~250 us tail of such
// rays with the rest of the chip idle.  Spawning level l+1 inside the trace epilogue removes the
// trace(l+1) -> shade(l) -> shadow(l) dependency, so the long rays of one stage overlap the bulk
// of the next.
//
// The reference recursion (binary ray tree per pixel) is unrolled level by level: level l holds
// every ray whose recursion depth is l; a ray's slot index is also its ray-tree node index, and
// each node remembers the slots of its two children in level l+1.  Shading stores the node-local
// colour; k_combine then walks the levels bottom-up and applies exactly the reference's
// c = c*(1-kr) + c_refl*kr ; c = c*(1-kt) + (c_refr (*) beer)*kt sequence, so the colour
// arithmetic is bit-identical to the recursive evaluation order.
#include "rt_traverse.cuh"
#include "rt_defer.cuh"
#include "rt_kernels.h"
#include <cstdlib>
#include <algorithm>

// ---- helpers -------------------------------------------------------------------------------------

// level-0 slot -> pixel: slots are laid out in 8x8 pixel tiles so a warp covers an 8x4 block
__device__ __forceinline__ void slot_to_pixel(const FrameParams &F, uint32_t i, int &x, int &y)
{
	const uint32_t w64 = (uint32_t)F.blk_w * 64u;
	const uint32_t tile = i >> 6, in = i & 63u;
	const uint32_t tiles_x = w64 >> 3;
	const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
	x = (int)(tx * 8u + (in & 7u));
	const uint32_t row = ty * 8u + (in >> 3);            // row among this shard's rows
	const uint32_t band = row / F.tile_rows + F.tile_first;   // row tile among this shard's tiles
	y = (int)(shard_tile(band, F.rank, F.world, F.serpentine) * F.tile_rows + row % F.tile_rows);
}

__device__ __forceinline__ uint8_t put8(float c)
{
	// Color::put: clamp with strict compares, truncate; NaN -> 0 (x86 cvttss2si low byte)
	if (c > 1.0f) return 255;
	if (c < 0.0f) return 0;
	const float v = c * 255;
	if (!(v == v)) return 0;
	return (uint8_t)v;
}

// Color(const Texture*, Coord2D), 3DElement.cpp:430-450
__device__ __forceinline__ F3 texel(const SceneDev &S, int tex, float cu, float cv)
{
	if (tex < 0)
		return f3(1.0f, 1.0f, 1.0f);
	const int4 T = __ldg(&S.textures[tex]);
	float whole;
	float nu = modff(cu, &whole), nv = modff(cv, &whole);
	if (nu < 0) nu += 1;
	if (nv < 0) nv += 1;
	const int x = (int)(short)(int)(nu * (float)T.x), y = (int)(short)(int)(nv * (float)T.y);
	const uint8_t *px = S.texels + (uint32_t)T.z + (y * T.x + x) * 3;
	return f3(px[2] / 255.0f, px[1] / 255.0f, px[0] / 255.0f);
}

// pow(float,float) / exp(float) of the host libm are within a fraction of an ulp of correctly
// rounded; FP64 evaluation rounded once reproduces them (DESIGN.md "transcendentals").
__device__ __forceinline__ float pow_ref(float a, float b) { return (float)pow((double)a, (double)b); }
__device__ __forceinline__ float exp_ref(float a) { return (float)exp((double)a); }

__device__ __forceinline__ RayD load_ray(const LevelBuf &L, uint32_t i)
{
	const float4 o = L.ray_o[i], d = L.ray_d[i];
	const uint2 m = L.ray_meta[i];
	RayD r;
	r.o = f3(o), r.d = f3(d), r.mtlrfr = o.w;
	r.skip = m.x, r.type = (uint8_t)(m.y & 0xFF), r.isInside = (uint8_t)((m.y >> 8) & 0xFF);
	return r;
}

template<bool STATS>
__device__ __forceinline__ void flush_stats(WaveState *ws, const TravStats &st)
{
	if (!STATS) return;
	unsigned long long n = st.nodes, t = st.tris, p = st.prims;
	for (int o = 16; o > 0; o >>= 1)
	{
		n += __shfl_down_sync(0xffffffffu, n, o);
		t += __shfl_down_sync(0xffffffffu, t, o);
		p += __shfl_down_sync(0xffffffffu, p, o);
	}
	if ((threadIdx.x & 31) == 0)
	{
		atomicAdd(&ws->nodes_visited, n);
		atomicAdd(&ws->tri_tests, t);
		atomicAdd(&ws->prim_tests, p);
	}
}

// RT_FLAG_STATS: SIMT utilisation a batch could reach at best = sum of the lanes' node visits / (32 x the longest lane)
template<bool STATS>
__device__ __forceinline__ void lane_stats(WaveState *ws, uint32_t kind, uint32_t mine)
{
	if (!STATS) return;
	uint32_t s = mine, m = mine;
	for (int o = 16; o > 0; o >>= 1)
	{
		s += __shfl_xor_sync(0xffffffffu, s, o);
		m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	}
	if ((threadIdx.x & 31) == 0)
	{
		atomicAdd(&ws->lane_sum[kind], (unsigned long long)s);
		atomicAdd(&ws->lane_cap[kind], 32ull * m);
	}
}

// ---- ray generation ------------------------------------------------------------------------------

// RayTracer::parallelRT, RayTracer.cpp:20-27: dir = cam.n + cam.u*(xcur*dp) + cam.v*(ycur*dp), then Ray() normalises
__device__ __forceinline__ F3 primary_dir(const FrameParams &F, uint32_t i, F3 &origin)
{
	const BatchFrame &B = F.frames[frame_of(F, i)];
	int x, y;
	slot_to_pixel(F, i, x, y);
	const int xcur = x - F.half_w, ycur = y - F.half_h;
	const float sx = (float)(xcur * F.dp), sy = (float)(ycur * F.dp);
	origin = f3(B.cam_pos);
	return normalize((f3(B.cam_n) + f3(B.cam_u) * sx) + f3(B.cam_v) * sy);
}

// level-0 slot -> address of its pixel in the framebuffer of its frame
__device__ __forceinline__ uint8_t *pixel_of(const FrameParams &F, uint32_t i)
{
	const BatchFrame &B = F.frames[frame_of(F, i)];
	int x, y;
	slot_to_pixel(F, i, x, y);
	return B.out + ((size_t)y * F.width + x) * 3;
}

__global__ void __launch_bounds__(256) k_raygen(const FrameParams *__restrict__ Fp, LevelBuf L, uint32_t n)
{
	const FrameParams &F = *Fp;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		F3 o;
		const F3 d = primary_dir(F, i, o);
		L.ray_o[i] = make_float4(o.x, o.y, o.z, 1.0f);
		L.ray_d[i] = make_float4(d.x, d.y, d.z, 1.0f);
		L.ray_meta[i] = make_uint2(RT_ID_NONE, (uint32_t)MY_RAY_BASERAY_ | (F.epoch << 16));
	}
}

// ---- queue helpers ---------------------------------------------------------------------------------

// Persistent warps: each warp pulls 32 consecutive work items at a time (one atomic per fetch), so
// SMs stay busy until the queue is dry and queue order (8x4 pixel tiles for primary rays, parent
// order for secondary rays, same light for shadow rays) keeps the lanes of a warp coherent.
__device__ __forceinline__ uint32_t warp_fetch(uint32_t *head, uint32_t batch)
{
	uint32_t base = 0;
	if ((threadIdx.x & 31) == 0) base = atomicAdd(head, batch);
	return __shfl_sync(0xffffffffu, base, 0);
}

// Rays per fetch.  A warp executes the union of its lanes' divergent paths, so when a queue holds
// fewer rays than the resident warps could take 32 at a time (deep recursion levels), handing each
// warp only a few rays spreads the same work over more schedulers and shortens the critical path.
__device__ __forceinline__ uint32_t fetch_batch(uint32_t n)
{
	const uint32_t warps = gridDim.x * (RT_BLOCK / 32);
	uint32_t b = (n + warps - 1) / warps;
	b = b < 2u ? 2u : b;
	return b > 32u ? 32u : b;
}

__device__ __forceinline__ uint32_t warp_append(uint32_t *counter, bool want)
{
	// warp-aggregated queue append: one atomic per warp, slots in lane order (keeps rays coherent)
	const uint32_t m = __ballot_sync(0xffffffffu, want);
	if (m == 0) return 0xFFFFFFFFu;
	const uint32_t lane = threadIdx.x & 31u, leader = __ffs((int)m) - 1;
	uint32_t base = 0;
	if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
	base = __shfl_sync(0xffffffffu, base, leader);
	return want ? base + __popc(m & ((1u << lane) - 1u)) : 0xFFFFFFFFu;
}

// the same for a converged subset `grp` of the warp (lanes that left the traversal together)
__device__ __forceinline__ uint32_t group_append(uint32_t *counter, bool want, uint32_t grp)
{
	const uint32_t m = __ballot_sync(grp, want);
	if (m == 0) return 0xFFFFFFFFu;
	const uint32_t lane = threadIdx.x & 31u, leader = __ffs((int)m) - 1;
	uint32_t base = 0;
	if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
	base = __shfl_sync(grp, base, leader);
	return want ? base + __popc(m & ((1u << lane) - 1u)) : 0xFFFFFFFFu;
}

// light k as seen from P: direction p2l and occlusion range (RayTracer.cpp:482-503)
__device__ __forceinline__ void light_dir(const DevLight &lit, const F3 &P, F3 &p2l, float &dis, float &lum)
{
	if (lit.type == RT_LIGHT_POINT)
	{
		const F3 v = f3(lit.position) - P;
		dis = dot(v, v);
		float step = lit.attenuation.x + lit.attenuation.z * dis;
		dis = sqrtf(dis);
		step += lit.attenuation.y * dis;
		lum = 1 / step;
		p2l = normalize(v);
	}
	else
	{
		dis = 1e10f;
		lum = 1.0f;   // parallel and spot lights are used unattenuated
		p2l = normalize(f3(lit.position));
	}
}

// ---- surface attributes: what the intersect operators leave in HitRes ------------------------------

struct Surface
{
	F3 N;
	float tu, tv, rfr;
	int mtl, tex;
	uint8_t isInside;
};

__device__ __forceinline__ Surface surface_attributes(const SceneDev &S, const RayD &ray, const F3 &P, uint32_t id)
{
	Surface s;
	s.tu = s.tv = 0.0f, s.rfr = 1.0f, s.tex = -1, s.isInside = 0;
	if (is_tri(id))
	{
		// Model::intersect epilogue, Model.cpp:794-807
		const uint32_t tri = id & 0x0FFFFFFFu, slot = __ldg(&S.tri_slot[tri]);
		float4 g0, g1, g2;
		load_tri(S.tri_geom, slot, g0, g1, g2);
		F3 bary = f3(0, 0, 0);
		triangle_t(ray.o, ray.d, f3(g0), f3(g1), f3(g2), &bary);
		const F3 n0 = f3(ldg4(&S.tri_norms[3 * tri])), n1 = f3(ldg4(&S.tri_norms[3 * tri + 1])), n2 = f3(ldg4(&S.tri_norms[3 * tri + 2]));
		s.N = normalize((n0 * bary.x + n1 * bary.y) + n2 * bary.z);
		const float2 c0 = __ldg(&S.tri_tcoords[3 * tri]), c1 = __ldg(&S.tri_tcoords[3 * tri + 1]), c2 = __ldg(&S.tri_tcoords[3 * tri + 2]);
		s.tu = (c0.x * bary.x + c1.x * bary.y) + c2.x * bary.z;
		s.tv = (c0.y * bary.x + c1.y * bary.y) + c2.y * bary.z;
		const DevPart &part = S.parts[__ldg(&S.tri_part[tri])];
		s.mtl = (int)__ldg(&part.material);
		s.tex = __ldg(&part.texture);
		s.rfr = __ldg(&S.materials[4 * s.mtl + 3]).w;
		return s;
	}
	const int4 meta = __ldg(&S.prim_meta[id]);
	const float4 g0 = ldg4(&S.prim_geom[4 * id]);
	s.mtl = meta.y;
	if (meta.x == RT_OBJ_SPHERE)
	{
		// Basic3DObject.cpp:146-160 (leaving through the far wall) and :179-187
		const bool leaving = ray.skip == id;   // only reachable with ray.isInside
		s.N = leaving ? normalize(f3(g0) - P) : normalize(P - f3(g0));
		s.isInside = (uint8_t)~ray.isInside;
		s.rfr = (leaving && ray.type == MY_RAY_REFRACTRAY_) ? 1.0f : __ldg(&S.materials[4 * s.mtl + 3]).w;
	}
	else if (meta.x == RT_OBJ_PLANE)
	{
		s.N = f3(ldg4(&S.prim_geom[4 * id + 1]));
		const float2 tc = plane_tcoord(ray.o, ray.d, f3(g0), f3(ldg4(&S.prim_geom[4 * id + 2])), f3(ldg4(&S.prim_geom[4 * id + 3])));
		s.tu = tc.x, s.tv = tc.y;
		s.tex = meta.z;
	}
	else
		s.N = box_normal(P, f3(g0), f3(ldg4(&S.prim_geom[4 * id + 3])));
	return s;
}

// cooperative cancel / scheduler abort, looked at whenever a warp fetches work
__device__ __forceinline__ bool frame_cancelled(const WaveState *ws, const FrameParams &F)
{
	return *(const volatile uint32_t *)&ws->overflow >= 2u || *(const volatile uint32_t *)&ws->stop_epoch == F.epoch;
}

// ---- fused wave kernel: closest hit of level `level` + shadow rays of level `level - 1` -------------

// DEFER: Model BVHs are walked with deferred triangle tests (rt_defer.cuh); the counting launches (STATS) keep the voted walk,
// whose node / triangle counts are the algorithmic ones (no speculative visits).
template<int WALK> struct WarpWork { typedef WarpDefer T; };
template<> struct WarpWork<2> { typedef WarpSteal T; };

template<bool STATS, int CTAS, int DEFER /* 0: voted walk, 1: deferred triangle tests, 2: work stealing inside the warp */>
__global__ void __launch_bounds__(RT_BLOCK, CTAS) k_wave(SceneDev S, const FrameParams *__restrict__ Fp, LevelBuf L, LevelBuf N, LevelBuf Lprev,
	WaveState *ws, uint32_t level, uint32_t traceOn, uint32_t shadowOn, float zNear)
{
	const FrameParams &F = *Fp;
	TravStats st = { 0, 0, 0 };
	__shared__ typename WarpWork<DEFER>::T wdefer[DEFER ? RT_BLOCK / 32 : 1];
	typename WarpWork<DEFER>::T *W = &wdefer[DEFER ? (threadIdx.x >> 5) : 0];

	// ---- phase A: closest hit + surface attributes + children ------------------------------------
	if (traceOn)
	{
		const uint32_t n = ws->count[level] < L.capacity ? ws->count[level] : L.capacity;
		const bool refraction = F.type != RT_TYPE_REFLECT;
		const bool deeper = level + 1 <= F.max_level;
		const uint32_t batch = fetch_batch(n);
		while (true)
		{
			const uint32_t base = warp_fetch(&ws->head_trace[level], batch);
			if (base >= n || frame_cancelled(ws, F))   // queue dry, or rt_stop
				break;
			const uint32_t lane = threadIdx.x & 31u;
			uint32_t i = lane < batch ? base + lane : 0xFFFFFFFFu;
			if (L.order && i < n) i = L.order[i];   // coherence binning: the k-th ray to trace is slot order[k]
			bool surface = false, wantFlec = false, wantFrac = false;
			float4 co = make_float4(0, 0, 0, 0), cdFlec = co, cdFrac = co;
			uint2 metaFlec = make_uint2(0, 0), metaFrac = metaFlec;
			float fracRfr = 1.0f;
			int4 aux = make_int4(-1, -1, -1, 0);
			// primary rays are made here (sched_flags bit 2: k_raygen did not run): no 40-byte record written
			// by one kernel and read back by the next
			const bool made = level == 0u && (F.sched_flags & 4u);
			RayD ray;
			ray.o = f3(0.0f, 0.0f, 0.0f), ray.d = f3(0.0f, 0.0f, 1.0f);
			ray.mtlrfr = 1.0f, ray.skip = RT_ID_NONE, ray.type = MY_RAY_BASERAY_, ray.isInside = 0;
			if (i < n)
			{
				if (made)
					ray.d = primary_dir(F, i, ray.o);
				else
					ray = load_ray(L, i);
			}
			Best best = { 1e20f, RT_ID_NONE, ray.skip };
			bool done = false;
			if (DEFER)
				trace_scene_warp<false>(S, ray, i < n, best, done, *W);   // the whole warp: idle lanes help with the triangle test rounds
			if (i < n)
			{
				const uint32_t nodes0 = st.nodes;
				if (!DEFER) trace_scene<false, STATS>(S, ray, best, done, st);
				if (STATS) atomicAdd(&ws->node_hist[min(11, 31 - __clz((int)(st.nodes - nodes0 + 1u)))], 1u);
				const F3 P = ray.o + ray.d * best.t;
				L.hit_p[i] = make_float4(P.x, P.y, P.z, best.t);
				L.hit_id[i] = make_uint2(best.id, best.newobj);
				// early cut of RayTracer.cpp:467-468: no surface -> Color(false), no light loop, no children
				surface = !(best.t > F.zFar || best.t < zNear);
				if (!surface)
					L.color[i] = make_float4(0.0f, 0.0f, 0.0f, 1e20f);
				else
				{
					const Surface sf = surface_attributes(S, ray, P, best.id);
					L.hit_n[i] = make_float4(sf.N.x, sf.N.y, sf.N.z, __int_as_float(sf.mtl));
					L.hit_uv[i] = make_float4(sf.tu, sf.tv, __int_as_float(sf.tex), 0.0f);
					const float4 mP = ldg4(&S.materials[4 * sf.mtl + 3]);   // shiness, reflect, refract, rfr
					const float bwc = made ? 1.0f : L.ray_d[i].w;
					if (made)
						L.ray_d[i] = make_float4(ray.d.x, ray.d.y, ray.d.z, 1.0f);   // the view direction of a surface is read again by k_shade
					aux.z = sf.mtl;
					co = make_float4(P.x, P.y, P.z, 1.0f);
					if (mP.y > 0.01f)
					{
						// reflection, RayTracer.cpp:548-563
						aux.w |= 1;
						const float bw = bwc * mP.y;
						if (deeper && !(bw < 1e-5f))
						{
							const float n_n = 2 * dot(ray.d, sf.N);
							const F3 r = normalize(ray.d - sf.N * n_n);
							wantFlec = true;
							cdFlec = make_float4(r.x, r.y, r.z, bw);
							metaFlec = make_uint2(best.newobj, refraction ? (uint32_t)MY_RAY_REFLECTRAY_ : 0u);
						}
					}
					if (refraction && mP.z > 0.01f)
					{
						// refraction, RayTracer.cpp:565-583
						aux.w |= 2;
						if (sf.isInside) aux.w |= 4;
						const float nn = ray.mtlrfr / sf.rfr;
						const float cosIn = -dot(ray.d, sf.N);
						const float cosOut2 = 1.0f - (nn * nn) * (1.0f - cosIn * cosIn);
						const float bw = bwc * mP.z;
						if (!(cosOut2 < 0.0f) && deeper && !(bw < 1e-5f))
						{
							const F3 l2 = ray.d * nn, l1 = sf.N * (nn * cosIn - sqrtf(cosOut2));
							const F3 r = normalize(l1 + l2);
							wantFrac = true;
							cdFrac = make_float4(r.x, r.y, r.z, bw);
							metaFrac = make_uint2(best.newobj, (uint32_t)MY_RAY_REFRACTRAY_ | ((uint32_t)sf.isInside << 8));
							fracRfr = sf.rfr;
						}
					}
				}
			}
			// queue appends: the whole warp takes part
			const uint32_t hslot = warp_append(&ws->n_hit[level], surface);
			if (surface)
				L.hit_list[hslot] = i + 1u;
			const uint32_t sFlec = warp_append(&ws->count[level + 1], wantFlec);
			if (wantFlec)
			{
				if (sFlec < N.capacity)
				{
					N.ray_o[sFlec] = co, N.ray_d[sFlec] = cdFlec, N.ray_meta[sFlec] = metaFlec;
					aux.x = (int)sFlec;
				}
				else
					ws->overflow = 1;
			}
			const uint32_t sFrac = warp_append(&ws->count[level + 1], wantFrac);
			if (wantFrac)
			{
				if (sFrac < N.capacity)
				{
					N.ray_o[sFrac] = make_float4(co.x, co.y, co.z, fracRfr), N.ray_d[sFrac] = cdFrac, N.ray_meta[sFrac] = metaFrac;
					aux.y = (int)sFrac;
				}
				else
					ws->overflow = 1;
			}
			if (i < n)
				L.aux[i] = aux;
			const uint32_t mf = __ballot_sync(0xffffffffu, wantFlec), mr = __ballot_sync(0xffffffffu, wantFrac);
			if ((threadIdx.x & 31) == 0)
			{
				if (mf) atomicAdd(&ws->n_reflect, (unsigned long long)__popc(mf));
				if (mr) atomicAdd(&ws->n_refract, (unsigned long long)__popc(mr));
			}
		}
	}

	// ---- phase B: shadow any-hit of the previous level's surfaces -----------------------------------
	// work item w = (enabled light w / n_hit, surface w % n_hit): a warp's rays go to one light
	if (shadowOn)
	{
		const uint32_t lp = level - 1u;
		const uint32_t nHit = ws->n_hit[lp];
		const uint32_t n = nHit * F.n_enabled;
		const uint32_t batch = fetch_batch(n);
		while (true)
		{
			const uint32_t base = warp_fetch(&ws->head_shadow[lp], batch);
			if (base >= n || frame_cancelled(ws, F))
				break;
			const uint32_t lane = threadIdx.x & 31u;
			const uint32_t w = lane < batch ? base + lane : 0xFFFFFFFFu;
			uint32_t k = 0, i = 0;
			RayD ray;
			ray.o = f3(0.0f, 0.0f, 0.0f), ray.d = f3(0.0f, 0.0f, 1.0f);
			ray.mtlrfr = 1.0f, ray.skip = RT_ID_NONE, ray.type = 0, ray.isInside = 0;
			float dis = 0.0f;
			if (w < n)
			{
				k = F.enabled_index[w / nHit], i = Lprev.hit_list[w % nHit] - 1u;
				const float4 hp = Lprev.hit_p[i];
				float lum;
				light_dir(F.lights[k], f3(hp), ray.d, dis, lum);
				ray.o = f3(hp);
				ray.skip = Lprev.hit_id[i].y;
				ray.type = (F.type == RT_TYPE_REFLECT || F.type == RT_TYPE_SHADOW) ? 0 : MY_RAY_SHADOWRAY_;
			}
			Best best = { dis, RT_ID_NONE, RT_ID_NONE };
			bool done = false;
			if (DEFER)
				trace_scene_warp<true>(S, ray, w < n, best, done, *W);
			if (w < n)
			{
				const uint32_t nodes0 = st.nodes;
				if (!DEFER) trace_scene<true, STATS>(S, ray, best, done, st);
				if (STATS) atomicAdd(&ws->node_hist[12 + min(11, 31 - __clz((int)(st.nodes - nodes0 + 1u)))], 1u);
				Lprev.shadow[(size_t)k * Lprev.capacity + i] = done ? 1 : 0;
			}
		}
	}
	flush_stats<STATS>(ws, st);
}

// (A lane-asynchronous variant -- per-warp ray pool, lanes swap rays in one by one -- was measured 35 % slower and removed:
// profiles/r2_fma_slab_ab.txt.)

// One finished closest-hit ray: hit record, early cut, surface attributes, child rays, queue appends (the epilogue
// of k_wave's phase A).  Called by ALL 32 lanes (lanes without a ray pass valid = false): appends are warp-aggregated.
__device__ __forceinline__ void trace_epilogue(const SceneDev &S, const FrameParams &F, const LevelBuf &L, const LevelBuf &N, WaveState *ws,
	uint32_t level, float zNear, bool valid, uint32_t i, const RayD &ray, float bwc, bool made, const Best &best)
{
	const bool refraction = F.type != RT_TYPE_REFLECT;
	const bool deeper = level + 1 <= F.max_level;
	bool surface = false, wantFlec = false, wantFrac = false;
	float4 co = make_float4(0, 0, 0, 0), cdFlec = co, cdFrac = co;
	uint2 metaFlec = make_uint2(0, 0), metaFrac = metaFlec;
	float fracRfr = 1.0f;
	int4 aux = make_int4(-1, -1, -1, 0);
	if (valid)
	{
		const F3 P = ray.o + ray.d * best.t;
		L.hit_p[i] = make_float4(P.x, P.y, P.z, best.t);
		L.hit_id[i] = make_uint2(best.id, best.newobj);
		// early cut of RayTracer.cpp:467-468: no surface -> Color(false), no light loop, no children
		surface = !(best.t > F.zFar || best.t < zNear);
		if (!surface)
			L.color[i] = make_float4(0.0f, 0.0f, 0.0f, 1e20f);
		else
		{
			const Surface sf = surface_attributes(S, ray, P, best.id);
			L.hit_n[i] = make_float4(sf.N.x, sf.N.y, sf.N.z, __int_as_float(sf.mtl));
			L.hit_uv[i] = make_float4(sf.tu, sf.tv, __int_as_float(sf.tex), 0.0f);
			const float4 mP = ldg4(&S.materials[4 * sf.mtl + 3]);   // shiness, reflect, refract, rfr
			if (made)
				L.ray_d[i] = make_float4(ray.d.x, ray.d.y, ray.d.z, 1.0f);   // the view direction of a surface is read again by k_shade
			aux.z = sf.mtl;
			co = make_float4(P.x, P.y, P.z, 1.0f);
			if (mP.y > 0.01f)
			{
				// reflection, RayTracer.cpp:548-563
				aux.w |= 1;
				const float bw = bwc * mP.y;
				if (deeper && !(bw < 1e-5f))
				{
					const float n_n = 2 * dot(ray.d, sf.N);
					const F3 r = normalize(ray.d - sf.N * n_n);
					wantFlec = true;
					cdFlec = make_float4(r.x, r.y, r.z, bw);
					metaFlec = make_uint2(best.newobj, refraction ? (uint32_t)MY_RAY_REFLECTRAY_ : 0u);
				}
			}
			if (refraction && mP.z > 0.01f)
			{
				// refraction, RayTracer.cpp:565-583
				aux.w |= 2;
				if (sf.isInside) aux.w |= 4;
				const float nn = ray.mtlrfr / sf.rfr;
				const float cosIn = -dot(ray.d, sf.N);
				const float cosOut2 = 1.0f - (nn * nn) * (1.0f - cosIn * cosIn);
				const float bw = bwc * mP.z;
				if (!(cosOut2 < 0.0f) && deeper && !(bw < 1e-5f))
				{
					const F3 l2 = ray.d * nn, l1 = sf.N * (nn * cosIn - sqrtf(cosOut2));
					const F3 r = normalize(l1 + l2);
					wantFrac = true;
					cdFrac = make_float4(r.x, r.y, r.z, bw);
					metaFrac = make_uint2(best.newobj, (uint32_t)MY_RAY_REFRACTRAY_ | ((uint32_t)sf.isInside << 8));
					fracRfr = sf.rfr;
				}
			}
		}
	}
	const uint32_t hslot = warp_append(&ws->n_hit[level], surface);
	if (surface)
		L.hit_list[hslot] = i + 1u;
	const uint32_t sFlec = warp_append(&ws->count[level + 1], wantFlec);
	if (wantFlec)
	{
		if (sFlec < N.capacity)
		{
			N.ray_o[sFlec] = co, N.ray_d[sFlec] = cdFlec, N.ray_meta[sFlec] = metaFlec;
			aux.x = (int)sFlec;
		}
		else
			ws->overflow = 1;
	}
	const uint32_t sFrac = warp_append(&ws->count[level + 1], wantFrac);
	if (wantFrac)
	{
		if (sFrac < N.capacity)
		{
			N.ray_o[sFrac] = make_float4(co.x, co.y, co.z, fracRfr), N.ray_d[sFrac] = cdFrac, N.ray_meta[sFrac] = metaFrac;
			aux.y = (int)sFrac;
		}
		else
			ws->overflow = 1;
	}
	if (valid)
		L.aux[i] = aux;
	const uint32_t mf = __ballot_sync(0xffffffffu, wantFlec), mr = __ballot_sync(0xffffffffu, wantFrac);
	if ((threadIdx.x & 31) == 0)
	{
		if (mf) atomicAdd(&ws->n_reflect, (unsigned long long)__popc(mf));
		if (mr) atomicAdd(&ws->n_refract, (unsigned long long)__popc(mr));
	}
}

// ---- two-stage wave kernel ------------------------------------------------------------------------------
// k_wave with the scene walk split into head and tail (rt_traverse.cuh "two stages"): per warp, rays that have to walk the
// last Model's BVH wait in a small shared-memory list until 32 of them are there; finished rays wait in a second list for
// an epilogue that also runs 32 at a time.  Everything is warp-converged, so the list lengths live in registers.
struct SplitLists
{
	uint4 todo[64];   // closest: slot, best.t, best.id, best.newobj   |   shadow: work index w in .x
	uint4 fin[64];    // closest: slot, best.t, best.id, best.newobj
};

// the ray in slot i of level L (or, for primary rays made in place, through its pixel)
__device__ __forceinline__ RayD wave_ray(const FrameParams &F, const LevelBuf &L, uint32_t i, bool made, float &bwc)
{
	RayD ray;
	if (made)
	{
		ray.d = primary_dir(F, i, ray.o);
		ray.mtlrfr = 1.0f, ray.skip = RT_ID_NONE, ray.type = MY_RAY_BASERAY_, ray.isInside = 0;
		bwc = 1.0f;
	}
	else
	{
		ray = load_ray(L, i);
		bwc = L.ray_d[i].w;
	}
	return ray;
}

template<bool STATS, int CTAS>
__global__ void __launch_bounds__(RT_BLOCK, CTAS) k_wave_split(SceneDev S, const FrameParams *__restrict__ Fp, LevelBuf L, LevelBuf N, LevelBuf Lprev,
	WaveState *ws, uint32_t level, uint32_t traceOn, uint32_t shadowOn, float zNear)
{
	__shared__ SplitLists lists[RT_BLOCK / 32];
	SplitLists &Q = lists[threadIdx.x >> 5];
	const FrameParams &F = *Fp;
	const uint32_t lane = threadIdx.x & 31u, lt = lanemask_lt();
	const bool hasTail = scene_has_tail(S);
	TravStats st = { 0, 0, 0 };

	// ---- phase A: closest hit + surface attributes + children ------------------------------------
	if (traceOn)
	{
		const uint32_t n = ws->count[level] < L.capacity ? ws->count[level] : L.capacity;
		const uint32_t batch = fetch_batch(n);
		const bool made = level == 0u && (F.sched_flags & 4u);
		uint32_t nTodo = 0, nFin = 0;
		bool dry = n == 0u;
		while (true)
		{
			if (nFin >= 32u || (dry && nTodo == 0u && nFin > 0u))
			{
				// ---- epilogue of 32 finished rays ----
				const uint32_t cnt = nFin < 32u ? nFin : 32u;
				const bool valid = lane < cnt;
				const uint4 e = Q.fin[nFin - cnt + (valid ? lane : 0u)];
				nFin -= cnt;
				RayD ray;
				float bwc = 1.0f;
				if (valid) ray = wave_ray(F, L, e.x, made, bwc);
				const Best best = { __uint_as_float(e.y), e.z, e.w };
				trace_epilogue(S, F, L, N, ws, level, zNear, valid, e.x, ray, bwc, made, best);
				__syncwarp();
			}
			else if (nTodo >= 32u || (dry && nTodo > 0u))
			{
				// ---- tail: 32 rays that passed the last Model's box test walk its BVH, full lanes ----
				const uint32_t cnt = nTodo < 32u ? nTodo : 32u;
				const bool valid = lane < cnt;
				const uint4 e = Q.todo[nTodo - cnt + (valid ? lane : 0u)];
				nTodo -= cnt;
				Best best = { __uint_as_float(e.y), e.z, e.w };
				if (valid)
				{
					float bwc;
					const RayD ray = wave_ray(F, L, e.x, made, bwc);
					bool done = false;
					trace_scene_tail<false, STATS>(S, ray, best, done, st);
				}
				__syncwarp();
				if (valid) Q.fin[nFin + lane] = make_uint4(e.x, __float_as_uint(best.t), best.id, best.newobj);
				nFin += cnt;
				__syncwarp();
			}
			else if (!dry)
			{
				// ---- head: 32 new rays, everything before the last Model's BVH ----
				const uint32_t base = warp_fetch(&ws->head_trace[level], batch);
				if (base >= n || frame_cancelled(ws, F))
				{
					dry = true;
					continue;
				}
				const uint32_t i = lane < batch ? base + lane : 0xFFFFFFFFu;
				const bool valid = i < n;
				bool enter = false;
				Best best = { 1e20f, RT_ID_NONE, RT_ID_NONE };
				if (valid)
				{
					float bwc;
					const RayD ray = wave_ray(F, L, i, made, bwc);
					best.newobj = ray.skip;
					bool done = false;
					enter = trace_scene_head<false, STATS>(S, ray, best, done, st, hasTail);
				}
				const uint32_t mT = __ballot_sync(0xffffffffu, valid && enter), mF = __ballot_sync(0xffffffffu, valid && !enter);
				const uint4 rec = make_uint4(i, __float_as_uint(best.t), best.id, best.newobj);
				if (valid && enter) Q.todo[nTodo + __popc(mT & lt)] = rec;
				if (valid && !enter) Q.fin[nFin + __popc(mF & lt)] = rec;
				nTodo += __popc(mT), nFin += __popc(mF);
				if (base + batch >= n) dry = true;
				__syncwarp();
			}
			else
				break;
		}
	}

	// ---- phase B: shadow any-hit of the previous level's surfaces -----------------------------------
	if (shadowOn)
	{
		__syncwarp();
		const uint32_t lp = level - 1u;
		const uint32_t nHit = ws->n_hit[lp];
		const uint32_t n = nHit * F.n_enabled;
		const uint32_t batch = fetch_batch(n);
		const uint8_t rayType = (F.type == RT_TYPE_REFLECT || F.type == RT_TYPE_SHADOW) ? 0 : MY_RAY_SHADOWRAY_;
		uint32_t nTodo = 0;
		bool dry = n == 0u;
		while (true)
		{
			const bool tail = nTodo >= 32u || (dry && nTodo > 0u);
			if (!tail && dry)
				break;
			uint32_t w = 0xFFFFFFFFu;
			if (tail)
			{
				const uint32_t cnt = nTodo < 32u ? nTodo : 32u;
				if (lane < cnt) w = Q.todo[nTodo - cnt + lane].x;
				nTodo -= cnt;
			}
			else
			{
				const uint32_t base = warp_fetch(&ws->head_shadow[lp], batch);
				if (base >= n || frame_cancelled(ws, F))
				{
					dry = true;
					continue;
				}
				if (lane < batch && base + lane < n) w = base + lane;
				if (base + batch >= n) dry = true;
			}
			bool enter = false;
			if (w != 0xFFFFFFFFu)
			{
				// the shadow ray of work item w = (enabled light w / n_hit, surface w % n_hit); made again for the tail (a few
				// L2 hits and one normalisation against 16 bytes of shared memory per waiting ray)
				const uint32_t k = F.enabled_index[w / nHit], i = Lprev.hit_list[w % nHit] - 1u;
				const float4 hp = Lprev.hit_p[i];
				RayD ray;
				float dis, lum;
				light_dir(F.lights[k], f3(hp), ray.d, dis, lum);
				ray.o = f3(hp);
				ray.mtlrfr = 1.0f;
				ray.skip = Lprev.hit_id[i].y;
				ray.type = rayType;
				ray.isInside = 0;
				Best best = { dis, RT_ID_NONE, RT_ID_NONE };
				bool done = false;
				if (tail)
					trace_scene_tail<true, STATS>(S, ray, best, done, st);
				else
					enter = trace_scene_head<true, STATS>(S, ray, best, done, st, hasTail);
				if (!enter)
					Lprev.shadow[(size_t)k * Lprev.capacity + i] = done ? 1 : 0;
			}
			__syncwarp();
			if (!tail)
			{
				const uint32_t mT = __ballot_sync(0xffffffffu, enter);
				if (enter) Q.todo[nTodo + __popc(mT & lt)].x = w;
				nTodo += __popc(mT);
				__syncwarp();
			}
		}
	}
	flush_stats<STATS>(ws, st);
}

// ---- whole-frame persistent scheduler ---------------------------------------------------------------
//
// One launch traces every ray of the frame.  Resident warps repeatedly take a batch of rays from
// the shallowest level that has unconsumed rays, trace them (closest hit), spawn the reflect /
// refract children into the next level's queue, and then trace the shadow rays of their own hits
// right away (one round per enabled light, lane i keeps working on the surface it just found).
// Level l+1 rays become visible to other warps the moment they are written, so the long rays of a
// level no longer hold back the next level: the critical path of a frame is one pixel's ray chain,
// not (levels x slowest ray).  That is what the per-level waves could not give, and what strong
// scaling of a 1080p frame over several GPUs needs.
//
// Publication protocol.  A producer reserves slots with one atomicAdd on count[l+1] (and adds them
// to `outstanding` at the same time), writes ray_o/ray_d, fences, and only then writes ray_meta,
// whose upper 16 bits carry this frame's epoch.  A consumer that was handed slot i spins until that
// epoch shows up (the producer is a running warp that is not waiting for anyone), fences, and reads
// the slot through L2 (__ldcg: another SM wrote it).  `outstanding` never under-counts, so
// "outstanding == 0" means the frame is complete and every warp may leave.
__device__ __forceinline__ uint32_t vload(const uint32_t *p) { return *(const volatile uint32_t *)p; }

// Release / acquire without wiping the L1.  __threadfence() is MEMBAR.SC.GPU + CCTL.IVALL on sm_100: the
// CCTL invalidates every L1 line of the SM, and k_frame fences several times per batch of 32 rays -- with
// 32 resident warps that flushed the SM's L1 (BVH nodes, triangles) every few hundred cycles.  Neither
// side needs the invalidate: producers only have to make their payload visible before the stamp
// (st.release.gpu = MEMBAR.ALL.GPU + strong store, no CCTL), and consumers read stamp and payload
// through L2 (ld.cg), so ordering the two loads (CTA-scope fence, MEMBAR.ALL.CTA) is enough.
// RT_LIGHT_FENCE=0 restores the __threadfence() pairs (A/B).
#ifndef RT_LIGHT_FENCE
#define RT_LIGHT_FENCE 1
#endif
__device__ __forceinline__ void consumer_fence()
{
#if RT_LIGHT_FENCE
	asm volatile("fence.acq_rel.cta;" ::: "memory");
#else
	__threadfence();
#endif
}
struct Publisher
{
	bool released = false;   // this lane's payload stores are already ordered before what it publishes next
	__device__ __forceinline__ void store(uint2 *p, uint2 v)
	{
#if RT_LIGHT_FENCE
		if (!released) asm volatile("st.release.gpu.global.v2.u32 [%0], {%1, %2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
		else asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
#else
		if (!released) __threadfence();
		*p = v;
#endif
		released = true;
	}
	__device__ __forceinline__ void store(uint32_t *p, uint32_t v)
	{
#if RT_LIGHT_FENCE
		if (!released) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
		else asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
#else
		if (!released) __threadfence();
		*p = v;
#endif
		released = true;
	}
};

// queue q of the claim scan -> level | light << 8: ray queues first (shallowest or deepest level first),
// then the shadow queues by (level, light)
__device__ __forceinline__ uint32_t queue_code(uint32_t q, uint32_t nL, uint32_t nEnabled, uint32_t schedFlags)
{
	if (q < nL)
		return (schedFlags & 1u) ? nL - 1u - q : q;
	const uint32_t s = q - nL, e = nEnabled ? nEnabled : 1u;
	return (s / e) | ((s % e) << 8);
}

template<bool STATS, int CTAS>
__global__ void __launch_bounds__(RT_BLOCK, CTAS) k_frame(SceneDev S, const FrameParams *__restrict__ Fp, LevelSet LS, WaveState *ws)
{
	const FrameParams &F = *Fp;
	const uint32_t lane = threadIdx.x & 31u;
	const bool refraction = F.type != RT_TYPE_REFLECT;
	const bool wantShadows = F.type != RT_TYPE_DEPTH && F.type != RT_TYPE_NORMAL && F.type != RT_TYPE_TEXTURE && F.type != RT_TYPE_MATERIAL;
	const bool genPrimary = (F.sched_flags & 4u) != 0u;   // level-0 rays are generated here, k_raygen did not run
	TravStats st = { 0, 0, 0 };
	uint32_t idleSpins = 0;
	// Work this warp has claimed but not done yet.  A claim (one atomicAdd on a queue head) may run
	// past what has been published so far; such a slot stays owned by its lane, which polls it without
	// blocking the lanes whose work is ready, until it is published or the frame is over.
	// pKind 0: closest-hit rays of level pLevel; pKind 1: shadow rays of level pLevel towards light pLight.
	uint32_t pKind = 0, pLevel = 0, pLight = 0, pSlot = 0xFFFFFFFFu;
	const uint32_t nPix0 = F.pix_per_frame * F.batch;   // level-0 slots of this launch (all frames of the batch)
	const uint32_t nL = F.max_level + 1u, nQ = nL + (wantShadows ? nL * F.n_enabled : 0u);
	const uint32_t myQueue = queue_code(lane, nL, F.n_enabled, F.sched_flags);
	uint32_t statNodes = 0, statKind = 0;   // RT_FLAG_STATS: node visits of the lane's last ray
	while (true)
	{
		if (STATS)
		{
			if (__ballot_sync(0xffffffffu, statNodes != 0u)) lane_stats<STATS>(ws, statKind, statNodes);
			statNodes = 0;
		}
		if (__ballot_sync(0xffffffffu, pSlot != 0xFFFFFFFFu) == 0u)
		{
			// ---- claim: the 32 lanes look at one queue each (rays of level q, then shadow rays per (level,
			// light)), so a claim costs one L2 round trip however many queues there are; the first queue in
			// that order wins -- rays before shadow rays (rays create work), shallowest level first.  Work is
			// consumed while it is being produced, so queues are usually short and taking whatever is there
			// would hand every warp 2-3 rays: pass 0 only takes full batches of 32, pass 1 (nothing full
			// anywhere) whatever exists, so the tail of the frame still drains. ----------------------------
			uint32_t kind = 0xFFFFFFFFu, level = 0, light = 0, base = 0, nb = 0;
			for (int pass = 0; pass < 2 && kind == 0xFFFFFFFFu; ++pass)
				for (uint32_t q0 = 0; q0 < nQ && kind == 0xFFFFFFFFu; q0 += 32u)
				{
					const uint32_t q = q0 + lane;
					uint32_t avail = 0, qLevel = 0, qLight = 0, cap = 0;
					uint32_t *head = nullptr;
					if (q < nQ)
					{
						uint32_t cnt;
						// which queue this lane looks at: decoded once per frame for the first 32 queues (myQueue)
						const uint32_t code = q0 == 0u ? myQueue : queue_code(q, nL, F.n_enabled, F.sched_flags);
						qLevel = code & 0xFFu, qLight = code >> 8;
						if (q < nL)
							cnt = vload(&ws->count[qLevel]), head = &ws->head_trace[qLevel];
						else
							cnt = vload(&ws->n_hit[qLevel]), head = &ws->head_light[qLevel][qLight];
						cap = LS.l[qLevel].capacity;
						cnt = cnt < cap ? cnt : cap;
						const uint32_t h = vload(head);
						avail = h < cnt ? cnt - h : 0u;
					}
					uint32_t cand = __ballot_sync(0xffffffffu, pass == 0 ? avail >= 32u : avail > 0u);
					while (cand && kind == 0xFFFFFFFFu)
					{
						const uint32_t src = __ffs((int)cand) - 1u;
						cand &= cand - 1u;
						uint32_t got = 0;
						const uint32_t want = avail > 32u ? 32u : avail;
						if (lane == src) got = atomicAdd(head, want);
						got = __shfl_sync(0xffffffffu, got, src);
						if (got < __shfl_sync(0xffffffffu, cap, src))
						{
							kind = src + q0 < nL ? 0u : 1u;
							level = __shfl_sync(0xffffffffu, qLevel, src), light = __shfl_sync(0xffffffffu, qLight, src);
							base = got, nb = __shfl_sync(0xffffffffu, want, src);
						}
					}
				}
			if (kind != 0xFFFFFFFFu)
			{
				pKind = kind, pLevel = level, pLight = light;
				if (lane < nb && base + lane < LS.l[level].capacity) pSlot = base + lane;
			}
		}
		const uint32_t level = pLevel;
		const LevelBuf &L = LS.l[level];
		const LevelBuf &N = LS.l[level + 1];
		// which of the owned slots are published?
		uint2 m = make_uint2(0, 0);
		uint32_t hitEntry = 0;
		bool ready = false;
		if (pSlot != 0xFFFFFFFFu)
		{
			if (pKind == 0u)
			{
				if (level == 0u && genPrimary)
				{
					// made in place below.  Two warps that race for the last batch can both be handed slots; the
					// loser's lie past the frame (the level's capacity may be larger than this frame): not rays
					if (pSlot < nPix0) m = make_uint2(RT_ID_NONE, (uint32_t)MY_RAY_BASERAY_ | (F.epoch << 16)), ready = true;
					else pSlot = 0xFFFFFFFFu;
				}
				else
				{
					m = __ldcg(&L.ray_meta[pSlot]);
					ready = (m.y >> 16) == F.epoch;
				}
			}
			else
			{
				hitEntry = __ldcg(&L.hit_list[pSlot]);
				ready = hitEntry != 0u;
			}
		}
		const uint32_t readyMask = __ballot_sync(0xffffffffu, ready);
		if (readyMask == 0u)
		{
			int out = 0;
			if (lane == 0) out = *(volatile int *)&ws->outstanding;
			out = __shfl_sync(0xffffffffu, out, 0);
			if (out <= 0 || frame_cancelled(ws, F))
				break;   // nothing in flight any more (or the scheduler gave up / rt_stop): unpublished slots will never be written
			// Retire: when the frame runs thin (a few long ray chains are left), a CTA that owns nothing and
			// finds nothing leaves for good if fewer than `retire_rays` rays per CTA would remain for it -- high CTA
			// indices first, the first `sms` CTAs never.  `outstanding` never under-counts and a ray tree is
			// bounded by max_level, so whoever stays can always finish the frame.  Leaving frees the SM slots
			// for the next frame's kernels (frames in flight) and thins out the pollers of the queue heads.
			const bool mayRetire = F.keep_div <= 1u ? blockIdx.x >= F.sms : (blockIdx.x + blockIdx.x / F.sms + F.keep_salt) % F.keep_div != 0u;
			const uint32_t retireRank = F.keep_div <= 1u ? blockIdx.x - F.sms : blockIdx.x;
			if ((F.sched_flags & 2u) && mayRetire && out < (int)((retireRank + 1u) * F.retire_rays)
				&& __ballot_sync(0xffffffffu, pSlot != 0xFFFFFFFFu) == 0u)
				break;
			if (++idleSpins > (1u << 22))
			{
				if (lane == 0) ws->overflow = 2u;   // scheduler stuck: fail loudly instead of hanging the GPU
				break;
			}
			// back off quickly: idle warps poll the very cache lines the busy warps' atomics hit
			__nanosleep(idleSpins < 4u ? 500 : (idleSpins < 16u ? 2000 : 8000));
			continue;
		}
		idleSpins = 0;
		if (frame_cancelled(ws, F))
			break;   // rt_stop or scheduler abort
		const uint32_t nb = __popc(readyMask);
		if (STATS && lane == 0)
		{
			unsigned long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			unsigned long long t0 = atomicCAS(&ws->t0, 0ull, now);
			if (t0 == 0ull) t0 = now;
			const unsigned long long bin = (now - t0) >> 14;
			atomicAdd(&ws->timeline[bin < 127ull ? bin : 127ull][(pKind ? 8u : 0u) + (pLevel < 7u ? pLevel : 7u)], nb);
		}

		if (pKind == 1u)
		{
			// ---- shadow any-hit rays: a warp's lanes go to the same light from neighbouring surfaces ------
			if (ready)
			{
				consumer_fence();
				const uint32_t i = hitEntry - 1u, k = F.enabled_index[pLight];
				const float4 hp = __ldcg(&L.hit_p[i]);
				RayD ray;
				float dis, lum;
				light_dir(F.lights[k], f3(hp), ray.d, dis, lum);
				ray.o = f3(hp);
				ray.mtlrfr = 1.0f;
				ray.skip = __ldcg(&L.hit_id[i]).y;
				ray.type = (F.type == RT_TYPE_REFLECT || F.type == RT_TYPE_SHADOW) ? 0 : MY_RAY_SHADOWRAY_;
				ray.isInside = 0;
				Best best = { dis, RT_ID_NONE, RT_ID_NONE };
				bool done = false;
				const uint32_t nodes0 = st.nodes;
				trace_scene<true, STATS>(S, ray, best, done, st);
				if (STATS) atomicAdd(&ws->node_hist[12 + min(11, 31 - __clz((int)(st.nodes - nodes0 + 1u)))], 1u);
				if (STATS) statNodes = st.nodes - nodes0, statKind = 1;
				L.shadow[(size_t)k * L.capacity + i] = done ? 1 : 0;
				pSlot = 0xFFFFFFFFu;
			}
			__syncwarp();
			if (lane == 0) atomicSub(&ws->outstanding, (int)nb);
			continue;
		}

		const float zNear = level == 0 ? F.zNear : 0.0f;
		const bool deeper = level + 1 <= F.max_level;
		const uint32_t i = ready ? pSlot : 0xFFFFFFFFu;
		if (ready) pSlot = 0xFFFFFFFFu;

		if (i != 0xFFFFFFFFu)
		{
			bool surface = false, wantFlec = false, wantFrac = false;
			float4 co = make_float4(0, 0, 0, 0), cdFlec = co, cdFrac = co;
			uint2 metaFlec = make_uint2(0, 0), metaFrac = metaFlec;
			float fracRfr = 1.0f;
			int4 aux = make_int4(-1, -1, -1, 0);
			float4 o4, d4;
			if (level == 0u && genPrimary)
			{
				// primary rays never exist in memory: no k_raygen launch, no 40-byte record written and read back
				F3 o;
				const F3 d = primary_dir(F, i, o);
				o4 = make_float4(o.x, o.y, o.z, 1.0f), d4 = make_float4(d.x, d.y, d.z, 1.0f);
			}
			else
			{
				consumer_fence();   // the stamp was seen: order the payload reads after it
				o4 = __ldcg(&L.ray_o[i]), d4 = __ldcg(&L.ray_d[i]);
			}
			RayD ray;
			ray.o = f3(o4), ray.d = f3(d4), ray.mtlrfr = o4.w;
			ray.skip = m.x, ray.type = (uint8_t)(m.y & 0xFF), ray.isInside = (uint8_t)((m.y >> 8) & 0xFF);
			Best best = { 1e20f, RT_ID_NONE, ray.skip };
			bool done = false;
			const uint32_t nodes0 = st.nodes;
			// EARLY: lanes whose ray is through leave the walk of the last scene item in groups, a few steps
			// after they finish, instead of waiting for the longest ray of the batch.  Everything below runs
			// per such group, so a ray's children and its shadow work are published when IT is done: the
			// frame's critical path is a chain of single rays, not a chain of batch maxima (measured on an
			// eighth of the C3 frame: 0.91 -> see DESIGN.md).
			trace_scene<false, STATS, true>(S, ray, best, done, st);
			if (STATS) atomicAdd(&ws->node_hist[min(11, 31 - __clz((int)(st.nodes - nodes0 + 1u)))], 1u);
			if (STATS) statNodes = st.nodes - nodes0, statKind = 0;
			const F3 P = ray.o + ray.d * best.t;
			L.hit_p[i] = make_float4(P.x, P.y, P.z, best.t);
			L.hit_id[i] = make_uint2(best.id, best.newobj);
			surface = !(best.t > F.zFar || best.t < zNear);
			if (!surface)
				L.color[i] = make_float4(0.0f, 0.0f, 0.0f, 1e20f);
			else
			{
				const Surface sf = surface_attributes(S, ray, P, best.id);
				L.hit_n[i] = make_float4(sf.N.x, sf.N.y, sf.N.z, __int_as_float(sf.mtl));
				L.hit_uv[i] = make_float4(sf.tu, sf.tv, __int_as_float(sf.tex), 0.0f);
				if (level == 0u && genPrimary)
					L.ray_d[i] = d4;   // the view direction of a surface is read again by k_shade
				const float4 mP = ldg4(&S.materials[4 * sf.mtl + 3]);   // shiness, reflect, refract, rfr
				const float bwc = d4.w;
				aux.z = sf.mtl;
				co = make_float4(P.x, P.y, P.z, 1.0f);
				if (mP.y > 0.01f)
				{
					aux.w |= 1;
					const float bw = bwc * mP.y;
					if (deeper && !(bw < 1e-5f))
					{
						const float n_n = 2 * dot(ray.d, sf.N);
						const F3 r = normalize(ray.d - sf.N * n_n);
						wantFlec = true;
						cdFlec = make_float4(r.x, r.y, r.z, bw);
						metaFlec = make_uint2(best.newobj, (refraction ? (uint32_t)MY_RAY_REFLECTRAY_ : 0u) | (F.epoch << 16));
					}
				}
				if (refraction && mP.z > 0.01f)
				{
					aux.w |= 2;
					if (sf.isInside) aux.w |= 4;
					const float nn = ray.mtlrfr / sf.rfr;
					const float cosIn = -dot(ray.d, sf.N);
					const float cosOut2 = 1.0f - (nn * nn) * (1.0f - cosIn * cosIn);
					const float bw = bwc * mP.z;
					if (!(cosOut2 < 0.0f) && deeper && !(bw < 1e-5f))
					{
						const F3 l2 = ray.d * nn, l1 = sf.N * (nn * cosIn - sqrtf(cosOut2));
						const F3 r = normalize(l1 + l2);
						wantFrac = true;
						cdFrac = make_float4(r.x, r.y, r.z, bw);
						metaFrac = make_uint2(best.newobj, (uint32_t)MY_RAY_REFRACTRAY_ | ((uint32_t)sf.isInside << 8) | (F.epoch << 16));
						fracRfr = sf.rfr;
					}
				}
			}

			// ---- new work: count it as outstanding BEFORE it can be consumed, then publish.  `grp` is whatever
			// set of lanes arrives here together; any partition of the batch into such groups is correct, each
			// group retires its own rays and adds its own children / shadow work in one atomic. ---------------
			const uint32_t grp = __activemask();
			const uint32_t gLeader = __ffs((int)grp) - 1;
			const uint32_t mf = __ballot_sync(grp, wantFlec), mr = __ballot_sync(grp, wantFrac), ms = __ballot_sync(grp, surface);
			const int nChildren = __popc(mf) + __popc(mr);
			const int nShadow = wantShadows ? __popc(ms) * (int)F.n_enabled : 0;
			const int nMine = __popc(grp);
			if (lane == gLeader && nChildren + nShadow != nMine) atomicAdd(&ws->outstanding, nChildren + nShadow - nMine);
			__syncwarp(grp);
			uint32_t sFlec = 0xFFFFFFFFu, sFrac = 0xFFFFFFFFu;
			if (nChildren)
			{
				sFlec = group_append(&ws->count[level + 1], wantFlec, grp);
				sFrac = group_append(&ws->count[level + 1], wantFrac, grp);
				int dropped = 0;
				if (wantFlec)
				{
					if (sFlec < N.capacity) { N.ray_o[sFlec] = co, N.ray_d[sFlec] = cdFlec; aux.x = (int)sFlec; }
					else { ws->overflow = 1; ++dropped; }
				}
				if (wantFrac)
				{
					if (sFrac < N.capacity) { N.ray_o[sFrac] = make_float4(co.x, co.y, co.z, fracRfr), N.ray_d[sFrac] = cdFrac; aux.y = (int)sFrac; }
					else { ws->overflow = 1; ++dropped; }
				}
				if (dropped) atomicSub(&ws->outstanding, dropped);
				if (lane == gLeader)
				{
					if (mf) atomicAdd(&ws->n_reflect, (unsigned long long)__popc(mf));
					if (mr) atomicAdd(&ws->n_refract, (unsigned long long)__popc(mr));
				}
			}
			L.aux[i] = aux;
			const uint32_t hslot = group_append(&ws->n_hit[level], surface, grp);
			// ---- publish: everything a consumer reads (child payloads, hit_p / hit_id / hit_n) is written; ONE
			// release point per lane, then the stamps of the children and the compacted surface entry ---------
			Publisher pub;
			if (wantFlec && sFlec < N.capacity) pub.store(&N.ray_meta[sFlec], metaFlec);
			if (wantFrac && sFrac < N.capacity) pub.store(&N.ray_meta[sFrac], metaFrac);
			if (surface) pub.store(&L.hit_list[hslot], i + 1u);
		}
	}
	flush_stats<STATS>(ws, st);
}

// ---- shading: light loop of every surface of every level ------------------------------------------

__global__ void __launch_bounds__(128) k_shade(SceneDev S, const FrameParams *__restrict__ Fp, LevelSet LS, const WaveState *__restrict__ ws, uint32_t resetHits)
{
	const FrameParams &F = *Fp;
	const uint32_t level = blockIdx.y;
	const LevelBuf &L = LS.l[level];
	const uint32_t n = ws->n_hit[level];
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x)
	{
		const uint32_t i = L.hit_list[h] - 1u;
		const float4 hp = L.hit_p[i], hn = L.hit_n[i], huv = L.hit_uv[i];
		const F3 P = f3(hp), Nn = f3(hn), rd = f3(L.ray_d[i]);
		const int mtl = __float_as_int(hn.w), tex = __float_as_int(huv.z);
		const F3 mA = f3(ldg4(&S.materials[4 * mtl])), mD = f3(ldg4(&S.materials[4 * mtl + 1])), mS = f3(ldg4(&S.materials[4 * mtl + 2]));
		const float shiness = ldg4(&S.materials[4 * mtl + 3]).x;
		const F3 vc = texel(S, tex, huv.x, huv.y);
		F3 mix_vd = f3(0, 0, 0), mix_vsc = f3(0, 0, 0);
		F3 mix_va = mixmul(mA, f3(F.env_light));
		for (uint32_t k = 0; k < F.n_lights; ++k)
		{
			const DevLight &lit = F.lights[k];
			if (!lit.enabled)
				continue;
			F3 p2l;
			float dis, lum;
			light_dir(lit, P, p2l, dis, lum);
			F3 la = f3(lit.ambient), ld = f3(lit.diffuse), ls = f3(lit.specular);
			if (lit.type == RT_LIGHT_POINT)
				la = la * lum, ld = ld * lum, ls = ls * lum;
			mix_va = mix_va + mixmul(mA, la);   // ambient is added before the shadow test
			if (L.shadow[(size_t)k * L.capacity + i])
				continue;
			float n_n = dot(Nn, p2l);
			if (n_n > 0)
				mix_vd = mix_vd + mixmul(mD, ld) * n_n;
			const F3 h2 = normalize(p2l - rd);
			n_n = dot(Nn, h2);
			if (n_n > 0)
				mix_vsc = mix_vsc + mixmul(mS, ls) * pow_ref(n_n, shiness);
		}
		const F3 c_all = mixmul(vc, mix_vd + mix_va) + mix_vsc;
		L.color[i] = make_float4(c_all.x, c_all.y, c_all.z, hp.w);
		if (resetHits)
			L.hit_list[h] = 0u;   // every entry is read exactly once, here: leave the list zeroed for the next frame (k_frame reads 0 as "not published")
	}
}

// ---- staged debug shaders (RayTracer::start types 1..6) -------------------------------------------
// RTcheck :48, RTdepth :81, RTnorm :94, RTtex :109, RTmtl :127, RTshd :222 of RayTracer.cpp: one
// closest-hit wave (plus one shadow wave for RTshd), then this kernel colours every pixel.
__device__ __forceinline__ float log_ref(float a) { return (float)log((double)a); }

__global__ void __launch_bounds__(128) k_debug(SceneDev S, const FrameParams *__restrict__ Fp, LevelBuf L, uint32_t n, uint8_t *__restrict__ out)
{
	const FrameParams &F = *Fp;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		int x, y;
		slot_to_pixel(F, i, x, y);
		F3 c = f3(0, 0, 0);
		if (F.type == RT_TYPE_CHECK)
		{
			const float v = ((y >> 6) & 1) == ((x >> 6) & 1) ? 1.0f : 0.0f;
			c = f3(v, v, v);
		}
		else
		{
			const float t = L.hit_p[i].w;
			if (F.type == RT_TYPE_DEPTH)
			{
				// Color::set, 3DElement.cpp:451-462
				if (t <= F.zNear) c = f3(1.0f, 0.0f, 0.0f);
				else if (t >= F.zFar) c = f3(0.0f, 0.0f, 0.0f);
				else
				{
					const float after = log_ref(t), mx = log_ref(F.zFar);
					const float g = (mx - after) / mx;
					c = f3(g, g, g);
				}
			}
			else if (t > F.zFar) c = f3(0.0f, 0.0f, 0.0f);
			else if (t < F.zNear) c = f3(1.0f, 1.0f, 1.0f);
			else
			{
				const float4 hn = L.hit_n[i], huv = L.hit_uv[i];
				const F3 Nn = f3(hn), P = f3(L.hit_p[i]), rd = f3(L.ray_d[i]);
				const int mtl = __float_as_int(hn.w), tex = __float_as_int(huv.z);
				if (F.type == RT_TYPE_NORMAL)   // Color(const Normal&): 0.5 * (n + 1) with a double multiply
					c = f3((float)(0.5 * (double)(Nn.x + 1)), (float)(0.5 * (double)(Nn.y + 1)), (float)(0.5 * (double)(Nn.z + 1)));
				else if (F.type == RT_TYPE_TEXTURE)
					c = tex >= 0 ? texel(S, tex, huv.x, huv.y) : f3(0.588f, 0.588f, 0.588f);
				else
				{
					const bool mtlStage = F.type == RT_TYPE_MATERIAL;
					const F3 mA = f3(ldg4(&S.materials[4 * mtl])), mD = f3(ldg4(&S.materials[4 * mtl + 1])), mS = f3(ldg4(&S.materials[4 * mtl + 2]));
					const float shiness = ldg4(&S.materials[4 * mtl + 3]).x;
					const F3 vc = texel(S, tex, huv.x, huv.y);
					F3 mix_vd = f3(0, 0, 0), mix_va = f3(0, 0, 0), mix_vsc = f3(0, 0, 0);
					for (uint32_t k = 0; k < F.n_lights; ++k)
					{
						const DevLight &lit = F.lights[k];
						if (!lit.enabled)
							continue;
						F3 p2l;
						float lum;
						// RTmtl tells point lights by position.alpha and sums the attenuation in another order
						if (mtlStage ? lit.position.w > RT_EPS6_BELOW : lit.type == RT_LIGHT_POINT)
						{
							const F3 v = f3(lit.position) - P;
							float dis = dot(v, v), step;
							if (mtlStage)
								step = (lit.attenuation.x + lit.attenuation.y * sqrtf(dis)) + lit.attenuation.z * dis;
							else
							{
								step = lit.attenuation.x + lit.attenuation.z * dis;
								dis = sqrtf(dis);
								step += lit.attenuation.y * dis;
							}
							lum = 1 / step;
							p2l = normalize(v);
						}
						else
						{
							lum = 1.0f;
							p2l = normalize(f3(lit.position));
						}
						const F3 la = f3(lit.ambient) * lum, ld = f3(lit.diffuse) * lum, ls = f3(lit.specular) * lum;
						mix_va = mix_va + mixmul(mA, la);
						if (!mtlStage && L.shadow[(size_t)k * L.capacity + i])
							continue;
						float n_n = dot(Nn, p2l);
						if (n_n > 0)
							mix_vd = mix_vd + mixmul(mD, ld) * n_n;
						const F3 h2 = normalize(p2l - rd);
						n_n = dot(Nn, h2);
						if (n_n > 0)
							mix_vsc = mix_vsc + mixmul(mS, ls) * pow_ref(n_n, shiness);
					}
					mix_va = mix_va + mixmul(mA, f3(F.env_light));   // environment ambient last in these stages
					c = mixmul(vc, mix_vd + mix_va) + mix_vsc;
				}
			}
		}
		uint8_t *o = out + ((size_t)y * F.width + x) * 3;
		o[0] = put8(c.x), o[1] = put8(c.y), o[2] = put8(c.z);
	}
}

// ---- B2: DrawObject::intersect of one object, reference order, no BVH ------------------------------
struct DevRay { float4 origin, direction; float mtlrfr; uint32_t type, is_inside, pad0; };
struct DevHit { float4 position, normal; float tu, tv; int material, texture; int id_object, id_sub, id_index, id_octant; float distance; float rfr; uint32_t is_inside, pad0; };

__global__ void k_intersect_object(SceneDev S, uint32_t primBegin, uint32_t primEnd, int modelIndex, const DevRay *rays, const DevHit *in,
	const uint32_t *skipIds, float minT, uint32_t *outIds, DevHit *out, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	RayD ray;
	ray.o = f3(rays[i].origin), ray.d = f3(rays[i].direction), ray.mtlrfr = rays[i].mtlrfr;
	ray.skip = skipIds[i], ray.type = (uint8_t)rays[i].type, ray.isInside = (uint8_t)rays[i].is_inside;
	const float hrDistance = in[i].distance;
	Best best = { hrDistance, RT_ID_NONE, ray.skip };
	bool done = false;
	// Sphere / Box / Plane, and BallPlane as its 16 slots in loop order (Basic3DObject.cpp:486-552)
	for (uint32_t p = primBegin; p < primEnd; ++p)
		test_prim<false>(S, ray, p, false, best, done);
	if (modelIndex >= 0)
	{
		// Model::intersect verbatim (Model.cpp:748-811): parts, enabled octants, triangles in file order
		const DevModel &M = S.models[modelIndex];
		const F3 idir = f3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
		float ans = border_test(ray.o, ray.d, idir, f3(M.border_min), f3(M.border_max));
		if (ans < hrDistance)
		{
			ans = hrDistance;
			bool stop = false;
			for (uint32_t a = M.part_begin; a < M.part_begin + M.part_count && !stop; ++a)
			{
				const DevPart &P = S.parts[a];
				uint32_t mask;
				if (!(border_test_ex(ray.o, ray.d, idir, f3(P.box_min), f3(P.box_max), &mask) < hrDistance))
					continue;
				for (uint32_t b = 0; b < 8 && !stop; ++b)
				{
					if (!(mask & (1u << b)))
						continue;
					for (uint32_t k = 0; k < P.tri_count && !stop; ++k)
					{
						const uint32_t tri = P.tri_begin + k, slot = S.tri_slot[tri];
						float4 g0, g1, g2;
						load_tri(S.tri_geom, slot, g0, g1, g2);
						if (!(__float_as_uint(g1.w) & (1u << b)))
							continue;   // not in this octant's list
						const uint32_t id = RT_ID_TRI | (b << 28) | tri;
						if (ray.skip == id)
							continue;
						const float t = triangle_t(ray.o, ray.d, f3(g0), f3(g1), f3(g2), nullptr);
						if (t < ans)
						{
							ans = t;
							best.t = t, best.id = best.newobj = id;
							if (t < minT)
								stop = true;
						}
					}
				}
			}
		}
	}
	outIds[i] = best.id;
	DevHit h = in[i];
	if (best.id != RT_ID_NONE && best.t < hrDistance)
	{
		const F3 P = ray.o + ray.d * best.t;
		const Surface sf = surface_attributes(S, ray, P, best.id);
		h.position = make_float4(P.x, P.y, P.z, 0), h.normal = make_float4(sf.N.x, sf.N.y, sf.N.z, 0);
		h.tu = sf.tu, h.tv = sf.tv, h.material = sf.mtl, h.texture = sf.tex;
		h.distance = best.t, h.rfr = sf.rfr, h.is_inside = sf.isInside;
	}
	out[i] = h;
}

void rtk_intersect_object(cudaStream_t st, const SceneDev &S, uint32_t primBegin, uint32_t primEnd, int modelIndex, const void *rays, const void *in,
	const uint32_t *skipIds, float minT, uint32_t *outIds, void *out, uint32_t n)
{
	if (n) k_intersect_object<<<(n + 127) / 128, 128, 0, st>>>(S, primBegin, primEnd, modelIndex, (const DevRay *)rays, (const DevHit *)in, skipIds, minT, outIds, (DevHit *)out, n);
}

// ---- end of frame: hit_list doubles as the publication flag of a surface (0 = not written) --------
__global__ void k_reset_hits(LevelSet LS, const WaveState *__restrict__ ws)
{
	const LevelBuf &L = LS.l[blockIdx.y];
	const uint32_t n = ws->n_hit[blockIdx.y] < L.capacity ? ws->n_hit[blockIdx.y] : L.capacity;
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x)
		L.hit_list[h] = 0u;
}

// ---- post-order combine --------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_combine(SceneDev S, const FrameParams *__restrict__ Fp, LevelBuf L, LevelBuf N,
	const WaveState *__restrict__ ws, uint32_t level, uint8_t *__restrict__ out)
{
	const FrameParams &F = *Fp;
	const uint32_t n = ws->count[level] < L.capacity ? ws->count[level] : L.capacity;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		float4 color = L.color[i];
		const int4 aux = L.aux[i];
		if (aux.z >= 0 && (aux.w & 3))
		{
			F3 c = f3(color);
			const float4 mP = ldg4(&S.materials[4 * aux.z + 3]);
			if (aux.w & 1)
			{
				c = c * (1 - mP.y);
				if (aux.x >= 0)
					c = c + f3(N.color[aux.x]) * mP.y;
			}
			if (aux.w & 2)
			{
				c = c * (1 - mP.z);
				if (aux.y >= 0)
				{
					const float4 cf = N.color[aux.y];
					F3 vcf = f3(1, 1, 1);
					if (aux.w & 4)
					{
						// Beer's law on the way out of the medium, RayTracer.cpp:585-590
						const F3 e = (f3(ldg4(&S.materials[4 * aux.z + 1])) * 0.15f) * (-cf.w);
						vcf = f3(exp_ref(e.x), exp_ref(e.y), exp_ref(e.z));
					}
					c = c + mixmul(f3(cf), vcf) * mP.z;
				}
			}
			color = make_float4(c.x, c.y, c.z, color.w);
			if (level > 0)
				L.color[i] = color;
		}
		if (level == 0)
		{
			uint8_t *o = pixel_of(F, i);
			o[0] = put8(color.x), o[1] = put8(color.y), o[2] = put8(color.z);
		}
	}
}

// The same combine as ONE launch: a thread per pixel walks its own ray tree (children are linked by
// aux.x / aux.y) depth first with an explicit stack and evaluates exactly the expression order of
// k_combine -- reflect child first, then the refract child -- so the colours are bit-identical to the
// level-by-level passes; nothing is written back to the levels.
struct ResolveFrame
{
	F3 c;
	float alpha, kr, kt;
	int4 aux;
};

__global__ void __launch_bounds__(256) k_resolve(SceneDev S, const FrameParams *__restrict__ Fp, LevelSet LS, const WaveState *__restrict__ ws, uint8_t *__restrict__ out)
{
	const FrameParams &F = *Fp;
	const uint32_t n = ws->count[0] < LS.l[0].capacity ? ws->count[0] : LS.l[0].capacity;
	for (uint32_t px = blockIdx.x * blockDim.x + threadIdx.x; px < n; px += gridDim.x * blockDim.x)
	{
		ResolveFrame st[RT_MAX_LEVELS + 1];
		int stage[RT_MAX_LEVELS + 1];
		int d = 0;
		uint32_t enter = px;      // slot to enter at depth d, or 0xFFFFFFFF when returning
		float4 ret = make_float4(0, 0, 0, 0);
		while (true)
		{
			if (enter != 0xFFFFFFFFu)
			{
				const LevelBuf &L = LS.l[d];
				const float4 color = L.color[enter];
				ResolveFrame &f = st[d];
				f.aux = L.aux[enter];
				f.c = f3(color), f.alpha = color.w;
				if (f.aux.z >= 0 && (f.aux.w & 3))
				{
					const float4 mP = ldg4(&S.materials[4 * f.aux.z + 3]);
					f.kr = mP.y, f.kt = mP.z;
					stage[d] = 0;
				}
				else
					stage[d] = 4;
				enter = 0xFFFFFFFFu;
			}
			ResolveFrame &f = st[d];
			int sg = stage[d];
			if (sg == 0)
			{
				sg = 2;
				if (f.aux.w & 1)
				{
					f.c = f.c * (1 - f.kr);
					if (f.aux.x >= 0) { stage[d] = 1; enter = (uint32_t)f.aux.x; ++d; continue; }
				}
			}
			if (sg == 1)
			{
				f.c = f.c + f3(ret) * f.kr;
				sg = 2;
			}
			if (sg == 2)
			{
				sg = 4;
				if (f.aux.w & 2)
				{
					f.c = f.c * (1 - f.kt);
					if (f.aux.y >= 0) { stage[d] = 3; enter = (uint32_t)f.aux.y; ++d; continue; }
				}
			}
			if (sg == 3)
			{
				F3 vcf = f3(1, 1, 1);
				if (f.aux.w & 4)
				{
					// Beer's law on the way out of the medium, RayTracer.cpp:585-590
					const F3 e = (f3(ldg4(&S.materials[4 * f.aux.z + 1])) * 0.15f) * (-ret.w);
					vcf = f3(exp_ref(e.x), exp_ref(e.y), exp_ref(e.z));
				}
				f.c = f.c + mixmul(f3(ret), vcf) * f.kt;
			}
			ret = make_float4(f.c.x, f.c.y, f.c.z, f.alpha);
			if (d == 0)
				break;
			--d;
		}
		uint8_t *o = pixel_of(F, px);
		o[0] = put8(ret.x), o[1] = put8(ret.y), o[2] = put8(ret.z);
	}
}

// ---- supersampling: integer mean of the sample frames (rt_render_supersampled) ------------------------
// out[b] = (sum over the n sample frames of frame_s[b]) / n for every byte of the `rows` image rows that start at byte
// offset rowStart[t] (one entry per row tile of the band): each sample was quantised by Color::put on its own, the mean
// truncates -- the arithmetic of averaging n reference renders in integer.  32-bit lanes of four bytes when the row
// pitch allows it.
struct AverageArgs
{
	const uint8_t *frames[RT_MAX_BATCH];
	uint8_t *out;
	uint32_t n, tiles, tileBytes;         // samples, row tiles of this band, bytes per row tile (tile_rows * 3 * width)
	uint32_t tileFirst, rank, world, serpentine, tileRows;
	size_t rowPitch;                      // 3 * width
};

__global__ void __launch_bounds__(256) k_average(AverageArgs A)
{
	const bool words = (A.tileBytes & 3u) == 0u && (A.rowPitch & 3u) == 0u;
	const uint32_t per = words ? A.tileBytes >> 2 : A.tileBytes;
	const uint64_t total = (uint64_t)per * A.tiles;
	for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total; w += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint32_t t = (uint32_t)(w / per), in = (uint32_t)(w % per);
		const size_t off = (size_t)shard_tile(A.tileFirst + t, A.rank, A.world, A.serpentine) * A.tileRows * A.rowPitch + (words ? (size_t)in * 4u : in);
		if (words)
		{
			uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
			for (uint32_t f = 0; f < A.n; ++f)
			{
				const uint32_t v = *(const uint32_t *)(A.frames[f] + off);
				s0 += v & 0xFFu, s1 += (v >> 8) & 0xFFu, s2 += (v >> 16) & 0xFFu, s3 += v >> 24;
			}
			*(uint32_t *)(A.out + off) = (s0 / A.n) | ((s1 / A.n) << 8) | ((s2 / A.n) << 16) | ((s3 / A.n) << 24);
		}
		else
		{
			uint32_t s0 = 0;
			for (uint32_t f = 0; f < A.n; ++f) s0 += A.frames[f][off];
			A.out[off] = (uint8_t)(s0 / A.n);
		}
	}
}

void rtk_average(cudaStream_t st, const uint8_t *const *frames, uint32_t n, uint8_t *out, int width, uint32_t tileRows, uint32_t tileFirst, uint32_t tiles,
	uint32_t rank, uint32_t world, uint32_t serpentine, unsigned sms)
{
	if (!tiles || !n) return;
	AverageArgs A;
	for (uint32_t f = 0; f < n; ++f) A.frames[f] = frames[f];
	A.out = out, A.n = n, A.tiles = tiles, A.rowPitch = (size_t)width * 3, A.tileBytes = (uint32_t)(tileRows * A.rowPitch);
	A.tileFirst = tileFirst, A.rank = rank, A.world = world, A.serpentine = serpentine, A.tileRows = tileRows;
	const uint64_t work = (uint64_t)(A.tileBytes / 4u + 1u) * tiles;
	k_average<<<(unsigned)std::min<uint64_t>((work + 255) / 256, (uint64_t)sms * 16), 256, 0, st>>>(A);
}

// ---- launchers -----------------------------------------------------------------------------------

static inline unsigned grid_for(uint32_t n, unsigned block, unsigned maxBlocks)
{
	unsigned g = (n + block - 1) / block;
	if (g < 1) g = 1;
	return g < maxBlocks ? g : maxBlocks;
}

void rtk_raygen(cudaStream_t st, const FrameParams *F, const LevelBuf &L, uint32_t n, unsigned sms)
{
	k_raygen<<<grid_for(n, 256, sms * 16), 256, 0, st>>>(F, L, n);
}

// Resident CTAs of 128 threads per SM for the traversal kernels: 8 (64 registers), 10 (48) or 12 (40).
// The kernels wait on memory latency rather than on issue slots, so more warps can pay for more spills.
static int traversal_ctas_per_sm()
{
	static const int v = []{ const char *e = getenv("RT_B200_OCC"); const int o = e ? atoi(e) : RT_CTAS_PER_SM; return (o == 4 || o == 6 || o == 10 || o == 12) ? o : RT_CTAS_PER_SM; }();
	return v;
}

// ---- coherence binning (see rt_kernels.h BinGrid) ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread3(uint32_t v)
{
	// 0b...cba -> 0b..c00b00a (up to 10 bits)
	v = (v | (v << 16)) & 0x030000FFu;
	v = (v | (v << 8)) & 0x0300F00Fu;
	v = (v | (v << 4)) & 0x030C30C3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}

// one atomic per distinct key of the warp: neighbouring rays mostly share their bin
__device__ __forceinline__ uint32_t warp_bin_add(uint32_t *hist, uint32_t key, bool valid)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t peers = __match_any_sync(0xffffffffu, valid ? key : 0xFFFFFFFFu);
	const uint32_t leader = __ffs((int)peers) - 1;
	uint32_t base = 0;
	if (valid && lane == leader) base = atomicAdd(&hist[key], (uint32_t)__popc(peers));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256) k_bin_keys(LevelBuf L, const WaveState *__restrict__ ws, uint32_t level, BinGrid G, uint32_t *hist)
{
	const uint32_t n = ws->count[level] < L.capacity ? ws->count[level] : L.capacity;
	const uint32_t top = (1u << G.bits) - 1u;
	for (uint32_t j0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; j0 < n; j0 += gridDim.x * blockDim.x)
	{
		const uint32_t j = j0 + (threadIdx.x & 31u);
		uint32_t key = 0;
		if (j < n)
		{
			const float4 o = L.ray_o[j], d = L.ray_d[j];
			const uint2 meta = L.ray_meta[j];
			const uint32_t qx = (uint32_t)fminf(fmaxf((o.x - G.lo[0]) * G.scale[0], 0.0f), (float)top);
			const uint32_t qy = (uint32_t)fminf(fmaxf((o.y - G.lo[1]) * G.scale[1], 0.0f), (float)top);
			const uint32_t qz = (uint32_t)fminf(fmaxf((o.z - G.lo[2]) * G.scale[2], 0.0f), (float)top);
			const uint32_t oct = (__float_as_uint(d.x) >> 31) | ((__float_as_uint(d.y) >> 31) << 1) | ((__float_as_uint(d.z) >> 31) << 2);
			uint32_t inside = (meta.y >> 8) & 1u;
			if (G.parts & 8u) inside = ((meta.y & 0xFFu) == (uint32_t)MY_RAY_REFRACTRAY_) ? 1u : 0u;
			const uint32_t cell = (spread3(qx) << 2) | (spread3(qy) << 1) | spread3(qz);
			key = ((((G.parts & 9u) ? inside << 3 : 0u) | ((G.parts & 2u) ? oct : 0u)) << (3u * G.bits)) | ((G.parts & 4u) ? cell : 0u);
			L.sort_key[j] = key;
		}
		warp_bin_add(hist, key, j < n);
	}
}

// exclusive prefix sum over the bins, one block (the histogram is L2-resident)
__global__ void __launch_bounds__(1024) k_bin_scan(uint32_t *hist, uint32_t total)
{
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

__global__ void __launch_bounds__(256) k_bin_scatter(LevelBuf L, const WaveState *__restrict__ ws, uint32_t level, uint32_t *hist)
{
	const uint32_t n = ws->count[level] < L.capacity ? ws->count[level] : L.capacity;
	for (uint32_t j0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; j0 < n; j0 += gridDim.x * blockDim.x)
	{
		const uint32_t j = j0 + (threadIdx.x & 31u);
		const uint32_t key = j < n ? L.sort_key[j] : 0u;
		const uint32_t pos = warp_bin_add(hist, key, j < n);
		if (j < n) L.order[pos] = j;
	}
}

void rtk_bin_rays(cudaStream_t st, const LevelBuf &L, const WaveState *ws, uint32_t level, const BinGrid &G, uint32_t *hist, unsigned sms)
{
	const uint32_t bins = 1u << (3u * G.bits + 4u);
	cudaMemsetAsync(hist, 0, sizeof(uint32_t) * bins, st);
	k_bin_keys<<<sms * 8, 256, 0, st>>>(L, ws, level, G, hist);
	k_bin_scan<<<1, 1024, 0, st>>>(hist, bins);
	k_bin_scatter<<<sms * 8, 256, 0, st>>>(L, ws, level, hist);
}

void rtk_wave(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelBuf &L, const LevelBuf &N, const LevelBuf &Lprev, WaveState *ws,
	uint32_t level, bool traceOn, bool shadowOn, float zNear, uint32_t maxItems, unsigned sms, bool stats, unsigned ctasPerSm, int walk)
{
	const int occ = traversal_ctas_per_sm();
	const unsigned g = grid_for(maxItems, RT_BLOCK, sms * (ctasPerSm && (int)ctasPerSm < occ ? ctasPerSm : occ));   // persistent: all CTAs resident
	if (walk == 2)
	{
		// two-stage walk: rays that enter the last Model's BVH are collected per warp, so its walk starts with full lanes
		if (stats) k_wave_split<true, RT_CTAS_PER_SM><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
		else if (occ == 6) k_wave_split<false, 6><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
		else k_wave_split<false, RT_CTAS_PER_SM><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
		return;
	}
	if (stats) k_wave<true, RT_CTAS_PER_SM, 0><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else if (walk == 3 && occ == 6) k_wave<false, 6, 1><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else if (walk == 3) k_wave<false, RT_CTAS_PER_SM, 1><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else if (walk == 4 && occ == 6) k_wave<false, 6, 2><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else if (walk == 4) k_wave<false, RT_CTAS_PER_SM, 2><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else if (occ == 4) k_wave<false, 4, 0><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else if (occ == 6) k_wave<false, 6, 0><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else if (occ == 10) k_wave<false, 10, 0><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else if (occ == 12) k_wave<false, 12, 0><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
	else k_wave<false, RT_CTAS_PER_SM, 0><<<g, RT_BLOCK, 0, st>>>(S, F, L, N, Lprev, ws, level, traceOn, shadowOn, zNear);
}

void rtk_frame(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelSet &LS, WaveState *ws, uint32_t nPix, unsigned sms, bool stats, unsigned ctasPerSm)
{
	// every CTA must be resident (consumers wait for producers): 8 CTAs of 128 threads fit per SM
	const int occ = traversal_ctas_per_sm();
	// a pipeline that shares the GPU with other frames in flight takes only its share of the resident CTA slots
	const unsigned g = grid_for(nPix, RT_BLOCK, sms * (ctasPerSm && (int)ctasPerSm < occ ? ctasPerSm : occ));
	if (stats) k_frame<true, RT_CTAS_PER_SM><<<g, RT_BLOCK, 0, st>>>(S, F, LS, ws);
	else if (occ == 4) k_frame<false, 4><<<g, RT_BLOCK, 0, st>>>(S, F, LS, ws);
	else if (occ == 6) k_frame<false, 6><<<g, RT_BLOCK, 0, st>>>(S, F, LS, ws);
	else if (occ == 10) k_frame<false, 10><<<g, RT_BLOCK, 0, st>>>(S, F, LS, ws);
	else if (occ == 12) k_frame<false, 12><<<g, RT_BLOCK, 0, st>>>(S, F, LS, ws);
	else k_frame<false, RT_CTAS_PER_SM><<<g, RT_BLOCK, 0, st>>>(S, F, LS, ws);
}

void rtk_shade(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelSet &LS, const WaveState *ws, uint32_t levels, uint32_t maxRays, unsigned sms, bool resetHits)
{
	const dim3 g(grid_for(maxRays, 128, sms * 8), levels);
	k_shade<<<g, 128, 0, st>>>(S, F, LS, ws, resetHits ? 1u : 0u);
}

void rtk_debug(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelBuf &L, uint32_t n, uint8_t *out, unsigned sms)
{
	k_debug<<<grid_for(n, 128, sms * 16), 128, 0, st>>>(S, F, L, n, out);
}

void rtk_reset_hits(cudaStream_t st, const LevelSet &LS, const WaveState *ws, uint32_t levels, uint32_t maxRays, unsigned sms)
{
	const dim3 g(grid_for(maxRays, 256, sms * 4), levels);
	k_reset_hits<<<g, 256, 0, st>>>(LS, ws);
}

void rtk_resolve(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelSet &LS, const WaveState *ws, uint8_t *out, uint32_t nPix, unsigned sms)
{
	k_resolve<<<grid_for(nPix, 256, sms * 32), 256, 0, st>>>(S, F, LS, ws, out);
}

void rtk_combine(cudaStream_t st, const SceneDev &S, const FrameParams *F, const LevelBuf &L, const LevelBuf &N, const WaveState *ws,
	uint32_t level, uint8_t *out, uint32_t maxRays, unsigned sms)
{
	k_combine<<<grid_for(maxRays, 256, sms * 32), 256, 0, st>>>(S, F, L, N, ws, level, out);
}
