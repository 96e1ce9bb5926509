// Headless render driver: the replacement for the reference's GLUT shell (main.cpp) on the
// one path this repo accelerates.  It builds a named synthetic scene through the object-model
// API, calls RayTracer::start() exactly as main.cpp:301 does, polls isFinish (main.cpp:252) and
// reads RayTracer::output (main.cpp:207-208).
//
// The SAME source is compiled twice:
//   * against raytrace_b200/host/ (the product)      -> raytrace_b200/bin/rt_render
//   * against /root/reference with -DRT_ARM_REFERENCE -> oracle/_ref/ref_render (test infra)
// so both arms are driven by identical code.  The arm-specific "taps" (primary hit ids, ray
// counters) come from render_taps.h, resolved by the include path of each build.
//
//   render --scene c1 --width 1088 --height 576 --level 1 [--type 0x80] [--threads 8]
//          [--n N] [--parts P] [--repeat K] [--out f.rgb] [--ids f.bin] [--counts]
//          [--tiles K --seed S | --tiles K --stratified] [--orbit K [--orbit-k k]] [--tmpdir D] [--gpus N]
// prints one JSON line on stdout.
#include "Scene.h"
#include "RayTracer.h"
#include "render_taps.h"
#include "../raytrace_b200/scenes/scenes.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static Scene scene;
static RayTracer rayt(scene);

static uint64_t fnv1a64(const uint8_t *p, size_t n)
{
	uint64_t h = 1469598103934665603ULL;
	for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
	return h;
}

static double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
	rtscenes::SceneArgs sa;
	int width = 1088, height = 576, level = 1, type = MY_MODEL_RAYTRACE, threads = 8, repeat = 1;
	int tiles = 0, seed = 0, warmup = 0, orbit = 0, orbitK = 0;
	bool counts = false;
	std::string out, ids;
	for (int i = 1; i < argc; ++i)
	{
		std::string k = argv[i];
		auto val = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", k.c_str()); exit(2); } return argv[++i]; };
		if (k == "--scene") sa.name = val();
		else if (k == "--n") sa.n = atoi(val());
		else if (k == "--parts") sa.parts = atoi(val());
		else if (k == "--tmpdir") sa.tmpdir = val();
		else if (k == "--width") width = atoi(val());
		else if (k == "--height") height = atoi(val());
		else if (k == "--level") level = atoi(val());
		else if (k == "--type") type = (int)strtol(val(), nullptr, 0);
		else if (k == "--threads") threads = atoi(val());
		else if (k == "--repeat") repeat = atoi(val());
		else if (k == "--out") out = val();
		else if (k == "--ids") ids = val();
		else if (k == "--counts") counts = true;
		else if (k == "--tiles") tiles = atoi(val());
		else if (k == "--seed") seed = atoi(val());
		else if (k == "--stratified") seed = -1;   // tiles spread evenly over the frame instead of seeded random ones
		else if (k == "--warmup") warmup = atoi(val());
		else if (k == "--orbit") orbit = atoi(val());       // pass r renders camera r % orbit of the orbit (scenes.h orbit_camera)
		else if (k == "--orbit-k") orbitK = atoi(val());    // first camera of the orbit to render
		else if (k == "--gpus") rt_taps::set_gpus(atoi(val()));
		else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
	}
	scene.cam.resize(width, height);
	if (!rtscenes::build(scene, sa)) { fprintf(stderr, "unknown scene %s\n", sa.name.c_str()); return 2; }
	rayt.maxLevel = (uint8_t)level;

	// frames larger than the reference's fixed 2048x2048x3 buffer: swap the public pointer
	// (SURVEY.md 8c "harness capabilities"); start() memsets 2048*2048*3 bytes regardless.
	const size_t need = (size_t)width * height * 3, fixed = (size_t)2048 * 2048 * 3;
	if (need > fixed) rt_taps::ensure_output(rayt, need);

	std::vector<double> walls, uses;
	rt_taps::Counts cnt;
	const Camera baseCam = scene.cam;
	if (tiles > 0)
	{
		// bounded CPU sample: `tiles` seeded 64x64 tiles through the per-pixel entry, `warmup`
		// untimed + `repeat` timed passes over the same tiles; rays counted in one extra pass
		// step_s = trace seconds of the sample, prepare_s = the reference's per-frame RTPrepare (timed apart: a
		// frame pays it once, a sample must not be charged all of it -- see render_taps.h)
		long px = 0;
		std::vector<double> prepares;
		std::vector<unsigned long long> raysPer;
		unsigned long long rays = 0;
		for (int r = 0; r < warmup + repeat; ++r)
		{
			double prep = 0, trace = 0;
			// timed pass s of `repeat` looks through camera floor(s * orbit / repeat): the passes are spread evenly over the
			// whole orbit, whatever their number (the cameras of an orbit differ in cost by up to 2x)
			if (orbit > 0) scene.cam = rtscenes::orbit_camera(baseCam, orbitK + (r >= warmup ? (int)(((long long)(r - warmup) * orbit) / repeat) : r), orbit);
			// with --counts every pass is counted (a relaxed atomic add per ray and object walk: noise next to a ray's
			// tens of microseconds), so rays and seconds belong to the same cameras
			rt_taps::Counts pc;
			px = rt_taps::render_tiles(scene, rayt, width, height, tiles, seed, threads, type, counts ? &pc : nullptr, &prep, &trace);
			if (r >= warmup)
			{
				walls.push_back(trace), prepares.push_back(prep);
				raysPer.push_back(pc.primary + pc.shadow + pc.reflect + pc.refract);
				rays += raysPer.back();
				cnt = pc;
			}
		}
		if (!raysPer.empty()) rays /= raysPer.size();   // rays_per_step = mean over the timed passes
		printf("{\"scene\":\"%s\",\"arm\":\"%s\",\"w\":%d,\"h\":%d,\"level\":%d,\"threads\":%d,\"tiles\":%d,\"pixels\":%ld,\"rays_per_step\":%llu,\"hash\":\"%016llx\",\"step_s\":[",
			sa.name.c_str(), rt_taps::arm(), width, height, level, threads, tiles, px, rays, (unsigned long long)fnv1a64(rayt.output, need));
		for (size_t i = 0; i < walls.size(); ++i) printf("%s%.6f", i ? "," : "", walls[i]);
		printf("],\"prepare_s\":[");
		for (size_t i = 0; i < prepares.size(); ++i) printf("%s%.6f", i ? "," : "", prepares[i]);
		printf("],\"rays_s\":[");
		for (size_t i = 0; i < raysPer.size(); ++i) printf("%s%llu", i ? "," : "", raysPer[i]);
		printf("],\"frame_tiles\":%d,\"stratified\":%s}\n", (width / 64) * (height / 64), seed < 0 ? "true" : "false");
		return 0;
	}
	for (int r = 0; r < repeat; ++r)
	{
		if (orbit > 0) scene.cam = rtscenes::orbit_camera(baseCam, orbitK + r, orbit);
		double t0 = now_s();
		rayt.start((uint8_t)type, (int8_t)threads);
		while (!rayt.isFinish) std::this_thread::sleep_for(std::chrono::microseconds(200));
		double t1 = now_s();
		walls.push_back(t1 - t0);
		uses.push_back((double)rayt.useTime);
	}
	const uint64_t h = fnv1a64(rayt.output, need);
	if (!out.empty())
	{
		FILE *f = fopen(out.c_str(), "wb");
		if (!f) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 2; }
		fwrite(rayt.output, 1, need, f);
		fclose(f);
	}
	if (!ids.empty())
	{
		std::vector<rt_taps::HitId> hid((size_t)width * height);
		rt_taps::primary_ids(scene, rayt, width, height, hid.data());
		FILE *f = fopen(ids.c_str(), "wb");
		if (!f) { fprintf(stderr, "cannot write %s\n", ids.c_str()); return 2; }
		fwrite(hid.data(), sizeof(rt_taps::HitId), hid.size(), f);
		fclose(f);
	}
	if (counts) cnt = rt_taps::count_rays(scene, rayt, type, threads);

	printf("{\"scene\":\"%s\",\"arm\":\"%s\",\"w\":%d,\"h\":%d,\"level\":%d,\"type\":%d,\"threads\":%d,\"hash\":\"%016llx\",\"wall_s\":[",
		sa.name.c_str(), rt_taps::arm(), width, height, level, type, threads, (unsigned long long)h);
	for (size_t i = 0; i < walls.size(); ++i) printf("%s%.6f", i ? "," : "", walls[i]);
	printf("],\"useTime_s\":[");
	for (size_t i = 0; i < uses.size(); ++i) printf("%s%.6f", i ? "," : "", uses[i]);
	printf("]");
	if (counts)
		printf(",\"rays\":{\"primary\":%llu,\"shadow\":%llu,\"reflect\":%llu,\"refract\":%llu,\"total\":%llu}",
			cnt.primary, cnt.shadow, cnt.reflect, cnt.refract, cnt.primary + cnt.shadow + cnt.reflect + cnt.refract);
	printf("%s}\n", rt_taps::extra_json(rayt).c_str());
	return 0;
}
