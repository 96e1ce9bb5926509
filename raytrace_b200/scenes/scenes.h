// Synthetic scene builders shared by BOTH arms of every parity test.
//
// This header is written only against the reference's public object-model API
// (Scene::AddLight/AddPlane/AddSphere/AddCube/AddBallPlane/AddModel/MovePos/ChgMtl,
// Scene::{EnvLight,Lights,Objects,MtlLiby,cam}; /root/reference/Scene.h:24-45), so the same
// source compiles against
//   * the reference's own headers  -> oracle/_ref/ref_render   (oracle/build_ref.sh), and
//   * raytrace_b200/host/ headers  -> raytrace_b200/bin/rt_render (the B200 product, also via host/capi.cpp).
// That both compile from one file is the drop-in check for SURVEY.md section 8 (b) B1.
//
// Scene ids follow SURVEY.md section 8 (d): c1..c5 plus small "t_*" cases for tests.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace rtscenes
{

struct SceneArgs
{
	std::string name = "c1";
	int n = 0;            // primary size knob (spheres per side, mesh cells per side); 0 = scene default
	int parts = 0;        // mesh parts per side; 0 = scene default
	std::string tmpdir = "/tmp";
};

inline std::wstring widen(const std::string &s) { return std::wstring(s.begin(), s.end()); }

// main.cpp:165-168 of the reference: the two default lights.
template<class SceneT> inline void default_lights(SceneT &scene)
{
	scene.EnvLight = Vertex(0.05f, 0.05f, 0.05f, 1.0f);
	auto l = scene.AddLight(MY_LIGHT_PARALLEL, Vertex(0.1f, 0.45f, 0.45f));
	scene.MovePos(MY_MODEL_LIGHT, l, Vertex(-45, 45, 0));
	scene.AddLight(MY_LIGHT_POINT, Vertex(0.15f, 0.55f, 0.3f), Vertex(0.0f, 0.0f, 1.0f, 256));
}

// Height field y = 1.5 + 0.6 sin3x cos2.5z + 0.25 sin(9x+1) sin7z on [-4,4]^2, `cells` x `cells`
// quads split in two, `pside` x `pside` usemtl parts, analytic normals, written as OBJ + MTL.
// 24-bit BMP (bottom-up, BGR, width a multiple of 4 so rows carry no padding): what Model::loadtex reads (Model.cpp:283-315)
inline void write_bmp(const std::string &path, int w, int h)
{
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
	const unsigned size = (unsigned)(w * h * 3);
	unsigned char head[54] = { 'B', 'M' };
	auto le32 = [&](int off, unsigned v) { head[off] = v & 255, head[off + 1] = (v >> 8) & 255, head[off + 2] = (v >> 16) & 255, head[off + 3] = (v >> 24) & 255; };
	le32(2, 54 + size), le32(10, 54), le32(14, 40), le32(18, (unsigned)w), le32(22, (unsigned)h);
	head[26] = 1, head[28] = 24;
	le32(34, size);
	fwrite(head, 1, 54, f);
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x)
		{
			// coloured checker with a gradient: every texel differs from its neighbours
			const int c = ((x / 2 + y / 2) & 1) ? 200 : 40;
			const unsigned char px[3] = { (unsigned char)((c + 7 * x) & 255), (unsigned char)((255 - c + 5 * y) & 255), (unsigned char)((90 + 11 * x + 13 * y) & 255) };
			fwrite(px, 1, 3, f);
		}
	fclose(f);
}

inline void write_heightfield(const std::string &obj, const std::string &mtl, int cells, int pside, const std::string &texture = std::string())
{
	FILE *fm = fopen(mtl.c_str(), "w");
	if (!fm) { fprintf(stderr, "cannot write %s\n", mtl.c_str()); exit(2); }
	fprintf(fm, "newmtl hfa\nKa 0.100000 0.100000 0.100000\nKd 0.100000 0.500000 0.800000\nKs 1.000000 1.000000 1.000000\nNs 100.000000\n");
	if (!texture.empty()) fprintf(fm, "map_Kd %s\n", texture.c_str());   // -> Model::loadtex of <texture without extension>.bmp (Model.cpp:260-278)
	fprintf(fm, "newmtl hfb\nKa 0.200000 0.100000 0.100000\nKd 0.800000 0.400000 0.100000\nKs 0.500000 0.500000 0.500000\nNs 20.000000\n");
	fclose(fm);
	FILE *fo = fopen(obj.c_str(), "w");
	if (!fo) { fprintf(stderr, "cannot write %s\n", obj.c_str()); exit(2); }
	const int nv = cells + 1;
	for (int j = 0; j < nv; ++j)
		for (int i = 0; i < nv; ++i)
		{
			const double x = -4.0 + 8.0 * i / cells, z = -4.0 + 8.0 * j / cells;
			const double y = 1.5 + 0.6 * sin(3 * x) * cos(2.5 * z) + 0.25 * sin(9 * x + 1) * sin(7 * z);
			fprintf(fo, "v %.6f %.6f %.6f\n", x, y, z);
		}
	for (int j = 0; j < nv; ++j)
		for (int i = 0; i < nv; ++i)
		{
			const double x = -4.0 + 8.0 * i / cells, z = -4.0 + 8.0 * j / cells;
			const double dx = 1.8 * cos(3 * x) * cos(2.5 * z) + 2.25 * cos(9 * x + 1) * sin(7 * z);
			const double dz = -1.5 * sin(3 * x) * sin(2.5 * z) + 1.75 * sin(9 * x + 1) * cos(7 * z);
			const double inv = 1.0 / sqrt(dx * dx + 1.0 + dz * dz);
			fprintf(fo, "vn %.6f %.6f %.6f\n", -dx * inv, inv, -dz * inv);
		}
	for (int j = 0; j < nv; ++j)
		for (int i = 0; i < nv; ++i)
			fprintf(fo, "vt %.6f %.6f\n", (double)i / cells, (double)j / cells);
	const int pc = cells / pside;   // cells per part side
	for (int pj = 0; pj < pside; ++pj)
		for (int pi = 0; pi < pside; ++pi)
		{
			fprintf(fo, "usemtl %s\n", ((pi + pj) & 1) ? "hfb" : "hfa");
			for (int j = pj * pc; j < (pj + 1) * pc; ++j)
				for (int i = pi * pc; i < (pi + 1) * pc; ++i)
				{
					const int a = j * nv + i + 1, b = a + 1, c = a + nv, d = c + 1;
					fprintf(fo, "f %d/%d/%d %d/%d/%d %d/%d/%d\n", a, a, a, c, c, c, b, b, b);
					fprintf(fo, "f %d/%d/%d %d/%d/%d %d/%d/%d\n", b, b, b, c, c, c, d, d, d);
				}
		}
	fclose(fo);
}

// A closed, coarse icosphere-like blob written with quads and "v//vn" faces, to exercise the
// loader's quad split (Model.cpp:88-117) and the no-tcoord parse path (Model.cpp:903-906).
inline void write_quadblob(const std::string &obj, int rings, int sectors)
{
	FILE *fo = fopen(obj.c_str(), "w");
	if (!fo) { fprintf(stderr, "cannot write %s\n", obj.c_str()); exit(2); }
	for (int r = 0; r <= rings; ++r)
		for (int s = 0; s < sectors; ++s)
		{
			const double th = 3.14159265358979323846 * (0.08 + 0.84 * r / rings), ph = 2 * 3.14159265358979323846 * s / sectors;
			const double rad = 1.0 + 0.25 * sin(3 * ph) * sin(2 * th);
			fprintf(fo, "v %.6f %.6f %.6f\n", rad * sin(th) * cos(ph), rad * cos(th), rad * sin(th) * sin(ph));
		}
	for (int r = 0; r <= rings; ++r)
		for (int s = 0; s < sectors; ++s)
		{
			const double th = 3.14159265358979323846 * (0.08 + 0.84 * r / rings), ph = 2 * 3.14159265358979323846 * s / sectors;
			fprintf(fo, "vn %.6f %.6f %.6f\n", sin(th) * cos(ph), cos(th), sin(th) * sin(ph));
		}
	for (int r = 0; r < rings; ++r)
	{
		if (r == rings / 2) fprintf(fo, "usemtl lower\n");
		for (int s = 0; s < sectors; ++s)
		{
			const int a = r * sectors + s + 1, b = r * sectors + (s + 1) % sectors + 1, c = b + sectors, d = a + sectors;
			fprintf(fo, "f %d//%d %d//%d %d//%d %d//%d\n", a, a, b, b, c, c, d, d);
		}
	}
	fclose(fo);
}

template<class SceneT> inline int add_heightfield(SceneT &scene, const SceneArgs &a, int cells, int pside, bool textured = false)
{
	char tag[64];
	snprintf(tag, sizeof tag, "/rt_hf_%d_%d%s", cells, pside, textured ? "_tex" : "");
	const std::string obj = a.tmpdir + tag + ".obj", mtl = a.tmpdir + tag + ".mtl", tex = a.tmpdir + "/rt_tex_16x12.bmp";
	if (textured) write_bmp(tex, 16, 12);
	write_heightfield(obj, mtl, cells, pside, textured ? tex : std::string());
	return scene.AddModel(widen(obj), widen(mtl));
}

// ---- C1: exactly main.cpp:165-174 ------------------------------------------------------
template<class SceneT> inline void build_c1(SceneT &scene, const SceneArgs &)
{
	default_lights(scene);
	scene.AddPlane();
	scene.AddSphere(1.0);
}

// ---- C2: plane + n x n spheres r=0.4, 4 point lights -----------------------
template<class SceneT> inline void build_c2(SceneT &scene, const SceneArgs &a)
{
	const int n = a.n > 0 ? a.n : 32;
	scene.EnvLight = Vertex(0.05f, 0.05f, 0.05f, 1.0f);
	const float lx[4] = { 12, -12, 12, -12 }, lz[4] = { 6, 6, -26, -26 };
	for (int k = 0; k < 4; ++k)
	{
		scene.AddLight(MY_LIGHT_POINT, Vertex(0.15f, 0.55f, 0.3f), Vertex(0.0f, 0.0f, 1.0f, 256));
		scene.Lights[k].position = Vertex(lx[k], 10, lz[k], 1.0f);
	}
	scene.AddPlane();
	for (int j = 0; j < n; ++j)
		for (int i = 0; i < n; ++i)
		{
			scene.AddSphere(0.4f);
			scene.Objects.back()->position = Vertex(-(n - 1) * 0.5f + i, 0.4f, 4.0f - j);
		}
}

// ---- C3: plane + height-field Model (n x n cells, parts x parts usemtl groups) ---------
template<class SceneT> inline void build_c3(SceneT &scene, const SceneArgs &a)
{
	const int cells = a.n > 0 ? a.n : 720, pside = a.parts > 0 ? a.parts : cells / 16;
	default_lights(scene);
	scene.Lights[1].position = Vertex(3, 9, 12, 1.0f);   // lifted off the ground plane
	scene.AddPlane();
	const int m = add_heightfield(scene, a, cells, pside);
	scene.ChgMtl(m, scene.MtlLiby[1]);
	scene.MovePos(MY_MODEL_OBJECT, m, Vertex(0, 0, 5));
}

// ---- C4/C5: mesh + 8x8 glass spheres + mirror spheres (refraction) ---------------------
template<class SceneT> inline void build_c4(SceneT &scene, const SceneArgs &a)
{
	const int cells = a.n > 0 ? a.n : 1440, pside = a.parts > 0 ? a.parts : cells / 16;
	default_lights(scene);
	scene.Lights[1].position = Vertex(3, 9, 12, 1.0f);
	scene.AddPlane();
	const int m = add_heightfield(scene, a, cells, pside);
	scene.ChgMtl(m, scene.MtlLiby[1]);
	scene.MovePos(MY_MODEL_OBJECT, m, Vertex(0, 0, 3));
	for (int j = 0; j < 8; ++j)
		for (int i = 0; i < 8; ++i)
		{
			const int s = scene.AddSphere(0.35f);
			scene.Objects.back()->position = Vertex(-3.5f + i, 2.9f + 0.25f * ((i + j) & 1), 7.5f - 0.9f * j);
			scene.ChgMtl(s, scene.MtlLiby[4]);   // "grass" = glass: reflect .15, refract .75, rfr 1.5
		}
	for (int i = 0; i < 6; ++i)
	{
		const int s = scene.AddSphere(0.8f);
		scene.Objects.back()->position = Vertex(-7.5f + 3.0f * i, 0.8f, 9.5f + ((i & 1) ? 1.0f : 0.0f));
		scene.ChgMtl(s, scene.MtlLiby[2]);       // mirror
	}
}

// ---- small test scenes -----------------------------------------------------------------
// every analytic primitive kind + glass + mirror + box, two default lights
template<class SceneT> inline void build_t_mixed(SceneT &scene, const SceneArgs &)
{
	default_lights(scene);
	scene.Lights[1].position = Vertex(-2, 7, 9, 1.0f);
	scene.AddPlane();
	int s = scene.AddSphere(1.0f);                       // default blue sphere
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(-2.5f, 0, 1));
	s = scene.AddSphere(1.2f);                           // glass
	scene.ChgMtl(s, scene.MtlLiby[4]);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(0.5f, 0.2f, 4));
	s = scene.AddSphere(0.9f);                           // mirror
	scene.ChgMtl(s, scene.MtlLiby[2]);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(3.2f, 0, 0.5f));
	int b = scene.AddCube(1.6f);                         // brass box
	scene.MovePos(MY_MODEL_OBJECT, b, Vertex(-0.5f, 0, -2.5f));
	b = scene.AddCube(1.0f);
	scene.ChgMtl(b, scene.MtlLiby[5]);                   // wall box
	scene.MovePos(MY_MODEL_OBJECT, b, Vertex(2.2f, 0.0f, 5.5f));
	s = scene.AddSphere(0.5f);                           // green reflective
	scene.ChgMtl(s, scene.MtlLiby[3]);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(-1.2f, 0, 6.5f));
}

// BallPlane lattice + plane + glass sphere
template<class SceneT> inline void build_t_ballplane(SceneT &scene, const SceneArgs &)
{
	default_lights(scene);
	scene.Lights[1].position = Vertex(2, 8, 10, 1.0f);
	scene.AddPlane();
	const int bp = scene.AddBallPlane(0.3f);
	scene.Objects[bp]->position = Vertex(0, 1.2f, 3);
	const int s = scene.AddSphere(0.8f);
	scene.ChgMtl(s, scene.MtlLiby[4]);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(2.5f, 0, 6));
}

// small mesh + spheres + plane (c3/c4 shape at test size); n = cells, parts = parts per side
template<class SceneT> inline void build_t_mesh(SceneT &scene, const SceneArgs &a)
{
	SceneArgs b = a;
	if (b.n <= 0) b.n = 48;
	if (b.parts <= 0) b.parts = 3;
	default_lights(scene);
	scene.Lights[1].position = Vertex(3, 9, 12, 1.0f);
	scene.AddPlane();
	int s = scene.AddSphere(0.7f);
	scene.ChgMtl(s, scene.MtlLiby[4]);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(-2.0f, 2.6f, 7));
	const int m = add_heightfield(scene, b, b.n, b.parts);
	scene.MovePos(MY_MODEL_OBJECT, m, Vertex(0, 0, 5));
	s = scene.AddSphere(0.9f);
	scene.ChgMtl(s, scene.MtlLiby[2]);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(3.0f, 1.9f, 6));
}

// two meshes (one reflective, one with MTL materials and quads), box between them
template<class SceneT> inline void build_t_twomesh(SceneT &scene, const SceneArgs &a)
{
	SceneArgs b = a;
	if (b.n <= 0) b.n = 32;
	if (b.parts <= 0) b.parts = 2;
	default_lights(scene);
	scene.Lights[1].position = Vertex(-3, 9, 12, 1.0f);
	scene.AddPlane();
	const int m = add_heightfield(scene, b, b.n, b.parts);
	scene.ChgMtl(m, scene.MtlLiby[3]);
	scene.MovePos(MY_MODEL_OBJECT, m, Vertex(-1, -0.4f, 2));
	const int c = scene.AddCube(1.2f);
	scene.MovePos(MY_MODEL_OBJECT, c, Vertex(3.5f, 0, 7));
	const std::string blob = b.tmpdir + "/rt_blob.obj";
	write_quadblob(blob, 10, 16);
	const int q = scene.AddModel(widen(blob), widen(b.tmpdir + "/rt_blob_missing.mtl"));
	scene.MovePos(MY_MODEL_OBJECT, q, Vertex(-6.0f, 4.5f, -4.0f));
}

// f-2: a mesh whose MTL names a texture (map_Kd -> 24-bit BMP through Model::loadtex, per-part texture on the device) next
// to a model turned by Model::zRotate (Model.cpp:351-393: y/z swap of vertices, normals, part boxes -- and a min/max z that
// is NOT re-ordered), a mirror sphere to see both again in a reflection
template<class SceneT> inline void build_t_textured(SceneT &scene, const SceneArgs &a)
{
	SceneArgs b = a;
	if (b.n <= 0) b.n = 32;
	if (b.parts <= 0) b.parts = 2;
	default_lights(scene);
	scene.Lights[1].position = Vertex(2, 9, 13, 1.0f);
	scene.AddPlane();
	const int m = add_heightfield(scene, b, b.n, b.parts, true);
	scene.MovePos(MY_MODEL_OBJECT, m, Vertex(-1.5f, 0, 4));
	const std::string blob = b.tmpdir + "/rt_blob.obj";
	write_quadblob(blob, 10, 16);
	const int q = scene.AddModel(widen(blob), widen(b.tmpdir + "/rt_blob_missing.mtl"));
	dynamic_cast<Model &>(*scene.Objects[q]).zRotate();
	scene.MovePos(MY_MODEL_OBJECT, q, Vertex(4.5f, 3.0f, 6.0f));
	const int s = scene.AddSphere(1.0f);
	scene.ChgMtl(s, scene.MtlLiby[2]);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(1.5f, 2.4f, 8.5f));
}

// ---- edge cases ------------------------------------------------------------------------------------
// nothing to hit: every ray misses (black frame, 1e20 distances)
template<class SceneT> inline void build_t_empty(SceneT &scene, const SceneArgs &)
{
	default_lights(scene);
}

// no lights at all: only the environment ambient term survives (RayTracer.cpp:472)
template<class SceneT> inline void build_t_nolight(SceneT &scene, const SceneArgs &)
{
	scene.EnvLight = Vertex(0.6f, 0.5f, 0.4f, 1.0f);
	scene.AddPlane();
	const int s = scene.AddSphere(1.3f);
	scene.ChgMtl(s, scene.MtlLiby[0]);
}

// 8 lights (the maximum, Scene.cpp:85) of all three kinds, one switched off, a hidden object, a
// spot light (shaded as a parallel light by the reference, RayTracer.cpp:496-503), overlapping and
// ground-piercing spheres, the camera moved and turned
template<class SceneT> inline void build_t_lights(SceneT &scene, const SceneArgs &)
{
	scene.EnvLight = Vertex(0.05f, 0.05f, 0.05f, 1.0f);
	for (int k = 0; k < 8; ++k)
	{
		const uint8_t type = k % 3 == 0 ? MY_LIGHT_PARALLEL : (k % 3 == 1 ? MY_LIGHT_POINT : MY_LIGHT_SPOT);
		scene.AddLight(type, Vertex(0.1f + 0.02f * k, 0.5f, 0.3f), type == MY_LIGHT_PARALLEL ? Vertex(1, 0, 0, 0.25f) : Vertex(0.2f, 0.05f, 0.6f, 40));
		scene.MovePos(MY_MODEL_LIGHT, k, Vertex(-20.0f + 9.0f * k, 33.0f * k, -2.0f * k));
	}
	scene.AddLight(MY_LIGHT_POINT, Vertex(1, 1, 1));           // ninth light: rejected (returns 0xff)
	scene.Switch(MY_MODEL_LIGHT, 3, false);
	scene.AddPlane();
	int s = scene.AddSphere(1.0f);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(-1.0f, -0.6f, 2));   // pierces the ground plane
	s = scene.AddSphere(1.0f);
	scene.ChgMtl(s, scene.MtlLiby[4]);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(0.2f, 0.1f, 2.5f));  // glass, overlaps the first sphere
	s = scene.AddSphere(0.8f);
	scene.MovePos(MY_MODEL_OBJECT, s, Vertex(2.5f, 0, 4));
	scene.Switch(MY_MODEL_OBJECT, s, false);                      // hidden: must not be uploaded
	const int b = scene.AddCube(1.5f);
	scene.ChgMtl(b, scene.MtlLiby[2]);
	scene.MovePos(MY_MODEL_OBJECT, b, Vertex(3.0f, 0, 1));
	scene.cam.move(1.5f, 0.5f, 3.0f);
	scene.cam.yaw(12.0f);
	scene.cam.pitch(-8.0f);
}

// camera inside a glass sphere, looking out through it at a mesh
template<class SceneT> inline void build_t_inside(SceneT &scene, const SceneArgs &a)
{
	SceneArgs b = a;
	if (b.n <= 0) b.n = 24;
	if (b.parts <= 0) b.parts = 2;
	default_lights(scene);
	scene.Lights[1].position = Vertex(4, 9, 14, 1.0f);
	scene.AddPlane();
	const int s = scene.AddSphere(2.0f);
	scene.ChgMtl(s, scene.MtlLiby[4]);
	scene.Objects[s]->position = Vertex(0, 4, 15);                // around the default camera
	const int m = add_heightfield(scene, b, b.n, b.parts);
	scene.ChgMtl(m, scene.MtlLiby[3]);
	scene.MovePos(MY_MODEL_OBJECT, m, Vertex(0, 0, 4));
}

// ---- camera orbit: the frame stream of the throughput benchmark ----------------------------------------
// Camera k of K: the scene's own camera dollied along a small closed path -- up to 1 unit to either side, 0.5 up and
// 1 forward, so the frames stay close to the configuration's own view and cost -- with Camera::move only
// (translations along u, v, n, 3DElement.cpp:585-590: both arms compute it with the same float additions).
// Camera 0 is exactly `base`: its frame is the configuration's own frame (the one the golden hashes pin).
template<class CameraT> inline CameraT orbit_camera(const CameraT &base, int k, int K)
{
	CameraT c = base;
	if (K > 1 && k % K != 0)
	{
		const double a = 2.0 * 3.14159265358979323846 * (k % K) / K;
		c.move((float)(1.0 * sin(a)), (float)(0.25 * (1.0 - cos(a))), (float)(0.5 * (1.0 - cos(a))));
	}
	return c;
}

// ---- jittered supersampling (BASELINE configs[4]) ----------------------------------------------------------
// side x side stratified sub-pixel offsets in [0,1)^2, fixed by `seed` (a 64-bit LCG: the table is the same in every
// arm, language and run).  out: side*side (dx, dy) pairs.
inline void stratified_table(int side, int seed, std::vector<float> &out)
{
	unsigned long long s = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)seed;
	out.clear();
	for (int j = 0; j < side; ++j)
		for (int i = 0; i < side; ++i)
		{
			s = s * 6364136223846793005ULL + 1442695040888963407ULL;
			const double a = (double)((s >> 40) & 0xFFFFFFULL) / (double)(1 << 24);
			s = s * 6364136223846793005ULL + 1442695040888963407ULL;
			const double b = (double)((s >> 40) & 0xFFFFFFULL) / (double)(1 << 24);
			out.push_back((float)((i + a) / side)), out.push_back((float)((j + b) / side));
		}
}

// sample camera: the forward vector offset by a sub-pixel step, n' = n + u*(dx*dp) + v*(dy*dp) (deliberately NOT
// re-normalised: primary rays are cam.n + cam.u*(xcur*dp) + cam.v*(ycur*dp), RayTracer.cpp:20-27, so this is the ray
// through (x + dx, y + dy))
template<class CameraT> inline CameraT jittered_camera(const CameraT &base, float dx, float dy)
{
	CameraT c = base;
	const double dp = tan(c.fovy * 3.1415926535897932384626433832795 / 360) / (c.height / 2);
	const Vertex n = c.n + c.u * (float)(dx * dp) + c.v * (float)(dy * dp);
	c.n.x = n.x, c.n.y = n.y, c.n.z = n.z, c.n.w = n.w;
	return c;
}

template<class SceneT> inline bool build(SceneT &scene, const SceneArgs &a)
{
	if (a.name == "c1") build_c1(scene, a);
	else if (a.name == "c2") build_c2(scene, a);
	else if (a.name == "c3") build_c3(scene, a);
	else if (a.name == "c4" || a.name == "c5") build_c4(scene, a);
	else if (a.name == "t_mixed") build_t_mixed(scene, a);
	else if (a.name == "t_ballplane") build_t_ballplane(scene, a);
	else if (a.name == "t_mesh") build_t_mesh(scene, a);
	else if (a.name == "t_twomesh") build_t_twomesh(scene, a);
	else if (a.name == "t_textured") build_t_textured(scene, a);
	else if (a.name == "t_empty") build_t_empty(scene, a);
	else if (a.name == "t_nolight") build_t_nolight(scene, a);
	else if (a.name == "t_lights") build_t_lights(scene, a);
	else if (a.name == "t_inside") build_t_inside(scene, a);
	else return false;
	return true;
}

}  // namespace rtscenes
