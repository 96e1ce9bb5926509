// OBJ / MTL / BMP loading for Model.  Behavioural twin of /root/reference/Model.cpp:14-400,870-910
// including its quirks, because they decide which triangles and bounds reach the tracer:
//   * tokens: at most five whitespace-separated words per 255-char line, the first one padded
//     with "****" and dispatched on its first three characters (Model.cpp:7-12,880-888);
//   * words of a short line keep the values of the previous line (the token array is reused);
//   * faces may be triangles or quads (split 0-1-2 / 0-2-3), "v/vt/vn" or "v//vn" (:84-137,903-906);
//   * `usemtl` starts a new part and RESETS the running bounds even before the first part is
//     closed, so triangles that precede the first usemtl are not covered by their part's box (:144-161);
//   * the whole model is rescaled so that its largest extent is 8, with the reference's
//     asymmetric extent pick (:185-197);
//   * materials are cumulative: `newmtl` does not reset the running material (:229-236).
#include "Model.h"
#include <atomic>
#include <cstdio>
#include <cstdlib>

static std::atomic<uint64_t> g_epoch{ 1 };

static inline int32_t tag3(std::string s)
{
	if (s.length() < 3)
		s += "****";
	return s[2] + 256 * s[1] + 65536 * s[0];
}
static constexpr int32_t tag(char a, char b, char c) { return a * 65536 + b * 256 + c; }

Model::Loader::Loader(const std::wstring &fname)
{
	std::string narrow;
	for (wchar_t c : fname) narrow.push_back((char)c);
	fp = fopen(narrow.c_str(), "r");
	line[0] = 0;
}

Model::Loader::~Loader()
{
	if (fp) fclose(fp);
}

// returns the number of words, -1 for an empty line, INT8_MIN at end of file
int8_t Model::Loader::read(std::string data[])
{
	if (!fp || !fgets(line, 256, fp))
		return INT8_MIN;
	char word[5][256];
	const int n = sscanf(line, "%s%s%s%s%s", word[0], word[1], word[2], word[3], word[4]);
	for (int k = 0; k < n; ++k)
		data[k] = k == 0 ? std::string(word[0]) + "****" : std::string(word[k]);
	return (int8_t)n;
}

int8_t Model::Loader::parseInt(const std::string &in, int out[])
{
	int n = sscanf(in.c_str(), "%d/%d/%d/%d", &out[0], &out[1], &out[2], &out[3]);
	if (n < 2)
		n = sscanf(in.c_str(), "%d//%d/%d", &out[0], &out[2], &out[3]);
	return (int8_t)n;
}

static inline void grow(Vertex &lo, Vertex &hi, const Vertex &p)
{
	// SSE min/max semantics: result = (a < b) ? a : b  /  (a > b) ? a : b  per lane
	lo = Vertex(lo.x < p.x ? lo.x : p.x, lo.y < p.y ? lo.y : p.y, lo.z < p.z ? lo.z : p.z, lo.w < p.w ? lo.w : p.w);
	hi = Vertex(hi.x > p.x ? hi.x : p.x, hi.y > p.y ? hi.y : p.y, hi.z > p.z ? hi.z : p.z, hi.w > p.w ? hi.w : p.w);
}

int32_t Model::loadobj(const std::wstring &objname, const uint8_t)
{
	Loader ldr(objname);
	std::string ele[5];
	int ti[16] = { 0 };
	bool firstPart = true;

	std::vector<Triangle> tris;
	vers.push_back(Vertex());     // OBJ indices are 1-based
	nors.push_back(Normal());
	txcs.push_back(Coord2D());
	const Vertex farMin(1000, 1000, 1000), farMax(-1000, -1000, -1000);
	VerMin = farMin, VerMax = farMax;

	for (;;)
	{
		const int8_t num = ldr.read(ele);
		if (num == INT8_MIN)
			break;
		if (num == -1)
			continue;
		switch (tag3(ele[0]))
		{
		case tag('v', '*', '*'):
			vers.push_back(Vertex(atof(ele[1].c_str()), atof(ele[2].c_str()), atof(ele[3].c_str())));
			break;
		case tag('v', 'n', '*'):
			nors.push_back(Normal(atof(ele[1].c_str()), atof(ele[2].c_str()), atof(ele[3].c_str())));
			break;
		case tag('v', 't', '*'):
			txcs.push_back(Coord2D(atof(ele[1].c_str()), atof(ele[2].c_str())));
			break;
		case tag('f', '*', '*'):
		{
			const bool quad = num > 4;
			int8_t got = ldr.parseInt(ele[1], &ti[0]) + ldr.parseInt(ele[2], &ti[3]) + ldr.parseInt(ele[3], &ti[6]);
			if (quad)
				got += ldr.parseInt(ele[4], &ti[9]);
			const bool textured = got >= (quad ? 12 : 9);
			auto corner = [&](int k, Vertex &p, Normal &n, Coord2D &t) { p = vers[ti[3 * k]], n = nors[ti[3 * k + 2]], t = textured ? txcs[ti[3 * k + 1]] : Coord2D(); };
			Vertex p[4]; Normal n[4]; Coord2D t[4];
			for (int k = 0; k < (quad ? 4 : 3); ++k)
				corner(k, p[k], n[k], t[k]);
			tris.push_back(Triangle(p[0], n[0], t[0], p[1], n[1], t[1], p[2], n[2], t[2]));
			if (quad)
				tris.push_back(Triangle(p[0], n[0], t[0], p[2], n[2], t[2], p[3], n[3], t[3]));
			for (int k = 0; k < (quad ? 4 : 3); ++k)
				grow(VerMin, VerMax, p[k]);
			break;
		}
		case tag('u', 's', 'e'):
		{
			if (!firstPart)
			{
				tris.shrink_to_fit();
				parts.push_back(std::move(tris));
				tris.clear();
				borders.push_back(VerMin);
				borders.push_back(VerMax);
			}
			VerMin = farMin, VerMax = farMax;
			firstPart = false;
			int8_t a = (int8_t)mtls.size();
			while (--a > 0)
				if (mtls[a].name == ele[1])
					break;
			part_mtl.push_back(a);
			break;
		}
		}
	}
	if (firstPart)
		part_mtl.push_back(0);
	tris.shrink_to_fit();
	parts.push_back(std::move(tris));
	borders.push_back(VerMin);
	borders.push_back(VerMax);

	// whole-model bounds and the extent-8 rescale
	VerMin = farMin, VerMax = farMax;
	for (const Vertex &b : borders)
		grow(VerMin, VerMax, b);
	const Vertex dif = VerMax - VerMin;
	float scale = dif.x > dif.y ? dif.x : dif.y;
	scale = dif.z > scale ? dif.z / 8.0 : scale / 8.0;
	VerMax /= scale, VerMin /= scale;
	for (Vertex &b : borders)
		b /= scale;
	for (auto &part : parts)
		for (Triangle &t : part)
			for (Vertex &p : t.points)
				p /= scale;
	return (int32_t)parts.size();
}

int32_t Model::loadmtl(const std::wstring &mtlname, const uint8_t code)
{
	Loader ldr(mtlname);
	std::string ele[5];
	int8_t curTex = -1;
	bool first = true;
	Material mtl;
	mtls.push_back(mtl);       // slot 0: the fallback material
	mtl_tex.push_back(curTex);
	for (;;)
	{
		const int8_t num = ldr.read(ele);
		if (num == INT8_MIN)
			break;
		if (num == -1)
			continue;
		auto f = [&](int k) { return (float)atof(ele[k].c_str()); };
		switch (tag3(ele[0]))
		{
		case tag('n', 'e', 'w'):
			if (!first)
			{
				mtls.push_back(mtl);
				mtl_tex.push_back(curTex);
			}
			mtl.name = ele[1];
			curTex = -1;
			first = false;
			break;
		case tag('K', 'a', '*'): mtl.SetMtl(MY_MODEL_AMBIENT, f(1), f(2), f(3)); break;
		case tag('K', 'd', '*'): mtl.SetMtl(MY_MODEL_DIFFUSE, f(1), f(2), f(3)); break;
		case tag('K', 's', '*'): mtl.SetMtl(MY_MODEL_SPECULAR, f(1), f(2), f(3)); break;
		case tag('K', 'e', '*'): mtl.SetMtl(MY_MODEL_EMISSION, f(1), f(2), f(3)); break;
		case tag('N', 's', '*'): mtl.SetMtl(MY_MODEL_SHINESS, 0, 0, 0, f(1)); break;
		case tag('m', 'a', 'p'):
			if (ele[0] == "map_Kd****")
			{
				// keep the base name, force a .bmp extension
				const auto slash = ele[1].find_last_of('\\'), dot = ele[1].find_last_of('.');
				ele[1] = ele[1].substr(slash + 1, dot - slash) + "bmp";
				int8_t a = (int8_t)texs.size();
				while (--a >= 0)
					if (texs[a].name == ele[1])
					{
						curTex = a;
						break;
					}
				if (a < 0)
				{
					loadtex(ele[1], code);
					curTex = (int8_t)(texs.size() - 1);
				}
			}
			break;
		}
	}
	mtls.push_back(mtl);
	mtl_tex.push_back(curTex);
	return 0;
}

// 24-bit BMP, rows taken as tightly packed BGR (no 4-byte row padding handling), Model.cpp:282-309
int32_t Model::loadtex(const std::string &texname, const uint8_t)
{
	FILE *fp = fopen(texname.c_str(), "rb");
	if (!fp)
		return -1;
	uint8_t head[54];
	if (fread(head, 1, 54, fp) != 54 || head[0] != 'B' || head[1] != 'M')
	{
		fclose(fp);
		return -1;
	}
	auto le32 = [&](int off) { return (int32_t)(head[off] | head[off + 1] << 8 | head[off + 2] << 16 | (uint32_t)head[off + 3] << 24); };
	const int32_t offBits = le32(10), width = le32(18), height = le32(22);
	int32_t size = le32(34);
	if (size == 0)
		size = width * height * 3;
	fseek(fp, offBits, SEEK_SET);
	std::vector<uint8_t> image(size > width * height * 3 ? size : width * height * 3, 0);
	if (fread(image.data(), 1, size, fp) == 0) { /* short files leave zeros */ }
	fclose(fp);
	texs.push_back(Texture(texname, (int16_t)width, (int16_t)height, image.data()));
	return 0;
}

void Model::reset()
{
	VerMin = VerMax = Vertex();
	mtl_tex.clear(), part_mtl.clear(), texs.clear(), mtls.clear();
	parts.clear(), borders.clear(), bboxs.clear();
	vers.clear(), nors.clear(), txcs.clear();
}

Model::~Model() {}

int32_t Model::loadOBJ(const std::wstring &objname, const std::wstring &mtlname, const uint8_t code)
{
	this->objname = objname, this->mtlname = mtlname;
	reset();
	loadmtl(mtlname, code);
	loadobj(objname, code);
	touch();
	return 1;
}

// y-up <-> z-up swap used by the reference's benchmark scene (Model.cpp:343-388)
void Model::zRotate()
{
	auto turn = [](Vertex &p) { std::swap(p.y, p.z); p.z *= -1; };
	for (Vertex &p : vers) turn(p);
	for (Normal &n : nors) turn(n);
	for (auto &part : parts)
		for (Triangle &t : part)
			for (int k = 0; k < 3; ++k)
				turn(t.points[k]), turn(t.norms[k]);
	for (size_t k = 0; k + 1 < borders.size(); k += 2)
	{
		turn(borders[k]), turn(borders[k + 1]);
		std::swap(borders[k].z, borders[k + 1].z);
	}
	turn(VerMin), turn(VerMax);   // note: z of min/max is NOT re-ordered here (Model.cpp:383-386)
	touch();
}

void Model::SetMtl(const Material &mtl)
{
	for (auto &m : mtls)
		m = mtl;
}

// Host part of Model::RTPrepare (Model.cpp:404,418-419): translated bounds.  The octant binning
// (Model.cpp:430-472) is evaluated on the GPU when the scene is uploaded.
void Model::RTPrepare()
{
	BorderMin = VerMin + position, BorderMax = VerMax + position;
	bboxs.clear();
	for (const Vertex &b : borders)
		bboxs.push_back(b + position);
}

void Model::touch() { geometryEpoch = g_epoch.fetch_add(1); }
