// Triangle-mesh object of the object model: OBJ/MTL/BMP loading and per-part bookkeeping.
// API twin of /root/reference/Model.h:6-52.  What differs by design: RTPrepare() no longer bins
// triangles into octants on the CPU (Model.cpp:402-480) -- the GPU builds an LBVH and replays the
// octant predicate on device -- and intersect() is evaluated on the GPU.
#pragma once
#include "3DElement.h"

class Model : public DrawObject
{
	class Loader
	{
		FILE *fp;
		char line[256];
	public:
		Loader(const std::wstring &fname);
		~Loader();
		int8_t read(std::string data[]);
		int8_t parseInt(const std::string &in, int out[]);
	};
	Vertex VerMin, VerMax, BorderMin, BorderMax;
	uint64_t geometryEpoch = 0;
public:
	std::vector<Vertex> vers;
	std::vector<Normal> nors;
	std::vector<Coord2D> txcs;
	std::vector<std::vector<Triangle>> parts;
	std::vector<Vertex> borders;   // per part: min, max (untranslated)
	std::vector<Vertex> bboxs;     // per part: min, max + position (filled by RTPrepare)
	std::vector<Material> mtls;
	std::vector<Texture> texs;
	std::vector<int8_t> part_mtl, mtl_tex;
	std::wstring objname, mtlname;
private:
	int32_t loadobj(const std::wstring &objname, const uint8_t code);
	int32_t loadmtl(const std::wstring &mtlname, const uint8_t code);
	int32_t loadtex(const std::string &texname, const uint8_t code);
	void reset();
public:
	Model(GLuint num = 0) : DrawObject(num) { type = MY_OBJECT_MODEL; }
	~Model() override;
	int32_t loadOBJ(const std::wstring &objname, const std::wstring &mtlname, const uint8_t code = 0x0);
	void zRotate();
	void SetMtl(const Material &mtl) override;
	void RTPrepare() override;

	const Vertex &getVerMin() const { return VerMin; }
	const Vertex &getVerMax() const { return VerMax; }
	// Bumped whenever triangle data changes (loadOBJ, zRotate, touch); lets the uploader keep
	// the device-resident triangles and BVH when only placement/materials changed.
	uint64_t epoch() const { return geometryEpoch; }
	void touch();
};
