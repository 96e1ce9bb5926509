// Flattens a Scene (the object graph RayTracer.cpp walks per ray) into the
// rt_scene_desc of include/rt_b200.h: the step that replaces per-ray virtual dispatch
// (RayTracer.cpp:458-465) by one upload of SoA-ready records.
#pragma once
#include "Scene.h"
#include "../../include/rt_b200.h"

class SceneFlattener
{
public:
	// Fills `desc` with pointers into this object's storage (valid until the next flatten()).
	// Heavy triangle arrays are only rebuilt when some Model's epoch changed.
	void flatten(const Scene &scene, rt_scene_desc &desc);
	// the C-ABI record of a Camera (3DElement.h:225-237)
	static void cameraRecord(const Camera &cam, rt_camera &out);

	// what the material / texture indices of the last flatten() refer to (for mapping device hits
	// back to HitRes::mtl / HitRes::tex pointers)
	std::vector<const Material *> materialPtrs;
	std::vector<const Texture *> texturePtrs;

private:
	std::vector<rt_light> lights;
	std::vector<rt_material> materials;
	std::vector<rt_texture> textures;
	std::vector<uint8_t> texels;
	std::vector<rt_prim> prims;
	std::vector<rt_model> models;
	std::vector<rt_part> parts;
	std::vector<rt_vec4> triPoints, triNorms;
	std::vector<float> triTcoords;
	std::vector<uint64_t> modelEpochs;   // epochs the triangle arrays were built from
	uint64_t geometryEpoch = 0;

	uint32_t addMaterial(const Material &m);
	int32_t addTexture(const Texture &t);
};
