#include "SceneUpload.h"
#include <stdexcept>

static inline rt_vec4 v4(const Vertex &v) { return rt_vec4{ v.x, v.y, v.z, v.w }; }

uint32_t SceneFlattener::addMaterial(const Material &m)
{
	rt_material r;
	r.ambient = v4(m.ambient), r.diffuse = v4(m.diffuse), r.specular = v4(m.specular), r.emission = v4(m.emission);
	r.shiness = m.shiness, r.reflect = m.reflect, r.refract = m.refract, r.rfr = m.rfr;
	materials.push_back(r);
	materialPtrs.push_back(&m);
	return (uint32_t)materials.size() - 1;
}

int32_t SceneFlattener::addTexture(const Texture &t)
{
	if (t.data == nullptr || t.w <= 0 || t.h <= 0)
		return -1;
	rt_texture r;
	r.w = t.w, r.h = t.h, r.offset = (uint32_t)texels.size(), r.pad0 = 0;
	// one spare texel: the reference can read one past the last row when a tiny negative
	// coordinate wraps to exactly 1.0 (3DElement.cpp:437-449)
	texels.insert(texels.end(), t.data, t.data + (size_t)t.w * t.h * 3);
	texels.insert(texels.end(), 4, 0);
	textures.push_back(r);
	texturePtrs.push_back(&t);
	return (int32_t)textures.size() - 1;
}

void SceneFlattener::cameraRecord(const Camera &cam, rt_camera &out)
{
	memset(&out, 0, sizeof out);
	out.u = v4(cam.u), out.v = v4(cam.v), out.n = v4(cam.n), out.position = v4(cam.position);
	out.width = cam.width, out.height = cam.height;
	out.fovy = cam.fovy, out.zNear = cam.zNear, out.zFar = cam.zFar;
}

void SceneFlattener::flatten(const Scene &scene, rt_scene_desc &desc)
{
	lights.clear(), materials.clear(), textures.clear(), texels.clear(), materialPtrs.clear(), texturePtrs.clear();
	prims.clear(), models.clear(), parts.clear();

	memset(&desc, 0, sizeof desc);
	const Camera &cam = scene.cam;
	cameraRecord(cam, desc.camera);
	desc.env_light = v4(scene.EnvLight);

	for (const Light &l : scene.Lights)
	{
		rt_light r;
		memset(&r, 0, sizeof r);
		r.position = v4(l.position), r.ambient = v4(l.ambient), r.diffuse = v4(l.diffuse), r.specular = v4(l.specular);
		r.attenuation = v4(l.attenuation);
		r.type = l.type, r.enabled = l.bLight ? 1u : 0u;
		lights.push_back(r);
	}

	// do the resident triangle arrays still describe the visible models?
	std::vector<uint64_t> epochs;
	for (DrawObject *o : scene.Objects)
		if (o->bShow && o->type == MY_OBJECT_MODEL)
			epochs.push_back(dynamic_cast<Model &>(*o).epoch());
	const bool rebuildTris = epochs != modelEpochs || geometryEpoch == 0;
	if (rebuildTris)
	{
		triPoints.clear(), triNorms.clear(), triTcoords.clear();
		modelEpochs = epochs;
		static uint64_t counter = 0;
		geometryEpoch = ++counter;
	}

	uint32_t triCursor = 0;
	for (size_t oi = 0; oi < scene.Objects.size(); ++oi)
	{
		DrawObject *o = scene.Objects[oi];
		if (!o->bShow)
			continue;
		rt_prim p;
		memset(&p, 0, sizeof p);
		p.object = (uint32_t)oi, p.texture = -1;
		p.position = v4(o->position);
		switch (o->type)
		{
		case MY_OBJECT_SPHERE:
		{
			const Sphere &s = dynamic_cast<Sphere &>(*o);
			p.kind = RT_OBJ_SPHERE, p.radius = s.getRadius(), p.radius_sqr = s.getRadiusSqr();
			p.material = addMaterial(o->mtl);
			prims.push_back(p);
			break;
		}
		case MY_OBJECT_CUBE:
		{
			const Box &b = dynamic_cast<Box &>(*o);
			p.kind = RT_OBJ_CUBE, p.a = v4(b.getMin()), p.b = v4(b.getMax());
			p.material = addMaterial(o->mtl);
			prims.push_back(p);
			break;
		}
		case MY_OBJECT_PLANE:
		{
			const Plane &pl = dynamic_cast<Plane &>(*o);
			p.kind = RT_OBJ_PLANE, p.a = v4(pl.normal), p.b = v4(pl.getAxisX()), p.c = v4(pl.getAxisY());
			p.material = addMaterial(o->mtl);
			p.texture = addTexture(pl.getTex());
			prims.push_back(p);
			break;
		}
		case MY_OBJECT_BALLPLANE:
		{
			const BallPlane &bp = dynamic_cast<BallPlane &>(*o);
			p.kind = RT_OBJ_SPHERE, p.radius = bp.getRadius(), p.radius_sqr = bp.getRadiusSqr();
			p.material = addMaterial(o->mtl);
			uint32_t slot = 0;
			for (const Vertex &c : bp.latticeCentres())
			{
				p.sub = ++slot;
				p.position = v4(c);
				prims.push_back(p);
			}
			break;
		}
		case MY_OBJECT_MODEL:
		{
			const Model &m = dynamic_cast<Model &>(*o);
			rt_model rm;
			memset(&rm, 0, sizeof rm);
			rm.object = (uint32_t)oi;
			rm.part_begin = (uint32_t)parts.size(), rm.part_count = (uint32_t)m.parts.size();
			rm.position = v4(m.position), rm.ver_min = v4(m.getVerMin()), rm.ver_max = v4(m.getVerMax());
			const uint32_t mtlBase = (uint32_t)materials.size();
			for (const Material &mm : m.mtls)
				addMaterial(mm);
			std::vector<int32_t> texIndex;
			for (const Texture &t : m.texs)
				texIndex.push_back(addTexture(t));
			for (size_t pi = 0; pi < m.parts.size(); ++pi)
			{
				const std::vector<Triangle> &part = m.parts[pi];
				if (part.size() > 32767)
					throw std::runtime_error("Model part exceeds 32767 triangles (clTri::numb is int16, 3DElement.h:125)");
				rt_part rp;
				rp.border_min = v4(m.borders[2 * pi]), rp.border_max = v4(m.borders[2 * pi + 1]);
				rp.tri_begin = triCursor, rp.tri_count = (uint32_t)part.size();
				const int8_t mnum = m.part_mtl[pi];
				rp.material = mtlBase + (uint32_t)mnum;
				const int8_t tnum = m.mtl_tex[mnum];
				rp.texture = tnum >= 0 ? texIndex[tnum] : -1;
				parts.push_back(rp);
				triCursor += rp.tri_count;
				if (rebuildTris)
					for (const Triangle &t : part)
						for (int k = 0; k < 3; ++k)
						{
							triPoints.push_back(v4(t.points[k]));
							triNorms.push_back(v4(t.norms[k]));
							triTcoords.push_back(t.tcoords[k].u);
							triTcoords.push_back(t.tcoords[k].v);
						}
			}
			models.push_back(rm);
			break;
		}
		default:
			throw std::runtime_error("Scene contains an object type the B200 path cannot flatten");
		}
	}

	desc.n_lights = (uint32_t)lights.size(), desc.lights = lights.data();
	desc.n_materials = (uint32_t)materials.size(), desc.materials = materials.data();
	desc.n_textures = (uint32_t)textures.size(), desc.textures = textures.data();
	desc.texel_bytes = texels.size(), desc.texels = texels.data();
	desc.n_prims = (uint32_t)prims.size(), desc.prims = prims.data();
	desc.n_models = (uint32_t)models.size(), desc.models = models.data();
	desc.n_parts = (uint32_t)parts.size(), desc.parts = parts.data();
	desc.n_tris = triCursor;
	desc.tri_points = triPoints.data(), desc.tri_norms = triNorms.data(), desc.tri_tcoords = triTcoords.data();
	desc.geometry_epoch = geometryEpoch;
}
