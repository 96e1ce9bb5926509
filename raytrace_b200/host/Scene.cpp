// Scene container: material library, light/object factories and the editing operations.
// Behavioural twin of /root/reference/Scene.cpp:6-301 (GL display lists dropped).
#include "Scene.h"

namespace
{
struct LibEntry { const char *name; float a[3], d[3], s[3], shine, reflect, refract, rfr; };
// the six library materials of Scene.cpp:13-66
const LibEntry kLibrary[] = {
	{ "brass",       { 0.329412f, 0.223529f, 0.027451f }, { 0.780392f, 0.568627f, 0.113725f }, { 0.992157f, 0.941176f, 0.807843f }, 27.8974f, 0.0f, 0.0f, 1.0f },
	{ "bas-sphere",  { 0.1f, 0.1f, 0.1f }, { 0.1f, 0.5f, 0.8f }, { 1.0f, 1.0f, 1.0f }, 100, 0.35f, 0.0f, 1.0f },
	{ "mirror",      { 0.1f, 0.1f, 0.1f }, { 0.1f, 0.1f, 0.1f }, { 1.0f, 1.0f, 1.0f }, 127, 0.95f, 0.0f, 1.0f },
	{ "green grass", { 0.1f, 0.4f, 0.1f }, { 0.1f, 0.5f, 0.1f }, { 1.0f, 1.0f, 1.0f }, 127, 0.55f, 0.0f, 1.0f },
	{ "grass",       { 0.1f, 0.1f, 0.1f }, { 0.9f, 0.9f, 0.9f }, { 1.0f, 1.0f, 1.0f }, 127, 0.15f, 0.75f, 1.5f },
	{ "wall",        { 0.2f, 0.2f, 0.2f }, { 0.7f, 0.7f, 0.7f }, { 0.5f, 0.5f, 0.5f }, 5, 0.15f, 0.0f, 1.0f },
};
}

Scene::Scene()
{
	EnvLight = Vertex(0.2f, 0.2f, 0.2f, 1.0f);
	for (const LibEntry &e : kLibrary)
	{
		Material m;
		m.name = e.name;
		m.SetMtl(MY_MODEL_AMBIENT, e.a[0], e.a[1], e.a[2]);
		m.SetMtl(MY_MODEL_DIFFUSE, e.d[0], e.d[1], e.d[2]);
		m.SetMtl(MY_MODEL_SPECULAR, e.s[0], e.s[1], e.s[2]);
		m.SetMtl(MY_MODEL_SHINESS, e.shine);
		m.reflect = e.reflect, m.refract = e.refract, m.rfr = e.rfr;
		MtlLiby.push_back(m);
	}
}

Scene::~Scene()
{
	for (DrawObject *o : Objects)
		delete o;
}

// comp = (ambient, diffuse, specular) split, normalised to sum 1 and scaled by the luminance in
// atte.alpha; at most 8 lights (Scene.cpp:83-97)
uint8_t Scene::AddLight(const uint8_t type, const Vertex &comp, const Vertex &atte)
{
	if (Lights.size() == 8)
		return 0xff;
	Light light(type);
	const float sum = comp.x + comp.y + comp.z;
	const Vertex share = comp / sum;
	light.SetProperty(MY_MODEL_AMBIENT, share.x, share.x, share.x);
	light.SetProperty(MY_MODEL_DIFFUSE, share.y, share.y, share.y);
	light.SetProperty(MY_MODEL_SPECULAR, share.z, share.z, share.z);
	light.SetProperty(MY_MODEL_ATTENUATION, atte.x, atte.y, atte.z);
	light.SetLumi(atte.alpha);
	Lights.push_back(light);
	return (uint8_t)(Lights.size() - 1);
}

uint8_t Scene::AddSphere(const float radius)
{
	Sphere *s = new Sphere(radius);
	s->position = Vertex(0.0, radius, 0.0);
	s->SetMtl(MtlLiby[1]);
	Objects.push_back(s);
	return (uint8_t)(Objects.size() - 1);
}

uint8_t Scene::AddCube(const float len)
{
	Box *b = new Box(len);
	b->position = Vertex(0.0, len / 2, 0.0);
	b->SetMtl(MtlLiby[0]);
	Objects.push_back(b);
	return (uint8_t)(Objects.size() - 1);
}

uint8_t Scene::AddModel(const std::wstring &objname, const std::wstring &mtlname, uint8_t code)
{
	Model *m = new Model();
	m->loadOBJ(objname, mtlname, code);
	Objects.push_back(m);
	return (uint8_t)(Objects.size() - 1);
}

uint8_t Scene::AddPlane()
{
	Plane *p = new Plane();
	Material m;
	m.reflect = 0.6f;
	p->SetMtl(m);
	Objects.push_back(p);
	return (uint8_t)(Objects.size() - 1);
}

uint8_t Scene::AddBallPlane(const float radius)
{
	BallPlane *b = new BallPlane(radius);
	b->SetMtl(MtlLiby[1]);
	Objects.push_back(b);
	return (uint8_t)(Objects.size() - 1);
}

bool Scene::ChgLightComp(const uint8_t type, const uint8_t num, const Vertex &v)
{
	if (num >= Lights.size())
		return false;
	Light &light = Lights[num];
	if (type == MY_LIGHT_LUMI)
	{
		light.SetLumi(light.attenuation.alpha * v.alpha);
		return true;
	}
	if (type != MY_LIGHT_COMPENT)
		return false;
	// re-weight the three components per colour channel, keeping their sum at the luminance
	Vertex ta = light.ambient * v.x, td = light.diffuse * v.y, ts = light.specular * v.z;
	float *pa = ta, *pd = td, *ps = ts;
	for (int ch = 0; ch < 3; ++ch)
	{
		const float norm = (pa[ch] + pd[ch] + ps[ch]) / light.attenuation.alpha;
		pa[ch] /= norm, ps[ch] /= norm, pd[ch] /= norm;
	}
	light.SetProperty(MY_MODEL_AMBIENT, ta.x, ta.y, ta.z);
	light.SetProperty(MY_MODEL_DIFFUSE, td.x, td.y, td.z);
	light.SetProperty(MY_MODEL_SPECULAR, ts.x, ts.y, ts.z);
	return true;
}

bool Scene::ChgMtl(const uint8_t num, const Material &mtl)
{
	if (num >= Objects.size())
		return false;
	Objects[num]->SetMtl(mtl);
	Plane *p = dynamic_cast<Plane *>(Objects[num]);
	if (mtl.name == "wall" && p != nullptr)
		p->setTex(Texture(false));
	return true;
}

bool Scene::ChgMtl(const uint8_t num, const Normal &clr)
{
	if (num >= Objects.size())
		return false;
	auto recolour = [clr](Vertex &c)
	{
		if (clr.w < 0.5f)
			c = clr * (c.r + c.g + c.b);   // keep brightness, change hue
		else
			c = clr;
	};
	if (Objects[num]->type == MY_OBJECT_MODEL)
		for (Material &m : dynamic_cast<Model &>(*Objects[num]).mtls)
			recolour(m.diffuse);
	else
		recolour(Objects[num]->mtl.diffuse);
	return true;
}

bool Scene::Delete(uint8_t type, const uint8_t num)
{
	if (type == MY_MODEL_OBJECT)
	{
		if (num >= Objects.size())
			return false;
		delete Objects[num];
		Objects.erase(Objects.begin() + num);
	}
	else if (type == MY_MODEL_LIGHT)
	{
		if (num >= Lights.size())
			return false;
		Lights.erase(Lights.begin() + num);
	}
	return true;
}

bool Scene::MovePos(const uint8_t type, const uint8_t num, const Vertex &v)
{
	if (type == MY_MODEL_LIGHT)
	{
		if (num >= Lights.size())
			return false;
		Lights[num].move(v.x, v.y, v.z);
		return true;
	}
	if (type != MY_MODEL_OBJECT || num >= Objects.size())
		return false;
	DrawObject *o = Objects[num];
	if (o->type == MY_OBJECT_PLANE)
		dynamic_cast<Plane &>(*o).rotate(v);
	else if (o->type == MY_OBJECT_BALLPLANE)
		dynamic_cast<BallPlane &>(*o).rotate(v);
	else
		o->position += v;
	return true;
}

bool Scene::Switch(uint8_t type, const uint8_t num, const bool isShow)
{
	const bool toggle = type & MY_MODEL_SWITCH;
	switch (type & 0x7f)
	{
	case MY_MODEL_LIGHT:
		if (num >= Lights.size())
			return false;
		Lights[num].bLight = toggle ? !Lights[num].bLight : isShow;
		return true;
	case MY_MODEL_OBJECT:
	{
		if (num >= Objects.size())
			return false;
		const bool old = Objects[num]->bShow;
		Objects[num]->bShow = toggle ? !old : isShow;
		return old;
	}
	}
	return false;
}
