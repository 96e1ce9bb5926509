// Host-side records: textures, materials, colours, lights, camera and the spherical-coordinate
// helpers that place the default lights and plane frames.  Behavioural twin of
// /root/reference/3DElement.cpp (cited per function); written for plain C++17, no SSE/GL.
#include "3DElement.h"

// ---- spherical helpers (3DElement.cpp:4-38 of the reference) ---------------------------------
// All trigonometry is evaluated in double and rounded once on assignment, like the reference.

float mod(const float &l, const float &r)
{
	float whole;
	std::modf(l / r, &whole);
	return l - whole * r;
}

static inline double deg2rad(const float deg) { return deg * PI / 180; }

void Coord_sph2car(float &angy, float &angz, const float dis, Vertex &v)
{
	v.z = dis * std::sin(deg2rad(angy)) * std::cos(deg2rad(angz));
	v.x = dis * std::sin(deg2rad(angy)) * std::sin(deg2rad(angz));
	v.y = dis * std::cos(deg2rad(angy));
}

void Coord_sph2car2(float &angy, float &angz, const float dis, Vertex &v)
{
	bool flipped = false;
	if (angz >= 180)
	{
		angz = mod(angz, 180);
		angy = mod(360 - angy, 360);
		flipped = true;
	}
	if (angy < 1e-6)
		angy = 360;
	v.z = dis * std::sin(deg2rad(angy)) * std::cos(deg2rad(angz));
	v.x = dis * std::sin(deg2rad(angy)) * std::sin(deg2rad(angz));
	if (flipped && mod(angy, 180) < 1e-6)
	{
		v.z *= -1;
		v.x *= -1;
	}
	v.y = dis * std::cos(deg2rad(angy));
}

void Coord_car2sph(const Vertex &v, float &angy, float &angz, float &dis)
{
	dis = v.length();
	angy = std::acos(v.y / dis) * 180 / PI;
	angz = std::atan2(v.x, v.z) * 180 / PI;
}

// ---- Texture (3DElement.cpp:242-336) -----------------------------------------------------------

Texture::Texture(bool check)
{
	w = h = 4;
	data = new uint8_t[48];
	if (!check)
	{
		name = "empty";
		memset(data, 0xff, 48);
		return;
	}
	name = "check";
	for (int row = 0; row < 4; ++row)
		for (int col = 0; col < 4; ++col)
		{
			const uint8_t shade = ((row & 0x2) == (col & 0x2)) ? 0xff : 0x7f;
			uint8_t *px = data + (4 * row + col) * 3;
			px[0] = px[1] = px[2] = shade;
		}
}

Texture::Texture(const std::string &iname, const int16_t iw, const int16_t ih, const uint8_t *img) : name(iname), w(iw), h(ih)
{
	const int32_t size = w * h * 3;
	data = new uint8_t[size];
	memcpy(data, img, size);
}

Texture::~Texture() { delete[] data; }

Texture::Texture(const Texture &t) : name(t.name), w(t.w), h(t.h)
{
	if (t.data)
	{
		const int32_t size = w * h * 3;
		data = new uint8_t[size];
		memcpy(data, t.data, size);
	}
}

Texture::Texture(Texture &&t) : name(std::move(t.name)), w(t.w), h(t.h), data(t.data) { t.data = nullptr; }

Texture &Texture::operator=(const Texture &t)
{
	if (this == &t)
		return *this;
	name = t.name;
	w = t.w, h = t.h;
	delete[] data;
	data = nullptr;
	if (t.data)
	{
		const int32_t size = w * h * 3;
		data = new uint8_t[size];
		memcpy(data, t.data, size);
	}
	return *this;
}

// ---- Material (3DElement.cpp:340-377) ----------------------------------------------------------

Material::Material()
{
	name = "simple";
	SetMtl(MY_MODEL_AMBIENT | MY_MODEL_DIFFUSE, 0.588f, 0.588f, 0.588f);
	SetMtl(MY_MODEL_EMISSION | MY_MODEL_SPECULAR, 0.0f, 0.0f, 0.0f);
	SetMtl(MY_MODEL_SHINESS, 10.0f);
	reflect = refract = 0.0f;
	rfr = 1.0f;
}

void Material::SetMtl(const uint8_t prop, const float r, const float g, const float b, const float a)
{
	SetMtl(prop, Vertex(r, g, b, a));
	SetMtl(prop, a);
}

void Material::SetMtl(const uint8_t prop, const Vertex &v)
{
	if (prop & MY_MODEL_AMBIENT) ambient = v;
	if (prop & MY_MODEL_DIFFUSE) diffuse = v;
	if (prop & MY_MODEL_EMISSION) emission = v;
	if (prop & MY_MODEL_SPECULAR) specular = v;
}

void Material::SetMtl(const uint8_t prop, const float val)
{
	if (prop & MY_MODEL_SHINESS) shiness = val;
}

// ---- Triangle (3DElement.cpp:381-404) ----------------------------------------------------------

Triangle::Triangle(const Vertex &va, const Vertex &vb, const Vertex &vc)
{
	points[0] = va, points[1] = vb, points[2] = vc;
}

Triangle::Triangle(const Vertex &va, const Normal &na, const Vertex &vb, const Normal &nb, const Vertex &vc, const Normal &nc)
{
	points[0] = va, points[1] = vb, points[2] = vc;
	norms[0] = na, norms[1] = nb, norms[2] = nc;
}

Triangle::Triangle(const Vertex &va, const Normal &na, const Coord2D &ta, const Vertex &vb, const Normal &nb, const Coord2D &tb,
	const Vertex &vc, const Normal &nc, const Coord2D &tc)
{
	points[0] = va, points[1] = vb, points[2] = vc;
	norms[0] = na, norms[1] = nb, norms[2] = nc;
	tcoords[0] = ta, tcoords[1] = tb, tcoords[2] = tc;
}

// ---- Color (3DElement.cpp:408-472) -------------------------------------------------------------
// alpha doubles as "distance of the hit this colour came from" (1e20 = none).

Color::Color(const bool white)
{
	r = g = b = white ? 1.0f : 0.0f;
	alpha = 1e20f;
}

Color::Color(const Vertex &v)
{
	r = v.x, g = v.y, b = v.z;
	alpha = 1e20f;
}

Color::Color(const Normal &n)
{
	r = 0.5 * (n.x + 1);
	g = 0.5 * (n.y + 1);
	b = 0.5 * (n.z + 1);
}

Color::Color(const Texture *tex, const Coord2D &coord)
{
	if (tex == nullptr)
	{
		r = g = b = 1.0f;
		return;
	}
	float whole;
	float fu = std::modf(coord.u, &whole), fv = std::modf(coord.v, &whole);
	if (fu < 0) fu += 1;
	if (fv < 0) fv += 1;
	const int16_t tx = (int16_t)(fu * tex->w), ty = (int16_t)(fv * tex->h);
	const uint8_t *px = tex->data + (ty * tex->w + tx) * 3;
	b = px[0] / 255.0f;
	g = px[1] / 255.0f;
	r = px[2] / 255.0f;
}

void Color::set(const float depth, const float mindepth, const float maxdepth)
{
	if (depth <= mindepth)
		r = 1.0f, g = b = 0.0f;
	else if (depth >= maxdepth)
		r = g = b = 0.0f;
	else
	{
		const float ld = std::log(depth), lm = std::log(maxdepth);
		r = g = b = (lm - ld) / lm;
	}
}

static inline uint8_t quantise(const float c) { return c > 1.0f ? 255 : (c < 0.0f ? 0 : (uint8_t)(c * 255)); }

void Color::put(uint8_t *addr)
{
	addr[0] = quantise(r), addr[1] = quantise(g), addr[2] = quantise(b);
}

void Color::get(uint8_t *addr)
{
	r = addr[0] / 255.0f, g = addr[1] / 255.0f, b = addr[2] / 255.0f;
}

// ---- Light (3DElement.cpp:504-569) -------------------------------------------------------------

Light::Light(const uint8_t type)
{
	this->type = type;
	bLight = true;
	rangy = 90, rangz = 0, rdis = 16;
	coang = exponent = 0;
	move(0, 0, 0);
	SetProperty(MY_MODEL_AMBIENT, 0.05f, 0.05f, 0.05f);
	SetProperty(MY_MODEL_DIFFUSE | MY_MODEL_SPECULAR, 1.0f, 1.0f, 1.0f);
	SetProperty(MY_MODEL_ATTENUATION, 1.0f, 0.0f, 0.0f);
	position.alpha = (type == MY_LIGHT_PARALLEL) ? 0.0f : 1.0f;   // w = 0 parallel, 1 point/spot
}

bool Light::turn() { return bLight = !bLight; }

void Light::move(const float dangy, const float dangz, const float ddis)
{
	rdis += ddis;
	if (rdis < 2) rdis = 2;
	else if (rdis > 64) rdis = 64;
	angy = rangy = mod(360 + rangy + dangy, 360);
	angz = rangz = mod(360 + rangz + dangz, 360);
	dis = rdis;
	Coord_sph2car2(angy, angz, dis, position);
}

void Light::SetProperty(const int16_t prop, const float r, const float g, const float b, const float a)
{
	const Vertex set(r, g, b, a);
	if (prop & MY_MODEL_AMBIENT) ambient = set;
	if (prop & MY_MODEL_DIFFUSE) diffuse = set;
	if (prop & MY_MODEL_SPECULAR) specular = set;
	if (prop & MY_MODEL_ATTENUATION) attenuation = set;
	if (prop & MY_MODEL_POSITION) position = set;
}

void Light::SetLumi(const float lum)
{
	const float ext = lum / attenuation.alpha;
	attenuation.alpha = lum;
	ambient *= ext;
	diffuse *= ext;
	specular *= ext;
}

// ---- Camera (3DElement.cpp:573-635) ------------------------------------------------------------

Camera::Camera(GLint w, GLint h)
{
	width = w, height = h;
	aspect = (float)w / h;
	fovy = 45.0, zNear = 1.0, zFar = 100.0;
	position = Vertex(0, 4, 15);
	u = Vertex(1, 0, 0);
	v = Vertex(0, 1, 0);
	n = Vertex(0, 0, -1);
}

void Camera::move(const float &x, const float &y, const float &z)
{
	position += u * x;
	position += v * y;
	position += n * z;
}

void Camera::yaw(const float angz)
{
	float ay = std::acos(n.y / 1) * 180 / PI, az = std::atan2(n.x, n.z) * 180 / PI;
	az -= angz;
	Coord_sph2car(ay, az, 1, n);
	ay = std::acos(u.y / 1) * 180 / PI;
	az = std::atan2(u.x, u.z) * 180 / PI;
	az -= angz;
	Coord_sph2car(ay, az, 1, u);
}

void Camera::pitch(float angy)
{
	float ay = std::acos(n.y / 1) * 180 / PI, az = std::atan2(n.x, n.z) * 180 / PI;
	if (ay - angy < 1.0) angy = ay - 1.0;
	if (ay - angy > 179.0) angy = ay - 179.0;
	ay -= angy;
	Coord_sph2car(ay, az, 1, n);
	v = u * n;
}

void Camera::resize(GLint w, GLint h)
{
	width = w, height = h;
	aspect = (float)w / h;
}
