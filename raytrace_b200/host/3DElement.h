// Host-side element math and records of the B200 build.
//
// Same class names, members and operator meanings as the reference's element layer
// (/root/reference/3DElement.h:34-237) so that code written against the reference object model
// compiles unchanged; the implementation is scalar C++ (no SSE, no OpenGL).  Host arithmetic that
// feeds the scene (light placement, plane frames, material setup) reproduces the reference's
// operation order so an identical scene description reaches the GPU:
//   * dot = (x0*y0 + x1*y1) + (x2*y2 + 0)      -- dpps mask 0x71, 3DElement.cpp:206-214
//   * v / s multiplies by 1/s                   -- 3DElement.cpp:137-147
//   * Normal(Vertex) divides by sqrt(dot)       -- 3DElement.cpp:218-238
// Build with -ffp-contract=off.  The per-ray work itself never runs here: it runs on the GPU
// behind include/rt_b200.h.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

typedef unsigned int GLuint;
typedef int GLint;

#define MY_OBJECT_SPHERE    0x1
#define MY_OBJECT_CUBE      0x2
#define MY_OBJECT_MODEL     0x3
#define MY_OBJECT_PLANE     0x4
#define MY_OBJECT_BALLPLANE 0x5
const char MY_OBJECT_NAME[][10] = { "ERROR", "sphere", "cube", "model", "plane", "ballplane" };

#define MY_LIGHT_PARALLEL 0x1
#define MY_LIGHT_POINT    0x2
#define MY_LIGHT_SPOT     0x3
const char MY_LIGHT_NAME[][10] = { "ERROR", "parallel", "point", "spot" };

#define MY_MODEL_AMBIENT     0x1
#define MY_MODEL_DIFFUSE     0x2
#define MY_MODEL_SPECULAR    0x4
#define MY_MODEL_SHINESS     0x8
#define MY_MODEL_EMISSION    0x10
#define MY_MODEL_POSITION    0x100
#define MY_MODEL_ATTENUATION 0x200

#define MY_RAY_BASERAY    0x1
#define MY_RAY_SHADOWRAY  0x2
#define MY_RAY_REFLECTRAY 0x3
#define MY_RAY_REFRACTRAY 0x4

#ifndef PI
#define PI 3.1415926535897932384
#endif

struct Coord2D
{
	float u = 0.0f, v = 0.0f;
	Coord2D() {}
	Coord2D(const float &iu, const float &iv) : u(iu), v(iv) {}
	Coord2D operator+(const Coord2D &c) const { return Coord2D(u + c.u, v + c.v); }
	Coord2D operator*(const float &n) const { return Coord2D(u * n, v * n); }
	operator float *() { return &u; }
};

// 4-lane value; the w lane rides along through + - * / exactly as the reference's __m128 does
// (a point light's w = 1 leaks into directions but never into dot/length).
struct alignas(16) Vertex
{
	union
	{
		struct { float x, y, z, w; };
		struct { float r, g, b, alpha; };
	};
	Vertex() : x(0), y(0), z(0), w(0) {}
	Vertex(const float ix, const float iy, const float iz, const float ia = 0) : x(ix), y(iy), z(iz), w(ia) {}
	operator float *() { return &x; }

	float length_sqr() const { return (x * x + y * y) + (z * z + 0.0f); }
	float length() const { return std::sqrt(length_sqr()); }
	Vertex mixmul(const Vertex &v) const { return Vertex(x * v.x, y * v.y, z * v.z, w * v.w); }
	Vertex muladd(const float &n, const Vertex &v) const { return Vertex(std::fma(x, n, v.x), std::fma(y, n, v.y), std::fma(z, n, v.z), std::fma(w, n, v.w)); }

	Vertex operator+(const Vertex &v) const { return Vertex(x + v.x, y + v.y, z + v.z, w + v.w); }
	Vertex operator-(const Vertex &v) const { return Vertex(x - v.x, y - v.y, z - v.z, w - v.w); }
	Vertex operator*(const float &n) const { return Vertex(x * n, y * n, z * n, w * n); }
	Vertex operator/(const float &n) const { const float rec = 1 / n; return Vertex(x * rec, y * rec, z * rec, w * rec); }
	Vertex &operator+=(const Vertex &v) { return *this = *this + v; }
	Vertex &operator-=(const Vertex &v) { return *this = *this - v; }
	Vertex &operator*=(const float &n) { return *this = *this * n; }
	Vertex &operator/=(const float &n) { return *this = *this / n; }
	// cross product; lane w = w*v.w - w*v.w like the shuffled SSE form (3DElement.cpp:190-197)
	Vertex operator*(const Vertex &v) const
	{
		return Vertex(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x, w * v.w - w * v.w);
	}
	float operator&(const Vertex &v) const { return (x * v.x + y * v.y) + (z * v.z + 0.0f); }
};

struct Normal : public Vertex
{
	Normal() : Vertex() {}
	Normal(const float &ix, const float &iy, const float &iz, const float &iw = 0.0f) : Vertex(ix, iy, iz, iw) {}
	Normal(const Vertex &v)   // normalise: every lane divided by sqrt(dot)
	{
		const float len = std::sqrt(v.length_sqr());
		x = v.x / len, y = v.y / len, z = v.z / len, w = v.w / len;
	}
};

class Texture
{
public:
	std::string name;
	int16_t w = 0, h = 0;
	uint8_t *data = nullptr;   // BGR8, rows of 3*w bytes (3DElement.cpp:430-450)
	Texture(bool check = false);
	Texture(const std::string &iname, const int16_t iw, const int16_t ih, const uint8_t *img);
	~Texture();
	Texture(const Texture &t);
	Texture(Texture &&t);
	Texture &operator=(const Texture &t);
};

class Material
{
public:
	Vertex ambient, diffuse, specular, emission;
	float shiness, reflect, refract, rfr;
	std::string name;
	Material();
	void SetMtl(const uint8_t prop, const float r, const float g, const float b, const float a = 1.0f);
	void SetMtl(const uint8_t prop, const Vertex &v);
	void SetMtl(const uint8_t prop, const float val);
};

struct alignas(16) clTri
{
	Vertex axisu, axisv, p0;
	int16_t numa = 0, numb = 0;
	clTri(const Vertex &u = Vertex(), const Vertex &v = Vertex(), const Vertex &p = Vertex()) : axisu(u), axisv(v), p0(p) {}
};

class Triangle
{
public:
	Vertex points[3];
	Normal norms[3];
	Coord2D tcoords[3];
	Triangle() {}
	Triangle(const Vertex &va, const Vertex &vb, const Vertex &vc);
	Triangle(const Vertex &va, const Normal &na, const Vertex &vb, const Normal &nb, const Vertex &vc, const Normal &nc);
	Triangle(const Vertex &va, const Normal &na, const Coord2D &ta, const Vertex &vb, const Normal &nb, const Coord2D &tb,
		const Vertex &vc, const Normal &nc, const Coord2D &tc);
};

class Color : public Vertex
{
public:
	Color(const bool white = false);
	Color(const float &ix, const float &iy, const float &iz) : Vertex(ix, iy, iz, 1e20f) {}
	Color(const Vertex &v);
	Color(const Normal &n);
	Color(const Texture *tex, const Coord2D &coord);
	void set(const float depth, const float mindepth, const float maxdepth);
	void put(uint8_t *addr);
	void get(uint8_t *addr);
};

class Ray
{
public:
	Vertex origin;
	Normal direction;
	float mtlrfr = 1.0f;
	uint8_t type, isInside = 0x0;
	Ray(const Vertex &o, const Normal &dir, const uint8_t type = 0x0) : origin(o), direction(dir), type(type) {}
};

class HitRes
{
public:
	Vertex position;
	Normal normal;
	Coord2D tcoord;
	Material *mtl = nullptr;
	Texture *tex = nullptr;
	intptr_t obj = (intptr_t)this;
	float distance, rfr = 1.0f;
	uint8_t isInside = 0x0;
	HitRes(bool b = false) { distance = b ? 1e8 : 1e20; }
	HitRes(float dis) : distance(dis) {}
	bool operator<(const HitRes &right) { return distance < right.distance; }
	operator bool() { return distance < 1e8; }
};

// The per-primitive operator interface of the reference (3DElement.h:185-202).  In this build
// the intersect operator of every shipped primitive is evaluated on the GPU; intersect() on the
// host forwards a single query to the device (SceneUpload.cpp) and is meant for probes/tests,
// the frame path never calls it.
class DrawObject
{
protected:
	GLuint GLListNum;
public:
	Vertex position;
	Material mtl;
	uint8_t type = 0;
	bool bShow = true;

	DrawObject(GLuint n = 0) : GLListNum(n) {}
	virtual ~DrawObject() {}
	virtual void SetMtl(const Material &mtl) { this->mtl = mtl; }
	void GLDraw() {}
	virtual void GLPrepare() {}
	virtual void RTPrepare() {}
	virtual HitRes intersect(const Ray &ray, const HitRes &hr, const float min = 0);
};

class Light
{
public:
	Vertex position, ambient, diffuse, specular, attenuation;
	float rangy, rangz, rdis, angy, angz, dis;
	float coang, exponent;
	uint8_t type;
	bool bLight;

	Light(const uint8_t type);
	bool turn();
	void move(const float dangy, const float dangz, const float ddis);
	void SetProperty(const int16_t prop, const float r, const float g, const float b, const float a = 1.0f);
	void SetLumi(const float lum);
};

class Camera
{
public:
	Normal u, v, n;   // right, up, toward
	Vertex position;
	GLint width, height;
	float fovy, aspect, zNear, zFar;
	Camera(GLint w = 1120, GLint h = 630);
	void move(const float &x, const float &y, const float &z);
	void yaw(const float angz);
	void pitch(float angy);
	void resize(GLint w, GLint h);
};

void Coord_sph2car(float &angy, float &angz, const float dis, Vertex &v);
void Coord_sph2car2(float &angy, float &angz, const float dis, Vertex &v);
void Coord_car2sph(const Vertex &v, float &angy, float &angz, float &dis);
float mod(const float &l, const float &r);
