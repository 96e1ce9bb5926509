// Host state of the analytic primitives (constructors and the spherical-angle frame of the
// plane-like objects).  Follows /root/reference/Basic3DObject.cpp:85-97,225-240,304-330,414-447.
// Intersection itself lives in raytrace_b200/csrc (device code).
#include "Basic3DObject.h"

Sphere::Sphere(const float r, GLuint lnum) : DrawObject(lnum)
{
	type = MY_OBJECT_SPHERE;
	radius = r;
	radius_sqr = r * r;
}

Box::Box(const float len, GLuint lnum) : DrawObject(lnum)
{
	type = MY_OBJECT_CUBE;
	width = height = length = len;
	const float half = len / 2;
	max = Vertex(half, half, half);
	min = max * -1;
}

// The reference's three-argument constructor never initialises `min` and negates `max`
// (Basic3DObject.cpp:234-240); kept so that scenes built through it upload the same box.
Box::Box(const float l, const float w, const float h, GLuint lnum) : DrawObject(lnum)
{
	type = MY_OBJECT_CUBE;
	length = l, width = w, height = h;
	max = Vertex(l / 2, w / 2, h / 2);
	max = max * -1;
}

// Shared by Plane and BallPlane (Basic3DObject.cpp:312-330 and :438-454): spherical angles
// (ang.x, ang.y) and distance ang.z give the foot point and the normal; the in-plane frame is
// the direction 90 degrees further along ang.x and its cross product with the normal.
static void plane_frame(Vertex &ang, const Vertex &delta, Vertex &position, Normal &normal, Normal &axisx, Normal &axisy)
{
	bool atOrigin = false;
	ang.x = mod(360 + ang.x - delta.y * 5, 360);
	ang.y = mod(360 + ang.y - delta.x * 5, 360);
	ang.z += delta.z;
	if (ang.z < 0.0f)
		ang.z = 0.0f;
	if (std::abs(ang.z) < 1e-5f)
		atOrigin = true, ang.z = 1;
	Coord_sph2car2(ang.x, ang.y, ang.z, position);
	normal = Normal(position * -1);
	if (atOrigin)
		ang.z = 0, position = Vertex();
	float ahead = mod(90 + ang.x, 360);
	Coord_sph2car2(ahead, ang.y, 1, axisy);
	axisx = axisy * normal;
}

Plane::Plane(GLuint lnum) : DrawObject(lnum)
{
	type = MY_OBJECT_PLANE;
	tex = Texture(true);
	rotate(Vertex(0, 36, 0));
}

void Plane::rotate(const Vertex &v) { plane_frame(ang, v, position, normal, axisx, axisy); }

BallPlane::BallPlane(const float r, GLuint lnum) : DrawObject(lnum)
{
	type = MY_OBJECT_BALLPLANE;
	radius = r;
	radius_sqr = r * r;
	rotate(Vertex(0, 36, 0));
}

void BallPlane::rotate(const Vertex &v) { plane_frame(ang, v, position, normal, axisx, axisy); }

std::vector<Vertex> BallPlane::latticeCentres() const
{
	std::vector<Vertex> out;
	for (auto cx = radius * -6; cx < radius * 8; cx += radius * 4)
		for (auto cy = radius * -6; cy < radius * 8; cy += radius * 4)
		{
			const Vertex offset = axisx * cx + axisy * cy;
			out.push_back(offset + position);
		}
	return out;
}
