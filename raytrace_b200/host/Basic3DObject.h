// Analytic primitives of the object model: Sphere, Box, Plane, BallPlane.
// API twin of /root/reference/Basic3DObject.h:5-68.  The GL tessellation/display-list half of
// the reference classes is dropped (out of scope); what is kept is the state
// the tracer reads, exposed through const accessors so SceneUpload.cpp can flatten it
// (fields that are private in the reference: radius, min/max, plane axes).
#pragma once
#include "3DElement.h"

class Sphere : public DrawObject
{
	float radius, radius_sqr;
public:
	Sphere(const float r = 1.0f, GLuint lnum = 0);
	float getRadius() const { return radius; }
	float getRadiusSqr() const { return radius_sqr; }
};

class Box : public DrawObject
{
	float width, height, length;
	Vertex min, max;
public:
	Box(const float len = 2.0, GLuint lnum = 0);
	Box(const float l, const float w, const float h, GLuint lnum = 0);
	const Vertex &getMin() const { return min; }
	const Vertex &getMax() const { return max; }
};

class Plane : public DrawObject
{
	Vertex ang;
	Normal axisx, axisy;
	Texture tex;
public:
	Normal normal;

	Plane(GLuint lnum = 0);
	void rotate(const Vertex &v);
	void setTex(const Texture &tex) { this->tex = tex; }
	const Normal &getAxisX() const { return axisx; }
	const Normal &getAxisY() const { return axisy; }
	const Texture &getTex() const { return tex; }
};

class BallPlane : public DrawObject
{
	Vertex ang;
	Normal axisx, axisy;
	float radius, radius_sqr;
public:
	Normal normal;

	BallPlane(const float r = 0.3f, GLuint lnum = 0);
	void rotate(const Vertex &v);
	float getRadius() const { return radius; }
	float getRadiusSqr() const { return radius_sqr; }
	// centres of the lattice spheres in the reference's loop order (Basic3DObject.cpp:492-497)
	std::vector<Vertex> latticeCentres() const;
};
