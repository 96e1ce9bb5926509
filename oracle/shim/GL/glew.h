// Test-infrastructure shim (ours, not reference code): lets the reference's GL preview
// code compile headlessly so the oracle build (oracle/build_ref.sh) can link the
// reference's own ray tracer.  Every GL entry point is a no-op; glGenLists hands out
// an incrementing id.  Symbol list taken from a grep over the five reference .cpp files.
#pragma once
typedef unsigned int GLuint;
typedef int GLint;
typedef unsigned short GLushort;
typedef unsigned int GLenum;
typedef float GLfloat;
typedef int GLsizei;
typedef double GLdouble;
enum {
	GL_AMBIENT = 1, GL_BGR_EXT, GL_CLAMP, GL_COMPILE, GL_CONSTANT_ATTENUATION, GL_DIFFUSE,
	GL_EMISSION, GL_FILL, GL_FLOAT, GL_FRONT_AND_BACK, GL_LIGHT0, GL_LIGHT_MODEL_AMBIENT = GL_LIGHT0 + 16,
	GL_LINE, GL_LINEAR, GL_LINEAR_ATTENUATION, GL_NEAREST, GL_NORMAL_ARRAY, GL_POSITION,
	GL_QUADRATIC_ATTENUATION, GL_QUADS, GL_REPEAT, GL_RESCALE_NORMAL, GL_RGB, GL_SHININESS,
	GL_SPECULAR, GL_TEXTURE_2D, GL_TEXTURE_COORD_ARRAY, GL_TEXTURE_MAG_FILTER,
	GL_TEXTURE_MIN_FILTER, GL_TEXTURE_WRAP_S, GL_TEXTURE_WRAP_T, GL_TRIANGLES,
	GL_UNSIGNED_BYTE, GL_UNSIGNED_SHORT, GL_VERTEX_ARRAY
};
#define RT_GL_NOOP(name) template<class... A> inline void name(A&&...) {}
RT_GL_NOOP(glMaterialfv) RT_GL_NOOP(glVertex3fv) RT_GL_NOOP(glBindTexture) RT_GL_NOOP(glTexParameteri)
RT_GL_NOOP(glEnableClientState) RT_GL_NOOP(glDisableClientState) RT_GL_NOOP(glNewList) RT_GL_NOOP(glEndList)
RT_GL_NOOP(glPushMatrix) RT_GL_NOOP(glPopMatrix) RT_GL_NOOP(glMaterialf) RT_GL_NOOP(glVertexPointer)
RT_GL_NOOP(glTexCoord2f) RT_GL_NOOP(glNormalPointer) RT_GL_NOOP(glNormal3fv) RT_GL_NOOP(glLightfv)
RT_GL_NOOP(glDrawElements) RT_GL_NOOP(glTexCoordPointer) RT_GL_NOOP(glTexCoord2fv) RT_GL_NOOP(glLightf)
RT_GL_NOOP(glGenTextures) RT_GL_NOOP(glEnd) RT_GL_NOOP(glEnable) RT_GL_NOOP(glDisable) RT_GL_NOOP(glCallList)
RT_GL_NOOP(glBegin) RT_GL_NOOP(glutSolidSphere) RT_GL_NOOP(glTranslatef) RT_GL_NOOP(glTexImage2D)
RT_GL_NOOP(glScalef) RT_GL_NOOP(glPolygonMode) RT_GL_NOOP(glDeleteTextures) RT_GL_NOOP(glutWireSphere)
RT_GL_NOOP(glutSolidCube) RT_GL_NOOP(gluLookAt) RT_GL_NOOP(gluBuild2DMipmaps) RT_GL_NOOP(glTranslated)
RT_GL_NOOP(glLightModelfv) RT_GL_NOOP(glDrawArrays)
inline GLuint glGenLists(GLsizei n) { static GLuint next = 1; GLuint r = next; next += n; return r; }
