// SYNTAX-CHECK STAND-IN, not Pangolin: the two types the reference's headers name.
#pragma once
typedef unsigned char GLubyte; typedef unsigned int GLuint; typedef float GLfloat; typedef double GLdouble; typedef int GLint;
namespace pangolin {
struct OpenGlMatrix { double m[16]; void SetIdentity() {} };
struct OpenGlRenderState {};
}  // namespace pangolin
