// host_shim/glad/glad.h — just the GL typedefs the reference's class declarations mention.  TEST INFRASTRUCTURE ONLY.
#pragma once
typedef unsigned int GLuint;
typedef int GLint;
typedef unsigned int GLenum;
