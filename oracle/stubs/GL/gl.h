// stub for cuda_gl_interop.h (pulled in by the reference globalstate.h:21); test-infrastructure only
typedef unsigned int GLuint;
typedef unsigned int GLenum;
