/*
 * gl_stubs.c -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * The reference ray tracer (Sh-Anand/Raytracer-in-CPP) links against GLEW/GLFW/OpenGL
 * for its interactive preview.  None of that is on the ray-tracing path, and this image
 * has no GL libraries, so the headless oracle build links the reference's own translation
 * units against these no-op definitions instead.  Nothing here is reference code.
 *
 * GLEW exposes every GL>1.1 entry point as a function-pointer *variable* named __glewXxx;
 * we define those variables (as void*) and point them at generic stubs.  On x86-64 SysV a
 * variadic-free stub taking (long,long,long,long) can stand in for any of the signatures
 * used, since all arguments are integer/pointer class or floats that we ignore.
 */
#include <stddef.h>

static unsigned g_next_id = 1;

static long stub_noop(void) { return 0; }
static long stub_one(void) { return 1; }
static long stub_new_id(void) { return (long)(g_next_id++); }

/* glGenBuffers / glGenVertexArrays / glGenTextures (GLsizei n, GLuint* ids) */
static void stub_gen(int n, unsigned *ids) {
  for (int i = 0; i < n; ++i) ids[i] = g_next_id++;
}
/* glGetShaderiv / glGetProgramiv (GLuint obj, GLenum pname, GLint* out): report success */
static void stub_getiv(unsigned obj, unsigned pname, int *out) {
  (void)obj; (void)pname;
  if (out) *out = 1;
}
/* glGetShaderInfoLog / glGetProgramInfoLog (obj, bufSize, GLsizei* len, char* log) */
static void stub_infolog(unsigned obj, int bufsize, int *len, char *log) {
  (void)obj;
  if (len) *len = 0;
  if (log && bufsize > 0) log[0] = 0;
}

#define FP(name, fn) void *name = (void *)(fn)

FP(__glewGetUniformLocation, stub_noop);
FP(__glewGetAttribLocation, stub_noop);
FP(__glewGetShaderiv, stub_getiv);
FP(__glewGetProgramiv, stub_getiv);
FP(__glewGetShaderInfoLog, stub_infolog);
FP(__glewGetProgramInfoLog, stub_infolog);
FP(__glewBindBuffer, stub_noop);
FP(__glewBufferData, stub_noop);
FP(__glewGenBuffers, stub_gen);
FP(__glewDeleteBuffers, stub_noop);
FP(__glewGenVertexArrays, stub_gen);
FP(__glewDeleteVertexArrays, stub_noop);
FP(__glewBindVertexArray, stub_noop);
FP(__glewUniformMatrix4fv, stub_noop);
FP(__glewUniformMatrix3fv, stub_noop);
FP(__glewUniformMatrix2fv, stub_noop);
FP(__glewUniform1i, stub_noop);
FP(__glewUniform2i, stub_noop);
FP(__glewUniform3i, stub_noop);
FP(__glewUniform4i, stub_noop);
FP(__glewUniform1f, stub_noop);
FP(__glewUniform2f, stub_noop);
FP(__glewUniform3f, stub_noop);
FP(__glewUniform4f, stub_noop);
FP(__glewUniform1iv, stub_noop);
FP(__glewUniform2iv, stub_noop);
FP(__glewUniform3iv, stub_noop);
FP(__glewUniform4iv, stub_noop);
FP(__glewUniform1fv, stub_noop);
FP(__glewUniform2fv, stub_noop);
FP(__glewUniform3fv, stub_noop);
FP(__glewUniform4fv, stub_noop);
FP(__glewShaderSource, stub_noop);
FP(__glewCompileShader, stub_noop);
FP(__glewCreateShader, stub_new_id);
FP(__glewDeleteShader, stub_noop);
FP(__glewAttachShader, stub_noop);
FP(__glewDetachShader, stub_noop);
FP(__glewCreateProgram, stub_new_id);
FP(__glewDeleteProgram, stub_noop);
FP(__glewLinkProgram, stub_noop);
FP(__glewUseProgram, stub_noop);
FP(__glewEnableVertexAttribArray, stub_noop);
FP(__glewDisableVertexAttribArray, stub_noop);
FP(__glewVertexAttribPointer, stub_noop);
FP(__glewVertexAttribIPointer, stub_noop);
FP(__glewPatchParameteri, stub_noop);
FP(__glewActiveTexture, stub_noop);
FP(__glewDebugMessageCallback, stub_noop);
FP(__glewMapBuffer, stub_noop);
FP(__glewUnmapBuffer, stub_one);
FP(__glewMapBufferRange, stub_noop);
FP(__glewTransformFeedbackVaryings, stub_noop);
FP(__glewBindBufferBase, stub_noop);
FP(__glewBindImageTexture, stub_noop);
FP(__glewGenerateMipmap, stub_noop);
FP(__glewTexImage3D, stub_noop);
FP(__glewGenFramebuffers, stub_gen);
FP(__glewBindFramebuffer, stub_noop);
FP(__glewDeleteFramebuffers, stub_noop);

unsigned char glewExperimental = 0;
unsigned glewInit(void) { return 0; }
const unsigned char *glewGetErrorString(unsigned e) { (void)e; return (const unsigned char *)"stub"; }
const unsigned char *glewGetString(unsigned e) { (void)e; return (const unsigned char *)"stub"; }

/* GL 1.1 entry points (plain functions in the headers) */
void glDrawElements(unsigned m, int c, unsigned t, const void *i) { (void)m; (void)c; (void)t; (void)i; }
void glDrawArrays(unsigned m, int f, int c) { (void)m; (void)f; (void)c; }
void glViewport(int x, int y, int w, int h) { (void)x; (void)y; (void)w; (void)h; }
void glEnable(unsigned c) { (void)c; }
void glDisable(unsigned c) { (void)c; }
void glGetIntegerv(unsigned p, int *d) { (void)p; if (d) *d = 0; }
void glDeleteTextures(int n, const unsigned *t) { (void)n; (void)t; }
void glBindTexture(unsigned t, unsigned x) { (void)t; (void)x; }
unsigned glGetError(void) { return 0; }
void glGenTextures(int n, unsigned *t) { stub_gen(n, t); }
void glClearColor(float r, float g, float b, float a) { (void)r; (void)g; (void)b; (void)a; }
void glClear(unsigned m) { (void)m; }
void glTexParameteri(unsigned a, unsigned b, int c) { (void)a; (void)b; (void)c; }
void glTexParameterf(unsigned a, unsigned b, float c) { (void)a; (void)b; (void)c; }
void glTexImage2D(unsigned a, int b, int c, int d, int e, int f, unsigned g, unsigned h, const void *p) {
  (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; (void)g; (void)h; (void)p;
}
void glTexImage1D(unsigned a, int b, int c, int d, int e, unsigned g, unsigned h, const void *p) {
  (void)a; (void)b; (void)c; (void)d; (void)e; (void)g; (void)h; (void)p;
}
void glPixelStorei(unsigned a, int b) { (void)a; (void)b; }
void glReadPixels(int x, int y, int w, int h, unsigned f, unsigned t, void *p) {
  (void)x; (void)y; (void)w; (void)h; (void)f; (void)t; (void)p;
}
void glGetTexImage(unsigned a, int b, unsigned c, unsigned d, void *p) { (void)a; (void)b; (void)c; (void)d; (void)p; }
void glDepthMask(unsigned char f) { (void)f; }
void glBlendFunc(unsigned s, unsigned d) { (void)s; (void)d; }
void glLineWidth(float w) { (void)w; }
void glPointSize(float w) { (void)w; }
void glPolygonMode(unsigned f, unsigned m) { (void)f; (void)m; }
void glCullFace(unsigned m) { (void)m; }
void glFinish(void) {}
void glFlush(void) {}
void glDrawBuffer(unsigned m) { (void)m; }
void glReadBuffer(unsigned m) { (void)m; }
void glGetFloatv(unsigned p, float *d) { (void)p; if (d) *d = 0; }
const unsigned char *glGetString(unsigned n) { (void)n; return (const unsigned char *)"stub"; }

/* GLFW: only glfwGetKey is referenced (Flyscene::simulate) */
int glfwGetKey(void *w, int k) { (void)w; (void)k; return 0; }
