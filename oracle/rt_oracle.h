/*
 * rt_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement of the reference hot path
 * (Sh-Anand/Raytracer-in-CPP: Flyscene::raytraceScene -> traceRay -> BoxTree/BoundingBox ->
 * lightStrikes -> phongShade -> recursion -> PPM quantiser) in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (raytracer-in-cpp_b200/) never links
 * or calls it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the port is
 * pinned against the reference itself compiled headless (oracle/_ref/ref_oracle, built from the
 * sources under /root/reference by oracle/Makefile): tests/test_oracle_vs_reference.py compares
 * float RGB / face id / t bit-for-bit with fixtures under tests/golden/ that
 * tests/golden/make_golden.py generated from ref_oracle, and live against ref_oracle when present.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  float kd[3];   /* Mtl::getDiffuse        dependencies/tucano/tucano/materials/mtl.hpp:108 */
  float ks[3];   /* Mtl::getSpecular */
  float ns;      /* Mtl::getShininess */
  float ni;      /* Mtl::getOpticalDensity */
  int32_t illum; /* Mtl::getIlluminationModel */
} OrMaterial;

/* Flat, already-baked scene: exactly the values the reference's inner loops read
 * (src/flyscene.cpp:788-792, 867-872, 712-713). */
typedef struct {
  int32_t n_faces;
  const float *verts;     /* [T][3][3] world-space (getShapeModelMatrix()*v).head<3>() */
  const float *fnormals;  /* [T][3]    Face::normal */
  const float *vnormals;  /* [T][3][3] mesh.getNormal(vertex_ids[k]) */
  const int32_t *mat_id;  /* [T] */
  int32_t n_mats;
  const OrMaterial *mats;
  float model_matrix[12]; /* 3x4 row-major Affine applied to the interpolated normal, src/flyscene.cpp:829 */
  /* optional analytic spheres (NOT in the reference: parity unpinned) */
  int32_t n_spheres;
  const float *spheres;        /* [S][4] centre xyz, radius */
  const int32_t *sphere_mat;   /* [S] */
} OrSceneDesc;

typedef struct {
  float eye[3];        /* Camera::getCenter          tucano/camera.hpp:115 */
  float view_inv[12];  /* getViewMatrix().inverse()  3x4 row-major */
  float viewport[4];   /* (0,0,w,h)                  tucano/camera.hpp:319 */
  float fovy;          /* degrees */
  float aspect;
} OrCamera;

typedef struct {
  int32_t area_light;    /* stdin flag 1, src/flyscene.cpp:31-32 */
  int32_t point_light;   /* stdin flag 2, src/flyscene.cpp:33-34 */
  int32_t max_depth;     /* rays at level >= max_depth shade as plain Phong; <0 = unbounded (reference) */
  int32_t usteps, vsteps;/* area grid, reference 5 x 5 (src/flyscene.cpp:971) */
  float area_len_x, area_len_y; /* 0.3, 0.15 */
  float light_color[3];  /* lightrep colour (1,1,0), src/flyscene.cpp:68 */
  int32_t capacity;      /* octree leaf capacity 1000, src/flyscene.cpp:86 */
  int32_t candidates;    /* 0 = reference octree candidates (faithful); 1 = every face (exact nearest hit) */
  int32_t recursion_guard; /* hard stop for unbounded mode (returns BACKGROUND), default 64 */
  /* spherical light mode (area_light = point_light = 0, src/flyscene.cpp:974-993): the reference's
   * random_device draws are replaced by u_k = (lowbias32(seed * 0x9E3779B9u + k) >> 8) / 2^24 */
  uint32_t sphere_seed;    /* default 1 */
  float sphere_radius;     /* lightrep.getBoundingSphereRadius(), 1.0000001f in the reference */
} OrParams;

typedef struct OrScene OrScene;

void or_default_params(OrParams *p);

OrScene *or_scene_create(const OrSceneDesc *desc, const OrParams *params);
void or_scene_destroy(OrScene *s);
void or_scene_root_box(const OrScene *s, float mn[3], float mx[3]);
/* octree statistics: leaves, inner nodes, face references, max leaf size */
void or_scene_octree_stats(const OrScene *s, int64_t out[4]);

/* a2 -- Camera::screenToWorld, tucano/camera.hpp:155-173 */
void or_screen_to_world(const OrCamera *cam, float i, float j, float out[3]);
/* a4 -- BoundingBox::boxIntersect, src/boundingBox.cpp:48-83 */
int or_box_intersect(const float mn[3], const float mx[3], const float o[3], const float dest[3]);
/* a5 -- BoxTree::intersect, src/boxTree.cpp:150-173; returns count, ids ascending & unique */
int or_octree_candidates(const OrScene *s, const float o[3], const float dest[3], int32_t *ids, int cap);
/* a7 -- Flyscene::rayTriangleIntersection, src/flyscene.cpp:787-819 (-72 = miss) */
float or_ray_triangle(const OrScene *s, const float o[3], const float d[3], int face);
/* a12 -- createSpherePoint / arealight::getPointLights; returns sample count (<=25) */
int or_light_samples(const OrParams *p, const float light[3], float *out /*[25][3]*/);
/* a8 -- lightStrikes, src/flyscene.cpp:912-954 */
int or_light_strikes(const OrScene *s, const float hit[3], const float *lights, int n, uint8_t *visible);
/* a9 -- phongShade, src/flyscene.cpp:822-859 */
void or_phong_shade(const OrScene *s, const float origin[3], const float hit[3], int face, const float *lights,
                    int n_lights, float rgb[3]);
/* a3 -- traceRay, src/flyscene.cpp:651-771.  Also reports the first hit (face id or -1, t). */
void or_trace_ray(const OrScene *s, const float o[3], const float d[3], int level, const float *lights,
                  int n_lights, float rgb[3], int32_t *face, float *t);
/* a13 -- writePPMImage quantiser, tucano/utils/ppmIO.hpp:145 */
int or_quantize(float c);

/* a1 -- the per-pixel loop of raytraceScene (src/flyscene.cpp:573-598,613-625) over a pixel list.
 * px/py: [n] pixel coordinates.  Outputs may be NULL.  threads<=0 -> 1. */
void or_render_pixels(const OrScene *s, const OrCamera *cam, const float *lights, int n_lights,
                      const int32_t *px, const int32_t *py, int64_t n, float *rgb /*[n][3]*/,
                      int32_t *face /*[n]*/, float *t /*[n]*/, uint8_t *rgb8 /*[n][3]*/, int threads);

/* ray census accumulated by or_render_pixels since the last reset */
void or_census_reset(void);
void or_census_get(int64_t out[3]); /* primary, shadow(any-hit queries), secondary */

#ifdef __cplusplus
}
#endif
#endif
