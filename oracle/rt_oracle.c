/*
 * rt_oracle.c -- TEST INFRASTRUCTURE.  See rt_oracle.h for scope and parity status.
 *
 * Plain-C restatement of the reference hot path.  Floating point follows the reference build
 * exactly: IEEE binary32, no FMA contraction (compile with
 * -ffp-contract=off), Eigen 3.3.7 reduction order for 3-vectors  a0*b0 + (a1*b1 + a2*b2)
 * (dependencies/eigen/include/Eigen/src/Core/Redux.h:91-104), `normalized()` = v / sqrt(v.v)
 * with zero left untouched (Eigen/src/Core/Dot.h:124-134).
 *
 * All file:line citations are relative to /root/reference.
 */
#define _GNU_SOURCE 1
#include "rt_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* tiny vector helpers in Eigen's evaluation order                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float *p) { return V(p[0], p[1], p[2]); }
static inline void st3(float *p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
static inline v3 cmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
/* Eigen redux_novec_unroller<.,.,0,3>: func(coeff0, func(coeff1, coeff2)) */
static inline float dot(v3 a, v3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline v3 cross(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 normalized(v3 a) {
  float z = dot(a, a);
  if (z > 0.f) { float n = sqrtf(z); return V(a.x / n, a.y / n, a.z / n); }
  return a;
}
/* Affine3f * Vector3f (Eigen/src/Geometry/Transform.h:1372-1392): evaluated as
 * (Matrix4f * Vector4f(v,1)).head<3>(); the 4x4 * 4x1 product is the SSE column-major packet
 * path (etor_product_packet_impl), i.e. res = col0*x; res = col1*y + res; res = col2*z + res;
 * res = col3*1 + res  -- left-to-right accumulation, no FMA. */
static inline v3 affine_point(const float m[12], v3 p) {
  v3 r;
  r.x = m[3] * 1.0f + (m[2] * p.z + (m[1] * p.y + m[0] * p.x));
  r.y = m[7] * 1.0f + (m[6] * p.z + (m[5] * p.y + m[4] * p.x));
  r.z = m[11] * 1.0f + (m[10] * p.z + (m[9] * p.y + m[8] * p.x));
  return r;
}
static inline float fmin_std(float a, float b) { return (b < a) ? b : a; } /* std::min(a,b) */
static inline float fmax_std(float a, float b) { return (a < b) ? b : a; } /* std::max(a,b) */

/* ------------------------------------------------------------------------------------------ */
/* scene + octree                                                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  float mn[3], mx[3];
  int is_leaf, is_empty;
  int child[8];   /* node indices, -1 when the node has no children */
  int n_children; /* 0 or 8 */
  int32_t *faces;
  int n_faces, cap_faces;
} Node;

struct OrScene {
  OrSceneDesc d;
  OrParams p;
  float *verts, *fnormals, *vnormals, *spheres;
  int32_t *mat_id, *sphere_mat;
  OrMaterial *mats;
  Node *nodes;
  int n_nodes, cap_nodes;
};

static int64_t g_census[3];
static pthread_mutex_t g_census_mtx = PTHREAD_MUTEX_INITIALIZER;

void or_default_params(OrParams *p) {
  memset(p, 0, sizeof(*p));
  p->area_light = 0;
  p->point_light = 1;
  p->max_depth = -1;
  p->usteps = 5;
  p->vsteps = 5;
  p->area_len_x = 0.3f;
  p->area_len_y = 0.15f;
  p->light_color[0] = 1.f; p->light_color[1] = 1.f; p->light_color[2] = 0.f;
  p->capacity = 1000;
  p->candidates = 0;
  p->recursion_guard = 64;
  p->sphere_seed = 1u;
  p->sphere_radius = 1.00000012f; /* 0x3f800001, dumped from the reference (tests/golden: light_radius) */
}

static int node_new(OrScene *s) {
  if (s->n_nodes == s->cap_nodes) {
    s->cap_nodes = s->cap_nodes ? s->cap_nodes * 2 : 64;
    s->nodes = (Node *)realloc(s->nodes, sizeof(Node) * (size_t)s->cap_nodes);
  }
  Node *n = &s->nodes[s->n_nodes];
  memset(n, 0, sizeof(*n));
  for (int i = 0; i < 8; ++i) n->child[i] = -1;
  return s->n_nodes++;
}
static void node_push(Node *n, int32_t f) {
  if (n->n_faces == n->cap_faces) {
    n->cap_faces = n->cap_faces ? n->cap_faces * 2 : 16;
    n->faces = (int32_t *)realloc(n->faces, sizeof(int32_t) * (size_t)n->cap_faces);
  }
  n->faces[n->n_faces++] = f;
}

/* the six BoxTree::axisTest* helpers (src/boxTree.cpp:364-456) come in three shapes */
static int axis_x(float a, float b, float fa, float fb, v3 v0, v3 v1, v3 h) {
  float p0 = a * v0.y - b * v0.z, p1 = a * v1.y - b * v1.z;
  float mx = fmax_std(p1, p0), mn = fmin_std(p1, p0);
  float rad = fa * h.y + fb * h.z;
  return !(mn > rad || mx < -rad);
}
static int axis_y(float a, float b, float fa, float fb, v3 v0, v3 v1, v3 h) {
  float p0 = -a * v0.x + b * v0.z, p1 = -a * v1.x + b * v1.z;
  float mx = fmax_std(p1, p0), mn = fmin_std(p1, p0);
  float rad = fa * h.x + fb * h.z;
  return !(mn > rad || mx < -rad);
}
static int axis_z(float a, float b, float fa, float fb, v3 v0, v3 v1, v3 h) {
  float p0 = a * v0.x - b * v0.y, p1 = a * v1.x - b * v1.y;
  float mx = fmax_std(p1, p0), mn = fmin_std(p1, p0);
  float rad = fa * h.x + fb * h.y;
  return !(mn > rad || mx < -rad);
}
/* BoxTree::planeBoxOverlap, src/boxTree.cpp:338-361 */
static int plane_box_overlap(v3 normal, v3 vert, v3 maxbox) {
  float n[3] = {normal.x, normal.y, normal.z}, vv[3] = {vert.x, vert.y, vert.z},
        mb[3] = {maxbox.x, maxbox.y, maxbox.z}, vmin[3], vmax[3];
  for (int i = 0; i < 3; ++i) {
    float v = vv[i];
    if (n[i] > 0.0f) { vmin[i] = -mb[i] - v; vmax[i] = mb[i] - v; }
    else { vmin[i] = mb[i] - v; vmax[i] = -mb[i] - v; }
  }
  if (dot(normal, V(vmin[0], vmin[1], vmin[2])) > 0.0f) return 0;
  if (dot(normal, V(vmax[0], vmax[1], vmax[2])) >= 0.0f) return 1;
  return 0;
}

/* BoxTree::clasifyFace, src/boxTree.cpp:203-336 (vertex-in-box, else the SAT variant that the
 * reference evaluates on *normalised* centre-relative vectors, :236-240) */
static int classify_face(const OrScene *s, const Node *nd, int face) {
  const float *vp = s->verts + (size_t)face * 9;
  v3 vert[3] = {ld3(vp), ld3(vp + 3), ld3(vp + 6)};
  int inside = 0;
  for (int k = 0; k < 3; ++k) {
    v3 v = vert[k];
    if (nd->mn[0] <= v.x && nd->mx[0] >= v.x && nd->mn[1] <= v.y && nd->mx[1] >= v.y &&
        nd->mn[2] <= v.z && nd->mx[2] >= v.z)
      inside++;
  }
  if (inside > 0) return 1;

  v3 c = V(nd->mn[0] + (nd->mx[0] - nd->mn[0]) / 2.f, nd->mn[1] + (nd->mx[1] - nd->mn[1]) / 2.f,
           nd->mn[2] + (nd->mx[2] - nd->mn[2]) / 2.f);
  v3 h = normalized(sub(ld3(nd->mx), c));
  v3 a = normalized(sub(vert[0], c)), b = normalized(sub(vert[1], c)), cc = normalized(sub(vert[2], c));
  v3 e0 = sub(b, a), e1 = sub(cc, b), e2 = sub(a, cc);
  float fex, fey, fez;

  fex = fabsf(e0.x); fey = fabsf(e0.y); fez = fabsf(e0.z);
  if (!axis_x(e0.z, e0.y, fez, fey, a, cc, h)) return 0; /* axisTestX01 */
  if (!axis_y(e0.z, e0.x, fez, fex, a, cc, h)) return 0; /* axisTestY02 */
  if (!axis_z(e0.y, e0.x, fey, fex, b, cc, h)) return 0; /* axisTestZ12 */

  fex = fabsf(e1.x); fey = fabsf(e1.y); fez = fabsf(e1.z);
  if (!axis_x(e1.z, e1.y, fez, fey, a, cc, h)) return 0; /* axisTestX01 */
  if (!axis_y(e1.z, e1.x, fez, fex, a, cc, h)) return 0; /* axisTestY02 */
  if (!axis_z(e1.y, e1.x, fey, fex, a, b, h)) return 0;  /* axisTestZ0  */

  fex = fabsf(e2.x); fey = fabsf(e2.y); fez = fabsf(e2.z);
  if (!axis_x(e2.z, e2.y, fez, fey, a, b, h)) return 0;  /* axisTestX02 */
  if (!axis_y(e2.z, e2.x, fez, fex, a, b, h)) return 0;  /* axisTestY1  */
  if (!axis_z(e2.y, e2.x, fey, fex, b, cc, h)) return 0; /* axisTestZ12 */

  float mn, mx;
  mn = fmin_std(fmin_std(a.x, b.x), cc.x); mx = fmax_std(fmax_std(a.x, b.x), cc.x);
  if (mn > h.x || mx < -h.x) return 0;
  mn = fmin_std(fmin_std(a.y, b.y), cc.y); mx = fmax_std(fmax_std(a.y, b.y), cc.y);
  if (mn > h.y || mx < -h.y) return 0;
  mn = fmin_std(fmin_std(a.z, b.z), cc.z); mx = fmax_std(fmax_std(a.z, b.z), cc.z);
  if (mn > h.z || mx < -h.z) return 0;

  v3 nrm = normalized(cross(sub(a, b), sub(a, cc)));
  if (!plane_box_overlap(nrm, a, h)) return 0;
  return 1;
}

/* BoxTree::split, src/boxTree.cpp:88-147 */
static void split(OrScene *s, int ni, int depth) {
  s->nodes[ni].is_leaf = 0;
  float mn[3], mx[3];
  memcpy(mn, s->nodes[ni].mn, 12); memcpy(mx, s->nodes[ni].mx, 12);
  float dx = (mx[0] - mn[0]) / 2, dy = (mx[1] - mn[1]) / 2, dz = (mx[2] - mn[2]) / 2;
  v3 bmin = ld3(mn), bmax = ld3(mx);
  v3 vx = V(dx, 0, 0), vy = V(0, dy, 0), vz = V(0, 0, dz);
  v3 lo[8], hi[8];
  /* b000 .. b111 exactly as written at :103-110 (left-to-right vector sums) */
  lo[0] = bmin;                          hi[0] = add(add(add(bmin, vx), vy), vz);
  lo[1] = add(bmin, vz);                 hi[1] = add(add(add(bmin, vx), vy), mul(2, vz));
  lo[2] = add(bmin, vy);                 hi[2] = add(add(add(bmin, vx), mul(2, vy)), vz);
  lo[3] = add(add(bmin, vy), vz);        hi[3] = add(add(add(bmin, vx), mul(2, vy)), mul(2, vz));
  lo[4] = add(bmin, vx);                 hi[4] = add(add(add(bmin, mul(2, vx)), vy), vz);
  lo[5] = add(add(bmin, vx), vz);        hi[5] = sub(bmax, vy);
  lo[6] = add(add(bmin, vx), vy);        hi[6] = sub(bmax, vz);
  lo[7] = add(add(add(bmin, vx), vy), vz); hi[7] = bmax;

  int kids[8];
  for (int k = 0; k < 8; ++k) {
    int ci = node_new(s); /* may realloc s->nodes */
    kids[k] = ci;
    st3(s->nodes[ci].mn, lo[k]); st3(s->nodes[ci].mx, hi[k]);
  }
  for (int k = 0; k < 8; ++k) s->nodes[ni].child[k] = kids[k];
  s->nodes[ni].n_children = 8;

  for (int k = 0; k < 8; ++k) {
    Node *ch = &s->nodes[kids[k]];
    const Node *par = &s->nodes[ni];
    for (int i = 0; i < par->n_faces; ++i)
      if (classify_face(s, ch, par->faces[i])) node_push(ch, par->faces[i]);
  }
  free(s->nodes[ni].faces);
  s->nodes[ni].faces = NULL; s->nodes[ni].n_faces = 0; s->nodes[ni].cap_faces = 0;

  const int cap = s->p.capacity;
  for (int k = 0; k < 8; ++k) {
    int ci = kids[k];
    if (s->nodes[ci].n_faces == 0 && s->nodes[ci].n_children == 0) s->nodes[ci].is_empty = 1;
    if (s->nodes[ci].n_faces < cap || depth <= 0) s->nodes[ci].is_leaf = 1;
    if (s->nodes[ci].n_faces > cap && depth > 0) split(s, ci, depth - 1);
  }
}

OrScene *or_scene_create(const OrSceneDesc *desc, const OrParams *params) {
  OrScene *s = (OrScene *)calloc(1, sizeof(OrScene));
  s->d = *desc;
  s->p = *params;
  const size_t T = (size_t)desc->n_faces;
  s->verts = (float *)malloc(T * 36 + 4); memcpy(s->verts, desc->verts, T * 36);
  s->fnormals = (float *)malloc(T * 12 + 4); memcpy(s->fnormals, desc->fnormals, T * 12);
  s->vnormals = (float *)malloc(T * 36 + 4); memcpy(s->vnormals, desc->vnormals, T * 36);
  s->mat_id = (int32_t *)malloc(T * 4 + 4); memcpy(s->mat_id, desc->mat_id, T * 4);
  s->mats = (OrMaterial *)malloc(sizeof(OrMaterial) * (size_t)(desc->n_mats + 1));
  memcpy(s->mats, desc->mats, sizeof(OrMaterial) * (size_t)desc->n_mats);
  if (desc->n_spheres > 0) {
    s->spheres = (float *)malloc((size_t)desc->n_spheres * 16);
    memcpy(s->spheres, desc->spheres, (size_t)desc->n_spheres * 16);
    s->sphere_mat = (int32_t *)malloc((size_t)desc->n_spheres * 4);
    memcpy(s->sphere_mat, desc->sphere_mat, (size_t)desc->n_spheres * 4);
  }
  s->d.verts = s->verts; s->d.fnormals = s->fnormals; s->d.vnormals = s->vnormals;
  s->d.mat_id = s->mat_id; s->d.mats = s->mats; s->d.spheres = s->spheres; s->d.sphere_mat = s->sphere_mat;

  /* BoundingBox::BoundingBox(Mesh&), src/boundingBox.cpp:14-43: note max starts at FLT_MIN (>0) */
  float xmin = FLT_MAX, ymin = FLT_MAX, zmin = FLT_MAX, xmax = FLT_MIN, ymax = FLT_MIN, zmax = FLT_MIN;
  for (size_t i = 0; i < T * 3; ++i) {
    float x = s->verts[3 * i], y = s->verts[3 * i + 1], z = s->verts[3 * i + 2];
    xmin = fmin_std(xmin, x); ymin = fmin_std(ymin, y); zmin = fmin_std(zmin, z);
    xmax = fmax_std(xmax, x); ymax = fmax_std(ymax, y); zmax = fmax_std(zmax, z);
  }
  /* BoxTree::BoxTree(Mesh&, capacity), src/boxTree.cpp:11-31 */
  int root = node_new(s);
  s->nodes[root].mn[0] = xmin; s->nodes[root].mn[1] = ymin; s->nodes[root].mn[2] = zmin;
  s->nodes[root].mx[0] = xmax; s->nodes[root].mx[1] = ymax; s->nodes[root].mx[2] = zmax;
  for (size_t i = 0; i < T; ++i) node_push(&s->nodes[root], (int32_t)i);
  if ((int)T > s->p.capacity) split(s, root, 15 /* MAX_DEPTH, src/boxTree.cpp:3 */);
  else if (T == 0) s->nodes[root].is_empty = 1;
  else s->nodes[root].is_leaf = 1;
  return s;
}

void or_scene_destroy(OrScene *s) {
  if (!s) return;
  for (int i = 0; i < s->n_nodes; ++i) free(s->nodes[i].faces);
  free(s->nodes); free(s->verts); free(s->fnormals); free(s->vnormals); free(s->mat_id);
  free(s->mats); free(s->spheres); free(s->sphere_mat); free(s);
}

void or_scene_root_box(const OrScene *s, float mn[3], float mx[3]) {
  memcpy(mn, s->nodes[0].mn, 12); memcpy(mx, s->nodes[0].mx, 12);
}

void or_scene_octree_stats(const OrScene *s, int64_t out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0;
  int *stack = (int *)malloc(sizeof(int) * (size_t)(s->n_nodes + 1));
  int sp = 0; stack[sp++] = 0;
  while (sp) {
    const Node *n = &s->nodes[stack[--sp]];
    if (n->is_leaf && !n->is_empty) { out[0]++; out[2] += n->n_faces; if (n->n_faces > out[3]) out[3] = n->n_faces; }
    else if (!n->is_empty) { out[1]++; for (int k = 0; k < n->n_children; ++k) stack[sp++] = n->child[k]; }
  }
  free(stack);
}

/* ------------------------------------------------------------------------------------------ */
/* a4  BoundingBox::boxIntersect, src/boundingBox.cpp:48-83                                    */
/* ------------------------------------------------------------------------------------------ */
int or_box_intersect(const float mn[3], const float mx[3], const float o[3], const float dest[3]) {
  float dx = dest[0] - o[0], dy = dest[1] - o[1], dz = dest[2] - o[2];
  float txmin = (mn[0] - o[0]) / dx, txmax = (mx[0] - o[0]) / dx;
  float tymin = (mn[1] - o[1]) / dy, tymax = (mx[1] - o[1]) / dy;
  float tzmin = (mn[2] - o[2]) / dz, tzmax = (mx[2] - o[2]) / dz;
  float tinx = fmin_std(txmin, txmax), toutx = fmax_std(txmin, txmax);
  float tiny = fmin_std(tymin, tymax), touty = fmax_std(tymin, tymax);
  float tinz = fmin_std(tzmin, tzmax), toutz = fmax_std(tzmin, tzmax);
  float tin = fmax_std(fmax_std(tinx, tiny), tinz);
  float tout = fmin_std(fmin_std(toutx, touty), toutz);
  if ((tin > tout) || (tout < 0)) return 0;
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* a5  BoxTree::intersect, src/boxTree.cpp:150-173.  The reference walks breadth-first and    */
/*     collects into a std::set; the visit order does not change the set, so walk depth-first  */
/*     and hand every candidate face to a callback.                                            */
/* ------------------------------------------------------------------------------------------ */
typedef void (*face_cb)(void *ctx, int face);

static void octree_visit(const OrScene *s, int ni, const float o[3], const float dest[3], face_cb cb, void *ctx) {
  const Node *n = &s->nodes[ni];
  if (!or_box_intersect(n->mn, n->mx, o, dest)) return;
  if (n->is_leaf && !n->is_empty) {
    for (int i = 0; i < n->n_faces; ++i) cb(ctx, n->faces[i]);
  } else if (!n->is_empty) {
    for (int k = 0; k < n->n_children; ++k) {
      const Node *c = &s->nodes[n->child[k]];
      if (!c->is_empty && or_box_intersect(c->mn, c->mx, o, dest)) octree_visit(s, n->child[k], o, dest, cb, ctx);
    }
  }
}
static void for_each_candidate(const OrScene *s, const float o[3], const float dest[3], face_cb cb, void *ctx) {
  if (s->p.candidates == 1) {
    for (int i = 0; i < s->d.n_faces; ++i) cb(ctx, i);
  } else {
    octree_visit(s, 0, o, dest, cb, ctx);
  }
}

typedef struct { int32_t *ids; int n, cap; } IdList;
static void collect_cb(void *ctx, int face) {
  IdList *l = (IdList *)ctx;
  if (l->n < l->cap) l->ids[l->n] = face;
  l->n++;
}
static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}
int or_octree_candidates(const OrScene *s, const float o[3], const float dest[3], int32_t *ids, int cap) {
  IdList l = {ids, 0, cap};
  octree_visit(s, 0, o, dest, collect_cb, &l);
  int n = l.n < cap ? l.n : cap;
  qsort(ids, (size_t)n, sizeof(int32_t), cmp_i32);
  int m = 0;
  for (int i = 0; i < n; ++i) if (i == 0 || ids[i] != ids[i - 1]) ids[m++] = ids[i];
  return (l.n > cap) ? -l.n : m;
}

/* ------------------------------------------------------------------------------------------ */
/* a7  Flyscene::rayTriangleIntersection, src/flyscene.cpp:787-819                             */
/* ------------------------------------------------------------------------------------------ */
static float ray_triangle(const OrScene *s, v3 o, v3 d, int face) {
  const float *vp = s->verts + (size_t)face * 9;
  v3 a = ld3(vp), b = ld3(vp + 3), c = ld3(vp + 6);
  v3 n = ld3(s->fnormals + (size_t)face * 3);
  if (dot(d, n) == 0) return -72;
  float t = (dot(n, a) - dot(o, n)) / dot(d, n);
  v3 P = add(o, mul(t, d));
  v3 v0 = sub(c, a), v1 = sub(b, a), v2 = sub(P, a);
  float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d02 = dot(v0, v2), d12 = dot(v1, v2);
  float invDenom = 1 / (d00 * d11 - d01 * d01);
  float u = (d11 * d02 - d01 * d12) * invDenom;
  float v = (d00 * d12 - d01 * d02) * invDenom;
  if ((u >= 0) && (v >= 0) && (u + v < 1)) return t;
  return -72;
}
float or_ray_triangle(const OrScene *s, const float o[3], const float d[3], int face) {
  return ray_triangle(s, ld3(o), ld3(d), face);
}

/* Analytic sphere (NOT in the reference; conventions chosen to match the triangle path:
 * un-normalised d, smallest root > 1e-5).  Returns -72 on miss. */
static float ray_sphere(const OrScene *s, v3 o, v3 d, int si) {
  const float *sp = s->spheres + (size_t)si * 4;
  v3 oc = sub(o, ld3(sp));
  float r = sp[3];
  float a = dot(d, d), hb = dot(oc, d), cc = dot(oc, oc) - r * r;
  float disc = hb * hb - a * cc;
  if (!(disc >= 0.f) || a == 0.f) return -72;
  float sq = sqrtf(disc);
  float t0 = (-hb - sq) / a, t1 = (-hb + sq) / a;
  if (t0 > 0.00001f) return t0;
  if (t1 > 0.00001f) return t1;
  return -72;
}

/* ------------------------------------------------------------------------------------------ */
/* nearest hit, src/flyscene.cpp:672-683 (ascending face id, strict <  => lowest id wins ties) */
/* ------------------------------------------------------------------------------------------ */
typedef struct { const OrScene *s; v3 o, d; float t; int best; } NearCtx;
static void nearest_cb(void *ctx, int face) {
  NearCtx *c = (NearCtx *)ctx;
  float is = ray_triangle(c->s, c->o, c->d, face);
  if (is != -72 && is > 0.00001f) {
    if (is < c->t || (is == c->t && face < c->best)) { c->t = is; c->best = face; }
  }
}
static int nearest_hit(const OrScene *s, v3 o, v3 d, float *t_out) {
  NearCtx c = {s, o, d, FLT_MAX, -1};
  float of[3] = {o.x, o.y, o.z};
  v3 de = add(d, o); /* octree.intersect(origin, direction+origin), :672 */
  float df[3] = {de.x, de.y, de.z};
  for_each_candidate(s, of, df, nearest_cb, &c);
  for (int si = 0; si < s->d.n_spheres; ++si) {
    float is = ray_sphere(s, o, d, si);
    if (is != -72 && is > 0.00001f && is < c.t) { c.t = is; c.best = s->d.n_faces + si; }
  }
  *t_out = c.t;
  return c.best;
}

/* ------------------------------------------------------------------------------------------ */
/* a8  Flyscene::lightStrikes, src/flyscene.cpp:912-954                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { const OrScene *s; v3 o, d; float t; } ShadowCtx;
static void shadow_cb(void *ctx, int face) {
  ShadowCtx *c = (ShadowCtx *)ctx;
  if (c->s->mats[c->s->mat_id[face]].illum == 9) return; /* :934-936 */
  float is = ray_triangle(c->s, c->o, c->d, face);
  if (is != -72 && is < c->t && (double)is > 0.00001) c->t = is;
}
static int light_strikes(const OrScene *s, v3 hit, const float *lights, int n, uint8_t *visible, int64_t *census) {
  int any = 0;
  for (int l = 0; l < n; ++l) {
    ShadowCtx c = {s, ld3(lights + 3 * l), V(0, 0, 0), FLT_MAX};
    c.d = sub(hit, c.o);
    float of[3] = {c.o.x, c.o.y, c.o.z}, hf[3] = {hit.x, hit.y, hit.z};
    if (census) census[1]++;
    if (or_box_intersect(s->nodes[0].mn, s->nodes[0].mx, of, hf)) {
      for_each_candidate(s, of, hf, shadow_cb, &c);
    }
    for (int si = 0; si < s->d.n_spheres; ++si) {
      if (s->mats[s->sphere_mat[si]].illum == 9) continue;
      float is = ray_sphere(s, c.o, c.d, si);
      if (is != -72 && is < c.t && (double)is > 0.00001) c.t = is;
    }
    if ((double)c.t >= 0.98) { any = 1; visible[l] = 1; }
    else visible[l] = 0;
  }
  return any;
}
int or_light_strikes(const OrScene *s, const float hit[3], const float *lights, int n, uint8_t *visible) {
  return light_strikes(s, ld3(hit), lights, n, visible, NULL);
}

/* ------------------------------------------------------------------------------------------ */
/* a12 createSpherePoint / createAreaLight / arealight::getPointLights                         */
/*     src/flyscene.cpp:956-972, arealight.hpp:15-25.  The random "spherical" mode (:974-995)  */
/*     seeds from std::random_device per point and is not reproducible: its 25 draws are       */
/*     replaced by a counter hash (rt_oracle.h, OrParams.sphere_seed); everything after the     */
/*     draw follows :981-990 (float randomno/theta/phi/x/y/z, double 2.0f*M_PI*u and acos,      */
/*     float sin/cos overloads, Vector3f(x,y,z)/5 + lightPoint).  PARITY UNPINNED for this mode.*/
/* ------------------------------------------------------------------------------------------ */
static uint32_t hash_lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
int or_light_samples(const OrParams *p, const float light[3], float *out) {
  if (p->point_light) { out[0] = light[0]; out[1] = light[1]; out[2] = light[2]; return 1; }
  if (p->area_light) {
    v3 c = ld3(light);
    v3 uvec = add(c, mul(p->area_len_x, V(1, 0, 0)));
    v3 vvec = add(c, mul(p->area_len_y, V(0, 1, 0)));
    int k = 0;
    for (int i = 0; i < p->usteps; ++i)
      for (int j = 0; j < p->vsteps; ++j) {
        /* ((i + 0.5) * (uvec/usteps)).x(): int->float division, double literal demoted to float
         * by Eigen's scalar promotion (the product is evaluated in float) */
        float sx = (float)(i + 0.5) * (uvec.x / (float)p->usteps);
        float sy = (float)(j + 0.5) * (vvec.y / (float)p->vsteps);
        out[3 * k] = sx; out[3 * k + 1] = sy; out[3 * k + 2] = uvec.z;
        ++k;
      }
    return k;
  }
  for (int k = 0; k < 25; ++k) {
    uint32_t h = hash_lowbias32(p->sphere_seed * 0x9E3779B9U + (uint32_t)k);
    float randomno = (float)(h >> 8) * (1.0f / 16777216.0f);
    float theta = (float)(2.0f * M_PI * randomno);
    float phi = (float)acos(2.0 * randomno - 1.0);
    float x = p->sphere_radius * sinf(phi) * cosf(theta);
    float y = p->sphere_radius * sinf(phi) * sinf(theta);
    float z = p->sphere_radius * cosf(phi);
    out[3 * k] = x / 5 + light[0]; out[3 * k + 1] = y / 5 + light[1]; out[3 * k + 2] = z / 5 + light[2];
  }
  return 25;
}

/* ------------------------------------------------------------------------------------------ */
/* a10 getInterpolatedNormal, src/flyscene.cpp:864-888; sphere: (hit - c)                      */
/* ------------------------------------------------------------------------------------------ */
static v3 interpolated_normal(const OrScene *s, v3 p, int face) {
  if (face >= s->d.n_faces) {
    const float *sp = s->spheres + (size_t)(face - s->d.n_faces) * 4;
    return sub(p, ld3(sp));
  }
  const float *vp = s->verts + (size_t)face * 9;
  v3 a = ld3(vp), b = ld3(vp + 3), c = ld3(vp + 6);
  v3 v0 = sub(b, a), v1 = sub(c, a), v2 = sub(p, a);
  const float *np = s->vnormals + (size_t)face * 9;
  v3 nA = ld3(np), nB = ld3(np + 3), nC = ld3(np + 6);
  float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
  float denom = d00 * d11 - d01 * d01;
  float v = (d11 * d20 - d01 * d21) / denom;
  float w = (d00 * d21 - d01 * d20) / denom;
  float u = 1.0f - v - w;
  return add(add(mul(u, nA), mul(v, nB)), mul(w, nC));
}

static const OrMaterial *face_material(const OrScene *s, int face) {
  if (face >= s->d.n_faces) return &s->mats[s->sphere_mat[face - s->d.n_faces]];
  return &s->mats[s->mat_id[face]];
}
static v3 face_normal(const OrScene *s, int face, v3 hit) {
  if (face >= s->d.n_faces) {
    const float *sp = s->spheres + (size_t)(face - s->d.n_faces) * 4;
    return normalized(sub(hit, ld3(sp)));
  }
  return ld3(s->fnormals + (size_t)face * 3);
}

/* ------------------------------------------------------------------------------------------ */
/* a9  Flyscene::phongShade, src/flyscene.cpp:822-859                                          */
/* ------------------------------------------------------------------------------------------ */
static v3 phong_shade(const OrScene *s, v3 origin, v3 hit, int face, const float *lightsp, int n_lights, int64_t *census) {
  v3 I = ld3(s->p.light_color);
  const OrMaterial *m = face_material(s, face);
  v3 kd = ld3(m->kd), ks = ld3(m->ks);
  v3 final = V(0, 0, 0);
  v3 normal = normalized(affine_point(s->d.model_matrix, interpolated_normal(s, hit, face)));
  float samples[25 * 3];
  uint8_t visible[25];
  for (int l = 0; l < n_lights; ++l) {
    float sum = 0;
    v3 colour = V(0, 0, 0);
    int ns = or_light_samples(&s->p, lightsp + 3 * l, samples);
    /* census convention (BASELINE.json north_star): in point mode the sample ray of a light is the
     * very same query as its gate ray (:699 vs :836), so it is counted once */
    light_strikes(s, hit, samples, ns, visible, s->p.point_light ? NULL : census);
    for (int i = 0; i < ns; ++i) {
      if (!visible[i]) continue;
      sum++;
      v3 ldir = normalized(sub(ld3(samples + 3 * i), hit));
      float costheta = fmax_std(0.0f, dot(ldir, normal));
      v3 diffuse = mul(costheta, cmul(I, kd)); /* (I.cwiseProduct(kd)) * costheta */
      v3 refl = normalized(sub(ldir, mul(2 * dot(ldir, normal), normal)));
      v3 eye = normalized(mul(-1.f, sub(hit, origin)));
      float cosphi = fmax_std(0.0f, dot(eye, mul(-1.f, refl)));
      v3 specular = mul(powf(cosphi, m->ns), cmul(I, ks));
      colour = add(colour, add(diffuse, specular));
    }
    float a = sum / (float)ns, b = 1.3f / (float)ns;
    final = add(final, mul(b, mul(a, colour))); /* colour*(sum/S) * (1.3f/S) */
  }
  return final;
}

void or_phong_shade(const OrScene *s, const float origin[3], const float hit[3], int face, const float *lights,
                    int n_lights, float rgb[3]) {
  v3 c = phong_shade(s, ld3(origin), ld3(hit), face, lights, n_lights, NULL);
  st3(rgb, c);
}

/* a11 Flyscene::fresnel, src/flyscene.cpp:890-910 */
static float fresnel(v3 I, v3 N, float ior) {
  float cosi = dot(I, N);
  float etai = 1, etat = ior;
  if (cosi > 0) { float tmp = etai; etai = etat; etat = tmp; }
  float sint = etai / etat * sqrtf(fmax_std(0.f, 1 - cosi * cosi));
  if (sint >= 1) return 1;
  float cost = sqrtf(fmax_std(0.f, 1 - sint * sint));
  cosi = fabsf(cosi);
  float Rs = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
  float Rp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
  return (Rs * Rs + Rp * Rp) / 2;
}

/* ------------------------------------------------------------------------------------------ */
/* a3  Flyscene::traceRay, src/flyscene.cpp:651-771                                            */
/* ------------------------------------------------------------------------------------------ */
static v3 refracted(v3 d, v3 n, float ni) {
  /* :722-724 / :747-749 -- c1 float, c2 through double pow/sqrt */
  float c1 = fabsf(dot(d, n));
  float inv = 1 / ni;
  double p1 = (double)inv * (double)inv;  /* pow(float,int) -> double, exact square */
  double p2 = (double)c1 * (double)c1;
  float c2 = (float)sqrt(1 - p1 * (1 - p2));
  return add(mul(inv, d), mul(inv * c1 - c2, n));
}

static v3 trace_ray(const OrScene *s, v3 o, v3 d, int level, const float *lights, int n_lights,
                    int32_t *face_out, float *t_out, int64_t *census) {
  const v3 BACKGROUND = V(1.f, 1.f, 1.f), SHADOW = V(0.f, 0.f, 0.f);
  if (face_out) *face_out = -1;
  if (t_out) *t_out = FLT_MAX;
  if (level > s->p.recursion_guard) return BACKGROUND;
  float of[3] = {o.x, o.y, o.z};
  v3 de = add(o, d);
  float df[3] = {de.x, de.y, de.z};
  if (census) census[level == 0 ? 0 : 2]++;
  if (!or_box_intersect(s->nodes[0].mn, s->nodes[0].mx, of, df)) {
    if (s->d.n_spheres == 0) return BACKGROUND; /* :655-665 */
  }
  float t;
  int best;
  if (!or_box_intersect(s->nodes[0].mn, s->nodes[0].mx, of, df)) {
    /* spheres live outside the triangle root box: test them alone */
    t = FLT_MAX; best = -1;
    for (int si = 0; si < s->d.n_spheres; ++si) {
      float is = ray_sphere(s, o, d, si);
      if (is != -72 && is > 0.00001f && is < t) { t = is; best = s->d.n_faces + si; }
    }
  } else {
    best = nearest_hit(s, o, d, &t);
  }
  if (best == -1) return BACKGROUND; /* :684-691 */
  if (face_out) *face_out = best;
  if (t_out) *t_out = t;

  v3 hit = add(o, mul(t, d)); /* :695 */
  v3 fn = face_normal(s, best, hit);

  uint8_t vis[25];
  if (!light_strikes(s, hit, lights, n_lights, vis, census)) return SHADOW; /* :699-710 */

  const OrMaterial *m = face_material(s, best);
  int imodel = m->illum;
  if (s->p.max_depth >= 0 && level >= s->p.max_depth) imodel = 2; /* depth cap (not in reference) */
  float fresnelIndex = 1;
  int have = 0; /* Color != (-1,-1,-1) */
  v3 Color = V(-1, -1, -1);

  if (imodel == 9) { /* :717-719 */
    v3 ph = phong_shade(s, o, hit, best, lights, n_lights, census);
    v3 ch = trace_ray(s, hit, d, level + 1, lights, n_lights, NULL, NULL, census);
    Color = add(mul(0.10f, ph), mul(0.90f, ch));
    have = 1;
  } else if (imodel == 6 || imodel == 7) { /* :721-731 */
    v3 rr = refracted(d, fn, m->ni);
    v3 ch = trace_ray(s, hit, rr, level + 1, lights, n_lights, NULL, NULL, census);
    if (imodel == 7) Color = add(mul(fresnelIndex, Color), mul(1 - fresnelIndex, ch));
    else Color = ch;
    have = 1;
  } else if (imodel > 2 && imodel < 7) { /* :733-744 */
    v3 refl = sub(d, mul(2 * dot(d, fn), fn));
    float rl[3] = {hit.x, hit.y, hit.z};
    v3 ph = phong_shade(s, o, hit, best, lights, n_lights, census);
    v3 ch = trace_ray(s, hit, refl, level + 1, rl, 1, NULL, NULL, census);
    Color = add(mul(0.15f, ph), mul(0.85f, ch));
    have = 1;
    if (imodel == 5) {
      fresnelIndex = fresnel(refl, fn, m->ni);
      return mul(fresnelIndex, Color);
    }
  }
  if (imodel == 6 || imodel == 7) { /* :746-756 */
    v3 rr = refracted(d, fn, m->ni);
    v3 ch = trace_ray(s, hit, rr, level + 1, lights, n_lights, NULL, NULL, census);
    if (imodel == 7) Color = add(mul(fresnelIndex, Color), mul(1 - fresnelIndex, ch));
    else {
      v3 ph = phong_shade(s, o, hit, best, lights, n_lights, census);
      Color = add(mul(0.2f, ph), mul(0.8f, ch));
    }
    have = 1;
  }
  (void)have;
  if (Color.x == -1 && Color.y == -1 && Color.z == -1) /* :758-760 */
    Color = phong_shade(s, o, hit, best, lights, n_lights, census);
  return Color;
}

void or_trace_ray(const OrScene *s, const float o[3], const float d[3], int level, const float *lights,
                  int n_lights, float rgb[3], int32_t *face, float *t) {
  v3 c = trace_ray(s, ld3(o), ld3(d), level, lights, n_lights, face, t, NULL);
  st3(rgb, c);
}

/* ------------------------------------------------------------------------------------------ */
/* a2  Camera::screenToWorld, tucano/camera.hpp:155-173; getPerspectiveScale :263-266          */
/* ------------------------------------------------------------------------------------------ */
void or_screen_to_world(const OrCamera *cam, float i, float j, float out[3]) {
  float nx = (float)(2.0 * (double)(i - cam->viewport[0]) / (double)cam->viewport[2] - 1.0);
  float ny = (float)(1.0 - 2.0 * (double)(j - cam->viewport[1]) / (double)cam->viewport[3]);
  float nz = -1.0f;
  float pscale = (float)((double)1.0f / tan((double)(cam->fovy / 2.0f) * (M_PI / (double)180.0f)));
  float scale = (float)(1.0 / (double)pscale);
  nx *= cam->aspect * scale;
  ny *= scale;
  v3 w = affine_point(cam->view_inv, V(nx, ny, nz));
  st3(out, w);
}

/* a13 ppmIO.hpp:145: min(255, (int)(255*c)) -- float multiply, truncation, no lower clamp */
int or_quantize(float c) {
  float v = 255 * c;
  int q;
  if (v != v) q = (int)0x80000000; /* x86 cvttss2si of NaN */
  else if (v >= 2147483648.0f || v < -2147483648.0f) q = (int)0x80000000;
  else q = (int)v;
  return q < 255 ? q : 255;
}

/* ------------------------------------------------------------------------------------------ */
/* a1  per-pixel loop of Flyscene::raytraceScene, src/flyscene.cpp:573-598, 613-625            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const OrScene *s; const OrCamera *cam; const float *lights; int n_lights;
  const int32_t *px, *py; int64_t n; float *rgb; int32_t *face; float *t; uint8_t *rgb8;
  int64_t *cursor; pthread_mutex_t *mtx;
} Job;

static void *render_worker(void *arg) {
  Job *j = (Job *)arg;
  int64_t census[3] = {0, 0, 0};
  v3 eye = ld3(j->cam->eye);
  for (;;) {
    pthread_mutex_lock(j->mtx);
    int64_t k0 = *j->cursor; *j->cursor += 64;
    pthread_mutex_unlock(j->mtx);
    if (k0 >= j->n) break;
    int64_t k1 = k0 + 64 < j->n ? k0 + 64 : j->n;
    for (int64_t k = k0; k < k1; ++k) {
      float screen[3];
      or_screen_to_world(j->cam, (float)j->px[k], (float)j->py[k], screen);
      v3 colour = V(1.f, 1.f, 1.f);
      int32_t face = -1; float t = FLT_MAX;
      int hit_box = or_box_intersect(j->s->nodes[0].mn, j->s->nodes[0].mx, j->cam->eye, screen); /* :576 */
      if (hit_box || j->s->d.n_spheres > 0) {
        v3 dir = sub(ld3(screen), eye); /* :619 */
        colour = trace_ray(j->s, eye, dir, 0, j->lights, j->n_lights, &face, &t, census);
      } else {
        census[0]++;
      }
      if (j->rgb) st3(j->rgb + 3 * k, colour);
      if (j->face) j->face[k] = face;
      if (j->t) j->t[k] = t;
      if (j->rgb8) {
        int q[3] = {or_quantize(colour.x), or_quantize(colour.y), or_quantize(colour.z)};
        for (int c = 0; c < 3; ++c) j->rgb8[3 * k + c] = (uint8_t)(q[c] < 0 ? 0 : q[c]);
      }
    }
  }
  pthread_mutex_lock(&g_census_mtx);
  for (int c = 0; c < 3; ++c) g_census[c] += census[c];
  pthread_mutex_unlock(&g_census_mtx);
  return NULL;
}

void or_render_pixels(const OrScene *s, const OrCamera *cam, const float *lights, int n_lights,
                      const int32_t *px, const int32_t *py, int64_t n, float *rgb, int32_t *face,
                      float *t, uint8_t *rgb8, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  int64_t cursor = 0;
  pthread_mutex_t mtx = PTHREAD_MUTEX_INITIALIZER;
  Job job = {s, cam, lights, n_lights, px, py, n, rgb, face, t, rgb8, &cursor, &mtx};
  pthread_t th[256];
  for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, render_worker, &job);
  for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
}

void or_census_reset(void) {
  pthread_mutex_lock(&g_census_mtx);
  g_census[0] = g_census[1] = g_census[2] = 0;
  pthread_mutex_unlock(&g_census_mtx);
}
void or_census_get(int64_t out[3]) {
  pthread_mutex_lock(&g_census_mtx);
  out[0] = g_census[0]; out[1] = g_census[1]; out[2] = g_census[2];
  pthread_mutex_unlock(&g_census_mtx);
}
