/* oracle/tpt_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or executed
 * from the product path; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it).
 *
 * A plain-C restatement of the reference's hot path (BlurryLight/tiny-path-tracer), operating on
 * the flattened scene arrays of include/tpt.h. Each function cites the reference lines it follows.
 * It is deliberately the NAIVE algorithm: recursive color() that keeps bouncing NaN "zombie" paths
 * and zero-weight paths to max_depth exactly like the reference, recursive hit() with the
 * reference's t_max handling and tie rules. The CUDA path takes exact shortcuts (early
 * termination, forward throughput); this file does not, so agreement between the two checks those
 * shortcuts as well.
 *
 * PARITY PINNED: tests/test_oracle_port.py checks this restatement bit for bit against the
 * reference itself (oracle/_ref, compiled from /root/reference) -- hit records and per-sample
 * radiance under the injected Philox stream -- and against the committed golden fixtures.
 *
 * Arithmetic follows the C++ expressions' promotions: unqualified sqrt()/pow()/M_PI/drand_r()
 * expressions are double, std::sqrt/std::sin/std::cos/std::atan2/std::asin on floats are the
 * float libm functions. Build: gcc -std=c11 -O2 -ffp-contract=off (oracle/Makefile `port`).
 */
#define _GNU_SOURCE
#include "tpt.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* std::sin / std::cos of a float are the host libm's sinf / cosf: faithful (< 1 ulp) but not correctly rounded, and the
 * last bit depends on the libm build. The CUDA parity path evaluates them in double and rounds once. Diagnostic switch
 * (default 0 = the reference's calls): with tpto_set_rounded_trig(1) this restatement does the same, which attributes a
 * residual pixel difference to that last bit or to something else (tools/gpu_diff_fuzz.py, profiles/r02_fuzz.txt). */
static int g_rounded_trig = 0;
void tpto_set_rounded_trig(int on) { g_rounded_trig = on; }
#define SINF(x) (g_rounded_trig ? (float)sin((double)(x)) : sinf(x))
#define COSF(x) (g_rounded_trig ? (float)cos((double)(x)) : cosf(x))


#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { float e[3]; } v3;
typedef struct { v3 o, d; float time; } ray_t;
typedef struct { float t, u, v; v3 p, n; int mat, prim; } rec_t; /* hit_record, headers/hitable.h:14-21 */

/* ---- headers/vec3.h:54-84,143-144 ---- */
static v3 V(float a, float b, float c) { v3 r = {{a, b, c}}; return r; }
static v3 vadd(v3 a, v3 b) { return V(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
static v3 vsub(v3 a, v3 b) { return V(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
static v3 vmul(v3 a, v3 b) { return V(a.e[0] * b.e[0], a.e[1] * b.e[1], a.e[2] * b.e[2]); }
static v3 vscale(v3 a, float s) { return V(a.e[0] * s, a.e[1] * s, a.e[2] * s); }
static v3 vdivs(v3 a, float s) { return V(a.e[0] / s, a.e[1] / s, a.e[2] / s); }
static v3 vneg(v3 a) { return V(-a.e[0], -a.e[1], -a.e[2]); }
static float vdot(v3 a, v3 b) { return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2]; }
static v3 vcross(v3 a, v3 b) {
  return V(a.e[1] * b.e[2] - b.e[1] * a.e[2], a.e[2] * b.e[0] - b.e[2] * a.e[0], a.e[0] * b.e[1] - b.e[0] * a.e[1]);
}
static float vsqlen(v3 a) { return a.e[0] * a.e[0] + a.e[1] * a.e[1] + a.e[2] * a.e[2]; }
static float vlen(v3 a) { return sqrtf(vsqlen(a)); }                 /* std::sqrt(float) */
static v3 vunit(v3 a) { return vdivs(a, vlen(a)); }
static v3 point_at(const ray_t *r, float t) { return vadd(r->o, vscale(r->d, t)); } /* headers/ray.h:13 */

/* ---- injected stream: identical to oracle/ref_inject.cc and the device Rng ---- */
typedef struct { uint32_t key[2], pixel, sample, stage, ndraw, block_id, block[4]; uint64_t total; } rng_t;

static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void tpto_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }

static double drand_r(rng_t *g) { /* stands in for src/utils.cc:28-32 under the injected stream */
  uint32_t blk = g->ndraw >> 2;
  if (blk != g->block_id) {
    uint32_t ctr[4] = {g->pixel, g->sample, g->stage, blk};
    philox4x32_10(ctr, g->key, g->block);
    g->block_id = blk;
  }
  uint32_t x = g->block[g->ndraw & 3];
  g->ndraw++;
  g->total++;
  return (double)(x >> 8) * (1.0 / 16777216.0);
}
static void rng_next_stage(rng_t *g) { g->stage++; g->ndraw = 0; g->block_id = 0xffffffffu; }

/* ---- scene access ---- */
typedef struct {
  const tpt_scene_desc *d;
  rng_t *g;
  uint64_t rays;
} ctx_t;

static int node_kind(const tpt_node *n) { return n->kind & 0xff; }
static int node_chain(const tpt_node *n) { return n->kind >> 16; }

/* ---- src/aabb.cc:3-19 ---- */
static int aabb_hit(const float mn[3], const float mx[3], const ray_t *r, float tmin, float tmax) {
  for (int i = 0; i < 3; i++) {
    float inv_D = 1.0f / r->d.e[i];
    float t0 = (mn[i] - r->o.e[i]) * inv_D;
    float t1 = (mx[i] - r->o.e[i]) * inv_D;
    if (inv_D < 0.0f) { float t = t0; t0 = t1; t1 = t; }
    tmin = t0 > tmin ? t0 : tmin;
    tmax = t1 < tmax ? t1 : tmax;
    if (tmin > tmax) return 0;
  }
  return 1;
}

/* ---- headers/utils.h:38-43 ---- */
static void get_uv_map(v3 p, float *u, float *v) {
  *u = (float)(atan2f(p.e[2], p.e[0]) / (2 * M_PI));
  *v = (float)(asinf(p.e[1]) / M_PI);
  *u += 0.5f;
  *v += 0.5f;
}

/* ---- src/sphere.cc:13-75 ---- */
static int sphere_hit_c(v3 center, v3 uv_center, float radius, const ray_t *r, float t_min, float t_max, rec_t *rec) {
  v3 oc = vsub(r->o, center);
  float a = vdot(r->d, r->d);
  float b = (float)(2.0 * vdot(r->d, oc));
  float c = vdot(oc, oc) - radius * radius;
  float discriminant = b * b - 4 * a * c;
  if (discriminant > 0) {
    float temp = (float)((-b - sqrt(b * b - 4 * a * c)) / (2 * a));
    if (temp < t_max && temp > t_min) {
      rec->t = temp;
      rec->p = point_at(r, temp);
      get_uv_map(vdivs(vsub(rec->p, uv_center), radius), &rec->u, &rec->v);
      rec->n = vdivs(vsub(rec->p, center), radius);
      return 1;
    }
    temp = (float)((-b + sqrt(b * b - 4 * a * c)) / (2 * a));
    if (temp < t_max && temp > t_min) {
      rec->t = temp;
      rec->p = point_at(r, temp);
      get_uv_map(vdivs(vsub(rec->p, uv_center), radius), &rec->u, &rec->v);
      rec->n = vdivs(vsub(rec->p, center), radius);
      return 1;
    }
  }
  return 0;
}

/* ---- src/rect_box.cc:8-24,51-68,75-91 : ka = plane axis, a/b = in-plane axes ---- */
static int rect_hit_c(int ka, int aa, int ba, const float p[5], const ray_t *r, float t_min, float t_max, rec_t *rec) {
  float t = (p[4] - r->o.e[ka]) / r->d.e[ka];
  if (t > t_max || t < t_min) return 0;
  float a = r->o.e[aa] + t * r->d.e[aa];
  float b = r->o.e[ba] + t * r->d.e[ba];
  if (a < p[0] || a > p[1] || b < p[2] || b > p[3]) return 0;
  rec->u = (a - p[0]) / (p[1] - p[0]);
  rec->v = (b - p[2]) / (p[3] - p[2]);
  rec->t = t;
  rec->p = point_at(r, t);
  rec->n = V(ka == 0 ? 1.f : 0.f, ka == 1 ? 1.f : 0.f, ka == 2 ? 1.f : 0.f);
  return 1;
}

static int prim_hit(const tpt_scene_desc *d, int id, const ray_t *r, float t_min, float t_max, rec_t *rec) {
  const tpt_prim *q = &d->prims[id];
  int ok = 0;
  switch (q->kind) {
  case TPT_PRIM_SPHERE: {
    v3 c = V(q->p[0], q->p[1], q->p[2]);
    ok = sphere_hit_c(c, c, q->p[3], r, t_min, t_max, rec);
    break;
  }
  case TPT_PRIM_MOVING_SPHERE: { /* headers/sphere.h:28-31: center(time); uv uses center0_ */
    v3 c0 = V(q->p[0], q->p[1], q->p[2]), c1 = V(q->p[4], q->p[5], q->p[6]);
    float f = (r->time - q->p[7]) / (q->p[8] - q->p[7]);
    v3 c = vadd(c0, vscale(vsub(c1, c0), f));
    ok = sphere_hit_c(c, c0, q->p[3], r, t_min, t_max, rec);
    break;
  }
  case TPT_PRIM_XY_RECT: ok = rect_hit_c(2, 0, 1, q->p, r, t_min, t_max, rec); break;
  case TPT_PRIM_XZ_RECT: ok = rect_hit_c(1, 0, 2, q->p, r, t_min, t_max, rec); break;
  case TPT_PRIM_YZ_RECT: ok = rect_hit_c(0, 1, 2, q->p, r, t_min, t_max, rec); break;
  default: break;
  }
  if (ok) {
    rec->mat = q->material;
    rec->prim = id;
    if (q->flags & TPT_PRIM_FLIP) rec->n = vneg(rec->n); /* headers/rect_box.h:50-57 */
  }
  return ok;
}

static int node_hit(ctx_t *cx, int i, const ray_t *r_in, int from_nops, float t_min, float t_max, rec_t *rec);

/* constant_medium::hit src/hitable.cc:98-128. `r` is the ray in the medium's own space; the
 * boundary is the sub-tree stored at nodes [first, end) (its chains extend the medium's). */
static int medium_hit(ctx_t *cx, int id, const ray_t *r, int from_nops, float t_min, float t_max, rec_t *rec) {
  const tpt_prim *q = &cx->d->prims[id];
  if (!cx->g) return 0; /* hit batches carry no stream */
  int32_t first;
  memcpy(&first, &q->p[1], 4);
  rec_t rec1, rec2;
  if (node_hit(cx, first, r, from_nops, -FLT_MAX, FLT_MAX, &rec1)) {
    if (node_hit(cx, first, r, from_nops, (float)(rec1.t + 0.0001), FLT_MAX, &rec2)) {
      if (rec1.t < t_min) rec1.t = t_min;
      if (rec2.t > t_max) rec2.t = t_max;
      if (rec1.t >= rec2.t) return 0;
      float distance_inside_boundary = (rec2.t - rec1.t) * vlen(r->d);
      float hit_distance = (float)((-1 / q->p[0]) * log(drand_r(cx->g)));
      if (hit_distance < distance_inside_boundary) {
        rec->t = rec1.t + hit_distance / vlen(r->d);
        rec->p = point_at(r, rec->t);
        float z = (float)drand_r(cx->g), y = (float)drand_r(cx->g), x = (float)drand_r(cx->g); /* right to left */
        v3 n = V(x, y, z);
        float k = (float)(1.0 / vlen(n)); /* vec3::make_unit_vector headers/vec3.h:95-100 */
        rec->n = vscale(n, k);
        rec->u = rec->v = 0;
        rec->mat = q->material;
        rec->prim = id;
        return 1;
      }
    }
  }
  return 0;
}

/* the hitable at node i, seen from a ray expressed in the space of chain `from` */
static int node_hit(ctx_t *cx, int i, const ray_t *r_in, int from_nops, float t_min, float t_max, rec_t *rec) {
  const tpt_scene_desc *d = cx->d;
  const tpt_node *nd = &d->nodes[i];
  const tpt_chain *ch = &d->chains[node_chain(nd)];
  /* wrappers between the parent's space and this node: translate::hit headers/rect_box.h:87-95,
   * rotate_y::hit src/rect_box.cc:171-195 (outermost first) */
  ray_t r = *r_in;
  for (int k = from_nops; k < ch->n_ops; k++) {
    const tpt_xform_op *op = &d->xform_ops[ch->first_op + k];
    if (op->kind == TPT_XF_TRANSLATE) {
      r.o = vsub(r.o, V(op->a, op->b, op->c));
    } else {
      float s = op->a, co = op->b;
      v3 o = r.o, dd = r.d;
      o.e[0] = co * r.o.e[0] - s * r.o.e[2];
      o.e[2] = s * r.o.e[0] + co * r.o.e[2];
      dd.e[0] = co * r.d.e[0] - s * r.d.e[2];
      dd.e[2] = s * r.d.e[0] + co * r.d.e[2];
      r.o = o;
      r.d = dd;
    }
  }
  int hit = 0;
  int kind = node_kind(nd);
  if (kind == TPT_NODE_LEAF) {
    if (d->prims[nd->end_or_prim].kind == TPT_PRIM_MEDIUM)
      hit = medium_hit(cx, nd->end_or_prim, &r, ch->n_ops, t_min, t_max, rec);
    else
      hit = prim_hit(d, nd->end_or_prim, &r, t_min, t_max, rec);
  } else if (kind == TPT_NODE_BVH) { /* src/hitable.cc:63-90 */
    if (aabb_hit(nd->bmin, nd->bmax, &r, t_min, t_max)) {
      int left = i + 1;
      int right = node_kind(&d->nodes[left]) == TPT_NODE_LEAF ? left + 1 : d->nodes[left].end_or_prim;
      rec_t lr, rr;
      int hl = node_hit(cx, left, &r, ch->n_ops, t_min, t_max, &lr);
      int hr = node_hit(cx, right, &r, ch->n_ops, t_min, t_max, &rr);
      if (hl && hr) { *rec = lr.t < rr.t ? lr : rr; hit = 1; }
      else if (hl) { *rec = lr; hit = 1; }
      else if (hr) { *rec = rr; hit = 1; }
    }
  } else { /* hitable_list::hit src/hitable_list.cc:38-51 */
    rec_t tmp;
    double closest_so_far = t_max;
    int j = i + 1, end = nd->end_or_prim;
    while (j < end) {
      if (node_hit(cx, j, &r, ch->n_ops, t_min, (float)closest_so_far, &tmp)) {
        hit = 1;
        closest_so_far = tmp.t;
        *rec = tmp;
      }
      j = node_kind(&d->nodes[j]) == TPT_NODE_LEAF ? j + 1 : d->nodes[j].end_or_prim;
    }
  }
  if (hit) { /* unwind the wrappers, innermost first */
    for (int k = ch->n_ops - 1; k >= from_nops; k--) {
      const tpt_xform_op *op = &d->xform_ops[ch->first_op + k];
      if (op->kind == TPT_XF_TRANSLATE) {
        rec->p = vadd(rec->p, V(op->a, op->b, op->c));
      } else {
        float s = op->a, co = op->b;
        v3 p = rec->p, n = rec->n;
        p.e[0] = co * rec->p.e[0] + s * rec->p.e[2];
        p.e[2] = -s * rec->p.e[0] + co * rec->p.e[2];
        n.e[0] = co * rec->n.e[0] + s * rec->n.e[2];
        n.e[2] = -s * rec->n.e[0] + co * rec->n.e[2];
        rec->p = p;
        rec->n = n;
      }
    }
  }
  return hit;
}

static int world_hit(ctx_t *cx, const ray_t *r, float t_min, float t_max, rec_t *rec) {
  return node_hit(cx, 0, r, 0, t_min, t_max, rec);
}

/* ---- textures: src/texture.cc:4-42, perlin src/utils.cc:160-225 ---- */
static float perlin_interp(const v3 c[2][2][2], float u, float v, float w) {
  float uu = (float)(6 * pow(u, 5) - 15 * pow(u, 4) + 10 * pow(u, 3));
  float vv = (float)(6 * pow(v, 5) - 15 * pow(v, 4) + 10 * pow(v, 3));
  float ww = (float)(6 * pow(w, 5) - 15 * pow(w, 4) + 10 * pow(w, 3));
  float accum = 0.0f;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++) {
        v3 wv = V(uu - i, vv - j, ww - k);
        accum += (i * uu + (1 - i) * (1 - uu)) * (j * vv + (1 - j) * (1 - vv)) * (k * ww + (1 - k) * (1 - ww)) *
                 vdot(c[i][j][k], wv);
      }
  return fabsf(accum);
}
static float perlin_noise_at(const tpt_perlin_tables *T, v3 p) {
  float u = p.e[0] - floorf(p.e[0]), v = p.e[1] - floorf(p.e[1]), w = p.e[2] - floorf(p.e[2]);
  int i = ((int)floorf(p.e[0])) & 255, j = ((int)floorf(p.e[1])) & 255, k = ((int)floorf(p.e[2])) & 255;
  v3 c[2][2][2];
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 2; b++)
      for (int cc = 0; cc < 2; cc++) {
        int idx = T->perm_x[(i + a) & 255] ^ T->perm_y[(j + b) & 255] ^ T->perm_z[(k + cc) & 255];
        c[a][b][cc] = V(T->ranvec[idx][0], T->ranvec[idx][1], T->ranvec[idx][2]);
      }
  return perlin_interp(c, u, v, w);
}
static float perlin_turb(const tpt_perlin_tables *T, v3 p) {
  float accum = 0, weight = 1.0f;
  v3 tmp = p;
  for (int i = 0; i < 5; i++) {
    accum += weight * perlin_noise_at(T, tmp);
    weight *= 0.5f;
    tmp = vscale(tmp, 2);
  }
  return fabsf(accum);
}
static v3 texture_value(const tpt_scene_desc *d, int id, float u, float v, v3 p) {
  const tpt_texture *t = &d->textures[id];
  switch (t->kind) {
  case TPT_TEX_CONSTANT: return V(t->color[0], t->color[1], t->color[2]);
  case TPT_TEX_CHECKER: {
    float s = SINF(10 * p.e[0]) * SINF(10 * p.e[1]) * SINF(10 * p.e[2]);
    return texture_value(d, isless(s, 0.0f) ? t->odd : t->even, u, v, p);
  }
  case TPT_TEX_PERLIN: {
    float g = 0.5f * (1 + SINF(t->scale * p.e[2] + 10 * perlin_turb(d->perlin, p)));
    return V(g, g, g); /* vec3(1,1,1) * 0.5 * (1 + sin(...)) */
  }
  default: {
    const tpt_image_desc *im = &d->images[t->image];
    int i = (int)(u * im->width), j = (int)((1 - v) * im->height);
    if (i < 0) i = 0;
    if (i > im->width - 1) i = im->width - 1;
    if (j > im->height - 1) j = im->height - 1;
    if (j < 0) j = 0;
    const uint8_t *px = im->rgb + 3 * i + 3 * im->width * j;
    return V(px[0] / 255.0f, px[1] / 255.0f, px[2] / 255.0f);
  }
  }
}
void tpto_texture_value(const tpt_scene_desc *d, int tex, const float *uvp, int n, float *out) {
  for (int i = 0; i < n; i++) {
    v3 c = texture_value(d, tex, uvp[5 * i], uvp[5 * i + 1], V(uvp[5 * i + 2], uvp[5 * i + 3], uvp[5 * i + 4]));
    out[3 * i] = c.e[0]; out[3 * i + 1] = c.e[1]; out[3 * i + 2] = c.e[2];
  }
}

/* ---- src/utils.cc:437-450, headers/utils.h:55-58 ---- */
typedef struct { v3 u, v, w; } onb_t;
static onb_t onb_from_w(v3 n) {
  onb_t b;
  b.w = n;
  v3 tmp = (fabsf(n.e[0]) > 0.9) ? V(0, 1, 0) : V(1, 0, 0); /* float promoted, compared with double 0.9 */
  b.v = vunit(vcross(n, tmp));
  b.u = vcross(b.v, b.w);
  return b;
}
static v3 onb_local(const onb_t *b, float x, float y, float z) {
  return vadd(vadd(vscale(b->u, x), vscale(b->v, y)), vscale(b->w, z));
}

/* ---- samplers src/utils.cc:13-27,427-435 (g++: constructor arguments right to left) ---- */
static v3 random_in_unit_disk(rng_t *g) {
  v3 p;
  do {
    float y = (float)drand_r(g);
    float x = (float)drand_r(g);
    p = vsub(vscale(V(x, y, 0), 2.0f), V(1, 1, 0));
  } while (vdot(p, p) >= 1.0);
  return p;
}
static v3 random_in_unit_sphere(rng_t *g) {
  v3 p;
  do {
    float z = (float)drand_r(g);
    float y = (float)drand_r(g);
    float x = (float)drand_r(g);
    p = vsub(vscale(V(x, y, z), 2.0f), V(1.0f, 1.0f, 1.0f));
  } while (vlen(p) >= 1.0);
  return p;
}
static v3 random_on_hemisphere(rng_t *g) {
  float r1 = (float)drand_r(g);
  float r2 = (float)drand_r(g);
  float phi = (float)(2 * M_PI * r1);
  float x = COSF(phi) * sqrtf(r2);
  float y = SINF(phi) * sqrtf(r2);
  float z = sqrtf(1 - r2);
  return V(x, y, z);
}

/* ---- light shapes: hitable_list::pdf_value/random src/hitable_list.cc:24-36 ---- */
static float light_pdf_value(ctx_t *cx, v3 origin, v3 direction) {
  const tpt_scene_desc *d = cx->d;
  float weight = (float)(1.0 / d->n_lights);
  float sum = 0;
  for (int i = 0; i < d->n_lights; i++) {
    const tpt_light *L = &d->lights[i];
    float pdf = 0.0f;
    ray_t r = {origin, direction, 0.0f};
    rec_t rec;
    if (L->kind == TPT_LIGHT_XZ_RECT) { /* src/rect_box.cc:26-37 */
      if (rect_hit_c(1, 0, 2, L->p, &r, 0.0001f, FLT_MAX, &rec)) {
        float rec_area = fabsf((L->p[1] - L->p[0]) * (L->p[3] - L->p[2]));
        float distance_squared = vsqlen(vscale(direction, rec.t));
        float cosine = fabsf(vdot(direction, rec.n) / vlen(direction));
        pdf = distance_squared / (cosine * rec_area);
      }
    } else if (L->kind == TPT_LIGHT_SPHERE) { /* src/sphere.cc:93-106 */
      v3 c = V(L->p[0], L->p[1], L->p[2]);
      float radius = L->p[3];
      if (sphere_hit_c(c, c, radius, &r, 0.001f, FLT_MAX, &rec)) {
        float tmp = (radius * radius) / vsqlen(vsub(c, origin));
        float cosine_theta_max = (float)sqrt(1 - tmp);
        float solid_angle = (float)(2 * M_PI * (1 - cosine_theta_max));
        pdf = isnan(solid_angle) ? 0 : 1 / solid_angle;
      }
    } /* else hitable::pdf_value -> 0 (headers/hitable.h:35-37) */
    sum += weight * pdf;
  }
  return sum;
}
static v3 light_random(ctx_t *cx, v3 origin) {
  const tpt_scene_desc *d = cx->d;
  int index = (int)(drand_r(cx->g) * d->n_lights);
  const tpt_light *L = &d->lights[index];
  if (L->kind == TPT_LIGHT_XZ_RECT) { /* src/rect_box.cc:39-43: z drawn first */
    float z = (float)(L->p[2] + drand_r(cx->g) * (L->p[3] - L->p[2]));
    float x = (float)(L->p[0] + drand_r(cx->g) * (L->p[1] - L->p[0]));
    return vsub(V(x, L->p[4], z), origin);
  }
  if (L->kind == TPT_LIGHT_SPHERE) { /* src/sphere.cc:108-120 */
    v3 c = V(L->p[0], L->p[1], L->p[2]);
    float radius = L->p[3];
    v3 direction = vsub(c, origin);
    onb_t uvw = onb_from_w(vunit(direction));
    float tmp = (radius * radius) / vsqlen(vsub(c, origin));
    float cosine_theta_max = (float)sqrt(1 - tmp);
    float r1 = (float)drand_r(cx->g);
    float r2 = (float)drand_r(cx->g);
    float z = 1 + r2 * (cosine_theta_max - 1);
    float x = (float)(cos(2 * M_PI * r1) * sqrtf(1 - z * z));
    float y = (float)(sin(2 * M_PI * r1) * sqrtf(1 - z * z));
    return onb_local(&uvw, x, y, z);
  }
  return V(1, 0, 0); /* hitable::random, headers/hitable.h:38 */
}

/* ---- optics src/utils.cc:34-56 ---- */
static v3 reflect(v3 v, v3 n) { return vsub(v, vscale(n, 2 * vdot(v, n))); }
static int refract(v3 v, v3 n, float ni_over_nt, v3 *refracted) {
  v3 unit_v = vunit(v);
  float dt = vdot(unit_v, n);
  float discriminant = (float)(1.0 - ni_over_nt * ni_over_nt * (1 - dt * dt));
  if (discriminant > 0) {
    *refracted = vsub(vscale(vsub(unit_v, vscale(n, dt)), ni_over_nt), vscale(n, (float)sqrt(discriminant)));
    return 1;
  }
  return 0;
}
static float schlick(float cosine, float ref_index) {
  float r0 = (1 - ref_index) / (1 + ref_index);
  r0 = r0 * r0;
  return (float)(r0 + (1 - r0) * pow((1 - cosine), 5));
}

/* ---- color(): src/utils.cc:58-94 (recursive, no shortcut) ---- */
static v3 color(ctx_t *cx, const ray_t *r, int depth, int max_depth, float t_min) {
  const tpt_scene_desc *d = cx->d;
  rec_t rec;
  cx->rays++;
  rng_next_stage(cx->g);
  if (!world_hit(cx, r, t_min, FLT_MAX, &rec)) {
    if (d->background == TPT_BG_SKY) { /* the commented gradient src/utils.cc:87-90 */
      v3 ud = vunit(r->d);
      float t = (float)((ud.e[1] + 1.0) * 0.5);
      return vscale(vadd(vscale(V(1.0f, 1.0f, 1.0f), 1 - t), vscale(V(0.5f, 0.7f, 1.0f), t)), 0.1f);
    }
    return V(0, 0, 0);
  }
  const tpt_material *m = &d->materials[rec.mat];
  v3 emitted = V(0, 0, 0);
  if (m->kind == TPT_MAT_DIFFUSE_LIGHT && vdot(rec.n, r->d) < 0) /* src/material.cc:79-86 */
    emitted = texture_value(d, m->texture, rec.u, rec.v, rec.p);
  if (!(depth < max_depth)) return emitted;
  if (m->kind == TPT_MAT_METAL) { /* src/material.cc:88-98 */
    v3 reflected = reflect(vunit(r->d), rec.n);
    v3 dir = vadd(reflected, vscale(random_in_unit_sphere(cx->g), m->fuzz));
    if (!(vdot(dir, rec.n) > 0)) return emitted;
    ray_t s = {rec.p, dir, r->time};
    return vmul(V(m->albedo[0], m->albedo[1], m->albedo[2]), color(cx, &s, depth + 1, max_depth, t_min));
  }
  if (m->kind == TPT_MAT_DIELECTRIC) { /* src/material.cc:19-70 */
    v3 outward_normal, refracted, reflected = reflect(r->d, rec.n);
    float ni_over_nt, reflect_prob, cosine;
    if (vdot(r->d, rec.n) > 0) {
      outward_normal = vneg(rec.n);
      ni_over_nt = m->ref_idx;
      cosine = vdot(r->d, rec.n) / vlen(r->d);
      cosine = sqrtf(1 - m->ref_idx * m->ref_idx * (1 - cosine * cosine));
    } else {
      outward_normal = rec.n;
      ni_over_nt = (float)(1.0 / m->ref_idx);
      cosine = -vdot(r->d, rec.n) / vlen(r->d);
    }
    if (refract(r->d, outward_normal, ni_over_nt, &refracted)) reflect_prob = schlick(cosine, m->ref_idx);
    else reflect_prob = 1.0;
    ray_t s = {rec.p, drand_r(cx->g) < reflect_prob ? reflected : refracted, r->time};
    return color(cx, &s, depth + 1, max_depth, t_min); /* attenuation (1,1,1) */
  }
  if (m->kind != TPT_MAT_LAMBERTIAN) return emitted; /* base material::scatter -> false */
  /* lambertian::scatter src/material.cc:3-9 + mixture pdf src/utils.cc:73-81 */
  v3 attenuation = texture_value(d, m->texture, rec.u, rec.v, rec.p);
  onb_t uvw = onb_from_w(rec.n);
  v3 dir;
  if (drand_r(cx->g) < 0.5) dir = light_random(cx, rec.p);
  else {
    v3 h = random_on_hemisphere(cx->g);
    dir = onb_local(&uvw, h.e[0], h.e[1], h.e[2]);
  }
  ray_t scattered = {rec.p, dir, r->time};
  float c1 = vdot(vunit(dir), uvw.w); /* cosine_pdf::value returns cos (headers/utils.h:77-83) */
  float cos_pdf = c1 > 0 ? c1 : 0;
  float pdf_value = (float)(0.5 * light_pdf_value(cx, rec.p, dir) + 0.5 * cos_pdf);
  float c2 = vdot(rec.n, vunit(dir)); /* lambertian::scattering_pdf src/material.cc:11-17 */
  float spdf = c2 < 0 ? 0 : (float)(c2 / M_PI);
  v3 inner = color(cx, &scattered, depth + 1, max_depth, t_min);
  return vadd(emitted, vdivs(vmul(vscale(attenuation, spdf), inner), pdf_value));
}

/* ---- camera::get_ray src/camera.cc:23-31 ---- */
static ray_t get_ray(const tpt_camera *c, float s, float t, rng_t *g) {
  v3 rd = vscale(random_in_unit_disk(g), c->lens_radius);
  v3 u = V(c->u[0], c->u[1], c->u[2]), v = V(c->v[0], c->v[1], c->v[2]);
  v3 offset = vadd(vscale(u, rd.e[0]), vscale(v, rd.e[1]));
  float time = (float)(c->time0 + drand_r(g) * (c->time1 - c->time0));
  v3 origin = V(c->origin[0], c->origin[1], c->origin[2]);
  v3 llc = V(c->lower_left_corner[0], c->lower_left_corner[1], c->lower_left_corner[2]);
  v3 hor = V(c->horizontal[0], c->horizontal[1], c->horizontal[2]);
  v3 ver = V(c->vertical[0], c->vertical[1], c->vertical[2]);
  ray_t r;
  r.o = vadd(origin, offset);
  r.d = vsub(vsub(vadd(vadd(llc, vscale(hor, s)), vscale(ver, t)), origin), offset);
  r.time = time;
  return r;
}

/* ================================ entry points ================================ */
void tpto_hit_batch(const tpt_scene_desc *d, const tpt_ray *rays, size_t n, float tmin, float tmax, tpt_hit *out) {
  ctx_t cx = {d, NULL, 0};
  for (size_t i = 0; i < n; i++) {
    ray_t r = {V(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].time};
    rec_t rec;
    memset(&out[i], 0, sizeof(out[i]));
    out[i].prim = out[i].mat = -1;
    if (world_hit(&cx, &r, tmin, tmax, &rec)) {
      out[i].hit = 1; out[i].prim = rec.prim; out[i].mat = rec.mat;
      out[i].t = rec.t; out[i].u = rec.u; out[i].v = rec.v;
      for (int c = 0; c < 3; c++) { out[i].p[c] = rec.p.e[c]; out[i].n[c] = rec.n.e[c]; }
    }
  }
}

typedef struct {
  const tpt_scene_desc *d;
  const tpt_camera *cam;
  const tpt_render_params *p;
  float *out_sum, *out_samples;
  int next_row;
  pthread_mutex_t mu;
  uint64_t rays, draws;
} job_t;

/* main.cpp:115-134: ns jittered samples per pixel, col += de_nan(color()), running sums per slice */
static void *render_rows(void *arg) {
  job_t *J = (job_t *)arg;
  const tpt_render_params *p = J->p;
  const int nx = p->nx, ny = p->ny, ns = p->ns;
  const int slices = p->slices > 0 ? p->slices : 1, per_slice = ns / slices;
  rng_t g;
  memset(&g, 0, sizeof(g));
  g.key[0] = p->seed_lo; g.key[1] = p->seed_hi;
  ctx_t cx = {J->d, &g, 0};
  for (;;) {
    pthread_mutex_lock(&J->mu);
    int j = J->next_row++;
    pthread_mutex_unlock(&J->mu);
    if (j >= ny) break;
    for (int i = 0; i < nx; i++) {
      v3 col = V(0, 0, 0);
      int count = 0;
      for (int k = 0; k < ns; k++) {
        ++count;
        g.pixel = (uint32_t)(j * nx + i); g.sample = (uint32_t)k; g.stage = 0; g.ndraw = 0; g.block_id = 0xffffffffu;
        float u = (float)(((float)i + drand_r(&g)) / (float)nx);
        float v = (float)(((float)j + drand_r(&g)) / (float)ny);
        ray_t r = get_ray(J->cam, u, v, &g);
        v3 c = color(&cx, &r, 0, p->max_depth, p->t_min);
        for (int q = 0; q < 3; q++) if (isnan(c.e[q])) c.e[q] = 0; /* de_nan headers/utils.h:100-109 */
        col = vadd(col, c);
        if (J->out_samples) memcpy(J->out_samples + (((size_t)j * nx + i) * ns + k) * 3, c.e, 12);
        if (count % per_slice == 0 && count / per_slice - 1 < slices)
          memcpy(J->out_sum + (((size_t)(count / per_slice - 1) * ny + j) * nx + i) * 3, col.e, 12);
      }
    }
  }
  pthread_mutex_lock(&J->mu);
  J->rays += cx.rays; J->draws += g.total;
  pthread_mutex_unlock(&J->mu);
  return NULL;
}

/* out_sum [slices][ny][nx][3] running sums, out_samples (optional) [ny][nx][ns][3]; stats: rays, draws */
int tpto_render(const tpt_scene_desc *d, const tpt_camera *cam, const tpt_render_params *p, int threads,
                float *out_sum, float *out_samples, uint64_t *stats) {
  if (!d || !cam || !p || p->ns <= 0 || (p->slices > 0 && p->ns / p->slices <= 0) || d->n_lights <= 0) return -1;
  job_t J;
  memset(&J, 0, sizeof(J));
  J.d = d; J.cam = cam; J.p = p; J.out_sum = out_sum; J.out_samples = out_samples;
  pthread_mutex_init(&J.mu, NULL);
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256];
  for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, render_rows, &J);
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&J.mu);
  if (stats) { stats[0] = J.rays; stats[1] = J.draws; }
  return 0;
}

/* main.cpp:135-139,176-182 */
void tpto_quantise(const float *sum_rgb, size_t n, float denom, uint8_t *out) {
  for (size_t i = 0; i < n; i++) {
    float c = sum_rgb[i] / denom;
    c = (float)sqrt(c);
    int q = (int)(255.99f * c);
    out[i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}
