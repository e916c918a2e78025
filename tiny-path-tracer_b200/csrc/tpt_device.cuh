// csrc/tpt_device.cuh -- device side of the path-tracing core (sm_100a), shared by the
// megakernel, the wavefront kernels and the intersect-batch kernel.
//
// Everything is templated on PAR:
//   PAR = true  (TPT_MODE_PARITY): the reference's arithmetic, operation by operation -- fp64
//         exactly where the C++ expression promotes to double, IEEE div/sqrt, no FMA contraction
//         (the parity translation unit is compiled with -fmad=false), the reference's
//         comparison forms (NaN behaviour included), the un-shrunk t_max BVH walk with its tie
//         rules.  Each function cites the reference lines it restates.
//   PAR = false (TPT_MODE_FAST): the same estimator in fp32 with FMA / fast intrinsics and a
//         culled stackless traversal.
// Both consume the SAME Philox stream in the SAME draw order, so a fast render is comparable
// sample by sample with a parity render (and with the reference under the injected stream).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "tpt.h"

namespace tptd {

#define TPT_DEV __device__ __forceinline__
#define TPT_MAX_FRAMES 32
#define TPT_MAX_BOUNDARY_FRAMES 8 // nesting inside a constant_medium boundary (sphere / box / small list)
#define TPT_MAX_IMAGES 8
#define TPT_MAX_RANGES 256
#define TPT_TEXF_IMAGE 1      // texture feature bits of a kernel build (texture_value / fill_hit)
#define TPT_TEXF_PROCEDURAL 2
#define TPT_TEXF_ALL 3
#ifndef TPT_PAR_SKIP_WALK
#define TPT_PAR_SKIP_WALK 1 // PARITY, large simple trees: per-lane skip-pointer walk instead of the lock-step frame replay
#endif
#ifndef TPT_EXACT_DOUBLE_ROOTS
#define TPT_EXACT_DOUBLE_ROOTS 1 // FAST mode, huge "wall" spheres: roots in double like the reference (0: IEEE fp32)
#endif
#ifndef TPT_CAMERA_BLOCK_LOOP
#define TPT_CAMERA_BLOCK_LOOP 0 // 1: camera stage draws taken block by block from one Philox expansion site in a loop instead of Rng::next() + the out-of-line refill. Measured (r02, same box, Cornell A): +0.2 % unculled, -2.1 % with the pixel-bundle test (the library default), parity +-0: left off
#endif
#ifndef TPT_PAR_SKIP_ORDERED
#define TPT_PAR_SKIP_ORDERED 0 // closest_hit_skip_ordered: nearest child first on a per-lane stack. Measured on B200 (random_scene, parity, 1600x1600x16): 393 Mpaths/s against 544 for the sequential walk with the same cull (443 without it) -- identical images; the stack, the look-ahead for the right child and the larger parity kernel cost more than the saved boxes
#endif
#ifndef TPT_PAR_SKIP_CULL
#define TPT_PAR_SKIP_CULL 1 // closest_hit_skip leaves out boxes that lie behind the best hit so far (plus a margin)
#endif
#ifndef TPT_EAGER_CAMERA_BLOCK
#define TPT_EAGER_CAMERA_BLOCK 2 // second Philox block of the camera stage expanded in line: 0 never, 1 always, 2 PARITY kernels only
#endif

// ------------------------------------------------------------------------------------------
// Scene view: one contiguous blob of 16-byte words (shared memory when it fits, else global),
// laid out exactly as the C structs of include/tpt.h.
// ------------------------------------------------------------------------------------------
struct SceneLayout { // host-computed, lives in kernel parameter (constant) space
  const float4 *blob_global;
  int blob_words; // float4 count
  int off_nodes, off_prims, off_chains, off_ops, off_mats, off_texs, off_lights, off_perlin;
  int off_mediums, n_mediums; // MEDIUM primitive ids in DFS order (int list); boundaries live behind n_nodes
  int off_fbvh, off_fleaf, n_fbvh, fbvh_time_ok; // FAST-mode SAH BVH over world-space leaf boxes (0 nodes = none)
  int tree_simple; // no medium and no bvh_node below a hitable_list: PARITY walks the tree with skip pointers (closest_hit_skip)
  int skip_ordered;    // closest_hit_skip may visit a bvh_node's children nearest first (hints in the blob's node words, tree shallow enough for the lane's stack)
  float skip_cull_t0, skip_cull_t1; // ... for ray times in this interval (moving spheres leave their boxes outside their own [time0, time1])
  float skip_cull_abs; // closest_hit_skip: absolute part of the margin behind the best hit beyond which a box is not entered (< 0: never cull)
  int n_nodes, n_prims, n_lights, background;
  cudaTextureObject_t images[TPT_MAX_IMAGES];
  int image_w[TPT_MAX_IMAGES], image_h[TPT_MAX_IMAGES];
};

// Small scenes (<= TPT_SMALL_MAX_PRIMS primitives, no moving spheres): a second, flat copy of the
// geometry ordered by (transform chain, primitive kind), passed BY VALUE in the kernel
// parameters. Parameter space is the constant bank: the warp-uniform brute-force closest hit
// below reads it through the uniform datapath (ULDC / constant operands), so a primitive test
// costs no load instruction on the main pipes. FAST mode only.
#ifndef TPT_RECT_UNROLL
#define TPT_RECT_UNROLL 1 // a group holds 1-3 rects per axis: unrolling only adds remainder-loop overhead
#endif
constexpr int kRectUnroll = TPT_RECT_UNROLL;
#define TPT_SMALL_MAX_PRIMS 48
#define TPT_SMALL_MAX_GROUPS 8
#define TPT_SMALL_MAX_OPS 16
#define TPT_SMALL_MAX_BOXES 6
struct SmallGroup {
  // the chain's translate / rotate_y wrappers composed on the host: o' = Ry*o + b, d' = Ry*d
  float cs, sn, bx, by, bz;
  int first_op, n_ops;            // the chain itself (kept for reference; fill_hit replays it exactly)
  int xy_end, xz_end, yz_end, sph_end; // the group's records are [begin, xy_end) xy_rects, ... spheres
  int begin;
  int box_begin, box_end;         // axis-aligned boxes (a `box` = list of its six faces) under this chain
  int pad0;
};
// `box` (src/rect_box.cc:93-115) recognised at scene upload: six rect faces of one axis-aligned
// block. One slab test finds the entry / exit face instead of six rectangle tests.
// Loose rects of one chain that are whole faces of a common block (the walls of a Cornell room) are
// gathered the same way; faces nobody provides stay -1 and face[7] marks the block as open.
struct SmallBox {
  float4 lo, hi;  // pmin, pmax
  int face[8];    // prim ids of the faces: [-x, +x, -y, +y, -z, +z], -1 = absent; [6] chain (host), [7] open
};
struct SmallScene {
  int n_groups, enabled, pad0, pad1;
  SmallGroup groups[TPT_SMALL_MAX_GROUPS];
  float4 ops[TPT_SMALL_MAX_OPS];     // tpt_xform_op
  float4 geo[TPT_SMALL_MAX_PRIMS];   // rect: a0,a1,b0,b1 ; sphere: cx,cy,cz,r
  float2 aux[TPT_SMALL_MAX_PRIMS];   // rect: k, prim id (int bits) ; sphere: r*r, prim id
  SmallBox boxes[TPT_SMALL_MAX_BOXES];
};

// PARITY mode, small trees: the reference's hitable tree (bvh_node / hitable_list / leaves) laid out
// for a warp-uniform replay of world->hit from the constant bank -- closest_hit_flat below. It holds
// the SAME boxes, primitives and nesting as the pre-order node array; only the bookkeeping differs
// (no frame stack, no per-node kind decoding).
//   boxes  every non-duplicate bvh_node in pre-order: box_, chain, index of the enclosing bvh_node
//   items  what hangs below the bvh_nodes (ranked in DFS order, stored grouped by chain): one leaf, or one hitable_list (nested lists
//          and `box` objects concatenated: a list hands closest_so_far down and its result up, so a
//          list of lists is one list) as the run prims[first, end); `parent` = its bvh_node
//   prims  the leaves of the runs, in list order
#define TPT_FLAT_MAX_BOXES 32
#define TPT_FLAT_MAX_PRIMS 48
#define TPT_FLAT_MAX_ITEMS 48
struct FlatBox {
  float4 lo, hi; // lo.w = chain, hi.w = parent box (int bits; -1: a child of nothing)
};
struct FlatPrim {
  float4 geo; // rect: a0,a1,b0,b1 ; sphere: cx,cy,cz,r
  int kind, prim, chain;
  float k;    // rect plane coordinate
};
struct FlatItem {
  int first, end, parent, rank; // rank = position in DFS order (items are STORED grouped by transform chain)
};
struct FlatTree {
  int enabled, n_boxes, n_items, n_prims;
  FlatBox boxes[TPT_FLAT_MAX_BOXES];
  FlatPrim prims[TPT_FLAT_MAX_PRIMS];
  FlatItem items[TPT_FLAT_MAX_ITEMS];
};

struct SceneView {
  const float4 *blob;      // shared-memory copy when it fits, else the global blob
  const SceneLayout *L;    // offsets / counts / texture objects (constant bank)
  const SmallScene *small; // constant bank; valid when a FAST kernel is instantiated with SMALL
  const FlatTree *flat;    // constant bank; valid when a PARITY kernel is instantiated with SMALL
  bool wide_loads = false; // blob is the GLOBAL copy: SAH BVH nodes / leaf records may be fetched with 256-bit loads
};

struct V3 {
  float x, y, z;
};
TPT_DEV V3 mk(float x, float y, float z) { return V3{x, y, z}; }
TPT_DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
TPT_DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
TPT_DEV V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
TPT_DEV V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
TPT_DEV V3 operator*(float s, V3 a) { return mk(a.x * s, a.y * s, a.z * s); }
TPT_DEV V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
TPT_DEV V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
// headers/vec3.h:77-84: ((x*x + y*y) + z*z), no contraction in the parity TU
TPT_DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
TPT_DEV V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
TPT_DEV float sqlen(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
TPT_DEV float length(V3 a) { return sqrtf(sqlen(a)); } // headers/vec3.h:44-46 (std::sqrt(float))
// PARITY: the IEEE quotient a / b, written so that a ZERO dividend does not take the hardware
// division's out-of-line slow path (FCHK rejects a zero or denormal operand; r02 capture: 38 % of the
// warp-level rect tests called it for the 2.7 lanes whose ray starts ON that rect's plane, k - o == 0,
// and every unit(cross(n, axis)) of an axis-aligned normal divides exact zeros). 0 / b is +-0 with the
// XOR of the signs unless b is 0 or NaN (then NaN): the same value on every input, no call.
TPT_DEV float pdiv(float a, float b) {
  const bool zero = a == 0.0f;
  // the substitute dividend is formed in opaque PTX: written in C++ the compiler proves that the quotient is
  // only used when a != 0, folds the select away and divides a / b after all (r02 capture: 36 M slow-path
  // calls per 16-spp frame remained at 2.7 lanes)
  float safe;
  asm("{ .reg .pred p; setp.eq.f32 p, %1, 0f00000000; selp.f32 %0, 0f3F800000, %1, p; }" : "=f"(safe) : "f"(a));
  const float q = safe / b;
  const float z = (b == 0.0f || b != b) ? __int_as_float(0x7fffffff)
                                        : __int_as_float((__float_as_int(a) ^ __float_as_int(b)) & (int)0x80000000);
  return zero ? z : q;
}

template <bool PAR> TPT_DEV V3 unit(V3 a) {         // headers/vec3.h:143-144: v / v.length()
  if (PAR) {
    const float l = length(a);
    return mk(pdiv(a.x, l), pdiv(a.y, l), pdiv(a.z, l));
  }
  float inv = rsqrtf(sqlen(a));
  return a * inv;
}

struct Ray {
  V3 o, d;
  float time;
};

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter = (pixel, sample, stage, block),
// key = (seed_lo, seed_hi); four 24-bit uniforms per block, generated lazily.
// stage 0 = camera draws (main.cpp:121-124), stage d+1 = draws of color() at depth d.
// ------------------------------------------------------------------------------------------
TPT_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                           uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// The same block function with the ten round keys (k + r * Weyl constant) precomputed on the host
// (RenderArgs::rk, kernel parameter space): the key schedule is uniform over the whole launch, so
// every round's key is a constant-bank operand of the XOR instead of two adds per round at every
// expansion site (a third of the block function's instructions).
TPT_DEV void philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t *rk, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// out-of-line block function for the rare refills (keeps the Rng state in registers: nothing
// takes its address)
static __device__ __noinline__ uint4 philox_block_slow(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                uint32_t k1) {
  uint32_t o[4];
  philox4x32_10(c0, c1, c2, c3, k0, k1, o);
  return make_uint4(o[0], o[1], o[2], o[3]);
}

struct Rng {
  const uint32_t *rk; // round keys: rk[2r] = seed_lo + r * 0x9E3779B9, rk[2r+1] = seed_hi + r * 0xBB67AE85
  uint32_t pixel, sample, stage, ndraw;
  uint32_t fresh; // draw index whose block is already loaded (set_stage: 0, set_stage_at(n): n & ~3)
  uint32_t b0, b1, b2, b3;
  TPT_DEV void begin(const uint32_t *round_keys, uint32_t pix, uint32_t smp) {
    rk = round_keys;
    pixel = pix;
    sample = smp;
    stage = 0;
    ndraw = 0;
    fresh = 0;
  }
  // Start a stage and generate its first block right away: camera (u, v, lens y, lens x),
  // lambertian (mixture pick, light index, 2 coordinates), dielectric (1) and the first round of
  // metal's rejection loop (3) all fit in four draws, so the block function is expanded at ONE
  // place per stage instead of at every draw (instruction-cache footprint) and next() is three
  // selects. Draws beyond four (rejection-loop retries, shutter time) take the out-of-line refill.
  TPT_DEV void set_stage(uint32_t s) {
    stage = s;
    ndraw = 0;
    fresh = 0;
    uint32_t o[4];
    philox4x32_10_rk(pixel, sample, stage, 0u, rk, o);
    b0 = o[0];
    b1 = o[1];
    b2 = o[2];
    b3 = o[3];
  }
  // resume a stage after `n` draws were already consumed (constant_medium::hit draws inside
  // world->hit, before the material's own draws of the same stage)
  TPT_DEV void set_stage_at(uint32_t s, uint32_t n) {
    stage = s;
    ndraw = n;
    fresh = n & ~3u;
    uint32_t o[4];
    philox4x32_10_rk(pixel, sample, stage, n >> 2, rk, o);
    b0 = o[0];
    b1 = o[1];
    b2 = o[2];
    b3 = o[3];
  }
  // the stage's NEXT block, expanded in line where a stage is known to run past four draws (the
  // camera stage: jitter u, v + lens y, x fill block 0 and a fifth of the lens samples are redrawn).
  // ndraw must be a multiple of four.
  TPT_DEV void next_block_inline() {
    fresh = ndraw;
    uint32_t o[4];
    philox4x32_10_rk(pixel, sample, stage, ndraw >> 2, rk, o);
    b0 = o[0];
    b1 = o[1];
    b2 = o[2];
    b3 = o[3];
  }
  TPT_DEV void refill() {
    uint4 o = philox_block_slow(pixel, sample, stage, ndraw >> 2, rk[0], rk[1]);
    b0 = o.x;
    b1 = o.y;
    b2 = o.z;
    b3 = o.w;
  }
  // next uniform in [0,1): (x >> 8) * 2^-24, exactly representable in fp32
  TPT_DEV float next() {
    uint32_t lane = ndraw & 3u;
    if (lane == 0 && ndraw != fresh) refill();
    uint32_t x = lane == 0 ? b0 : (lane == 1 ? b1 : (lane == 2 ? b2 : b3));
    ndraw++;
    return (float)(x >> 8) * 5.9604644775390625e-8f;
  }
};

// ------------------------------------------------------------------------------------------
// Transform chains: translate / rotate_y wrappers between the root and a node, outermost first
// ------------------------------------------------------------------------------------------
struct XRay { // the ray expressed in the space of `chain`
  V3 o, d;
  V3 inv; // 1/d. PARITY: the IEEE quotient 1.0f / d of src/aabb.cc:5, taken once per chain instead of at every box
  int chain;
};

template <bool PAR, bool INV = true> TPT_DEV void to_chain(const SceneView &S, const Ray &r, int chain, XRay &x) {
  if (x.chain == chain) return;
  V3 o = r.o, d = r.d;
  if (chain != 0) {
    float4 c = S.blob[S.L->off_chains + chain];
    int first = __float_as_int(c.x), n = __float_as_int(c.y);
    // chains hold 1-2 wrappers: unrolled copies only spread the FAST kernels' hot code over more
    // I-cache lines (there the chain is entered once per shaded hit); the PARITY walk enters it at every
    // node of the reference's tree and keeps the compiler's unrolling
#pragma unroll(PAR ? 4 : 1)
    for (int i = 0; i < n; i++) {
      float4 op = S.blob[S.L->off_ops + first + i];
      if (__float_as_int(op.x) == TPT_XF_TRANSLATE) {
        // headers/rect_box.h:89: moved_r(origin - offset_, direction)
        o = o - mk(op.y, op.z, op.w);
      } else {
        // src/rect_box.cc:175-179: x' = cos*x - sin*z ; z' = sin*x + cos*z (origin and direction)
        float s = op.y, co = op.z;
        float ox = co * o.x - s * o.z, oz = s * o.x + co * o.z;
        float dx = co * d.x - s * d.z, dz = s * d.x + co * d.z;
        o.x = ox;
        o.z = oz;
        d.x = dx;
        d.z = dz;
      }
    }
  }
  x.o = o;
  x.d = d;
  x.chain = chain;
  if (INV) x.inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
}

// hit point / normal back to world space: wrappers unwind innermost first
// (src/rect_box.cc:185-189 then headers/rect_box.h:91)
TPT_DEV void from_chain(const SceneView &S, int chain, V3 &p, V3 &n) {
  if (chain == 0) return;
  float4 c = S.blob[S.L->off_chains + chain];
  int first = __float_as_int(c.x), cnt = __float_as_int(c.y);
#pragma unroll 1
  for (int i = cnt - 1; i >= 0; i--) {
    float4 op = S.blob[S.L->off_ops + first + i];
    if (__float_as_int(op.x) == TPT_XF_TRANSLATE) {
      p = p + mk(op.y, op.z, op.w);
    } else {
      float s = op.y, co = op.z;
      float px = co * p.x + s * p.z, pz = -s * p.x + co * p.z;
      float nx = co * n.x + s * n.z, nz = -s * n.x + co * n.z;
      p.x = px;
      p.z = pz;
      n.x = nx;
      n.z = nz;
    }
  }
}

// ------------------------------------------------------------------------------------------
// AABB slab test. PAR: src/aabb.cc:3-19 verbatim (1.0f/d per axis, swap on negative, ternary
// updates so NaN keeps the old bound, reject on tmin > tmax).
// ------------------------------------------------------------------------------------------
template <bool PAR>
TPT_DEV bool aabb_hit(const XRay &x, float4 lo, float4 hi, float tmin, float tmax) {
  if (PAR) {
    // `inv_D = 1.0f / r.direction()[i]` depends on the ray alone: the same IEEE quotient, computed
    // where the ray enters the chain's space (to_chain) instead of at each of the tree's boxes
    const float o[3] = {x.o.x, x.o.y, x.o.z}, iv[3] = {x.inv.x, x.inv.y, x.inv.z};
    const float mn[3] = {lo.x, lo.y, lo.z}, mx[3] = {hi.x, hi.y, hi.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float inv_d = iv[i];
      float t0 = (mn[i] - o[i]) * inv_d;
      float t1 = (mx[i] - o[i]) * inv_d;
      if (inv_d < 0.0f) {
        float t = t0;
        t0 = t1;
        t1 = t;
      }
      tmin = t0 > tmin ? t0 : tmin;
      tmax = t1 < tmax ? t1 : tmax;
      if (tmin > tmax) return false;
    }
    return true;
  } else {
    float ax = (lo.x - x.o.x) * x.inv.x, bx = (hi.x - x.o.x) * x.inv.x;
    float ay = (lo.y - x.o.y) * x.inv.y, by = (hi.y - x.o.y) * x.inv.y;
    float az = (lo.z - x.o.z) * x.inv.z, bz = (hi.z - x.o.z) * x.inv.z;
    float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    return t0 <= t1;
  }
}

// ------------------------------------------------------------------------------------------
// Primitive tests: return true and the accepted t. `incl` semantics are the primitive's own:
// rects accept t_min <= t <= t_max (src/rect_box.cc:11,55,78), spheres t_min < t < t_max
// (src/sphere.cc:24,33).
// ------------------------------------------------------------------------------------------
TPT_DEV V3 moving_center(float4 a, float4 b, float4 c, float time) {
  // headers/sphere.h:28-31: center0 + (time - time0)/(time1 - time0) * (center1 - center0)
  float f = (time - b.w) / (c.x - b.w);
  V3 c0 = mk(a.x, a.y, a.z), c1 = mk(b.x, b.y, b.z);
  return c0 + f * (c1 - c0);
}

// EXACT (fast mode only), for the huge "wall" / ground spheres (radius >= 500), where the last bit of
// t decides whether a ray that starts ON the sphere re-hits it at t ~ 0.001 and which of two nearly
// coincident surfaces wins: IEEE fp32 sqrt / divide, and from radius 1e4 up the reference's own root
// arithmetic -- sqrt and the division in double, rounded once (src/sphere.cc:23,32; r02: with fp32
// roots 47 of 2474 surface-start rays of sphere_cornell_box, whose walls are 1e5-radius spheres, picked
// another wall than the reference). Ordinary spheres take the approximate MUFU forms.
template <bool PAR, bool EXACT = true>
TPT_DEV bool sphere_test(V3 center, float radius, const XRay &x, float tmin, float tmax, float &t) {
  // src/sphere.cc:15-41. The quadratic's coefficients and discriminant are evaluated with
  // explicitly rounded fp32 operations in BOTH modes (never contracted into FMAs): the
  // discriminant of a 1000- or 1e5-radius "wall" sphere is a catastrophic cancellation, and only
  // the reference's own rounding sequence reproduces which of two such spheres wins.
  V3 oc = x.o - center;
  float a = __fadd_rn(__fadd_rn(__fmul_rn(x.d.x, x.d.x), __fmul_rn(x.d.y, x.d.y)), __fmul_rn(x.d.z, x.d.z));
  float b = 2.0f * __fadd_rn(__fadd_rn(__fmul_rn(x.d.x, oc.x), __fmul_rn(x.d.y, oc.y)), __fmul_rn(x.d.z, oc.z));
  float c = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(oc.x, oc.x), __fmul_rn(oc.y, oc.y)), __fmul_rn(oc.z, oc.z)),
                      -__fmul_rn(radius, radius));
  float disc = __fadd_rn(__fmul_rn(b, b), -__fmul_rn(__fmul_rn(4.0f, a), c));
  if (disc > 0) {
    // FAST: double only for the truly gigantic spheres (1e5-radius "walls", src/utils.cc:267-271); a
    // 1000-radius ground sphere is tested by every ray of its scene and matched the reference on every
    // ray with IEEE fp32 roots already (r02: double roots there cost two_perlin_spheres 30 %)
    if (PAR || (EXACT && TPT_EXACT_DOUBLE_ROOTS && radius >= 1.0e4f)) {
      // `sqrt` (unqualified) and the division are evaluated in double, rounded once to float
      double sq = sqrt((double)disc);
      float temp = (float)((-(double)b - sq) / (double)(2 * a));
      if (temp < tmax && temp > tmin) {
        t = temp;
        return true;
      }
      temp = (float)((-(double)b + sq) / (double)(2 * a));
      if (temp < tmax && temp > tmin) {
        t = temp;
        return true;
      }
    } else if (EXACT) {
      float sq = __fsqrt_rn(disc);
      float two_a = 2.0f * a;
      float temp = __fdiv_rn(-b - sq, two_a);
      if (temp < tmax && temp > tmin) {
        t = temp;
        return true;
      }
      temp = __fdiv_rn(-b + sq, two_a);
      if (temp < tmax && temp > tmin) {
        t = temp;
        return true;
      }
    } else {
      float sq = sqrtf(disc);
      float inv = __frcp_rn(2.0f * a); // one reciprocal, two multiplies
      float temp = (-b - sq) * inv;
      if (temp < tmax && temp > tmin) {
        t = temp;
        return true;
      }
      temp = (-b + sq) * inv;
      if (temp < tmax && temp > tmin) {
        t = temp;
        return true;
      }
    }
  }
  return false;
}

// FAST mode, huge "wall" spheres only: the exact-root variant as an out-of-line call. It is rare
// (no such sphere in most scenes) and its IEEE sqrt / divide sequences are long; inlined at every
// sphere site they sat between the hot instructions. The render kernels' hot loop is about as large
// as the SM's 32 KB instruction cache (ncu: 93 % hit rate, GPC instruction fetch at 79 % of its
// peak before this), so rarely executed code is kept out of line throughout this file.
static __device__ __noinline__ float2 sphere_test_exact_slow(float cx, float cy, float cz, float radius, float ox, float oy,
                                                         float oz, float dx, float dy, float dz, float tmin, float tmax) {
  XRay x;
  x.o = mk(ox, oy, oz);
  x.d = mk(dx, dy, dz);
  x.chain = 0;
  float t = 0.f;
  const bool hit = sphere_test<false, true>(mk(cx, cy, cz), radius, x, tmin, tmax, t);
  return make_float2(hit ? 1.0f : 0.0f, t);
}
TPT_DEV bool sphere_test_exact_fast(V3 c, float radius, const XRay &x, float tmin, float tmax, float &t) {
  const float2 r = sphere_test_exact_slow(c.x, c.y, c.z, radius, x.o.x, x.o.y, x.o.z, x.d.x, x.d.y, x.d.z, tmin, tmax);
  t = r.y;
  return r.x != 0.0f;
}

// axis: 0 = xy_rect (plane z=k), 1 = xz_rect (y=k), 2 = yz_rect (x=k); p = a0,a1,b0,b1,k
template <bool PAR>
TPT_DEV bool rect_test(int axis, float4 p, float k, const XRay &x, float tmin, float tmax, float &t) {
  float ok, dk, oa, da, ob, db;
  if (axis == 0) {
    ok = x.o.z; dk = x.d.z; oa = x.o.x; da = x.d.x; ob = x.o.y; db = x.d.y;
  } else if (axis == 1) {
    ok = x.o.y; dk = x.d.y; oa = x.o.x; da = x.d.x; ob = x.o.z; db = x.d.z;
  } else {
    ok = x.o.x; dk = x.d.x; oa = x.o.y; da = x.d.y; ob = x.o.z; db = x.d.z;
  }
  float tt = PAR ? pdiv(k - ok, dk) : (k - ok) / dk;
  if (tt > tmax || tt < tmin) return false;
  float a = oa + tt * da;
  float b = ob + tt * db;
  if (a < p.x || a > p.y || b < p.z || b > p.w) return false;
  t = tt;
  return true;
}

template <bool PAR>
TPT_DEV bool prim_test(const SceneView &S, int prim, const XRay &x, float time, float tmin,
                       float tmax, float &t) {
  const float4 *P = S.blob + S.L->off_prims + 4 * prim;
  float4 h = P[0], a = P[1];
  int kind = __float_as_int(h.x);
  if (kind > TPT_PRIM_YZ_RECT) return false; // MEDIUM: handled by medium_test where a stream exists
  if (kind == TPT_PRIM_SPHERE) {
    return sphere_test<PAR>(mk(a.x, a.y, a.z), a.w, x, tmin, tmax, t);
  } else if (kind == TPT_PRIM_MOVING_SPHERE) {
    float4 b = P[2], c = P[3];
    return sphere_test<PAR>(moving_center(a, b, c, time), a.w, x, tmin, tmax, t);
  } else {
    float k = P[2].x;
    return rect_test<PAR>(kind - TPT_PRIM_XY_RECT, a, k, x, tmin, tmax, t);
  }
}

// ------------------------------------------------------------------------------------------
// Closest hit.
// PAR: exact emulation of the recursive reference walk on the pre-order node array with a small
// frame stack. A frame is an open bvh_node or hitable_list:
//   bvh_node::hit   (src/hitable.cc:63-90): box test with the incoming t_max, BOTH children get
//                   that same t_max, results merged with `left.t < right.t ? left : right`
//   hitable_list::hit (src/hitable_list.cc:38-51): children get closest_so_far, any success
//                   replaces the record.
// A virtual root LIST frame carries the caller's t_max.
// FAST: stackless skip-pointer walk, t_max shrinks to the best hit, duplicate sub-trees skipped.
// ------------------------------------------------------------------------------------------
struct Frame {
  float tmax_in, best_t;
  int best_prim, end_list; // end index | (is_list << 31)
};

template <bool PAR, bool MED>
TPT_DEV bool walk_range(const SceneView &S, const Ray &r, int first, int end_all, float tmin, float tmax, float &t_out,
                        int &prim_out, Rng *g);

// constant_medium::hit (src/hitable.cc:98-128). `x` is the ray in the medium's own space, the
// boundary sub-tree [bf, be) is walked from the world ray (its nodes carry absolute chains; t is
// the same in every space). Draws: one for the free-flight distance, three for the "arbitrary"
// normal -- consumed inside world->hit, i.e. before the material's draws of the same stage.
template <bool PAR>
TPT_DEV bool medium_test(const SceneView &S, int prim, const XRay &x, const Ray &r, float tmin, float tmax, float &t,
                         Rng *g) {
  if (!g) return false; // closest-hit batches have no stream: media are refused by the API
  const float4 *P = S.blob + S.L->off_prims + 4 * prim;
  const float4 a = P[1];
  const float density = a.x;
  const int bf = __float_as_int(a.y), be = __float_as_int(a.z);
  float t1, t2;
  int dummy;
  if (!walk_range<PAR, false>(S, r, bf, be, -FLT_MAX, FLT_MAX, t1, dummy, nullptr)) return false;
  const float tmin2 = PAR ? (float)((double)t1 + 0.0001) : t1 + 0.0001f;
  if (!walk_range<PAR, false>(S, r, bf, be, tmin2, FLT_MAX, t2, dummy, nullptr)) return false;
  if (t1 < tmin) t1 = tmin;
  if (t2 > tmax) t2 = tmax;
  if (t1 >= t2) return false;
  const float len = length(x.d);
  const float inside = (t2 - t1) * len;
  const float u = g->next();
  // (-1 / density_) * std::log(drand_r()): float * double -> float
  const float hit_distance = PAR ? (float)((double)(-1 / density) * log((double)u)) : (-1.0f / density) * __logf(u);
  if (hit_distance < inside) {
    t = t1 + hit_distance / len;
    g->next(); // rec.normal = vec3(drand_r(), drand_r(), drand_r()): never used (isotropic is an
    g->next(); // absorber at HEAD) but the stream advances
    g->next();
    return true;
  }
  return false;
}

template <bool PAR, bool MED>
TPT_DEV bool any_prim_test(const SceneView &S, int prim, const XRay &x, const Ray &r, float tmin, float tmax, float &t,
                           Rng *g) {
  if (MED && __float_as_int(S.blob[S.L->off_prims + 4 * prim].x) == TPT_PRIM_MEDIUM)
    return medium_test<PAR>(S, prim, x, r, tmin, tmax, t, g);
  return prim_test<PAR>(S, prim, x, r.time, tmin, tmax, t);
}

// hit() of the sub-tree stored at nodes [first, end_all)
template <bool PAR, bool MED>
TPT_DEV bool walk_range(const SceneView &S, const Ray &r, int first, int end_all, float tmin, float tmax, float &t_out,
                        int &prim_out, Rng *g) {
  XRay x;
  x.chain = -1;
  const float4 *N = S.blob + S.L->off_nodes;
  const int n = end_all;
  if (PAR) {
    Frame fr[MED ? TPT_MAX_FRAMES : TPT_MAX_BOUNDARY_FRAMES];
    int sp = 1;
    fr[0].tmax_in = tmax;
    fr[0].best_t = 0.f;
    fr[0].best_prim = -1;
    fr[0].end_list = n | (int)0x80000000;
    // The walk is kept in lock step across the lanes of a warp: `i` only moves by decisions every
    // converged lane shares. A lane whose box test fails does not jump ahead on its own (that put
    // the lanes of a warp at different nodes: 10 of 32 active in this loop on the Cornell tree);
    // it marks itself dead until the end of that sub-tree and idles through it, and the sub-tree
    // is skipped only when no lane wants it. A dead lane records nothing, so the frames it opens
    // close empty and the result is the one the private walk would produce.
    int i = first;
    int dead_until = first; // this lane takes part in node i iff i >= dead_until
    for (;;) {
      // close finished groups, handing their result to the parent
      while (sp > 1 && i == (fr[sp - 1].end_list & 0x7fffffff)) {
        Frame f = fr[--sp];
        if (f.best_prim >= 0) {
          Frame &P = fr[sp - 1];
          bool take = (P.end_list < 0) || (P.best_prim < 0) || !(P.best_t < f.best_t);
          if (take) {
            P.best_t = f.best_t;
            P.best_prim = f.best_prim;
          }
        }
      }
      if (i >= n) break;
      const bool alive = i >= dead_until;
      Frame &P = fr[sp - 1];
      float ctx = (P.end_list < 0 && P.best_prim >= 0) ? P.best_t : P.tmax_in;
      float4 n0 = N[2 * i], n1 = N[2 * i + 1];
      int kind = __float_as_int(n0.w);
      int chain = kind >> 16;
      int k = kind & 0xff;
      to_chain<PAR>(S, r, chain, x);
      if (k == TPT_NODE_LEAF) {
        float t;
        int prim = __float_as_int(n1.w);
        if (alive && any_prim_test<PAR, MED>(S, prim, x, r, tmin, ctx, t, g)) {
          bool take = (P.end_list < 0) || (P.best_prim < 0) || !(P.best_t < t);
          if (take) {
            P.best_t = t;
            P.best_prim = prim;
          }
        }
        i++;
      } else {
        int end = __float_as_int(n1.w);
        const bool enter = alive && !(k == TPT_NODE_BVH && !aabb_hit<PAR>(x, n0, n1, tmin, ctx));
        if (alive && !enter) dead_until = end;
        if (!__any_sync(__activemask(), enter)) {
          i = end;
        } else {
          Frame &F = fr[sp++];
          F.tmax_in = ctx;
          F.best_t = 0.f;
          F.best_prim = -1;
          F.end_list = end | (k == TPT_NODE_LIST ? (int)0x80000000 : 0);
          i++;
        }
      }
    }
    t_out = fr[0].best_t;
    prim_out = fr[0].best_prim;
    return fr[0].best_prim >= 0;
  } else {
    float best = tmax;
    int best_prim = -1;
    int i = first;
    while (i < n) {
      float4 n0 = N[2 * i], n1 = N[2 * i + 1];
      int kind = __float_as_int(n0.w);
      int k = kind & 0xff;
      int payload = __float_as_int(n1.w);
      if (kind & TPT_NODE_DUP) {
        i = (k == TPT_NODE_LEAF) ? i + 1 : payload;
        continue;
      }
      to_chain<PAR>(S, r, kind >> 16, x);
      if (k == TPT_NODE_LEAF) {
        float t;
        if (any_prim_test<PAR, MED>(S, payload, x, r, tmin, best, t, g)) {
          best = t;
          best_prim = payload;
        }
        i++;
      } else {
        i = aabb_hit<PAR>(x, n0, n1, tmin, best) ? i + 1 : payload;
      }
    }
    t_out = best;
    prim_out = best_prim;
    return best_prim >= 0;
  }
}

// ------------------------------------------------------------------------------------------
// PARITY mode, small trees: world->hit replayed from the FlatTree with warp-uniform control flow.
// Why it is the same function as the recursive reference (and as walk_range<true> above):
//  * bvh_node::hit (src/hitable.cc:63-90) hands its own t_max to both children, so below the
//    bvh_nodes every leaf / list sees the caller's t_max, whatever its siblings found: the items
//    are independent of each other. A node's box is consulted only when every bvh_node above it
//    let the ray in -- `reach` bit b = box b passed AND its parent is reached.
//  * hitable_list::hit (src/hitable_list.cc:38-51) is replayed literally per item: children in
//    list order against closest_so_far, any success replaces the record (rect bounds inclusive,
//    sphere bounds strict -- the primitive tests themselves are the ones walk_range uses).
//  * the nested merge `left.t < right.t ? left : right` is a tournament whose winner is the
//    smallest t, the LAST one in DFS order among equals -- unless a candidate's t is NaN (a rect
//    "hit" of a ray lying in its plane through 0/0), where `<` stops being an order. Such a ray
//    is handed to walk_range, the literal frame-by-frame replay.
//  * a one-element bvh_node hits its child twice and merges two identical records: the duplicate
//    is not replayed.
// A lane outside a box idles through that sub-tree (its result is masked); a sub-tree is skipped
// when no lane of the warp reaches it.
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ float2 closest_hit_replay(const float4 *blob, const SceneLayout *L, float ox, float oy, float oz,
                                                         float dx, float dy, float dz, float time, float tmin, float tmax) {
  SceneView S;
  S.blob = blob;
  S.L = L;
  S.small = nullptr;
  S.flat = nullptr;
  Ray r;
  r.o = mk(ox, oy, oz);
  r.d = mk(dx, dy, dz);
  r.time = time;
  float t = 0.f;
  int prim = -1;
  walk_range<true, true>(S, r, 0, L->n_nodes, tmin, tmax, t, prim, nullptr);
  return make_float2(t, __int_as_float(prim));
}

TPT_DEV bool closest_hit_flat(const SceneView &S, const Ray &r, float tmin, float tmax, float &t_out, int &prim_out) {
  const FlatTree &F = *S.flat;
  XRay x;
  x.chain = -1;
  unsigned reach = 0u;
  for (int b = 0; b < F.n_boxes; b++) {
    const float4 lo = F.boxes[b].lo, hi = F.boxes[b].hi;
    const int parent = __float_as_int(hi.w);
    const bool open = parent < 0 || ((reach >> parent) & 1u);
    if (!__any_sync(__activemask(), open)) continue;
    to_chain<true>(S, r, __float_as_int(lo.w), x);
    if (open && aabb_hit<true>(x, lo, hi, tmin, tmax)) reach |= 1u << b;
  }
  x.chain = -1; // the primitive tests below need no 1/d: their chain changes skip the three quotients
  float best_t = 0.f;
  int best_prim = -1, best_rank = -1;
  bool unordered = false;
  for (int it = 0; it < F.n_items; it++) {
    const FlatItem I = F.items[it];
    const bool act = I.parent < 0 || ((reach >> I.parent) & 1u);
    if (!__any_sync(__activemask(), act)) continue;
    float run_t = tmax; // closest_so_far of this list (a lone leaf: the caller's t_max)
    int run_prim = -1;
    for (int k = I.first; k < I.end; k++) {
      const FlatPrim P = F.prims[k];
      to_chain<true, false>(S, r, P.chain, x);
      float t;
      bool hit;
      if (P.kind == TPT_PRIM_SPHERE) hit = sphere_test<true>(mk(P.geo.x, P.geo.y, P.geo.z), P.geo.w, x, tmin, run_t, t);
      else hit = rect_test<true>(P.kind - TPT_PRIM_XY_RECT, P.geo, P.k, x, tmin, run_t, t);
      if (act && hit) {
        run_t = t;
        run_prim = P.prim;
      }
    }
    if (run_prim >= 0) {
      unordered = unordered || isnan(run_t);
      // the items are visited grouped by transform chain (each chain is entered once per ray), not in DFS
      // order: "the smallest t wins, the LAST in DFS order among equals" is applied through the rank
      if (best_prim < 0 || run_t < best_t || (run_t == best_t && I.rank > best_rank)) {
        best_t = run_t;
        best_prim = run_prim;
        best_rank = I.rank;
      }
    }
  }
  if (unordered) {
    const float2 w = closest_hit_replay(S.blob, S.L, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.time, tmin, tmax);
    best_t = w.x;
    best_prim = __float_as_int(w.y);
  }
  t_out = best_t;
  prim_out = best_prim;
  return best_prim >= 0;
}

// ------------------------------------------------------------------------------------------
// PARITY mode, trees of any size whose hitable_lists hold no bvh_node (SceneLayout::tree_simple, e.g.
// random_scene's 485-leaf tree): the same argument as closest_hit_flat, walked per lane over the
// pre-order node array with skip pointers instead of a frame stack. Below the bvh_nodes every leaf /
// list sees the caller's t_max, so the merged result is "smallest t, last in DFS order among equals":
// a sequential scan in DFS order with `take unless best.t < t`; a list (nested lists concatenate) is
// one run against its own closest_so_far; a failed box skips its sub-tree; a NaN candidate hands the
// ray to the literal replay.
// ------------------------------------------------------------------------------------------
TPT_DEV bool closest_hit_skip(const SceneView &S, const Ray &r, float tmin, float tmax, float &t_out, int &prim_out) {
  const float4 *N = S.blob + S.L->off_nodes;
  const int n = S.L->n_nodes;
  XRay x;
  x.chain = -1;
  float best_t = 0.f, run_t = tmax;
  int best_prim = -1, run_prim = -1, list_end = -1;
  bool unordered = false;
  // The reference tests every box against the caller's t_max (bvh_node::hit never shrinks it), so it enters
  // boxes that lie wholly behind the closest hit found so far. Nothing in such a box can change the result
  // (a candidate replaces the best only with t <= best_t), so the walk leaves them out: a box is tested
  // against best_t plus a margin (0.1 % + 1e-5 of the scene's extent) that is orders of magnitude wider than
  // the rounding of the slab test and of the primitives' own t, and narrow enough to cut the node visits.
  // With t_max' < t_max AABB::hit differs only by refusing boxes whose entry distance is >= t_max'.
  const float cull_abs = (r.time >= S.L->skip_cull_t0 && r.time <= S.L->skip_cull_t1) ? S.L->skip_cull_abs : -1.f;
  float box_max = tmax;
#define TPT_SKIP_TOOK() if (TPT_PAR_SKIP_CULL && cull_abs >= 0.f) box_max = fminf(tmax, fmaf(best_t, 1.001f, cull_abs))
  for (int i = 0; i < n;) {
    if (i == list_end) { // the (outermost) list closes: its record goes up to the enclosing bvh_node
      if (run_prim >= 0) {
        unordered = unordered || isnan(run_t);
        if (best_prim < 0 || !(best_t < run_t)) {
          best_t = run_t;
          best_prim = run_prim;
          TPT_SKIP_TOOK();
        }
      }
      list_end = -1;
    }
    const float4 n0 = N[2 * i], n1 = N[2 * i + 1];
    const int kind = __float_as_int(n0.w), k = kind & 0xff, payload = __float_as_int(n1.w);
    if (kind & TPT_NODE_DUP) { // second visit of a one-element bvh_node's child: the same record twice
      i = k == TPT_NODE_LEAF ? i + 1 : payload;
      continue;
    }
    if (k == TPT_NODE_LEAF) {
      to_chain<true>(S, r, kind >> 16, x);
      float t;
      if (list_end >= 0) {
        if (prim_test<true>(S, payload, x, r.time, tmin, run_t, t)) {
          run_t = t;
          run_prim = payload;
        }
      } else if (prim_test<true>(S, payload, x, r.time, tmin, tmax, t)) {
        unordered = unordered || isnan(t);
        if (best_prim < 0 || !(best_t < t)) {
          best_t = t;
          best_prim = payload;
          TPT_SKIP_TOOK();
        }
      }
      i++;
    } else if (k == TPT_NODE_LIST) {
      if (list_end < 0) {
        list_end = payload;
        run_t = tmax;
        run_prim = -1;
      }
      i++;
    } else {
      to_chain<true>(S, r, kind >> 16, x);
      i = aabb_hit<true>(x, n0, n1, tmin, box_max) ? i + 1 : payload;
    }
  }
  if (list_end >= 0 && run_prim >= 0) {
    unordered = unordered || isnan(run_t);
    if (best_prim < 0 || !(best_t < run_t)) {
      best_t = run_t;
      best_prim = run_prim;
    }
  }
  if (unordered) {
    const float2 w = closest_hit_replay(S.blob, S.L, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.time, tmin, tmax);
    best_t = w.x;
    best_prim = __float_as_int(w.y);
  }
#undef TPT_SKIP_TOOK
  t_out = best_t;
  prim_out = best_prim;
  return best_prim >= 0;
}

// The same walk, nearest child first. The result of world->hit below the bvh_nodes is "smallest t, last in
// DFS order among equals" whatever the order the candidates are produced in, so the walk is free to visit
// the child the ray meets first (hint bits put into the blob's copy of the node words at upload: the axis
// along which the two children's boxes are furthest apart, and which of them is the lower one), keep the
// other on a small per-lane stack, and apply the DFS rule through the candidates' pre-order index. Found
// early, the near hit makes the "behind the best hit" test above refuse most of the far boxes.
// A hitable_list is still one sequential run against its own closest_so_far.
#define TPT_NODE_HINT 0x200       // node word: hint valid
#define TPT_NODE_HINT_LOWER 0x1000 // the first (left) child is the lower one along axis (kind >> 10) & 3
#define TPT_SKIP_STACK 32
TPT_DEV bool closest_hit_skip_ordered(const SceneView &S, const Ray &r, float tmin, float tmax, float &t_out, int &prim_out) {
  const float4 *N = S.blob + S.L->off_nodes;
  XRay x;
  x.chain = -1;
  float best_t = 0.f;
  int best_prim = -1, best_rank = -1;
  bool unordered = false;
  const float cull_abs = (r.time >= S.L->skip_cull_t0 && r.time <= S.L->skip_cull_t1) ? S.L->skip_cull_abs : -1.f;
  float box_max = tmax;
  int stack[TPT_SKIP_STACK];
  int sp = 0, cur = 0;
  if (S.L->n_nodes <= 0) return false;
  for (;;) {
    const float4 n0 = N[2 * cur], n1 = N[2 * cur + 1];
    const int kind = __float_as_int(n0.w), k = kind & 0xff, payload = __float_as_int(n1.w);
    int next = -1;
    float cand_t = 0.f;
    int cand_prim = -1;
    if (kind & TPT_NODE_DUP) {
      // second copy of a one-element bvh_node's child: the same record twice, nothing to add
    } else if (k == TPT_NODE_BVH) {
      to_chain<true>(S, r, kind >> 16, x);
      if (aabb_hit<true>(x, n0, n1, tmin, box_max)) {
        const int l = cur + 1;
        const int lkind = __float_as_int(N[2 * l].w);
        const int rgt = (lkind & 0xff) == TPT_NODE_LEAF ? l + 1 : __float_as_int(N[2 * l + 1].w);
        const int axis = (kind >> 10) & 3;
        const float da = axis == 0 ? x.d.x : (axis == 1 ? x.d.y : x.d.z);
        const bool right_first = (kind & TPT_NODE_HINT) && ((da >= 0.f) != ((kind & TPT_NODE_HINT_LOWER) != 0));
        stack[sp++] = right_first ? l : rgt;
        next = right_first ? rgt : l;
      }
    } else if (k == TPT_NODE_LEAF) {
      to_chain<true>(S, r, kind >> 16, x);
      float t;
      if (prim_test<true>(S, payload, x, r.time, tmin, tmax, t)) {
        cand_t = t;
        cand_prim = payload;
      }
    } else { // hitable_list (nested lists concatenate): one run against its own closest_so_far
      float run_t = tmax;
      for (int i = cur + 1; i < payload;) {
        const float4 m0 = N[2 * i];
        const int mkind = __float_as_int(m0.w), mk_ = mkind & 0xff;
        if (mk_ == TPT_NODE_LEAF) {
          if (!(mkind & TPT_NODE_DUP)) {
            to_chain<true>(S, r, mkind >> 16, x);
            const int prim = __float_as_int(N[2 * i + 1].w);
            float t;
            if (prim_test<true>(S, prim, x, r.time, tmin, run_t, t)) {
              run_t = t;
              cand_prim = prim;
            }
          }
          i++;
        } else if (mk_ == TPT_NODE_LIST && !(mkind & TPT_NODE_DUP)) {
          i++;
        } else {
          i = __float_as_int(N[2 * i + 1].w); // (tree_simple: no bvh_node below a list; a duplicate list adds nothing)
        }
      }
      cand_t = run_t;
    }
    if (cand_prim >= 0) {
      unordered = unordered || isnan(cand_t);
      if (best_prim < 0 || cand_t < best_t || (cand_t == best_t && cur > best_rank)) {
        best_t = cand_t;
        best_prim = cand_prim;
        best_rank = cur;
        if (TPT_PAR_SKIP_CULL && cull_abs >= 0.f) box_max = fminf(tmax, fmaf(best_t, 1.001f, cull_abs));
      }
    }
    if (next < 0) {
      if (sp == 0) break;
      next = stack[--sp];
    }
    cur = next;
  }
  if (unordered) {
    const float2 w = closest_hit_replay(S.blob, S.L, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.time, tmin, tmax);
    best_t = w.x;
    best_prim = __float_as_int(w.y);
  }
  t_out = best_t;
  prim_out = best_prim;
  return best_prim >= 0;
}

// world->hit: the root tree is nodes [0, n_nodes)
template <bool PAR>
TPT_DEV bool closest_hit(const SceneView &S, const Ray &r, float tmin, float tmax, float &t_out, int &prim_out,
                         Rng *g = nullptr) {
  if (PAR && TPT_PAR_SKIP_WALK && g == nullptr && S.L->tree_simple)
    return (TPT_PAR_SKIP_ORDERED && S.L->skip_ordered) ? closest_hit_skip_ordered(S, r, tmin, tmax, t_out, prim_out)
                                                       : closest_hit_skip(S, r, tmin, tmax, t_out, prim_out);
  return walk_range<PAR, true>(S, r, 0, S.L->n_nodes, tmin, tmax, t_out, prim_out, g);
}

// FAST mode: media are tested after the surfaces, in their DFS order, against the closest surface
// found (statistically identical to the reference's list order: a free-flight distance beyond a
// closer surface loses to it either way)
TPT_DEV void fast_media_pass(const SceneView &S, const Ray &r, float tmin, float &best, int &best_prim, Rng *g) {
  const int *ids = reinterpret_cast<const int *>(S.blob + S.L->off_mediums);
  XRay x;
  x.chain = -1;
  for (int m = 0; m < S.L->n_mediums; m++) {
    const int prim = ids[m];
    to_chain<false>(S, r, __float_as_int(S.blob[S.L->off_prims + 4 * prim].z), x);
    float t;
    if (medium_test<false>(S, prim, x, r, tmin, best, t, g)) {
      best = t;
      best_prim = prim;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Small scenes (Cornell: 19 primitives under a BVH that prunes almost nothing -- the reference
// itself tests 8.6 of its 9 boxes per ray): every lane tests every primitive in the same order.
// Control flow and shared-memory addresses are warp-uniform (broadcast loads, no divergence);
// only the "closer?" update is predicated. FAST mode only -- ties resolve to the first leaf.
// ------------------------------------------------------------------------------------------
// rect with compile-time axes: plane coordinate K, in-plane coordinates A and B (0=x,1=y,2=z)
template <int K, int A, int B>
TPT_DEV void small_rects(const SmallScene &Q, int begin, int end, const float (&o)[3], const float (&d)[3],
                         const float (&inv)[3], float tmin, float &best, int &best_prim) {
#pragma unroll kRectUnroll
  for (int i = begin; i < end; i++) {
    const float4 g = Q.geo[i];
    const float2 w = Q.aux[i];
    const float t = (w.x - o[K]) * inv[K];
    const float a = fmaf(t, d[A], o[A]);
    const float b = fmaf(t, d[B], o[B]);
    // t_min <= t <= t_max and inside the rectangle, bounds inclusive (src/rect_box.cc:11-16)
    const bool ok = (t >= tmin) & (t <= best) & (a >= g.x) & (a <= g.y) & (b >= g.z) & (b <= g.w);
    best = ok ? t : best;
    best_prim = ok ? __float_as_int(w.y) : best_prim;
  }
}

TPT_DEV bool closest_hit_uniform(const SceneView &S, const Ray &r, float tmin, float tmax, float &t_out,
                                 int &prim_out) {
  const SmallScene &Q = *S.small;
  float best = tmax;
  int best_prim = -1;
  const float inv_y = 1.0f / r.d.y; // translate / rotate_y leave the y component of a direction alone: one reciprocal for every group
  for (int gi = 0; gi < Q.n_groups; gi++) {
    const SmallGroup &G = Q.groups[gi];
    // ray into the group's space with the composed transform (headers/rect_box.h:89 and
    // src/rect_box.cc:175-179 applied in chain order, folded into one y-rotation + offset)
    const float o[3] = {G.cs * r.o.x - G.sn * r.o.z + G.bx, r.o.y + G.by, G.sn * r.o.x + G.cs * r.o.z + G.bz};
    const float d[3] = {G.cs * r.d.x - G.sn * r.d.z, r.d.y, G.sn * r.d.x + G.cs * r.d.z};
    const float inv[3] = {1.0f / d[0], inv_y, 1.0f / d[2]};
    small_rects<2, 0, 1>(Q, G.begin, G.xy_end, o, d, inv, tmin, best, best_prim);  // xy_rect: z = k
    small_rects<1, 0, 2>(Q, G.xy_end, G.xz_end, o, d, inv, tmin, best, best_prim); // xz_rect: y = k
    small_rects<0, 1, 2>(Q, G.xz_end, G.yz_end, o, d, inv, tmin, best, best_prim); // yz_rect: x = k
    for (int bi = G.box_begin; bi < G.box_end; bi++) {
      const SmallBox &B = Q.boxes[bi];
      const float x0 = (B.lo.x - o[0]) * inv[0], x1 = (B.hi.x - o[0]) * inv[0];
      const float y0 = (B.lo.y - o[1]) * inv[1], y1 = (B.hi.y - o[1]) * inv[1];
      const float z0 = (B.lo.z - o[2]) * inv[2], z1 = (B.hi.z - o[2]) * inv[2];
      const float nx = fminf(x0, x1), fx = fmaxf(x0, x1);
      const float ny = fminf(y0, y1), fy = fmaxf(y0, y1);
      const float nz = fminf(z0, z1), fz = fmaxf(z0, z1);
      const float tn = fmaxf(fmaxf(nx, ny), nz), tf = fminf(fminf(fx, fy), fz);
      // the closest face hit with t >= t_min is the entry point, or the exit when the origin is inside
      if (B.face[7]) {
        // open block (some faces absent, e.g. the five walls of a Cornell room): the only places a
        // ray can meet the faces are its entry and exit points of the block, in that order; a
        // crossing whose face is absent is skipped
        if ((tn <= tf) & (tf >= tmin)) {
          bool settled = false;
          if (tn >= tmin) {
            settled = tn > best; // both crossings lie behind the closest hit so far
            if (!settled) {
              const int axis = tn == nx ? 0 : (tn == ny ? 1 : 2);
              const float t_lo = axis == 0 ? x0 : (axis == 1 ? y0 : z0);
              const int f = B.face[2 * axis + (tn == t_lo ? 0 : 1)];
              if (f >= 0) {
                best = tn;
                best_prim = f;
                settled = true;
              }
            }
          }
          if (!settled && tf <= best) {
            const int axis = tf == fx ? 0 : (tf == fy ? 1 : 2);
            const float t_lo = axis == 0 ? x0 : (axis == 1 ? y0 : z0);
            const int f = B.face[2 * axis + (tf == t_lo ? 0 : 1)];
            if (f >= 0) {
              best = tf;
              best_prim = f;
            }
          }
        }
        continue;
      }
      const bool entry = tn >= tmin;
      const float t = entry ? tn : tf;
      const bool ok = (tn <= tf) & (t >= tmin) & (t <= best);
      if (ok) {
        const int axis = entry ? (tn == nx ? 0 : (tn == ny ? 1 : 2)) : (tf == fx ? 0 : (tf == fy ? 1 : 2));
        const float t_lo = axis == 0 ? x0 : (axis == 1 ? y0 : z0);
        best = t;
        best_prim = B.face[2 * axis + (t == t_lo ? 0 : 1)];
      }
    }
    if (G.yz_end < G.sph_end) {
      XRay x;
      x.o = mk(o[0], o[1], o[2]);
      x.d = mk(d[0], d[1], d[2]);
      for (int i = G.yz_end; i < G.sph_end; i++) {
        const float4 g = Q.geo[i];
        const float2 w = Q.aux[i]; // w.x != 0: huge sphere, exact roots
        float t;
        const bool hit = (w.x != 0.0f) ? sphere_test_exact_fast(mk(g.x, g.y, g.z), g.w, x, tmin, best, t)
                                       : sphere_test<false, false>(mk(g.x, g.y, g.z), g.w, x, tmin, best, t);
        if (hit) {
          best = t;
          best_prim = __float_as_int(w.y);
        }
      }
    }
  }
  t_out = best;
  prim_out = best_prim;
  return best_prim >= 0;
}

// ------------------------------------------------------------------------------------------
// Larger scenes, FAST mode: a binned-SAH BVH2 built by the library at scene upload over the
// WORLD-space boxes of the leaves (the reference's own tree is a random-axis median split that
// costs 54 box + 12 primitive tests per ray on random_scene). 64-byte nodes hold both
// children's boxes; traversal is per lane, nearest child first, t_max shrinking, short stack.
// Closest hit is independent of the acceleration structure (ties aside), so the estimator is
// unchanged; PARITY mode keeps replaying the reference's tree.
// ------------------------------------------------------------------------------------------
// ordinary spheres in FAST mode: contracted discriminant of the half-b form, MUFU roots
TPT_DEV bool sphere_test_quick(V3 center, float radius, const XRay &x, float tmin, float tmax, float &t) {
  const V3 oc = x.o - center;
  const float a = dot(x.d, x.d), hb = dot(oc, x.d), c = dot(oc, oc) - radius * radius;
  const float disc = hb * hb - a * c;
  if (disc > 0) {
    const float sq = sqrtf(disc), inv = 1.0f / a;
    float temp = (-hb - sq) * inv;
    if (temp < tmax && temp > tmin) {
      t = temp;
      return true;
    }
    temp = (-hb + sq) * inv;
    if (temp < tmax && temp > tmin) {
      t = temp;
      return true;
    }
  }
  return false;
}

// "while-while" traversal: every lane first descends to its next leaf, the warp reconverges, and
// then the leaf tests run together (the one-loop form ran them at 3-4 active lanes per warp).
// The state is a struct so that a kernel can interleave traversal steps of many rays with other
// work (render_mega_bvh_kernel postpones shading until enough lanes of the warp wait for it).
#define TPT_FBVH_DONE 0x7fffffff
#define TPT_FBVH_EMPTY 0x7ffffffe // BVH4: unused child slot
#ifndef TPT_FBVH_WIDE
#define TPT_FBVH_WIDE 0 // 1: four-wide nodes (collapse_to_bvh4 in tpt_api.cu), 0: the binary tree. Measured on B200 (profiles/r02_tuning_sweeps.txt): the per-lane BVH4 walk is 14 % SLOWER than the binary one on random_scene (1 798 vs 2 088 Mpaths/s; 1 888 with unsorted pushes) and 11 % on oneweek_final -- 28 live box words per node spill at 80 registers, and at 128 registers / 2 CTAs it is slower still (1 599)
#endif
#ifndef TPT_FBVH_SORT
#define TPT_FBVH_SORT 1 // BVH4: deferred children pushed far-to-near (0: in slot order)
#endif
#define TPT_FBVH_STACK (TPT_FBVH_WIDE ? 48 : 32)
#ifndef TPT_FBVH_LD256
#define TPT_FBVH_LD256 1 // 64-byte nodes / leaf records of the GLOBAL blob: two 256-bit loads (sm_100: LDG.E.256) instead of four 128-bit ones
#endif
// 32 bytes of the read-only scene blob, 32-byte aligned, global memory only
TPT_DEV void ld256(const float4 *p, float4 &lo, float4 &hi) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
      : "l"(p));
}
struct FbvhTrav {
  float ix, iy, iz, ox, oy, oz; // t = p * inv + (-o * inv)
  float best;
  int best_prim, node, sp;
  // the node stack is a separate local array handed to step(): inside the struct its dynamic
  // indexing kept the scalar members in local memory too (-9 % / -16 % measured)
  TPT_DEV void start(const Ray &r, float tmax) {
    ix = 1.0f / r.d.x; iy = 1.0f / r.d.y; iz = 1.0f / r.d.z;
    ox = -r.o.x * ix; oy = -r.o.y * iy; oz = -r.o.z * iz;
    best = tmax;
    best_prim = -1;
    node = 0;
    sp = 0;
  }
  TPT_DEV bool done() const { return node == TPT_FBVH_DONE; }
  // one inner node of the binary tree: both children's boxes, nearer child first
  TPT_DEV void inner_one(const float4 *N, float tmin, int *stack, bool wide) {
    float4 a, b, c, d;
    if (TPT_FBVH_LD256 && wide) {
      ld256(N + 4 * node, a, b);
      ld256(N + 4 * node + 2, c, d);
    } else {
      a = N[4 * node]; b = N[4 * node + 1]; c = N[4 * node + 2]; d = N[4 * node + 3];
    }
    // child 0: lo = (a.x,a.y,a.z) hi = (a.w,b.x,b.y) ; child 1: lo = (b.z,b.w,c.x) hi = (c.y,c.z,c.w)
    float t0x = fmaf(a.x, ix, ox), t1x = fmaf(a.w, ix, ox);
    float t0y = fmaf(a.y, iy, oy), t1y = fmaf(b.x, iy, oy);
    float t0z = fmaf(a.z, iz, oz), t1z = fmaf(b.y, iz, oz);
    float n0 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
    float f0 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best));
    t0x = fmaf(b.z, ix, ox); t1x = fmaf(c.y, ix, ox);
    t0y = fmaf(b.w, iy, oy); t1y = fmaf(c.z, iy, oy);
    t0z = fmaf(c.x, iz, oz); t1z = fmaf(c.w, iz, oz);
    float n1 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
    float f1 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best));
    const bool h0 = n0 <= f0, h1 = n1 <= f1;
    int i0 = __float_as_int(d.x), i1 = __float_as_int(d.y);
    if (h0 && h1) {
      if (n1 < n0) { int t = i0; i0 = i1; i1 = t; } // i0 = nearer
      stack[sp++] = i1;
      node = i0;
    } else if (h0) node = i0;
    else if (h1) node = i1;
    else node = sp > 0 ? stack[--sp] : TPT_FBVH_DONE;
  }
  // the primitives of the leaf in `node`, then the next deferred node
  TPT_DEV void leaf(const SceneView &S, const Ray &r, float tmin, int *stack) {
    // leaf: ~node = first << 3 | (count - 1); records of 4 float4 (tpt_api.cu, off_fleaf)
    const int code = ~node, count = (code & 7) + 1;
    const float4 *Q = S.blob + S.L->off_fleaf + 4 * (code >> 3);
    XRay x;
    x.chain = -1;
    for (int k = 0; k < count; k++, Q += 4) {
      float4 g, h;
      if (TPT_FBVH_LD256 && S.wide_loads) ld256(Q, g, h);
      else { g = Q[0]; h = Q[1]; }
      const int kf = __float_as_int(h.x), kind = kf & 0xff;
      to_chain<false>(S, r, __float_as_int(h.z), x);
      float t;
      bool hit;
      if (kind <= TPT_PRIM_MOVING_SPHERE) {
        V3 cen = mk(g.x, g.y, g.z);
        if (kind == TPT_PRIM_MOVING_SPHERE) {
          const float4 m = Q[2];
          cen = moving_center(g, make_float4(m.x, m.y, m.z, h.w), make_float4(m.w, 0.f, 0.f, 0.f), r.time);
        }
        hit = (kf & 0x100) ? sphere_test_exact_fast(cen, g.w, x, tmin, best, t)
                           : sphere_test_quick(cen, g.w, x, tmin, best, t);
      } else {
        hit = rect_test<false>(kind - TPT_PRIM_XY_RECT, g, h.w, x, tmin, best, t);
      }
      if (hit) {
        best = t;
        best_prim = __float_as_int(h.y);
      }
    }
    node = sp > 0 ? stack[--sp] : TPT_FBVH_DONE;
  }
  // descend to the next leaf (or run out of nodes), then test that leaf's primitives
  TPT_DEV void step(const SceneView &S, const Ray &r, float tmin, int *stack) {
    const float4 *N = S.blob + S.L->off_fbvh;
#if TPT_FBVH_WIDE
    while ((unsigned)node < (unsigned)TPT_FBVH_EMPTY) {
      const float4 *M = N + 8 * node;
      const float4 lx = M[0], ly = M[1], lz = M[2], hx = M[3], hy = M[4], hz = M[5];
      const int4 ch = *reinterpret_cast<const int4 *>(M + 6);
      // entry distance of child k, or "miss"; packed with the slot number in the two low mantissa bits
      // so that unsigned order = distance order (entry distances are >= tmin > 0)
      unsigned key[4];
      const float clx[4] = {lx.x, lx.y, lx.z, lx.w}, cly[4] = {ly.x, ly.y, ly.z, ly.w}, clz[4] = {lz.x, lz.y, lz.z, lz.w};
      const float chx[4] = {hx.x, hx.y, hx.z, hx.w}, chy[4] = {hy.x, hy.y, hy.z, hy.w}, chz[4] = {hz.x, hz.y, hz.z, hz.w};
      const int cref[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float t0x = fmaf(clx[k], ix, ox), t1x = fmaf(chx[k], ix, ox);
        const float t0y = fmaf(cly[k], iy, oy), t1y = fmaf(chy[k], iy, oy);
        const float t0z = fmaf(clz[k], iz, oz), t1z = fmaf(chz[k], iz, oz);
        const float nr = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
        const float fr = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best));
        const bool hit = (nr <= fr) & (cref[k] != TPT_FBVH_EMPTY);
        key[k] = hit ? ((__float_as_uint(nr) & ~3u) | (unsigned)k) : 0xffffffffu;
      }
#if TPT_FBVH_SORT
#define TPT_CE(a, b) { const unsigned lo_ = min(key[a], key[b]), hi_ = max(key[a], key[b]); key[a] = lo_; key[b] = hi_; }
      TPT_CE(0, 1) TPT_CE(2, 3) TPT_CE(0, 2) TPT_CE(1, 3) TPT_CE(1, 2)
#undef TPT_CE
      if (key[0] == 0xffffffffu) {
        node = sp > 0 ? stack[--sp] : TPT_FBVH_DONE;
        continue;
      }
      // farthest first, so that the nearest deferred child is popped first
#pragma unroll
      for (int k = 3; k >= 1; k--)
        if (key[k] != 0xffffffffu) {
          const unsigned slot = key[k] & 3u;
          stack[sp++] = slot == 0 ? cref[0] : (slot == 1 ? cref[1] : (slot == 2 ? cref[2] : cref[3]));
        }
      {
        const unsigned slot = key[0] & 3u;
        node = slot == 0 ? cref[0] : (slot == 1 ? cref[1] : (slot == 2 ? cref[2] : cref[3]));
      }
#else
      const unsigned nearest = min(min(key[0], key[1]), min(key[2], key[3]));
      if (nearest == 0xffffffffu) {
        node = sp > 0 ? stack[--sp] : TPT_FBVH_DONE;
        continue;
      }
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (key[k] != 0xffffffffu && key[k] != nearest) stack[sp++] = cref[k];
      {
        const unsigned slot = nearest & 3u;
        node = slot == 0 ? cref[0] : (slot == 1 ? cref[1] : (slot == 2 ? cref[2] : cref[3]));
      }
#endif
    }
#else
    while ((unsigned)node < (unsigned)TPT_FBVH_DONE) inner_one(N, tmin, stack, S.wide_loads);
#endif
    if (node < 0) leaf(S, r, tmin, stack);
  }
};

TPT_DEV bool closest_hit_fbvh(const SceneView &S, const Ray &r, float tmin, float tmax, float &t_out, int &prim_out) {
  FbvhTrav tv;
  int stack[TPT_FBVH_STACK];
  tv.start(r, tmax);
  while (!tv.done()) tv.step(S, r, tmin, stack);
  t_out = tv.best;
  prim_out = tv.best_prim;
  return tv.best_prim >= 0;
}

// Vote-scheduled traversal (binary tree). In the while-while form a warp's descent lasts as long as its
// slowest lane's: the number of inner nodes between two leaves is close to geometrically distributed, the
// maximum of 32 such draws is ~3.5 times their mean, and the descent loop of the BVH scenes ran at 9.8 of 32
// lanes, the leaf tests at 5-8 (profiles/r02_ncu_full_random_scene_lines.txt). Here a lane that has reached a
// leaf simply WAITS (node < 0 is its whole state) while the others keep descending, and the warp switches
// to the leaf code once no more than TPT_VOTE_NUM / (TPT_VOTE_NUM + TPT_VOTE_DEN) of its walking lanes still
// descend; both code paths then run with most of their lanes. Must be called by all 32 lanes (has_ray = the
// lane brought a ray). Measured (sweep 9 of profiles/r02_tuning_sweeps.txt): the descent's lane occupancy
// doubles, the kernel's time does not move on random_scene; the "trace" variant of the media scenes gains 5 %.
#ifndef TPT_FBVH_VOTE
#define TPT_FBVH_VOTE 0 // plain wavefront kernels of the SAH BVH scenes: 1 = vote-scheduled walk, 0 = per-lane while-while
#endif
#ifndef TPT_VOTE_NUM
#define TPT_VOTE_NUM 1
#endif
#ifndef TPT_VOTE_DEN
#define TPT_VOTE_DEN 2 // descend while more than NUM / (NUM + DEN) of the walking lanes still do
#endif
TPT_DEV bool closest_hit_fbvh_vote(const SceneView &S, const Ray &r, bool has_ray, float tmin, float tmax, float &t_out, int &prim_out) {
  FbvhTrav tv;
  int stack[TPT_FBVH_STACK];
  tv.start(r, tmax);
  if (!has_ray) tv.node = TPT_FBVH_DONE;
  const float4 *N = S.blob + S.L->off_fbvh;
  for (;;) {
    const int walking = __popc(__ballot_sync(0xffffffffu, !tv.done()));
    if (walking == 0) break;
    // descend while more than this many lanes still do; the others wait at their leaf (or are done)
    const int limit = walking * TPT_VOTE_NUM / (TPT_VOTE_NUM + TPT_VOTE_DEN);
    bool inner = (unsigned)tv.node < (unsigned)TPT_FBVH_DONE;
    while (__popc(__ballot_sync(0xffffffffu, inner)) > limit) {
      if (inner) tv.inner_one(N, tmin, stack, S.wide_loads);
      inner = (unsigned)tv.node < (unsigned)TPT_FBVH_DONE;
    }
    if (tv.node < 0) tv.leaf(S, r, tmin, stack);
  }
  t_out = tv.best;
  prim_out = tv.best_prim;
  return tv.best_prim >= 0;
}

// ------------------------------------------------------------------------------------------
// hit_record of the winning leaf, recomputed from (prim, t) with the leaf's own arithmetic
// ------------------------------------------------------------------------------------------
struct HitRec {
  float t, u, v;
  V3 p, n;
  int mat, prim;
};

template <bool PAR> TPT_DEV float round_atan2(float y, float x) {
  // std::atan2(float,float) -> atan2f. PAR evaluates in double and rounds once (closest
  // reproducible stand-in for glibc's atan2f; differences are <= 1 ulp).
  if (PAR) return (float)atan2((double)y, (double)x);
  return atan2f(y, x);
}
template <bool PAR> TPT_DEV float round_asin(float x) {
  if (PAR) return (float)asin((double)x);
  return asinf(x);
}

// headers/utils.h:38-43
template <bool PAR> TPT_DEV void get_uv_map(V3 p, float &u, float &v) {
  if (PAR) {
    u = (float)((double)round_atan2<PAR>(p.z, p.x) / (2 * 3.14159265358979323846));
    v = (float)((double)round_asin<PAR>(p.y) / 3.14159265358979323846);
  } else {
    u = round_atan2<PAR>(p.z, p.x) * 0.15915494309189535f;
    v = round_asin<PAR>(p.y) * 0.3183098861837907f;
  }
  u += 0.5f;
  v += 0.5f;
}

// LEAN (FAST kernels of small scenes whose textures are all constant_texture, e.g. the Cornell box):
// nobody reads (u, v), so the atan2f / asinf sequence is not even compiled in
template <bool PAR, int TEX = TPT_TEXF_ALL>
TPT_DEV void fill_hit(const SceneView &S, const Ray &r, int prim, float t, bool want_uv, HitRec &h) {
  const float4 *P = S.blob + S.L->off_prims + 4 * prim;
  float4 hd = P[0], a = P[1];
  int kind = __float_as_int(hd.x);
  int chain = __float_as_int(hd.z);
  int flags = __float_as_int(hd.w);
  h.t = t;
  h.prim = prim;
  h.mat = __float_as_int(hd.y);
  if (!PAR && !want_uv && kind != TPT_PRIM_MOVING_SPHERE) {
    // FAST, nobody reads (u, v): t is the same in every space, so the hit point is taken on the WORLD ray
    // and the normal from what tpt_scene_create stored in the record's last word (rects: the face normal
    // carried back to world space, flip included; spheres: the world-space centre, w = +-1 for the flip).
    // No trip into the primitive's transform chain and back (r01 capture: to_chain / from_chain ran at 8-9
    // of 32 lanes in the shade stage, 4.9 % of the kernel's warp instructions).
    const float4 w = P[3];
    h.p = r.o + t * r.d;
    h.u = 0.f;
    h.v = 0.f;
    if (kind == TPT_PRIM_SPHERE) h.n = (h.p - mk(w.x, w.y, w.z)) * (w.w / a.w);
    else h.n = mk(w.x, w.y, w.z);
    return;
  }
  XRay x;
  x.chain = -1;
  to_chain<PAR>(S, r, chain, x);
  h.p = x.o + t * x.d; // ray::point_at_parameter, headers/ray.h:13
  h.u = 0.f;
  h.v = 0.f;
  if (kind == TPT_PRIM_SPHERE || kind == TPT_PRIM_MOVING_SPHERE) {
    V3 c = mk(a.x, a.y, a.z);
    V3 cn = c;
    if (kind == TPT_PRIM_MOVING_SPHERE) cn = moving_center(a, P[2], P[3], r.time);
    h.n = (h.p - cn) / a.w;                                 // src/sphere.cc:28,59
    if ((TEX & TPT_TEXF_IMAGE) && want_uv) get_uv_map<PAR>((h.p - c) / a.w, h.u, h.v); // src/sphere.cc:27,60 (center0_)
  } else {
    float k = P[2].x;
    (void)k;
    if (kind == TPT_PRIM_XY_RECT) { // src/rect_box.cc:13-22
      h.u = (h.p.x - a.x) / (a.y - a.x);
      h.v = (h.p.y - a.z) / (a.w - a.z);
      h.n = mk(0, 0, 1);
    } else if (kind == TPT_PRIM_XZ_RECT) { // src/rect_box.cc:57-66
      h.u = (h.p.x - a.x) / (a.y - a.x);
      h.v = (h.p.z - a.z) / (a.w - a.z);
      h.n = mk(0, 1, 0);
    } else { // src/rect_box.cc:80-89
      h.u = (h.p.y - a.x) / (a.y - a.x);
      h.v = (h.p.z - a.z) / (a.w - a.z);
      h.n = mk(1, 0, 0);
    }
  }
  if (flags & TPT_PRIM_FLIP) h.n = -h.n; // headers/rect_box.h:53 (negation commutes with rotate_y)
  from_chain(S, chain, h.p, h.n);
}

// ------------------------------------------------------------------------------------------
// Textures (src/texture.cc, src/utils.cc:160-225)
// ------------------------------------------------------------------------------------------
template <bool PAR> TPT_DEV float round_sin(float x) {
  if (PAR) return (float)sin((double)x); // std::sin(float) -> sinf, correctly-rounded stand-in
  // FAST: MUFU.SIN is only accurate near the origin (absolute error grows with |x|; texture arguments
  // reach 10 * coordinate). Two-constant reduction to [-pi, pi] first: 2 pi = 6.2831855f - 1.7484556e-7f.
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, x);
  r = fmaf(k, 1.7484555314695172e-7f, r);
  return __sinf(r);
}
template <bool PAR> TPT_DEV float round_cos(float x) {
  if (PAR) return (float)cos((double)x);
  return __cosf(x);
}

template <bool PAR> TPT_DEV float perlin_fade(float t) {
  // src/utils.cc:207-210: 6 t^5 - 15 t^4 + 10 t^3 with std::pow(float,int) -> double
  if (PAR) {
    double d = (double)t;
    double d3 = d * d * d, d4 = d3 * d, d5 = d4 * d;
    return (float)(6 * d5 - 15 * d4 + 10 * d3);
  }
  return t * t * t * (t * (t * 6.f - 15.f) + 10.f);
}

template <bool PAR> TPT_DEV float perlin_noise_at(const SceneView &S, V3 p) {
  // src/utils.cc:160-191 + perlin_interpolate :205-225
  float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
  float u = p.x - fx, v = p.y - fy, w = p.z - fz;
  int i = ((int)fx) & 255, j = ((int)fy) & 255, k = ((int)fz) & 255;
  const float4 *ranvec = S.blob + S.L->off_perlin;
  const int *perm = reinterpret_cast<const int *>(S.blob + S.L->off_perlin + 256);
  float uu = perlin_fade<PAR>(u), vv = perlin_fade<PAR>(v), ww = perlin_fade<PAR>(w);
  float accum = 0.0f;
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        int idx = perm[(i + a) & 255] ^ perm[256 + ((j + b) & 255)] ^ perm[512 + ((k + c) & 255)];
        float4 g = ranvec[idx];
        V3 wv = mk(uu - a, vv - b, ww - c);
        float wa = a * uu + (1 - a) * (1 - uu);
        float wb = b * vv + (1 - b) * (1 - vv);
        float wc = c * ww + (1 - c) * (1 - ww);
        accum += wa * wb * wc * dot(mk(g.x, g.y, g.z), wv);
      }
  return fabsf(accum);
}

template <bool PAR> TPT_DEV float perlin_turb(const SceneView &S, V3 p) {
  // src/utils.cc:193-203, depth 5
  float accum = 0, weight = 1.0f;
  V3 tmp = p;
  for (int i = 0; i < 5; i++) {
    accum += weight * perlin_noise_at<PAR>(S, tmp);
    weight *= 0.5f;
    tmp = tmp * 2.0f;
  }
  return fabsf(accum);
}

// TEX = the texture features of the scene (checked at upload): TPT_TEXF_IMAGE (image textures: uv of
// sphere hits + texel fetch), TPT_TEXF_PROCEDURAL (checker, Perlin marble). A kernel build carries only
// the code of the features its scene has; TEX == 0 (every texture a constant_texture, e.g. the Cornell
// box) leaves out a third of the kernel's instructions, never executed on such a scene.
// The hot loop of the render kernels is about as large as the SM's 32 KB instruction cache (ncu on
// the Cornell frame: 92.8 % hit rate and the GPC's instruction fetch at 79 % of its peak with
// everything compiled in; 97.5 % and 43 % without the dead code), so what a scene cannot reach is
// kept out of its kernel.
template <bool PAR, int TEX = TPT_TEXF_ALL> TPT_DEV V3 texture_value(const SceneView &S, int tex, float u, float v, V3 p) {
  if (TEX == 0) {
    const float4 t0 = S.blob[S.L->off_texs + 2 * tex];
    return mk(t0.y, t0.z, t0.w);
  }
  for (int guard = 0; guard < 8; guard++) {
    float4 t0 = S.blob[S.L->off_texs + 2 * tex], t1 = S.blob[S.L->off_texs + 2 * tex + 1];
    int kind = __float_as_int(t0.x);
    if (kind == TPT_TEX_CONSTANT) return mk(t0.y, t0.z, t0.w);
    if ((TEX & TPT_TEXF_PROCEDURAL) && kind == TPT_TEX_CHECKER) { // src/texture.cc:4-16
      float s = round_sin<PAR>(10 * p.x) * round_sin<PAR>(10 * p.y) * round_sin<PAR>(10 * p.z);
      tex = (s < 0.0f) ? __float_as_int(t1.x) : __float_as_int(t1.y);
      continue;
    }
    if ((TEX & TPT_TEXF_PROCEDURAL) && kind == TPT_TEX_PERLIN) { // src/texture.cc:18-25
      float scale = t1.z;
      float s = round_sin<PAR>(scale * p.z + 10 * perlin_turb<PAR>(S, p));
      float g = 0.5f * (1 + s);
      return mk(g, g, g);
    }
    if (TEX & TPT_TEXF_IMAGE) {
      // TPT_TEX_IMAGE: src/texture.cc:27-42, nearest texel, clamp, /255.0f
      int img = __float_as_int(t1.w);
      int W = S.L->image_w[img], H = S.L->image_h[img];
      int i = (int)(u * W);
      int j = (int)((1 - v) * H);
      if (i < 0) i = 0;
      if (i > W - 1) i = W - 1;
      if (j > H - 1) j = H - 1;
      if (j < 0) j = 0;
      uchar4 c = tex2D<uchar4>(S.L->images[img], (float)i, (float)j);
      return mk((int)c.x / 255.0f, (int)c.y / 255.0f, (int)c.z / 255.0f);
    }
    break; // a texture kind this build was not compiled for: tpt_scene_create never selects it for such a scene
  }
  return mk(0, 0, 0);
}

TPT_DEV bool texture_needs_uv(const SceneView &S, int tex) {
  // only image textures read (u,v); checker children can be images
  for (int guard = 0; guard < 4; guard++) {
    float4 t0 = S.blob[S.L->off_texs + 2 * tex];
    int kind = __float_as_int(t0.x);
    if (kind == TPT_TEX_IMAGE) return true;
    if (kind == TPT_TEX_CHECKER) return true; // conservative
    return false;
  }
  return false;
}

// ------------------------------------------------------------------------------------------
// Sampling helpers
// ------------------------------------------------------------------------------------------
struct Onb {
  V3 u, v, w;
};
template <bool PAR> TPT_DEV Onb onb_from_w(V3 n) {
  // src/utils.cc:437-450. `std::abs(normal.x()) > 0.9` compares in double.
  Onb b;
  b.w = n;
  bool x_axis = PAR ? ((double)fabsf(n.x) > 0.9) : (fabsf(n.x) > 0.9f);
  V3 tmp = x_axis ? mk(0, 1, 0) : mk(1, 0, 0);
  b.v = unit<PAR>(cross(n, tmp));
  b.u = cross(b.v, b.w);
  return b;
}
TPT_DEV V3 onb_local(const Onb &b, float x, float y, float z) { return x * b.u + y * b.v + z * b.w; }

#define TPT_PI_D 3.14159265358979323846
#define TPT_PI_F 3.14159265358979323846f

// src/utils.cc:427-435
template <bool PAR> TPT_DEV V3 random_on_hemisphere(Rng &g) {
  float r1 = g.next();
  float r2 = g.next();
  float phi = PAR ? (float)(2 * TPT_PI_D * (double)r1) : (2.f * TPT_PI_F) * r1;
  float sr = sqrtf(r2);
  float x, y;
  if (PAR) {
    // std::cos(float) / std::sin(float): one double sincos (shared argument reduction, half the code
    // of two calls), each result rounded once to float
    double sd, cd;
    sincos((double)phi, &sd, &cd);
    x = (float)cd * sr;
    y = (float)sd * sr;
  } else {
    float s, c;
    __sincosf(phi, &s, &c);
    x = c * sr;
    y = s * sr;
  }
  float z = sqrtf(1 - r2);
  return mk(x, y, z);
}

// src/utils.cc:21-27 (g++ evaluates the ctor arguments right to left: z, y, x)
TPT_DEV V3 random_in_unit_sphere(Rng &g) {
  V3 p;
  do {
    float z = g.next();
    float y = g.next();
    float x = g.next();
    p = 2.0f * mk(x, y, z) - mk(1.0f, 1.0f, 1.0f);
  } while (length(p) >= 1.0f);
  return p;
}

// hitable_list::random -> xz_rect::random | sphere::random
// (src/hitable_list.cc:33-36, src/rect_box.cc:39-43, src/sphere.cc:108-120)
template <bool PAR> TPT_DEV V3 light_random(const SceneView &S, V3 origin, Rng &g) {
  float pick = g.next();
  int index = PAR ? (int)((double)pick * (double)S.L->n_lights) : (int)(pick * (float)S.L->n_lights);
  float4 l0 = S.blob[S.L->off_lights + 2 * index], l1 = S.blob[S.L->off_lights + 2 * index + 1];
  int kind = __float_as_int(l0.x);
  if (kind == TPT_LIGHT_XZ_RECT) {
    float x0 = l0.y, x1 = l0.z, z0 = l0.w, z1 = l1.x, k = l1.y;
    float rz = g.next(); // third ctor argument first
    float rx = g.next();
    V3 pt;
    if (PAR) {
      pt = mk((float)((double)x0 + (double)rx * (double)(x1 - x0)), k,
              (float)((double)z0 + (double)rz * (double)(z1 - z0)));
    } else {
      pt = mk(x0 + rx * (x1 - x0), k, z0 + rz * (z1 - z0));
    }
    return pt - origin;
  } else if (kind == TPT_LIGHT_SPHERE) {
    V3 c = mk(l0.y, l0.z, l0.w);
    float radius = l1.x;
    V3 direction = c - origin;
    Onb uvw = onb_from_w<PAR>(unit<PAR>(direction));
    float tmp = (radius * radius) / sqlen(c - origin);
    float cmax = sqrtf(1 - tmp); // NaN when the origin is inside the sphere
    float r1 = g.next();
    float r2 = g.next();
    float z = 1 + r2 * (cmax - 1);
    float sq = sqrtf(1 - z * z);
    float x, y;
    if (PAR) {
      double sd, cd;
      sincos(2 * TPT_PI_D * (double)r1, &sd, &cd);
      x = (float)(cd * (double)sq);
      y = (float)(sd * (double)sq);
    } else {
      float s, co;
      __sincosf((2.f * TPT_PI_F) * r1, &s, &co);
      x = co * sq;
      y = s * sq;
    }
    return onb_local(uvw, x, y, z);
  }
  return mk(1, 0, 0); // hitable::random, headers/hitable.h:38
}

// hitable_list::pdf_value (src/hitable_list.cc:24-31) over xz_rect::pdf_value
// (src/rect_box.cc:26-37) and sphere::pdf_value (src/sphere.cc:93-106)
template <bool PAR> TPT_DEV float light_pdf(const SceneView &S, V3 origin, V3 dir) {
  float weight = PAR ? (float)(1.0 / (double)S.L->n_lights) : 1.0f / (float)S.L->n_lights;
  float sum = 0;
  XRay x;
  x.o = origin;
  x.d = dir;
  x.chain = 0;
  for (int i = 0; i < S.L->n_lights; i++) {
    float4 l0 = S.blob[S.L->off_lights + 2 * i], l1 = S.blob[S.L->off_lights + 2 * i + 1];
    int kind = __float_as_int(l0.x);
    float pdf = 0.0f;
    if (kind == TPT_LIGHT_XZ_RECT) {
      float t;
      float4 p = make_float4(l0.y, l0.z, l0.w, l1.x);
      if (rect_test<PAR>(1, p, l1.y, x, 0.0001f, FLT_MAX, t)) {
        float area = fabsf((p.y - p.x) * (p.w - p.z));
        float dist2 = sqlen(t * dir);
        float cosine = fabsf(dot(dir, mk(0, 1, 0)) / length(dir));
        pdf = dist2 / (cosine * area);
      }
    } else if (kind == TPT_LIGHT_SPHERE) {
      float t;
      V3 c = mk(l0.y, l0.z, l0.w);
      float radius = l1.x;
      bool sph_hit;
      if (PAR) sph_hit = sphere_test<PAR, true>(c, radius, x, 0.001f, FLT_MAX, t);
      else if (radius >= 500.0f) sph_hit = sphere_test_exact_fast(c, radius, x, 0.001f, FLT_MAX, t);
      else sph_hit = sphere_test<PAR, false>(c, radius, x, 0.001f, FLT_MAX, t);
      if (sph_hit) {
        float tmp = (radius * radius) / sqlen(c - origin);
        float cmax = sqrtf(1 - tmp);
        float solid = PAR ? (float)(2 * TPT_PI_D * (double)(1 - cmax)) : (2.f * TPT_PI_F) * (1 - cmax);
        pdf = isnan(solid) ? 0.0f : 1 / solid;
      }
    }
    sum += weight * pdf;
  }
  return sum;
}

// ------------------------------------------------------------------------------------------
// Optics (src/utils.cc:34-56)
// ------------------------------------------------------------------------------------------
TPT_DEV V3 reflect(V3 v, V3 n) { return v - 2 * dot(v, n) * n; }

template <bool PAR> TPT_DEV bool refract(V3 v, V3 n, float ni_over_nt, V3 &refracted) {
  V3 uv = unit<PAR>(v);
  float dt = dot(uv, n);
  float disc = PAR ? (float)(1.0 - (double)(ni_over_nt * ni_over_nt * (1 - dt * dt)))
                   : 1.0f - ni_over_nt * ni_over_nt * (1 - dt * dt);
  if (disc > 0) {
    refracted = ni_over_nt * (uv - n * dt) - n * sqrtf(disc);
    return true;
  }
  return false;
}

template <bool PAR> TPT_DEV float schlick(float cosine, float ref_index) {
  float r0 = (1 - ref_index) / (1 + ref_index);
  r0 = r0 * r0;
  if (PAR) {
    // pow(double, 5): four double multiplications are within 2 ulp(double) of the exact power, like
    // libm's pow -- the float the sum rounds to is the same except on ~2^-28 of the inputs -- at a
    // twentieth of the code of the general pow
    const double m = (double)(1 - cosine), m2 = m * m;
    return (float)((double)r0 + (double)(1 - r0) * (m2 * m2 * m));
  }
  float m = 1 - cosine;
  float m2 = m * m;
  return r0 + (1 - r0) * (m2 * m2 * m);
}

// ------------------------------------------------------------------------------------------
// Camera (src/camera.cc:23-31) and pixel jitter (main.cpp:121-122)
// ------------------------------------------------------------------------------------------
struct CamView {
  V3 origin, llc, vertical, horizontal, u, v, w;
  float lens_radius, time0, time1;
};

TPT_DEV float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; } // Rng::next()'s mapping

template <bool PAR>
TPT_DEV Ray camera_sample(const CamView &C, int i, int j, int nx, int ny, Rng &g) {
#if TPT_CAMERA_BLOCK_LOOP
  // The camera stage's draws in stream order: jitter u, v; lens pairs (y, x) until one lies in the unit
  // disk (random_in_unit_disk, src/utils.cc:13-19: vec3(drand_r(), drand_r(), 0) -> y drawn first); the
  // shutter time when the shutter is open. They are taken block by block from ONE expansion site of the
  // Philox function inside a loop: every lane runs the trip for block 0 (u, v, first pair), the 21 % whose
  // first lens sample falls outside the disk -- or every lane, with an open shutter -- run another.
  // Same draws at the same stream positions as g.next() would deliver (pairs never straddle a block).
  // Meant to replace a block expansion plus an out-of-line refill that runs at 4.6 of 32 lanes in nearly
  // every camera chunk (r02 capture: 4.3 % of the fast kernel's warp instructions, + 2.1 % in next()'s
  // refill test); measured, it does not pay (see the macro) -- kept as the alternative it was tested as.
  const bool need_time = C.time1 != C.time0;
  float r_u = 0.f, r_v = 0.f, px = 0.f, py = 0.f, rt = 0.f;
  int phase = 0; // 0: u, v and the first pair; 1: another pair; 2: the shutter time; 3: done
  for (uint32_t blk = 0;; blk++) {
    uint32_t o[4];
    philox4x32_10_rk(g.pixel, g.sample, 0u, blk, g.rk, o);
    int k = 0; // draws of this block consumed so far
    if (phase == 0) {
      r_u = u01(o[0]);
      r_v = u01(o[1]);
      py = 2.0f * u01(o[2]) - 1.0f;
      px = 2.0f * u01(o[3]) - 1.0f;
      k = 4;
      phase = (px * px + py * py >= 1.0f) ? 1 : (need_time ? 2 : 3);
    } else {
      if (phase == 1) {
        py = 2.0f * u01(o[0]) - 1.0f;
        px = 2.0f * u01(o[1]) - 1.0f;
        k = 2;
        if (px * px + py * py >= 1.0f) {
          py = 2.0f * u01(o[2]) - 1.0f;
          px = 2.0f * u01(o[3]) - 1.0f;
          k = 4;
        }
        phase = (px * px + py * py >= 1.0f) ? 1 : (need_time ? 2 : 3);
      }
      if (phase == 2 && k < 4) {
        rt = u01(k == 0 ? o[0] : o[2]);
        phase = 3;
      }
    }
    if (phase == 3) break;
  }
  g.stage = 0u;
  float s, t;
  if (PAR) {
    s = (float)(((double)(float)i + (double)r_u) / (double)(float)nx);
    t = (float)(((double)(float)j + (double)r_v) / (double)(float)ny);
  } else {
    s = ((float)i + r_u) / (float)nx;
    t = ((float)j + r_v) / (float)ny;
  }
#else
  g.set_stage(0u);
  float r_u = g.next();
  float r_v = g.next();
  float s, t;
  if (PAR) {
    s = (float)(((double)(float)i + (double)r_u) / (double)(float)nx);
    t = (float)(((double)(float)j + (double)r_v) / (double)(float)ny);
  } else {
    s = ((float)i + r_u) / (float)nx;
    t = ((float)j + r_v) / (float)ny;
  }
  // random_in_unit_disk, src/utils.cc:13-19: vec3(drand_r(), drand_r(), 0) -> y drawn first
  float px, py;
  {
    float y = g.next();
    float x = g.next();
    px = 2.0f * x - 1.0f;
    py = 2.0f * y - 1.0f;
  }
  if (TPT_EAGER_CAMERA_BLOCK == 1 || (TPT_EAGER_CAMERA_BLOCK == 2 && PAR)) g.next_block_inline();
  while (px * px + py * py >= 1.0f) {
    float y = g.next();
    float x = g.next();
    px = 2.0f * x - 1.0f;
    py = 2.0f * y - 1.0f;
  }
  const bool need_time = C.time1 != C.time0;
  float rt = need_time ? g.next() : 0.f;
#endif
  float rdx = C.lens_radius * px, rdy = C.lens_radius * py;
  V3 offset = C.u * rdx + C.v * rdy;
  Ray r;
  if (need_time) {
    r.time = PAR ? (float)((double)C.time0 + (double)rt * (double)(C.time1 - C.time0))
                 : C.time0 + rt * (C.time1 - C.time0);
  } else {
    r.time = C.time0; // time0 + drand*0: the draw is the last of its stage, skipping it is exact
  }
  r.o = C.origin + offset;
  r.d = C.llc + s * C.horizontal + t * C.vertical - C.origin - offset;
  return r;
}

// ------------------------------------------------------------------------------------------
// One bounce of color() (src/utils.cc:58-94), iterative form.
// Path state: ray, throughput T (product of attenuation*scattering_pdf/pdf_value of the
// diffuse bounces and attenuation of the specular ones), depth. Emission is non-zero only on
// materials that never scatter (diffuse_light), so `emitted + w*color(next)` collapses to
// "radiance = T * emitted at the terminal vertex"; a path whose T is 0-or-NaN in every
// channel can stop: every continuation yields 0 after de_nan (headers/utils.h:100-109).
// Returns true while the path continues.
// ------------------------------------------------------------------------------------------
struct PathState {
  Ray ray;
  V3 T;
  int depth;
};

TPT_DEV bool dead_channel(float t) { return t == 0.0f || isnan(t); }

// May a ray hit anything at all? Used on freshly generated camera rays so that the ones that leave
// the scene's bounds (64 % of the headline frame: the camera sees past the box) never enter the
// extend queue. PARITY: only when the root is a bvh_node, with the reference's own first test
// (bvh_node::hit -> box_.hit(r, t_min, FLT_MAX), src/hitable.cc:65) -- exactly the rays for which
// world->hit returns false there. FAST: the root's bounds, padded, NaN-safe (a NaN keeps the ray).
template <bool PAR> TPT_DEV bool may_hit_world(const SceneView &S, const Ray &r, float t_min) {
  const float4 n0 = S.blob[S.L->off_nodes], n1 = S.blob[S.L->off_nodes + 1];
  const int kind = __float_as_int(n0.w);
  if (PAR) {
    if ((kind & 0xff) != TPT_NODE_BVH || (kind >> 16) != 0) return true;
    XRay x;
    x.o = r.o;
    x.d = r.d;
    x.inv = mk(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    x.chain = 0;
    return aabb_hit<true>(x, n0, n1, t_min, FLT_MAX);
  }
  if ((kind >> 16) != 0) return true;
  const float px = 1e-3f * fmaxf(fabsf(n0.x), fabsf(n1.x)) + 1e-3f, py = 1e-3f * fmaxf(fabsf(n0.y), fabsf(n1.y)) + 1e-3f;
  const float pz = 1e-3f * fmaxf(fabsf(n0.z), fabsf(n1.z)) + 1e-3f;
  const float ix = 1.0f / r.d.x, iy = 1.0f / r.d.y, iz = 1.0f / r.d.z;
  const float ax = (n0.x - px - r.o.x) * ix, bx = (n1.x + px - r.o.x) * ix;
  const float ay = (n0.y - py - r.o.y) * iy, by = (n1.y + py - r.o.y) * iy;
  const float az = (n0.z - pz - r.o.z) * iz, bz = (n1.z + pz - r.o.z) * iz;
  const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t_min));
  const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
  return !(t0 > t1); // any NaN compares false -> keep the ray
}

// radiance of a ray that leaves the scene (src/utils.cc:85-90)
template <bool PAR> TPT_DEV V3 background_radiance(const SceneView &S, const Ray &r, V3 T) {
  if (S.L->background != TPT_BG_SKY) return mk(0, 0, 0);
  V3 ud = unit<PAR>(r.d);
  float tt = PAR ? (float)(((double)ud.y + 1.0) * 0.5) : (ud.y + 1.0f) * 0.5f;
  V3 sky = 0.1f * ((1 - tt) * mk(1.0f, 1.0f, 1.0f) + tt * mk(0.5f, 0.7f, 1.0f));
  return T * sky;
}

// FAST mode without a usable SAH BVH (a shutter interval outside the one the moving spheres' boxes
// were built for, or a tree the builder gave up on): the culled walk over the reference's own tree.
// Rare, and several KB of code: kept out of line so that the SAH-BVH kernels do not carry it in
// their hot loop's instruction-cache lines.
static __device__ __noinline__ float2 closest_hit_reftree_slow(const float4 *blob, const SceneLayout *L, float ox, float oy, float oz,
                                                               float dx, float dy, float dz, float time, float tmin, float tmax) {
  SceneView S;
  S.blob = blob;
  S.L = L;
  S.small = nullptr;
  S.flat = nullptr;
  Ray r;
  r.o = mk(ox, oy, oz);
  r.d = mk(dx, dy, dz);
  r.time = time;
  float t = tmax;
  int prim = -1;
  walk_range<false, false>(S, r, 0, L->n_nodes, tmin, tmax, t, prim, nullptr);
  return make_float2(t, __int_as_float(prim));
}

// extend(): world->hit plus every outcome that ends the path without a scatter() call
// (src/utils.cc:61-66,82-86). Returns TPT_EXT_DONE with the sample's radiance, or the kind of the
// scattering material (LAMBERTIAN / METAL / DIELECTRIC) with (prim, t) of the hit.
#define TPT_EXT_DONE (-1)
// MEDIA is a compile-time switch: only kernels instantiated for scenes with participating media
// carry the stream through world->hit (taking the Rng's address costs registers everywhere else).
// Everything extend() does once the closest surface hit (any_hit, t, prim) is known: the media
// pass of FAST mode, the miss / lamp / absorber / depth-limit endings, else the material kind.
template <bool PAR, bool MEDIA, int TEX = TPT_TEXF_ALL>
TPT_DEV int extend_finish(const SceneView &S, const PathState &ps, int max_depth, float t_min, bool any_hit, float &t,
                          int &prim, V3 &radiance, Rng &g, uint32_t &ndraw_out) {
  if (!PAR && MEDIA) {
    if (!any_hit) {
      t = FLT_MAX;
      prim = -1;
    }
    fast_media_pass(S, ps.ray, t_min, t, prim, &g);
    any_hit = prim >= 0;
  }
  ndraw_out = MEDIA ? g.ndraw : 0u;
  radiance = mk(0, 0, 0);
  if (!any_hit) {
    radiance = background_radiance<PAR>(S, ps.ray, ps.T); // black at HEAD, or the commented gradient
    return TPT_EXT_DONE;
  }
  // tpt_material = {kind, texture, albedo[3], fuzz, ref_idx, pad}
  const int mat = __float_as_int(S.blob[S.L->off_prims + 4 * prim].y);
  const float4 m0 = S.blob[S.L->off_mats + 2 * mat];
  const int mkind = __float_as_int(m0.x);
  if (mkind == TPT_MAT_DIFFUSE_LIGHT) {
    // src/material.cc:79-86: one-sided; base scatter() is false -> return emitted
    const int mtex = __float_as_int(m0.y);
    HitRec h;
    fill_hit<PAR, TEX>(S, ps.ray, prim, t, (TEX & TPT_TEXF_IMAGE) && texture_needs_uv(S, mtex), h);
    if (dot(h.n, ps.ray.d) < 0) radiance = ps.T * texture_value<PAR, TEX>(S, mtex, h.u, h.v, h.p);
    return TPT_EXT_DONE;
  }
  if (mkind == TPT_MAT_ABSORBER || mkind == TPT_MAT_ISOTROPIC || ps.depth >= max_depth) return TPT_EXT_DONE; // emitted == 0
  return mkind;
}

template <bool PAR, bool SMALL, bool MEDIA, int TEX = TPT_TEXF_ALL>
TPT_DEV int extend(const SceneView &S, const PathState &ps, int max_depth, float t_min, float &t, int &prim,
                   V3 &radiance, Rng &g, uint32_t &ndraw_out) {
  // scenes with participating media draw inside world->hit: the stage's stream starts here
  Rng *gp = nullptr;
  if (MEDIA) {
    g.set_stage((uint32_t)ps.depth + 1u);
    gp = &g;
  }
  bool any_hit;
  if constexpr (PAR) {
    if constexpr (SMALL && !MEDIA) any_hit = closest_hit_flat(S, ps.ray, t_min, FLT_MAX, t, prim);
    else any_hit = closest_hit<PAR>(S, ps.ray, t_min, FLT_MAX, t, prim, gp);
  } else {
    if (SMALL) any_hit = closest_hit_uniform(S, ps.ray, t_min, FLT_MAX, t, prim);
    else if (S.L->n_fbvh > 0 && S.L->fbvh_time_ok) any_hit = closest_hit_fbvh(S, ps.ray, t_min, FLT_MAX, t, prim);
    else {
      const float2 w = closest_hit_reftree_slow(S.blob, S.L, ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, ps.ray.d.x, ps.ray.d.y, ps.ray.d.z,
                                                ps.ray.time, t_min, FLT_MAX);
      t = w.x;
      prim = __float_as_int(w.y);
      any_hit = prim >= 0;
    }
  }
  return extend_finish<PAR, MEDIA, TEX>(S, ps, max_depth, t_min, any_hit, t, prim, radiance, g, ndraw_out);
}

// shade(): material::scatter + the mixture-pdf step of color() for the hit (prim, t).
// Returns true while the path continues (ps holds the next ray, throughput, depth).
template <bool PAR, int TEX = TPT_TEXF_ALL>
TPT_DEV bool shade(const SceneView &S, PathState &ps, Rng &g, int prim, float t, uint32_t ndraw0) {
  // stage d+1 = the draws color() makes at depth d; ndraw0 of them were already taken inside
  // world->hit by participating media (0 in scenes without any)
  if (ndraw0 == 0u) g.set_stage((uint32_t)ps.depth + 1u);
  else g.set_stage_at((uint32_t)ps.depth + 1u, ndraw0);
  const int mat = __float_as_int(S.blob[S.L->off_prims + 4 * prim].y);
  const float4 m0 = S.blob[S.L->off_mats + 2 * mat];
  const float4 m1 = S.blob[S.L->off_mats + 2 * mat + 1];
  const int mkind = __float_as_int(m0.x);
  const int mtex = __float_as_int(m0.y);
  bool want_uv = (TEX & TPT_TEXF_IMAGE) && mkind == TPT_MAT_LAMBERTIAN && texture_needs_uv(S, mtex);
  HitRec h;
  fill_hit<PAR, TEX>(S, ps.ray, prim, t, want_uv, h);
  if (mkind == TPT_MAT_METAL) { // src/material.cc:88-98
    V3 reflected = reflect(unit<PAR>(ps.ray.d), h.n);
    V3 dir = reflected + m1.y * random_in_unit_sphere(g); // fuzz_
    if (!(dot(dir, h.n) > 0)) return false;
    ps.T = ps.T * mk(m0.z, m0.w, m1.x); // attenuation = albedo_, no emitted term (src/utils.cc:67-72)
    ps.ray.o = h.p;
    ps.ray.d = dir;
  } else if (mkind == TPT_MAT_DIELECTRIC) { // src/material.cc:19-70
    float ref_idx = m1.z;
    V3 d = ps.ray.d;
    V3 reflected = reflect(d, h.n);
    V3 outward;
    float ni_over_nt, cosine;
    float ddn = dot(d, h.n);
    if (ddn > 0) {
      outward = -h.n;
      ni_over_nt = ref_idx;
      cosine = ddn / length(d);
      cosine = sqrtf(1 - ref_idx * ref_idx * (1 - cosine * cosine));
    } else {
      outward = h.n;
      ni_over_nt = PAR ? (float)(1.0 / (double)ref_idx) : 1.0f / ref_idx;
      cosine = -ddn / length(d);
    }
    V3 refracted;
    float reflect_prob;
    if (refract<PAR>(d, outward, ni_over_nt, refracted))
      reflect_prob = schlick<PAR>(cosine, ref_idx);
    else
      reflect_prob = 1.0f;
    float xi = g.next();
    ps.ray.o = h.p;
    ps.ray.d = (xi < reflect_prob) ? reflected : refracted; // attenuation (1,1,1)
  } else {
    // lambertian: src/material.cc:3-17 + mixture sampling src/utils.cc:73-81
    V3 atten = texture_value<PAR, TEX>(S, mtex, h.u, h.v, h.p);
    Onb uvw = onb_from_w<PAR>(h.n);
    V3 dir;
    if (g.next() < 0.5f) {
      dir = light_random<PAR>(S, h.p, g);
    } else {
      V3 l = random_on_hemisphere<PAR>(g);
      dir = onb_local(uvw, l.x, l.y, l.z);
    }
    V3 udir = unit<PAR>(dir);
    float lpdf = light_pdf<PAR>(S, h.p, dir);
    float c1 = dot(udir, uvw.w);
    float cpdf = c1 > 0 ? c1 : 0.0f; // cosine_pdf::value returns cos, not cos/pi
    float pdf_value = PAR ? (float)(0.5 * (double)lpdf + 0.5 * (double)cpdf) : 0.5f * lpdf + 0.5f * cpdf;
    float c2 = dot(h.n, udir);
    float spdf;
    if (c2 < 0)
      spdf = 0;
    else
      spdf = PAR ? (float)((double)c2 / TPT_PI_D) : c2 * 0.3183098861837907f;
    V3 w = (atten * spdf) / pdf_value;
    ps.T = ps.T * w;
    ps.ray.o = h.p;
    ps.ray.d = dir;
  }
  ps.depth++;
  return !(dead_channel(ps.T.x) && dead_channel(ps.T.y) && dead_channel(ps.T.z));
}

// one bounce = extend + shade (megakernel form)
template <bool PAR, bool SMALL, bool MEDIA, int TEX = TPT_TEXF_ALL>
TPT_DEV bool bounce(const SceneView &S, PathState &ps, Rng &g, int max_depth, float t_min, V3 &radiance) {
  float t;
  int prim;
  uint32_t ndraw0;
  int cls = extend<PAR, SMALL, MEDIA, TEX>(S, ps, max_depth, t_min, t, prim, radiance, g, ndraw0);
  if (cls == TPT_EXT_DONE) return false;
  return shade<PAR, TEX>(S, ps, g, prim, t, ndraw0);
}

} // namespace tptd
