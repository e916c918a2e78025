// csrc/tpt_api.cu -- the extern "C" layer of include/tpt.h: scene upload, render dispatch,
// resolve/quantise kernel, statistics. Compiled with -fmad=false (the resolve kernel restates
// main.cpp:135-139 exactly). No CPU fallback anywhere: without a device every compute entry
// point returns TPT_ERR_NO_DEVICE.
#include "tpt.h"
#include "tpt_launch.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <functional>
#include <string>
#include <thread>
#include <mutex>
#include <vector>

using namespace tptd;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string &msg) {
  g_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char *what) {
  int code = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? TPT_ERR_NO_DEVICE : TPT_ERR_CUDA;
  return fail(code, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);                                            \
  } while (0)

#define TPT_MAX_BATCHES 256 // work-stealing batches of one multi-GPU render
#ifndef TPT_FBVH_LEAF
#define TPT_FBVH_LEAF 2 // primitives per leaf of the SAH BVH (<= 8: the leaf code keeps count - 1 in three bits)
#endif

struct ResolveArgs {
  const float *acc; // [n_ranges][npix][3]
  int n_ranges, npix, ns, slices, per_slice;
  int slice_last_range[TPT_MAX_RANGES];
  int nx, tiles_x, part_count;
  unsigned owned[TPT_MAX_BATCHES / 32]; // bit b: this device rendered the tiles t with t % part_count == b
  float *sum_rgb;       // [slices][npix][3] or null
  uint8_t *rgb8;        // [npix][3] or null
  uint8_t *rgb8_slices; // [slices][npix][3] or null
};

// int(255.99f * c) with x86 cvttss2si semantics for out-of-range / NaN (-> INT_MIN), then the
// writer's clamp to [0,255] (main.cpp:137-139, 176-182)
__device__ __forceinline__ uint8_t quantise(float sum, float denom) {
  float c = sum / denom;  // col /= float(ns)           main.cpp:135
  c = sqrtf(c);           // sqrt gamma                 main.cpp:136
  float v = 255.99f * c;  //                            main.cpp:137
  int q = (v >= -2147483648.f && v < 2147483648.f) ? (int)v : INT_MIN;
  q = q < 0 ? 0 : (q > 255 ? 255 : q);
  return (uint8_t)q;
}

__global__ void resolve_kernel(const __grid_constant__ ResolveArgs R) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= R.npix) return;
  // pixels of tiles another part renders: the accumulators were never written (and are never
  // read); their outputs are zero so that parts can be gathered by addition
  const int py = pix / R.nx, px = pix - py * R.nx;
  const int batch = ((py / TPT_TILE) * R.tiles_x + (px / TPT_TILE)) % R.part_count;
  const bool mine = (R.owned[batch >> 5] >> (batch & 31)) & 1u;
  float rx = 0.f, ry = 0.f, rz = 0.f;
  int slice = 0;
  for (int r = 0; r < (mine ? R.n_ranges : 0); r++) {
    const float *a = R.acc + ((size_t)r * R.npix + pix) * 3;
    rx += a[0];
    ry += a[1];
    rz += a[2];
    if (slice < R.slices && r == R.slice_last_range[slice]) {
      size_t o = ((size_t)slice * R.npix + pix) * 3;
      if (R.sum_rgb) {
        R.sum_rgb[o] = rx;
        R.sum_rgb[o + 1] = ry;
        R.sum_rgb[o + 2] = rz;
      }
      if (R.rgb8_slices) { // main.cpp:205-209
        float den = float(R.per_slice * (slice + 1));
        R.rgb8_slices[o] = quantise(rx, den);
        R.rgb8_slices[o + 1] = quantise(ry, den);
        R.rgb8_slices[o + 2] = quantise(rz, den);
      }
      slice++;
    }
  }
  if (!mine) {
    for (int sl = 0; sl < R.slices; sl++) {
      size_t o = ((size_t)sl * R.npix + pix) * 3;
      if (R.sum_rgb) R.sum_rgb[o] = R.sum_rgb[o + 1] = R.sum_rgb[o + 2] = 0.f;
      if (R.rgb8_slices) R.rgb8_slices[o] = R.rgb8_slices[o + 1] = R.rgb8_slices[o + 2] = 0;
    }
  }
  if (R.rgb8) {
    float den = float(R.ns);
    R.rgb8[(size_t)pix * 3] = quantise(rx, den);
    R.rgb8[(size_t)pix * 3 + 1] = quantise(ry, den);
    R.rgb8[(size_t)pix * 3 + 2] = quantise(rz, den);
  }
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct tpt_scene {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  // Pool of the image products (sums, 8-bit pictures): the one tpt_render_multi opens to GPU 0 for
  // the NVLink gather. The big per-range accumulators stay in the device's default pool, which is
  // never peer-mapped (a peer-mapped pool refused allocations above ~2 GB: "out of memory").
  cudaMemPool_t products = nullptr;
  bool products_private = false; // some product had to come from the default pool: gather through staged copies
  SceneLayout layout{};
  float4 *d_blob = nullptr;
  size_t blob_bytes = 0;
  bool use_smem = false;
  SmallScene small{}; // flat (chain, kind)-ordered geometry for the uniform brute-force closest hit
  FlatTree flat{};    // PARITY mode: the reference's own tree for the warp-uniform replay (closest_hit_flat)
  std::vector<cudaArray_t> arrays;
  std::vector<cudaTextureObject_t> textures;
  bool has_lights = false;
  int tex_mask = TPT_TEXF_ALL; // TPT_TEXF_* features present among the scene's textures (0: all constant -> lean kernel builds)
  int n_mediums = 0;
  bool fbvh_has_moving = false; // the fast BVH boxes moving spheres over [fbvh_t0, fbvh_t1] only
  float fbvh_t0 = 0, fbvh_t1 = 0;
  // world bounds for the pixel-bundle test: the root node's box when it is stated in world space
  bool root_box_ok = false;
  float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
  bool any_moving = false; // node boxes of moving spheres hold for [moving_t0, moving_t1] only
  float moving_t0 = 0, moving_t1 = 0;
  int background = 0;
  // render products (device)
  float *d_acc = nullptr;
  size_t acc_bytes = 0;
  unsigned long long *d_counters = nullptr;
  float *d_sum = nullptr;
  uint8_t *d_rgb8 = nullptr, *d_rgb8_slices = nullptr;
  size_t sum_bytes = 0, rgb8_bytes = 0, rgb8_slices_bytes = 0;
  int last_nx = 0, last_ny = 0, last_slices = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  tpt_stats stats{};
};

namespace {

template <typename T> void append(std::vector<unsigned char> &blob, const T *src, size_t count, size_t pad_to = 16) {
  size_t bytes = sizeof(T) * count;
  size_t at = blob.size();
  blob.resize(at + ((bytes + pad_to - 1) / pad_to) * pad_to, 0);
  if (bytes) std::memcpy(blob.data() + at, src, bytes);
}

int validate_desc(const tpt_scene_desc *d, int &depth_out) {
  if (!d) return fail(TPT_ERR_INVALID, "null scene description");
  if (d->api_version != TPT_API_VERSION) return fail(TPT_ERR_INVALID, "api_version mismatch");
  if (d->n_nodes <= 0 || !d->nodes) return fail(TPT_ERR_INVALID, "scene has no nodes");
  if (d->n_prims <= 0 || !d->prims) return fail(TPT_ERR_INVALID, "scene has no primitives");
  if (d->n_chains <= 0 || !d->chains) return fail(TPT_ERR_INVALID, "chain 0 (identity) is required");
  if (d->n_materials <= 0 || !d->materials) return fail(TPT_ERR_INVALID, "scene has no materials");
  if (d->n_images > TPT_MAX_IMAGES) return fail(TPT_ERR_UNSUPPORTED, "too many image textures");
  if (d->n_textures < 0 || (d->n_textures > 0 && !d->textures)) return fail(TPT_ERR_INVALID, "bad texture table");
  if (d->n_xform_ops < 0 || (d->n_xform_ops > 0 && !d->xform_ops)) return fail(TPT_ERR_INVALID, "bad transform-op table");
  if (d->n_images < 0 || (d->n_images > 0 && !d->images)) return fail(TPT_ERR_INVALID, "bad image table");
  if (d->n_lights < 0 || (d->n_lights > 0 && !d->lights)) return fail(TPT_ERR_INVALID, "bad light list");
  // structural walk: every group's end must nest properly (root tree, then medium boundaries)
  const int n_root = d->n_root_nodes > 0 ? d->n_root_nodes : d->n_nodes;
  if (n_root > d->n_nodes) return fail(TPT_ERR_INVALID, "n_root_nodes exceeds n_nodes");
  std::vector<int> ends;
  int depth = 0;
  for (int i = 0; i < d->n_nodes; i++) {
    while (!ends.empty() && ends.back() == i) ends.pop_back();
    if (i == n_root && !ends.empty()) return fail(TPT_ERR_INVALID, "root tree does not close at n_root_nodes");
    const tpt_node &n = d->nodes[i];
    int k = n.kind & 0xff, chain = n.kind >> 16;
    if (chain < 0 || chain >= d->n_chains) return fail(TPT_ERR_INVALID, "node chain out of range");
    if (k == TPT_NODE_LEAF) {
      if (n.end_or_prim < 0 || n.end_or_prim >= d->n_prims) return fail(TPT_ERR_INVALID, "leaf prim out of range");
    } else if (k == TPT_NODE_BVH || k == TPT_NODE_LIST) {
      int limit = ends.empty() ? d->n_nodes : ends.back();
      if (n.end_or_prim <= i || n.end_or_prim > limit) return fail(TPT_ERR_INVALID, "group end does not nest");
      ends.push_back(n.end_or_prim);
      if ((int)ends.size() > depth) depth = (int)ends.size();
    } else {
      return fail(TPT_ERR_INVALID, "unknown node kind");
    }
  }
  if (depth + 2 > TPT_MAX_FRAMES) return fail(TPT_ERR_UNSUPPORTED, "hitable tree nests deeper than TPT_MAX_FRAMES");
  depth_out = depth;
  for (int i = 0; i < d->n_prims; i++) {
    const tpt_prim &p = d->prims[i];
    if (p.kind < TPT_PRIM_SPHERE || p.kind > TPT_PRIM_MEDIUM) return fail(TPT_ERR_UNSUPPORTED, "unknown primitive kind");
    if (p.kind == TPT_PRIM_MEDIUM) {
      int32_t bf, be;
      std::memcpy(&bf, &p.p[1], 4);
      std::memcpy(&be, &p.p[2], 4);
      if (bf < n_root || be <= bf || be > d->n_nodes) return fail(TPT_ERR_INVALID, "medium boundary range out of bounds");
      if (!(p.p[0] > 0)) return fail(TPT_ERR_INVALID, "medium density must be positive");
      {
        // the parity walk of a boundary sub-tree keeps TPT_MAX_BOUNDARY_FRAMES frames (one is the virtual
        // root): a boundary nested deeper than that (a bvh_node or lists of lists) is refused here
        std::vector<int> open_ends;
        int deepest = 0;
        for (int k = bf; k < be; k++) {
          while (!open_ends.empty() && open_ends.back() == k) open_ends.pop_back();
          if ((d->nodes[k].kind & 0xff) != TPT_NODE_LEAF) {
            open_ends.push_back(d->nodes[k].end_or_prim);
            deepest = std::max(deepest, (int)open_ends.size());
          }
        }
        if (deepest + 1 > TPT_MAX_BOUNDARY_FRAMES)
          return fail(TPT_ERR_UNSUPPORTED, "medium boundary nests deeper than TPT_MAX_BOUNDARY_FRAMES");
      }
      for (int k = bf; k < be; k++)
        if ((d->nodes[k].kind & 0xff) == TPT_NODE_LEAF && d->prims[d->nodes[k].end_or_prim].kind == TPT_PRIM_MEDIUM)
          return fail(TPT_ERR_UNSUPPORTED, "medium inside a medium boundary");
    }
    if (p.material < 0 || p.material >= d->n_materials) return fail(TPT_ERR_INVALID, "prim material out of range");
    if (p.chain < 0 || p.chain >= d->n_chains) return fail(TPT_ERR_INVALID, "prim chain out of range");
  }
  for (int i = 0; i < d->n_chains; i++) {
    const tpt_chain &c = d->chains[i];
    if (c.n_ops < 0 || c.first_op < 0 || (long long)c.first_op + c.n_ops > d->n_xform_ops)
      return fail(TPT_ERR_INVALID, "chain ops out of range");
  }
  bool needs_perlin = false;
  for (int i = 0; i < d->n_textures; i++) {
    const tpt_texture &t = d->textures[i];
    if (t.kind == TPT_TEX_CHECKER && (t.odd < 0 || t.odd >= d->n_textures || t.even < 0 || t.even >= d->n_textures))
      return fail(TPT_ERR_INVALID, "checker children out of range");
    if (t.kind == TPT_TEX_IMAGE && (t.image < 0 || t.image >= d->n_images))
      return fail(TPT_ERR_INVALID, "image index out of range");
    if (t.kind == TPT_TEX_PERLIN) needs_perlin = true;
    if (t.kind < TPT_TEX_CONSTANT || t.kind > TPT_TEX_IMAGE) return fail(TPT_ERR_UNSUPPORTED, "unknown texture kind");
  }
  if (needs_perlin && !d->perlin) return fail(TPT_ERR_INVALID, "perlin texture without tables");
  for (int i = 0; i < d->n_materials; i++) {
    const tpt_material &m = d->materials[i];
    if (m.kind < TPT_MAT_LAMBERTIAN || m.kind > TPT_MAT_ISOTROPIC) return fail(TPT_ERR_UNSUPPORTED, "unknown material kind");
    if ((m.kind == TPT_MAT_LAMBERTIAN || m.kind == TPT_MAT_DIFFUSE_LIGHT) && (m.texture < 0 || m.texture >= d->n_textures))
      return fail(TPT_ERR_INVALID, "material texture out of range");
  }
  for (int i = 0; i < d->n_images; i++)
    if (!d->images[i].rgb || d->images[i].width <= 0 || d->images[i].height <= 0)
      return fail(TPT_ERR_INVALID, "image without pixels");
  return TPT_OK;
}

// ---------------------------------------------------------------------------------------------
// FAST-mode acceleration structure for scenes above the brute-force size: binned-SAH BVH2 over
// the world-space boxes of the non-duplicate leaves. Host side, once per scene upload.
// ---------------------------------------------------------------------------------------------
struct Box3 {
  float lo[3], hi[3];
  void reset() {
    for (int c = 0; c < 3; c++) { lo[c] = FLT_MAX; hi[c] = -FLT_MAX; }
  }
  void grow(const Box3 &b) {
    for (int c = 0; c < 3; c++) { lo[c] = std::min(lo[c], b.lo[c]); hi[c] = std::max(hi[c], b.hi[c]); }
  }
  float area() const {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.f * (dx * dy + dy * dz + dz * dx);
  }
};

struct FastBvh {
  std::vector<float> nodes; // 16 floats per node
  std::vector<int32_t> leaf_prims;
  float moving_t0 = -FLT_MAX, moving_t1 = FLT_MAX; // ray times ALL the moving-sphere boxes are valid for (intersection of the spheres' own intervals)
  bool has_moving = false;
};

// object-space box of a leaf -> world space through the inverse of its wrapper chain
// (translate: +offset, rotate_y: x = c x' + s z', z = -s x' + c z'), innermost wrapper first
// a box in the object space of transform chain `chain` -> world space (the box of its eight corners)
Box3 chain_to_world(const tpt_scene_desc *d, const Box3 &in, int chain) {
  const tpt_chain &ch = d->chains[chain];
  Box3 out;
  out.reset();
  for (int corner = 0; corner < 8; corner++) {
    double p[3] = {corner & 1 ? in.hi[0] : in.lo[0], corner & 2 ? in.hi[1] : in.lo[1], corner & 4 ? in.hi[2] : in.lo[2]};
    for (int k = ch.n_ops - 1; k >= 0; k--) {
      const tpt_xform_op &op = d->xform_ops[ch.first_op + k];
      if (op.kind == TPT_XF_TRANSLATE) {
        p[0] += op.a; p[1] += op.b; p[2] += op.c;
      } else {
        double x = op.b * p[0] + op.a * p[2], z = -op.a * p[0] + op.b * p[2];
        p[0] = x; p[2] = z;
      }
    }
    Box3 b;
    for (int c = 0; c < 3; c++) b.lo[c] = b.hi[c] = (float)p[c];
    out.grow(b);
  }
  return out;
}

Box3 world_box(const tpt_scene_desc *d, const tpt_node &leaf) {
  Box3 in;
  for (int c = 0; c < 3; c++) { in.lo[c] = leaf.bmin[c]; in.hi[c] = leaf.bmax[c]; }
  Box3 out = chain_to_world(d, in, leaf.kind >> 16);
  for (int c = 0; c < 3; c++) { // conservative padding: fp32 slab test vs the primitive's own arithmetic
    float pad = 1e-4f + 1e-5f * std::max(std::fabs(out.lo[c]), std::fabs(out.hi[c]));
    out.lo[c] -= pad;
    out.hi[c] += pad;
  }
  return out;
}

struct BuildItem { Box3 box; float cen[3]; int prim; };

int32_t build_fbvh_rec(std::vector<BuildItem> &items, int begin, int end, FastBvh &out, Box3 &bounds_out);

// object-space bounds of a surface primitive from its OWN parameters (sphere: centre +- radius; moving sphere: both end
// positions, valid for ray times inside [time0, time1]; rects: the reference's 0.0002-thick slab, src/rect_box.cc)
static bool prim_own_box(const tpt_prim &p, Box3 &b) {
  b.reset();
  auto add = [&b](float x, float y, float z) {
    Box3 q;
    q.lo[0] = q.hi[0] = x; q.lo[1] = q.hi[1] = y; q.lo[2] = q.hi[2] = z;
    b.grow(q);
  };
  const float *q = p.p;
  if (p.kind == TPT_PRIM_SPHERE || p.kind == TPT_PRIM_MOVING_SPHERE) {
    const float r = std::fabs(q[3]);
    add(q[0] - r, q[1] - r, q[2] - r);
    add(q[0] + r, q[1] + r, q[2] + r);
    if (p.kind == TPT_PRIM_MOVING_SPHERE) {
      add(q[4] - r, q[5] - r, q[6] - r);
      add(q[4] + r, q[5] + r, q[6] + r);
    }
  } else if (p.kind == TPT_PRIM_XY_RECT) {
    add(q[0], q[2], q[4] - 1e-4f); add(q[1], q[3], q[4] + 1e-4f);
  } else if (p.kind == TPT_PRIM_XZ_RECT) {
    add(q[0], q[4] - 1e-4f, q[2]); add(q[1], q[4] + 1e-4f, q[3]);
  } else if (p.kind == TPT_PRIM_YZ_RECT) {
    add(q[4] - 1e-4f, q[0], q[2]); add(q[4] + 1e-4f, q[1], q[3]);
  } else {
    return false;
  }
  for (int c = 0; c < 3; c++)
    if (!std::isfinite(b.lo[c]) || !std::isfinite(b.hi[c])) return false;
  return true;
}

// closest_hit_skip's cull assumes what the reference's own bounding_box() guarantees and a hand-made description may
// not: every primitive lies inside the box of every bvh_node above it. True iff that holds for the root tree (all
// bvh_nodes in world space; primitive bounds carried out of their transform chains), with a slack of 1e-5 of the extent.
static bool boxes_contain_their_prims(const tpt_scene_desc *d, int n_root) {
  std::vector<int> open; // bvh_nodes whose sub-tree the scan is in
  for (int i = 0; i < n_root; i++) {
    while (!open.empty() && i >= d->nodes[open.back()].end_or_prim) open.pop_back();
    const tpt_node &nd = d->nodes[i];
    const int k = nd.kind & 0xff;
    if (k == TPT_NODE_BVH) {
      if ((nd.kind >> 16) != 0) return false;
      open.push_back(i);
    } else if (k == TPT_NODE_LEAF && !(nd.kind & TPT_NODE_DUP)) {
      const tpt_prim &p = d->prims[nd.end_or_prim];
      Box3 own;
      if (!prim_own_box(p, own)) return false;
      const Box3 w = chain_to_world(d, own, nd.kind >> 16);
      for (int a : open) {
        const tpt_node &an = d->nodes[a];
        for (int c = 0; c < 3; c++) {
          const float slack = 1e-5f * std::max(1.0f, std::max(std::fabs(an.bmin[c]), std::fabs(an.bmax[c])));
          if (!(w.lo[c] >= an.bmin[c] - slack) || !(w.hi[c] <= an.bmax[c] + slack)) return false;
        }
      }
    }
  }
  return true;
}

// returns the child reference: >= 0 inner node index, < 0 ~((first << 3) | (count - 1))
int32_t make_child(std::vector<BuildItem> &items, int begin, int end, FastBvh &out, Box3 &box) {
  int n = end - begin;
  if (n <= TPT_FBVH_LEAF) {
    box.reset();
    int first = (int)out.leaf_prims.size();
    for (int i = begin; i < end; i++) {
      out.leaf_prims.push_back(items[i].prim);
      box.grow(items[i].box);
    }
    return ~((first << 3) | (n - 1));
  }
  return build_fbvh_rec(items, begin, end, out, box);
}

// bin of a centroid along an axis of extent ext starting at lo; any non-finite quotient lands in bin 0
static int bin_of(int nb, float cen, float lo, float ext) {
  const float q = (float)nb * (cen - lo) / ext;
  if (!(q > 0.f)) return 0;
  return q >= (float)(nb - 1) ? nb - 1 : (int)q;
}

int32_t build_fbvh_rec(std::vector<BuildItem> &items, int begin, int end, FastBvh &out, Box3 &bounds_out) {
  Box3 cb, bb;
  cb.reset();
  bb.reset();
  for (int i = begin; i < end; i++) {
    Box3 c;
    for (int k = 0; k < 3; k++) c.lo[k] = c.hi[k] = items[i].cen[k];
    cb.grow(c);
    bb.grow(items[i].box);
  }
  bounds_out = bb;
  // binned SAH over the axis / plane with the lowest cost
  const int NB = 16;
  int best_axis = -1, best_split = -1;
  float best_cost = FLT_MAX;
  for (int axis = 0; axis < 3; axis++) {
    float ext = cb.hi[axis] - cb.lo[axis];
    if (!(ext > 0)) continue;
    Box3 bins[NB];
    int cnt[NB] = {0};
    for (auto &b : bins) b.reset();
    for (int i = begin; i < end; i++) {
      int k = bin_of(NB, items[i].cen[axis], cb.lo[axis], ext);
      bins[k].grow(items[i].box);
      cnt[k]++;
    }
    float right_area[NB];
    int right_cnt[NB];
    Box3 acc;
    acc.reset();
    int c = 0;
    for (int k = NB - 1; k > 0; k--) {
      acc.grow(bins[k]);
      c += cnt[k];
      right_area[k] = c ? acc.area() : 0.f;
      right_cnt[k] = c;
    }
    acc.reset();
    c = 0;
    for (int k = 0; k < NB - 1; k++) {
      acc.grow(bins[k]);
      c += cnt[k];
      if (c == 0 || right_cnt[k + 1] == 0) continue;
      float cost = acc.area() * c + right_area[k + 1] * right_cnt[k + 1];
      if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = k; }
    }
  }
  int mid;
  if (best_axis < 0) {
    mid = (begin + end) / 2; // coincident centroids
  } else {
    float ext = cb.hi[best_axis] - cb.lo[best_axis];
    auto it = std::partition(items.begin() + begin, items.begin() + end, [&](const BuildItem &b) {
      return bin_of(NB, b.cen[best_axis], cb.lo[best_axis], ext) <= best_split;
    });
    mid = (int)(it - items.begin());
    if (mid == begin || mid == end) mid = (begin + end) / 2;
  }
  int me = (int)(out.nodes.size() / 16);
  out.nodes.resize(out.nodes.size() + 16, 0.f);
  Box3 b0, b1;
  int32_t c0 = make_child(items, begin, mid, out, b0);
  int32_t c1 = make_child(items, mid, end, out, b1);
  float *n = &out.nodes[(size_t)me * 16];
  n[0] = b0.lo[0]; n[1] = b0.lo[1]; n[2] = b0.lo[2]; n[3] = b0.hi[0];
  n[4] = b0.hi[1]; n[5] = b0.hi[2]; n[6] = b1.lo[0]; n[7] = b1.lo[1];
  n[8] = b1.lo[2]; n[9] = b1.hi[0]; n[10] = b1.hi[1]; n[11] = b1.hi[2];
  std::memcpy(&n[12], &c0, 4);
  std::memcpy(&n[13], &c1, 4);
  return me;
}

// BVH2 -> BVH4: every node keeps absorbing its largest inner child's two children until it has four
// (surface-area heuristic for which child to open). Half the levels of the binary tree: half the
// dependent node fetches per ray and, in a warp whose lanes walk different sub-trees, half the loop
// trips over which their counts can differ. Node = 8 float4, structure-of-arrays over the four
// children: lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] child[4] unused[4]; a child reference is
// >= 0 (inner node), < 0 (leaf code as in the binary tree) or TPT_FBVH_EMPTY.
void collapse_to_bvh4(const FastBvh &b2, std::vector<float> &out, int &max_depth) {
  struct Kid { Box3 box; int32_t ref; };
  auto kids_of = [&b2](int32_t node, Kid k[2]) {
    const float *n = &b2.nodes[(size_t)node * 16];
    for (int c = 0; c < 2; c++) {
      for (int a = 0; a < 3; a++) { k[c].box.lo[a] = n[6 * c + a]; k[c].box.hi[a] = n[6 * c + 3 + a]; }
      std::memcpy(&k[c].ref, &n[12 + c], 4);
    }
  };
  max_depth = 0;
  std::function<int32_t(int32_t, int)> emit = [&](int32_t node2, int depth) -> int32_t {
    max_depth = std::max(max_depth, depth);
    std::vector<Kid> kids(2);
    kids_of(node2, kids.data());
    while (kids.size() < 4) {
      int open = -1;
      float best = -1.f;
      for (size_t i = 0; i < kids.size(); i++)
        if (kids[i].ref >= 0 && kids[i].box.area() > best) { best = kids[i].box.area(); open = (int)i; }
      if (open < 0) break;
      Kid two[2];
      kids_of(kids[open].ref, two);
      kids[open] = two[0];
      kids.push_back(two[1]);
    }
    const int me = (int)(out.size() / 32);
    out.resize(out.size() + 32, 0.f);
    std::vector<int32_t> refs(4, TPT_FBVH_EMPTY);
    for (size_t i = 0; i < kids.size(); i++) refs[i] = kids[i].ref >= 0 ? emit(kids[i].ref, depth + 1) : kids[i].ref;
    float *n = &out[(size_t)me * 32];
    for (int i = 0; i < 4; i++) {
      const bool used = i < (int)kids.size();
      for (int a = 0; a < 3; a++) {
        n[4 * a + i] = used ? kids[i].box.lo[a] : FLT_MAX;       // an empty slot: a point at +FLT_MAX, never hit
        n[12 + 4 * a + i] = used ? kids[i].box.hi[a] : FLT_MAX;
      }
      std::memcpy(&n[24 + i], &refs[i], 4);
    }
    return me;
  };
  if (!b2.nodes.empty()) emit(0, 1);
}

// surface primitives reachable from the root tree (media and their boundaries excluded)
std::vector<int> root_surface_prims(const tpt_scene_desc *d, int n_root) {
  std::vector<int> ids;
  std::vector<char> seen(d->n_prims, 0);
  int skip_until = -1;
  for (int i = 0; i < n_root; i++) {
    const tpt_node &nd = d->nodes[i];
    int k = nd.kind & 0xff;
    if (i < skip_until) continue;
    if (nd.kind & TPT_NODE_DUP) {
      if (k != TPT_NODE_LEAF) skip_until = nd.end_or_prim;
      continue;
    }
    if (k != TPT_NODE_LEAF || seen[nd.end_or_prim] || d->prims[nd.end_or_prim].kind == TPT_PRIM_MEDIUM) continue;
    seen[nd.end_or_prim] = 1;
    ids.push_back(nd.end_or_prim);
  }
  return ids;
}
int count_root_surface_prims(const tpt_scene_desc *d, int n_root) { return (int)root_surface_prims(d, n_root).size(); }

void build_fast_bvh(const tpt_scene_desc *d, int n_root, FastBvh &out) {
  std::vector<BuildItem> items;
  std::vector<char> seen(d->n_prims, 0);
  int skip_until = -1;
  for (int i = 0; i < n_root; i++) {
    const tpt_node &nd = d->nodes[i];
    int k = nd.kind & 0xff;
    if (i < skip_until) continue;
    if (nd.kind & TPT_NODE_DUP) {
      if (k != TPT_NODE_LEAF) skip_until = nd.end_or_prim;
      continue;
    }
    if (k != TPT_NODE_LEAF || seen[nd.end_or_prim] || d->prims[nd.end_or_prim].kind == TPT_PRIM_MEDIUM) continue;
    seen[nd.end_or_prim] = 1;
    BuildItem it;
    it.box = world_box(d, nd);
    for (int c = 0; c < 3; c++) it.cen[c] = 0.5f * (it.box.lo[c] + it.box.hi[c]);
    // a leaf box that is not a finite, ordered interval (a caller's NaN / infinite coordinates or transform)
    // cannot be binned: no SAH tree for this scene, FAST mode walks the caller's own tree instead
    for (int c = 0; c < 3; c++)
      if (!std::isfinite(it.box.lo[c]) || !std::isfinite(it.box.hi[c]) || !(it.box.lo[c] <= it.box.hi[c]) || !std::isfinite(it.cen[c])) {
        out = FastBvh();
        return;
      }
    it.prim = nd.end_or_prim;
    const tpt_prim &p = d->prims[it.prim];
    if (p.kind == TPT_PRIM_MOVING_SPHERE) {
      out.has_moving = true;
      out.moving_t0 = std::max(out.moving_t0, std::min(p.p[7], p.p[8]));
      out.moving_t1 = std::min(out.moving_t1, std::max(p.p[7], p.p[8]));
    }
    items.push_back(it);
  }
  if (items.size() < 2) return;
  Box3 root;
  build_fbvh_rec(items, 0, (int)items.size(), out, root);
}

// Group the non-duplicate leaves by transform chain and primitive kind (xy, xz, yz rect, sphere).
// Closest-hit does not depend on the order except on exact ties, which FAST mode does not
// promise to resolve like the reference.
void build_small_scene(const tpt_scene_desc *d, int n_root, bool smem_ok, SmallScene &Q) {
  std::memset(&Q, 0, sizeof(Q));
  const std::vector<int> surface = root_surface_prims(d, n_root);
  std::vector<char> is_surface(d->n_prims, 0);
  for (int id : surface) is_surface[id] = 1;
  if (!smem_ok || (int)surface.size() > TPT_SMALL_MAX_PRIMS || d->n_chains > TPT_SMALL_MAX_GROUPS ||
      d->n_xform_ops > TPT_SMALL_MAX_OPS)
    return;
  for (int id : surface)
    if (d->prims[id].kind == TPT_PRIM_MOVING_SPHERE) return;
  for (int i = 0; i < d->n_xform_ops; i++) {
    const tpt_xform_op &op = d->xform_ops[i];
    int32_t kind = op.kind;
    float kf;
    std::memcpy(&kf, &kind, 4);
    Q.ops[i] = make_float4(kf, op.a, op.b, op.c);
  }
  // ---- recognise `box` lists: six non-duplicate leaf children of one LIST, same chain, that are
  // exactly the faces of an axis-aligned block (src/rect_box.cc:93-115) ----
  std::vector<int> in_box(d->n_prims, -1);
  int nb = 0;
  for (int i = 0; i < n_root && nb < TPT_SMALL_MAX_BOXES; i++) {
    const tpt_node &g = d->nodes[i];
    if ((g.kind & 0xff) != TPT_NODE_LIST || (g.kind & TPT_NODE_DUP) || g.end_or_prim - i - 1 != 6) continue;
    int ids[6];
    bool ok = true;
    for (int c = 0; c < 6 && ok; c++) {
      const tpt_node &l = d->nodes[i + 1 + c];
      ok = (l.kind & 0xff) == TPT_NODE_LEAF && (l.kind >> 16) == (g.kind >> 16);
      ids[c] = l.end_or_prim;
      if (ok && in_box[ids[c]] >= 0) ok = false;
    }
    if (!ok) continue;
    // candidate extents from the first xy_rect found
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    int face[6] = {-1, -1, -1, -1, -1, -1};
    float zs[2], ys[2], xs[2];
    int nz = 0, ny = 0, nx = 0;
    for (int c = 0; c < 6; c++) {
      const tpt_prim &p = d->prims[ids[c]];
      if (p.kind == TPT_PRIM_XY_RECT && nz < 2) zs[nz++] = p.p[4];
      else if (p.kind == TPT_PRIM_XZ_RECT && ny < 2) ys[ny++] = p.p[4];
      else if (p.kind == TPT_PRIM_YZ_RECT && nx < 2) xs[nx++] = p.p[4];
      else ok = false;
    }
    if (!ok || nx != 2 || ny != 2 || nz != 2) continue;
    lo[0] = std::min(xs[0], xs[1]); hi[0] = std::max(xs[0], xs[1]);
    lo[1] = std::min(ys[0], ys[1]); hi[1] = std::max(ys[0], ys[1]);
    lo[2] = std::min(zs[0], zs[1]); hi[2] = std::max(zs[0], zs[1]);
    if (!(lo[0] < hi[0] && lo[1] < hi[1] && lo[2] < hi[2])) continue;
    for (int c = 0; c < 6 && ok; c++) {
      const tpt_prim &p = d->prims[ids[c]];
      int axis, a, b; // plane axis and the two in-plane axes in the order of p.p[0..3]
      if (p.kind == TPT_PRIM_XY_RECT) { axis = 2; a = 0; b = 1; }
      else if (p.kind == TPT_PRIM_XZ_RECT) { axis = 1; a = 0; b = 2; }
      else { axis = 0; a = 1; b = 2; }
      ok = p.p[0] == lo[a] && p.p[1] == hi[a] && p.p[2] == lo[b] && p.p[3] == hi[b];
      int side = p.p[4] == lo[axis] ? 0 : 1;
      if (ok && face[2 * axis + side] >= 0) ok = false;
      face[2 * axis + side] = ids[c];
    }
    if (!ok) continue;
    SmallBox &B = Q.boxes[nb];
    B.lo = make_float4(lo[0], lo[1], lo[2], 0.f);
    B.hi = make_float4(hi[0], hi[1], hi[2], 0.f);
    for (int f = 0; f < 6; f++) {
      B.face[f] = face[f];
      in_box[face[f]] = nb;
    }
    B.face[6] = g.kind >> 16; // chain, consumed below
    nb++;
  }
  // ---- loose rects of one chain that are whole faces of a common axis-aligned block (the five
  // walls of a Cornell room): two perpendicular faces fix the block, every rect of the chain that
  // is exactly one of its faces joins; >= 4 faces make it worth one slab test instead of that many
  // rectangle tests. Absent faces stay -1 (closest_hit_uniform skips a crossing without a face).
  // TPT_SMALL_OPEN_BLOCKS=0 in the environment keeps the rects separate (A/B measurements). ----
  const char *open_env = std::getenv("TPT_SMALL_OPEN_BLOCKS");
  const bool open_blocks = !(open_env && open_env[0] == '0');
  auto rect_axes = [](int kind, int &axis, int &a, int &b) {
    if (kind == TPT_PRIM_XY_RECT) { axis = 2; a = 0; b = 1; }
    else if (kind == TPT_PRIM_XZ_RECT) { axis = 1; a = 0; b = 2; }
    else { axis = 0; a = 1; b = 2; }
  };
  auto extent = [&](const tpt_prim &p, int along, float &lo, float &hi) { // in-plane extent along an axis
    int axis, a, b;
    rect_axes(p.kind, axis, a, b);
    if (along == a) { lo = p.p[0]; hi = p.p[1]; }
    else { lo = p.p[2]; hi = p.p[3]; }
  };
  auto match_faces = [&](const std::vector<int> &cand, const float lo[3], const float hi[3], int face[6]) {
    int count = 0;
    for (int f = 0; f < 6; f++) face[f] = -1;
    for (int id : cand) {
      const tpt_prim &p = d->prims[id];
      int axis, a, b;
      rect_axes(p.kind, axis, a, b);
      if (!(p.p[0] == lo[a] && p.p[1] == hi[a] && p.p[2] == lo[b] && p.p[3] == hi[b])) continue;
      if (p.p[4] != lo[axis] && p.p[4] != hi[axis]) continue;
      const int slot = 2 * axis + (p.p[4] == lo[axis] ? 0 : 1);
      if (face[slot] >= 0) continue; // a second rect on the same face stays a rect
      face[slot] = id;
      count++;
    }
    return count;
  };
  for (int c = 0; open_blocks && c < d->n_chains; c++) {
    while (nb < TPT_SMALL_MAX_BOXES) {
      std::vector<int> cand;
      for (int id : surface) {
        const tpt_prim &p = d->prims[id];
        if (p.chain == c && in_box[id] < 0 && p.kind >= TPT_PRIM_XY_RECT && p.kind <= TPT_PRIM_YZ_RECT) cand.push_back(id);
      }
      int best_count = 0, best_face[6];
      float best_lo[3], best_hi[3];
      for (int i : cand)
        for (int j : cand) {
          const tpt_prim &pi = d->prims[i], &pj = d->prims[j];
          int ai, aa, ab, aj, ja, jb;
          rect_axes(pi.kind, ai, aa, ab);
          rect_axes(pj.kind, aj, ja, jb);
          if (ai == aj) continue;
          const int ac = 3 - ai - aj; // the axis both rects extend along
          float lo[3], hi[3], l2, h2;
          extent(pi, ac, lo[ac], hi[ac]);
          extent(pj, ac, l2, h2);
          if (l2 != lo[ac] || h2 != hi[ac]) continue;
          extent(pi, aj, lo[aj], hi[aj]);
          extent(pj, ai, lo[ai], hi[ai]);
          if ((pj.p[4] != lo[aj] && pj.p[4] != hi[aj]) || (pi.p[4] != lo[ai] && pi.p[4] != hi[ai])) continue;
          if (!(lo[0] < hi[0] && lo[1] < hi[1] && lo[2] < hi[2])) continue;
          int face[6];
          const int count = match_faces(cand, lo, hi, face);
          if (count > best_count) {
            best_count = count;
            for (int f = 0; f < 6; f++) best_face[f] = face[f];
            for (int k = 0; k < 3; k++) { best_lo[k] = lo[k]; best_hi[k] = hi[k]; }
          }
        }
      if (best_count < 4) break;
      SmallBox &B = Q.boxes[nb];
      B.lo = make_float4(best_lo[0], best_lo[1], best_lo[2], 0.f);
      B.hi = make_float4(best_hi[0], best_hi[1], best_hi[2], 0.f);
      for (int f = 0; f < 6; f++) {
        B.face[f] = best_face[f];
        if (best_face[f] >= 0) in_box[best_face[f]] = nb;
      }
      B.face[6] = c;
      B.face[7] = best_count < 6 ? 1 : 0;
      nb++;
    }
  }
  int n = 0, ng = 0, nbox_sorted = 0;
  SmallBox sorted[TPT_SMALL_MAX_BOXES];
  const int order[4] = {TPT_PRIM_XY_RECT, TPT_PRIM_XZ_RECT, TPT_PRIM_YZ_RECT, TPT_PRIM_SPHERE};
  for (int c = 0; c < d->n_chains; c++) {
    SmallGroup G{};
    G.first_op = d->chains[c].first_op;
    G.n_ops = d->chains[c].n_ops;
    G.begin = n;
    {
      // compose the wrappers (outermost first): translate: o -= t ; rotate_y: (x,z) -> (c x - s z, s x + c z)
      double cs = 1.0, sn = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
      for (int k = 0; k < G.n_ops; k++) {
        const tpt_xform_op &op = d->xform_ops[G.first_op + k];
        if (op.kind == TPT_XF_TRANSLATE) {
          bx -= op.a;
          by -= op.b;
          bz -= op.c;
        } else {
          double s2 = op.a, c2 = op.b;
          double ncs = c2 * cs - s2 * sn, nsn = s2 * cs + c2 * sn;
          double nbx = c2 * bx - s2 * bz, nbz = s2 * bx + c2 * bz;
          cs = ncs;
          sn = nsn;
          bx = nbx;
          bz = nbz;
        }
      }
      G.cs = (float)cs;
      G.sn = (float)sn;
      G.bx = (float)bx;
      G.by = (float)by;
      G.bz = (float)bz;
    }
    int ends[4];
    for (int o = 0; o < 4; o++) {
      for (int i = 0; i < d->n_prims; i++) {
        const tpt_prim &p = d->prims[i];
        if (!is_surface[i] || p.chain != c || p.kind != order[o] || in_box[i] >= 0) continue;
        int32_t id = i;
        float idf;
        std::memcpy(&idf, &id, 4);
        if (p.kind == TPT_PRIM_SPHERE) {
          Q.geo[n] = make_float4(p.p[0], p.p[1], p.p[2], p.p[3]);
          Q.aux[n] = make_float2(p.p[3] >= 500.0f ? 1.0f : 0.0f, idf); // huge "wall" spheres: exact roots
        } else {
          Q.geo[n] = make_float4(p.p[0], p.p[1], p.p[2], p.p[3]);
          Q.aux[n] = make_float2(p.p[4], idf);
        }
        n++;
      }
      ends[o] = n;
    }
    G.box_begin = nbox_sorted;
    for (int b = 0; b < nb; b++)
      if (Q.boxes[b].face[6] == c) sorted[nbox_sorted++] = Q.boxes[b];
    G.box_end = nbox_sorted;
    if (n == G.begin && G.box_begin == G.box_end) continue; // nothing lives directly under this chain
    G.xy_end = ends[0];
    G.xz_end = ends[1];
    G.yz_end = ends[2];
    G.sph_end = ends[3];
    Q.groups[ng++] = G;
  }
  for (int b = 0; b < nbox_sorted; b++) Q.boxes[b] = sorted[b];
  Q.n_groups = ng;
  Q.enabled = 1;
}

// PARITY mode: the reference's tree as a FlatTree (tpt_device.cuh) when it is small and has the shape
// closest_hit_flat replays: bvh_nodes on top, leaves and hitable_lists (nested lists and `box`
// objects concatenate) below them, no bvh_node under a list, no media, no moving spheres.
// TPT_PARITY_FLAT=0 in the environment keeps the generic frame-stack replay (A/B and test use).
void build_flat_tree(const tpt_scene_desc *d, int n_root, bool smem_ok, FlatTree &F) {
  std::memset(&F, 0, sizeof(F));
  if (const char *e = std::getenv("TPT_PARITY_FLAT"))
    if (e[0] == '0') return;
  if (!smem_ok) return;
  bool ok = true;
  int nb = 0, np = 0, ni = 0;
  auto put_prim = [&](int id, int chain) {
    const tpt_prim &p = d->prims[id];
    if (np >= TPT_FLAT_MAX_PRIMS || p.kind == TPT_PRIM_MEDIUM || p.kind == TPT_PRIM_MOVING_SPHERE) {
      ok = false;
      return;
    }
    FlatPrim &q = F.prims[np++];
    q.geo = make_float4(p.p[0], p.p[1], p.p[2], p.p[3]);
    q.kind = p.kind;
    q.prim = id;
    q.chain = chain;
    q.k = p.p[4];
  };
  // recursive descent over the pre-order array: [i, end) are the children of one group
  std::function<void(int, int, int, bool)> walk = [&](int i, int end, int parent, bool in_list) {
    while (ok && i < end) {
      const tpt_node &nd = d->nodes[i];
      const int k = nd.kind & 0xff, chain = nd.kind >> 16;
      const int next = k == TPT_NODE_LEAF ? i + 1 : nd.end_or_prim;
      if (nd.kind & TPT_NODE_DUP) { // second visit of a one-element bvh_node's child: same record twice
        i = next;
        continue;
      }
      if (k == TPT_NODE_BVH) {
        if (in_list || nb >= TPT_FLAT_MAX_BOXES) {
          ok = false;
          return;
        }
        const int b = nb++;
        int32_t pi = parent;
        float cf, pf;
        std::memcpy(&cf, &chain, 4);
        std::memcpy(&pf, &pi, 4);
        F.boxes[b].lo = make_float4(nd.bmin[0], nd.bmin[1], nd.bmin[2], cf);
        F.boxes[b].hi = make_float4(nd.bmax[0], nd.bmax[1], nd.bmax[2], pf);
        walk(i + 1, nd.end_or_prim, b, false);
      } else {
        const bool opens = !in_list; // a leaf or list directly below a bvh_node (or the root itself): one item
        if (opens) {
          if (ni >= TPT_FLAT_MAX_ITEMS) {
            ok = false;
            return;
          }
          F.items[ni].first = np;
          F.items[ni].parent = parent;
        }
        if (k == TPT_NODE_LEAF) put_prim(nd.end_or_prim, d->prims[nd.end_or_prim].chain);
        else walk(i + 1, nd.end_or_prim, parent, true);
        if (opens) F.items[ni++].end = np;
      }
      i = next;
    }
  };
  walk(0, n_root, -1, false);
  if (!ok || ni == 0) {
    std::memset(&F, 0, sizeof(F));
    return;
  }
  // visit order: grouped by the transform chain of the item's first primitive (stable), DFS rank kept for ties
  for (int i = 0; i < ni; i++) F.items[i].rank = i;
  std::stable_sort(F.items, F.items + ni, [&F](const FlatItem &a, const FlatItem &b) {
    return F.prims[a.first].chain < F.prims[b.first].chain;
  });
  F.n_boxes = nb;
  F.n_items = ni;
  F.n_prims = np;
  F.enabled = 1;
}

// Render products come from the device's stream-ordered memory pool (cudaMallocAsync): the pool
// keeps freed blocks (release threshold = max, set in tpt_scene_create), so creating a scene,
// rendering and destroying it again -- the end-to-end pattern of a short-lived caller -- does not
// pay cudaMalloc / cudaFree device synchronisations and page (un)mapping every time.
// `pool`: the device's products pool (see products_pool) or nullptr for the default pool
int ensure(cudaStream_t st, void **p, size_t &have, size_t want, cudaMemPool_t pool = nullptr, bool *fell_back = nullptr) {
  if (have >= want && *p) return TPT_OK;
  if (*p) cudaFreeAsync(*p, st);
  *p = nullptr;
  have = 0;
  cudaError_t e = pool ? cudaMallocFromPoolAsync(p, want, pool, st) : cudaMallocAsync(p, want, st);
  if (e != cudaSuccess && pool) { // e.g. a product too large to be peer-mapped: private memory, staged gather
    cudaGetLastError();
    e = cudaMallocAsync(p, want, st);
    if (fell_back) *fell_back = true;
  }
  if (e != cudaSuccess) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    char what[160];
    std::snprintf(what, sizeof(what), "cudaMallocAsync(%zu bytes; device reports %zu of %zu free)", want, free_b, total_b);
    return cuda_fail(e, what);
  }
  have = want;
  return TPT_OK;
}

struct Plan {
  RenderArgs args;
  ResolveArgs res;
  int blocks = 0, blocks_per_sm = 0;
  bool parity = false;
  bool wavefront = false;
  size_t npix = 0;
};

int make_plan(tpt_scene *s, const tpt_camera *cam, const tpt_render_params *p, Plan &plan) {
  if (!s || !cam || !p) return fail(TPT_ERR_INVALID, "null argument");
  if (p->nx <= 0 || p->ny <= 0 || p->ns <= 0 || p->max_depth < 0) return fail(TPT_ERR_INVALID, "bad image/sample parameters");
  if ((long long)p->nx * p->ny > (1LL << 28)) return fail(TPT_ERR_UNSUPPORTED, "image too large");
  if (p->max_depth > 65535) return fail(TPT_ERR_UNSUPPORTED, "max_depth above 65535 (the slot's depth word carries flags above bit 20)");
  int slices = p->slices > 0 ? p->slices : 1;
  int per_slice = p->ns / slices;
  if (per_slice <= 0) return fail(TPT_ERR_INVALID, "ns < slices (the reference would divide by zero, main.cpp:113,127)");
  if (p->mode != TPT_MODE_PARITY && p->mode != TPT_MODE_FAST) return fail(TPT_ERR_INVALID, "unknown mode");
  if (p->kernel != TPT_KERNEL_MEGA && p->kernel != TPT_KERNEL_WAVEFRONT) return fail(TPT_ERR_UNSUPPORTED, "unknown kernel variant");
  plan.wavefront = p->kernel == TPT_KERNEL_WAVEFRONT;
  if (plan.wavefront && (p->nx > 32767 || p->ny > 32767))
    return fail(TPT_ERR_UNSUPPORTED, "wavefront variant packs pixel coordinates in 16 bits: use TPT_KERNEL_MEGA above 32767");
  if (p->part_count <= 0 || p->part_index < 0 || p->part_index >= p->part_count) return fail(TPT_ERR_INVALID, "bad part_index/part_count");
  if (p->part_count > TPT_MAX_BATCHES) return fail(TPT_ERR_UNSUPPORTED, "part_count above TPT_MAX_BATCHES");
  if (!s->has_lights) return fail(TPT_ERR_INVALID, "light-sampling list is empty (color() needs light_shape, main.cpp:99-106)");
  plan.parity = p->mode == TPT_MODE_PARITY;
  RenderArgs &A = plan.args;
  std::memset(&A, 0, sizeof(A));
  A.scene = s->layout;
  if (plan.parity) A.flat = s->flat; // the two tables share storage: a launch reads the one of its mode
  else A.small = s->small;
  // moving-sphere boxes of the fast BVH cover the spheres' own [time0,time1]; a shutter interval
  // outside it extrapolates the centres, so fall back to the reference tree for that render
  if (s->fbvh_has_moving && !(std::min(cam->time0, cam->time1) >= s->fbvh_t0 && std::max(cam->time0, cam->time1) <= s->fbvh_t1))
    A.scene.fbvh_time_ok = 0;
  auto v3 = [](const float *f) { return V3{f[0], f[1], f[2]}; };
  A.cam.origin = v3(cam->origin);
  A.cam.llc = v3(cam->lower_left_corner);
  A.cam.vertical = v3(cam->vertical);
  A.cam.horizontal = v3(cam->horizontal);
  A.cam.u = v3(cam->u);
  A.cam.v = v3(cam->v);
  A.cam.w = v3(cam->w);
  A.cam.lens_radius = cam->lens_radius;
  A.cam.time0 = cam->time0;
  A.cam.time1 = cam->time1;
  A.nx = p->nx;
  A.ny = p->ny;
  A.ns = p->ns;
  A.max_depth = p->max_depth;
  A.t_min = p->t_min;
  for (uint32_t r = 0; r < 10; r++) { // Philox key schedule, uniform over the launch
    A.rk[2 * r] = p->seed_lo + r * 0x9E3779B9u;
    A.rk[2 * r + 1] = p->seed_hi + r * 0xBB67AE85u;
  }
  // Pixel-bundle bounds test (RenderArgs::cull): black background, world-space root box, shutter
  // inside the interval the moving spheres' boxes were built for. reserved[2] = 1 turns it off.
  // The same rectangle also tells whether ANY pixel of the frame can look past the scene: only then
  // does the wavefront kernel's generate step test fresh camera rays against the scene's bounds and
  // redraw the ones that miss (RenderArgs::camera_tries); a frame-filling or interior camera skips
  // the test altogether.
  A.cull = 0;
  A.camera_tries = 1;
  A.tex_mask = s->tex_mask;
  const bool want_cull = p->reserved[2] == 0 && s->background == TPT_BG_BLACK;
  if (s->root_box_ok &&
      (!s->any_moving || (std::min(cam->time0, cam->time1) >= s->moving_t0 && std::max(cam->time0, cam->time1) <= s->moving_t1))) {
    auto dotd = [](const double *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    double q[3];
    for (int k = 0; k < 3; k++) q[k] = (double)cam->lower_left_corner[k] - cam->origin[k];
    const double fd = -dotd(q, cam->w);
    const double hlen[3] = {cam->horizontal[0], cam->horizontal[1], cam->horizontal[2]};
    const double vlen[3] = {cam->vertical[0], cam->vertical[1], cam->vertical[2]};
    // the frame must be the one camera_with_blur builds (src/camera.cc:2-21): horizontal along u,
    // vertical along v, focus plane perpendicular to w at distance fd > 0
    const double h_off = std::fabs(dotd(hlen, cam->v)) + std::fabs(dotd(hlen, cam->w));
    const double v_off = std::fabs(dotd(vlen, cam->u)) + std::fabs(dotd(vlen, cam->w));
    const double span = std::fabs(dotd(hlen, cam->u)) + std::fabs(dotd(vlen, cam->v));
    if (fd > 0 && std::isfinite(fd) && span > 0 && h_off <= 1e-4 * span && v_off <= 1e-4 * span) {
      // Camera coordinates: x along u, y along v, z forward (-w). The rays of pixel column px run
      // from a lens point (|x| <= R at z = 0) through the column's footprint [fx0, fx1] on the
      // focus plane z = fd, so at depth z = l * fd they satisfy
      //     x >= l*fx0 - |1-l|*R >= l*(fx0 - R) - R        (likewise x <= l*(fx1 + R) + R):
      // two planes bound them all. The column cannot see the box when its eight corners are all on
      // the wrong side of one of the planes; rows alike; nothing is visible when the box is behind
      // the lens plane. The visible set is therefore a pixel rectangle. Conservative: the box is
      // padded as in may_hit_world, the footprint widened by `slack`.
      const double inv_fd = 1.0 / fd;
      const double xl = dotd(q, cam->u), yb = dotd(q, cam->v);
      const double hl = dotd(hlen, cam->u), vl = dotd(vlen, cam->v);
      const double slack = 1e-4 * span + std::fabs((double)cam->lens_radius) * 1.001;
      double X[8], Y[8], Z[8];
      bool finite = true, behind = true;
      for (int c = 0; c < 8; c++) {
        double pw[3];
        for (int k = 0; k < 3; k++) {
          const double lo = s->root_lo[k], hi = s->root_hi[k];
          const double pad = 1e-3 * std::max(std::fabs(lo), std::fabs(hi)) + 1e-3;
          pw[k] = ((c >> k) & 1 ? hi + pad : lo - pad) - cam->origin[k];
        }
        X[c] = dotd(pw, cam->u);
        Y[c] = dotd(pw, cam->v);
        Z[c] = -dotd(pw, cam->w);
        finite = finite && std::isfinite(X[c]) && std::isfinite(Y[c]) && std::isfinite(Z[c]);
        behind = behind && Z[c] < 0;
      }
      auto visible_span = [&](int n, double first, double len, const double *C, int &i0, int &i1) {
        i0 = n;
        i1 = -1;
        for (int i = 0; i < n; i++) {
          const double a = first + (double)i / n * len, b = first + (double)(i + 1) / n * len;
          const double lo = std::min(a, b) - slack, hi = std::max(a, b) + slack;
          bool out_lo = true, out_hi = true;
          for (int c = 0; c < 8; c++) {
            const double l = Z[c] * inv_fd;
            out_lo = out_lo && (C[c] < l * lo - slack);
            out_hi = out_hi && (C[c] > l * hi + slack);
          }
          if (!(out_lo || out_hi)) {
            i0 = std::min(i0, i);
            i1 = std::max(i1, i);
          }
        }
      };
      if (finite) {
        A.cull = 1;
        if (behind) {
          A.cull_x0 = A.cull_y0 = 1;
          A.cull_x1 = A.cull_y1 = 0; // empty rectangle
        } else {
          visible_span(p->nx, xl, hl, X, A.cull_x0, A.cull_x1);
          visible_span(p->ny, yb, vl, Y, A.cull_y0, A.cull_y1);
        }
        // nothing to cull (frame-filling or interior camera): run the kernel build without the test
        if (A.cull_x0 <= 0 && A.cull_y0 <= 0 && A.cull_x1 >= p->nx - 1 && A.cull_y1 >= p->ny - 1) A.cull = 0;
        else A.camera_tries = TPT_WAVE_CAMERA_TRIES;
        if (!want_cull) A.cull = 0;
      }
    }
  }
  // sample ranges: each slice is cut into `subs` sub-ranges (finer bins => shorter kernel tail);
  // samples beyond slices*per_slice (ns not divisible) form one tail range.
  // Automatic choice (fast mode): enough bins that every resident path slot works through >= 32 of
  // them -- a slot runs a bin's samples sequentially, so the launch's tail is about one bin long.
  // Matters when a launch covers a fraction of the frame (1/N of the tiles per GPU, 1/(8N) per
  // stolen batch): at 8 GPUs a 256-sample bin would be a fifth of the whole launch.
  int subs = p->reserved[1];
  int tail = p->ns - slices * per_slice;
  if (subs <= 0) {
    subs = 1;
    {
      // both modes: PARITY sums a bin's samples in the reference's order (main.cpp:119-126) and the
      // bins of a pixel in range order -- another association of the same fp32 sum (~1e-7
      // relative, gate 2 allows 1e-4); reserved[1] = 1 keeps one bin per pixel and slice, i.e. the
      // reference's strictly sequential sum
      // resident path slots per SM: 2 CTAs x 1152 (fast) / 768 (parity) on small scenes, 3 x 512 otherwise
      const bool small_scene = plan.parity ? s->flat.enabled != 0 : s->small.enabled != 0;
      const double slots = (double)s->prop.multiProcessorCount *
                           (small_scene ? TPT_SMALL_MIN_BLOCKS * wave_slots(plan.parity, true, false, s->tex_mask == 0) : 1536.0);
      const double local_pixels = (double)p->nx * p->ny / p->part_count;
      long long want = (long long)std::ceil(32.0 * slots / std::max(local_pixels, 1.0) / slices);
      long long by_samples = std::max(1, per_slice / 8);                       // >= 8 samples per bin
      long long by_table = std::max(1, (TPT_MAX_RANGES - (tail ? 1 : 0)) / slices);
      long long by_memory = std::max<long long>(1, (long long)((2.0 * (1 << 30)) / (12.0 * p->nx * p->ny * slices)));
      subs = (int)std::max<long long>(1, std::min(std::min(want, by_samples), std::min(by_table, by_memory)));
    }
  }
  while (subs > 1 && slices * subs + (tail ? 1 : 0) > TPT_MAX_RANGES) subs--;
  if (slices * subs + (tail ? 1 : 0) > TPT_MAX_RANGES) return fail(TPT_ERR_UNSUPPORTED, "too many slices");
  if (subs > per_slice) subs = per_slice;
  ResolveArgs &R = plan.res;
  std::memset(&R, 0, sizeof(R));
  int r = 0;
  for (int sl = 0; sl < slices; sl++) {
    for (int q = 0; q < subs; q++) {
      A.range_bounds[r] = sl * per_slice + (int)((long long)per_slice * q / subs);
      r++;
    }
    R.slice_last_range[sl] = r - 1;
  }
  A.range_bounds[r] = slices * per_slice;
  if (tail) {
    r++;
    A.range_bounds[r] = p->ns;
  }
  A.n_ranges = r;
  A.tiles_x = (p->nx + TPT_TILE - 1) / TPT_TILE;
  A.tiles_y = (p->ny + TPT_TILE - 1) / TPT_TILE;
  A.part_index = p->part_index;
  A.part_count = p->part_count;
  long long n_tiles = (long long)A.tiles_x * A.tiles_y;
  long long local_tiles = (n_tiles - p->part_index + p->part_count - 1) / p->part_count;
  if (local_tiles < 0) local_tiles = 0;
  A.n_bins = (unsigned long long)local_tiles * TPT_TILE * TPT_TILE * (unsigned long long)A.n_ranges;
  if (A.n_bins >= (1ULL << 32)) return fail(TPT_ERR_UNSUPPORTED, "too many bins for one launch");
  plan.npix = (size_t)p->nx * p->ny;
  R.n_ranges = A.n_ranges;
  R.nx = p->nx;
  R.tiles_x = A.tiles_x;
  R.part_count = p->part_count;
  R.owned[p->part_index >> 5] |= 1u << (p->part_index & 31);
  R.npix = (int)plan.npix;
  R.ns = p->ns;
  R.slices = slices;
  R.per_slice = per_slice;
  return TPT_OK;
}

int run_plan(tpt_scene *s, Plan &plan, bool want_sum, bool want_rgb8, bool want_slices) {
  CK(cudaSetDevice(s->device));
  RenderArgs &A = plan.args;
  ResolveArgs &R = plan.res;
  int rc;
  size_t acc_want = (size_t)A.n_ranges * plan.npix * 3 * sizeof(float);
  if ((rc = ensure(s->stream, (void **)&s->d_acc, s->acc_bytes, acc_want)) != TPT_OK) return rc;
  size_t sum_want = (size_t)R.slices * plan.npix * 3 * sizeof(float);
  if ((rc = ensure(s->stream, (void **)&s->d_sum, s->sum_bytes, sum_want, s->products, &s->products_private)) != TPT_OK) return rc;
  if ((rc = ensure(s->stream, (void **)&s->d_rgb8, s->rgb8_bytes, plan.npix * 3, s->products, &s->products_private)) != TPT_OK) return rc;
  if ((rc = ensure(s->stream, (void **)&s->d_rgb8_slices, s->rgb8_slices_bytes, (size_t)R.slices * plan.npix * 3, s->products, &s->products_private)) != TPT_OK) return rc;
  A.acc = s->d_acc;
  A.counters = s->d_counters;
  R.acc = s->d_acc;
  R.sum_rgb = s->d_sum;
  R.rgb8 = s->d_rgb8;
  R.rgb8_slices = want_slices ? s->d_rgb8_slices : nullptr;
  (void)want_sum;
  (void)want_rgb8;

  size_t smem = s->use_smem ? s->blob_bytes : 0;
  int bps = 0;
  const bool media = s->n_mediums > 0;
  const bool small = (plan.parity ? s->flat.enabled != 0 : s->small.enabled != 0) && !media;
  const bool trace = TPT_TRACE_ENABLE && !plan.parity && (media || TPT_TRACE_ALL) && A.scene.n_fbvh > 0 && A.scene.fbvh_time_ok; // SAH BVH + media: 2-CTA variant with dynamic ray hand-out
  if (plan.wavefront)
    CK(plan.parity ? wave_occupancy_parity(A, small, s->use_smem, media, trace, &bps) : wave_occupancy_fast(A, small, s->use_smem, media, trace, &bps));
  else
    CK(plan.parity ? mega_occupancy_parity(A, s->use_smem, small, media, smem, &bps)
                   : mega_occupancy_fast(A, s->use_smem, small, media, smem, &bps));
  if (bps < 1) return fail(TPT_ERR_CUDA, "render kernel does not fit on an SM");
  plan.blocks_per_sm = bps;
  plan.blocks = bps * s->prop.multiProcessorCount;

  CK(cudaMemsetAsync(s->d_counters, 0, 8 * sizeof(unsigned long long), s->stream));
  CK(cudaEventRecord(s->ev[0], s->stream));
  if (plan.wavefront)
    CK(plan.parity ? launch_wave_parity(A, small, s->use_smem, media, trace, plan.blocks, s->stream) : launch_wave_fast(A, small, s->use_smem, media, trace, plan.blocks, s->stream));
  else
    CK(plan.parity ? launch_mega_parity(A, s->use_smem, small, media, plan.blocks, s->stream)
                   : launch_mega_fast(A, s->use_smem, small, media, plan.blocks, s->stream));
  CK(cudaEventRecord(s->ev[1], s->stream));
  int rb = (int)((plan.npix + 255) / 256);
  resolve_kernel<<<rb, 256, 0, s->stream>>>(R);
  CK(cudaGetLastError());
  CK(cudaEventRecord(s->ev[2], s->stream));
  CK(cudaStreamSynchronize(s->stream));
  unsigned long long c[5];
  CK(cudaMemcpy(c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
  float ms_render = 0, ms_resolve = 0;
  CK(cudaEventElapsedTime(&ms_render, s->ev[0], s->ev[1]));
  CK(cudaEventElapsedTime(&ms_resolve, s->ev[1], s->ev[2]));
  tpt_stats &st = s->stats;
  std::memset(&st, 0, sizeof(st));
  st.paths = c[3];
  st.rays = c[1];
  st.nan_samples = c[2];
  st.culled_paths = c[4];
  st.render_ms = ms_render;
  st.resolve_ms = ms_resolve;
  st.kernel_launches = 2;
  st.sm_count = s->prop.multiProcessorCount;
  st.blocks = plan.blocks;
  st.threads_per_block = !plan.wavefront ? TPT_MEGA_THREADS
                                          : (plan.parity ? wave_threads_parity(A, small, s->use_smem, media, trace)
                                                         : wave_threads_fast(A, small, s->use_smem, media, trace));
  st.reserved[0] = A.n_ranges; // sample ranges per pixel (accumulator planes written by the kernel)
  st.h2d_bytes = s->blob_bytes + sizeof(RenderArgs) + sizeof(ResolveArgs); // scene blob + launch arguments
  s->last_nx = A.nx;
  s->last_ny = A.ny;
  s->last_slices = R.slices;
  return TPT_OK;
}

int fetch(tpt_scene *s, tpt_image *out) {
  if (!out) return TPT_OK;
  CK(cudaSetDevice(s->device));
  size_t npix = (size_t)s->last_nx * s->last_ny;
  if (npix == 0) return fail(TPT_ERR_INVALID, "nothing rendered yet");
  double t0 = now_ms();
  uint64_t bytes = 0;
  if (out->sum_rgb) {
    size_t b = (size_t)s->last_slices * npix * 3 * sizeof(float);
    CK(cudaMemcpy(out->sum_rgb, s->d_sum, b, cudaMemcpyDeviceToHost));
    bytes += b;
  }
  if (out->rgb8) {
    CK(cudaMemcpy(out->rgb8, s->d_rgb8, npix * 3, cudaMemcpyDeviceToHost));
    bytes += npix * 3;
  }
  if (out->rgb8_slices) {
    size_t b = (size_t)s->last_slices * npix * 3;
    CK(cudaMemcpy(out->rgb8_slices, s->d_rgb8_slices, b, cudaMemcpyDeviceToHost));
    bytes += b;
  }
  s->stats.d2h_ms = now_ms() - t0;
  s->stats.d2h_bytes = bytes;
  return TPT_OK;
}

// ---------------------------------------------------------------------------------------------
// In-process multi-GPU render: static split + work stealing + gather.
// The frame is cut into NB = 8 x n_gpus batches; batch b = the tiles t with t % NB == b (every
// batch is a uniform sample of the frame, so batches cost about the same). One host thread per
// GPU first renders its static share (7/8 of the batches, b % n_gpus == g, in ONE launch) and then steals the
// remaining batches from a shared atomic counter, so a slower or busier GPU simply takes fewer.
// Every batch is one launch of the persistent kernel into that GPU's own accumulators (bins are
// disjoint); each GPU then resolves its partial frame, GPU 0 pulls the others over NVLink
// (cudaMemcpyPeerAsync) and adds them -- tiles a GPU did not render are zero, so the sum is an
// exact gather -- and one download serves the caller. No NCCL: nothing else is exchanged.
// ---------------------------------------------------------------------------------------------
__global__ void add_f32_kernel(float *dst, const float *src, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
__global__ void add_u8_kernel(uint8_t *dst, const uint8_t *src, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (uint8_t)(dst[i] + src[i]);
}

// Tile-owned gather: every pixel is fetched from the ONE GPU that rendered its batch (GPU 0 reads (n-1)/n of
// one frame over NVLink instead of n-1 whole frames of mostly zeros through the add kernels above).
struct GatherArgs {
  const float *sum[TPT_MAX_GPUS];
  const uint8_t *rgb8[TPT_MAX_GPUS];
  const uint8_t *rgb8_slices[TPT_MAX_GPUS];
  unsigned char owner[TPT_MAX_BATCHES]; // GPU that rendered batch b
  int npix, nx, tiles_x, n_batches, slices;
  float *dst_sum;
  uint8_t *dst_rgb8, *dst_rgb8_slices;
};
__global__ void gather_owned_kernel(const __grid_constant__ GatherArgs G) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= G.npix) return;
  const int py = pix / G.nx, px = pix - py * G.nx;
  const int g = G.owner[((py / TPT_TILE) * G.tiles_x + (px / TPT_TILE)) % G.n_batches];
  if (g == 0) return; // GPU 0's own tiles are already in place
  for (int sl = 0; sl < G.slices; sl++) {
    const size_t o = ((size_t)sl * G.npix + pix) * 3;
    G.dst_sum[o] = G.sum[g][o];
    G.dst_sum[o + 1] = G.sum[g][o + 1];
    G.dst_sum[o + 2] = G.sum[g][o + 2];
    if (G.dst_rgb8_slices) {
      G.dst_rgb8_slices[o] = G.rgb8_slices[g][o];
      G.dst_rgb8_slices[o + 1] = G.rgb8_slices[g][o + 1];
      G.dst_rgb8_slices[o + 2] = G.rgb8_slices[g][o + 2];
    }
  }
  const size_t o = (size_t)pix * 3;
  G.dst_rgb8[o] = G.rgb8[g][o];
  G.dst_rgb8[o + 1] = G.rgb8[g][o + 1];
  G.dst_rgb8[o + 2] = G.rgb8[g][o + 2];
}

struct MultiWorker {
  tpt_scene *s = nullptr;
  int rc = TPT_OK;
  std::string err;
  int batches = 0, stolen = 0;
  double busy_ms = 0;
  unsigned owned[TPT_MAX_BATCHES / 32] = {};
};

int prepare_buffers(tpt_scene *s, Plan &plan, bool want_slices) {
  CK(cudaSetDevice(s->device));
  RenderArgs &A = plan.args;
  ResolveArgs &R = plan.res;
  int rc;
  size_t acc_want = (size_t)A.n_ranges * plan.npix * 3 * sizeof(float);
  if ((rc = ensure(s->stream, (void **)&s->d_acc, s->acc_bytes, acc_want)) != TPT_OK) return rc;
  if ((rc = ensure(s->stream, (void **)&s->d_sum, s->sum_bytes, (size_t)R.slices * plan.npix * 3 * sizeof(float), s->products, &s->products_private)) != TPT_OK) return rc;
  if ((rc = ensure(s->stream, (void **)&s->d_rgb8, s->rgb8_bytes, plan.npix * 3, s->products, &s->products_private)) != TPT_OK) return rc;
  if ((rc = ensure(s->stream, (void **)&s->d_rgb8_slices, s->rgb8_slices_bytes, (size_t)R.slices * plan.npix * 3, s->products, &s->products_private)) != TPT_OK) return rc;
  R.acc = s->d_acc;
  R.sum_rgb = s->d_sum;
  R.rgb8 = s->d_rgb8;
  R.rgb8_slices = want_slices ? s->d_rgb8_slices : nullptr;
  CK(cudaMemsetAsync(s->d_counters, 0, TPT_MAX_BATCHES * 8 * sizeof(unsigned long long), s->stream));
  return TPT_OK;
}

// one launch for batch `batch`, or (group > 1) for the batches batch, batch + stride, ..., `group` of them
int launch_batch(tpt_scene *s, const tpt_camera *cam, const tpt_render_params *p, int batch, int n_batches, int group = 1,
                 int stride = 0) {
  tpt_render_params bp = *p;
  bp.part_index = batch;
  bp.part_count = n_batches;
  Plan plan;
  int rc = make_plan(s, cam, &bp, plan);
  if (rc != TPT_OK) return rc;
  RenderArgs &A = plan.args;
  if (group > 1) {
    A.part_group = group;
    A.part_stride = stride;
    const unsigned long long n_tiles = (unsigned long long)A.tiles_x * A.tiles_y;
    const unsigned long long periods = (n_tiles + n_batches - 1) / n_batches; // tiles past the frame decode to rows >= ny and are skipped
    A.n_bins = periods * group * TPT_TILE * TPT_TILE * (unsigned long long)A.n_ranges;
    if (A.n_bins >= (1ULL << 32)) return fail(TPT_ERR_UNSUPPORTED, "too many bins for one launch");
  }
  A.acc = s->d_acc;
  A.counters = s->d_counters + (size_t)batch * 8;
  size_t smem = s->use_smem ? s->blob_bytes : 0;
  int bps = 0;
  const bool media = s->n_mediums > 0;
  const bool small = (plan.parity ? s->flat.enabled != 0 : s->small.enabled != 0) && !media;
  const bool trace = TPT_TRACE_ENABLE && !plan.parity && (media || TPT_TRACE_ALL) && A.scene.n_fbvh > 0 && A.scene.fbvh_time_ok; // SAH BVH + media: 2-CTA variant with dynamic ray hand-out
  if (plan.wavefront)
    CK(plan.parity ? wave_occupancy_parity(A, small, s->use_smem, media, trace, &bps) : wave_occupancy_fast(A, small, s->use_smem, media, trace, &bps));
  else
    CK(plan.parity ? mega_occupancy_parity(A, s->use_smem, small, media, smem, &bps)
                   : mega_occupancy_fast(A, s->use_smem, small, media, smem, &bps));
  if (bps < 1) return fail(TPT_ERR_CUDA, "render kernel does not fit on an SM");
  int blocks = bps * s->prop.multiProcessorCount;
  if (plan.wavefront)
    CK(plan.parity ? launch_wave_parity(A, small, s->use_smem, media, trace, blocks, s->stream) : launch_wave_fast(A, small, s->use_smem, media, trace, blocks, s->stream));
  else
    CK(plan.parity ? launch_mega_parity(A, s->use_smem, small, media, blocks, s->stream)
                   : launch_mega_fast(A, s->use_smem, small, media, blocks, s->stream));
  s->stats.blocks = blocks;
  s->stats.threads_per_block = !plan.wavefront ? TPT_MEGA_THREADS
                                               : (plan.parity ? wave_threads_parity(A, small, s->use_smem, media, trace)
                                                              : wave_threads_fast(A, small, s->use_smem, media, trace));
  return TPT_OK;
}

int render_multi(tpt_scene *const *scenes, int n, const tpt_camera *cam, const tpt_render_params *p, tpt_image *out) {
  if (!scenes || n < 1 || !cam || !p) return fail(TPT_ERR_INVALID, "null argument");
  for (int i = 0; i < n; i++) {
    if (!scenes[i]) return fail(TPT_ERR_INVALID, "null scene");
    for (int j = 0; j < i; j++)
      if (scenes[j]->device == scenes[i]->device) return fail(TPT_ERR_INVALID, "two scenes on the same device");
  }
  if (p->part_count != 1 || p->part_index != 0) return fail(TPT_ERR_INVALID, "tpt_render_multi partitions the frame itself");
  if (n > TPT_MAX_GPUS) return fail(TPT_ERR_UNSUPPORTED, "more than TPT_MAX_GPUS scenes");
  const double t_begin = now_ms();
  const int n_batches = std::min(TPT_MAX_BATCHES, 8 * n);
  // static share: 7 of every GPU's 8 batches (a multiple of n: every GPU gets the same), one launch; the last
  // eighth of the frame is handed out through the counter. (r02: with 3/4 static every GPU ended up taking
  // exactly its two stealable batches anyway -- identical GPUs, interleaved tiles -- and each extra launch
  // costs its own ramp-up and tail, ~0.5 ms of a 37 ms frame on 8 GPUs.)
  const int n_static = (n_batches * 7 / 8) / n * n;
  Plan plan0;
  int rc = make_plan(scenes[0], cam, p, plan0);
  if (rc != TPT_OK) return rc;
  const bool want_slices = out && out->rgb8_slices;
  std::atomic<int> next(n_static);
  std::vector<MultiWorker> W(n);
  std::vector<std::thread> threads;
  for (int g = 0; g < n; g++) {
    W[g].s = scenes[g];
    threads.emplace_back([&, g]() {
      MultiWorker &w = W[g];
      tpt_scene *s = w.s;
      Plan plan; // sized like a batch launch (same sample ranges as launch_batch will use)
      tpt_render_params pp = *p;
      pp.part_index = 0;
      pp.part_count = n_batches;
      if ((w.rc = make_plan(s, cam, &pp, plan)) != TPT_OK || (w.rc = prepare_buffers(s, plan, want_slices)) != TPT_OK) {
        w.err = g_error;
        return;
      }
      cudaEventRecord(s->ev[0], s->stream);
      // static share: batches g, g + n, ... below n_static, rendered by ONE launch (r02: eight launches of 1/64
      // of the frame each paid eight kernel tails -- 39.5 ms of GPU time per 2048-spp frame against 35.8 ms
      // for one launch of the same tiles)
      const int n_mine = n_static / n;
      if (n_mine > 0) w.rc = launch_batch(s, cam, p, g, n_batches, n_mine, n);
      for (int b = g; b < n_static; b += n) {
        w.batches++;
        w.owned[b >> 5] |= 1u << (b & 31);
      }
      // steal: when this GPU has drained its static share it takes the next free batch, one at a time
      while (w.rc == TPT_OK) {
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) {
          w.rc = TPT_ERR_CUDA;
          g_error = "stream synchronize failed while stealing";
          break;
        }
        int b = next.fetch_add(1);
        if (b >= n_batches) break;
        w.rc = launch_batch(s, cam, p, b, n_batches);
        w.batches++;
        w.stolen++;
        w.owned[b >> 5] |= 1u << (b & 31);
      }
      if (w.rc != TPT_OK) {
        w.err = g_error;
        return;
      }
      cudaEventRecord(s->ev[1], s->stream);
      ResolveArgs R = plan.res;
      R.part_count = n_batches;
      std::memcpy(R.owned, w.owned, sizeof(R.owned));
      int rb = (int)((plan.npix + 255) / 256);
      resolve_kernel<<<rb, 256, 0, s->stream>>>(R);
      cudaEventRecord(s->ev[2], s->stream);
      cudaError_t e = cudaStreamSynchronize(s->stream);
      if (e != cudaSuccess) {
        w.rc = TPT_ERR_CUDA;
        w.err = cudaGetErrorString(e);
        return;
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]);
      w.busy_ms = ms;
      s->last_nx = p->nx;
      s->last_ny = p->ny;
      s->last_slices = plan.res.slices;
    });
  }
  for (auto &t : threads) t.join();
  for (int g = 0; g < n; g++)
    if (W[g].rc != TPT_OK) return fail(W[g].rc, "GPU " + std::to_string(scenes[g]->device) + ": " + W[g].err);
  // ---- gather on GPU 0 over NVLink ----
  tpt_scene *s0 = scenes[0];
  CK(cudaSetDevice(s0->device));
  const size_t npix = plan0.npix;
  const size_t sum_n = (size_t)plan0.res.slices * npix * 3, rgb_n = npix * 3;
  double t_gather = now_ms();
  bool all_peers = n > 1;
  for (int g = 1; g < n && all_peers; g++) { // peer access + pool access for every GPU: then ONE gather kernel
    int can = 0;
    cudaDeviceCanAccessPeer(&can, s0->device, scenes[g]->device);
    if (can) {
      cudaError_t e = cudaDeviceEnablePeerAccess(scenes[g]->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) can = 0;
    }
    cudaMemAccessDesc acc_desc{};
    acc_desc.location.type = cudaMemLocationTypeDevice;
    acc_desc.location.id = s0->device;
    acc_desc.flags = cudaMemAccessFlagsProtReadWrite;
    if (!can || !scenes[g]->products || scenes[g]->products_private ||
        cudaMemPoolSetAccess(scenes[g]->products, &acc_desc, 1) != cudaSuccess) {
      cudaGetLastError();
      all_peers = false;
    }
  }
  if (all_peers) {
    GatherArgs G;
    std::memset(&G, 0, sizeof(G));
    for (int g = 0; g < n; g++) {
      G.sum[g] = scenes[g]->d_sum;
      G.rgb8[g] = scenes[g]->d_rgb8;
      G.rgb8_slices[g] = scenes[g]->d_rgb8_slices;
      for (int b = 0; b < n_batches; b++)
        if ((W[g].owned[b >> 5] >> (b & 31)) & 1u) G.owner[b] = (unsigned char)g;
    }
    G.npix = (int)npix;
    G.nx = p->nx;
    G.tiles_x = plan0.args.tiles_x;
    G.n_batches = n_batches;
    G.slices = plan0.res.slices;
    G.dst_sum = s0->d_sum;
    G.dst_rgb8 = s0->d_rgb8;
    G.dst_rgb8_slices = want_slices ? s0->d_rgb8_slices : nullptr;
    gather_owned_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s0->stream>>>(G);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s0->stream));
  } else if (n > 1) {
    float *tmp_f = nullptr; // only when a peer is not directly addressable
    uint8_t *tmp_b = nullptr;
    for (int g = 1; g < n; g++) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, s0->device, scenes[g]->device);
      if (can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(scenes[g]->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) can = 0;
      }
      if (can) {
        // the products live in GPU g's stream-ordered pool: pools are private to their device
        // until access is granted explicitly (cudaDeviceEnablePeerAccess does not cover them)
        cudaMemAccessDesc acc_desc{};
        acc_desc.location.type = cudaMemLocationTypeDevice;
        acc_desc.location.id = s0->device;
        acc_desc.flags = cudaMemAccessFlagsProtReadWrite;
        if (!scenes[g]->products || scenes[g]->products_private ||
            cudaMemPoolSetAccess(scenes[g]->products, &acc_desc, 1) != cudaSuccess) {
          cudaGetLastError();
          can = 0;
        }
      }
      const float *src_f = scenes[g]->d_sum;
      const uint8_t *src_b = scenes[g]->d_rgb8, *src_s = scenes[g]->d_rgb8_slices;
      if (can) {
        // GPU 0 reads the peer's buffers in place over NVLink: the add IS the transfer
        add_f32_kernel<<<(unsigned)((sum_n + 255) / 256), 256, 0, s0->stream>>>(s0->d_sum, src_f, sum_n);
        add_u8_kernel<<<(unsigned)((rgb_n + 255) / 256), 256, 0, s0->stream>>>(s0->d_rgb8, src_b, rgb_n);
        if (want_slices)
          add_u8_kernel<<<(unsigned)((sum_n + 255) / 256), 256, 0, s0->stream>>>(s0->d_rgb8_slices, src_s, sum_n);
      } else {
        if (!tmp_f) {
          CK(cudaMalloc((void **)&tmp_f, sum_n * sizeof(float)));
          CK(cudaMalloc((void **)&tmp_b, std::max(rgb_n, sum_n)));
        }
        CK(cudaMemcpyPeerAsync(tmp_f, s0->device, src_f, scenes[g]->device, sum_n * sizeof(float), s0->stream));
        add_f32_kernel<<<(unsigned)((sum_n + 255) / 256), 256, 0, s0->stream>>>(s0->d_sum, tmp_f, sum_n);
        CK(cudaMemcpyPeerAsync(tmp_b, s0->device, src_b, scenes[g]->device, rgb_n, s0->stream));
        add_u8_kernel<<<(unsigned)((rgb_n + 255) / 256), 256, 0, s0->stream>>>(s0->d_rgb8, tmp_b, rgb_n);
        if (want_slices) {
          CK(cudaMemcpyPeerAsync(tmp_b, s0->device, src_s, scenes[g]->device, sum_n, s0->stream));
          add_u8_kernel<<<(unsigned)((sum_n + 255) / 256), 256, 0, s0->stream>>>(s0->d_rgb8_slices, tmp_b, sum_n);
        }
      }
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s0->stream));
    if (tmp_f) cudaFree(tmp_f);
    if (tmp_b) cudaFree(tmp_b);
  }
  t_gather = now_ms() - t_gather;
  // ---- statistics: sums over GPUs and batches, time = slowest GPU ----
  tpt_stats st;
  std::memset(&st, 0, sizeof(st));
  for (int g = 0; g < n; g++) {
    CK(cudaSetDevice(scenes[g]->device));
    std::vector<unsigned long long> c((size_t)n_batches * 8);
    CK(cudaMemcpy(c.data(), scenes[g]->d_counters, c.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int b = 0; b < n_batches; b++) {
      st.paths += c[(size_t)b * 8 + 3];
      st.rays += c[(size_t)b * 8 + 1];
      st.nan_samples += c[(size_t)b * 8 + 2];
      st.culled_paths += c[(size_t)b * 8 + 4];
    }
    st.render_ms = std::max(st.render_ms, W[g].busy_ms);
    st.kernel_launches += (n_static / n > 0 ? 1 : 0) + W[g].stolen + 1; // static share + stolen batches + resolve
    st.reserved[g < 4 ? g : 3] = W[g].batches; // batches taken by the first GPUs (load-balance evidence)
    if (g < TPT_MAX_GPUS) {
      st.multi_batches[g] = W[g].batches;
      st.multi_stolen[g] = W[g].stolen;
      st.multi_busy_ms[g] = W[g].busy_ms;
    }
  }
  st.multi_gpus = n;
  st.multi_batches_total = n_batches;
  st.multi_gather_ms = t_gather;
  st.resolve_ms = t_gather;
  st.sm_count = s0->prop.multiProcessorCount;
  st.blocks = s0->stats.blocks;
  st.threads_per_block = s0->stats.threads_per_block;
  st.h2d_bytes = (s0->blob_bytes + sizeof(RenderArgs) + sizeof(ResolveArgs)) * n;
  s0->stats = st;
  rc = fetch(s0, out);
  s0->stats.wall_ms = now_ms() - t_begin;
  return rc;
}

} // namespace

extern "C" {

int tpt_api_version(void) { return TPT_API_VERSION; }

int tpt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char *tpt_last_error(void) { return g_error.c_str(); }

int tpt_device_warm(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TPT_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return fail(TPT_ERR_INVALID, "device ordinal out of range");
  CK(cudaSetDevice(device));
  CK(cudaFree(nullptr)); // forces context creation
  return TPT_OK;
}

static int scene_create_impl(const tpt_scene_desc *d, int device, tpt_scene **out, tpt_scene **partial);

int tpt_scene_create(const tpt_scene_desc *d, int device, tpt_scene **out) {
  // every early return of the builder below leaves what it had allocated so far in `partial`: released here,
  // so a failed create leaks neither the stream, the events, device memory nor the texture arrays
  tpt_scene *partial = nullptr;
  const int rc = scene_create_impl(d, device, out, &partial);
  if (rc != TPT_OK && partial) {
    const std::string keep = g_error;
    tpt_scene_destroy(partial);
    cudaGetLastError();
    g_error = keep;
    if (out) *out = nullptr;
  }
  return rc;
}

static int scene_create_impl(const tpt_scene_desc *d, int device, tpt_scene **out, tpt_scene **partial) {
  if (!out) return fail(TPT_ERR_INVALID, "null out pointer");
  *out = nullptr;
  int depth = 0;
  int rc = validate_desc(d, depth);
  if (rc != TPT_OK) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TPT_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return fail(TPT_ERR_INVALID, "device ordinal out of range");
  CK(cudaSetDevice(device));
  tpt_scene *s = new tpt_scene();
  *partial = s;
  s->device = device;
  {
    // cudaGetDeviceProperties is a slow driver query (3 ms typical, 100+ ms now and then): once
    // per device and process; the same for the memory pool's release threshold
    static std::mutex mu;
    struct DeviceOnce { int device; cudaDeviceProp prop; cudaMemPool_t products; };
    static std::vector<DeviceOnce> cache;
    std::lock_guard<std::mutex> lock(mu);
    const DeviceOnce *hit = nullptr;
    for (auto &c : cache)
      if (c.device == device) hit = &c;
    if (!hit) {
      cudaDeviceProp prop;
      cudaError_t pe = cudaGetDeviceProperties(&prop, device);
      if (pe != cudaSuccess) return fail(TPT_ERR_CUDA, cudaGetErrorString(pe));
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaMemPool_t products = nullptr;
      cudaMemPoolProps pp{};
      pp.allocType = cudaMemAllocationTypePinned;
      pp.handleTypes = cudaMemHandleTypeNone;
      pp.location.type = cudaMemLocationTypeDevice;
      pp.location.id = device;
      if (cudaMemPoolCreate(&products, &pp) == cudaSuccess) {
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(products, cudaMemPoolAttrReleaseThreshold, &keep);
      } else {
        cudaGetLastError();
        products = nullptr; // products fall back to the default pool, the gather to staged copies
      }
      cache.push_back(DeviceOnce{device, prop, products});
      hit = &cache.back();
    }
    s->prop = hit->prop;
    s->products = hit->products;
  }
  if (s->prop.major < 10) return fail(TPT_ERR_NO_DEVICE, "device is not sm_100-class; kernels are built for sm_100a only");
  CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  for (auto &ev : s->ev) CK(cudaEventCreate(&ev));

  // ---- blob: the C structs back to back, each table padded to 16 bytes ----
  std::vector<unsigned char> blob;
  SceneLayout &L = s->layout;
  auto words = [&]() { return (int)(blob.size() / 16); };
  L.off_nodes = words();
  append(blob, d->nodes, d->n_nodes);
  L.off_prims = words();
  append(blob, d->prims, d->n_prims);
  {
    // FAST mode reads a world-space normal (rects) / centre (spheres) from the last four words of a
    // primitive's record (fill_hit): the object-space value carried out of its transform chain,
    // innermost wrapper first (src/rect_box.cc:185-189, headers/rect_box.h:91). The caller's description
    // is not touched; moving spheres and media keep their own data there.
    for (int i = 0; i < d->n_prims; i++) {
      const tpt_prim &p = d->prims[i];
      if (p.kind == TPT_PRIM_MOVING_SPHERE || p.kind == TPT_PRIM_MEDIUM) continue;
      const bool sph = p.kind == TPT_PRIM_SPHERE;
      double v[3] = {0, 0, 0};
      if (sph) { v[0] = p.p[0]; v[1] = p.p[1]; v[2] = p.p[2]; }
      else v[p.kind == TPT_PRIM_XY_RECT ? 2 : (p.kind == TPT_PRIM_XZ_RECT ? 1 : 0)] = 1.0;
      const tpt_chain &ch = d->chains[p.chain];
      for (int k = ch.n_ops - 1; k >= 0; k--) {
        const tpt_xform_op &op = d->xform_ops[ch.first_op + k];
        if (op.kind == TPT_XF_TRANSLATE) {
          if (sph) { v[0] += op.a; v[1] += op.b; v[2] += op.c; }
        } else {
          const double x = op.b * v[0] + op.a * v[2], z = -op.a * v[0] + op.b * v[2];
          v[0] = x;
          v[2] = z;
        }
      }
      const float sign = (p.flags & TPT_PRIM_FLIP) ? -1.0f : 1.0f;
      float w[4] = {(float)v[0], (float)v[1], (float)v[2], sign};
      if (!sph)
        for (int c = 0; c < 3; c++) w[c] *= sign;
      std::memcpy(blob.data() + (size_t)L.off_prims * 16 + (size_t)i * sizeof(tpt_prim) + 16 + 8 * sizeof(float), w, sizeof(w));
    }
  }
  L.off_chains = words();
  {
    std::vector<int32_t> padded((size_t)d->n_chains * 4, 0);
    for (int i = 0; i < d->n_chains; i++) {
      padded[4 * i] = d->chains[i].first_op;
      padded[4 * i + 1] = d->chains[i].n_ops;
    }
    append(blob, padded.data(), padded.size());
  }
  L.off_ops = words();
  append(blob, d->xform_ops, d->n_xform_ops);
  L.off_mats = words();
  append(blob, d->materials, d->n_materials);
  L.off_texs = words();
  append(blob, d->textures, d->n_textures);
  L.off_lights = words();
  append(blob, d->lights, d->n_lights);
  L.off_perlin = words();
  if (d->perlin) {
    std::vector<float> rv(256 * 4, 0.f);
    for (int i = 0; i < 256; i++)
      for (int c = 0; c < 3; c++) rv[4 * i + c] = d->perlin->ranvec[i][c];
    append(blob, rv.data(), rv.size());
    append(blob, d->perlin->perm_x, 256);
    append(blob, d->perlin->perm_y, 256);
    append(blob, d->perlin->perm_z, 256);
  }
  const int n_root = d->n_root_nodes > 0 ? d->n_root_nodes : d->n_nodes;
  std::vector<int32_t> mediums; // MEDIUM primitives in DFS order of the root tree
  for (int i = 0; i < n_root; i++) {
    const tpt_node &nd = d->nodes[i];
    if ((nd.kind & 0xff) == TPT_NODE_LEAF && !(nd.kind & TPT_NODE_DUP) && d->prims[nd.end_or_prim].kind == TPT_PRIM_MEDIUM &&
        std::find(mediums.begin(), mediums.end(), nd.end_or_prim) == mediums.end())
      mediums.push_back(nd.end_or_prim);
  }
  L.off_mediums = words();
  append(blob, mediums.data(), mediums.size());
  L.n_mediums = (int)mediums.size();
  s->n_mediums = L.n_mediums;
  FastBvh fb;
  if (count_root_surface_prims(d, n_root) > TPT_SMALL_MAX_PRIMS) build_fast_bvh(d, n_root, fb);
  int fbvh_depth = 0;
  std::vector<float> wide;
  if (TPT_FBVH_WIDE) collapse_to_bvh4(fb, wide, fbvh_depth);
  // the traversal stack holds at most 3 deferred children per level (1 in the binary form): a tree too
  // deep for it (a degenerate, sliver-by-sliver split) falls back to the reference's own tree
  if (TPT_FBVH_WIDE && 3 * fbvh_depth + 4 > TPT_FBVH_STACK) {
    fb = FastBvh();
    wide.clear();
  }
  if (!TPT_FBVH_WIDE && !fb.nodes.empty()) {
    // binary tree: the traversal defers at most one child per level. A tree deeper than the stack (binned SAH can
    // split 1 : n-1 level after level on geometrically spaced primitives) falls back to the reference's own tree
    // instead of overflowing the per-lane stack.
    std::function<int(int32_t)> depth_of = [&](int32_t node) -> int {
      if (node < 0) return 0;
      int32_t c0, c1;
      std::memcpy(&c0, &fb.nodes[(size_t)node * 16 + 12], 4);
      std::memcpy(&c1, &fb.nodes[(size_t)node * 16 + 13], 4);
      return 1 + std::max(depth_of(c0), depth_of(c1));
    };
    if (depth_of(0) + 2 > TPT_FBVH_STACK) fb = FastBvh();
  }
  // nodes and leaf records are 64 bytes each and 64-byte aligned: a lane fetches one with two 256-bit loads
  // (two L1 wavefronts instead of four: the BVH kernels' L1 data pipe was as busy as their issue slots)
  while (words() % 4 != 0) blob.insert(blob.end(), 16, 0);
  L.off_fbvh = words();
  if (TPT_FBVH_WIDE) append(blob, wide.data(), wide.size());
  else append(blob, fb.nodes.data(), fb.nodes.size());
  while (words() % 4 != 0) blob.insert(blob.end(), 16, 0);
  L.off_fleaf = words();
  {
    // leaf records in leaf order, 4 float4 each (read by closest_hit_fbvh):
    //   q0 = primitive words p[0..3]; q1 = {kind | exact << 8, primitive id, chain, rect k or time0};
    //   q2 = moving sphere {center1, time1}; q3 = padding
    std::vector<float> rec(fb.leaf_prims.size() * 16, 0.f);
    for (size_t i = 0; i < fb.leaf_prims.size(); i++) {
      const int32_t id = fb.leaf_prims[i];
      const tpt_prim &p = d->prims[id];
      float *q = &rec[i * 16];
      std::memcpy(q, p.p, 16);
      const bool sph = p.kind == TPT_PRIM_SPHERE || p.kind == TPT_PRIM_MOVING_SPHERE;
      int32_t kf = p.kind | (sph && p.p[3] >= 500.0f ? 0x100 : 0); // huge "wall" spheres: exact roots
      std::memcpy(q + 4, &kf, 4);
      std::memcpy(q + 5, &id, 4);
      std::memcpy(q + 6, &p.chain, 4);
      q[7] = p.kind == TPT_PRIM_MOVING_SPHERE ? p.p[7] : p.p[4];
      if (p.kind == TPT_PRIM_MOVING_SPHERE) {
        q[8] = p.p[4]; q[9] = p.p[5]; q[10] = p.p[6]; q[11] = p.p[8];
      }
    }
    append(blob, rec.data(), rec.size());
  }
  L.n_fbvh = TPT_FBVH_WIDE ? (int)(wide.size() / 32) : (int)(fb.nodes.size() / 16);
  {
    // world bounds = the root node's box, when the root sits outside any transform wrapper
    const tpt_node &root = d->nodes[0];
    s->root_box_ok = (root.kind >> 16) == 0;
    for (int k = 0; k < 3; k++) {
      s->root_lo[k] = root.bmin[k];
      s->root_hi[k] = root.bmax[k];
      if (!std::isfinite(root.bmin[k]) || !std::isfinite(root.bmax[k]) || !(root.bmin[k] <= root.bmax[k])) s->root_box_ok = false;
    }
    s->moving_t0 = -FLT_MAX; // intersection of the moving spheres' own [time0, time1]
    s->moving_t1 = FLT_MAX;
    for (int i = 0; i < d->n_prims; i++)
      if (d->prims[i].kind == TPT_PRIM_MOVING_SPHERE) {
        s->any_moving = true;
        s->moving_t0 = std::max(s->moving_t0, std::min(d->prims[i].p[7], d->prims[i].p[8]));
        s->moving_t1 = std::min(s->moving_t1, std::max(d->prims[i].p[7], d->prims[i].p[8]));
      }
    s->background = d->background;
  }
  L.fbvh_time_ok = 1;
  {
    // closest_hit_skip (PARITY) needs: no medium, no bvh_node anywhere below a hitable_list
    bool simple = mediums.empty();
    int list_until = -1;
    for (int i = 0; simple && i < n_root; i++) {
      const int k = d->nodes[i].kind & 0xff;
      if (i >= list_until) list_until = -1;
      if (k == TPT_NODE_BVH && list_until >= 0) simple = false;
      if (k == TPT_NODE_LIST && list_until < 0) list_until = d->nodes[i].end_or_prim;
    }
    if (const char *e = std::getenv("TPT_PARITY_SKIP")) // TPT_PARITY_SKIP=0: keep the frame replay (A/B and test use)
      if (e[0] == '0') simple = false;
    L.tree_simple = simple ? 1 : 0;
    // margin of the skip walk's "behind the best hit" test: 1e-5 of the root box's largest extent (no culling
    // without a finite root box, or with TPT_PARITY_SKIP_CULL=0)
    L.skip_cull_abs = -1.f;
    L.skip_cull_t0 = -FLT_MAX;
    L.skip_cull_t1 = FLT_MAX;
    for (int i = 0; i < d->n_prims; i++)
      if (d->prims[i].kind == TPT_PRIM_MOVING_SPHERE) { // a moving sphere stays inside its boxes for ray times within its own interval only
        L.skip_cull_t0 = std::max(L.skip_cull_t0, std::min(d->prims[i].p[7], d->prims[i].p[8]));
        L.skip_cull_t1 = std::min(L.skip_cull_t1, std::max(d->prims[i].p[7], d->prims[i].p[8]));
      }
    if (s->root_box_ok && simple && boxes_contain_their_prims(d, n_root)) {
      float ext = 0.f;
      for (int k = 0; k < 3; k++) ext = std::max(ext, s->root_hi[k] - s->root_lo[k]);
      if (std::isfinite(ext)) L.skip_cull_abs = 1e-5f * ext;
    }
    if (const char *e = std::getenv("TPT_PARITY_SKIP_CULL"))
      if (e[0] == '0') L.skip_cull_abs = -1.f;
    // closest_hit_skip_ordered: nearest-child hints in the blob's copy of the bvh_node words (the caller's
    // description is not touched) and the nesting depth of the bvh_nodes = the most children the walk defers
    L.skip_ordered = 0;
    if (simple && L.skip_cull_abs >= 0.f) {
      int depth = 0, max_depth = 0;
      std::vector<int> ends;
      for (int i = 0; i < n_root; i++) {
        while (!ends.empty() && i >= ends.back()) { ends.pop_back(); depth--; }
        const tpt_node &nd = d->nodes[i];
        if ((nd.kind & 0xff) != TPT_NODE_BVH) {
          if ((nd.kind & 0xff) == TPT_NODE_LIST) i = nd.end_or_prim - 1; // nothing below a list is deferred
          continue;
        }
        ends.push_back(nd.end_or_prim);
        max_depth = std::max(max_depth, ++depth);
        if (nd.kind & TPT_NODE_DUP) continue;
        const int l = i + 1;
        if (l >= nd.end_or_prim) continue;
        const tpt_node &ln = d->nodes[l];
        const int r = (ln.kind & 0xff) == TPT_NODE_LEAF ? l + 1 : ln.end_or_prim;
        if (r >= nd.end_or_prim) continue;
        const tpt_node &rn = d->nodes[r];
        if ((rn.kind & TPT_NODE_DUP) || (ln.kind >> 16) != (nd.kind >> 16) || (rn.kind >> 16) != (nd.kind >> 16)) continue;
        int axis = -1;
        float sep = 0.f;
        bool lower = true;
        for (int k = 0; k < 3; k++) {
          const float cl = 0.5f * (ln.bmin[k] + ln.bmax[k]), cr = 0.5f * (rn.bmin[k] + rn.bmax[k]);
          if (!std::isfinite(cl) || !std::isfinite(cr)) { axis = -1; break; }
          if (std::fabs(cl - cr) > sep) { sep = std::fabs(cl - cr); axis = k; lower = cl <= cr; }
        }
        if (axis < 0) continue;
        int32_t kind;
        unsigned char *w = blob.data() + (size_t)L.off_nodes * 16 + (size_t)i * sizeof(tpt_node) + offsetof(tpt_node, kind);
        std::memcpy(&kind, w, 4);
        kind |= 0x200 | (axis << 10) | (lower ? 0x1000 : 0); // TPT_NODE_HINT, axis, TPT_NODE_HINT_LOWER (tpt_device.cuh)
        std::memcpy(w, &kind, 4);
      }
      L.skip_ordered = max_depth + 2 <= 32 ? 1 : 0; // TPT_SKIP_STACK
      if (const char *e = std::getenv("TPT_PARITY_SKIP_ORDERED"))
        if (e[0] == '0') L.skip_ordered = 0;
    }
  }
  s->fbvh_has_moving = fb.has_moving;
  s->fbvh_t0 = fb.moving_t0;
  s->fbvh_t1 = fb.moving_t1;
  L.blob_words = words();
  L.n_nodes = n_root; // world->hit walks the root tree only; boundaries are reached through their medium
  L.n_prims = d->n_prims;
  L.n_lights = d->n_lights;
  L.background = d->background;
  s->has_lights = d->n_lights > 0;
  s->tex_mask = 0; // TPT_NO_LEAN=1 in the environment keeps the full kernels (A/B measurements)
  for (int i = 0; i < d->n_textures; i++) {
    if (d->textures[i].kind == TPT_TEX_IMAGE) s->tex_mask |= TPT_TEXF_IMAGE;
    if (d->textures[i].kind == TPT_TEX_CHECKER || d->textures[i].kind == TPT_TEX_PERLIN) s->tex_mask |= TPT_TEXF_PROCEDURAL;
  }
  if (const char *e = std::getenv("TPT_NO_LEAN"))
    if (e[0] == '1') s->tex_mask = TPT_TEXF_ALL;
  s->blob_bytes = blob.size();
  CK(cudaMallocAsync((void **)&s->d_blob, blob.size(), s->stream));
  CK(cudaMemcpyAsync(s->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  L.blob_global = s->d_blob;
  s->use_smem = blob.size() <= 64 * 1024;
  build_small_scene(d, n_root, s->use_smem, s->small);
  build_flat_tree(d, n_root, s->use_smem && mediums.empty(), s->flat);

  // ---- image textures: RGB -> RGBA8 cudaArray, point sampling, clamp, unnormalised coords ----
  for (int i = 0; i < d->n_images; i++) {
    const tpt_image_desc &im = d->images[i];
    std::vector<uchar4> rgba((size_t)im.width * im.height);
    for (size_t k = 0; k < rgba.size(); k++)
      rgba[k] = make_uchar4(im.rgb[3 * k], im.rgb[3 * k + 1], im.rgb[3 * k + 2], 255);
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr;
    CK(cudaMallocArray(&arr, &fmt, im.width, im.height));
    CK(cudaMemcpy2DToArray(arr, 0, 0, rgba.data(), (size_t)im.width * 4, (size_t)im.width * 4, im.height,
                           cudaMemcpyHostToDevice));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    s->arrays.push_back(arr);
    s->textures.push_back(tex);
    L.images[i] = tex;
    L.image_w[i] = im.width;
    L.image_h[i] = im.height;
  }
  CK(cudaMallocAsync((void **)&s->d_counters, TPT_MAX_BATCHES * 8 * sizeof(unsigned long long), s->stream));
  s->stats.h2d_bytes = blob.size();
  *out = s;
  *partial = nullptr;
  return TPT_OK;
}

void tpt_scene_destroy(tpt_scene *s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (auto t : s->textures) cudaDestroyTextureObject(t);
  for (auto a : s->arrays) cudaFreeArray(a);
  if (s->stream) {
    if (s->d_acc) cudaFreeAsync(s->d_acc, s->stream);
    if (s->d_sum) cudaFreeAsync(s->d_sum, s->stream);
    if (s->d_rgb8) cudaFreeAsync(s->d_rgb8, s->stream);
    if (s->d_rgb8_slices) cudaFreeAsync(s->d_rgb8_slices, s->stream);
    if (s->d_blob) cudaFreeAsync(s->d_blob, s->stream);
    if (s->d_counters) cudaFreeAsync(s->d_counters, s->stream);
    cudaStreamSynchronize(s->stream);
  }
  for (auto ev : s->ev)
    if (ev) cudaEventDestroy(ev);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

int tpt_intersect_batch(const tpt_scene *cs, const tpt_ray *rays, size_t n, float tmin, float tmax, int mode,
                        tpt_hit *out) {
  tpt_scene *s = const_cast<tpt_scene *>(cs);
  if (!s || (!rays && n) || (!out && n)) return fail(TPT_ERR_INVALID, "null argument");
  if (mode != TPT_MODE_PARITY && mode != TPT_MODE_FAST) return fail(TPT_ERR_INVALID, "unknown mode");
  if (n == 0) return TPT_OK;
  if (s->n_mediums > 0)
    return fail(TPT_ERR_UNSUPPORTED, "constant_medium::hit draws random numbers: a scene with media has no deterministic hit batch");
  static_assert(sizeof(tpt_ray) == 28, "tpt_ray layout");
  CK(cudaSetDevice(s->device));
  float *d_rays = nullptr;
  tpt_hit *d_out = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_rays, n * sizeof(tpt_ray));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, n * sizeof(tpt_hit));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rays, rays, n * sizeof(tpt_ray), cudaMemcpyHostToDevice, s->stream);
  if (e != cudaSuccess) { // nothing allocated so far outlives a failed call
    cudaFree(d_rays);
    cudaFree(d_out);
    return cuda_fail(e, "intersect batch: device buffers");
  }
  IntersectArgs A;
  A.scene = s->layout;
  if (mode == TPT_MODE_PARITY) A.flat = s->flat;
  else A.small = s->small;
  A.rays = d_rays;
  A.n = n;
  A.tmin = tmin;
  A.tmax = tmax;
  A.out = d_out;
  e = mode == TPT_MODE_PARITY ? launch_intersect_parity(A, s->use_smem, s->flat.enabled != 0, s->stream)
                                          : launch_intersect_fast(A, s->use_smem, s->small.enabled != 0, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * sizeof(tpt_hit), cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(d_rays);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "intersect batch");
  return TPT_OK;
}

int tpt_render_device(tpt_scene *s, const tpt_camera *cam, const tpt_render_params *p) {
  Plan plan;
  int rc = make_plan(s, cam, p, plan);
  if (rc != TPT_OK) return rc;
  double t0 = now_ms();
  rc = run_plan(s, plan, true, true, true);
  if (rc == TPT_OK) s->stats.wall_ms = now_ms() - t0;
  return rc;
}

int tpt_render_fetch(tpt_scene *s, tpt_image *out) {
  if (!s) return fail(TPT_ERR_INVALID, "null scene");
  return fetch(s, out);
}

int tpt_render(tpt_scene *s, const tpt_camera *cam, const tpt_render_params *p, tpt_image *out) {
  Plan plan;
  int rc = make_plan(s, cam, p, plan);
  if (rc != TPT_OK) return rc;
  double t0 = now_ms();
  rc = run_plan(s, plan, out && out->sum_rgb, out && out->rgb8, out && out->rgb8_slices);
  if (rc != TPT_OK) return rc;
  rc = fetch(s, out);
  s->stats.wall_ms = now_ms() - t0;
  return rc;
}

int tpt_render_multi(tpt_scene *const *scenes, int n_scenes, const tpt_camera *cam, const tpt_render_params *params,
                     tpt_image *out) {
  return render_multi(scenes, n_scenes, cam, params, out);
}

int tpt_device_buffers(const tpt_scene *s, void **sum_rgb, size_t *sum_bytes, void **rgb8, size_t *rgb8_bytes) {
  if (!s) return fail(TPT_ERR_INVALID, "null scene");
  size_t npix = (size_t)s->last_nx * s->last_ny;
  if (npix == 0) return fail(TPT_ERR_INVALID, "nothing rendered yet");
  if (sum_rgb) *sum_rgb = s->d_sum;
  if (sum_bytes) *sum_bytes = (size_t)s->last_slices * npix * 3 * sizeof(float);
  if (rgb8) *rgb8 = s->d_rgb8;
  if (rgb8_bytes) *rgb8_bytes = npix * 3;
  return TPT_OK;
}

int tpt_get_stats(const tpt_scene *s, tpt_stats *out) {
  if (!s || !out) return fail(TPT_ERR_INVALID, "null argument");
  *out = s->stats;
  return TPT_OK;
}

int tpt_debug_philox(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TPT_ERR_NO_DEVICE, "no CUDA device available");
  }
  CK(cudaSetDevice(device));
  uint32_t *d = nullptr;
  CK(cudaMalloc((void **)&d, 16));
  cudaError_t e = launch_philox_probe(ctr, key, d, 0);
  if (e == cudaSuccess) e = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "philox probe");
  return TPT_OK;
}

int tpt_debug_fp32_peak(int device, double *tflops, double *ms) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TPT_ERR_NO_DEVICE, "no CUDA device available");
  }
  if (!tflops) return fail(TPT_ERR_INVALID, "null argument");
  CK(cudaSetDevice(device));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, iters = 1 << 16;
  float *sink = nullptr;
  cudaEvent_t e0, e1;
  cudaEvent_t none = nullptr;
  e0 = e1 = none;
  cudaError_t e = cudaMalloc((void **)&sink, 4);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e != cudaSuccess) {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(sink);
    return cuda_fail(e, "fp32 peak probe: setup");
  }
  float best = 0.f;
  for (int rep = 0; rep < 4 && e == cudaSuccess; rep++) { // first pass warms the clocks up
    cudaEventRecord(e0, 0);
    e = launch_fp32_peak_probe(blocks, iters, sink, 0);
    cudaEventRecord(e1, 0);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float t = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, e0, e1);
    if (rep > 0 && (best == 0.f || t < best)) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) return cuda_fail(e, "fp32 peak probe");
  const double flop = (double)blocks * 256.0 * (double)iters * 64.0 * 2.0;
  *tflops = flop / ((double)best * 1e-3) / 1e12;
  if (ms) *ms = best;
  return TPT_OK;
}

int tpt_debug_small_scene(const tpt_scene_desc *d, int32_t out[64]) {
  int depth = 0;
  int rc = validate_desc(d, depth);
  if (rc != TPT_OK) return rc;
  if (!out) return fail(TPT_ERR_INVALID, "null argument");
  const int n_root = d->n_root_nodes > 0 ? d->n_root_nodes : d->n_nodes;
  SmallScene Q;
  build_small_scene(d, n_root, true, Q);
  std::memset(out, 0, 64 * sizeof(int32_t));
  out[0] = Q.enabled;
  out[1] = Q.n_groups;
  int nbox = 0, nrect = 0, nsph = 0;
  for (int g = 0; g < Q.n_groups; g++) {
    const SmallGroup &G = Q.groups[g];
    nrect += G.yz_end - G.begin;
    nsph += G.sph_end - G.yz_end;
    nbox = std::max(nbox, G.box_end);
  }
  out[2] = nrect;
  out[3] = nsph;
  out[4] = nbox;
  for (int b = 0; b < nbox && b < TPT_SMALL_MAX_BOXES; b++)
    for (int f = 0; f < 8; f++) out[8 + 8 * b + f] = Q.boxes[b].face[f];
  return TPT_OK;
}

int tpt_debug_texture(const tpt_scene *cs, int texture, const float *uvp, size_t n, int mode, float *out_rgb) {
  tpt_scene *s = const_cast<tpt_scene *>(cs);
  if (!s || !uvp || !out_rgb) return fail(TPT_ERR_INVALID, "null argument");
  if (n == 0) return TPT_OK;
  CK(cudaSetDevice(s->device));
  float *d_in = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_in, n * 5 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, n * 3 * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(d_in, uvp, n * 5 * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d_in);
    cudaFree(d_out);
    return cuda_fail(e, "texture probe: device buffers");
  }
  TextureProbeArgs A;
  A.scene = s->layout;
  A.texture = texture;
  A.uvp = d_in;
  A.n = n;
  A.out = d_out;
  e = mode == TPT_MODE_PARITY ? launch_texture_probe_parity(A, s->stream)
                                          : launch_texture_probe_fast(A, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  if (e == cudaSuccess) e = cudaMemcpy(out_rgb, d_out, n * 3 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "texture probe");
  return TPT_OK;
}

} // extern "C"
