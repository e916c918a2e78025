// host/tpt_flatten.cc -- see tpt_flatten.h
#include "tpt_flatten.h"

#include <cstring>
#include <algorithm>
#include <limits>

namespace tpt {

tpt_scene_desc FlatScene::desc() const {
  tpt_scene_desc d;
  std::memset(&d, 0, sizeof(d));
  d.api_version = TPT_API_VERSION;
  d.n_nodes = (int32_t)nodes.size();
  d.n_prims = (int32_t)prims.size();
  d.n_chains = (int32_t)chains.size();
  d.n_xform_ops = (int32_t)xform_ops.size();
  d.n_materials = (int32_t)materials.size();
  d.n_textures = (int32_t)textures.size();
  d.n_images = (int32_t)images.size();
  d.n_lights = (int32_t)lights.size();
  d.nodes = nodes.data();
  d.prims = prims.data();
  d.chains = chains.data();
  d.xform_ops = xform_ops.data();
  d.materials = materials.data();
  d.textures = textures.data();
  d.images = images.data();
  d.perlin = has_perlin ? &perlin : nullptr;
  d.lights = lights.data();
  d.background = background;
  d.n_root_nodes = n_root_nodes;
  return d;
}

Flattener::Flattener(FlatScene &out) : out_(out) {
  out_.chains.push_back(tpt_chain{0, 0}); // chain 0 = identity
  chain_ids_[std::vector<int32_t>()] = 0;
}

void Flattener::fail(const std::string &why) {
  if (error_.empty()) error_ = why;
}

int Flattener::current_chain() {
  std::vector<int32_t> key;
  for (const tpt_xform_op &op : stack_) {
    int32_t raw[4];
    std::memcpy(raw, &op, sizeof(raw));
    key.insert(key.end(), raw, raw + 4);
  }
  auto it = chain_ids_.find(key);
  if (it != chain_ids_.end()) return it->second;
  tpt_chain c;
  c.first_op = (int32_t)out_.xform_ops.size();
  c.n_ops = (int32_t)stack_.size();
  for (const tpt_xform_op &op : stack_) out_.xform_ops.push_back(op);
  out_.chains.push_back(c);
  int id = (int)out_.chains.size() - 1;
  chain_ids_[key] = id;
  return id;
}

static void set_bounds(tpt_node &n, const AABB &b) {
  for (int c = 0; c < 3; c++) {
    n.bmin[c] = b.min_[c];
    n.bmax[c] = b.max_[c];
  }
}

int Flattener::begin_group(int kind, const AABB &bounds) {
  tpt_node n;
  std::memset(&n, 0, sizeof(n));
  set_bounds(n, bounds);
  // the chain id of a group rides in bits 16.. of `kind` (BVH nodes under a transform test
  // their box against the transformed ray)
  n.kind = kind | (dup_next_ ? TPT_NODE_DUP : 0) | (current_chain() << 16);
  dup_next_ = false;
  n.end_or_prim = -1;
  out_.nodes.push_back(n);
  if (++depth_ > out_.max_depth) out_.max_depth = depth_;
  return (int)out_.nodes.size() - 1;
}

void Flattener::end_group(int node_index) {
  tpt_node &g = out_.nodes[node_index];
  g.end_or_prim = (int32_t)out_.nodes.size();
  --depth_;
  if ((g.kind & 0xff) == TPT_NODE_LIST) {
    // The reference never box-tests a hitable_list (src/hitable_list.cc:38-51); the FAST walk and the
    // scene-bounds tests do. hitable_list::bounding_box(0, 0) only covers a moving_sphere's t = 0
    // position, so the node's box is rebuilt here as the union of its children's NODE boxes (a
    // moving sphere's leaf box spans its own [time0, time1], a bvh_node carries the box its ctor
    // built). A child stated in another chain's space cannot be united: then the list is unbounded.
    const float m = std::numeric_limits<float>::max();
    float lo[3] = {m, m, m}, hi[3] = {-m, -m, -m};
    bool bounded = g.end_or_prim > node_index + 1;
    for (int i = node_index + 1; bounded && i < g.end_or_prim;) {
      const tpt_node &c = out_.nodes[i];
      if ((c.kind >> 16) != (g.kind >> 16)) bounded = false;
      for (int k = 0; k < 3; k++) {
        lo[k] = std::min(lo[k], c.bmin[k]);
        hi[k] = std::max(hi[k], c.bmax[k]);
      }
      i = (c.kind & 0xff) == TPT_NODE_LEAF ? i + 1 : c.end_or_prim;
    }
    for (int k = 0; k < 3; k++) {
      g.bmin[k] = bounded ? lo[k] : -m;
      g.bmax[k] = bounded ? hi[k] : m;
    }
  }
}

void Flattener::leaf(const hitable *self, int prim_kind, const float *params, int n_params,
                     const material *mat, const AABB &bounds) {
  int chain = current_chain();
  auto key = std::make_pair(self, std::make_pair(chain, flip_ ? 1 : 0));
  int prim_id;
  auto it = prim_ids_.find(key);
  if (it != prim_ids_.end()) {
    prim_id = it->second;
  } else {
    tpt_prim p;
    std::memset(&p, 0, sizeof(p));
    p.kind = prim_kind;
    p.material = material_id(mat);
    p.chain = chain;
    p.flags = flip_ ? TPT_PRIM_FLIP : 0;
    for (int i = 0; i < n_params && i < 12; i++) p.p[i] = params[i];
    out_.prims.push_back(p);
    prim_id = (int)out_.prims.size() - 1;
    prim_ids_[key] = prim_id;
  }
  tpt_node n;
  std::memset(&n, 0, sizeof(n));
  set_bounds(n, bounds);
  n.kind = TPT_NODE_LEAF | (dup_next_ ? TPT_NODE_DUP : 0) | (chain << 16);
  dup_next_ = false;
  n.end_or_prim = prim_id;
  out_.nodes.push_back(n);
}

// constant_medium(boundary, density, texture) (headers/hitable.h:58-69, src/hitable.cc:92-128):
// a MEDIUM primitive in the world tree; its boundary is an ordinary sub-tree kept apart, because
// constant_medium::hit queries it on its own (twice, with its own t ranges).
void Flattener::medium(const hitable *self, const hitable *boundary, float density, const material *phase) {
  // emit the boundary into a scratch node list (same transform chain / flip context)
  std::vector<tpt_node> saved;
  saved.swap(out_.nodes);
  int saved_depth = depth_;
  depth_ = 0;
  boundary->emit(*this);
  depth_ = saved_depth;
  std::vector<tpt_node> sub;
  sub.swap(out_.nodes);
  out_.nodes.swap(saved);
  for (const tpt_node &n : sub)
    if ((n.kind & 0xff) == TPT_NODE_LEAF && out_.prims[n.end_or_prim].kind == TPT_PRIM_MEDIUM) {
      fail("constant_medium inside a constant_medium boundary is not supported");
      return;
    }
  float params[3] = {density, 0.f, 0.f}; // p[1], p[2] patched with the boundary's node range at the end
  AABB b;
  if (!boundary->bounding_box(0, 0, b)) {
    float m = 3.4028234663852886e38f;
    b = AABB(vec3(-m, -m, -m), vec3(m, m, m));
  }
  int before = (int)out_.prims.size();
  leaf(self, TPT_PRIM_MEDIUM, params, 3, phase, b);
  if ((int)out_.prims.size() > before) { // first time this medium is seen
    boundaries_.push_back(sub);
    boundary_prims_.push_back(before);
  }
}

void Flattener::push_xform(const tpt_xform_op &op) { stack_.push_back(op); }
void Flattener::pop_xform() { stack_.pop_back(); }

int Flattener::material_id(const material *m) {
  if (!m) {
    // a leaf without material (the reference's light-list shapes pass nullptr): absorber
    static const material none;
    m = &none;
  }
  auto it = mat_ids_.find(m);
  if (it != mat_ids_.end()) return it->second;
  int id = m->emit(*this);
  mat_ids_[m] = id;
  return id;
}

int Flattener::texture_id(const texture *t) {
  if (!t) {
    fail("material without texture");
    return -1;
  }
  auto it = tex_ids_.find(t);
  if (it != tex_ids_.end()) return it->second;
  int id = t->emit(*this);
  tex_ids_[t] = id;
  return id;
}

int Flattener::add_image(const unsigned char *rgb, int w, int h) {
  if (!rgb || w <= 0 || h <= 0) {
    // the reference would segfault on the first lookup (src/utils.cc:400-401): report instead
    fail("image_texture without pixel data");
    return -1;
  }
  auto it = image_ids_.find(rgb);
  if (it != image_ids_.end()) return it->second;
  out_.image_data.emplace_back(rgb, rgb + (size_t)w * h * 3);
  tpt_image_desc d;
  d.rgb = nullptr; // patched after all images are collected (vector may reallocate)
  d.width = w;
  d.height = h;
  out_.images.push_back(d);
  int id = (int)out_.images.size() - 1;
  image_ids_[rgb] = id;
  return id;
}

bool flatten_scene(const hitable *world, const hitable *light_shape, int background,
                   FlatScene &out, std::string &err) {
  out = FlatScene();
  out.background = background;
  if (!world) {
    err = "null world";
    return false;
  }
  Flattener f(out);
  world->emit(f);
  if (!f.ok()) {
    err = f.error();
    return false;
  }
  // medium boundaries go behind the world tree; group ends become absolute indices again
  out.n_root_nodes = (int)out.nodes.size();
  for (size_t m = 0; m < f.boundaries_.size(); m++) {
    int base = (int)out.nodes.size();
    for (tpt_node n : f.boundaries_[m]) {
      if ((n.kind & 0xff) != TPT_NODE_LEAF) n.end_or_prim += base;
      out.nodes.push_back(n);
    }
    int32_t first = base, end = (int32_t)out.nodes.size();
    tpt_prim &p = out.prims[f.boundary_prims_[m]];
    std::memcpy(&p.p[1], &first, 4);
    std::memcpy(&p.p[2], &end, 4);
  }
  for (size_t i = 0; i < out.images.size(); i++) out.images[i].rgb = out.image_data[i].data();
  if (out.has_perlin) {
    for (int i = 0; i < 256; i++) {
      for (int c = 0; c < 3; c++) out.perlin.ranvec[i][c] = perlin_noise::random_vec3_[i][c];
      out.perlin.perm_x[i] = perlin_noise::permute_x_[i];
      out.perlin.perm_y[i] = perlin_noise::permute_y_[i];
      out.perlin.perm_z[i] = perlin_noise::permute_z_[i];
    }
  }
  // light-sampling shapes (main.cpp:99-106)
  std::vector<const hitable *> shapes;
  if (auto *l = dynamic_cast<const hitable_list *>(light_shape)) {
    for (int i = 0; i < l->list_size_; i++) shapes.push_back(l->list_[i]);
  } else if (light_shape) {
    shapes.push_back(light_shape);
  }
  for (const hitable *s : shapes) {
    tpt_light L;
    std::memset(&L, 0, sizeof(L));
    if (auto *r = dynamic_cast<const xz_rect *>(s)) {
      L.kind = TPT_LIGHT_XZ_RECT;
      L.p[0] = r->x0_;
      L.p[1] = r->x1_;
      L.p[2] = r->z0_;
      L.p[3] = r->z1_;
      L.p[4] = r->k_;
    } else if (auto *sp = dynamic_cast<const sphere *>(s)) {
      L.kind = TPT_LIGHT_SPHERE;
      L.p[0] = sp->center_.x();
      L.p[1] = sp->center_.y();
      L.p[2] = sp->center_.z();
      L.p[3] = sp->radius_;
    } else {
      L.kind = TPT_LIGHT_OTHER; // hitable base class: pdf 0, direction (1,0,0)
    }
    out.lights.push_back(L);
  }
  return true;
}

std::vector<tpt_light> derive_light_list(const FlatScene &flat) {
  std::vector<tpt_light> out;
  for (const tpt_prim &p : flat.prims) {
    if (p.material < 0 || p.material >= (int)flat.materials.size()) continue;
    if (flat.materials[p.material].kind != TPT_MAT_DIFFUSE_LIGHT || p.chain != 0) continue;
    tpt_light L;
    std::memset(&L, 0, sizeof(L));
    if (p.kind == TPT_PRIM_XZ_RECT) {
      L.kind = TPT_LIGHT_XZ_RECT;
      for (int i = 0; i < 5; i++) L.p[i] = p.p[i];
    } else if (p.kind == TPT_PRIM_SPHERE) {
      L.kind = TPT_LIGHT_SPHERE;
      for (int i = 0; i < 4; i++) L.p[i] = p.p[i];
    } else {
      continue;
    }
    out.push_back(L);
  }
  return out;
}

tpt_camera make_camera_desc(const camera_with_blur &cam) {
  tpt_camera c;
  for (int i = 0; i < 3; i++) {
    c.origin[i] = cam.origin_[i];
    c.lower_left_corner[i] = cam.lower_left_corner_[i];
    c.vertical[i] = cam.vertical_[i];
    c.horizontal[i] = cam.horizontal_[i];
    c.u[i] = cam.u_[i];
    c.v[i] = cam.v_[i];
    c.w[i] = cam.w_[i];
  }
  c.lens_radius = cam.lens_radius_;
  c.time0 = cam.time0;
  c.time1 = cam.time1;
  return c;
}

} // namespace tpt
