// oracle/flatten_ref.cc -- the REFERENCE-SIDE BINDING of include/tpt.h, compiled for real.
//
// This is the file INTEGRATION.md section 2 tells a maintainer of BlurryLight/tiny-path-tracer to
// add: it is built against the UNMODIFIED reference headers (/root/reference/headers/*.h) and linked
// with the reference's own objects (src/*.cc) and with tiny-path-tracer_b200/lib/libtpt.so. It walks
// a `hitable*` built by the reference's own scene builders (src/utils.cc:96-414) with dynamic_cast
// over the public members the reference exposes (headers/hitable.h:51-53, hitable_list.h:6-7,
// rect_box.h:59-116, sphere.h, material.h, texture.h), fills a tpt_scene_desc, and replaces the block
// main.cpp:109-175 with tpt_scene_create + tpt_render. Nothing here computes a pixel.
//
// It lives under oracle/ because it can only be compiled where /root/reference exists; the tests
// (tests/test_ref_binding.py) use it to prove that libtpt.so is a drop-in for the REAL classes:
// the description it produces equals the one the repo's own front end (host/tpt_flatten.cc over the
// host/tpt_scene.h classes) produces, byte for byte, and the pictures rendered through it are
// bit-identical.
//
// Flattening rules (the contract of include/tpt.h, restated from host/tpt_flatten.h):
//   pre-order nodes (bvh_node: left then right; hitable_list: in order; box: its list);
//   a one-element bvh_node (left_ == right_, src/hitable.cc:41-42) emits its child twice, the
//   second copy flagged TPT_NODE_DUP; flip_normal toggles a flag; translate / rotate_y push one op
//   on the chain; prim / material / texture ids in order of first encounter; a hitable_list's box =
//   the union of its children's node boxes when they share its chain, else unbounded.
#include "camera.h"
#include "hitable.h"
#include "hitable_list.h"
#include "material.h"
#include "perlin_noise.h"
#include "rect_box.h"
#include "sphere.h"
#include "texture.h"
#include "utils.h"

#include "../tiny-path-tracer_b200/host/tpt_scene_programs.h" // test scene family, compiled here against the reference's classes

#include "tpt.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <typeinfo>
#include <vector>

namespace {

struct Desc {
  std::vector<tpt_node> nodes;
  std::vector<tpt_prim> prims;
  std::vector<tpt_chain> chains;
  std::vector<tpt_xform_op> ops;
  std::vector<tpt_material> materials;
  std::vector<tpt_texture> textures;
  std::vector<tpt_image_desc> images;
  std::vector<tpt_light> lights;
  tpt_perlin_tables perlin;
  bool has_perlin = false;
  std::string error;
};

class Walker {
public:
  explicit Walker(Desc &d) : d_(d) {
    d_.chains.push_back(tpt_chain{0, 0});
    chain_ids_[std::vector<int32_t>()] = 0;
  }

  void visit(hitable *h) {
    if (!d_.error.empty()) return;
    if (auto *n = dynamic_cast<bvh_node *>(h)) {
      const int me = open(TPT_NODE_BVH, n->box_.min_, n->box_.max_);
      visit(n->left_);
      if (n->right_ == n->left_) dup_next_ = true;
      visit(n->right_);
      close(me);
    } else if (auto *l = dynamic_cast<hitable_list *>(h)) {
      const float m = std::numeric_limits<float>::max();
      const int me = open(TPT_NODE_LIST, vec3(-m, -m, -m), vec3(m, m, m));
      for (int i = 0; i < l->list_size_; i++) visit(l->list_[i]);
      close(me);
    } else if (auto *b = dynamic_cast<box *>(h)) {
      visit(b->list_ptr_); // src/rect_box.cc:117-119
    } else if (auto *f = dynamic_cast<flip_normal *>(h)) {
      flip_ = !flip_;
      visit(f->ptr_);
      flip_ = !flip_;
    } else if (auto *t = dynamic_cast<translate *>(h)) {
      stack_.push_back(tpt_xform_op{TPT_XF_TRANSLATE, t->offset_.x(), t->offset_.y(), t->offset_.z()});
      visit(t->ptr_);
      stack_.pop_back();
    } else if (auto *r = dynamic_cast<rotate_y *>(h)) {
      stack_.push_back(tpt_xform_op{TPT_XF_ROTATE_Y, r->sin_theta_, r->cos_theta_, 0.0f});
      visit(r->ptr_);
      stack_.pop_back();
    } else if (auto *s = dynamic_cast<sphere *>(h)) {
      const float p[4] = {s->center_.x(), s->center_.y(), s->center_.z(), s->radius_};
      leaf(h, TPT_PRIM_SPHERE, p, 4, s->mat_ptr_, 0, 0);
    } else if (auto *ms = dynamic_cast<moving_sphere *>(h)) {
      const float p[9] = {ms->center0_.x(), ms->center0_.y(), ms->center0_.z(), ms->radius_, ms->center1_.x(),
                          ms->center1_.y(), ms->center1_.z(), ms->time0_, ms->time1_};
      leaf(h, TPT_PRIM_MOVING_SPHERE, p, 9, ms->mat_ptr_, ms->time0_, ms->time1_);
    } else if (auto *xy = dynamic_cast<xy_rect *>(h)) {
      const float p[5] = {xy->x0_, xy->x1_, xy->y0_, xy->y1_, xy->k_};
      leaf(h, TPT_PRIM_XY_RECT, p, 5, xy->mat_ptr_, 0, 0);
    } else if (auto *xz = dynamic_cast<xz_rect *>(h)) {
      const float p[5] = {xz->x0_, xz->x1_, xz->z0_, xz->z1_, xz->k_};
      leaf(h, TPT_PRIM_XZ_RECT, p, 5, xz->mat_ptr_, 0, 0);
    } else if (auto *yz = dynamic_cast<yz_rect *>(h)) {
      const float p[5] = {yz->y0_, yz->y1_, yz->z0_, yz->z1_, yz->k_};
      leaf(h, TPT_PRIM_YZ_RECT, p, 5, yz->mat_ptr_, 0, 0);
    } else {
      // constant_medium is reachable the same way (boundary_ / density_ / phase_funcion_ are public,
      // headers/hitable.h:66-68); this binding covers the scenes main.cpp can select at HEAD
      d_.error = std::string("hitable class outside this binding: ") + typeid(*h).name();
    }
  }

private:
  int chain() {
    std::vector<int32_t> key;
    for (const tpt_xform_op &op : stack_) {
      int32_t raw[4];
      std::memcpy(raw, &op, sizeof(raw));
      key.insert(key.end(), raw, raw + 4);
    }
    auto it = chain_ids_.find(key);
    if (it != chain_ids_.end()) return it->second;
    tpt_chain c{(int32_t)d_.ops.size(), (int32_t)stack_.size()};
    for (const tpt_xform_op &op : stack_) d_.ops.push_back(op);
    d_.chains.push_back(c);
    return chain_ids_[key] = (int)d_.chains.size() - 1;
  }
  int open(int kind, const vec3 &lo, const vec3 &hi) {
    tpt_node n;
    std::memset(&n, 0, sizeof(n));
    for (int c = 0; c < 3; c++) {
      n.bmin[c] = lo[c];
      n.bmax[c] = hi[c];
    }
    n.kind = kind | (dup_next_ ? TPT_NODE_DUP : 0) | (chain() << 16);
    dup_next_ = false;
    n.end_or_prim = -1;
    d_.nodes.push_back(n);
    return (int)d_.nodes.size() - 1;
  }
  void close(int me) {
    tpt_node &g = d_.nodes[me];
    g.end_or_prim = (int32_t)d_.nodes.size();
    if ((g.kind & 0xff) != TPT_NODE_LIST) return;
    const float m = std::numeric_limits<float>::max();
    float lo[3] = {m, m, m}, hi[3] = {-m, -m, -m};
    bool bounded = g.end_or_prim > me + 1;
    for (int i = me + 1; bounded && i < g.end_or_prim;) {
      const tpt_node &c = d_.nodes[i];
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
  void leaf(hitable *self, int kind, const float *params, int n, material *mat, float t0, float t1) {
    const int ch = chain();
    auto key = std::make_pair((const void *)self, std::make_pair(ch, flip_ ? 1 : 0));
    int id;
    auto it = prim_ids_.find(key);
    if (it != prim_ids_.end()) {
      id = it->second;
    } else {
      tpt_prim p;
      std::memset(&p, 0, sizeof(p));
      p.kind = kind;
      p.material = material_id(mat);
      p.chain = ch;
      p.flags = flip_ ? TPT_PRIM_FLIP : 0;
      for (int i = 0; i < n; i++) p.p[i] = params[i];
      d_.prims.push_back(p);
      id = prim_ids_[key] = (int)d_.prims.size() - 1;
    }
    AABB b;
    const float m = std::numeric_limits<float>::max();
    if (!self->bounding_box(t0, t1, b)) b = AABB(vec3(-m, -m, -m), vec3(m, m, m));
    tpt_node nd;
    std::memset(&nd, 0, sizeof(nd));
    for (int c = 0; c < 3; c++) {
      nd.bmin[c] = b.min_[c];
      nd.bmax[c] = b.max_[c];
    }
    nd.kind = TPT_NODE_LEAF | (dup_next_ ? TPT_NODE_DUP : 0) | (ch << 16);
    dup_next_ = false;
    nd.end_or_prim = id;
    d_.nodes.push_back(nd);
  }
  int texture_id(texture *t) {
    auto it = tex_ids_.find(t);
    if (it != tex_ids_.end()) return it->second;
    tpt_texture o;
    std::memset(&o, 0, sizeof(o));
    o.odd = o.even = o.image = -1;
    if (auto *c = dynamic_cast<constant_texture *>(t)) {
      o.kind = TPT_TEX_CONSTANT;
      o.color[0] = c->color_.r();
      o.color[1] = c->color_.g();
      o.color[2] = c->color_.b();
    } else if (auto *k = dynamic_cast<checker_texture *>(t)) {
      o.kind = TPT_TEX_CHECKER;
      o.odd = texture_id(k->odd_);
      o.even = texture_id(k->even_);
    } else if (auto *p = dynamic_cast<perlin_noise_texture *>(t)) {
      o.kind = TPT_TEX_PERLIN;
      o.scale = p->scale_;
      d_.has_perlin = true;
    } else if (auto *im = dynamic_cast<image_texture *>(t)) {
      o.kind = TPT_TEX_IMAGE;
      auto f = image_ids_.find(im->data_);
      if (f == image_ids_.end()) {
        d_.images.push_back(tpt_image_desc{im->data_, im->width_, im->height_});
        f = image_ids_.insert({im->data_, (int)d_.images.size() - 1}).first;
      }
      o.image = f->second;
    } else {
      d_.error = "texture class outside this binding";
    }
    d_.textures.push_back(o);
    return tex_ids_[t] = (int)d_.textures.size() - 1;
  }
  int material_id(material *m) {
    auto it = mat_ids_.find(m);
    if (it != mat_ids_.end()) return it->second;
    tpt_material o;
    std::memset(&o, 0, sizeof(o));
    o.kind = TPT_MAT_ABSORBER; // a null material or the base class: scatter() false, emitted() 0
    o.texture = -1;
    if (auto *l = dynamic_cast<lambertian *>(m)) {
      o.kind = TPT_MAT_LAMBERTIAN;
      o.texture = texture_id(l->albedo_);
    } else if (auto *me = dynamic_cast<metal *>(m)) {
      o.kind = TPT_MAT_METAL;
      o.albedo[0] = me->albedo_.r();
      o.albedo[1] = me->albedo_.g();
      o.albedo[2] = me->albedo_.b();
      o.fuzz = me->fuzz_;
    } else if (auto *di = dynamic_cast<dielectric *>(m)) {
      o.kind = TPT_MAT_DIELECTRIC;
      o.ref_idx = di->ref_idx_;
    } else if (auto *dl = dynamic_cast<diffuse_light *>(m)) {
      o.kind = TPT_MAT_DIFFUSE_LIGHT;
      o.texture = texture_id(dl->emit_);
    } else if (auto *iso = dynamic_cast<isotropic *>(m)) {
      o.kind = TPT_MAT_ISOTROPIC;
      o.texture = texture_id(iso->albedo_);
    }
    d_.materials.push_back(o);
    return mat_ids_[m] = (int)d_.materials.size() - 1;
  }

  Desc &d_;
  std::vector<tpt_xform_op> stack_;
  std::map<std::vector<int32_t>, int> chain_ids_;
  std::map<std::pair<const void *, std::pair<int, int>>, int> prim_ids_;
  std::map<material *, int> mat_ids_;
  std::map<texture *, int> tex_ids_;
  std::map<unsigned char *, int> image_ids_;
  bool flip_ = false, dup_next_ = false;
};

// the light-sampling list handed to color(): main.cpp:99-106
void describe_lights(hitable_list &hlist, Desc &d) {
  for (int i = 0; i < hlist.list_size_; i++) {
    tpt_light L;
    std::memset(&L, 0, sizeof(L));
    if (auto *r = dynamic_cast<xz_rect *>(hlist.list_[i])) {
      L.kind = TPT_LIGHT_XZ_RECT;
      L.p[0] = r->x0_;
      L.p[1] = r->x1_;
      L.p[2] = r->z0_;
      L.p[3] = r->z1_;
      L.p[4] = r->k_;
    } else if (auto *s = dynamic_cast<sphere *>(hlist.list_[i])) {
      L.kind = TPT_LIGHT_SPHERE;
      L.p[0] = s->center_.x();
      L.p[1] = s->center_.y();
      L.p[2] = s->center_.z();
      L.p[3] = s->radius_;
    } else {
      L.kind = TPT_LIGHT_OTHER;
    }
    d.lights.push_back(L);
  }
}

tpt_scene_desc as_desc(Desc &d) {
  tpt_scene_desc s;
  std::memset(&s, 0, sizeof(s));
  s.api_version = TPT_API_VERSION;
  s.n_nodes = (int32_t)d.nodes.size();
  s.n_prims = (int32_t)d.prims.size();
  s.n_chains = (int32_t)d.chains.size();
  s.n_xform_ops = (int32_t)d.ops.size();
  s.n_materials = (int32_t)d.materials.size();
  s.n_textures = (int32_t)d.textures.size();
  s.n_images = (int32_t)d.images.size();
  s.n_lights = (int32_t)d.lights.size();
  s.nodes = d.nodes.data();
  s.prims = d.prims.data();
  s.chains = d.chains.data();
  s.xform_ops = d.ops.data();
  s.materials = d.materials.data();
  s.textures = d.textures.data();
  s.images = d.images.data();
  s.lights = d.lights.data();
  if (d.has_perlin) { // the live static tables, AFTER scene construction (src/perlin_noise.cc:3-21)
    for (int i = 0; i < 256; i++) {
      for (int c = 0; c < 3; c++) d.perlin.ranvec[i][c] = perlin_noise::random_vec3_[i][c];
      d.perlin.perm_x[i] = perlin_noise::permute_x_[i];
      d.perlin.perm_y[i] = perlin_noise::permute_y_[i];
      d.perlin.perm_z[i] = perlin_noise::permute_z_[i];
    }
    s.perlin = &d.perlin;
  }
  s.background = TPT_BG_BLACK; // src/utils.cc:85-86 at HEAD
  s.n_root_nodes = 0;
  return s;
}

// scene selection as main.cpp:71-81 offers it (one line active, the others commented out)
hitable *build_on_this_thread(const std::string &name, unsigned char *img, int iw, int ih);
// built on a FRESH thread: drand_r's thread_local mt19937 (src/utils.cc:29) then starts at its default
// seed, exactly as on the main thread of a fresh Path_tracer process (bvh_node's axis choice and
// random_scene() draw from it)
hitable *build_reference_scene(const std::string &name, unsigned char *img, int iw, int ih) {
  hitable *world = nullptr;
  std::thread t([&] { world = build_on_this_thread(name, img, iw, ih); });
  t.join();
  return world;
}
hitable *build_on_this_thread(const std::string &name, unsigned char *img, int iw, int ih) {
  if (name == "cornell_box") return cornell_box();
  if (name == "sphere_cornell_box") return sphere_cornell_box();
  if (name == "random_scene") return random_scene();
  if (name == "two_perlin_spheres") return two_perlin_spheres();
  if (name == "light_spheres") return light_spheres();
  if (name == "earth" && img) return new sphere(vec3(0, 0, 0), 3, new lambertian(new image_texture(img, iw, ih))); // main.cpp:78-81
  // random programs over the reference's classes (test scene family)
  if (name.rfind("programL:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 9, nullptr, 10), false, true);
  if (name.rfind("program:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 8, nullptr, 10));
  return nullptr;
}

std::string g_error;

} // namespace

extern "C" {

const char *tptbind_last_error(void) { return g_error.c_str(); }

// Flatten the reference-built scene `name` and copy the description's tables out, back to back, for a
// byte-wise comparison with the repo's own front end: nodes | prims | chains | ops | materials |
// textures | lights. Returns the number of bytes (<= cap), or -1. counts[8] = the eight n_* fields.
long tptbind_describe(const char *name, unsigned char *image, int iw, int ih, unsigned char *out, long cap, int32_t counts[8]) {
  hitable *world = build_reference_scene(name, image, iw, ih);
  if (!world) {
    g_error = "unknown scene";
    return -1;
  }
  Desc d;
  Walker(d).visit(world);
  if (!d.error.empty()) {
    g_error = d.error;
    return -1;
  }
  xz_rect light_shape(-100, 100, -150, -50, 298, nullptr); // main.cpp:99-106
  sphere sphere_shape(vec3(120, -50, 40), 120, nullptr);
  hitable *a[2] = {&light_shape, &sphere_shape};
  hitable_list hlist(a, 2);
  describe_lights(hlist, d);
  tpt_scene_desc s = as_desc(d);
  const int32_t n[8] = {s.n_nodes, s.n_prims, s.n_chains, s.n_xform_ops, s.n_materials, s.n_textures, s.n_images, s.n_lights};
  std::memcpy(counts, n, sizeof(n));
  std::vector<unsigned char> blob;
  auto put = [&blob](const void *p, size_t bytes) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    blob.insert(blob.end(), b, b + bytes);
  };
  put(s.nodes, sizeof(tpt_node) * s.n_nodes);
  put(s.prims, sizeof(tpt_prim) * s.n_prims);
  put(s.chains, sizeof(tpt_chain) * s.n_chains);
  put(s.xform_ops, sizeof(tpt_xform_op) * s.n_xform_ops);
  put(s.materials, sizeof(tpt_material) * s.n_materials);
  put(s.textures, sizeof(tpt_texture) * s.n_textures);
  put(s.lights, sizeof(tpt_light) * s.n_lights);
  if ((long)blob.size() > cap) {
    g_error = "buffer too small";
    return -1;
  }
  std::memcpy(out, blob.data(), blob.size());
  return (long)blob.size();
}

// What main() does from main.cpp:71 to :175 with the sample loop replaced:
//   world = <builder>() ; camera cam(...) ; hlist = {xz_rect, sphere} ; [flatten ; tpt_scene_create ; tpt_render]
int tptbind_render(const char *name, unsigned char *image, int iw, int ih, const float lookfrom[3], const float lookat[3],
                   float vfov, float aperture, float focus_dist, float t0, float t1, const tpt_render_params *rp,
                   float *sum_rgb, uint8_t *rgb8, tpt_stats *stats) {
  hitable *world = build_reference_scene(name, image, iw, ih);
  if (!world) {
    g_error = "unknown scene";
    return TPT_ERR_INVALID;
  }
  camera cam(vec3(lookfrom[0], lookfrom[1], lookfrom[2]), vec3(lookat[0], lookat[1], lookat[2]), vec3(0, 1, 0), vfov,
             float(rp->nx) / (float)rp->ny, aperture, focus_dist, t0, t1); // main.cpp:90-91
  xz_rect light_shape(-100, 100, -150, -50, 298, nullptr);
  sphere sphere_shape(vec3(120, -50, 40), 120, nullptr);
  hitable *a[2] = {&light_shape, &sphere_shape};
  hitable_list hlist(a, 2);

  Desc d;
  Walker(d).visit(world);
  if (!d.error.empty()) {
    g_error = d.error;
    return TPT_ERR_UNSUPPORTED;
  }
  describe_lights(hlist, d);
  tpt_scene_desc desc = as_desc(d);
  tpt_camera c; // the public fields the ctor filled in (headers/camera.h:14-20)
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

  tpt_scene *scene = nullptr;
  int rc = tpt_scene_create(&desc, rp->device, &scene);
  if (rc != TPT_OK) {
    g_error = tpt_last_error();
    return rc;
  }
  tpt_image img;
  img.sum_rgb = sum_rgb;
  img.rgb8 = rgb8;
  img.rgb8_slices = nullptr;
  rc = tpt_render(scene, &c, rp, &img);
  if (rc != TPT_OK) g_error = tpt_last_error();
  if (rc == TPT_OK && stats) tpt_get_stats(scene, stats);
  tpt_scene_destroy(scene);
  return rc;
}

} // extern "C"
