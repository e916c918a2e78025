// host/tpt_scene.cc -- construction side of the scene classes + their flattening hooks.
// Reference behaviour followed (cited per function); no intersection / shading code here.
#include "tpt_scene.h"
#include "tpt_flatten.h"

#include <chrono>
#include <cstdlib>
#include <limits>
#include <random>

// ------------------------------------------------------------------------------------ RNG
namespace {
std::mt19937 &host_engine() {
  static thread_local std::mt19937 engine; // default seed 5489, like src/utils.cc:29
  return engine;
}
} // namespace

double drand_r(double min, double max) {
  std::uniform_real_distribution<double> dist(min, max); // fresh per call: src/utils.cc:30
  return dist(host_engine());
}
void drand_r_reset() { host_engine() = std::mt19937(); }

// --------------------------------------------------------------------------------- perlin
std::array<vec3, 256> perlin_noise::random_vec3_;
std::array<int, 256> perlin_noise::permute_x_;
std::array<int, 256> perlin_noise::permute_y_;
std::array<int, 256> perlin_noise::permute_z_;

// src/perlin_noise.cc:3-21: 256 unit gradients from drand_r, three permutations shuffled by a
// default_random_engine seeded from the wall clock (so tables differ run to run; the flattener
// uploads whatever is live after scene construction).
perlin_noise::perlin_noise() {
  for (vec3 &g : random_vec3_) {
    float a = static_cast<float>(2 * drand_r() - 1);
    float b = static_cast<float>(2 * drand_r() - 1);
    float c = static_cast<float>(2 * drand_r() - 1);
    g = unit_vector(vec3(a, b, c));
  }
  for (int i = 0; i < 256; i++) permute_x_[i] = permute_y_[i] = permute_z_[i] = i;
  auto seed = std::chrono::high_resolution_clock::now().time_since_epoch().count();
  std::default_random_engine engine(seed);
  std::shuffle(permute_x_.begin(), permute_x_.end(), engine);
  std::shuffle(permute_y_.begin(), permute_y_.end(), engine);
  std::shuffle(permute_z_.begin(), permute_z_.end(), engine);
}

// ------------------------------------------------------------------------ texture flattening
int constant_texture::emit(tpt::Flattener &f) const {
  tpt_texture t{};
  t.kind = TPT_TEX_CONSTANT;
  t.color[0] = color_.r();
  t.color[1] = color_.g();
  t.color[2] = color_.b();
  t.odd = t.even = t.image = -1;
  return f.add_texture(t);
}
int checker_texture::emit(tpt::Flattener &f) const {
  tpt_texture t{};
  t.kind = TPT_TEX_CHECKER;
  t.odd = f.texture_id(odd_);
  t.even = f.texture_id(even_);
  t.image = -1;
  return f.add_texture(t);
}
int perlin_noise_texture::emit(tpt::Flattener &f) const {
  tpt_texture t{};
  t.kind = TPT_TEX_PERLIN;
  t.scale = scale_;
  t.odd = t.even = t.image = -1;
  f.note_perlin();
  return f.add_texture(t);
}
int image_texture::emit(tpt::Flattener &f) const {
  tpt_texture t{};
  t.kind = TPT_TEX_IMAGE;
  t.odd = t.even = -1;
  t.image = f.add_image(data_, width_, height_);
  return f.add_texture(t);
}

// ----------------------------------------------------------------------- material flattening
static tpt_material blank_material(int kind) {
  tpt_material m{};
  m.kind = kind;
  m.texture = -1;
  return m;
}
int material::emit(tpt::Flattener &f) const { return f.add_material(blank_material(TPT_MAT_ABSORBER)); }
int isotropic::emit(tpt::Flattener &f) const {
  tpt_material m = blank_material(TPT_MAT_ISOTROPIC);
  m.texture = f.texture_id(albedo_);
  return f.add_material(m);
}
int lambertian::emit(tpt::Flattener &f) const {
  tpt_material m = blank_material(TPT_MAT_LAMBERTIAN);
  m.texture = f.texture_id(albedo_);
  return f.add_material(m);
}
int metal::emit(tpt::Flattener &f) const {
  tpt_material m = blank_material(TPT_MAT_METAL);
  m.albedo[0] = albedo_.r();
  m.albedo[1] = albedo_.g();
  m.albedo[2] = albedo_.b();
  m.fuzz = fuzz_;
  return f.add_material(m);
}
int dielectric::emit(tpt::Flattener &f) const {
  tpt_material m = blank_material(TPT_MAT_DIELECTRIC);
  m.ref_idx = ref_idx_;
  return f.add_material(m);
}
int diffuse_light::emit(tpt::Flattener &f) const {
  tpt_material m = blank_material(TPT_MAT_DIFFUSE_LIGHT);
  m.texture = f.texture_id(emit_);
  return f.add_material(m);
}

// ------------------------------------------------------------------------- bounding boxes
// rect boxes are padded by 1e-4 along the thin axis (src/rect_box.cc:3-6,45-49,70-73); the
// padding is a double subtraction rounded to float.
bool xy_rect::bounding_box(float, float, AABB &b) const {
  b = AABB(vec3(x0_, y0_, k_ - 0.0001), vec3(x1_, y1_, k_ + 0.0001));
  return true;
}
bool xz_rect::bounding_box(float, float, AABB &b) const {
  b = AABB(vec3(x0_, k_ - 0.0001, z0_), vec3(x1_, k_ + 0.0001, z1_));
  return true;
}
bool yz_rect::bounding_box(float, float, AABB &b) const {
  b = AABB(vec3(k_ - 0.0001, y0_, z0_), vec3(k_ + 0.0001, y1_, z1_));
  return true;
}
bool sphere::bounding_box(float, float, AABB &b) const { // src/sphere.cc:87-91
  vec3 r(radius_, radius_, radius_);
  b = AABB(center_ - r, center_ + r);
  return true;
}
bool moving_sphere::bounding_box(float t0, float t1, AABB &b) const { // src/sphere.cc:77-85
  vec3 r(radius_, radius_, radius_);
  b = surrounding_box(AABB(center(t0) - r, center(t0) + r), AABB(center(t1) - r, center(t1) + r));
  return true;
}
bool hitable_list::bounding_box(float t0, float t1, AABB &b) const { // src/hitable_list.cc:3-22
  if (list_size_ < 1) return false;
  AABB tmp;
  if (!list_[0]->bounding_box(t0, t1, tmp)) return false;
  b = tmp;
  for (int i = 1; i < list_size_; i++) {
    if (!list_[i]->bounding_box(t0, t1, tmp)) return false;
    b = surrounding_box(b, tmp);
  }
  return true;
}
bool bvh_node::bounding_box(float, float, AABB &b) const {
  b = box_;
  return true;
}

// ------------------------------------------------------------------------------ BVH build
namespace {
// ordering used by the reference's build (src/hitable.cc:10-30): a precedes b unless
// a.min[axis] >= b.min[axis]; a leaf without a box aborts the program.
struct by_box_min {
  int axis;
  bool operator()(const hitable *a, const hitable *b) const {
    AABB ba, bb;
    if (!a->bounding_box(0, 0, ba) || !b->bounding_box(0, 0, bb)) {
      std::cout << "No bounding box in BVH node constructor" << std::endl;
      std::exit(-1);
    }
    return !std::isgreaterequal(ba.min()[axis], bb.min()[axis]);
  }
};
} // namespace

// src/hitable.cc:32-56: pick a random axis, sort the range along it, split at the median;
// one element -> both children are that element, two -> one each.
bvh_node::bvh_node(hitable **l, int n, float time0, float time1) {
  int axis = int(drand_r(0, 3.0));
  if (axis > 2) axis = 2; // axis==0 -> x, 1 -> y, anything else -> z (src/hitable.cc:34-40)
  std::sort(l, l + n, by_box_min{axis});
  if (n == 1) {
    left_ = right_ = l[0];
  } else if (n == 2) {
    left_ = l[0];
    right_ = l[1];
  } else {
    left_ = new bvh_node(l, n / 2, time0, time1);
    right_ = new bvh_node(l + n / 2, n - n / 2, time0, time1);
  }
  AABB bl, br;
  if (!left_->bounding_box(time0, time1, bl) || !right_->bounding_box(time0, time1, br))
    std::cout << "no bounding box in bvh constructor" << std::endl;
  box_ = surrounding_box(bl, br);
}

// --------------------------------------------------------------------- box / rotate_y ctors
// six faces, outward normals; order and flip pattern of src/rect_box.cc:93-115
box::box(vec3 pmin, vec3 pmax, material *mat) : point_min_(pmin), point_max_(pmax) {
  hitable **faces = new hitable *[6];
  faces[0] = new xy_rect(pmin.x(), pmax.x(), pmin.y(), pmax.y(), pmax.z(), mat);
  faces[1] = new flip_normal(new xy_rect(pmin.x(), pmax.x(), pmin.y(), pmax.y(), pmin.z(), mat));
  faces[2] = new xz_rect(pmin.x(), pmax.x(), pmin.z(), pmax.z(), pmax.y(), mat);
  faces[3] = new flip_normal(new xz_rect(pmin.x(), pmax.x(), pmin.z(), pmax.z(), pmin.y(), mat));
  faces[4] = new yz_rect(pmin.y(), pmax.y(), pmin.z(), pmax.z(), pmax.x(), mat);
  faces[5] = new flip_normal(new yz_rect(pmin.y(), pmax.y(), pmin.z(), pmax.z(), pmin.x(), mat));
  list_ptr_ = new hitable_list(faces, 6);
}

// src/rect_box.cc:121-169: sin/cos of the angle in float; the rotated box is the sum of the
// per-axis min/max of the rotation-matrix columns scaled by the child's box extents.
rotate_y::rotate_y(hitable *p, float angle) : ptr_(p) {
  float radians = angle / 180.0f * M_PI;
  sin_theta_ = std::sin(radians);
  cos_theta_ = std::cos(radians);
  has_box_ = ptr_->bounding_box(0, 0, box_);
  vec3 col_x(cos_theta_, 0, -sin_theta_), col_y(0, 1, 0), col_z(sin_theta_, 0, cos_theta_);
  vec3 xa = col_x * box_.max().x(), xb = col_x * box_.min().x();
  vec3 ya = col_y * box_.max().y(), yb = col_y * box_.min().y();
  vec3 za = col_z * box_.max().z(), zb = col_z * box_.min().z();
  vec3 lo = vec_min(xa, xb) + vec_min(ya, yb) + vec_min(za, zb);
  vec3 hi = vec_max(xa, xb) + vec_max(ya, yb) + vec_max(za, zb);
  box_ = AABB(lo, hi);
}

// --------------------------------------------------------------------------------- camera
// src/camera.cc:2-21. theta and tan() are evaluated the way the reference's expressions promote:
// vfov*M_PI/180 in double -> float; tan(float) resolves to the double C function.
camera_with_blur::camera_with_blur(vec3 lookfrom, vec3 lookat, vec3 vup, float vfov, float aspect,
                                   float aperture, float focus_dist, float t0, float t1) {
  time0 = t0;
  time1 = t1;
  lens_radius_ = aperture / 2;
  float theta = (double)vfov * M_PI / 180;
  float half_height = ::tan((double)(theta / 2));
  float half_width = aspect * half_height;
  origin_ = lookfrom;
  w_ = unit_vector(lookfrom - lookat);
  u_ = unit_vector(cross(vup, w_));
  v_ = unit_vector(cross(w_, u_));
  lower_left_corner_ =
      origin_ - half_width * focus_dist * u_ - half_height * focus_dist * v_ - w_ * focus_dist;
  horizontal_ = 2 * half_width * u_ * focus_dist;
  vertical_ = 2 * half_height * v_ * focus_dist;
}

// ------------------------------------------------------------------------ hitable flattening
static AABB huge_box() {
  float m = std::numeric_limits<float>::max();
  return AABB(vec3(-m, -m, -m), vec3(m, m, m));
}
static AABB box_or_huge(const hitable *h) {
  AABB b;
  return h->bounding_box(0, 0, b) ? b : huge_box();
}

void hitable_list::emit(tpt::Flattener &f) const {
  int me = f.begin_group(TPT_NODE_LIST, list_size_ > 0 ? box_or_huge(this) : huge_box());
  for (int i = 0; i < list_size_; i++) list_[i]->emit(f);
  f.end_group(me);
}
void bvh_node::emit(tpt::Flattener &f) const {
  int me = f.begin_group(TPT_NODE_BVH, box_);
  left_->emit(f);
  if (right_ == left_) f.mark_next_dup(); // src/hitable.cc:41-42
  right_->emit(f);
  f.end_group(me);
}
void sphere::emit(tpt::Flattener &f) const {
  float p[4] = {center_.x(), center_.y(), center_.z(), radius_};
  f.leaf(this, TPT_PRIM_SPHERE, p, 4, mat_ptr_, box_or_huge(this));
}
void moving_sphere::emit(tpt::Flattener &f) const {
  float p[9] = {center0_.x(), center0_.y(), center0_.z(), radius_, center1_.x(),
                center1_.y(), center1_.z(), time0_,       time1_};
  AABB b;
  bounding_box(time0_, time1_, b);
  f.leaf(this, TPT_PRIM_MOVING_SPHERE, p, 9, mat_ptr_, b);
}
void xy_rect::emit(tpt::Flattener &f) const {
  float p[5] = {x0_, x1_, y0_, y1_, k_};
  f.leaf(this, TPT_PRIM_XY_RECT, p, 5, mat_ptr_, box_or_huge(this));
}
void xz_rect::emit(tpt::Flattener &f) const {
  float p[5] = {x0_, x1_, z0_, z1_, k_};
  f.leaf(this, TPT_PRIM_XZ_RECT, p, 5, mat_ptr_, box_or_huge(this));
}
void yz_rect::emit(tpt::Flattener &f) const {
  float p[5] = {y0_, y1_, z0_, z1_, k_};
  f.leaf(this, TPT_PRIM_YZ_RECT, p, 5, mat_ptr_, box_or_huge(this));
}
void flip_normal::emit(tpt::Flattener &f) const {
  f.toggle_flip();
  ptr_->emit(f);
  f.toggle_flip();
}
void box::emit(tpt::Flattener &f) const { list_ptr_->emit(f); } // src/rect_box.cc:117-119
void translate::emit(tpt::Flattener &f) const {
  tpt_xform_op op{TPT_XF_TRANSLATE, offset_.x(), offset_.y(), offset_.z()};
  f.push_xform(op);
  ptr_->emit(f);
  f.pop_xform();
}
void rotate_y::emit(tpt::Flattener &f) const {
  tpt_xform_op op{TPT_XF_ROTATE_Y, sin_theta_, cos_theta_, 0.0f};
  f.push_xform(op);
  ptr_->emit(f);
  f.pop_xform();
}
void constant_medium::emit(tpt::Flattener &f) const { f.medium(this, boundary_, density_, phase_funcion_); }

// ---------------------------------------------------------------------------- scene builders
namespace {
struct scene_list { // grows a hitable* array the way the builders fill `list[i++]`
  explicit scene_list(int cap) : items(new hitable *[cap]) {}
  void add(hitable *h) { items[n++] = h; }
  hitable **items;
  int n = 0;
};
lambertian *matte(float r, float g, float b) { return new lambertian(new constant_texture(vec3(r, g, b))); }
} // namespace

// src/utils.cc:96-141. The draw order matters (scene must equal the reference's for seed 5489):
// choose_mat, then the centre (g++ evaluates the vec3 ctor arguments right to left, so z before
// x), then the per-material draws inside nested constructor calls. The nested `new` expressions
// are therefore kept as single expressions with the reference's argument structure.
hitable *random_scene() {
  scene_list s(501);
  texture *checker = new checker_texture(new constant_texture({0.1, 0.1, 0.1}),
                                         new constant_texture({0.9, 0.9, 0.9}));
  s.add(new sphere(vec3(0, -1000, 0), 1000, new lambertian(checker)));
  for (int a = -11; a < 11; a++) {
    for (int b = -11; b < 11; b++) {
      float choose_mat = drand_r();
      vec3 center(a + 0.9 * drand_r(), 0.2, b + 0.9 * drand_r());
      if ((center - vec3(4, 0.2, 0)).length() > 0.9) {
        if (choose_mat < 0.8) {
          s.add(new moving_sphere(center, center + vec3(0, 0.5 * drand_r(), 0), 0.0, 1.0, 0.2,
                                  new lambertian(new constant_texture(
                                      vec3(drand_r() * drand_r(), drand_r() * drand_r(),
                                           drand_r() * drand_r())))));
        } else if (choose_mat < 0.95) {
          s.add(new moving_sphere(center, center + vec3(-0.5f + drand_r(), -0.5f + drand_r(), 0.0),
                                  0.0, 1.0, 0.2,
                                  new metal(vec3(0.5 * (1 + drand_r()), 0.5 * (1 + drand_r()),
                                                 0.5 * (1 + drand_r())),
                                            0.5 * drand_r())));
        } else {
          s.add(new sphere(center, 0.2, new dielectric(1.5)));
        }
      }
    }
  }
  s.add(new sphere(vec3(0, 1, 0), 1.0, new dielectric(1.5)));
  s.add(new sphere(vec3(-4, 1, 0), 1.0, matte(0.4, 0.2, 0.1)));
  s.add(new sphere(vec3(4, 1, 0), 1.0, new metal(vec3(0.7, 0.6, 0.5), 0.0)));
  return new bvh_node(s.items, s.n, 0, 1);
}

hitable *two_checker_spheres() { // src/utils.cc:143-153
  texture *checker = new checker_texture(new constant_texture({0.1, 0.1, 0.1}),
                                         new constant_texture({0.9, 0.9, 0.9}));
  scene_list s(3);
  s.add(new sphere(vec3(0, 10, 0), 10, new lambertian(checker)));
  s.add(new sphere(vec3(0, -10, 0), 10, new lambertian(checker)));
  return new hitable_list(s.items, s.n);
}

hitable *two_perlin_spheres() { // src/utils.cc:227-234
  texture *marble = new perlin_noise_texture(2.0f);
  scene_list s(3);
  s.add(new sphere(vec3(0, 2, 0), 2, new lambertian(marble)));
  s.add(new sphere(vec3(0, -1000, 0), 1000, new lambertian(marble)));
  return new hitable_list(s.items, s.n);
}

hitable *light_spheres() { // src/utils.cc:242-255
  texture *glow = new constant_texture(vec3(4, 4, 4));
  scene_list s(10);
  s.add(new sphere(vec3(-1, 1, 0), 1, matte(0.3, 0.4, 0.5)));
  s.add(new sphere(vec3(-3, 1, 2), 1, new diffuse_light(glow)));
  s.add(new sphere(vec3(0, -1000, 0), 1000, new lambertian(new perlin_noise_texture(4.0f))));
  s.add(new sphere(vec3(-3, 1, -2), 1, new diffuse_light(glow)));
  s.add(new xy_rect(3, 5, 1, 3, -2, new diffuse_light(glow)));
  return new hitable_list(s.items, s.n);
}

namespace {
struct cornell_palette {
  material *red = matte(0.65, 0.05, 0.05);
  material *white = matte(0.73, 0.73, 0.73);
  material *green = matte(0.12, 0.45, 0.15);
  material *light = new diffuse_light(new constant_texture(vec3(20, 20, 20)));
};
// the five walls + ceiling lamp shared by cornell_box / cornell_box_smoke (src/utils.cc:297-304)
void cornell_shell(scene_list &s, const cornell_palette &c) {
  s.add(new yz_rect(-300, 300, -300, 300, -300, c.green));
  s.add(new flip_normal(new yz_rect(-300, 300, -300, 300, 300, c.red)));
  s.add(new xz_rect(-300, 300, -300, 300, -300, c.white));
  s.add(new flip_normal(new xz_rect(-300, 300, -300, 300, 300, c.white)));
  s.add(new xy_rect(-300, 300, -300, 300, -300, c.white));
  s.add(new flip_normal(new xz_rect(-100, 100, -150, -50, 298, c.light)));
}
} // namespace

hitable *sphere_cornell_box() { // src/utils.cc:257-285
  cornell_palette c;
  scene_list s(100);
  s.add(new sphere(vec3(-1e5, 0, 0), 1e5 - 300, c.green));
  s.add(new sphere(vec3(1e5, 0, 0), 1e5 - 300, c.red));
  s.add(new sphere(vec3(0, 1e5, 0), 1e5 - 300, c.white));
  s.add(new sphere(vec3(0, -1e5, 0), 1e5 - 300, c.white));
  s.add(new sphere(vec3(0, 0, -1e5), 1e5 - 300, c.white));
  s.add(new xz_rect(-150, 150, -150, 150, 300, c.light));
  s.add(new sphere(vec3(-150, -200, -100), 100, new dielectric(1.5)));
  s.add(new sphere(vec3(150, -200, 100), 100, new metal(vec3(0.7, 0.6, 0.5), 0.0)));
  return new bvh_node(s.items, s.n, 0, 0);
}

hitable *cornell_box() { // src/utils.cc:287-319: the headline scene
  cornell_palette c;
  scene_list s(100);
  cornell_shell(s, c);
  material *aluminum = new metal(vec3(0.8, 0.85, 0.88), 0.0);
  s.add(new translate(new rotate_y(new box(vec3(0, 0, 0), vec3(200, 350, 75), aluminum), 35.0f),
                      vec3(-200, -300, -100)));
  s.add(new translate(new rotate_y(new box(vec3(0, 0, 0), vec3(180, 180, 180), c.white), -25.0f),
                      vec3(30, -300, -50)));
  s.add(new sphere(vec3(120, -50, 40), 70, new dielectric(1.5)));
  return new bvh_node(s.items, s.n, 0, 0);
}

hitable *cornell_box_smoke() { // src/utils.cc:321-357 (constant_medium: not accelerated yet)
  cornell_palette c;
  scene_list s(100);
  cornell_shell(s, c);
  hitable *tall = new translate(
      new rotate_y(new box(vec3(0, 0, 0), vec3(200, 350, 75), c.white), 45.0f), vec3(-200, -300, -100));
  hitable *cube = new translate(
      new rotate_y(new box(vec3(0, 0, 0), vec3(180, 180, 180), c.white), -15.0f), vec3(30, -300, -50));
  s.add(new constant_medium(tall, 0.05, new constant_texture(vec3(1.0, 1.0, 1.0))));
  s.add(new constant_medium(cube, 0.01, new constant_texture(vec3(0.1, 0.0, 0.0))));
  hitable *mist = new sphere(vec3(0, 0, 0), 1000, new dielectric(1.5));
  s.add(new constant_medium(mist, 0.0001, new constant_texture(vec3(1.0, 1.0, 1.0))));
  return new hitable_list(s.items, s.n);
}

// src/utils.cc:359-414: 20x20 random-height boxes under a bvh, lamp, moving sphere, glass, metal
// (fuzz 10 clamps to 1), a glass sphere filled with blue fog, a global mist, the earth, a marble
// sphere and 1000 small spheres in a rotated + translated bvh. Draw order (scene must equal the
// reference's for seed 5489): one draw per box height, then three per small sphere with the vec3
// constructor arguments evaluated right to left.
hitable *oneweek_final_with(unsigned char *tex_data, int nx, int ny) {
  const int nb = 20;
  scene_list top(30), boxes(10000), balls(10000);
  material *white = matte(0.73, 0.73, 0.73);
  material *ground = matte(0.48, 0.83, 0.53);
  for (int i = 0; i < nb; i++) {
    for (int j = 0; j < nb; j++) {
      float w = 100;
      float x0 = -1000 + i * w;
      float z0 = -1000 + j * w;
      float y0 = 0;
      float x1 = x0 + w;
      float y1 = 100 * (drand_r() + 0.01);
      float z1 = z0 + w;
      boxes.add(new box(vec3(x0, y0, z0), vec3(x1, y1, z1), ground));
    }
  }
  top.add(new bvh_node(boxes.items, boxes.n, 0, 1));
  material *light = new diffuse_light(new constant_texture(vec3(7, 7, 7)));
  top.add(new xz_rect(123, 423, 147, 412, 554, light));
  vec3 center(400, 400, 200);
  top.add(new moving_sphere(center, center + vec3(30, 0, 0), 0, 1, 50, matte(0.7, 0.3, 0.1)));
  top.add(new sphere(vec3(260, 150, 45), 50, new dielectric(1.5)));
  top.add(new sphere(vec3(0, 150, 145), 50, new metal(vec3(0.8, 0.8, 0.9), 10.0)));
  hitable *boundary = new sphere(vec3(360, 150, 145), 70, new dielectric(1.5));
  top.add(boundary);
  top.add(new constant_medium(boundary, 0.2, new constant_texture(vec3(0.2, 0.4, 0.9))));
  boundary = new sphere(vec3(0, 0, 0), 5000, new dielectric(1.5));
  top.add(new constant_medium(boundary, 0.0001, new constant_texture(vec3(1.0, 1.0, 1.0))));
  material *emat = new lambertian(new image_texture(tex_data, nx, ny));
  top.add(new sphere(vec3(400, 200, 400), 100, emat));
  texture *pertext = new perlin_noise_texture(0.1);
  top.add(new sphere(vec3(220, 280, 300), 80, new lambertian(pertext)));
  const int ns = 1000;
  for (int j = 0; j < ns; j++)
    balls.add(new sphere(vec3(165 * drand_r(), 165 * drand_r(), 165 * drand_r()), 10, white));
  top.add(new translate(new rotate_y(new bvh_node(balls.items, ns, 0.0, 1.0), 15), vec3(-100, 270, 395)));
  return new hitable_list(top.items, top.n);
}

hitable *oneweek_final() {
  int nx = 0, ny = 0, nn = 0;
  unsigned char *tex_data = load_image_texture("earthmap.jpg", nx, ny, nn);
  return oneweek_final_with(tex_data, nx, ny); // a missing picture is reported by the flattener
}
