// host/tpt_scene.h -- the scene-description front end (the "plugin surface" that stays C++).
//
// Class names, constructor signatures and public data members follow the reference
// (headers/hitable.h, hitable_list.h, sphere.h, rect_box.h, material.h, texture.h,
// perlin_noise.h, camera.h, utils.h) so scene-building code written against it compiles
// unchanged. What is deliberately DIFFERENT: these classes do not intersect or shade.
// hitable::hit / pdf_value / random, material::scatter / emitted, texture::value and
// camera::get_ray have no host implementation -- that work is the hot path and lives in
// csrc/ as sm_100a CUDA. Instead every node knows how to describe itself to the flattener
// (`emit`), which produces the POD arrays of include/tpt.h.
#ifndef TPT_HOST_SCENE_H_
#define TPT_HOST_SCENE_H_

#include "tpt_math.h"

#include <array>
#include <string>

namespace tpt {
class Flattener;
}

// uniform double in [min,max): thread_local default-seeded std::mt19937, as the reference's
// src/utils.cc:28-32. Used only while BUILDING scenes (random_scene, bvh axis choice, perlin
// gradients); rendering uses the counter-based Philox stream on the device.
double drand_r(double min = 0.0, double max = 1.0);
// restart this thread's generator at the default seed (what a fresh thread would see)
void drand_r_reset();

// ------------------------------------------------------------------------------- textures
class texture {
public:
  virtual ~texture() {}
  virtual int emit(tpt::Flattener &f) const = 0; // returns texture index
};

class constant_texture : public texture {
public:
  constant_texture() {}
  constant_texture(vec3 c) : color_(c) {}
  int emit(tpt::Flattener &f) const override;
  vec3 color_;
};

class checker_texture : public texture {
public:
  checker_texture() {}
  checker_texture(texture *t0, texture *t1) : odd_(t0), even_(t1) {}
  int emit(tpt::Flattener &f) const override;
  texture *odd_ = nullptr;
  texture *even_ = nullptr;
};

// Gradient-noise tables. Static, shared by every instance and re-randomised by every ctor
// (the reference's behaviour, src/perlin_noise.cc:3-21: gradients from drand_r, permutations
// shuffled with a wall-clock seed). The flattener snapshots them after the scene is built.
class perlin_noise {
public:
  perlin_noise();
  static std::array<vec3, 256> random_vec3_;
  static std::array<int, 256> permute_x_;
  static std::array<int, 256> permute_y_;
  static std::array<int, 256> permute_z_;
};

class perlin_noise_texture : public texture {
public:
  perlin_noise_texture(float scale = 1.0) : scale_(scale) {}
  int emit(tpt::Flattener &f) const override;
  perlin_noise noise_;
  float scale_;
};

class image_texture : public texture {
public:
  image_texture() {}
  image_texture(unsigned char *pixels, int width, int height)
      : data_(pixels), width_(width), height_(height) {}
  int emit(tpt::Flattener &f) const override;
  unsigned char *data_ = nullptr;
  int width_ = 0, height_ = 0;
};

// ------------------------------------------------------------------------------ materials
class material {
public:
  virtual ~material() {}
  virtual int emit(tpt::Flattener &f) const; // base class == absorber (scatter false, emits 0)
};

class lambertian : public material {
public:
  lambertian(texture *albedo) : albedo_(albedo) {}
  int emit(tpt::Flattener &f) const override;
  texture *albedo_;
};

class metal : public material {
public:
  metal(const vec3 &albedo, float fuzz) : albedo_(albedo) { fuzz_ = (fuzz < 1 && fuzz >= 0) ? fuzz : 1; }
  int emit(tpt::Flattener &f) const override;
  vec3 albedo_;
  float fuzz_;
};

class dielectric : public material {
public:
  dielectric(float ri) : ref_idx_(ri) {}
  int emit(tpt::Flattener &f) const override;
  float ref_idx_;
};

class diffuse_light : public material {
public:
  diffuse_light(texture *a) : emit_(a) {}
  int emit(tpt::Flattener &f) const override;
  texture *emit_;
};

// At the reference's HEAD isotropic::scatter has a stale signature and never overrides
// material::scatter (headers/material.h:74-80 vs :13-16), i.e. it behaves as an absorber.
class isotropic : public material {
public:
  isotropic(texture *t) : albedo_(t) {}
  int emit(tpt::Flattener &f) const override;
  texture *albedo_;
};

// ------------------------------------------------------------------------------- hitables
class hitable {
public:
  virtual ~hitable() {}
  virtual bool bounding_box(float t0, float t1, AABB &box) const = 0;
  // describe this node (and its subtree) to the flattener
  virtual void emit(tpt::Flattener &f) const = 0;
};

class hitable_list : public hitable {
public:
  hitable_list() {}
  hitable_list(hitable **l, int n) : list_(l), list_size_(n) {}
  bool bounding_box(float t0, float t1, AABB &box) const override;
  void emit(tpt::Flattener &f) const override;
  hitable **list_ = nullptr;
  int list_size_ = 0;
};

class bvh_node : public hitable {
public:
  bvh_node() {}
  // random-axis median split over l[0..n) (sorts l in place), as src/hitable.cc:32-56
  bvh_node(hitable **l, int n, float time0, float time1);
  bool bounding_box(float t0, float t1, AABB &box) const override;
  void emit(tpt::Flattener &f) const override;
  hitable *left_ = nullptr;
  hitable *right_ = nullptr;
  AABB box_;
};

class sphere : public hitable {
public:
  sphere() {}
  sphere(vec3 center, float radius, material *m) : center_(center), radius_(radius), mat_ptr_(m) {}
  bool bounding_box(float t0, float t1, AABB &box) const override;
  void emit(tpt::Flattener &f) const override;
  vec3 center_;
  float radius_ = 0;
  material *mat_ptr_ = nullptr;
};

class moving_sphere : public hitable {
public:
  moving_sphere() {}
  moving_sphere(vec3 center0, vec3 center1, float t0, float t1, float radius, material *m)
      : center0_(center0), center1_(center1), radius_(radius), mat_ptr_(m), time0_(t0), time1_(t1) {}
  vec3 center(float time) const {
    return center0_ + (time - time0_) / (time1_ - time0_) * (center1_ - center0_);
  }
  bool bounding_box(float t0, float t1, AABB &box) const override;
  void emit(tpt::Flattener &f) const override;
  vec3 center0_, center1_;
  float radius_ = 0;
  material *mat_ptr_ = nullptr;
  float time0_ = 0, time1_ = 1;
};

class xy_rect : public hitable {
public:
  xy_rect() {}
  xy_rect(float x0, float x1, float y0, float y1, float k, material *mat)
      : mat_ptr_(mat), x0_(x0), x1_(x1), y0_(y0), y1_(y1), k_(k) {}
  bool bounding_box(float t0, float t1, AABB &box) const override;
  void emit(tpt::Flattener &f) const override;
  material *mat_ptr_ = nullptr;
  float x0_ = 0, x1_ = 0, y0_ = 0, y1_ = 0, k_ = 0;
};

class xz_rect : public hitable {
public:
  xz_rect() {}
  xz_rect(float x0, float x1, float z0, float z1, float k, material *mat)
      : mat_ptr_(mat), x0_(x0), x1_(x1), z0_(z0), z1_(z1), k_(k) {}
  bool bounding_box(float t0, float t1, AABB &box) const override;
  void emit(tpt::Flattener &f) const override;
  material *mat_ptr_ = nullptr;
  float x0_ = 0, x1_ = 0, z0_ = 0, z1_ = 0, k_ = 0;
};

class yz_rect : public hitable {
public:
  yz_rect() {}
  yz_rect(float y0, float y1, float z0, float z1, float k, material *mat)
      : mat_ptr_(mat), y0_(y0), y1_(y1), z0_(z0), z1_(z1), k_(k) {}
  bool bounding_box(float t0, float t1, AABB &box) const override;
  void emit(tpt::Flattener &f) const override;
  material *mat_ptr_ = nullptr;
  float y0_ = 0, y1_ = 0, z0_ = 0, z1_ = 0, k_ = 0;
};

class flip_normal : public hitable {
public:
  flip_normal(hitable *p) : ptr_(p) {}
  bool bounding_box(float t0, float t1, AABB &box) const override {
    return ptr_->bounding_box(t0, t1, box);
  }
  void emit(tpt::Flattener &f) const override;
  hitable *ptr_;
};

class box : public hitable {
public:
  box() {}
  box(vec3 pmin, vec3 pmax, material *mat);
  bool bounding_box(float, float, AABB &b) const override {
    b = AABB(point_min_, point_max_);
    return true;
  }
  void emit(tpt::Flattener &f) const override;
  vec3 point_min_, point_max_;
  hitable *list_ptr_ = nullptr;
};

class translate : public hitable {
public:
  translate(hitable *p, const vec3 &offset) : ptr_(p), offset_(offset) {}
  bool bounding_box(float t0, float t1, AABB &b) const override {
    if (!ptr_->bounding_box(t0, t1, b)) return false;
    b = AABB(b.min() + offset_, b.max() + offset_);
    return true;
  }
  void emit(tpt::Flattener &f) const override;
  hitable *ptr_;
  vec3 offset_;
};

class rotate_y : public hitable {
public:
  rotate_y(hitable *p, float angle);
  bool bounding_box(float, float, AABB &b) const override {
    b = box_;
    return has_box_;
  }
  void emit(tpt::Flattener &f) const override;
  hitable *ptr_;
  float sin_theta_, cos_theta_;
  bool has_box_;
  AABB box_;
};

// Participating medium (headers/hitable.h:58-69): flattened to a MEDIUM primitive whose boundary
// sub-tree is kept behind the world tree (tpt_flatten.cc).
class constant_medium : public hitable {
public:
  constant_medium(hitable *boundary, float density, texture *tex)
      : boundary_(boundary), density_(density), phase_funcion_(new isotropic(tex)) {}
  bool bounding_box(float t0, float t1, AABB &b) const override {
    return boundary_->bounding_box(t0, t1, b);
  }
  void emit(tpt::Flattener &f) const override;
  hitable *boundary_;
  float density_;
  material *phase_funcion_;
};

// --------------------------------------------------------------------------------- camera
// Thin-lens camera with a shutter interval: the ctor (frame construction) stays on the host,
// get_ray is part of the device sample loop.
class camera_with_blur {
public:
  camera_with_blur(vec3 lookfrom, vec3 lookat, vec3 vup, float vfov, float aspect, float aperture,
                   float focus_dist, float t0, float t1);
  vec3 origin_, lower_left_corner_, vertical_, horizontal_;
  vec3 u_, v_, w_;
  float lens_radius_;
  float time0, time1;
};
using camera = camera_with_blur;

// ---------------------------------------------------------------------------- scene builders
hitable *random_scene();
hitable *two_checker_spheres();
hitable *two_perlin_spheres();
hitable *light_spheres();
hitable *sphere_cornell_box();
hitable *cornell_box();
hitable *cornell_box_smoke();
hitable *oneweek_final();                                   // loads ./earthmap.jpg like the reference
hitable *oneweek_final_with(unsigned char *rgb, int w, int h); // same scene, texture bytes supplied

// decoded picture, 3 bytes per pixel; supports binary PPM (P6) and baseline JPEG
unsigned char *load_image_texture(std::string filename, int &width, int &height, int &channels);

#endif // TPT_HOST_SCENE_H_
