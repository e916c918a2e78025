// host/tpt_flatten.h -- hitable tree -> POD arrays of include/tpt.h (the boundary's producer).
//
// Walks the tree in pre-order (bvh: left then right; list: in order; box: its six faces in the
// order src/rect_box.cc:100-112 builds them) and assigns
//   * prim ids   = DFS leaf order (a leaf reachable twice through a one-element bvh_node keeps
//                  one id; its second node is flagged TPT_NODE_DUP),
//   * material / texture ids = order of first encounter,
//   * chain ids  = distinct sequences of translate / rotate_y wrappers, outermost first,
//                  kept as ORDERED op lists (not baked into a matrix) so the device applies
//                  them with the reference's arithmetic (headers/rect_box.h:87-95,
//                  src/rect_box.cc:171-195).
#ifndef TPT_HOST_FLATTEN_H_
#define TPT_HOST_FLATTEN_H_

#include "tpt.h"
#include "tpt_scene.h"

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace tpt {

struct FlatScene {
  std::vector<tpt_node> nodes;
  std::vector<tpt_prim> prims;
  std::vector<tpt_chain> chains;
  std::vector<tpt_xform_op> xform_ops;
  std::vector<tpt_material> materials;
  std::vector<tpt_texture> textures;
  std::vector<tpt_image_desc> images;
  std::vector<std::vector<uint8_t>> image_data; // owned copies behind images[i].rgb
  tpt_perlin_tables perlin;
  bool has_perlin = false;
  std::vector<tpt_light> lights;
  int background = TPT_BG_BLACK;
  int max_depth = 0; // deepest BVH/LIST nesting (frames the parity walk needs)
  int n_root_nodes = 0; // nodes behind this index are medium boundaries

  tpt_scene_desc desc() const; // pointers into this object; valid while it lives unmodified
};

class Flattener {
public:
  explicit Flattener(FlatScene &out);

  // ---- called by hitable::emit ----
  int begin_group(int kind, const AABB &bounds); // returns node index
  void end_group(int node_index);
  void leaf(const hitable *self, int prim_kind, const float *params, int n_params,
            const material *mat, const AABB &bounds);
  void push_xform(const tpt_xform_op &op);
  void pop_xform();
  // constant_medium: `boundary` is emitted into a side list that ends up behind the root tree
  void medium(const hitable *self, const hitable *boundary, float density, const material *phase);
  void toggle_flip() { flip_ = !flip_; }
  void mark_next_dup() { dup_next_ = true; }
  void fail(const std::string &why);

  // ---- called by material::emit / texture::emit ----
  int material_id(const material *m);
  int texture_id(const texture *t);
  int add_material(const tpt_material &m) {
    out_.materials.push_back(m);
    return (int)out_.materials.size() - 1;
  }
  int add_texture(const tpt_texture &t) {
    out_.textures.push_back(t);
    return (int)out_.textures.size() - 1;
  }
  int add_image(const unsigned char *rgb, int w, int h);
  void note_perlin() { out_.has_perlin = true; }

  bool ok() const { return error_.empty(); }
  const std::string &error() const { return error_; }

private:
  int current_chain();
  FlatScene &out_;
  std::vector<tpt_xform_op> stack_;
  std::map<std::vector<int32_t>, int> chain_ids_; // key: raw bits of the op list
  std::map<const material *, int> mat_ids_;
  std::map<const texture *, int> tex_ids_;
  std::map<const unsigned char *, int> image_ids_;
  std::map<std::pair<const hitable *, std::pair<int, int>>, int> prim_ids_;
  std::vector<std::vector<tpt_node>> boundaries_; // one pre-order sub-tree per medium
  std::vector<int> boundary_prims_;               // the MEDIUM primitive that owns boundaries_[i]
  friend bool flatten_scene(const hitable *, const hitable *, int, FlatScene &, std::string &);
  bool flip_ = false;
  bool dup_next_ = false;
  int depth_ = 0;
  std::string error_;
};

// Flatten `world` plus the light-sampling shapes (a hitable_list of xz_rect / sphere as in
// main.cpp:99-106, or a single shape, or nullptr). Returns false and sets `err` on failure.
bool flatten_scene(const hitable *world, const hitable *light_shape, int background,
                   FlatScene &out, std::string &err);

// Light-sampling list derived from the scene itself instead of the hard-coded shapes of
// main.cpp:99-106 (the design regret README.md:17 names): every diffuse_light
// primitive that hitable::random / pdf_value can represent -- an xz_rect or a sphere outside any
// translate / rotate_y wrapper (src/rect_box.cc:26-43, src/sphere.cc:93-120). Empty when the scene
// has none (callers keep the reference list then).
std::vector<tpt_light> derive_light_list(const FlatScene &flat);

// tpt_camera from the host camera's public fields
tpt_camera make_camera_desc(const camera_with_blur &cam);

} // namespace tpt

#endif
