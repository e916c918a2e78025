/* include/tpt.h -- C-ABI of the B200-native path-tracing core (libtpt.so).
 *
 * Drop-in boundary for the per-pixel sample loop of BlurryLight/tiny-path-tracer:
 * it replaces exactly main.cpp:109-175 (the `render_a_pixel` lambda, its std::async dispatch and
 * the wait) and everything that loop calls: color() src/utils.cc:58-94, hitable::hit
 * (src/hitable.cc:63-90, src/hitable_list.cc:38-51, src/sphere.cc:13-75, src/rect_box.cc:8-91,
 * 171-195, headers/rect_box.h:50-57,87-95, src/aabb.cc:3-19), material::scatter / emitted /
 * scattering_pdf (src/material.cc), the pdf classes (headers/utils.h:65-98,
 * headers/rect_box.h:118-131), texture::value (src/texture.cc) and camera::get_ray
 * (src/camera.cc:23-31).
 *
 * The reference has no FFI: main() calls C++ virtuals directly. The host keeps the reference's
 * C++ scene-description classes (tiny-path-tracer_b200/host), flattens the hitable tree into the
 * POD arrays below, and calls these entry points. Plain pointers and sizes only; no exceptions
 * cross the boundary; every function returns TPT_OK (0) or a negative tpt_status and records a
 * message retrievable with tpt_last_error().
 *
 * There is NO CPU fallback: every compute entry point fails with TPT_ERR_NO_DEVICE when no
 * sm_100 device is usable.
 */
#ifndef TPT_H_
#define TPT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPT_API_VERSION 2 /* 2: tpt_stats grew the per-GPU fields of tpt_render_multi */

typedef enum tpt_status {
  TPT_OK = 0,
  TPT_ERR_INVALID = -1,     /* bad argument / malformed scene description          */
  TPT_ERR_CUDA = -2,        /* a CUDA runtime call failed (message has the detail)  */
  TPT_ERR_NO_DEVICE = -3,   /* no usable GPU: there is deliberately no CPU path     */
  TPT_ERR_UNSUPPORTED = -4, /* scene uses a construct outside the accelerated path  */
  TPT_ERR_NOMEM = -5
} tpt_status;

/* ---------------------------------------------------------------------------------------------
 * Flattened scene (producer: host flattener; mirrors the reference's hitable tree 1:1).
 *
 * `nodes` is the hitable tree in PRE-ORDER. Each record is 32 bytes (two float4):
 *   BVH   : bounds = bvh_node::box_ (headers/hitable.h:53); `end` = index one past its subtree.
 *           Children are the two sub-trees that follow (left first). A bvh_node built from one
 *           element has left_ == right_ (src/hitable.cc:41-42): the flattener emits that child
 *           twice, the second copy flagged TPT_NODE_DUP.
 *   LIST  : hitable_list (and `box`, which forwards to a 6-face list, src/rect_box.cc:93-119);
 *           children follow in list order; `end` as above.
 *   LEAF  : `prim` indexes `prims`; bounds = the leaf's bounding_box().
 * flip_normal / translate / rotate_y are not nodes: a leaf carries its flip parity and the id of
 * its transform chain (the translate / rotate_y wrappers between the root and the leaf, outermost
 * first); BVH nodes that live under a transform carry the chain id too.
 *
 * Bounds are what the reference's bounding_box() returns, so every primitive lies inside the boxes of the
 * nodes above it. PARITY mode honours the boxes exactly as given (a primitive is tested whenever the boxes
 * above it are crossed, wherever it is), and speeds its walk up only after verifying that containment;
 * FAST mode builds its own acceleration structure from the leaf bounds and relies on it.
 * ------------------------------------------------------------------------------------------- */
enum { TPT_NODE_BVH = 0, TPT_NODE_LIST = 1, TPT_NODE_LEAF = 2 };
enum { TPT_NODE_DUP = 0x100 }; /* or'ed into tpt_node.kind */

typedef struct tpt_node {
  float bmin[3];
  int32_t kind; /* TPT_NODE_* | flags */
  float bmax[3];
  int32_t end_or_prim; /* BVH/LIST: one past the subtree; LEAF: primitive index */
} tpt_node;        /* 32 bytes */

/* one translate / rotate_y wrapper (headers/rect_box.h:77-116) */
enum { TPT_XF_TRANSLATE = 0, TPT_XF_ROTATE_Y = 1 };
typedef struct tpt_xform_op {
  int32_t kind;
  float a, b, c; /* TRANSLATE: offset_.xyz ; ROTATE_Y: sin_theta_, cos_theta_, unused */
} tpt_xform_op;

typedef struct tpt_chain {
  int32_t first_op; /* index into xform_ops, outermost wrapper first */
  int32_t n_ops;    /* 0 for the identity chain (chain 0 is always identity) */
} tpt_chain;

enum {
  TPT_PRIM_SPHERE = 0,        /* p = cx,cy,cz,radius                                (headers/sphere.h:18-20)  */
  TPT_PRIM_MOVING_SPHERE = 1, /* p = c0.xyz, radius, c1.xyz, time0, time1          (headers/sphere.h:36-39)  */
  TPT_PRIM_XY_RECT = 2,       /* p = x0,x1,y0,y1,k                                  (headers/rect_box.h:14)   */
  TPT_PRIM_XZ_RECT = 3,       /* p = x0,x1,z0,z1,k                                  (headers/rect_box.h:29)   */
  TPT_PRIM_YZ_RECT = 4,       /* p = y0,y1,z0,z1,k                                  (headers/rect_box.h:41)   */
  TPT_PRIM_MEDIUM = 5         /* constant_medium (headers/hitable.h:58-69): p[0] = density_, p[1], p[2] = the
                                 boundary's node range [first, end) as int32 bit patterns; `material` = its
                                 phase function. The boundary sub-tree lives BEHIND the root tree in `nodes`
                                 (indices >= n_root_nodes) and is only reached through this primitive. */
};
enum { TPT_PRIM_FLIP = 1 }; /* odd number of flip_normal wrappers above the leaf */

typedef struct tpt_prim {
  int32_t kind;
  int32_t material; /* index into materials */
  int32_t chain;    /* index into chains */
  int32_t flags;
  float p[12];
} tpt_prim; /* 64 bytes */

enum {
  TPT_MAT_LAMBERTIAN = 0,    /* texture                       (headers/material.h:27-37) */
  TPT_MAT_METAL = 1,         /* albedo, fuzz                  (headers/material.h:39-52) */
  TPT_MAT_DIELECTRIC = 2,    /* ref_idx                       (headers/material.h:53-59) */
  TPT_MAT_DIFFUSE_LIGHT = 3, /* texture                       (headers/material.h:61-72) */
  TPT_MAT_ABSORBER = 4,      /* base material: scatter()==false, emitted()==0 (headers/material.h:13-24) */
  TPT_MAT_ISOTROPIC = 5      /* isotropic(texture) (headers/material.h:74-80). At the reference's HEAD its scatter()
                                has a stale signature and never overrides material::scatter, so it behaves as an
                                absorber; the texture is carried but not evaluated */
};
typedef struct tpt_material {
  int32_t kind;
  int32_t texture; /* index into textures (LAMBERTIAN, DIFFUSE_LIGHT) or -1 */
  float albedo[3]; /* METAL */
  float fuzz;      /* METAL (already clamped by the ctor, headers/material.h:41-46) */
  float ref_idx;   /* DIELECTRIC */
  int32_t pad;
} tpt_material; /* 32 bytes */

enum {
  TPT_TEX_CONSTANT = 0, /* color                                   (headers/texture.h:15-24) */
  TPT_TEX_CHECKER = 1,  /* odd, even = texture indices             (headers/texture.h:25-33) */
  TPT_TEX_PERLIN = 2,   /* scale                                   (headers/texture.h:35-41) */
  TPT_TEX_IMAGE = 3     /* image index                             (headers/texture.h:42-51) */
};
typedef struct tpt_texture {
  int32_t kind;
  float color[3];
  int32_t odd, even; /* CHECKER */
  float scale;       /* PERLIN */
  int32_t image;     /* IMAGE: index into images */
} tpt_texture;       /* 32 bytes */

typedef struct tpt_image_desc {
  const uint8_t *rgb; /* width*height*3 bytes, row-major, exactly what stbi_load returned (3 channels) */
  int32_t width, height;
} tpt_image_desc;

/* the static tables of perlin_noise (headers/perlin_noise.h:13-16), copied AFTER scene
 * construction because every perlin_noise ctor re-randomises them (src/perlin_noise.cc:3-21) */
typedef struct tpt_perlin_tables {
  float ranvec[256][3];
  int32_t perm_x[256], perm_y[256], perm_z[256];
} tpt_perlin_tables;

/* light-sampling shapes: the hitable_list handed to color() as `light_shape`
 * (main.cpp:99-106). Only xz_rect and sphere override pdf_value()/random()
 * (src/rect_box.cc:26-43, src/sphere.cc:93-120); anything else behaves like the hitable base
 * class (pdf 0, direction (1,0,0): headers/hitable.h:35-38). */
enum { TPT_LIGHT_XZ_RECT = 0, TPT_LIGHT_SPHERE = 1, TPT_LIGHT_OTHER = 2 };
typedef struct tpt_light {
  int32_t kind;
  float p[5]; /* XZ_RECT: x0,x1,z0,z1,k ; SPHERE: cx,cy,cz,r */
  int32_t pad[2];
} tpt_light; /* 32 bytes */

enum { TPT_BG_BLACK = 0, /* HEAD: src/utils.cc:86 */
       TPT_BG_SKY = 1    /* the commented gradient src/utils.cc:87-90: 0.1*((1-t)*white + t*(0.5,0.7,1)) */ };

typedef struct tpt_scene_desc {
  int32_t api_version; /* TPT_API_VERSION */
  int32_t n_nodes, n_prims, n_chains, n_xform_ops, n_materials, n_textures, n_images, n_lights;
  const tpt_node *nodes;
  const tpt_prim *prims;
  const tpt_chain *chains;
  const tpt_xform_op *xform_ops;
  const tpt_material *materials;
  const tpt_texture *textures;
  const tpt_image_desc *images;
  const tpt_perlin_tables *perlin; /* may be NULL when no PERLIN texture exists */
  const tpt_light *lights;
  int32_t background;
  int32_t n_root_nodes; /* nodes [0, n_root_nodes) are the world tree; 0 means n_nodes (no medium boundaries) */
} tpt_scene_desc;

/* public fields of camera_with_blur after its ctor ran on the host (headers/camera.h:14-20) */
typedef struct tpt_camera {
  float origin[3], lower_left_corner[3], vertical[3], horizontal[3];
  float u[3], v[3], w[3];
  float lens_radius, time0, time1;
} tpt_camera;

/* ray = origin, direction (not normalised), time (headers/ray.h:15-17) */
typedef struct tpt_ray {
  float o[3], d[3], time;
} tpt_ray; /* 28 bytes */

/* hit_record (headers/hitable.h:14-21) + the ids the reference cannot report */
typedef struct tpt_hit {
  int32_t hit;  /* 0/1 */
  int32_t prim; /* DFS leaf index (box faces count individually), -1 on miss */
  int32_t mat;  /* material index, -1 on miss */
  float t, u, v, p[3], n[3];
} tpt_hit; /* 48 bytes */

enum {
  TPT_MODE_PARITY = 0, /* reference arithmetic: fp64 where the reference promotes, no FMA contraction,
                          un-shrunk t_max BVH walk with the reference's tie rules */
  TPT_MODE_FAST = 1    /* same estimator, fp32 + FMA + fast intrinsics, culled traversal */
};
enum {
  TPT_KERNEL_MEGA = 0,     /* persistent megakernel, one path per lane with regeneration */
  TPT_KERNEL_WAVEFRONT = 1 /* generate / extend / shade queues */
};

typedef struct tpt_render_params {
  int32_t nx, ny, ns, max_depth; /* main.cpp:31-37; max_depth <= 65535 */
  int32_t slices;   /* bonus_pic when allow_bonus_pic, else 1 (main.cpp:111-114); must divide into ns>=slices */
  int32_t mode;     /* TPT_MODE_*   */
  int32_t kernel;   /* TPT_KERNEL_* */
  uint32_t seed_lo, seed_hi; /* Philox key */
  float t_min;      /* 0.001f (src/utils.cc:61) */
  /* static split of the image across workers: this call renders tiles with
   * (tile_index % part_count) == part_index ; tiles are TPT_TILE x TPT_TILE pixels, row-major */
  int32_t part_index, part_count;
  int32_t device;   /* CUDA device ordinal for single-device calls */
  int32_t reserved[4]; /* zero for defaults. [1] > 0: sample sub-ranges per slice (work granularity);
                          [2] = 1: trace every path, also those of pixels none of whose rays can reach the
                          scene's bounds (by default such (pixel, sample) pairs are finished untraced with
                          their exact result, 0 on the black background -- tpt_stats.culled_paths) */
} tpt_render_params;

#define TPT_TILE 16
#define TPT_MAX_GPUS 16

/* caller-owned HOST buffers; any pointer may be NULL to skip that product */
typedef struct tpt_image {
  float *sum_rgb;   /* [slices][ny][nx][3] running radiance sums (pixel_sample_cols, main.cpp:127-133);
                       row 0 = bottom row (v=(j+xi)/ny, main.cpp:122) */
  uint8_t *rgb8;    /* [ny][nx][3] final picture: int(255.99f*sqrt(sum/ns)) clamped (main.cpp:135-139,176-182) */
  uint8_t *rgb8_slices; /* [slices][ny][nx][3] bonus pictures (main.cpp:191-215) */
} tpt_image;

typedef struct tpt_stats {
  uint64_t paths;         /* (pixel,sample) pairs traced by the last render */
  uint64_t rays;          /* world->hit queries actually traced            */
  uint64_t nan_samples;   /* samples zeroed by de_nan (headers/utils.h:100-109) */
  double render_ms;       /* device time of the path-tracing kernel(s), CUDA events on the library stream */
  double resolve_ms;      /* device time of the resolve / quantise kernel  */
  double h2d_ms, d2h_ms;  /* scene/camera upload, image download           */
  double wall_ms;         /* host wall clock of the whole call             */
  uint64_t h2d_bytes, d2h_bytes;
  int32_t kernel_launches; /* kernels launched by the last call            */
  int32_t sm_count;
  int32_t blocks, threads_per_block;
  int32_t reserved[4];     /* [0] = sample ranges per pixel of the last single-device render
                              (tpt_render_multi: batches taken by GPUs 0..3) */
  uint64_t culled_paths;   /* of `paths`: those of pixels whose whole ray bundle misses the scene's bounds;
                              world->hit is false for every ray they can generate, so they are finished
                              without tracing (0 rays counted for them). See tpt_render_params.reserved[2]. */
  /* tpt_render_multi only (multi_gpus = 0 after a single-device call): what each GPU did */
  int32_t multi_gpus;
  int32_t multi_batches_total;            /* 8 x n_gpus batches of interleaved tiles                      */
  int32_t multi_batches[TPT_MAX_GPUS];    /* batches rendered by GPU g: static share + stolen            */
  int32_t multi_stolen[TPT_MAX_GPUS];     /* of those, taken from the shared work-stealing counter       */
  double multi_busy_ms[TPT_MAX_GPUS];     /* device time of GPU g's batches (CUDA events on its stream)  */
  double multi_gather_ms;                 /* peer gather on GPU 0 (host clock around its stream)         */
} tpt_stats;

typedef struct tpt_scene tpt_scene; /* opaque: owns the device copies */

int tpt_api_version(void);
int tpt_device_count(void);
/* creates the CUDA context of `device` now instead of inside the first tpt_scene_create (a few hundred ms in a
 * fresh process): a driver calls it from a second host thread while it parses config.ini and builds the scene */
int tpt_device_warm(int device);
const char *tpt_last_error(void);

/* copies every array of `desc` (host and device side); the caller may free its arrays afterwards */
int tpt_scene_create(const tpt_scene_desc *desc, int device, tpt_scene **out);
void tpt_scene_destroy(tpt_scene *scene);

/* gate-1 entry point: world->hit(r, tmin, tmax, rec) for a batch of rays (host buffers) */
int tpt_intersect_batch(const tpt_scene *scene, const tpt_ray *rays, size_t n, float tmin, float tmax,
                        int mode, tpt_hit *out);

/* the sample loop. Host buffers in/out: uploads camera+params, renders, resolves, downloads. */
int tpt_render(tpt_scene *scene, const tpt_camera *cam, const tpt_render_params *params, tpt_image *out);

/* device-resident variant: same work, the products stay in HBM (fetch later with tpt_render_fetch) */
int tpt_render_device(tpt_scene *scene, const tpt_camera *cam, const tpt_render_params *params);
int tpt_render_fetch(tpt_scene *scene, tpt_image *out);

/* In-process multi-GPU: `scenes[g]` is the same scene created on GPU g (tpt_scene_create with
 * device = g). The frame is cut into 8 x n batches of interleaved 16x16 tiles; every GPU renders a
 * static share (7 of its 8 batches, one launch) and then steals the remaining batches from a shared
 * counter (one host thread per GPU inside the call); GPU 0 fetches every tile from the GPU that rendered
 * it over NVLink (one kernel over peer memory, no NCCL) and the image is downloaded once. The result is bit-identical to a single-GPU tpt_render (Philox is keyed on
 * pixel and sample). params->part_index/part_count must be 0/1. Statistics are read from
 * scenes[0]. */
int tpt_render_multi(tpt_scene *const *scenes, int n_scenes, const tpt_camera *cam, const tpt_render_params *params,
                     tpt_image *out);

/* device addresses of the last render's products (valid until the next render / destroy): lets a
 * multi-GPU caller combine the disjoint per-GPU tiles over NVLink before one download */
int tpt_device_buffers(const tpt_scene *scene, void **sum_rgb, size_t *sum_bytes, void **rgb8, size_t *rgb8_bytes);

int tpt_get_stats(const tpt_scene *scene, tpt_stats *out);

/* known-answer probes used by the unit tests (each is one tiny kernel launch) */
int tpt_debug_philox(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int tpt_debug_texture(const tpt_scene *scene, int texture, const float *uvp /* n x 5: u,v,px,py,pz */,
                      size_t n, int mode, float *out_rgb);
/* measured FP32 FMA throughput of `device` (TFLOP/s, 2 FLOP per FMA; best of three ~35 ms launches of
 * independent FMA chains at full occupancy): the denominator bench.py's FP32 roofline is quoted on */
int tpt_debug_fp32_peak(int device, double *tflops, double *ms);

/* host only, no device needed: what tpt_scene_create's small-scene table (the constant-bank copy the
 * warp-uniform closest hit reads) would hold for this scene. out[0] enabled, [1] groups, [2] loose
 * rects, [3] spheres, [4] blocks; then 8 ints per block: six face primitive ids in the order
 * -x +x -y +y -z +z (-1 = absent), the chain, 1 when the block is open */
int tpt_debug_small_scene(const tpt_scene_desc *scene, int32_t out[64]);

#ifdef __cplusplus
}
#endif
#endif /* TPT_H_ */
