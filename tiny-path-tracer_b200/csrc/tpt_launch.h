// csrc/tpt_launch.h -- kernel argument blocks and the launcher entry points exported by the two
// kernel translation units (parity / fast) to the C-ABI layer (tpt_api.cu).
#pragma once
#include "tpt_device.cuh"

#define TPT_MEGA_THREADS 256
#ifndef TPT_WAVE_SPLIT_GEN
#define TPT_WAVE_SPLIT_GEN 2 // generate as its own phase after shade: 0 never, 1 always, 2 in the LARGE-scene parity kernels only (see render_wave_kernel)
#endif
#ifndef TPT_TRACE_ENABLE
#define TPT_TRACE_ENABLE 1
#endif
#ifndef TPT_TRACE_ALL
#define TPT_TRACE_ALL 0 // 1: every SAH BVH scene takes the "trace" variant in fast mode, 0: only scenes with participating media
#endif
#ifndef TPT_TRACE_VOTE
#define TPT_TRACE_VOTE 1 // "trace" variant: vote-scheduled walk between the ray hand-outs (closest_hit_fbvh_vote; oneweek_final +5 %)
#endif
#ifndef TPT_TRACE_CARRY
#define TPT_TRACE_CARRY 1 // "trace" variant: walks still running when the phase's rays are handed out carry over to the next iteration
#endif
#ifndef TPT_TRACE_CARRY_K
#define TPT_TRACE_CARRY_K 8 // ... once fewer than this many lanes of the warp are still walking (8 / 16 / 24: 1 962 / 1 954 / 1 937 Mpaths/s on oneweek_final together with the vote)
#endif
#define TPT_SLOT_PARKED (1 << 20) // depth word of a slot whose walk is parked / finished after having been parked
#define TPT_SLOT_LATE (1 << 21)
#ifndef TPT_TRACE_THREADS
#define TPT_TRACE_THREADS TPT_WAVE_THREADS
#endif
#ifndef TPT_TRACE_SLOTS
#define TPT_TRACE_SLOTS 768    // path slots per CTA of the BVH ("trace") wavefront variant: 3 rays per lane to hand out
#endif
#ifndef TPT_TRACE_MIN_BLOCKS
#define TPT_TRACE_MIN_BLOCKS 2 // 2 CTAs/SM: 128 registers for the traversal state, 2 x 78 KB of slots (more starves L1 of the BVH)
#endif
#ifndef TPT_WAVE_MIN_BLOCKS
#define TPT_WAVE_MIN_BLOCKS 3 // __launch_bounds__ min blocks per SM: 80 registers, 3 CTAs (measured best)
#endif
#ifndef TPT_PAR_CAMERA_TRIES
#define TPT_PAR_CAMERA_TRIES 0 // parity kernels redraw camera rays that miss the root bvh_node's box inside generate (like the fast kernels)
#endif
#ifndef TPT_PAR_LEAN
#define TPT_PAR_LEAN 1 // parity wavefront kernel without the texture / uv code for scenes whose textures are all constant
#endif
#ifndef TPT_WAVE_PAR_MIN_BLOCKS
#define TPT_WAVE_PAR_MIN_BLOCKS TPT_WAVE_MIN_BLOCKS // parity kernels
#endif
#ifndef TPT_WAVE_LEAN_MIN_BLOCKS
#define TPT_WAVE_LEAN_MIN_BLOCKS 3 // lean small-scene kernels; 4 (64 registers, four 52 KB CTAs per SM) gains 1 % without the camera redraws and loses 6 % with them (72 B of spills)
#endif
#ifndef TPT_WAVE_DYNAMIC
#define TPT_WAVE_DYNAMIC 1 // merged shade+generate phase: chunks handed out through a shared counter, longest first
#endif
#ifndef TPT_WAVE_CAMERA_TRIES
#define TPT_WAVE_CAMERA_TRIES 3 // frames where some pixel can look past the scene: camera rays that miss its bounds are redrawn inside generate (up to this many per visit) instead of spending an extend pass
#endif
#ifndef TPT_WAVE_THREADS
#define TPT_WAVE_THREADS 256
#endif
#ifndef TPT_WAVE_SLOTS
#define TPT_WAVE_SLOTS 512 // path slots per CTA (two per thread)
#endif

#ifndef TPT_SMALL_THREADS
#define TPT_SMALL_THREADS 384 // small-scene (constant-bank intersector) kernels: 384 threads, 2 CTAs/SM
#endif
#ifndef TPT_SMALL_MIN_BLOCKS
#define TPT_SMALL_MIN_BLOCKS 2
#endif
#ifndef TPT_SMALL_SLOTS_FAST
#define TPT_SMALL_SLOTS_FAST 1152 // three path slots per thread
#endif
#ifndef TPT_SMALL_SLOTS_FAST_TEX
#define TPT_SMALL_SLOTS_FAST_TEX 1024 // textured small scenes: their tables (Perlin: +7 KB) must still leave room for two CTAs (r02 sweep: 768 / 960 / 1024 slots = 6 647 / 6 842 / 6 947 Mpaths/s on two_perlin_spheres)
#endif
#ifndef TPT_SMALL_SLOTS_PAR
#define TPT_SMALL_SLOTS_PAR 768
#endif

namespace tptd {

// CTA shape of a wavefront kernel instantiation: threads, path slots, minimum resident CTAs per SM.
// Chosen per kernel family from measurements on B200 (profiles/r02_tuning_sweeps.txt):
//  * small scenes (everything on chip, warp-uniform intersectors): the hot code is about as large as
//    the SM's instruction cache, and CTAs sit in different phases of it. Two CTAs of 384 threads
//    instead of three of 256 keep fewer phases' code live at once (same 24 warps per SM):
//    parity +6 %, and with three slots per thread the fast kernels gain another 3-5 % (fuller
//    chunks, fewer barriers per path); parity prefers two slots per thread.
//  * SAH-BVH / large scenes: 256 x 512 x 3 (r01 sweep), the media build 256 x 768 x 2 (128 registers).
__host__ __device__ constexpr int wave_threads(bool small, bool trace) { return trace ? TPT_TRACE_THREADS : (small ? TPT_SMALL_THREADS : TPT_WAVE_THREADS); }
// Two CTAs must fit the SM's 227 KB of shared memory together with their copies of the scene tables:
// 1152 slots (108 KB of path state) leave room for the 4 KB tables of a constant-texture scene only; a
// scene with Perlin tables (+7 KB) would drop to ONE resident CTA (r02: two_perlin_spheres 6 860 -> 4 720
// Mpaths/s), so the textured builds take 1024 slots (room for 14 KB of tables).
__host__ __device__ constexpr int wave_slots(bool par, bool small, bool trace, bool lean) {
  return trace ? TPT_TRACE_SLOTS : (small ? (par ? TPT_SMALL_SLOTS_PAR : (lean ? TPT_SMALL_SLOTS_FAST : TPT_SMALL_SLOTS_FAST_TEX)) : TPT_WAVE_SLOTS);
}
__host__ __device__ constexpr int wave_min_blocks(bool par, bool small, bool trace, bool lean) {
  return trace ? TPT_TRACE_MIN_BLOCKS : (small ? TPT_SMALL_MIN_BLOCKS : (lean ? TPT_WAVE_LEAN_MIN_BLOCKS : (par ? TPT_WAVE_PAR_MIN_BLOCKS : TPT_WAVE_MIN_BLOCKS)));
}

struct IntersectArgs {
  SceneLayout scene;
  union { // one of the two per launch: FAST kernels read `small`, PARITY kernels read `flat`
    SmallScene small;
    FlatTree flat;
  };
  const float *rays; // n x 7
  size_t n;
  float tmin, tmax;
  tpt_hit *out;
};

struct RenderArgs {
  SceneLayout scene;
  union { // one of the two per launch: FAST kernels read `small`, PARITY kernels read `flat`
    SmallScene small;
    FlatTree flat;
  };
  CamView cam;
  int nx, ny, ns, max_depth;
  float t_min;
  uint32_t rk[20]; // Philox round keys: rk[2r] = seed_lo + r * 0x9E3779B9, rk[2r+1] = seed_hi + r * 0xBB67AE85
  int n_ranges;                          // sample ranges per pixel (slices x sub-ranges)
  int range_bounds[TPT_MAX_RANGES + 1];  // range r covers samples [bounds[r], bounds[r+1])
  int tiles_x, tiles_y, part_index, part_count;
  int part_group, part_stride;           // > 1: this launch owns parts part_index + r * part_stride, r < part_group (one launch for a GPU's static share)
  unsigned long long n_bins;             // local tiles x TILE^2 x n_ranges
  float *acc;                            // [n_ranges][ny*nx][3] per-range radiance sums
  unsigned long long *counters;          // [0] next bin, [1] rays, [2] NaN samples, [3] paths, [4] paths resolved by the bundle test
  // Pixel-bundle bounds test (black background only): the rectangle of pixels [cull_x0, cull_x1] x
  // [cull_y0, cull_y1] outside of which no ray the camera can generate -- any lens point, any
  // jitter -- reaches the scene's padded root box (computed on the host, make_plan). A (pixel,
  // sample range) bin outside it is finished when it is taken: world->hit is false for all its
  // rays, so the bin's sum is exactly 0.
  int cull;
  int cull_x0, cull_x1, cull_y0, cull_y1;
  int camera_tries; // wavefront generate: camera rays drawn per visit while they miss the scene's bounds (1 = no test)
  int tex_mask; // TPT_TEXF_* features of the scene's textures: small-scene kernels are built per feature set (0 = every texture is a constant_texture)
};

struct TextureProbeArgs {
  SceneLayout scene;
  int texture;
  const float *uvp; // n x 5
  size_t n;
  float *out; // n x 3
};

cudaError_t launch_intersect_parity(const IntersectArgs &A, bool smem, bool small, cudaStream_t st);
cudaError_t launch_intersect_fast(const IntersectArgs &A, bool smem, bool small, cudaStream_t st);
// `media`: scene has constant_medium primitives (the stream is threaded through world->hit)
// `small`: scene has few enough primitives for the warp-uniform brute-force closest hit
cudaError_t mega_occupancy_parity(const RenderArgs &A, bool smem, bool small, bool media, size_t smem_bytes, int *blocks_per_sm);
cudaError_t mega_occupancy_fast(const RenderArgs &A, bool smem, bool small, bool media, size_t smem_bytes, int *blocks_per_sm);
cudaError_t launch_mega_parity(const RenderArgs &A, bool smem, bool small, bool media, int blocks, cudaStream_t st);
cudaError_t launch_mega_fast(const RenderArgs &A, bool smem, bool small, bool media, int blocks, cudaStream_t st);
// persistent wavefront variant (`smem`: scene tables staged in shared memory, else read through L1)
// `trace` (FAST only): SAH-BVH scenes, rays handed out dynamically inside the extend phase
cudaError_t wave_occupancy_parity(const RenderArgs &A, bool small, bool smem, bool media, bool trace, int *blocks_per_sm);
cudaError_t wave_occupancy_fast(const RenderArgs &A, bool small, bool smem, bool media, bool trace, int *blocks_per_sm);
cudaError_t launch_wave_parity(const RenderArgs &A, bool small, bool smem, bool media, bool trace, int blocks, cudaStream_t st);
cudaError_t launch_wave_fast(const RenderArgs &A, bool small, bool smem, bool media, bool trace, int blocks, cudaStream_t st);
// threads per CTA of the variant the arguments select (statistics)
int wave_threads_parity(const RenderArgs &A, bool small, bool smem, bool media, bool trace);
int wave_threads_fast(const RenderArgs &A, bool small, bool smem, bool media, bool trace);
cudaError_t launch_texture_probe_parity(const TextureProbeArgs &A, cudaStream_t st);
cudaError_t launch_texture_probe_fast(const TextureProbeArgs &A, cudaStream_t st);
// iters x 64 FMAs per thread, 256 threads per block (tpt_debug_fp32_peak)
cudaError_t launch_fp32_peak_probe(int blocks, int iters, float *sink, cudaStream_t st);
cudaError_t launch_philox_probe(const uint32_t ctr[4], const uint32_t key[2], uint32_t *d_out, cudaStream_t st);

} // namespace tptd
