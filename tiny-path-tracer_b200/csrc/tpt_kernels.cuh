// csrc/tpt_kernels.cuh -- kernels built on tpt_device.cuh. Included by exactly two translation
// units: tpt_render_parity.cu (-fmad=false, IEEE div/sqrt) and tpt_render_fast.cu
// (-use_fast_math); each instantiates the kernels for its own PAR value and exports the
// launchers declared in tpt_launch.h.
#pragma once
#include "tpt_device.cuh"
#include "tpt_launch.h"

namespace tptd {

// ------------------------------------------------------------------------------------------
// scene blob -> shared memory
// ------------------------------------------------------------------------------------------
template <bool SMEM>
TPT_DEV const float4 *stage_scene(const SceneLayout &L, float4 *sblob) {
  if (!SMEM) return L.blob_global;
  for (int i = threadIdx.x; i < L.blob_words; i += blockDim.x) sblob[i] = __ldg(L.blob_global + i);
  __syncthreads();
  return sblob;
}

// ------------------------------------------------------------------------------------------
// Gate-1 kernel: world->hit(r, tmin, tmax, rec) for a batch of rays
// ------------------------------------------------------------------------------------------
template <bool PAR, bool SMEM, bool SMALL>
__global__ void __launch_bounds__(128) intersect_kernel(const __grid_constant__ IntersectArgs A) {
  extern __shared__ float4 sblob[];
  SceneView S;
  S.blob = stage_scene<SMEM>(A.scene, sblob);
  S.wide_loads = !SMEM;
  S.L = &A.scene;
  S.small = &A.small;
  S.flat = &A.flat;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= A.n) return;
  const float *q = A.rays + 7 * idx;
  Ray r;
  r.o = mk(q[0], q[1], q[2]);
  r.d = mk(q[3], q[4], q[5]);
  r.time = q[6];
  tpt_hit out;
  out.hit = 0;
  out.prim = -1;
  out.mat = -1;
  out.t = out.u = out.v = 0.f;
  out.p[0] = out.p[1] = out.p[2] = 0.f;
  out.n[0] = out.n[1] = out.n[2] = 0.f;
  float t;
  int prim;
  bool any_hit;
  if constexpr (SMALL && PAR) any_hit = closest_hit_flat(S, r, A.tmin, A.tmax, t, prim);
  else if constexpr (SMALL && !PAR) any_hit = closest_hit_uniform(S, r, A.tmin, A.tmax, t, prim);
  else if (!PAR && A.scene.n_fbvh > 0) any_hit = closest_hit_fbvh(S, r, A.tmin, A.tmax, t, prim);
  else any_hit = closest_hit<PAR>(S, r, A.tmin, A.tmax, t, prim, nullptr);
  if (any_hit) {
    HitRec h;
    fill_hit<PAR>(S, r, prim, t, true, h);
    out.hit = 1;
    out.prim = prim;
    out.mat = h.mat;
    out.t = h.t;
    out.u = h.u;
    out.v = h.v;
    out.p[0] = h.p.x;
    out.p[1] = h.p.y;
    out.p[2] = h.p.z;
    out.n[0] = h.n.x;
    out.n[1] = h.n.y;
    out.n[2] = h.n.z;
  }
  A.out[idx] = out;
}

// One (pixel, sample range) bin per lane from the global work counter, warp-aggregated: the lanes
// that `need` one share a single atomicAdd. A finished bin is stored first (one store per bin).
struct LaneBin {
  bool have_bin, exhausted;
  int px, py, k, k_end;
  unsigned acc_index;
  V3 acc;
};
// tile owned by this launch: part `part_index` of `part_count` (tiles t with t % part_count == part_index), or --
// tpt_render_multi's static share in ONE launch -- the union of `part_group` such parts, `part_stride` apart
TPT_DEV unsigned part_tile(const RenderArgs &A, unsigned tile_local) {
  if (A.part_group <= 1) return (unsigned)A.part_index + tile_local * (unsigned)A.part_count;
  const unsigned q = tile_local / (unsigned)A.part_group, r = tile_local - q * (unsigned)A.part_group;
  return (unsigned)A.part_index + r * (unsigned)A.part_stride + q * (unsigned)A.part_count;
}
// no ray of pixel (px, py) can reach the scene's bounds (RenderArgs::cull_*, computed in make_plan)
TPT_DEV bool pixel_bundle_misses(const RenderArgs &A, int px, int py) {
  return px < A.cull_x0 || px > A.cull_x1 || py < A.cull_y0 || py > A.cull_y1;
}

template <bool CULL>
TPT_DEV void lane_bin_refill(const RenderArgs &A, LaneBin &B, bool need, unsigned lane, unsigned bins_per_tile,
                             unsigned long long &n_paths, unsigned long long &n_culled) {
  const unsigned FULL = 0xffffffffu;
  if (need && B.have_bin) { // bin finished: one store per (pixel, range)
    float *o = A.acc + (size_t)B.acc_index * 3;
    o[0] = B.acc.x;
    o[1] = B.acc.y;
    o[2] = B.acc.z;
    B.have_bin = false;
  }
  for (;;) { // until every lane that needs a bin holds one it has to trace, or the counter is dry
    unsigned m = __ballot_sync(FULL, need);
    if (!m) break;
    unsigned long long base = 0;
    int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(A.counters + 0, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (need) {
      unsigned long long b = base + __popc(m & ((1u << lane) - 1u));
      if (b >= A.n_bins) {
        B.exhausted = true;
        need = false;
      } else {
        unsigned bb = (unsigned)b;
        unsigned tile_local = bb / bins_per_tile;
        unsigned rem = bb - tile_local * bins_per_tile;
        unsigned range = rem / (unsigned)(TPT_TILE * TPT_TILE);
        unsigned pit = rem - range * (unsigned)(TPT_TILE * TPT_TILE);
        unsigned tile = part_tile(A, tile_local);
        unsigned ty = tile / (unsigned)A.tiles_x, tx = tile - ty * (unsigned)A.tiles_x;
        B.px = (int)(tx * TPT_TILE + (pit & (TPT_TILE - 1)));
        B.py = (int)(ty * TPT_TILE + (pit / TPT_TILE));
        if (B.px < A.nx && B.py < A.ny) {
          B.k = A.range_bounds[range];
          B.k_end = A.range_bounds[range + 1];
          B.acc_index = range * (unsigned)(A.nx * A.ny) + (unsigned)(B.py * A.nx + B.px);
          if (CULL && pixel_bundle_misses(A, B.px, B.py)) {
            float *o = A.acc + (size_t)B.acc_index * 3; // the bin's sum is exactly 0: finished
            o[0] = 0.f;
            o[1] = 0.f;
            o[2] = 0.f;
            n_paths += (unsigned long long)(B.k_end - B.k);
            n_culled += (unsigned long long)(B.k_end - B.k);
          } else {
            B.have_bin = true;
            B.acc = mk(0, 0, 0);
            need = false;
          }
        }
      }
    }
    if (!CULL) break; // without the bundle test a lane asks once per outer iteration (out-of-frame pixels of a partial tile: it asks again next time)
  }
}

// ------------------------------------------------------------------------------------------
// Persistent megakernel with per-lane path regeneration.
//
// Work unit ("bin") = (pixel, sample range). A lane owns one bin at a time and runs its samples
// sequentially -- so the per-pixel accumulation `col += de_nan(sample)` happens in the
// reference's order (main.cpp:119-126) in a register, and is stored once. When a lane's path
// ends it immediately starts its next sample (or grabs the next bin from the global
// work counter, warp-aggregated), so a warp stays full regardless of path length: 64 % of the
// headline frame's paths are single-ray misses while others bounce 15+ times.
// Bins are ordered tile-major with the pixel index fastest: the 32 lanes of a warp work on 32
// neighbouring pixels of one tile.
// ------------------------------------------------------------------------------------------
template <bool PAR, bool SMEM, bool SMALL, bool MEDIA, bool CULL, int TEX = TPT_TEXF_ALL>
__global__ void __launch_bounds__(TPT_MEGA_THREADS) render_mega_kernel(const __grid_constant__ RenderArgs A) {
  extern __shared__ float4 sblob[];
  SceneView S;
  S.blob = stage_scene<SMEM>(A.scene, sblob);
  S.wide_loads = !SMEM;
  S.L = &A.scene;
  S.small = &A.small;
  S.flat = &A.flat;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;

  bool active = false;
  LaneBin B;
  B.have_bin = false;
  B.exhausted = false;
  B.px = B.py = B.k = B.k_end = 0;
  B.acc_index = 0;
  B.acc = mk(0, 0, 0);
  PathState ps;
  Rng rng;
  unsigned long long n_rays = 0, n_nan = 0, n_paths = 0, n_culled = 0;
  const unsigned bins_per_tile = (unsigned)(TPT_TILE * TPT_TILE) * (unsigned)A.n_ranges;

  for (;;) {
    lane_bin_refill<CULL>(A, B, !active && !B.exhausted && (!B.have_bin || B.k >= B.k_end), lane, bins_per_tile, n_paths, n_culled);
    if (!active && B.have_bin && B.k < B.k_end) { // next sample of this lane's pixel
      rng.begin(A.rk, (uint32_t)(B.py * A.nx + B.px), (uint32_t)B.k);
      ps.ray = camera_sample<PAR>(A.cam, B.px, B.py, A.nx, A.ny, rng);
      ps.T = mk(1.f, 1.f, 1.f);
      ps.depth = 0;
      active = true;
      n_paths++;
    }
    if (__all_sync(FULL, !active && B.exhausted)) break;
    if (active) {
      V3 rad;
      n_rays++;
      if (!bounce<PAR, SMALL, MEDIA, TEX>(S, ps, rng, A.max_depth, A.t_min, rad)) {
        // col += de_nan(tmp): main.cpp:126, headers/utils.h:100-109
        bool nan_any = isnan(rad.x) || isnan(rad.y) || isnan(rad.z) || isnan(ps.T.x) ||
                       isnan(ps.T.y) || isnan(ps.T.z);
        if (nan_any) n_nan++;
        B.acc.x += isnan(rad.x) ? 0.f : rad.x;
        B.acc.y += isnan(rad.y) ? 0.f : rad.y;
        B.acc.z += isnan(rad.z) ? 0.f : rad.z;
        active = false;
        B.k++;
      }
    }
  }
  // statistics: one atomic per warp and counter
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_rays += __shfl_xor_sync(FULL, n_rays, o);
    n_nan += __shfl_xor_sync(FULL, n_nan, o);
    n_paths += __shfl_xor_sync(FULL, n_paths, o);
    n_culled += __shfl_xor_sync(FULL, n_culled, o);
  }
  if (lane == 0) {
    atomicAdd(A.counters + 1, n_rays);
    atomicAdd(A.counters + 2, n_nan);
    atomicAdd(A.counters + 3, n_paths);
    if (n_culled) atomicAdd(A.counters + 4, n_culled);
  }
}

// ------------------------------------------------------------------------------------------
// Persistent WAVEFRONT kernel: generate / extend / shade queues in shared memory.
//
// A CTA owns wave_slots() path slots (two or three per thread, tpt_launch.h) whose state -- ray, throughput, per-pixel
// accumulator, sample bookkeeping, pending hit -- lives in shared memory as structure-of-arrays.
// Every iteration has two phases separated by __syncthreads() (three in the parity kernels, where
// generate runs after shade as a phase of its own, see SPLIT_GEN below):
//   extend          each thread intersects its slots (warp-uniform brute force on small scenes).
//                   Paths that end here (miss, lamp, absorber, depth limit) add their radiance to
//                   the slot's accumulator and go to the GENERATE queue, the others to the queue of
//                   their material. Queues are compacted with __ballot_sync/__popc and one or two
//                   packed 32-bit shared-memory atomics per warp.
//   shade+generate  warps pull 32-item chunks from the queues: a warp shades 32 lambertian (or 32
//                   dielectric, or 32 metal) hits together instead of diverging over the material
//                   switch, or finishes 32 samples together (k++, store the bin when its range is
//                   complete, grab the next bin from the global work counter, shoot the next camera
//                   ray). Paths that die inside scatter() are queued for the next iteration's
//                   generate step (queue counters are double-buffered by iteration parity).
// Same device functions, same Philox stream and same per-pixel summation order as the
// megakernel: the two variants produce bit-identical images and differ only in scheduling.
//
// Evaluated alternative (profiles/r01_wave_warp_autonomous_rejected_metrics.csv): one autonomous
// queue scheduler per WARP, no block barriers. Its warps drift into different stages, the SM's
// instruction working set becomes the whole kernel (~80 KB) and the "no instruction" stall grows
// from 1.6 to 7.2 cycles per issued instruction; it ran 35 % slower than this lock-step form.
// ------------------------------------------------------------------------------------------
#define TPT_WAVE_NQ 4 // queues: 0 lambertian, 1 metal, 2 dielectric (= TPT_MAT_*), 3 generate
__host__ __device__ constexpr int wave_state_words(bool media) { return media ? 21 : 20; } // 32-bit words of path state per slot

template <bool PAR, bool SMALL, bool SMEM, bool MEDIA, bool CULL, bool TRACE = false, int TEX = TPT_TEXF_ALL>
__global__ void __launch_bounds__(wave_threads(SMALL, TRACE), wave_min_blocks(PAR, SMALL, TRACE, TEX == 0))
render_wave_kernel(const __grid_constant__ RenderArgs A) {
  // shade and generate as ONE phase (a warp takes material chunks and generate chunks from one list:
  // better balance, one barrier fewer) or as two (each phase's code stays hot in the instruction
  // cache). Measured on the Cornell frame: the split form won by 8 % while the kernel carried the
  // texture code it never runs; with the lean build the merged form wins by 6 % (fast mode). The
  // parity kernels are several times larger and keep the split (merged: -6 %).
  // r02: the small-scene parity kernels (flat replay of the reference tree, lean build) are small enough
  // for the merged form as well: +4.6 % / +5.5 % (Cornell A / B).
  constexpr bool SPLIT_GEN = TPT_WAVE_SPLIT_GEN == 1 || (TPT_WAVE_SPLIT_GEN == 2 && PAR && !SMALL);
  extern __shared__ float4 sblob[];
  constexpr int NSLOT = wave_slots(PAR, SMALL, TRACE, TEX == 0);
  constexpr int THREADS = wave_threads(SMALL, TRACE);
  constexpr int NWARP = THREADS / 32;
  // structure-of-arrays slot state: field f of slot s at sf[f * NSLOT + s]
  // F_DEPTH < 0 marks a slot without a live path; F_NDRAW (draws a medium took inside world->hit)
  // exists only in the media builds. 20 words per slot, 96 B with the queues: 48 KB per 512-slot CTA.
  enum { F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_TIME, F_TX, F_TY, F_TZ, F_AX, F_AY, F_AZ,
         F_PIXEL, F_K, F_KEND, F_ACCIDX, F_DEPTH, F_HPRIM, F_HT, F_NDRAW };
  constexpr int F_COUNT = wave_state_words(MEDIA);
  // dynamic shared memory: [slot state | queues | scene blob]. State and queues sit at
  // compile-time offsets (plain LDS/STS with immediate offsets); the blob, whose size is only known
  // at run time, comes last.
  constexpr int STATE_BYTES = F_COUNT * NSLOT * 4, QUEUE_BYTES = 2 * TPT_WAVE_NQ * NSLOT * 2;
  static_assert((STATE_BYTES + QUEUE_BYTES) % 16 == 0, "scene blob must stay 16-byte aligned");
  float *sf = reinterpret_cast<float *>(sblob);
  int *si = reinterpret_cast<int *>(sblob);
  // queue[parity][q][NSLOT] slot ids; counters packed 2 x 16 bit in two 32-bit words per parity
  // (q_cnt): word 0 = lambertian | generate << 16 (almost every warp feeds both), word 1 = metal |
  // dielectric << 16 (touched only by warps that saw one). 32-bit shared atomics are native
  // (ATOMS.ADD); the 64-bit add this replaces compiled to a compare-and-swap spin loop
  // (ATOMS.CAST.SPIN.64) that eight warps contended for.
  unsigned short *queue = reinterpret_cast<unsigned short *>(reinterpret_cast<char *>(sblob) + STATE_BYTES);
  SceneView S;
  S.blob = stage_scene<SMEM>(A.scene, sblob + (STATE_BYTES + QUEUE_BYTES) / 16); // > 64 KB: stays in global memory
  S.wide_loads = !SMEM;
  S.L = &A.scene;
  S.small = &A.small;
  S.flat = &A.flat;
  __shared__ unsigned q_cnt[2][2];
  __shared__ int n_idle;
  __shared__ int trace_next; // TRACE: next slot whose ray nobody has taken yet
  __shared__ int next_chunk; // merged shade+generate phase: next 32-item chunk nobody has taken yet
  // Merged phase: warps take chunks from one list through a shared counter, the expensive ones
  // first (lambertian, dielectric, metal, then the many short generate chunks as fillers) -- longest
  // processing time first -- instead of a fixed round-robin share.
  constexpr bool DYNAMIC = !SPLIT_GEN && TPT_WAVE_DYNAMIC;
#define SF(f, s) sf[(f) * NSLOT + (s)]
#define SI(f, s) si[(f) * NSLOT + (s)]
#define QUEUE(par, q) (queue + ((par) * TPT_WAVE_NQ + (q)) * NSLOT)

  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int warp = tid >> 5;
  const unsigned bins_per_tile = (unsigned)(TPT_TILE * TPT_TILE) * (unsigned)A.n_ranges;
  unsigned long long n_rays = 0, n_nan = 0, n_paths = 0, n_culled = 0;

  int stack[TRACE ? TPT_FBVH_STACK + 3 : 1]; // TRACE: the lane's node stack + a parked walk (slot, node, stack pointer)
  stack[TRACE ? TPT_FBVH_STACK : 0] = -1;
  for (int s = tid; s < NSLOT; s += THREADS) {
    SI(F_DEPTH, s) = -1;
    SI(F_K, s) = 0;
    SI(F_KEND, s) = -1; // no bin yet
    QUEUE(0, 3)[s] = (unsigned short)s;
  }
  if (tid == 0) {
    q_cnt[0][0] = (unsigned)NSLOT << 16; // everything starts in GENERATE
    q_cnt[0][1] = 0;
    q_cnt[1][0] = 0;
    q_cnt[1][1] = 0;
    n_idle = 0;
    trace_next = 0;
    next_chunk = 0;
  }
  __syncthreads();

  for (int par = 0;; par ^= 1) {
    // ------------------------------------------------------------------ extend
    if (tid == 0) { // next iteration's counters (nobody touches them before the barrier)
      q_cnt[par ^ 1][0] = 0;
      q_cnt[par ^ 1][1] = 0;
      next_chunk = 0;
    }
    if (TRACE) {
      // Closest hits through the SAH BVH, rays handed out dynamically: a lane that finishes its ray
      // takes the next untraced slot from the CTA-wide counter instead of idling until the longest
      // traversal of its warp ends. Results go to F_HT / F_HPRIM; the rest of extend() follows below.
      // Measured (BVH scenes): the hand-out itself is worth 2-4 %; the variant's gain on
      // oneweek_final (+15 %) comes from running 2 CTAs/SM at 128 registers (no spills around the
      // media pass). random_scene is faster on the plain 3-CTA form, so only media scenes come here.
      //
      // TPT_TRACE_CARRY: a phase still ends with its longest rays, walked by a few lanes per warp while
      // everybody else waits at the barrier (ray lengths spread over an order of magnitude, and a CTA
      // has only three rays per lane to hand out). With the carry-over a warp LEAVES the phase once the
      // counter is dry and fewer than TPT_TRACE_CARRY_K of its lanes still walk: those lanes park their
      // walk (node, stack pointer and slot number in the tail of the lane's stack array, best hit so far
      // in the slot) and resume it in the next iteration's extend phase, next to fresh rays. A parked
      // slot carries TPT_SLOT_PARKED in its depth word and sits in no queue; TPT_SLOT_LATE marks one
      // that finished in this phase, so that the hand-out does not give it away a second time.
      FbvhTrav tv;
      tv.node = TPT_FBVH_DONE;
      int ts = -1;
      Ray tr;
      tr.o = tr.d = mk(0, 0, 0);
      tr.time = 0.f;
      bool dry = false, got_fresh = false;
      if (TPT_TRACE_CARRY) {
        ts = stack[TPT_FBVH_STACK];
        if (ts >= 0) { // resume the parked walk
          tr.o = mk(SF(F_OX, ts), SF(F_OY, ts), SF(F_OZ, ts));
          tr.d = mk(SF(F_DX, ts), SF(F_DY, ts), SF(F_DZ, ts));
          tr.time = SF(F_TIME, ts);
          tv.start(tr, SF(F_HT, ts));
          tv.best_prim = SI(F_HPRIM, ts);
          tv.node = stack[TPT_FBVH_STACK + 1];
          tv.sp = stack[TPT_FBVH_STACK + 2];
        }
      }
      for (;;) {
        const bool need = tv.done();
        if (need && ts >= 0) {
          SI(F_HPRIM, ts) = tv.best_prim;
          SF(F_HT, ts) = tv.best;
          if (TPT_TRACE_CARRY) {
            const int d = SI(F_DEPTH, ts);
            if (d & TPT_SLOT_PARKED) SI(F_DEPTH, ts) = (d ^ TPT_SLOT_PARKED) | TPT_SLOT_LATE;
          }
          ts = -1;
        }
        const unsigned m = __ballot_sync(FULL, need);
        if (m && !dry) {
          int base = 0;
          const int leader = __ffs(m) - 1;
          if ((int)lane == leader) base = atomicAdd(&trace_next, __popc(m));
          base = __shfl_sync(FULL, base, leader);
          if (base >= NSLOT) dry = true;
          bool took = false;
          if (need) {
            const int c = base + __popc(m & lt_mask);
            if (c < NSLOT && (unsigned)SI(F_DEPTH, c) < (unsigned)TPT_SLOT_PARKED) {
              ts = c;
              took = true;
              tr.o = mk(SF(F_OX, c), SF(F_OY, c), SF(F_OZ, c));
              tr.d = mk(SF(F_DX, c), SF(F_DY, c), SF(F_DZ, c));
              tr.time = SF(F_TIME, c);
              tv.start(tr, FLT_MAX);
            }
          }
          if (TPT_TRACE_CARRY && !got_fresh) got_fresh = __ballot_sync(FULL, took) != 0u;
        }
        if (dry) {
          const unsigned walking = __ballot_sync(FULL, !tv.done());
          // (a warp that received no fresh ray in this phase finishes what it has: the end of the launch)
          if (walking == 0u || (TPT_TRACE_CARRY && got_fresh && __popc(walking) < TPT_TRACE_CARRY_K)) break;
        }
#if TPT_TRACE_VOTE
        // vote-scheduled walk (closest_hit_fbvh_vote) between the hand-outs: lanes that reached a leaf wait
        // while more than a share of the warp's walking lanes still descend, then the leaf tests run
        // together; the loop returns to the hand-out above after every leaf pass
        {
          const float4 *N = S.blob + S.L->off_fbvh;
          const int walking = __popc(__ballot_sync(FULL, !tv.done()));
          const int limit = walking * TPT_VOTE_NUM / (TPT_VOTE_NUM + TPT_VOTE_DEN);
          bool inner = (unsigned)tv.node < (unsigned)TPT_FBVH_DONE;
          while (__popc(__ballot_sync(FULL, inner)) > limit) {
            if (inner) tv.inner_one(N, A.t_min, stack, S.wide_loads);
            inner = (unsigned)tv.node < (unsigned)TPT_FBVH_DONE;
          }
          if (tv.node < 0) tv.leaf(S, tr, A.t_min, stack);
        }
#else
        if (!tv.done()) tv.step(S, tr, A.t_min, stack);
#endif
      }
      if (TPT_TRACE_CARRY) {
        if (ts >= 0) { // park the walk
          SI(F_DEPTH, ts) |= TPT_SLOT_PARKED;
          SF(F_HT, ts) = tv.best;
          SI(F_HPRIM, ts) = tv.best_prim;
          stack[TPT_FBVH_STACK + 1] = tv.node;
          stack[TPT_FBVH_STACK + 2] = tv.sp;
        }
        stack[TPT_FBVH_STACK] = ts;
      }
      __syncthreads();
    }
    // (r02: reserving queue space once per warp for all its slot groups instead of once per group -- one pair of
    // shared atomics instead of up to three -- was measured and dropped: the seven words carried across the
    // intersector spill, -1.1 % fast / -2.4 % parity)
    for (int s0 = warp * 32; s0 < NSLOT; s0 += THREADS) {
      const int s = s0 + (int)lane;
      int cls = -2; // -2: nothing to do
      int depth_s = SI(F_DEPTH, s);
      if (TRACE && TPT_TRACE_CARRY) {
        if (depth_s >= 0 && (depth_s & TPT_SLOT_PARKED)) depth_s = -1; // its walk goes on in the next iteration
        else if (depth_s >= 0 && (depth_s & TPT_SLOT_LATE)) SI(F_DEPTH, s) = depth_s ^= TPT_SLOT_LATE;
      }
      // SAH BVH scenes: the closest hit is found first, by all 32 lanes together (vote-scheduled walk)
      constexpr bool VOTE = TPT_FBVH_VOTE && !PAR && !SMALL && !TRACE;
      const bool voted = VOTE && A.scene.n_fbvh > 0 && A.scene.fbvh_time_ok;
      float t_v = 0.f;
      int prim_v = -1;
      if (voted) {
        Ray vr;
        vr.o = mk(SF(F_OX, s), SF(F_OY, s), SF(F_OZ, s));
        vr.d = mk(SF(F_DX, s), SF(F_DY, s), SF(F_DZ, s));
        vr.time = SF(F_TIME, s);
        closest_hit_fbvh_vote(S, vr, depth_s >= 0, A.t_min, FLT_MAX, t_v, prim_v);
      }
      if (depth_s >= 0) {
        PathState ps;
        ps.ray.o = mk(SF(F_OX, s), SF(F_OY, s), SF(F_OZ, s));
        ps.ray.d = mk(SF(F_DX, s), SF(F_DY, s), SF(F_DZ, s));
        ps.ray.time = SF(F_TIME, s);
        ps.T = mk(SF(F_TX, s), SF(F_TY, s), SF(F_TZ, s));
        ps.depth = depth_s;
        float t;
        int prim;
        V3 rad;
        n_rays++;
        Rng rng; // only participating media draw during extend
        const int pk = SI(F_PIXEL, s);
        rng.begin(A.rk, (uint32_t)((pk >> 16) * A.nx + (pk & 0xffff)), (uint32_t)SI(F_K, s));
        uint32_t ndraw0;
        if (TRACE) {
          t = SF(F_HT, s);
          prim = SI(F_HPRIM, s);
          if (MEDIA) rng.set_stage((uint32_t)ps.depth + 1u);
          cls = extend_finish<PAR, MEDIA, TEX>(S, ps, A.max_depth, A.t_min, prim >= 0, t, prim, rad, rng, ndraw0);
        } else if (voted) {
          t = t_v;
          prim = prim_v;
          if (MEDIA) rng.set_stage((uint32_t)ps.depth + 1u);
          cls = extend_finish<PAR, MEDIA, TEX>(S, ps, A.max_depth, A.t_min, prim >= 0, t, prim, rad, rng, ndraw0);
        } else {
          cls = extend<PAR, SMALL, MEDIA, TEX>(S, ps, A.max_depth, A.t_min, t, prim, rad, rng, ndraw0);
        }
        if (cls == TPT_EXT_DONE) {
          // col += de_nan(tmp): main.cpp:126, headers/utils.h:100-109
          if (isnan(rad.x) || isnan(rad.y) || isnan(rad.z)) n_nan++;
          SF(F_AX, s) += isnan(rad.x) ? 0.f : rad.x;
          SF(F_AY, s) += isnan(rad.y) ? 0.f : rad.y;
          SF(F_AZ, s) += isnan(rad.z) ? 0.f : rad.z;
          SI(F_DEPTH, s) = -1;
        } else {
          SI(F_HPRIM, s) = prim;
          SF(F_HT, s) = t;
          if (MEDIA) SI(F_NDRAW, s) = (int)ndraw0;
        }
      }
      // queue of this slot: TPT_MAT_LAMBERTIAN=0, METAL=1, DIELECTRIC=2, done -> 3 (generate)
      const int q = cls == TPT_EXT_DONE ? 3 : cls;
      const unsigned m0 = __ballot_sync(FULL, q == 0), m1 = __ballot_sync(FULL, q == 1);
      const unsigned m2 = __ballot_sync(FULL, q == 2), m3 = __ballot_sync(FULL, q == 3);
      if (m0 | m1 | m2 | m3) {
        const unsigned add_a = (unsigned)__popc(m0) | ((unsigned)__popc(m3) << 16);
        const unsigned add_b = (unsigned)__popc(m1) | ((unsigned)__popc(m2) << 16);
        unsigned base_a = 0, base_b = 0;
        if (lane == 0) {
          if (add_a) base_a = atomicAdd(&q_cnt[par][0], add_a);
          if (add_b) base_b = atomicAdd(&q_cnt[par][1], add_b);
        }
        base_a = __shfl_sync(FULL, base_a, 0);
        base_b = __shfl_sync(FULL, base_b, 0);
        if (q >= 0) {
          const unsigned mine = q == 0 ? m0 : q == 1 ? m1 : q == 2 ? m2 : m3;
          const unsigned base = (q == 0 || q == 3) ? base_a : base_b;
          const int at = (int)((base >> ((q >= 2) ? 16 : 0)) & 0xffffu) + __popc(mine & lt_mask);
          QUEUE(par, q)[at] = (unsigned short)s;
        }
      }
    }
    __syncthreads();
    // -------------------------------------------------------- shade + generate
    {
      if (TRACE && tid == 0) trace_next = 0;
      const unsigned word_a = q_cnt[par][0], word_b = q_cnt[par][1];
      const int c0 = (int)(word_a & 0xffffu), c3 = (int)(word_a >> 16);
      const int c1 = (int)(word_b & 0xffffu), c2 = (int)(word_b >> 16);
      const int t0 = (c0 + 31) >> 5, t1 = (c1 + 31) >> 5, t2 = (c2 + 31) >> 5, t3 = (c3 + 31) >> 5;
      // static share (parity kernels): GENERATE chunks first (the most numerous and the most uniform),
      // then the materials; dynamic hand-out (fast kernels): materials first, generate chunks as fillers
      for (int pass = 0; pass < (SPLIT_GEN ? 2 : 1); pass++) {
      if (SPLIT_GEN && pass == 1) __syncthreads();
      const int lo_t = !SPLIT_GEN ? 0 : (pass == 0 ? t3 : 0), hi_t = !SPLIT_GEN ? t0 + t1 + t2 + t3 : (pass == 0 ? t0 + t1 + t2 + t3 : t3);
      auto take_chunk = [&]() {
        int v = 0;
        if (lane == 0) v = atomicAdd(&next_chunk, 1);
        return __shfl_sync(FULL, v, 0);
      };
      for (int wt = DYNAMIC ? take_chunk() : lo_t + warp; wt < hi_t; wt = DYNAMIC ? take_chunk() : wt + NWARP) {
        int q, chunk, cnt;
        if (DYNAMIC) {
          if (wt < t0) { q = 0; chunk = wt; cnt = c0; }
          else if (wt < t0 + t2) { q = 2; chunk = wt - t0; cnt = c2; }
          else if (wt < t0 + t2 + t1) { q = 1; chunk = wt - t0 - t2; cnt = c1; }
          else { q = 3; chunk = wt - t0 - t2 - t1; cnt = c3; }
        } else if (wt < t3) { q = 3; chunk = wt; cnt = c3; }
        else if (wt < t3 + t0) { q = 0; chunk = wt - t3; cnt = c0; }
        else if (wt < t3 + t0 + t2) { q = 2; chunk = wt - t3 - t0; cnt = c2; }
        else { q = 1; chunk = wt - t3 - t0 - t2; cnt = c1; }
        const int idx = chunk * 32 + (int)lane;
        const bool mine = idx < cnt;
        const int s = mine ? QUEUE(par, q)[idx] : 0;
        if (q != 3) {
          // ------------------------------------------------------------- shade
          bool died = false;
          if (mine) {
            PathState ps;
            ps.ray.o = mk(SF(F_OX, s), SF(F_OY, s), SF(F_OZ, s));
            ps.ray.d = mk(SF(F_DX, s), SF(F_DY, s), SF(F_DZ, s));
            ps.ray.time = SF(F_TIME, s);
            ps.T = mk(SF(F_TX, s), SF(F_TY, s), SF(F_TZ, s));
            ps.depth = SI(F_DEPTH, s);
            Rng rng;
            const int pk = SI(F_PIXEL, s);
            rng.begin(A.rk, (uint32_t)((pk >> 16) * A.nx + (pk & 0xffff)), (uint32_t)SI(F_K, s));
            bool alive = shade<PAR, TEX>(S, ps, rng, SI(F_HPRIM, s), SF(F_HT, s), MEDIA ? (uint32_t)SI(F_NDRAW, s) : 0u);
            if (alive) {
              SF(F_OX, s) = ps.ray.o.x; SF(F_OY, s) = ps.ray.o.y; SF(F_OZ, s) = ps.ray.o.z;
              SF(F_DX, s) = ps.ray.d.x; SF(F_DY, s) = ps.ray.d.y; SF(F_DZ, s) = ps.ray.d.z;
              SF(F_TX, s) = ps.T.x; SF(F_TY, s) = ps.T.y; SF(F_TZ, s) = ps.T.z;
              SI(F_DEPTH, s) = ps.depth;
            } else {
              died = true; // contributes 0 (metal absorbed, or throughput 0 / NaN in every channel)
              SI(F_DEPTH, s) = -1;
              if (isnan(ps.T.x) || isnan(ps.T.y) || isnan(ps.T.z)) n_nan++;
            }
          }
          // dead paths regenerate in the NEXT iteration: push into the other parity's GENERATE queue
          const unsigned md = __ballot_sync(FULL, died);
          if (md) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&q_cnt[par ^ 1][0], (unsigned)__popc(md) << 16);
            base = __shfl_sync(FULL, base, 0);
            if (died) QUEUE(par ^ 1, 3)[(int)(base >> 16) + __popc(md & lt_mask)] = (unsigned short)s;
          }
        } else {
          // ---------------------------------------------------------- generate
          int k = 0, k_end = -1, pixel = 0;
          bool have_bin = false;
          if (mine) {
            k_end = SI(F_KEND, s);
            have_bin = k_end >= 0;
            k = SI(F_K, s) + (have_bin ? 1 : 0); // the sample that just ended
            pixel = SI(F_PIXEL, s);
            if (have_bin && k >= k_end) { // bin finished: one store per (pixel, sample range)
              float *o = A.acc + (size_t)(unsigned)SI(F_ACCIDX, s) * 3;
              o[0] = SF(F_AX, s);
              o[1] = SF(F_AY, s);
              o[2] = SF(F_AZ, s);
              have_bin = false;
            }
          }
          bool need = mine && !have_bin;
          bool exhausted = false;
          for (;;) { // grab bins until every lane that needs one has a valid pixel or the counter is dry
            unsigned m = __ballot_sync(FULL, need);
            if (!m) break;
            int leader = __ffs(m) - 1;
            unsigned long long b0 = 0;
            if ((int)lane == leader) b0 = atomicAdd(A.counters + 0, (unsigned long long)__popc(m));
            b0 = __shfl_sync(FULL, b0, leader);
            if (need) {
              unsigned long long b = b0 + __popc(m & lt_mask);
              if (b >= A.n_bins) {
                exhausted = true;
                need = false;
              } else {
                unsigned bb = (unsigned)b;
                unsigned tile_local = bb / bins_per_tile;
                unsigned rem = bb - tile_local * bins_per_tile;
                unsigned range = rem / (unsigned)(TPT_TILE * TPT_TILE);
                unsigned pit = rem - range * (unsigned)(TPT_TILE * TPT_TILE);
                unsigned tile = part_tile(A, tile_local);
                unsigned ty = tile / (unsigned)A.tiles_x, tx = tile - ty * (unsigned)A.tiles_x;
                int px = (int)(tx * TPT_TILE + (pit & (TPT_TILE - 1)));
                int py = (int)(ty * TPT_TILE + (pit / TPT_TILE));
                if (px < A.nx && py < A.ny) {
                  const unsigned acc_index = range * (unsigned)(A.nx * A.ny) + (unsigned)(py * A.nx + px);
                  if (CULL && pixel_bundle_misses(A, px, py)) {
                    // no ray of this pixel reaches the scene: the bin's sum is exactly 0, take another
                    float *o = A.acc + (size_t)acc_index * 3;
                    o[0] = 0.f;
                    o[1] = 0.f;
                    o[2] = 0.f;
                    const int cnt = A.range_bounds[range + 1] - A.range_bounds[range];
                    n_paths += (unsigned long long)cnt;
                    n_culled += (unsigned long long)cnt;
                  } else {
                    need = false;
                    have_bin = true;
                    pixel = px | (py << 16); // packed: no division when the camera ray is generated
                    k = A.range_bounds[range];
                    k_end = A.range_bounds[range + 1];
                    SI(F_PIXEL, s) = pixel;
                    SI(F_KEND, s) = k_end;
                    SI(F_ACCIDX, s) = (int)acc_index;
                    SF(F_AX, s) = 0.f;
                    SF(F_AY, s) = 0.f;
                    SF(F_AZ, s) = 0.f;
                  }
                }
              }
            }
          }
          if (mine) {
            if (have_bin) {
              Rng rng;
              const int py = pixel >> 16, px = pixel & 0xffff;
              rng.begin(A.rk, (uint32_t)(py * A.nx + px), (uint32_t)k);
              Ray r = camera_sample<PAR>(A.cam, px, py, A.nx, A.ny, rng);
              // A camera ray that cannot hit the scene's bounds (64 % of the headline frame looks past
              // the box) ends its path right here -- world->hit is false for it -- and the lane draws
              // its pixel's next sample instead of spending an extend pass on it. Bounded retries:
              // the last ray of the budget (or of the bin) goes to extend like any other.
              if (!MEDIA && (!PAR || TPT_PAR_CAMERA_TRIES)) { // PARITY: the reference's own root-box test (may_hit_world<true>), same outcome as the extend pass it replaces
                for (int attempt = 1; attempt < A.camera_tries && k + 1 < k_end && !may_hit_world<PAR>(S, r, A.t_min); attempt++) {
                  V3 bg = background_radiance<PAR>(S, r, mk(1.f, 1.f, 1.f));
                  SF(F_AX, s) += isnan(bg.x) ? 0.f : bg.x; // col += de_nan(tmp), as in extend
                  SF(F_AY, s) += isnan(bg.y) ? 0.f : bg.y;
                  SF(F_AZ, s) += isnan(bg.z) ? 0.f : bg.z;
                  n_paths++;
                  n_rays++;
                  k++;
                  rng.begin(A.rk, (uint32_t)(py * A.nx + px), (uint32_t)k);
                  r = camera_sample<PAR>(A.cam, px, py, A.nx, A.ny, rng);
                }
              }
              SF(F_OX, s) = r.o.x; SF(F_OY, s) = r.o.y; SF(F_OZ, s) = r.o.z;
              SF(F_DX, s) = r.d.x; SF(F_DY, s) = r.d.y; SF(F_DZ, s) = r.d.z;
              SF(F_TIME, s) = r.time;
              SF(F_TX, s) = 1.f; SF(F_TY, s) = 1.f; SF(F_TZ, s) = 1.f;
              SI(F_DEPTH, s) = 0; // live
              SI(F_K, s) = k;
              n_paths++;
            } else if (exhausted) {
              SI(F_KEND, s) = -1;
              atomicAdd(&n_idle, 1);
            }
          }
        }
      }
      }
    }
    __syncthreads();
    if (n_idle >= NSLOT) break; // every slot idle: the work counter is dry and all paths ended
  }
#undef SF
#undef SI
#undef QUEUE
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_rays += __shfl_xor_sync(FULL, n_rays, o);
    n_nan += __shfl_xor_sync(FULL, n_nan, o);
    n_paths += __shfl_xor_sync(FULL, n_paths, o);
    n_culled += __shfl_xor_sync(FULL, n_culled, o);
  }
  if (lane == 0) {
    atomicAdd(A.counters + 1, n_rays);
    atomicAdd(A.counters + 2, n_nan);
    atomicAdd(A.counters + 3, n_paths);
    if (n_culled) atomicAdd(A.counters + 4, n_culled);
  }
}

// ------------------------------------------------------------------------------------------
// Known-answer probes
// ------------------------------------------------------------------------------------------
template <bool PAR> __global__ void texture_probe_kernel(const __grid_constant__ TextureProbeArgs A) {
  SceneView S;
  S.blob = A.scene.blob_global;
  S.wide_loads = true;
  S.L = &A.scene;
  S.small = nullptr;
  S.flat = nullptr;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= A.n) return;
  const float *q = A.uvp + 5 * idx;
  V3 c = texture_value<PAR>(S, A.texture, q[0], q[1], mk(q[2], q[3], q[4]));
  A.out[3 * idx] = c.x;
  A.out[3 * idx + 1] = c.y;
  A.out[3 * idx + 2] = c.z;
}

// ------------------------------------------------------------------------------------------
// Launchers (one set per translation unit; TPT_SUFFIX = parity | fast)
// ------------------------------------------------------------------------------------------
#define TPT_CAT2(a, b) a##b
#define TPT_CAT(a, b) TPT_CAT2(a, b)
#define TPT_FN(name) TPT_CAT(name, TPT_SUFFIX)

template <typename K> static cudaError_t allow_smem(K kernel, size_t bytes) {
  if (bytes + 1024 <= 48 * 1024) return cudaSuccess; // the 48 KB default covers static + dynamic shared memory together
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t TPT_FN(launch_intersect_)(const IntersectArgs &A, bool smem, bool small, cudaStream_t st) {
  size_t bytes = smem ? (size_t)A.scene.blob_words * 16 : 0;
  int blocks = (int)((A.n + 127) / 128);
  if (blocks == 0) return cudaSuccess;
  typedef void (*fn)(IntersectArgs);
  fn k = smem ? (small ? (fn)intersect_kernel<TPT_PAR, true, true> : (fn)intersect_kernel<TPT_PAR, true, false>)
              : (fn)intersect_kernel<TPT_PAR, false, false>;
  cudaError_t e;
  if (smem && (e = allow_smem(k, bytes)) != cudaSuccess) return e;
  k<<<blocks, 128, bytes, st>>>(A);
  return cudaGetLastError();
}

// kernel variant tables: [smem][small] or the media build (generic walk, scene staged when it
// fits), each with and without the pixel-bundle bounds test compiled in (RenderArgs::cull)
typedef void (*mega_fn)(RenderArgs);
template <bool CULL> static mega_fn mega_variant_c(bool smem, bool small, bool media, bool lean) {
#if !TPT_PAR
  if (lean && small && smem && !media) return render_mega_kernel<false, true, true, false, CULL, 0>;
#endif
  (void)lean;
  if (media) return smem ? render_mega_kernel<TPT_PAR, true, false, true, CULL> : render_mega_kernel<TPT_PAR, false, false, true, CULL>;
  if (smem) return small ? render_mega_kernel<TPT_PAR, true, true, false, CULL> : render_mega_kernel<TPT_PAR, true, false, false, CULL>;
  return render_mega_kernel<TPT_PAR, false, false, false, CULL>;
}
static mega_fn mega_variant(const RenderArgs &A, bool smem, bool small, bool media) {
  return A.cull ? mega_variant_c<true>(smem, small, media, A.tex_mask == 0) : mega_variant_c<false>(smem, small, media, A.tex_mask == 0);
}

cudaError_t TPT_FN(mega_occupancy_)(const RenderArgs &A, bool smem, bool small, bool media, size_t smem_bytes, int *blocks_per_sm) {
  mega_fn k = mega_variant(A, smem, small, media);
  cudaError_t e;
  if (smem && (e = allow_smem(k, smem_bytes)) != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, TPT_MEGA_THREADS, smem ? smem_bytes : 0);
}

cudaError_t TPT_FN(launch_mega_)(const RenderArgs &A, bool smem, bool small, bool media, int blocks, cudaStream_t st) {
  size_t bytes = smem ? (size_t)A.scene.blob_words * 16 : 0;
  mega_variant(A, smem, small, media)<<<blocks, TPT_MEGA_THREADS, bytes, st>>>(A);
  return cudaGetLastError();
}

typedef void (*wave_fn)(RenderArgs);
struct WaveVariant {
  wave_fn fn;
  int threads, slots;
};
template <bool SMALL, bool SMEM, bool MEDIA, bool CULL, bool TRACE, int TEX> static WaveVariant wave_pick() {
  return WaveVariant{render_wave_kernel<TPT_PAR, SMALL, SMEM, MEDIA, CULL, TRACE, TEX>, wave_threads(SMALL, TRACE),
                     wave_slots(TPT_PAR, SMALL, TRACE, TEX == 0)};
}
// `trace` (FAST only): closest hits through the library's SAH BVH with dynamic ray hand-out; the
// scene tables are read through L1 there and shared memory holds TPT_TRACE_SLOTS path slots.
template <bool CULL> static WaveVariant wave_variant_c(bool small, bool smem, bool media, bool trace, int tex) {
  constexpr int ALL = TPT_TEXF_ALL;
#if TPT_PAR && TPT_PAR_LEAN
  if (tex == 0 && small && smem && !media) return wave_pick<true, true, false, CULL, false, 0>();
#endif
#if !TPT_PAR
  // small scenes: one build per texture feature set (constant only | + image | + procedural | both)
  if (small && smem && !media && !trace) {
    if (tex == 0) return wave_pick<true, true, false, CULL, false, 0>();
    if (tex == TPT_TEXF_IMAGE) return wave_pick<true, true, false, CULL, false, TPT_TEXF_IMAGE>();
    if (tex == TPT_TEXF_PROCEDURAL) return wave_pick<true, true, false, CULL, false, TPT_TEXF_PROCEDURAL>();
    return wave_pick<true, true, false, CULL, false, ALL>();
  }
  if (trace) return media ? wave_pick<false, false, true, CULL, true, ALL>() : wave_pick<false, false, false, CULL, true, ALL>();
#endif
  (void)trace;
  (void)tex;
  if (media) return smem ? wave_pick<false, true, true, CULL, false, ALL>() : wave_pick<false, false, true, CULL, false, ALL>();
  if (!smem) return wave_pick<false, false, false, CULL, false, ALL>();
  return small ? wave_pick<true, true, false, CULL, false, ALL>() : wave_pick<false, true, false, CULL, false, ALL>();
}
static WaveVariant wave_variant(const RenderArgs &A, bool small, bool smem, bool media, bool trace) {
  return A.cull ? wave_variant_c<true>(small, smem, media, trace, A.tex_mask) : wave_variant_c<false>(small, smem, media, trace, A.tex_mask);
}
static size_t wave_smem_bytes(const RenderArgs &A, const WaveVariant &v, bool smem, bool media, bool trace) {
  const size_t per_slot = (size_t)wave_state_words(media) * 4 + 2 * TPT_WAVE_NQ * 2;
  if (trace && !TPT_PAR) return (size_t)v.slots * per_slot;
  return (smem ? (size_t)A.scene.blob_words * 16 : 0) + (size_t)v.slots * per_slot;
}

cudaError_t TPT_FN(wave_occupancy_)(const RenderArgs &A, bool small, bool smem, bool media, bool trace, int *blocks_per_sm) {
  const WaveVariant v = wave_variant(A, small, smem, media, trace);
  size_t bytes = wave_smem_bytes(A, v, smem, media, trace);
  cudaError_t e = allow_smem(v.fn, bytes);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, v.fn, v.threads, bytes);
}

cudaError_t TPT_FN(launch_wave_)(const RenderArgs &A, bool small, bool smem, bool media, bool trace, int blocks, cudaStream_t st) {
  const WaveVariant v = wave_variant(A, small, smem, media, trace);
  v.fn<<<blocks, v.threads, wave_smem_bytes(A, v, smem, media, trace), st>>>(A);
  return cudaGetLastError();
}

int TPT_FN(wave_threads_)(const RenderArgs &A, bool small, bool smem, bool media, bool trace) {
  return wave_variant(A, small, smem, media, trace).threads;
}

cudaError_t TPT_FN(launch_texture_probe_)(const TextureProbeArgs &A, cudaStream_t st) {
  int blocks = (int)((A.n + 127) / 128);
  if (blocks == 0) return cudaSuccess;
  texture_probe_kernel<TPT_PAR><<<blocks, 128, 0, st>>>(A);
  return cudaGetLastError();
}

} // namespace tptd
