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
template <bool PAR, bool SMEM>
__global__ void __launch_bounds__(128) intersect_kernel(const __grid_constant__ IntersectArgs A) {
  extern __shared__ float4 sblob[];
  SceneView S;
  S.blob = stage_scene<SMEM>(A.scene, sblob);
  S.L = &A.scene;
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
  if (closest_hit<PAR>(S, r, A.tmin, A.tmax, t, prim)) {
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
template <bool PAR, bool SMEM>
__global__ void __launch_bounds__(TPT_MEGA_THREADS) render_mega_kernel(const __grid_constant__ RenderArgs A) {
  extern __shared__ float4 sblob[];
  SceneView S;
  S.blob = stage_scene<SMEM>(A.scene, sblob);
  S.L = &A.scene;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;

  bool have_bin = false, exhausted = false, active = false;
  int px = 0, py = 0, k = 0, k_end = 0;
  unsigned acc_index = 0;
  V3 acc = mk(0, 0, 0);
  PathState ps;
  Rng rng;
  unsigned long long n_rays = 0, n_nan = 0, n_paths = 0;
  const unsigned bins_per_tile = (unsigned)(TPT_TILE * TPT_TILE) * (unsigned)A.n_ranges;

  for (;;) {
    bool need = !active && !exhausted && (!have_bin || k >= k_end);
    if (need && have_bin) { // bin finished: one store per (pixel, range)
      float *o = A.acc + (size_t)acc_index * 3;
      o[0] = acc.x;
      o[1] = acc.y;
      o[2] = acc.z;
      have_bin = false;
    }
    unsigned m = __ballot_sync(FULL, need);
    if (m) {
      unsigned long long base = 0;
      int leader = __ffs(m) - 1;
      if ((int)lane == leader) base = atomicAdd(A.counters + 0, (unsigned long long)__popc(m));
      base = __shfl_sync(FULL, base, leader);
      if (need) {
        unsigned long long b = base + __popc(m & ((1u << lane) - 1u));
        if (b >= A.n_bins) {
          exhausted = true;
        } else {
          unsigned bb = (unsigned)b;
          unsigned tile_local = bb / bins_per_tile;
          unsigned rem = bb - tile_local * bins_per_tile;
          unsigned range = rem / (unsigned)(TPT_TILE * TPT_TILE);
          unsigned pit = rem - range * (unsigned)(TPT_TILE * TPT_TILE);
          unsigned tile = (unsigned)A.part_index + tile_local * (unsigned)A.part_count;
          unsigned ty = tile / (unsigned)A.tiles_x, tx = tile - ty * (unsigned)A.tiles_x;
          px = (int)(tx * TPT_TILE + (pit & (TPT_TILE - 1)));
          py = (int)(ty * TPT_TILE + (pit / TPT_TILE));
          if (px < A.nx && py < A.ny) {
            have_bin = true;
            k = A.range_bounds[range];
            k_end = A.range_bounds[range + 1];
            acc = mk(0, 0, 0);
            acc_index = range * (unsigned)(A.nx * A.ny) + (unsigned)(py * A.nx + px);
          }
        }
      }
    }
    if (!active && have_bin && k < k_end) { // next sample of this lane's pixel
      rng.begin(A.seed_lo, A.seed_hi, (uint32_t)(py * A.nx + px), (uint32_t)k);
      ps.ray = camera_sample<PAR>(A.cam, px, py, A.nx, A.ny, rng);
      ps.T = mk(1.f, 1.f, 1.f);
      ps.depth = 0;
      active = true;
      n_paths++;
    }
    if (__all_sync(FULL, !active && exhausted)) break;
    if (active) {
      V3 rad;
      n_rays++;
      if (!bounce<PAR>(S, ps, rng, A.max_depth, A.t_min, rad)) {
        // col += de_nan(tmp): main.cpp:126, headers/utils.h:100-109
        bool nan_any = isnan(rad.x) || isnan(rad.y) || isnan(rad.z) || isnan(ps.T.x) ||
                       isnan(ps.T.y) || isnan(ps.T.z);
        if (nan_any) n_nan++;
        acc.x += isnan(rad.x) ? 0.f : rad.x;
        acc.y += isnan(rad.y) ? 0.f : rad.y;
        acc.z += isnan(rad.z) ? 0.f : rad.z;
        active = false;
        k++;
      }
    }
  }
  // statistics: one atomic per warp and counter
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_rays += __shfl_xor_sync(FULL, n_rays, o);
    n_nan += __shfl_xor_sync(FULL, n_nan, o);
    n_paths += __shfl_xor_sync(FULL, n_paths, o);
  }
  if (lane == 0) {
    atomicAdd(A.counters + 1, n_rays);
    atomicAdd(A.counters + 2, n_nan);
    atomicAdd(A.counters + 3, n_paths);
  }
}

// ------------------------------------------------------------------------------------------
// Known-answer probes
// ------------------------------------------------------------------------------------------
template <bool PAR> __global__ void texture_probe_kernel(const __grid_constant__ TextureProbeArgs A) {
  SceneView S;
  S.blob = A.scene.blob_global;
  S.L = &A.scene;
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
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t TPT_FN(launch_intersect_)(const IntersectArgs &A, bool smem, cudaStream_t st) {
  size_t bytes = smem ? (size_t)A.scene.blob_words * 16 : 0;
  int blocks = (int)((A.n + 127) / 128);
  if (blocks == 0) return cudaSuccess;
  cudaError_t e;
  if (smem) {
    if ((e = allow_smem(intersect_kernel<TPT_PAR, true>, bytes)) != cudaSuccess) return e;
    intersect_kernel<TPT_PAR, true><<<blocks, 128, bytes, st>>>(A);
  } else {
    intersect_kernel<TPT_PAR, false><<<blocks, 128, 0, st>>>(A);
  }
  return cudaGetLastError();
}

cudaError_t TPT_FN(mega_occupancy_)(bool smem, size_t smem_bytes, int *blocks_per_sm) {
  cudaError_t e;
  if (smem) {
    if ((e = allow_smem(render_mega_kernel<TPT_PAR, true>, smem_bytes)) != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, render_mega_kernel<TPT_PAR, true>,
                                                         TPT_MEGA_THREADS, smem_bytes);
  }
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, render_mega_kernel<TPT_PAR, false>,
                                                       TPT_MEGA_THREADS, 0);
}

cudaError_t TPT_FN(launch_mega_)(const RenderArgs &A, bool smem, int blocks, cudaStream_t st) {
  size_t bytes = smem ? (size_t)A.scene.blob_words * 16 : 0;
  if (smem)
    render_mega_kernel<TPT_PAR, true><<<blocks, TPT_MEGA_THREADS, bytes, st>>>(A);
  else
    render_mega_kernel<TPT_PAR, false><<<blocks, TPT_MEGA_THREADS, 0, st>>>(A);
  return cudaGetLastError();
}

cudaError_t TPT_FN(launch_texture_probe_)(const TextureProbeArgs &A, cudaStream_t st) {
  int blocks = (int)((A.n + 127) / 128);
  if (blocks == 0) return cudaSuccess;
  texture_probe_kernel<TPT_PAR><<<blocks, 128, 0, st>>>(A);
  return cudaGetLastError();
}

} // namespace tptd
