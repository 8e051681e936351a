// common.cuh -- shared device/host helpers for libatlite_b200 (sm_100a)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/atlite_b200.h"

namespace atl {

// ---------------------------------------------------------------- errors
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
extern int64_t g_launches;

#define ATL_CUDA(call)                                  \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return atl::cuda_fail(_e, #call); \
  } while (0)

#define ATL_REQUIRE(cond, msg)        \
  do {                                \
    if (!(cond)) {                    \
      atl::set_error(msg);            \
      return ATL_ERR_INVALID;         \
    }                                 \
  } while (0)

// ---------------------------------------------------------------- tiling
// The grid (ny, nx) is covered by warp tiles of 32 (x) x 4 (y) cells.  Lane l of
// the owning warp handles column x = 32*tx + l and the four rows 4*ty .. 4*ty+3,
// so every field load of a warp is one fully coalesced 128-byte row segment and
// needs no alignment of nx (works for any cutout width).  A CTA is 4 warps = 4
// consecutive tiles (adjacent in x: 512 contiguous bytes per row and field).
constexpr int TILE_X = 32;
constexpr int TILE_Y = 4;
constexpr int TILE_CELLS = TILE_X * TILE_Y;
constexpr int WARPS_PER_CTA = 4;
constexpr int CTA_THREADS = 32 * WARPS_PER_CTA;

struct GridDev {
  int nx, ny;
  int n_tx, n_ty;  // tiles per row / column
  int64_t S;       // ny * nx
};

inline GridDev make_grid(int ny, int nx) {
  GridDev g;
  g.nx = nx;
  g.ny = ny;
  g.n_tx = (nx + TILE_X - 1) / TILE_X;
  g.n_ty = (ny + TILE_Y - 1) / TILE_Y;
  g.S = (int64_t)ny * nx;
  return g;
}

// Device view of an aggregation plan (see plan.cu).
struct PlanDev {
  const int32_t* tile_slot_ptr;  // [n_tiles + 1]
  const int32_t* slot_row;       // [n_slots] bus index of each (tile, bus) slot
  const float4* slot_w4;         // [n_slots * 32] weights: lane l -> rows 0..3 of column l
  const int32_t* active_tiles;   // [n_active]
  int32_t n_active;
  int32_t n_bus;
};

struct TileGeom {
  int x;           // column handled by this lane
  int y0;          // first row of the tile
  int base;        // y0 * nx + x  (flat cell offset of row 0)
  unsigned valid;  // bit r set: cell (y0 + r, x) lies inside the grid
};

#ifdef __CUDACC__
__device__ __forceinline__ TileGeom make_geom(int tile, int lane, const GridDev& gd) {
  TileGeom g;
  const int tx = tile % gd.n_tx, ty = tile / gd.n_tx;
  g.x = tx * TILE_X + lane;
  g.y0 = ty * TILE_Y;
  g.base = g.y0 * gd.nx + g.x;
  unsigned v = 0;
  if (g.x < gd.nx) {
#pragma unroll
    for (int r = 0; r < TILE_Y; ++r)
      if (g.y0 + r < gd.ny) v |= 1u << r;
  }
  g.valid = v;
  return g;
}

__device__ __forceinline__ float warp_sum(float p) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  return p;
}

// Streaming (read-once) global load: bypass L1 allocation so the small, hot
// weight / table lines stay resident.
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// Load the 4 rows of one field for this lane at time step t (slab-relative).
__device__ __forceinline__ void load4(const float* __restrict__ f, int64_t S, int nx,
                                      const TileGeom& g, int t, float (&o)[4]) {
  const float* p = f + (int64_t)t * S + g.base;
#pragma unroll
  for (int r = 0; r < TILE_Y; ++r) o[r] = (g.valid >> r) & 1u ? ld_stream(p + r * nx) : 0.f;
}
__device__ __forceinline__ void load4(const double* __restrict__ f, int64_t S, int nx,
                                      const TileGeom& g, int t, float (&o)[4]) {
  const double* p = f + (int64_t)t * S + g.base;
#pragma unroll
  for (int r = 0; r < TILE_Y; ++r)
    o[r] = (g.valid >> r) & 1u ? (float)ld_stream(p + r * nx) : 0.f;
}

// Reduce the 4 per-cell values of every lane into the tile's (tile, bus) slots
// and add the warp totals to out_row[bus].  Weight lines are L1-resident after
// the first time step.  A NaN/Inf anywhere in the warp takes the exact path in
// which zero weights do not touch the value (a sparse matrix never multiplies
// entries it does not store: aggregate.py:25 / scipy CSR product).
__device__ __forceinline__ void reduce_slots(const float (&v)[4], int s_beg, int s_end,
                                             const PlanDev& plan, float* __restrict__ out_row,
                                             int lane) {
  const float chk = (v[0] + v[1]) + (v[2] + v[3]);
  const bool bad = !(fabsf(chk) <= 3.0e38f);
  if (!__any_sync(0xffffffffu, bad)) {
    for (int s = s_beg; s < s_end; ++s) {
      const float4 w = __ldg(plan.slot_w4 + (size_t)s * 32 + lane);
      float p = fmaf(w.x, v[0], fmaf(w.y, v[1], fmaf(w.z, v[2], w.w * v[3])));
      p = warp_sum(p);
      if (lane == 0) atomicAdd(out_row + __ldg(plan.slot_row + s), p);
    }
  } else {
    for (int s = s_beg; s < s_end; ++s) {
      const float4 w = __ldg(plan.slot_w4 + (size_t)s * 32 + lane);
      float p = 0.f;
      if (w.x != 0.f) p += w.x * v[0];
      if (w.y != 0.f) p += w.y * v[1];
      if (w.z != 0.f) p += w.z * v[2];
      if (w.w != 0.f) p += w.w * v[3];
      p = warp_sum(p);
      if (lane == 0) atomicAdd(out_row + __ldg(plan.slot_row + s), p);
    }
  }
}
#endif  // __CUDACC__

}  // namespace atl

// Opaque handle definitions shared between translation units.
struct AtlPlan {
  int device;
  atl::GridDev grid;
  int32_t n_bus;
  int64_t nnz;
  int32_t n_tiles, n_active;
  int64_t n_slots;
  bool fused;
  // device arrays
  int32_t* d_tile_slot_ptr = nullptr;
  int32_t* d_slot_row = nullptr;
  float4* d_slot_w4 = nullptr;
  int32_t* d_active = nullptr;
  // CSR copy for the two-pass fallback / generic SpMM
  int64_t* d_indptr = nullptr;
  int32_t* d_indices = nullptr;
  float* d_vals = nullptr;
  atl::PlanDev dev() const {
    atl::PlanDev p;
    p.tile_slot_ptr = d_tile_slot_ptr;
    p.slot_row = d_slot_row;
    p.slot_w4 = d_slot_w4;
    p.active_tiles = d_active;
    p.n_active = n_active;
    p.n_bus = n_bus;
    return p;
  }
};
