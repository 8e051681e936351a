// kernels.cuh -- generic "per-cell physics -> {shape reduce | per-cell store |
// per-cell time sum}" kernels, templated on a physics functor.
//
// A physics functor `Phys` provides
//   static constexpr bool kVec;   lane layout (see common.cuh: TileGeomT<VEC>)
//   struct Cell;   per-thread constants (geometry of the lane's 4 cells)
//   struct Raw;    the raw field values of one time step (4 cells)
//   static constexpr int kSmemFloats;        CTA-shared lookup tables
//   static constexpr int kBatch, kMinBlocks; time steps per batch / CTAs per SM
//   __device__ void stage(float* smem) const;            (whole CTA, before use)
//   __device__ void init(Cell&, const Geom&, const float* smem) const;
//   __device__ void load(const Cell&, const Geom&, int64_t tb, Raw&) const;
//        tb = byte offset of the time slab (t * S * 4), maintained by the caller
//   __device__ void compute(const Cell&, const Geom&, int t, const Raw&,
//                           float (&v)[4], const float* smem) const;
//        values of out-of-grid lanes are unspecified (finite border copies): they
//        carry zero weight in every slot and are never stored
// `t` is relative to the slab the functor's field pointers address.
//
// Loop structure: a warp owns one 32x4 tile and walks a block of `tb`
// consecutive time steps.  Memory latency is hidden by occupancy (PREFETCH=0:
// few registers, many resident warps) or additionally by a register double
// buffer (PREFETCH=1: the loads of step t+1 are issued before the arithmetic
// of step t).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace atl {

// A warp owns one 32x4 tile and walks `tb` consecutive time steps in batches of
// B.  Per batch: the arithmetic of the B resident steps runs first, then the
// loads of the NEXT batch are issued into the now-dead raw registers, and only
// then the shuffle-reduce + atomics of the current batch execute -- so B steps of
// loads are in flight across the whole reduce phase at no extra register cost.
// Steps are reduced two at a time, slots four at a time, in one transposed
// butterfly (reduce_slots2g; G = 0 selects the older pairwise reduce_slots2).  B is chosen per
// physics so that B x (#fields) 16-byte loads per lane cover the HBM latency
// (PV: 2 x 5, wind: 4 x 2, SpMM: 4 x 1).  MINB = CTAs/SM the register allocator
// must allow.
template <class Phys, int B, int MINB, int G = 0>
__global__ void __launch_bounds__(CTA_THREADS, MINB)
    k_fused_reduce(const Phys phys, const GridDev gd, const PlanDev plan,
                   float* __restrict__ out, int nt, int tb) {
  extern __shared__ float smem[];
  phys.stage(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ai = blockIdx.x * WARPS_PER_CTA + warp;
  if (ai >= plan.n_active) return;
  const int tile = __ldg(plan.active_tiles + ai);
  const auto g = make_geom<Phys::kVec>(tile, lane, gd);
  const int s_beg = __ldg(plan.tile_slot_ptr + tile);
  const int s_end = __ldg(plan.tile_slot_ptr + tile + 1);
  const int t0 = blockIdx.y * tb;
  const int t1 = min(nt, t0 + tb);
  const int nb = plan.n_bus;

  typename Phys::Cell c;
  phys.init(c, g, smem);
  typename Phys::Raw r[B];
  float v[B][4];
  const int64_t S4 = gd.S * 4;
  int64_t sb = (int64_t)t0 * S4;  // byte offset of the next slab to LOAD
  const int nfull = (t1 - t0) / B;
  int t = t0;
  if (nfull > 0) {
#pragma unroll
    for (int j = 0; j < B; ++j) phys.load(c, g, sb + j * S4, r[j]);
    sb += B * S4;
#pragma unroll 1
    for (int k = 0; k < nfull; ++k, t += B) {
#pragma unroll
      for (int j = 0; j < B; ++j) {
        phys.compute(c, g, t + j, r[j], v[j], smem);
        zero_invalid(g, v[j]);
      }
      if (k + 1 < nfull) {
#pragma unroll
        for (int j = 0; j < B; ++j) phys.load(c, g, sb + j * S4, r[j]);
        sb += B * S4;
      }
#pragma unroll
      for (int j = 0; j + 1 < B; j += 2) {
        if (G)  // slots in groups of four (transposed butterfly)
          reduce_slots2g(v[j], v[j + 1], s_beg, s_end, plan, out + (size_t)(t + j) * nb, lane);
        else
          reduce_slots2(v[j], v[j + 1], s_beg, s_end, plan, out + (size_t)(t + j) * nb, lane);
      }
      if (B & 1) reduce_slots(v[B - 1], s_beg, s_end, plan, out + (size_t)(t + B - 1) * nb, lane);
    }
  }
  // tail: fewer than B steps left
#pragma unroll 1
  for (; t < t1; ++t, sb += S4) {
    phys.load(c, g, sb, r[0]);
    phys.compute(c, g, t, r[0], v[0], smem);
    zero_invalid(g, v[0]);
    reduce_slots(v[0], s_beg, s_end, plan, out + (size_t)t * nb, lane);
  }
}

// MODE 0: store per-cell values out[(t - t_begin), y, x]
// MODE 1: accumulate the (NaN-skipping) time sum into out[y, x]
// Rolling software pipeline over B register sets: as soon as step t has been
// evaluated from set j its registers are reloaded with step t + B, so B - 1 steps
// of arithmetic (and the other warps) cover every load.  Load indices are clamped
// to the block's last step instead of predicated (at most B - 1 redundant loads
// and evaluations per time block, their results discarded).
template <class Phys, int MODE, int B, int MINB>
__global__ void __launch_bounds__(CTA_THREADS, MINB)
    k_cells(const Phys phys, const GridDev gd, float* __restrict__ out, int t_begin,
            int t_end, int tb) {
  extern __shared__ float smem[];
  phys.stage(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * WARPS_PER_CTA + warp;
  if (tile >= gd.n_tx * gd.n_ty) return;
  const auto g = make_geom<Phys::kVec>(tile, lane, gd);
  const int t0 = t_begin + blockIdx.y * tb;
  const int t1 = min(t_end, t0 + tb);
  if (t0 >= t1) return;

  typename Phys::Cell c;
  phys.init(c, g, smem);
  typename Phys::Raw r[B];
  float v[4];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t S4 = gd.S * 4;
  const int tl = t1 - 1;
#pragma unroll
  for (int j = 0; j < B; ++j) phys.load(c, g, (int64_t)min(t0 + j, tl) * S4, r[j]);
#pragma unroll 1
  for (int t = t0; t < t1; t += B) {
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const bool live = t + j < t1;
      phys.compute(c, g, min(t + j, tl), r[j], v, smem);
      phys.load(c, g, (int64_t)min(t + j + B, tl) * S4, r[j]);
      if (MODE == 0) {
        if (live) store4(out + (int64_t)(t + j - t_begin) * gd.S_out, gd, g, v);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += (live && v[q] == v[q]) ? v[q] : 0.f;
      }
    }
  }
  if (MODE == 1) atomic_add4(out, gd, g, acc);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Deterministic reduce mode (atl_set_deterministic): bitwise-repeatable results.
bool deterministic();

// Generic CSR gather SpMM: out[t, row] = sum_k val[k] * dense[t, col[k]]
// (aggregate.py:24-32 on an already materialised field).  One warp per
// (row, t); used for plans that do not tile well and as second pass of the
// two-pass fallback.
__global__ void k_csr_spmm(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                           const float* __restrict__ val, const float* __restrict__ dense,
                           int64_t S, float* __restrict__ out, int n_bus, int nt);

// Run-time tuning knobs (environment: ATL_VARIANT, ATL_TB), for experiments.
struct Tuning {
  int variant = 0;
  int tb = 0;
};
const Tuning& tuning();

__global__ void k_gather_slots(const float* __restrict__ partial, const int32_t* __restrict__ row_slot_ptr,
                               const int32_t* __restrict__ row_slots, float* __restrict__ out,
                               int n_bus, int64_t n_slots, int nt);
int launch_gather_slots(const AtlPlan* plan, const float* partial, int64_t nt, float* out,
                        cudaStream_t st);

inline int pick_tb(int n_cta_x, int64_t nt) {
  // time steps per CTA: aim for ~24 CTAs per resident slot (148 SMs x 5 CTAs) so
  // the tail wave is small, but keep the per-CTA prologue (geometry, slot setup)
  // amortised over >= 16 steps.  Measured optimum 32..128 (profiles/r1_tb_sweep.log).
  const int64_t want = 148LL * 5 * 24;
  int64_t tb = (nt * n_cta_x + want - 1) / want;
  if (tb < 16) tb = 16;
  if (tb > 128) tb = 128;
  tb += tb & 1;
  if (tb > nt) tb = nt > 0 ? nt : 1;
  return (int)tb;
}

template <class Phys>
int launch_cells(const Phys& phys, const GridDev& gd, float* out, int64_t t_begin,
                 int64_t t_end, bool timesum, cudaStream_t st) {
  if (t_end <= t_begin) return ATL_OK;
  const int n_tiles = gd.n_tx * gd.n_ty;
  const int gx = (n_tiles + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  // deterministic time sums: one time block, so every cell receives exactly one add
  const int tb = (timesum && deterministic()) ? (int)(t_end - t_begin) : pick_tb(gx, t_end - t_begin);
  const int gy = (int)((t_end - t_begin + tb - 1) / tb);
  dim3 grid(gx, gy);
  const size_t smem = Phys::kSmemFloats * sizeof(float);
  GridDev go = gd;
  go.out_vec = (gd.nx % 4 == 0 && aligned16(out)) ? 1 : 0;
  if (timesum)
    k_cells<Phys, 1, Phys::kBatch, Phys::kMinBlocks>
        <<<grid, CTA_THREADS, smem, st>>>(phys, go, out, (int)t_begin, (int)t_end, tb);
  else
    k_cells<Phys, 0, Phys::kBatch, Phys::kMinBlocks>
        <<<grid, CTA_THREADS, smem, st>>>(phys, go, out, (int)t_begin, (int)t_end, tb);
  ++g_launches;
  ATL_CUDA(cudaGetLastError());
  return ATL_OK;
}

int launch_csr_spmm(const AtlPlan* plan, const float* dense, int64_t nt, float* out,
                    cudaStream_t st);

// Fused path of one slab.  `Phys` must use the plan's lane layout.
template <class Phys>
int launch_fused(const Phys& phys, const AtlPlan* plan, float* out, int64_t nt, cudaStream_t st) {
  const bool det = deterministic() && plan->n_slots > 0;
  float* acc = out;  // where the per-(slot, step) partial sums are accumulated
  PlanDev pd = plan->dev();
  if (det) {
    ATL_CUDA(cudaMallocAsync((void**)&acc, (size_t)nt * plan->n_slots * sizeof(float), st));
    pd = plan->dev_partial();
  }
  ATL_CUDA(cudaMemsetAsync(acc, 0, (size_t)nt * pd.n_bus * sizeof(float), st));
  if (plan->n_active == 0) {
    if (det) ATL_CUDA(cudaFreeAsync(acc, st));
    if (det) ATL_CUDA(cudaMemsetAsync(out, 0, (size_t)nt * plan->n_bus * sizeof(float), st));
    return ATL_OK;
  }
  const int gx = (plan->n_active + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  const int tb = tuning().tb > 0 ? tuning().tb : pick_tb(gx, nt);
  dim3 grid(gx, (unsigned)((nt + tb - 1) / tb));
  const size_t smem = Phys::kSmemFloats * sizeof(float);
  const GridDev gd = plan->grid;
#define ATL_LAUNCH_FUSED(B, MINB, ...) \
  k_fused_reduce<Phys, B, MINB, ##__VA_ARGS__><<<grid, CTA_THREADS, smem, st>>>(phys, gd, pd, acc, (int)nt, tb)
  switch (tuning().variant) {  // A/B experiments (ATL_VARIANT); 0 = the functor's own choice
    case 1: ATL_LAUNCH_FUSED(Phys::kBatch, Phys::kMinBlocks, 0); break;  // pairwise reduce
    case 2: ATL_LAUNCH_FUSED(4, 6, 1); break;                            // deeper batch
    default: ATL_LAUNCH_FUSED(Phys::kBatch, Phys::kMinBlocks, 1); break;
  }
#undef ATL_LAUNCH_FUSED
  ++g_launches;
  ATL_CUDA(cudaGetLastError());
  if (det) {
    int rc = launch_gather_slots(plan, acc, nt, out, st);
    cudaFreeAsync(acc, st);
    return rc;
  }
  return ATL_OK;
}

// Two-pass fallback for matrices that do not tile (e.g. one bus per cell):
// materialise a block of per-cell values, then CSR-gather it.
template <class Phys>
int launch_two_pass(const Phys& phys, const AtlPlan* plan, float* out, int64_t nt,
                    cudaStream_t st) {
  const int64_t S = plan->grid.S_out;  // the scratch cube is unpadded
  int64_t blk = (256LL << 20) / (S * 4);  // <= 256 MiB scratch
  if (blk < 1) blk = 1;
  if (blk > nt) blk = nt;
  float* scratch = nullptr;
  ATL_CUDA(cudaMallocAsync((void**)&scratch, (size_t)blk * S * sizeof(float), st));
  for (int64_t t = 0; t < nt; t += blk) {
    const int64_t n = (nt - t < blk) ? nt - t : blk;
    int rc = launch_cells(phys, plan->grid, scratch, t, t + n, false, st);
    if (rc == ATL_OK) rc = launch_csr_spmm(plan, scratch, n, out + (size_t)t * plan->n_bus, st);
    if (rc != ATL_OK) {
      cudaFreeAsync(scratch, st);
      return rc;
    }
  }
  ATL_CUDA(cudaFreeAsync(scratch, st));
  return ATL_OK;
}

// Dispatch on the lane layout.  `make(vec_tag)` builds the functor for a layout:
// make(std::true_type{}) -> Phys<VEC>, make(std::false_type{}) -> Phys<SCALAR>.
// `ptrs_aligned`: every field pointer is 16-byte aligned (needed for VEC).
template <class Make>
int dispatch_reduce(Make make, const AtlPlan* plan, bool ptrs_aligned, float* out, int64_t nt,
                    cudaStream_t st) {
  if (nt <= 0) return ATL_OK;
  ATL_REQUIRE(nt < (1LL << 31), "slab too long");
  if (plan->fused) {
    if (plan->vec) {
      ATL_REQUIRE(ptrs_aligned,
                  "field pointers must be 16-byte aligned (pitch % 4 == 0 uses 128-bit loads)");
      return launch_fused(make(std::true_type{}), plan, out, nt, st);
    }
    return launch_fused(make(std::false_type{}), plan, out, nt, st);
  }
  if (plan->vec && ptrs_aligned) return launch_two_pass(make(std::true_type{}), plan, out, nt, st);
  return launch_two_pass(make(std::false_type{}), plan, out, nt, st);
}

template <class Make>
int dispatch_cells(Make make, const GridDev& gd, bool ptrs_aligned, float* out, int64_t nt,
                   bool timesum, cudaStream_t st) {
  if (gd.pitch % 4 == 0 && ptrs_aligned)
    return launch_cells(make(std::true_type{}), gd, out, 0, nt, timesum, st);
  return launch_cells(make(std::false_type{}), gd, out, 0, nt, timesum, st);
}

}  // namespace atl
