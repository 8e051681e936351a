// wind.cu -- hub-height extrapolation + power-curve interpolation, fused with
// the shape reduce.  Follows wind.py:75-112 (extrapolate_wind_speed) and
// convert.py:648-649 (np.interp(v_hub, V, POW / P)).
//
// Algorithmic traffic: 8 B per cell-timestep (wnd + roughness | shear
// exponent), 4 B in the wnd{hub}m fast lane (wind.py:75-78).
#include <cmath>
#include <cstring>
#include <vector>

#include "kernels.cuh"

namespace atl {

// Power curve in shared memory (NK = power of two > n_knots):
//   xcmp[NK]   knot abscissae rounded UP to float -> `xcmp[j] <= x` is exactly the
//              float64 comparison `V[j] <= x` of numpy's binary search for a float
//              x; padded with +inf so a branch-free binary search counts the
//              knots <= x
//   seg[NK+1]  float4 {x0, f0, slope, -} indexed by that COUNT c: c = 0 -> left
//              clamp f[0]; 1 <= c < n -> segment [V[c-1], V[c]) with x0 = V[c-1]
//              (nearest float), slope = 0 on zero-width (duplicate-knot)
//              segments; c >= n -> right clamp f[n-1]      (np.interp semantics)
template <bool VEC>
struct WindPhys {
  static constexpr bool kVec = VEC;
  using Geom = TileGeomT<VEC>;
  const float* wnd;
  const float* aux;
  const float* curve;  // device, 4 * NK floats
  int64_t S;
  int method;
  int n_knots, NK;
  float lg2_to, lg2_from, lg2_ratio;
  float x_lo, x_hi;
  int use_lut, n_stage;  // LUT mode: the bucket table follows xcmp | seg in `curve`
  float inv_w;

  struct Cell {};
  struct Raw {
    float w[4], a[4];
  };
  static constexpr int kSmemFloats = 256 + 4 * 257 + 2 * 1025 + 2;  // xcmp | seg | LUT (<= 1025 x 8 B)
  static constexpr int kBatch = 2, kMinBlocks = 6;

  __device__ void stage(float* smem) const {
    for (int i = threadIdx.x; i < n_stage; i += blockDim.x) smem[i] = curve[i];
    __syncthreads();
  }
  __device__ void init(Cell&, const Geom&, const float*) const {}
  __device__ void load(const Cell&, const Geom& g, int64_t tb, Raw& r) const {
    load4(wnd, tb, g, r.w);
    if (method != ATL_WIND_NONE) load4(aux, tb, g, r.a);
  }
  // np.interp for the lane's 4 values at once: find cnt = #knots <= x, then
  // evaluate segment seg[cnt].  `xcmp[j] <= x` with xcmp rounded UP to float is
  // exactly NumPy's float64 comparison, so the same segment is selected --
  // including duplicate knots (cut-out) and the clamped ends.
  //  LUT mode (the normal case): a uniform grid on [V0, Vn-1], shifted by half a
  //  bucket, in which every bucket holds at most ONE distinct knot value, none
  //  within 1e-3 of a bucket edge (host-checked; else the fallback).  A bucket
  //  stores that knot (rounded up) and the segment offsets for x below / from it
  //  on: one 8-byte load, one compare, one 16-byte load.
  //  Fallback: branch-free binary search, the four searches in lock step.
  __device__ __forceinline__ void interp4(const float (&x)[4], float (&r)[4], const float* sm) const {
    const float* xcmp = sm;
    const float4* seg = reinterpret_cast<const float4*>(sm + NK);
    if (use_lut) {
      const float2* lut = reinterpret_cast<const float2*>(sm + NK + 4 * (NK + 1));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xc = fminf(fmaxf(x[i], x_lo), x_hi);
        const int b = __float2int_rd(fmaf(xc - x_lo, inv_w, 0.5f));
        const float2 e = lut[b];  // {knot inside the bucket (+inf if none), packed seg offsets}
        const unsigned pk = __float_as_uint(e.y);
        const unsigned off = (x[i] >= e.x) ? (pk >> 16) : (pk & 0xffffu);  // bytes into seg[]
        const float4 s = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(seg) + off);
        const float y = fmaf(s.z, xc - s.x, s.y);
        r[i] = (x[i] != x[i]) ? x[i] : y;
      }
      return;
    }
    int cnt[4] = {0, 0, 0, 0};  // number of knots <= x  (NaN compares false -> 0)
#pragma unroll 1
    for (int step = NK >> 1; step >= 1; step >>= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (xcmp[cnt[i] + step - 1] <= x[i]) cnt[i] += step;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 s = seg[min(cnt[i], n_knots)];
      // clamp keeps inf * 0 out of the clamped ends; NaN is restored below
      const float xc = fminf(fmaxf(x[i], x_lo), x_hi);
      const float y = fmaf(s.z, xc - s.x, s.y);
      r[i] = (x[i] != x[i]) ? x[i] : y;
    }
  }
  __device__ void compute(const Cell&, const Geom&, int, const Raw& r, float (&v)[4],
                          const float* sm) const {
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[i] = r.w[i];
      if (method == ATL_WIND_LOG) {
        // v * ln(to/z0) / ln(from/z0) = v * (lg2 to - lg2 z0) / (lg2 from - lg2 z0)
        const float L = __log2f(r.a[i]);
        x[i] = x[i] * __fdividef(lg2_to - L, lg2_from - L);
      } else if (method == ATL_WIND_POWER) {
        x[i] = x[i] * exp2f(r.a[i] * lg2_ratio);  // v * (to/from)^alpha
      }
    }
    interp4(x, v, sm);
  }
};

}  // namespace atl

using namespace atl;

struct AtlWindOp {
  int device;
  GridDev grid;
  int method;
  int n_knots, NK;
  float lg2_to, lg2_from, lg2_ratio;
  float x_lo, x_hi;
  int use_lut, n_stage;
  float inv_w;
  float* d_curve = nullptr;
};

template <bool VEC>
static WindPhys<VEC> make_phys(const AtlWindOp* op, const AtlWindFields* f) {
  WindPhys<VEC> p;
  p.wnd = f->wnd;
  p.aux = f->aux;
  p.curve = op->d_curve;
  p.S = op->grid.S;
  p.method = op->method;
  p.n_knots = op->n_knots;
  p.NK = op->NK;
  p.lg2_to = op->lg2_to;
  p.lg2_from = op->lg2_from;
  p.lg2_ratio = op->lg2_ratio;
  p.x_lo = op->x_lo;
  p.x_hi = op->x_hi;
  p.use_lut = op->use_lut;
  p.n_stage = op->n_stage;
  p.inv_w = op->inv_w;
  return p;
}

static int check_fields(const AtlWindOp* op, const AtlWindFields* f) {
  ATL_REQUIRE(op && f && f->wnd, "NULL argument");
  ATL_REQUIRE(op->method == ATL_WIND_NONE || f->aux,
              "roughness / wnd_shear_exp field missing for the chosen method");
  return ATL_OK;
}

extern "C" {

int atl_wind_create(int device, const AtlWindConfig* cfg, AtlWindOp** op_out) {
  ATL_REQUIRE(cfg && op_out, "NULL argument");
  *op_out = nullptr;
  ATL_REQUIRE(cfg->ny > 0 && cfg->nx > 0, "bad grid");
  ATL_REQUIRE(cfg->n_knots >= 1 && cfg->n_knots <= 255, "n_knots must be in [1, 255]");
  ATL_REQUIRE(cfg->V && cfg->POW_norm, "power curve missing");
  ATL_REQUIRE(cfg->method >= ATL_WIND_NONE && cfg->method <= ATL_WIND_POWER, "bad method");
  for (int i = 1; i < cfg->n_knots; ++i)
    ATL_REQUIRE(cfg->V[i] >= cfg->V[i - 1], "wind speed knots must be non-decreasing");
  if (cfg->method != ATL_WIND_NONE)
    ATL_REQUIRE(cfg->from_height > 0 && cfg->to_height > 0, "heights must be positive");

  int NK = 2;  // strictly more slots than knots: the search counts up to NK-1
  while (NK <= cfg->n_knots) NK <<= 1;
  // xcmp[NK] followed by seg[NK + 1] (float4), see WindPhys; NK * 4 bytes keeps
  // the float4 part 16-byte aligned (NK >= 4)
  if (NK < 4) NK = 4;
  std::vector<float> curve((size_t)NK + 4 * ((size_t)NK + 1), 0.f);
  const int n = cfg->n_knots;
  for (int j = 0; j < NK; ++j) {
    if (j < n) {
      const double x = cfg->V[j];
      float xc = (float)x;
      if ((double)xc < x) xc = nextafterf(xc, INFINITY);  // round up
      curve[j] = xc;
    } else {
      curve[j] = INFINITY;
    }
  }
  float* seg = curve.data() + NK;
  for (int c = 0; c <= NK; ++c) {
    float x0 = 0.f, f0 = 0.f, sl = 0.f;
    if (c == 0) {
      x0 = (float)cfg->V[0];
      f0 = (float)cfg->POW_norm[0];
    } else if (c >= n) {
      x0 = (float)cfg->V[n - 1];
      f0 = (float)cfg->POW_norm[n - 1];
    } else {
      const int j = c - 1;
      x0 = (float)cfg->V[j];
      f0 = (float)cfg->POW_norm[j];
      if (cfg->V[j + 1] > cfg->V[j])
        sl = (float)((cfg->POW_norm[j + 1] - cfg->POW_norm[j]) / (cfg->V[j + 1] - cfg->V[j]));
    }
    seg[4 * c + 0] = x0;
    seg[4 * c + 1] = f0;
    seg[4 * c + 2] = sl;
    seg[4 * c + 3] = 0.f;
  }
  // ---- uniform-bucket LUT (see interp4)
  int use_lut = 0;
  float inv_w = 0.f;
  {
    const double lo = cfg->V[0], hi = cfg->V[n - 1];
    for (int NB = 32; NB <= 1024 && !use_lut && hi > lo; NB *= 2) {
      const double wdt = (hi - lo) / NB;
      std::vector<int> bucket_of(n);
      std::vector<double> knot_in((size_t)NB + 1, std::nan(""));
      bool ok = true;
      for (int j = 0; j < n && ok; ++j) {
        const double pos = (cfg->V[j] - lo) / wdt + 0.5;
        const int b = (int)std::floor(pos);
        const double frac = pos - b;
        if (b < 0 || b > NB || frac < 1e-3 || frac > 1.0 - 1e-3) ok = false;
        else if (!std::isnan(knot_in[b]) && knot_in[b] != cfg->V[j]) ok = false;  // 2 distinct knots
        else {
          knot_in[b] = cfg->V[j];
          bucket_of[j] = b;
        }
      }
      if (!ok) continue;
      const size_t base = curve.size();  // NK + 4 (NK + 1): even, so the float2 table is 8-byte aligned
      curve.resize(base + 2 * ((size_t)NB + 1), 0.f);
      int c_lo = 0;  // knots in buckets < b
      for (int b = 0; b <= NB; ++b) {
        int c_hi = c_lo;
        while (c_hi < n && bucket_of[c_hi] == b) ++c_hi;
        float thr = INFINITY;
        if (!std::isnan(knot_in[b])) {
          thr = (float)knot_in[b];
          if ((double)thr < knot_in[b]) thr = nextafterf(thr, INFINITY);  // round up
        }
        const uint32_t pk = ((uint32_t)(c_hi * 16) << 16) | (uint32_t)(c_lo * 16);
        float pkf;
        std::memcpy(&pkf, &pk, 4);
        curve[base + 2 * (size_t)b] = thr;
        curve[base + 2 * (size_t)b + 1] = pkf;
        c_lo = c_hi;
      }
      use_lut = 1;
      inv_w = (float)(1.0 / wdt);
    }
  }

  AtlWindOp* op = new AtlWindOp();
  op->device = device;
  ATL_REQUIRE(cfg->pitch == 0 || cfg->pitch >= cfg->nx, "pitch must be >= nx");
  op->grid = make_grid(cfg->ny, cfg->nx, cfg->pitch);
  op->method = cfg->method;
  op->n_knots = n;
  op->NK = NK;
  op->lg2_to = op->lg2_from = op->lg2_ratio = 0.f;
  op->x_lo = (float)cfg->V[0];
  op->x_hi = (float)cfg->V[n - 1];
  op->use_lut = use_lut;
  op->n_stage = (int)curve.size();
  op->inv_w = inv_w;
  if (cfg->method != ATL_WIND_NONE) {
    op->lg2_to = (float)std::log2(cfg->to_height);
    op->lg2_from = (float)std::log2(cfg->from_height);
    op->lg2_ratio = (float)std::log2(cfg->to_height / cfg->from_height);
  }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&op->d_curve, curve.size() * 4);
  if (e == cudaSuccess)
    e = cudaMemcpy(op->d_curve, curve.data(), curve.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    atl_wind_destroy(op);
    return cuda_fail(e, "atl_wind_create");
  }
  *op_out = op;
  return ATL_OK;
}

void atl_wind_destroy(AtlWindOp* op) {
  if (!op) return;
  cudaSetDevice(op->device);
  cudaFree(op->d_curve);
  delete op;
}

int atl_wind_op_info(const AtlWindOp* op, int32_t* device, int32_t* ny, int32_t* nx) {
  ATL_REQUIRE(op, "NULL argument");
  if (device) *device = op->device;
  if (ny) *ny = op->grid.ny;
  if (nx) *nx = op->grid.nx;
  return ATL_OK;
}

int atl_wind_reduce(const AtlWindOp* op, const AtlPlan* plan, const AtlWindFields* f,
                    int64_t nt, float* out_dev, void* stream) {
  int rc = check_fields(op, f);
  if (rc) return rc;
  ATL_REQUIRE(plan && out_dev, "NULL argument");
  ATL_REQUIRE(plan->grid.nx == op->grid.nx && plan->grid.ny == op->grid.ny &&
                  plan->grid.pitch == op->grid.pitch,
              "plan / operator grid (or pitch) mismatch");
  ATL_CUDA(cudaSetDevice(op->device));
  auto make = [&](auto vec) { return make_phys<decltype(vec)::value>(op, f); };
  return dispatch_reduce(make, plan, aligned16(f->wnd) && aligned16(f->aux), out_dev, nt,
                         (cudaStream_t)stream);
}

int atl_wind_cells(const AtlWindOp* op, const AtlWindFields* f, int64_t nt, float* out_dev,
                   void* stream) {
  int rc = check_fields(op, f);
  if (rc) return rc;
  ATL_REQUIRE(out_dev, "NULL argument");
  ATL_CUDA(cudaSetDevice(op->device));
  auto make = [&](auto vec) { return make_phys<decltype(vec)::value>(op, f); };
  return dispatch_cells(make, op->grid, aligned16(f->wnd) && aligned16(f->aux), out_dev, nt, false,
                        (cudaStream_t)stream);
}

int atl_wind_timesum(const AtlWindOp* op, const AtlWindFields* f, int64_t nt, float* out_dev,
                     void* stream) {
  int rc = check_fields(op, f);
  if (rc) return rc;
  ATL_REQUIRE(out_dev, "NULL argument");
  ATL_CUDA(cudaSetDevice(op->device));
  auto make = [&](auto vec) { return make_phys<decltype(vec)::value>(op, f); };
  return dispatch_cells(make, op->grid, aligned16(f->wnd) && aligned16(f->aux), out_dev, nt, true,
                        (cudaStream_t)stream);
}

}  // extern "C"
