// wind.cu -- hub-height extrapolation + power-curve interpolation, fused with
// the shape reduce.  Follows wind.py:75-112 (extrapolate_wind_speed) and
// convert.py:648-649 (np.interp(v_hub, V, POW / P)).
//
// Algorithmic traffic: 8 B per cell-timestep (wnd + roughness | shear
// exponent), 4 B in the wnd{hub}m fast lane (wind.py:75-78).
#include <cmath>
#include <vector>

#include "kernels.cuh"

namespace atl {

// Power curve in shared memory: 4 arrays of NK (power of two >= n_knots):
//   xcmp  knot abscissae rounded UP to float  -> `x >= xcmp[j]` is exactly the
//         float64 comparison `x >= V[j]` of numpy's binary search for float x
//   xval  knot abscissae rounded to nearest   (for x - V[j])
//   f     POW[j] / P
//   slope (f[j+1]-f[j]) / (V[j+1]-V[j])  (0 on zero-width segments)
// padded with +inf (NK > n_knots) so a branch-free binary search counts the
// knots <= x.
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

  struct Cell {};
  struct Raw {
    float w[4], a[4];
  };
  static constexpr int kSmemFloats = 4 * 256;

  __device__ void stage(float* smem) const {
    for (int i = threadIdx.x; i < 4 * NK; i += blockDim.x) smem[i] = curve[i];
    __syncthreads();
  }
  __device__ void init(Cell&, const Geom&, const float*) const {}
  __device__ void load(const Cell&, const Geom& g, int t, Raw& r) const {
    load4(wnd, S, g, t, r.w);
    if (method != ATL_WIND_NONE) load4(aux, S, g, t, r.a);
  }
  __device__ __forceinline__ float interp(float x, const float* sm) const {
    const float* xcmp = sm;
    const float* xval = sm + NK;
    const float* f = sm + 2 * NK;
    const float* slope = sm + 3 * NK;
    int cnt = 0;  // number of knots <= x
    for (int step = NK >> 1; step >= 1; step >>= 1)
      if (xcmp[cnt + step - 1] <= x) cnt += step;
    // np.interp: x < V[0] -> f[0]; x >= V[n-1] -> f[n-1]; else linear on [j, j+1)
    const int j = max(cnt - 1, 0);
    float r = fmaf(slope[j], x - xval[j], f[j]);
    r = (cnt == 0) ? f[0] : r;
    r = (cnt >= n_knots) ? f[n_knots - 1] : r;
    return (x != x) ? x : r;
  }
  __device__ void compute(const Cell&, const Geom& g, int, const Raw& r, float (&v)[4],
                          const float* sm) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x = r.w[i];
      if (method == ATL_WIND_LOG) {
        // v * ln(to/z0) / ln(from/z0) = v * (lg2 to - lg2 z0) / (lg2 from - lg2 z0)
        const float L = __log2f(r.a[i]);
        x = x * __fdividef(lg2_to - L, lg2_from - L);
      } else if (method == ATL_WIND_POWER) {
        x = x * exp2f(r.a[i] * lg2_ratio);  // v * (to/from)^alpha
      }
      const float p = interp(x, sm);
      v[i] = ((g.valid >> i) & 1u) ? p : 0.f;
    }
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
  std::vector<float> curve((size_t)4 * NK);
  const int n = cfg->n_knots;
  for (int j = 0; j < NK; ++j) {
    if (j < n) {
      const double x = cfg->V[j];
      float xc = (float)x;
      if ((double)xc < x) xc = nextafterf(xc, INFINITY);  // round up
      curve[j] = xc;
      curve[NK + j] = (float)x;
      curve[2 * NK + j] = (float)cfg->POW_norm[j];
      double sl = 0.0;
      if (j + 1 < n && cfg->V[j + 1] > cfg->V[j])
        sl = (cfg->POW_norm[j + 1] - cfg->POW_norm[j]) / (cfg->V[j + 1] - cfg->V[j]);
      curve[3 * NK + j] = (float)sl;
    } else {
      curve[j] = INFINITY;
      curve[NK + j] = 0.f;
      curve[2 * NK + j] = 0.f;
      curve[3 * NK + j] = 0.f;
    }
  }
  AtlWindOp* op = new AtlWindOp();
  op->device = device;
  op->grid = make_grid(cfg->ny, cfg->nx);
  op->method = cfg->method;
  op->n_knots = n;
  op->NK = NK;
  op->lg2_to = op->lg2_from = op->lg2_ratio = 0.f;
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
  ATL_REQUIRE(plan->grid.nx == op->grid.nx && plan->grid.ny == op->grid.ny,
              "plan / operator grid mismatch");
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
