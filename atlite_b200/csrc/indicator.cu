// indicator.cu -- shapes -> (n_shapes, ny*nx) overlap matrix on a regular grid.
//
// Replaces the step in front of the hot path, Cutout.indicatormatrix ->
// gis.py:104-145 compute_indicatormatrix: I[i, j] = area(shape_i ∩ cell_j) /
// area(cell_j), where the cells are the boxes of Cutout.grid (cutout.py:355-376:
// centre (x, y), half-width (dx/2, dy/2)).  The reference loops over shapely
// intersections in Python; on a regular grid the same areas follow exactly from
// the polygon edges alone.
//
// Method (signed-area accumulation, the exact-coverage technique of vector
// rasterisers): in grid units (cells = unit squares) and for one row band
// j <= v < j+1, the part of cell i covered by a polygon is
//     cov(j, i) = - sum over edges  dir * ∫ clamp(i + 1 - u_e(v), 0, 1) dv
// (dir = +1 for an edge running towards +v; counter-clockwise rings).  For one
// edge piece the integral is 0 left of it, its full height h right of it, and a
// closed-form quadratic inside the cells it crosses, so an edge only touches the
// cells it passes through if the DIFFERENCES D(i) = F(i) - F(i-1) are
// accumulated and a prefix sum along x follows.
//
//   k_edges   one thread per edge: walks the rows / cells it crosses, adds the
//             differences into the shape's bounding-box buffer.  Sums are kept
//             as 2^-52 fixed point in 64-bit integer atomics: associative, so the
//             result (and with it the sparsity pattern) is identical run to run.
//   k_scan    one warp per buffer row: prefix sum -> coverage in [0, 1], counts
//             the entries above the keep threshold.
//   k_emit    one warp per buffer row: ordered compaction into CSR
//             (column = iy * nx + ix, the order of cutout.grid).
//
// Holes carry the opposite sign; ring orientation is normalised on the host.
// Not bandwidth-critical (O(edges + bounding-box cells)); float64 arithmetic.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace atl {

constexpr double kFix = 4503599627370496.0;  // 2^52: one ulp of a coverage fraction; sums wrap
                                              // harmlessly (mod 2^64) as long as the final coverage is < 2048
constexpr double kKeep = 1e-10;           // coverage fractions at or below this are dropped

struct EdgeDev {
  double u0, v0, u1, v1;  // grid units relative to the cutout's lower-left cell corner
  int32_t shape;          // index into the batch's box arrays
  int32_t sign;           // +1 / -1: ring weight x orientation normalisation
};

struct BoxDev {
  int32_t i0, j0, w, h;  // bounding box in cells (clipped to the grid)
  int64_t off;           // offset of the box buffer (w*h entries)
  int64_t row0;          // index of its first row in the batch's row list
};

// Mean of c(t) = clamp(1 - t, 0, 1) over [t0, t1] (t = u - i, the cell-local
// coordinate).  Built from the lengths of the pieces t <= 0 (c = 1) and
// 0 < t < 1 (c linear) rather than from an antiderivative difference, so nearly
// vertical edges (t1 - t0 -> 0) lose no precision.
__device__ __forceinline__ double cell_mean(double t0, double t1) {
  const double span = t1 - t0;
  if (!(span > 0.0)) return fmin(fmax(1.0 - t0, 0.0), 1.0);
  const double la = fmax(fmin(t1, 0.0) - t0, 0.0);
  const double a = fmax(t0, 0.0), b = fmin(t1, 1.0);
  const double lb = fmax(b - a, 0.0);
  return (la + lb * (1.0 - 0.5 * (a + b))) / span;
}

__global__ void k_edges(const EdgeDev* __restrict__ edges, int64_t n_edges,
                        const BoxDev* __restrict__ boxes, unsigned long long* __restrict__ buf) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const EdgeDev ed = edges[e];
  if (ed.v0 == ed.v1) return;  // horizontal: no contribution
  const BoxDev bx = boxes[ed.shape];
  const bool up = ed.v1 > ed.v0;
  const double sgn = (up ? -1.0 : 1.0) * (double)ed.sign;
  const double va = up ? ed.v0 : ed.v1, vb = up ? ed.v1 : ed.v0;
  const double ua = up ? ed.u0 : ed.u1, ub = up ? ed.u1 : ed.u0;
  const double m = (ub - ua) / (vb - va);  // du/dv
  int jlo = (int)floor(va), jhi = (int)ceil(vb) - 1;
  jlo = max(jlo, bx.j0);
  jhi = min(jhi, bx.j0 + bx.h - 1);
  for (int j = jlo; j <= jhi; ++j) {
    const double vlo = fmax(va, (double)j), vhi = fmin(vb, (double)j + 1.0);
    const double h = vhi - vlo;
    if (!(h > 0.0)) continue;
    // end points of the piece inside this row band (exact at the edge's own ends)
    const double p = (vlo == va) ? ua : ua + m * (vlo - va);
    const double q = (vhi == vb) ? ub : ua + m * (vhi - va);
    const double umin = fmin(p, q) - bx.i0, umax = fmax(p, q) - bx.i0;
    unsigned long long* row = buf + bx.off + (int64_t)(j - bx.j0) * bx.w;
    if (umax <= 0.0) {  // entirely left of the box: full height for every cell
      atomicAdd(row, (unsigned long long)llrint(sgn * h * kFix));
      continue;
    }
    if (umin >= (double)bx.w) continue;  // entirely right of it
    const int ia = max((int)floor(umin), 0);
    const int ib = min((int)floor(umax), bx.w - 1);
    double prev = 0.0;
    for (int i = ia; i <= ib + 1 && i < bx.w; ++i) {
      const double F = (i > ib) ? h : h * cell_mean(umin - i, umax - i);
      const double d = F - prev;
      prev = F;
      if (d != 0.0) atomicAdd(row + i, (unsigned long long)llrint(sgn * d * kFix));
    }
  }
}

// rows[r] = {box index}; one warp per row.  In place: fixed-point differences ->
// coverage as double bits.  row_count[r] = entries kept.
__global__ void k_scan(const int32_t* __restrict__ row_box, int64_t n_rows,
                       const BoxDev* __restrict__ boxes, unsigned long long* __restrict__ buf,
                       int32_t* __restrict__ row_count) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const BoxDev bx = boxes[row_box[r]];
  unsigned long long* row = buf + bx.off + (r - bx.row0) * bx.w;
  long long carry = 0;
  int cnt = 0;
  for (int base = 0; base < bx.w; base += 32) {
    const int i = base + lane;
    long long x = i < bx.w ? (long long)row[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    x += carry;
    carry = __shfl_sync(0xffffffffu, x, 31);
    if (i < bx.w) {
      double c = (double)x * (1.0 / kFix);
      c = fmin(fmax(c, 0.0), 1.0);
      if (!(c > kKeep)) c = 0.0;
      row[i] = (unsigned long long)__double_as_longlong(c);
      cnt += c > 0.0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) row_count[r] = cnt;
}

__global__ void k_emit(const int32_t* __restrict__ row_box, int64_t n_rows,
                       const BoxDev* __restrict__ boxes, const unsigned long long* __restrict__ buf,
                       const int64_t* __restrict__ row_out, int nx, int32_t* __restrict__ indices,
                       double* __restrict__ data) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const BoxDev bx = boxes[row_box[r]];
  const int j = (int)(r - bx.row0);
  const unsigned long long* row = buf + bx.off + (int64_t)j * bx.w;
  int64_t out = row_out[r];
  for (int base = 0; base < bx.w; base += 32) {
    const int i = base + lane;
    const double c = i < bx.w ? __longlong_as_double((long long)row[i]) : 0.0;
    const unsigned keep = __ballot_sync(0xffffffffu, c > 0.0);
    if (c > 0.0) {
      const int64_t o = out + __popc(keep & ((1u << lane) - 1u));
      indices[o] = (bx.j0 + j) * nx + bx.i0 + i;
      data[o] = c;
    }
    out += __popc(keep);
  }
}

}  // namespace atl

using namespace atl;

struct AtlIndicator {
  int32_t n_shapes = 0, ny = 0, nx = 0;
  std::vector<int64_t> indptr;
  std::vector<int32_t> indices;
  std::vector<double> data;
};

namespace {

struct HostBox {
  int32_t i0 = 0, j0 = 0, w = 0, h = 0;
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

// One batch of shapes [s_beg, s_end): all of their box buffers fit the scratch budget.
int run_batch(AtlIndicator* R, int s_beg, int s_end, const std::vector<HostBox>& hb,
              const std::vector<EdgeDev>& all_edges, const std::vector<int64_t>& shape_edge_ptr) {
  const int nb = s_end - s_beg;
  std::vector<BoxDev> boxes(nb);
  std::vector<int32_t> row_box;
  int64_t off = 0;
  for (int k = 0; k < nb; ++k) {
    const HostBox& b = hb[s_beg + k];
    boxes[k] = BoxDev{b.i0, b.j0, b.w, b.h, off, (int64_t)row_box.size()};
    off += (int64_t)b.w * b.h;
    row_box.insert(row_box.end(), b.h, k);
  }
  const int64_t n_rows = (int64_t)row_box.size();
  const int64_t e_beg = shape_edge_ptr[s_beg], e_end = shape_edge_ptr[s_end];
  std::vector<int64_t> counts_cum(n_rows + 1, 0);
  if (n_rows > 0 && off > 0) {
    std::vector<EdgeDev> edges(all_edges.begin() + e_beg, all_edges.begin() + e_end);
    for (auto& e : edges) e.shape -= s_beg;
    DevBuf<EdgeDev> d_edges;
    DevBuf<BoxDev> d_boxes;
    DevBuf<int32_t> d_row_box, d_row_count, d_indices;
    DevBuf<unsigned long long> d_buf;
    DevBuf<int64_t> d_row_out;
    DevBuf<double> d_data;
    ATL_CUDA(d_edges.alloc(edges.size()));
    ATL_CUDA(d_boxes.alloc(nb));
    ATL_CUDA(d_row_box.alloc(n_rows));
    ATL_CUDA(d_row_count.alloc(n_rows));
    ATL_CUDA(d_row_out.alloc(n_rows));
    ATL_CUDA(d_buf.alloc(off));
    ATL_CUDA(cudaMemcpy(d_edges.p, edges.data(), edges.size() * sizeof(EdgeDev), cudaMemcpyHostToDevice));
    ATL_CUDA(cudaMemcpy(d_boxes.p, boxes.data(), nb * sizeof(BoxDev), cudaMemcpyHostToDevice));
    ATL_CUDA(cudaMemcpy(d_row_box.p, row_box.data(), n_rows * 4, cudaMemcpyHostToDevice));
    ATL_CUDA(cudaMemset(d_buf.p, 0, off * 8));
    const int T = 128;
    if (!edges.empty()) {
      k_edges<<<(unsigned)((edges.size() + T - 1) / T), T>>>(d_edges.p, (int64_t)edges.size(), d_boxes.p, d_buf.p);
      ++g_launches;
    }
    const unsigned row_blocks = (unsigned)((n_rows * 32 + T - 1) / T);
    k_scan<<<row_blocks, T>>>(d_row_box.p, n_rows, d_boxes.p, d_buf.p, d_row_count.p);
    ++g_launches;
    ATL_CUDA(cudaGetLastError());
    std::vector<int32_t> counts(n_rows);
    ATL_CUDA(cudaMemcpy(counts.data(), d_row_count.p, n_rows * 4, cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r < n_rows; ++r) counts_cum[r + 1] = counts_cum[r] + counts[r];
    const int64_t nnz = counts_cum[n_rows];
    if (nnz > 0) {
      ATL_CUDA(d_indices.alloc(nnz));
      ATL_CUDA(d_data.alloc(nnz));
      ATL_CUDA(cudaMemcpy(d_row_out.p, counts_cum.data(), n_rows * 8, cudaMemcpyHostToDevice));
      k_emit<<<row_blocks, T>>>(d_row_box.p, n_rows, d_boxes.p, d_buf.p, d_row_out.p, R->nx, d_indices.p,
                                d_data.p);
      ++g_launches;
      ATL_CUDA(cudaGetLastError());
      const size_t base = R->indices.size();
      R->indices.resize(base + nnz);
      R->data.resize(base + nnz);
      ATL_CUDA(cudaMemcpy(R->indices.data() + base, d_indices.p, nnz * 4, cudaMemcpyDeviceToHost));
      ATL_CUDA(cudaMemcpy(R->data.data() + base, d_data.p, nnz * 8, cudaMemcpyDeviceToHost));
    }
  }
  // CSR row pointers of the batch's shapes
  for (int k = 0; k < nb; ++k) {
    const int64_t r0 = boxes[k].row0, r1 = r0 + boxes[k].h;
    R->indptr[s_beg + k + 1] = R->indptr[s_beg + k] + (counts_cum[r1] - counts_cum[r0]);
  }
  return ATL_OK;
}

}  // namespace

extern "C" {

int atl_indicator_compute(int device, int32_t ny, int32_t nx, double x0, double dx, double y0,
                          double dy, int32_t n_shapes, const int64_t* shape_ring_ptr,
                          const int64_t* ring_ptr, const int8_t* ring_is_hole, const double* xy,
                          AtlIndicator** out) {
  ATL_REQUIRE(out, "NULL argument");
  *out = nullptr;
  ATL_REQUIRE(ny > 0 && nx > 0 && (int64_t)ny * nx < (1LL << 31), "bad grid");
  ATL_REQUIRE(dx > 0 && dy > 0, "cell sizes must be positive (coordinates ascending)");
  ATL_REQUIRE(n_shapes >= 0 && shape_ring_ptr && (n_shapes == 0 || (ring_ptr && ring_is_hole && xy)),
              "NULL argument");
  ATL_CUDA(cudaSetDevice(device));
  const double uo = x0 - 0.5 * dx, vo = y0 - 0.5 * dy;  // lower-left corner of cell (0, 0)

  std::vector<EdgeDev> edges;
  std::vector<int64_t> shape_edge_ptr(n_shapes + 1, 0);
  std::vector<HostBox> hb(n_shapes);
  for (int s = 0; s < n_shapes; ++s) {
    double ulo = INFINITY, uhi = -INFINITY, vlo = INFINITY, vhi = -INFINITY;
    ATL_REQUIRE(shape_ring_ptr[s + 1] >= shape_ring_ptr[s], "shape_ring_ptr must be non-decreasing");
    for (int64_t r = shape_ring_ptr[s]; r < shape_ring_ptr[s + 1]; ++r) {
      const int64_t a = ring_ptr[r], b = ring_ptr[r + 1];
      ATL_REQUIRE(b >= a, "ring_ptr must be non-decreasing");
      if (b - a < 3) continue;  // degenerate ring: no area
      // orientation from the shoelace sum in grid units (about the first vertex)
      double area2 = 0.0;
      const double ux0 = (xy[2 * a] - uo) / dx, vy0 = (xy[2 * a + 1] - vo) / dy;
      for (int64_t k = a; k < b; ++k) {
        const int64_t k1 = (k + 1 < b) ? k + 1 : a;
        const double u0 = (xy[2 * k] - uo) / dx - ux0, v0 = (xy[2 * k + 1] - vo) / dy - vy0;
        const double u1 = (xy[2 * k1] - uo) / dx - ux0, v1 = (xy[2 * k1 + 1] - vo) / dy - vy0;
        ATL_REQUIRE(std::isfinite(u0) && std::isfinite(v0), "non-finite vertex");
        area2 += u0 * v1 - u1 * v0;
      }
      if (area2 == 0.0) continue;
      const int32_t sign = (ring_is_hole[r] ? -1 : 1) * (area2 > 0.0 ? 1 : -1);
      for (int64_t k = a; k < b; ++k) {
        const int64_t k1 = (k + 1 < b) ? k + 1 : a;
        EdgeDev e;
        e.u0 = (xy[2 * k] - uo) / dx;
        e.v0 = (xy[2 * k + 1] - vo) / dy;
        e.u1 = (xy[2 * k1] - uo) / dx;
        e.v1 = (xy[2 * k1 + 1] - vo) / dy;
        e.shape = s;
        e.sign = sign;
        if (!ring_is_hole[r]) {
          ulo = std::min(ulo, e.u0);
          uhi = std::max(uhi, e.u0);
          vlo = std::min(vlo, e.v0);
          vhi = std::max(vhi, e.v0);
        }
        if (e.v0 != e.v1 || e.u0 != e.u1) edges.push_back(e);
      }
    }
    shape_edge_ptr[s + 1] = (int64_t)edges.size();
    HostBox& b = hb[s];
    if (uhi > ulo && vhi > vlo) {
      const double i0 = std::max(std::floor(ulo), 0.0), i1 = std::min(std::ceil(uhi), (double)nx);
      const double j0 = std::max(std::floor(vlo), 0.0), j1 = std::min(std::ceil(vhi), (double)ny);
      if (i1 > i0 && j1 > j0) {
        b.i0 = (int32_t)i0;
        b.j0 = (int32_t)j0;
        b.w = (int32_t)(i1 - i0);
        b.h = (int32_t)(j1 - j0);
      }
    }
  }

  AtlIndicator* R = new AtlIndicator();
  R->n_shapes = n_shapes;
  R->ny = ny;
  R->nx = nx;
  R->indptr.assign((size_t)n_shapes + 1, 0);
  int64_t budget = 64LL << 20;  // box cells per batch (512 MiB of scratch)
  if (const char* b = getenv("ATL_INDICATOR_BUDGET")) budget = std::max<int64_t>(atoll(b), 1);  // tests
  int rc = ATL_OK;
  for (int s = 0; s < n_shapes && rc == ATL_OK;) {
    int e = s;
    int64_t cells = 0;
    while (e < n_shapes && (e == s || cells + (int64_t)hb[e].w * hb[e].h <= budget)) {
      cells += (int64_t)hb[e].w * hb[e].h;
      ++e;
    }
    rc = run_batch(R, s, e, hb, edges, shape_edge_ptr);
    s = e;
  }
  if (rc == ATL_OK) {
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) rc = cuda_fail(ce, "atl_indicator_compute");
  }
  if (rc != ATL_OK) {
    delete R;
    return rc;
  }
  *out = R;
  return ATL_OK;
}

int atl_indicator_nnz(const AtlIndicator* ind, int64_t* nnz_out) {
  ATL_REQUIRE(ind && nnz_out, "NULL argument");
  *nnz_out = (int64_t)ind->indices.size();
  return ATL_OK;
}

int atl_indicator_export(const AtlIndicator* ind, int64_t* indptr_out, int32_t* indices_out,
                         double* data_out) {
  ATL_REQUIRE(ind && indptr_out, "NULL argument");
  ATL_REQUIRE(ind->indices.empty() || (indices_out && data_out), "NULL argument");
  std::memcpy(indptr_out, ind->indptr.data(), ind->indptr.size() * 8);
  if (!ind->indices.empty()) {
    std::memcpy(indices_out, ind->indices.data(), ind->indices.size() * 4);
    std::memcpy(data_out, ind->data.data(), ind->data.size() * 8);
  }
  return ATL_OK;
}

void atl_indicator_destroy(AtlIndicator* ind) { delete ind; }

}  // extern "C"
