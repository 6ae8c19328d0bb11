// Table-driven assembly for higher-order elements (flux P_fd, pressure P_pd): every stored matrix
// entry gathers its (at most two) cell contributions, every right-hand-side row its source list,
// in a fixed order -- deterministic, no atomics.  The lists are built on the host (generic.py).
#pragma once

#include "assemble.cuh"

namespace nxfx {

constexpr int kRhFlag = 1 << 30;  // value = coef * R*h of the cell (else: coef)
constexpr int kVertexFlag = 1 << 30;

// h of every cell (cells run u -> v along their graph edge, mesh.py:293-309)
__global__ void __launch_bounds__(kThreads)
cell_length_kernel(Net g, double* __restrict__ cell_h) {
  const int64_t nc = (int64_t)g.E * g.N;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(c / g.N), j = (int)(c - (int64_t)e * g.N);
    const int4 t = g.slot_uvl[g.edge_slot[e]];
    cell_h[c] = seg_length(load_vertex(g.x2, vertex_id(g, e, t.x, t.y, j)),
                           load_vertex(g.x2, vertex_id(g, e, t.x, t.y, j + 1)));
  }
}

template <bool ACC>
__global__ void __launch_bounds__(kThreads)
assemble_generic_kernel(int64_t nnz, const int2* __restrict__ src_id, const double2* __restrict__ src_coef,
                        const double* __restrict__ cell_h, const double* __restrict__ R_cell, double R_const,
                        double* __restrict__ vals) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
    const int2 id = src_id[k];
    const double2 co = src_coef[k];
    auto term = [&](int s, double c) {
      if (s < 0) return 0.0;
      if (!(s & kRhFlag)) return c;
      const int cell = s & (kRhFlag - 1);
      return __dmul_rn(__dmul_rn(R_cell ? R_cell[cell] : R_const, cell_h[cell]), c);
    };
    double v = term(id.x, co.x);
    if (id.y >= 0) v = __dadd_rn(v, term(id.y, co.y));
    if (ACC) vals[k] += v; else vals[k] = v;
  }
}

template <bool ACC>
__global__ void __launch_bounds__(kThreads)
rhs_generic_kernel(int n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ id,
                   const double* __restrict__ coef, const double* __restrict__ cell_h,
                   const double* __restrict__ f_cell, double f_const, const double2* __restrict__ x2,
                   double* __restrict__ b) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double s = 0.0;
  for (int k = ptr[r]; k < ptr[r + 1]; ++k) {
    const int i = id[k];
    double t;
    if (i & kVertexFlag) t = __dmul_rn(coef[k], x2[2 * (size_t)(i & (kVertexFlag - 1)) + 1].y);  // p_bc
    else t = __dmul_rn(__dmul_rn(f_cell ? f_cell[i] : f_const, cell_h[i]), coef[k]);
    s = __dadd_rn(s, t);
  }
  if (ACC) b[r] += s; else b[r] = s;
}

}  // namespace nxfx
