// Mesh generation, symbolic pattern and numeric assembly kernels.
//
// Design (DESIGN.md "Assembly"): the assembly is ROW-OWNED.  One thread owns one matrix row,
// evaluates the (at most two) element tensors that touch it from the vertex coordinates, and
// stages the row's entries in shared memory at its offset inside the block's CSR tile; the block
// then streams the tile to HBM with fully coalesced stores.  No atomics, every stored value is
// written exactly once, and a value that has two cell contributions (the mass-matrix diagonal at a
// vertex shared by two cells of one graph edge) is the IEEE sum of two operands, so the result is
// bit-reproducible and independent of any scheduling.
//
// Forms restated here: assembly.py:253-255 (mass, +-grad . tangent), :258-262 (right-hand side),
// :271-277 (multiplier coupling); explicit zeros of the multiplier blocks as stored by DOLFINx.
// All arithmetic uses the _rn intrinsics so that ptxas cannot contract a*b+c into an FMA: the
// values are bit-identical to the NumPy oracle.
#pragma once

#include "ctx.cuh"

namespace nxfx {

struct Net {
  int32_t n_nodes, E, N, n_bif;
  int32_t nq, poff, loff, ndofs;
  const int4* __restrict__ slot_uvl;
  const int32_t* __restrict__ slot_edge;
  const int32_t* __restrict__ edge_slot;
  const int32_t* __restrict__ bif_ptr;
  const int32_t* __restrict__ bif_inc;
  const double2* __restrict__ x2;  // [nv] vertex records {x, y | z, p_bc} (32 B each)
};

struct Coef {
  const double* __restrict__ R_cell;  // [nc] or null
  const double* __restrict__ f_cell;  // [nc] or null
  double R_const, f_const;
  double* __restrict__ cell_rh;  // [nc] out
};

constexpr double kThird = 1.0 / 3.0;
constexpr double kSixth = 1.0 / 6.0;
constexpr int kMaxFluxRow = 7;  // 3 mass + 2 pressure + 2 multiplier entries (flux degree 1)

__device__ __forceinline__ int vertex_id(const Net& g, int e, int u, int v, int a) {
  // mesh.py:276,292: graph nodes first, then N-1 interior points per edge in edge order
  return a == 0 ? u : (a == g.N ? v : g.n_nodes + e * (g.N - 1) + (a - 1));
}

struct VertexRec {
  double x, y, z, p;
};

// one 32-byte sector per vertex: coordinates and the boundary pressure travel together
__device__ __forceinline__ VertexRec load_vertex(const double2* __restrict__ x2, int v) {
  const double2 a = __ldg(x2 + 2 * (size_t)v), b = __ldg(x2 + 2 * (size_t)v + 1);
  return VertexRec{a.x, a.y, b.x, b.y};
}

__device__ __forceinline__ double seg_length(const VertexRec& v0, const VertexRec& v1) {
  const double dx = __dsub_rn(v1.x, v0.x), dy = __dsub_rn(v1.y, v0.y), dz = __dsub_rn(v1.z, v0.z);
  return __dsqrt_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// ---- vertices ------------------------------------------------------------------------------
// mesh.py:275,290: interior point k (1..N-1) of edge (u,v) = start*(1-w) + end*w, w = k*(1/N)
__global__ void __launch_bounds__(kThreads)
pad_nodes_kernel(int n_nodes, int gdim, const double* __restrict__ pos, double* __restrict__ x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes * 3) return;
  int node = i / 3, d = i - node * 3;
  x[4 * (size_t)node + d] = d < gdim ? pos[(size_t)node * gdim + d] : 0.0;
}

// p_bc interpolated into P1 (assembly.py:225-234) is stored in the 4th slot of the vertex records
__global__ void __launch_bounds__(kThreads)
set_pbc_kernel(int64_t nv, const double* __restrict__ pbc, double* __restrict__ x) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nv) x[4 * i + 3] = pbc[i];
}

__global__ void __launch_bounds__(kThreads)
interior_vertices_kernel(int n_nodes, int E, int N, const int32_t* __restrict__ eu,
                         const int32_t* __restrict__ ev, double* __restrict__ x) {
  const int64_t total = (int64_t)E * (N - 1) * 3;
  const double step = 1.0 / (double)N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t vert = i / 3;
    const int d = (int)(i - vert * 3);
    const int e = (int)(vert / (N - 1));
    const int k = (int)(vert - (int64_t)e * (N - 1)) + 1;
    const double w = __dmul_rn((double)k, step);
    const double s = x[4 * (size_t)eu[e] + d], t = x[4 * (size_t)ev[e] + d];
    x[4 * ((size_t)n_nodes + vert) + d] = __dadd_rn(__dmul_rn(s, __dsub_rn(1.0, w)), __dmul_rn(t, w));
  }
}

// ---- row decode ----------------------------------------------------------------------------
// Flux row `r` (edge slot, local vertex a): entries in ascending column order
//   [mass a-1, a, a+1] [pressure cells a-1, a] [multipliers of u and/or v]
// Column decode (symbolic phase); the numeric kernel below emits values in the same order.
__device__ __forceinline__ int flux_row_cols(const Net& g, int r, int* cols) {
  const int N = g.N, np1 = N + 1;
  const int slot = r / np1, a = r - slot * np1;
  const int4 t = g.slot_uvl[slot];
  const int e = g.slot_edge[slot];
  int n = 0;
  if (a > 0) cols[n++] = r - 1;
  cols[n++] = r;
  if (a < N) cols[n++] = r + 1;
  const int pb = g.poff + e * N;
  if (a > 0) cols[n++] = pb + a - 1;
  if (a < N) cols[n++] = pb + a;
  const bool hu = t.z >= 0 && a <= 1, hv = t.w >= 0 && a >= N - 1;
  if (hu && hv) {
    cols[n++] = g.loff + min(t.z, t.w);
    cols[n++] = g.loff + max(t.z, t.w);
  } else if (hu) {
    cols[n++] = g.loff + t.z;
  } else if (hv) {
    cols[n++] = g.loff + t.w;
  }
  return n;
}

__device__ __forceinline__ int flux_row_len(const Net& g, int r) {
  const int N = g.N, np1 = N + 1;
  const int slot = r / np1, a = r - slot * np1;
  const int4 t = g.slot_uvl[slot];
  const int inner = (a > 0) + (a < N);
  return 1 + 2 * inner + (t.z >= 0 && a <= 1) + (t.w >= 0 && a >= N - 1);
}

// ---- symbolic ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) row_len_kernel(Net g, int32_t* __restrict__ len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > g.ndofs) return;
  int n = 0;
  if (r < g.nq) n = flux_row_len(g, r);
  else if (r < g.loff) n = 2;
  else if (r < g.ndofs) n = 2 * (g.bif_ptr[r - g.loff + 1] - g.bif_ptr[r - g.loff]);
  len[r] = n;  // len[ndofs] = 0 so that the exclusive scan yields rowptr[ndofs] = nnz
}

__global__ void __launch_bounds__(kThreads)
fill_cols_kernel(Net g, const int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= g.ndofs) return;
  int32_t* out = colidx + rowptr[r];
  const int np1 = g.N + 1;
  if (r < g.nq) {
    int cols[kMaxFluxRow];
    const int n = flux_row_cols(g, r, cols);
    for (int i = 0; i < n; ++i) out[i] = cols[i];
  } else if (r < g.loff) {
    const int cell = r - g.poff, e = cell / g.N, j = cell - e * g.N;
    const int fb = g.edge_slot[e] * np1;
    out[0] = fb + j;
    out[1] = fb + j + 1;
  } else {
    const int i0 = g.bif_ptr[r - g.loff], i1 = g.bif_ptr[r - g.loff + 1];
    for (int k = i0; k < i1; ++k) {
      const int inc = g.bif_inc[k], e = inc >> 1;
      const int fb = g.edge_slot[e] * np1;
      const int c0 = (inc & 1) ? fb + g.N - 1 : fb;  // in-edge: last cell; out-edge: first cell
      out[2 * (k - i0)] = c0;
      out[2 * (k - i0) + 1] = c0 + 1;
    }
  }
}

// ---- numeric -------------------------------------------------------------------------------
// One thread per matrix row.  A flux-row thread (slot, a) owns cell a of its graph edge (a < N):
// it loads the two vertex records of that cell, evaluates m_a = R_a h_a and hands m_a and the
// boundary pressure of the cell's end vertex to the next row of the same edge by warp shuffle
// (the previous cell's contribution to the shared-vertex mass entry), so every cell is evaluated
// once per warp.  Entries are staged in shared memory at the row's CSR offset and streamed out
// coalesced.
template <bool ACC>
__global__ void __launch_bounds__(kTileRows)
assemble_rows_kernel(Net g, Coef c, const int32_t* __restrict__ rowptr, double* __restrict__ vals,
                     double* __restrict__ b, int lhs, int rhs) {
  __shared__ double sm[kTileCap];
  const int r0 = blockIdx.x * kTileRows;
  const int r = r0 + threadIdx.x;
  const int rend = min(r0 + kTileRows, g.ndofs);
  const int sbase = rowptr[r0];
  const int tnnz = rowptr[rend] - sbase;
  const int lane = threadIdx.x & 31;
  const int N = g.N;
  int p = r < g.ndofs ? rowptr[r] - sbase : 0;  // running position inside the tile
  auto put = [&](double v) {
    if (p < kTileCap) sm[p] = v;
    else if (lhs) {
      if (ACC) vals[(size_t)sbase + p] += v; else vals[(size_t)sbase + p] = v;
    }
    ++p;
  };
  // ---- flux rows: geometry phase (all lanes take part in the shuffles) ----------------------
  const bool isflux = r < g.nq;
  int a = 0, e = 0;
  int4 t = make_int4(0, 0, -1, -1);
  double mR = 0.0, pA = 0.0, pNext = 0.0;
  if (isflux) {
    const int np1 = N + 1;
    const int slot = r / np1;
    a = r - slot * np1;
    t = g.slot_uvl[slot];
    e = g.slot_edge[slot];
    if (a < N) {
      const VertexRec v0 = load_vertex(g.x2, vertex_id(g, e, t.x, t.y, a));
      const VertexRec v1 = load_vertex(g.x2, vertex_id(g, e, t.x, t.y, a + 1));
      const double R = c.R_cell ? c.R_cell[(size_t)e * N + a] : c.R_const;
      mR = __dmul_rn(R, seg_length(v0, v1));
      c.cell_rh[(size_t)e * N + a] = mR;
      pA = v0.p;
      pNext = v1.p;
    }
  }
  const double mPrev = __shfl_up_sync(0xffffffffu, mR, 1);
  const double pPrev = __shfl_up_sync(0xffffffffu, pNext, 1);
  if (isflux) {
    double mL = 0.0;
    if (a > 0) {
      if (lane > 0) {  // the previous lane owns cell a-1 of the same edge
        mL = mPrev;
        if (a == N) pA = pPrev;
      } else {  // first lane of the warp: evaluate cell a-1 here
        const VertexRec v0 = load_vertex(g.x2, vertex_id(g, e, t.x, t.y, a - 1));
        const VertexRec v1 = load_vertex(g.x2, vertex_id(g, e, t.x, t.y, a));
        const double R = c.R_cell ? c.R_cell[(size_t)e * N + a - 1] : c.R_const;
        mL = __dmul_rn(R, seg_length(v0, v1));
        pA = v1.p;
      }
    }
    // mass block (assembly.py:253): R h [[1/3,1/6],[1/6,1/3]] per cell
    if (a > 0) put(__dmul_rn(mL, kSixth));
    put(a == 0 ? __dmul_rn(mR, kThird)
               : (a == N ? __dmul_rn(mL, kThird)
                         : __dadd_rn(__dmul_rn(mL, kThird), __dmul_rn(mR, kThird))));
    if (a < N) put(__dmul_rn(mR, kSixth));
    // a[i][P] = -int p dv/ds: -B^T (assembly.py:255)
    if (a > 0) put(-1.0);
    if (a < N) put(1.0);
    // multiplier columns (assembly.py:273,277); the cell's other flux dof stores an explicit 0.0
    const bool hu = t.z >= 0 && a <= 1, hv = t.w >= 0 && a >= N - 1;
    const double vu = a == 0 ? -1.0 : 0.0, vv = a == N ? 1.0 : 0.0;
    if (hu && hv) {
      const bool ufirst = t.z < t.w;
      put(ufirst ? vu : vv);
      put(ufirst ? vv : vu);
    } else if (hu) {
      put(vu);
    } else if (hv) {
      put(vv);
    }
    if (rhs) {
      // assembly.py:258-260: +p_bc at in_marker vertices (boundary END of an edge), -p_bc at
      // out_marker vertices (boundary START of an edge)
      double bv = 0.0;
      if (a == 0 && t.z < 0) bv = -pA;
      if (a == N && t.w < 0) bv = pA;
      if (ACC) b[r] += bv; else b[r] = bv;
    }
  } else if (r < g.loff) {
    // a[P][i] = +int phi dq/ds: B = [-1, +1] (assembly.py:254);  L[P] = int f phi (assembly.py:262)
    put(-1.0);
    put(1.0);
    if (rhs) {
      double bv = 0.0;
      if (c.f_cell || c.f_const != 0.0) {
        const int cell = r - g.poff, ce = cell / N, j = cell - ce * N;
        const int4 ct = g.slot_uvl[g.edge_slot[ce]];
        const double h = seg_length(load_vertex(g.x2, vertex_id(g, ce, ct.x, ct.y, j)),
                                    load_vertex(g.x2, vertex_id(g, ce, ct.x, ct.y, j + 1)));
        bv = __dmul_rn(c.f_cell ? c.f_cell[cell] : c.f_const, h);
      }
      if (ACC) b[r] += bv; else b[r] = bv;
    }
  } else if (r < g.ndofs) {
    // a[LM][c] = +mu q at in-edges, -mu q at out-edges (assembly.py:272,276)
    const int i0 = g.bif_ptr[r - g.loff], i1 = g.bif_ptr[r - g.loff + 1];
    for (int k = i0; k < i1; ++k) {
      const bool in = g.bif_inc[k] & 1;
      put(in ? 0.0 : -1.0);
      put(in ? 1.0 : 0.0);
    }
    if (rhs && !ACC) b[r] = 0.0;
  }
  __syncthreads();
  if (lhs) {
    const int m = min(tnnz, kTileCap);
    for (int i = threadIdx.x; i < m; i += kTileRows) {
      if (ACC) vals[(size_t)sbase + i] += sm[i]; else vals[(size_t)sbase + i] = sm[i];
    }
  }
}

// post_processing.py:36-51: per cell the two flux dofs of its edge
__global__ void __launch_bounds__(kThreads)
global_flux_kernel(Net g, const double* __restrict__ x, double* __restrict__ out) {
  const int64_t total = (int64_t)g.E * g.N;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < total;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(cell / g.N), j = (int)(cell - (int64_t)e * g.N);
    const int q0 = g.edge_slot[e] * (g.N + 1) + j;
    out[2 * cell] = x[q0];
    out[2 * cell + 1] = x[q0 + 1];
  }
}

}  // namespace nxfx
