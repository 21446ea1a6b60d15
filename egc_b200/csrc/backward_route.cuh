// Deterministic routing of the min/max gradients (EGC_BWD_DETERMINISTIC): a compare-and-add gather over the CSC.
//
// The default routing kernel (k_route_minmax, backward_pass1.cuh) walks the target rows and adds every slot
// gradient to its single winning source with fp32 atomics - the order of the additions into one d_bases row changes
// from run to run, like the reference's own `scatter_add_` backward (the reference warns about it,
// /root/reference/hyperparameters.md:3).  Here every SOURCE column j walks its CSC entries in order; entry e
// (target i = rowidx[e], CSR position csr2csc[e]) receives feature f of slot s iff the forward recorded that position
// as the winner, saved_arg[i][s][f] == csr2csc[e] (first-wins ties and duplicate edges are decided by the position,
// exactly as in the forward).  The sum over a column's entries runs in CSC order, chunk partials of long columns are
// merged in chunk order: bit-reproducible, no atomics.  Cost: two gathered rows (arg + gradient) per entry and slot.
// Included by aggregate_api.cu only.
#pragma once

#include "aggregate_fast.cuh"

namespace egc {

struct RouteCscParams {
  const int32_t* colptr;
  const int32_t* rowidx;
  const int32_t* csr2csc;       // CSR position of every CSC entry
  const float* csc_val_lin;     // CSC order, or null
  int n_cols;
  int n_long, n_chunks;
  const int32_t* long_rows;
  const int32_t* long_chunk_ptr;
  const int32_t* chunk_row;
  const int32_t* chunk_begin;
  float* partials;              // [n_chunks][BD]
  const int32_t* saved_arg;     // [n_dst][n_arg][BD]
  const float* t_route;         // [n_dst][n_arg][BD]
  float* d_bases;               // [n_cols][BD], accumulated into
  int n_arg, BD, nvec, G, n_pass;
  int mode;                     // 0: chunks of long columns + normal columns, 1: merge of the long columns
};

template <int VEC>
__global__ void __launch_bounds__(kAggThreads, 4) k_route_csc(const __grid_constant__ RouteCscParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kAggWarps + warp;
  int colj, begin = 0, end = 0, chunk_id = -1, long_idx = -1;
  if (p.mode == 0) {
    if (gw < p.n_chunks) {
      chunk_id = gw;
      colj = p.chunk_row[gw];
      begin = p.chunk_begin[gw];
      end = min(begin + EGC_CHUNK_EDGES, p.colptr[colj + 1]);
    } else {
      colj = gw - p.n_chunks;
      if (colj >= p.n_cols) return;
      begin = p.colptr[colj];
      end = p.colptr[colj + 1];
      if (end - begin > EGC_CHUNK_EDGES) return;                 // long column: its chunks + the merge launch
    }
  } else {
    long_idx = gw;
    if (long_idx >= p.n_long) return;
    colj = p.long_rows[long_idx];
  }
  const int G = p.G, NG = 32 / G, g = lane / G;
  const int64_t row_stride = static_cast<int64_t>(p.n_arg) * p.BD;
  constexpr int U = 4;                                           // entries in flight per lane group

  for (int pass = 0; pass < p.n_pass; ++pass) {
    const int piece = pass * 32 + (lane & (G - 1));
    const bool active = piece < p.nvec;
    const int foff = min(piece, p.nvec - 1) * VEC;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;

    if (p.mode == 0) {
      for (int e0 = begin + g; e0 < end; e0 += U * NG) {         // group g takes entries g, g + NG, ... in order
        int pos[U];
        float lw[U];
        int64_t base[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int e = e0 + u * NG;
          const bool ok = e < end;
          const int ec = ok ? e : end - 1;
          pos[u] = ok ? __ldg(p.csr2csc + ec) : -2;              // -2 never equals a recorded position (>= 0 or -1)
          lw[u] = p.csc_val_lin != nullptr ? __ldg(p.csc_val_lin + ec) : 1.f;
          base[u] = static_cast<int64_t>(__ldg(p.rowidx + ec)) * row_stride + foff;
        }
        for (int s = 0; s < p.n_arg; ++s) {
          int a[U][VEC];
          float v[U][VEC];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t o = base[u] + static_cast<int64_t>(s) * p.BD;
            if constexpr (VEC == 4) {
              const int4 t = __ldg(reinterpret_cast<const int4*>(p.saved_arg + o));
              a[u][0] = t.x; a[u][1] = t.y; a[u][2] = t.z; a[u][3] = t.w;
            } else {
              a[u][0] = __ldg(p.saved_arg + o);
            }
            ld_row<VEC>(v[u], p.t_route + o);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
              if (a[u][k] == pos[u]) acc[k] = __fadd_rn(acc[k], __fmul_rn(v[u][k], lw[u]));
          }
        }
      }
      for (int off = G; off < 32; off <<= 1) {                   // fixed merge order of the lane groups
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = __fadd_rn(acc[k], __shfl_xor_sync(kFull, acc[k], off));
      }
    } else {
      const int c0 = p.long_chunk_ptr[long_idx], c1 = p.long_chunk_ptr[long_idx + 1];
      for (int c = c0; c < c1; ++c) {                            // chunk order
        float t[VEC];
        ld_plain<VEC>(t, p.partials + static_cast<int64_t>(c) * p.BD + foff);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = __fadd_rn(acc[k], t[k]);
      }
    }

    if (!(active && lane < G)) continue;
    if (chunk_id >= 0) {
      st_row<VEC>(p.partials + static_cast<int64_t>(chunk_id) * p.BD + foff, acc);
      continue;
    }
    float* dst = p.d_bases + static_cast<int64_t>(colj) * p.BD + foff;
    float r[VEC];
    ld_plain<VEC>(r, dst);
#pragma unroll
    for (int k = 0; k < VEC; ++k) r[k] = __fadd_rn(r[k], acc[k]);
    st_row<VEC>(dst, r);
  }
}

static int launch_route_csc(RouteCscParams p, bool vec4, cudaStream_t st) {
  for (int mode = 0; mode < 2; ++mode) {
    p.mode = mode;
    const int64_t tasks = mode == 0 ? static_cast<int64_t>(p.n_chunks) + p.n_cols : p.n_long;
    if (tasks <= 0) continue;
    {
      LaunchScope ls(mode ? "k_route_csc_merge" : "k_route_csc", st);
      if (vec4) k_route_csc<4><<<ceil_div(tasks, kAggWarps), kAggThreads, 0, st>>>(p);
      else k_route_csc<1><<<ceil_div(tasks, kAggWarps), kAggThreads, 0, st>>>(p);
    }
    EGC_LAUNCH_CHECK("k_route_csc");
  }
  return EGC_OK;
}

}  // namespace egc
