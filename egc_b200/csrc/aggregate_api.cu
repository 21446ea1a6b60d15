// extern "C" entry points of the fused aggregation forward / backward, the source-side (CSC)
// backward pass and the column-sum helper.
#include <algorithm>

#include "aggregate.cuh"
#include "colsum.cuh"

namespace egc {

int fill_agg_params(AggParams& p, const egc_layer_desc& d, bool vec4, bool bwd) {
  p.H = d.heads; p.B = d.bases; p.D = d.dim; p.A = d.n_aggr;
  p.BD = d.bases * d.dim; p.HD = d.heads * d.dim; p.AB = d.n_aggr * d.bases; p.HAB = d.heads * p.AB;
  for (int a = 0; a < EGC_MAX_AGGR; ++a) p.aggr[a] = a < d.n_aggr ? d.aggr[a] : -1;
  p.sigmoid = d.sigmoid;
  const int vec = vec4 ? 4 : 1;
  p.nvec = (p.BD + vec - 1) / vec;
  int g = 1;
  while (g < 32 && g < p.nvec) g <<= 1;
  p.G = g;
  p.n_pass = p.nvec > 32 ? (p.nvec + 31) / 32 : 1;
  auto a4 = [](int v) { return (v + 3) & ~3; };
  int off = 0;
  p.sm_agg = off; off = a4(off + p.A * p.BD);
  p.sm_w = off; off = a4(off + p.HAB);
  p.sm_g = p.sm_mean = p.sm_var = p.sm_amx = p.sm_amn = 0;
  if (bwd) {
    p.sm_g = off; off = a4(off + p.HD);
    p.sm_mean = off; off = a4(off + p.BD);
    p.sm_var = off; off = a4(off + p.BD);
    p.sm_amx = off; off = a4(off + p.BD);
    p.sm_amn = off; off = a4(off + p.BD);
  }
  p.sm_per_warp = off;
  return off * kAggWarps * static_cast<int>(sizeof(float));
}

static int validate_desc(const egc_layer_desc* d, const char* who) {
  EGC_REQUIRE(d != nullptr, "%s: null descriptor", who);
  EGC_REQUIRE(d->n_dst > 0 && d->n_src > 0, "%s: n_dst=%d n_src=%d", who, d->n_dst, d->n_src);
  EGC_REQUIRE(d->heads > 0 && d->bases > 0 && d->dim > 0, "%s: heads=%d bases=%d dim=%d", who, d->heads, d->bases, d->dim);
  EGC_REQUIRE(d->n_aggr >= 1 && d->n_aggr <= EGC_MAX_AGGR, "%s: n_aggr=%d", who, d->n_aggr);
  for (int a = 0; a < d->n_aggr; ++a)
    EGC_REQUIRE(d->aggr[a] >= EGC_AGGR_SUM && d->aggr[a] <= EGC_AGGR_STD, "%s: unknown aggregator code %d", who, d->aggr[a]);
  return EGC_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static void set_plan(AggParams& p, const egc_row_plan* plan) {
  p.n_long = plan ? plan->n_long : 0;
  p.n_chunks = plan ? plan->n_chunks : 0;
  p.long_rows = plan ? plan->long_rows : nullptr;
  p.long_chunk_ptr = plan ? plan->long_chunk_ptr : nullptr;
  p.chunk_row = plan ? plan->chunk_row : nullptr;
  p.chunk_begin = plan ? plan->chunk_begin : nullptr;
}

static int validate_plan(const egc_row_plan* plan, const char* who) {
  if (plan == nullptr || plan->n_long == 0) return EGC_OK;
  EGC_REQUIRE(plan->n_long > 0 && plan->n_chunks >= 2 * plan->n_long, "%s: inconsistent row plan", who);
  EGC_REQUIRE(plan->long_rows && plan->long_chunk_ptr && plan->chunk_row && plan->chunk_begin, "%s: row plan with null arrays", who);
  return EGC_OK;
}

// ---------------------------------------------------------------------------------------------
// source-side (CSC) backward pass: d_bases[j] = sum over column j of the target-side streams
// ---------------------------------------------------------------------------------------------
struct ScatterParams {
  const int32_t* colptr;
  const int32_t* rowidx;
  const float* val_sym;     // CSC order
  const float* val_lin;     // CSC order
  int n_cols;
  int n_long, n_chunks;
  const int32_t* long_rows;
  const int32_t* long_chunk_ptr;
  const int32_t* chunk_row;
  const int32_t* chunk_begin;
  float* partials;          // [n_chunks][n_ts][BD]
  const float* tstreams;    // [n_dst, n_ts, BD]
  const float* bases;       // [n_cols, BD]
  float* d_bases;           // [n_cols, BD]
  int n_ts, ts_sym, ts_lin, ts_sq;
  int BD, nvec, G, n_pass;
  int routed;               // d_bases already holds atomically routed min/max gradients
  int mode;
};

constexpr int kScatterUnroll = 4;

template <int TSMASK, int VEC, bool LINW>
__global__ void __launch_bounds__(kAggThreads) k_scatter_bwd(const __grid_constant__ ScatterParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kAggWarps + warp;
  int colj, begin, end, chunk_id = -1, long_idx = -1;
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
      if (end - begin > EGC_CHUNK_EDGES) return;
    }
  } else {
    long_idx = gw;
    if (long_idx >= p.n_long) return;
    colj = p.long_rows[long_idx];
    begin = p.colptr[colj];
    end = p.colptr[colj + 1];
  }
  const int G = p.G, NG = 32 / G, g = lane / G;
  const int64_t row_stride = static_cast<int64_t>(p.n_ts) * p.BD;

  for (int pass = 0; pass < p.n_pass; ++pass) {
    const int piece = pass * 32 + (lane & (G - 1));
    const bool active = piece < p.nvec;
    const int foff = piece * VEC;
    float a_sym[VEC], a_lin[VEC], a_sq[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { a_sym[k] = 0.f; a_lin[k] = 0.f; a_sq[k] = 0.f; }

    if (p.mode == 0) {
      const float* __restrict__ src = p.tstreams + foff;
      for (int e0 = begin; e0 < end; e0 += 32) {
        const int n_here = min(32, end - e0);
        const bool have = lane < n_here;
        const int my_row = have ? __ldg(p.rowidx + e0 + lane) : 0;
        float my_vs = 0.f, my_vl = 0.f;
        if constexpr (TSMASK & 1) my_vs = have ? __ldg(p.val_sym + e0 + lane) : 0.f;
        if constexpr (LINW) my_vl = have ? __ldg(p.val_lin + e0 + lane) : 0.f;
        const int steps = (n_here + NG - 1) / NG;
        for (int s = 0; s < steps; s += kScatterUnroll) {
          float xs[kScatterUnroll][VEC], xl[kScatterUnroll][VEC], xq[kScatterUnroll][VEC];
          bool ok[kScatterUnroll];
#pragma unroll
          for (int u = 0; u < kScatterUnroll; ++u) {
            const int idx = (s + u) * NG + g;
            const int i = __shfl_sync(kFull, my_row, idx & 31);
            ok[u] = active && (s + u) < steps && idx < n_here;
            if (ok[u]) {
              const float* r = src + static_cast<int64_t>(i) * row_stride;
              if constexpr (TSMASK & 1) ld_row<VEC>(xs[u], r + p.ts_sym * p.BD);
              if constexpr (TSMASK & 2) ld_row<VEC>(xl[u], r + p.ts_lin * p.BD);
              if constexpr (TSMASK & 4) ld_row<VEC>(xq[u], r + p.ts_sq * p.BD);
            }
          }
#pragma unroll
          for (int u = 0; u < kScatterUnroll; ++u) {
            const int idx = (s + u) * NG + g;
            float vs = 0.f, vl = 1.f;
            if constexpr (TSMASK & 1) vs = __shfl_sync(kFull, my_vs, idx & 31);
            if constexpr (LINW) vl = __shfl_sync(kFull, my_vl, idx & 31);
            if (ok[u]) {
#pragma unroll
              for (int k = 0; k < VEC; ++k) {
                if constexpr (TSMASK & 1) a_sym[k] = __fadd_rn(a_sym[k], __fmul_rn(xs[u][k], vs));
                if constexpr (TSMASK & 2) a_lin[k] = __fadd_rn(a_lin[k], LINW ? __fmul_rn(xl[u][k], vl) : xl[u][k]);
                if constexpr (TSMASK & 4) a_sq[k] = __fadd_rn(a_sq[k], LINW ? __fmul_rn(xq[u][k], vl) : xq[u][k]);
              }
            }
          }
        }
      }
      for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          if constexpr (TSMASK & 1) a_sym[k] += __shfl_xor_sync(kFull, a_sym[k], off);
          if constexpr (TSMASK & 2) a_lin[k] += __shfl_xor_sync(kFull, a_lin[k], off);
          if constexpr (TSMASK & 4) a_sq[k] += __shfl_xor_sync(kFull, a_sq[k], off);
        }
      }
    } else if (active) {
      const int c0 = p.long_chunk_ptr[long_idx], c1 = p.long_chunk_ptr[long_idx + 1];
      for (int c = c0; c < c1; ++c) {
        const float* q = p.partials + static_cast<int64_t>(c) * row_stride + foff;
        float t[VEC];
        if constexpr (TSMASK & 1) { ld_plain<VEC>(t, q + p.ts_sym * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_sym[k] += t[k]; }
        if constexpr (TSMASK & 2) { ld_plain<VEC>(t, q + p.ts_lin * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_lin[k] += t[k]; }
        if constexpr (TSMASK & 4) { ld_plain<VEC>(t, q + p.ts_sq * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_sq[k] += t[k]; }
      }
    }

    const bool writer = active && lane < G;
    if (!writer) continue;
    if (chunk_id >= 0) {
      float* q = p.partials + static_cast<int64_t>(chunk_id) * row_stride + foff;
      if constexpr (TSMASK & 1) st_row<VEC>(q + p.ts_sym * p.BD, a_sym);
      if constexpr (TSMASK & 2) st_row<VEC>(q + p.ts_lin * p.BD, a_lin);
      if constexpr (TSMASK & 4) st_row<VEC>(q + p.ts_sq * p.BD, a_sq);
      continue;
    }
    float* dst = p.d_bases + static_cast<int64_t>(colj) * p.BD + foff;
    float r[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) r[k] = 0.f;
    if (p.routed) ld_plain<VEC>(r, dst);
    if constexpr (TSMASK & 4) {
      float xj[VEC];
      ld_row<VEC>(xj, p.bases + static_cast<int64_t>(colj) * p.BD + foff);
#pragma unroll
      for (int k = 0; k < VEC; ++k) r[k] += 2.f * xj[k] * a_sq[k];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if constexpr (TSMASK & 1) r[k] += a_sym[k];
      if constexpr (TSMASK & 2) r[k] += a_lin[k];
    }
    st_row<VEC>(dst, r);
  }
}

template <int TSMASK, int VEC, bool LINW>
static int launch_scatter_one(const ScatterParams& p, cudaStream_t st) {
  const int64_t tasks = p.mode == 0 ? static_cast<int64_t>(p.n_chunks) + p.n_cols : p.n_long;
  if (tasks <= 0) return EGC_OK;
  {
    LaunchScope egc_ls_(p.mode ? "k_scatter_bwd_merge" : "k_scatter_bwd", st);
    k_scatter_bwd<TSMASK, VEC, LINW><<<ceil_div(tasks, kAggWarps), kAggThreads, 0, st>>>(p);
  }
  EGC_LAUNCH_CHECK("k_scatter_bwd");
  return EGC_OK;
}

template <int VEC, bool LINW>
static int launch_scatter_mask(const ScatterParams& p, int tsmask, cudaStream_t st) {
  switch (tsmask) {
    case 1: return launch_scatter_one<1, VEC, LINW>(p, st);
    case 2: return launch_scatter_one<2, VEC, LINW>(p, st);
    case 3: return launch_scatter_one<3, VEC, LINW>(p, st);
    case 4: return launch_scatter_one<4, VEC, LINW>(p, st);
    case 5: return launch_scatter_one<5, VEC, LINW>(p, st);
    case 6: return launch_scatter_one<6, VEC, LINW>(p, st);
    case 7: return launch_scatter_one<7, VEC, LINW>(p, st);
  }
  set_error("scatter_bwd: bad stream mask %d", tsmask);
  return EGC_ERR_UNSUPPORTED;
}

static int launch_scatter(const ScatterParams& p, int tsmask, bool vec4, bool linw, cudaStream_t st) {
  if (vec4) return linw ? launch_scatter_mask<4, true>(p, tsmask, st) : launch_scatter_mask<4, false>(p, tsmask, st);
  return linw ? launch_scatter_mask<1, true>(p, tsmask, st) : launch_scatter_mask<1, false>(p, tsmask, st);
}

// which target-side streams does this aggregator list need?  bit0 sym, bit1 lin, bit2 sq
static int stream_mask_of(const egc_layer_desc& d, bool& has_route) {
  int m = 0;
  has_route = false;
  for (int a = 0; a < d.n_aggr; ++a) {
    switch (d.aggr[a]) {
      case EGC_AGGR_SUM: case EGC_AGGR_MEAN: m |= 2; break;
      case EGC_AGGR_SYMNORM: m |= 1; break;
      case EGC_AGGR_VAR: case EGC_AGGR_STD: m |= 2 | 4; break;
      case EGC_AGGR_MAX: case EGC_AGGR_MIN: has_route = true; break;
    }
  }
  return m;
}

struct BwdLayout {
  size_t ts_bytes, csr_part_bytes, csc_part_bytes, colsum_bytes, total;
  int n_ts, ts_sym, ts_lin, ts_sq, tsmask;
  bool has_route;
};

static BwdLayout bwd_layout(const egc_layer_desc& d, const egc_row_plan* csr_plan, const egc_row_plan* csc_plan) {
  BwdLayout L{};
  L.tsmask = stream_mask_of(d, L.has_route);
  int s = 0;
  L.ts_sym = (L.tsmask & 1) ? s++ : -1;
  L.ts_lin = (L.tsmask & 2) ? s++ : -1;
  L.ts_sq = (L.tsmask & 4) ? s++ : -1;
  L.n_ts = s;
  const size_t bd = static_cast<size_t>(d.bases) * d.dim;
  L.ts_bytes = align_up(static_cast<size_t>(d.n_dst) * std::max(L.n_ts, 1) * bd * 4, 256);
  const int mask = prim_mask_of(d);
  L.csr_part_bytes = align_up(static_cast<size_t>(csr_plan ? csr_plan->n_chunks : 0) * n_slots_of_mask(mask) * bd * 4, 256);
  L.csc_part_bytes = align_up(static_cast<size_t>(csc_plan ? csc_plan->n_chunks : 0) * std::max(L.n_ts, 1) * bd * 4, 256);
  L.colsum_bytes = align_up(colsum_workspace_bytes(d.n_dst, d.heads * d.dim), 256);
  L.total = L.ts_bytes + L.csr_part_bytes + L.csc_part_bytes + L.colsum_bytes + 256;
  return L;
}

}  // namespace egc

using namespace egc;

extern "C" {

size_t egc_aggregate_fwd_workspace_bytes(const egc_layer_desc* desc, const egc_row_plan* plan) {
  if (desc == nullptr) return 0;
  const int mask = prim_mask_of(*desc);
  if (mask <= 0) return 0;
  const size_t bd = static_cast<size_t>(desc->bases) * desc->dim;
  return align_up(static_cast<size_t>(plan ? plan->n_chunks : 0) * n_slots_of_mask(mask) * bd * 4, 256) + 256;
}

int egc_aggregate_fwd(const egc_layer_desc* desc, const int32_t* rowptr, const int32_t* col, const float* val_sym,
                      const float* val_lin, const egc_row_plan* plan, const float* bases, const float* weightings,
                      const float* bias, float* out, float* agg_out, int32_t* arg_out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (int rc = validate_desc(desc, "egc_aggregate_fwd")) return rc;
  if (int rc = validate_plan(plan, "egc_aggregate_fwd")) return rc;
  EGC_REQUIRE(rowptr && col && bases, "egc_aggregate_fwd: null graph / bases pointer");
  EGC_REQUIRE(out == nullptr || weightings != nullptr, "egc_aggregate_fwd: weightings required to produce out");
  EGC_REQUIRE(out || agg_out || arg_out, "egc_aggregate_fwd: no output requested");
  const int mask = prim_mask_of(*desc);
  EGC_REQUIRE(!(mask & P_SYM) || val_sym != nullptr, "egc_aggregate_fwd: symnorm requested without val_sym");
  EGC_REQUIRE(!((mask & P_SYM) && val_lin), "egc_aggregate_fwd: val_lin cannot be combined with symnorm (ref :253-254)");
  EGC_REQUIRE(workspace_bytes >= egc_aggregate_fwd_workspace_bytes(desc, plan) && (workspace || !(plan && plan->n_chunks)),
              "egc_aggregate_fwd: workspace too small");
  AggParams p{};
  const int bd = desc->bases * desc->dim;
  const bool vec4 = (bd % 4 == 0) && aligned16(bases) && aligned16(out) && aligned16(agg_out) && aligned16(workspace);
  const int smem = fill_agg_params(p, *desc, vec4, false);
  EGC_REQUIRE(smem <= 200 * 1024, "egc_aggregate_fwd: layer too wide for the shared-memory staging (%d bytes)", smem);
  p.rowptr = rowptr; p.col = col; p.val_sym = val_sym; p.val_lin = val_lin; p.n_rows = desc->n_dst;
  set_plan(p, plan);
  p.partials = static_cast<float*>(workspace);
  p.n_slots = n_slots_of_mask(mask);
  p.bases = bases; p.weightings = weightings ? weightings : bases; p.bias = bias;
  p.out = out; p.agg_out = agg_out; p.arg_out = arg_out;
  cudaStream_t st = as_stream(stream);
  auto launch = vec4 ? launch_aggregate_fwd_v4 : launch_aggregate_fwd_v1;
  p.mode = 0;
  if (int rc = launch(p, mask, val_lin != nullptr, smem, st)) return rc;
  if (p.n_long > 0) {
    p.mode = 1;
    if (int rc = launch(p, mask, val_lin != nullptr, smem, st)) return rc;
  }
  return EGC_OK;
}

size_t egc_aggregate_bwd_workspace_bytes(const egc_layer_desc* desc, int32_t nnz, const egc_row_plan* csr_plan,
                                         const egc_row_plan* csc_plan, int32_t flags) {
  (void)nnz; (void)flags;
  if (desc == nullptr || prim_mask_of(*desc) <= 0) return 0;
  return bwd_layout(*desc, csr_plan, csc_plan).total;
}

int egc_aggregate_bwd(const egc_layer_desc* desc, const int32_t* rowptr, const int32_t* col, const float* val_sym,
                      const float* val_lin, const egc_row_plan* csr_plan, const int32_t* colptr,
                      const int32_t* rowidx, const int32_t* csr2csc, const float* csc_val_sym,
                      const float* csc_val_lin, const egc_row_plan* csc_plan, const float* bases,
                      const float* weightings, const float* grad_out, float* d_weightings, float* d_bases,
                      float* d_bias, int32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
  (void)csr2csc;
  if (int rc = validate_desc(desc, "egc_aggregate_bwd")) return rc;
  if (int rc = validate_plan(csr_plan, "egc_aggregate_bwd")) return rc;
  if (int rc = validate_plan(csc_plan, "egc_aggregate_bwd")) return rc;
  EGC_REQUIRE(rowptr && col && colptr && rowidx && bases && weightings && grad_out && d_weightings && d_bases && workspace,
              "egc_aggregate_bwd: null pointer");
  if (flags & EGC_BWD_DETERMINISTIC) {
    bool has_route = false;
    stream_mask_of(*desc, has_route);
    if (has_route) {
      set_error("egc_aggregate_bwd: EGC_BWD_DETERMINISTIC routing of min/max gradients is not built yet");
      return EGC_ERR_UNSUPPORTED;
    }
  }
  const int mask = prim_mask_of(*desc);
  EGC_REQUIRE(!(mask & P_SYM) || (val_sym && csc_val_sym), "egc_aggregate_bwd: symnorm requested without val_sym / csc_val_sym");
  EGC_REQUIRE((val_lin == nullptr) == (csc_val_lin == nullptr), "egc_aggregate_bwd: val_lin and csc_val_lin must come together");
  EGC_REQUIRE(!((mask & P_SYM) && val_lin), "egc_aggregate_bwd: val_lin cannot be combined with symnorm");
  const BwdLayout L = bwd_layout(*desc, csr_plan, csc_plan);
  EGC_REQUIRE(workspace_bytes >= L.total, "egc_aggregate_bwd: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  float* tstreams = reinterpret_cast<float*>(ws);
  float* csr_part = reinterpret_cast<float*>(ws + L.ts_bytes);
  float* csc_part = reinterpret_cast<float*>(ws + L.ts_bytes + L.csr_part_bytes);
  void* colsum_ws = ws + L.ts_bytes + L.csr_part_bytes + L.csc_part_bytes;

  const int bd = desc->bases * desc->dim;
  const bool vec4 = (bd % 4 == 0) && aligned16(bases) && aligned16(d_bases) && aligned16(workspace) && aligned16(grad_out);

  if (L.has_route || L.tsmask == 0)
    EGC_CUDA(cudaMemsetAsync(d_bases, 0, static_cast<size_t>(desc->n_src) * bd * sizeof(float), st));

  // pass 1: per target row (CSR)
  AggParams p{};
  const int smem = fill_agg_params(p, *desc, vec4, true);
  EGC_REQUIRE(smem <= 200 * 1024, "egc_aggregate_bwd: layer too wide for the shared-memory staging (%d bytes)", smem);
  p.rowptr = rowptr; p.col = col; p.val_sym = val_sym; p.val_lin = val_lin; p.n_rows = desc->n_dst;
  set_plan(p, csr_plan);
  p.partials = csr_part;
  p.n_slots = n_slots_of_mask(mask);
  p.bases = bases; p.weightings = weightings; p.grad_out = grad_out;
  p.d_weightings = d_weightings; p.tstreams = tstreams; p.d_bases = d_bases;
  p.n_ts = L.n_ts; p.ts_sym = L.ts_sym; p.ts_lin = L.ts_lin; p.ts_sq = L.ts_sq;
  auto launch = vec4 ? launch_aggregate_bwd_v4 : launch_aggregate_bwd_v1;
  p.mode = 0;
  if (int rc = launch(p, mask, val_lin != nullptr, smem, st)) return rc;
  if (p.n_long > 0) {
    p.mode = 1;
    if (int rc = launch(p, mask, val_lin != nullptr, smem, st)) return rc;
  }

  // pass 2: per source column (CSC), atomic-free
  if (L.tsmask != 0) {
    ScatterParams s{};
    s.colptr = colptr; s.rowidx = rowidx; s.val_sym = csc_val_sym; s.val_lin = csc_val_lin; s.n_cols = desc->n_src;
    s.n_long = csc_plan ? csc_plan->n_long : 0;
    s.n_chunks = csc_plan ? csc_plan->n_chunks : 0;
    s.long_rows = csc_plan ? csc_plan->long_rows : nullptr;
    s.long_chunk_ptr = csc_plan ? csc_plan->long_chunk_ptr : nullptr;
    s.chunk_row = csc_plan ? csc_plan->chunk_row : nullptr;
    s.chunk_begin = csc_plan ? csc_plan->chunk_begin : nullptr;
    s.partials = csc_part;
    s.tstreams = tstreams; s.bases = bases; s.d_bases = d_bases;
    s.n_ts = L.n_ts; s.ts_sym = L.ts_sym; s.ts_lin = L.ts_lin; s.ts_sq = L.ts_sq;
    s.BD = bd; s.nvec = p.nvec; s.G = p.G; s.n_pass = p.n_pass;
    s.routed = L.has_route ? 1 : 0;
    s.mode = 0;
    if (int rc = launch_scatter(s, L.tsmask, vec4, val_lin != nullptr, st)) return rc;
    if (s.n_long > 0) {
      s.mode = 1;
      if (int rc = launch_scatter(s, L.tsmask, vec4, val_lin != nullptr, st)) return rc;
    }
  }

  if (d_bias != nullptr) {
    if (int rc = colsum_f32(grad_out, desc->n_dst, desc->heads * desc->dim, d_bias, colsum_ws, L.colsum_bytes, st)) return rc;
  }
  return EGC_OK;
}

}  // extern "C"
