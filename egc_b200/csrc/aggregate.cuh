// Fused multi-aggregator CSR SpMM + combination: shared device machinery (sm_100a).
//
// One warp owns one task: a whole CSR row (<= EGC_CHUNK_EDGES nnz), one chunk of a long row, or the
// merge of a long row's chunk partials.  A basis row (B*D floats) is spread over the lanes of a
// "group" of G lanes in VEC-wide pieces (VEC = 4 -> 128-bit gathers); when the row is narrower than
// 32*VEC floats the warp's 32/G groups walk different neighbours in parallel and are merged with
// xor-shuffles at the end.  All requested aggregators are produced from ONE gather of each
// neighbour row: primitives sum / symnorm-sum / sum of squares / max(+arg) / min(+arg).
#pragma once

#include <float.h>

#include "common.cuh"

namespace egc {

enum Prim : int { P_SUM = 1, P_SYM = 2, P_SQ = 4, P_MAX = 8, P_MIN = 16 };

constexpr int kAggWarps = 8;
constexpr int kAggThreads = kAggWarps * 32;
constexpr int kGatherUnroll = 8;      // independent 128-bit gathers in flight per lane
constexpr float kStdEps = 1e-5f;      // ref optimized_layers.py:244,273

struct AggParams {
  // target-major CSR (or the CSC when used by the scatter pass)
  const int32_t* rowptr;
  const int32_t* col;
  const float* val_sym;
  const float* val_lin;
  int n_rows;
  // long-row plan + scratch for chunk partials [n_chunks][n_slots][BD]
  int n_long, n_chunks;
  const int32_t* long_rows;
  const int32_t* long_chunk_ptr;
  const int32_t* chunk_row;
  const int32_t* chunk_begin;
  float* partials;
  int n_slots;
  // features
  const float* bases;        // [n_src, BD]
  const float* weightings;   // [n_rows, HAB]
  const float* bias;         // [HD] or null
  float* out;                // [n_rows, HD]
  float* agg_out;            // [n_rows, A, BD] or null
  int32_t* arg_out;          // [n_rows, A, BD] or null
  // backward (pass 1)
  const float* grad_out;     // [n_rows, HD]
  float* d_weightings;       // [n_rows, HAB]
  float* tstreams;           // [n_rows, n_ts, BD]
  float* d_bases;            // [n_src, BD], pre-zeroed when min/max gradients are routed atomically
  int n_ts, ts_sym, ts_lin, ts_sq;   // stream slots (-1: absent)
  // shape
  int H, B, D, A, BD, HD, AB, HAB;
  int aggr[EGC_MAX_AGGR];
  int sigmoid;
  // lane geometry
  int nvec;      // VEC-wide pieces per basis row
  int G;         // lanes per group (power of two, <= 32)
  int n_pass;    // passes of 32 pieces when nvec > 32
  // per-warp shared memory layout (float offsets)
  int sm_agg, sm_w, sm_g, sm_mean, sm_var, sm_amx, sm_amn, sm_per_warp;
  int mode;      // 0: chunk tasks then row tasks, 1: merge tasks (one per long row)
};

// ---------------------------------------------------------------------------------------------
// vector load / store of VEC consecutive floats
// ---------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ld_row(float (&v)[VEC], const float* __restrict__ p) {
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = __ldg(p + k);
  }
}

template <int VEC>
__device__ __forceinline__ void ld_plain(float (&v)[VEC], const float* p) {
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = p[k];
  }
}

template <int VEC>
__device__ __forceinline__ void st_row(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) p[k] = v[k];
  }
}

__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gmem_src) {
  unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// accumulator of the aggregation primitives for VEC features
// ---------------------------------------------------------------------------------------------
template <int MASK, int VEC, bool LINW>
struct Acc {
  float sum[VEC], sym[VEC], sq[VEC], mx[VEC], mn[VEC];
  int amx[VEC], amn[VEC];

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      sum[k] = 0.f; sym[k] = 0.f; sq[k] = 0.f;
      mx[k] = -INFINITY; mn[k] = INFINITY;
      amx[k] = -1; amn[k] = -1;
    }
  }

  // one neighbour; e = nnz position (for first-wins arg tracking).  Products and sums are rounded
  // separately (no FMA contraction) so a sequential walk reproduces the reference's fp32 results.
  __device__ __forceinline__ void add(const float (&x)[VEC], float vs, float vl, int e) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float xl = LINW ? __fmul_rn(x[k], vl) : x[k];
      if constexpr (MASK & P_SUM) sum[k] = __fadd_rn(sum[k], xl);
      if constexpr (MASK & P_SYM) sym[k] = __fadd_rn(sym[k], __fmul_rn(x[k], vs));
      if constexpr (MASK & P_SQ) {
        float s = __fmul_rn(x[k], x[k]);
        if (LINW) s = __fmul_rn(s, vl);
        sq[k] = __fadd_rn(sq[k], s);
      }
      if constexpr (MASK & P_MAX) { if (xl > mx[k]) { mx[k] = xl; amx[k] = e; } }
      if constexpr (MASK & P_MIN) { if (xl < mn[k]) { mn[k] = xl; amn[k] = e; } }
    }
  }

  // ties between partial results: the smaller nnz position wins (-1 = empty compares as +inf)
  __device__ __forceinline__ void merge_max(int k, float o, int oa) {
    if (o > mx[k] || (o == mx[k] && static_cast<unsigned>(oa) < static_cast<unsigned>(amx[k]))) { mx[k] = o; amx[k] = oa; }
  }
  __device__ __forceinline__ void merge_min(int k, float o, int oa) {
    if (o < mn[k] || (o == mn[k] && static_cast<unsigned>(oa) < static_cast<unsigned>(amn[k]))) { mn[k] = o; amn[k] = oa; }
  }

  __device__ __forceinline__ void merge_xor(int off) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if constexpr (MASK & P_SUM) sum[k] = __fadd_rn(sum[k], __shfl_xor_sync(kFull, sum[k], off));
      if constexpr (MASK & P_SYM) sym[k] = __fadd_rn(sym[k], __shfl_xor_sync(kFull, sym[k], off));
      if constexpr (MASK & P_SQ) sq[k] = __fadd_rn(sq[k], __shfl_xor_sync(kFull, sq[k], off));
      if constexpr (MASK & P_MAX) {
        float o = __shfl_xor_sync(kFull, mx[k], off);
        int oa = __shfl_xor_sync(kFull, amx[k], off);
        merge_max(k, o, oa);
      }
      if constexpr (MASK & P_MIN) {
        float o = __shfl_xor_sync(kFull, mn[k], off);
        int oa = __shfl_xor_sync(kFull, amn[k], off);
        merge_min(k, o, oa);
      }
    }
  }

  static constexpr int n_slots() {
    return ((MASK & P_SUM) ? 1 : 0) + ((MASK & P_SYM) ? 1 : 0) + ((MASK & P_SQ) ? 1 : 0) +
           ((MASK & P_MAX) ? 2 : 0) + ((MASK & P_MIN) ? 2 : 0);
  }

  // partial <-> scratch; `p` points at slot 0 of this lane's features, slots are BD floats apart
  __device__ __forceinline__ void store(float* p, int BD) const {
    int s = 0;
    if constexpr (MASK & P_SUM) { st_row<VEC>(p + s * BD, sum); ++s; }
    if constexpr (MASK & P_SYM) { st_row<VEC>(p + s * BD, sym); ++s; }
    if constexpr (MASK & P_SQ) { st_row<VEC>(p + s * BD, sq); ++s; }
    if constexpr (MASK & P_MAX) {
      st_row<VEC>(p + s * BD, mx); ++s;
      float t[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) t[k] = __int_as_float(amx[k]);
      st_row<VEC>(p + s * BD, t); ++s;
    }
    if constexpr (MASK & P_MIN) {
      st_row<VEC>(p + s * BD, mn); ++s;
      float t[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) t[k] = __int_as_float(amn[k]);
      st_row<VEC>(p + s * BD, t); ++s;
    }
  }

  __device__ __forceinline__ void merge_from(const float* p, int BD) {
    int s = 0;
    float t[VEC], u[VEC];
    if constexpr (MASK & P_SUM) {
      ld_plain<VEC>(t, p + s * BD); ++s;
#pragma unroll
      for (int k = 0; k < VEC; ++k) sum[k] = __fadd_rn(sum[k], t[k]);
    }
    if constexpr (MASK & P_SYM) {
      ld_plain<VEC>(t, p + s * BD); ++s;
#pragma unroll
      for (int k = 0; k < VEC; ++k) sym[k] = __fadd_rn(sym[k], t[k]);
    }
    if constexpr (MASK & P_SQ) {
      ld_plain<VEC>(t, p + s * BD); ++s;
#pragma unroll
      for (int k = 0; k < VEC; ++k) sq[k] = __fadd_rn(sq[k], t[k]);
    }
    if constexpr (MASK & P_MAX) {
      ld_plain<VEC>(t, p + s * BD); ++s;
      ld_plain<VEC>(u, p + s * BD); ++s;
#pragma unroll
      for (int k = 0; k < VEC; ++k) merge_max(k, t[k], __float_as_int(u[k]));
    }
    if constexpr (MASK & P_MIN) {
      ld_plain<VEC>(t, p + s * BD); ++s;
      ld_plain<VEC>(u, p + s * BD); ++s;
#pragma unroll
      for (int k = 0; k < VEC; ++k) merge_min(k, t[k], __float_as_int(u[k]));
    }
  }
};

// ---------------------------------------------------------------------------------------------
// walk nnz [begin, end) of one row; lanes of group g = lane / G take neighbours g, g + NG, ...
// ---------------------------------------------------------------------------------------------
template <int MASK, int VEC, bool LINW>
__device__ __forceinline__ void accumulate_range(Acc<MASK, VEC, LINW>& acc, const AggParams& p, int begin, int end,
                                                 int lane, int foff, bool active) {
  const int G = p.G, NG = 32 / G, g = lane / G;
  const float* __restrict__ src = p.bases + foff;
  for (int e0 = begin; e0 < end; e0 += 32) {
    const int n_here = min(32, end - e0);
    const bool have = lane < n_here;
    const int my_col = have ? __ldg(p.col + e0 + lane) : 0;
    float my_vs = 0.f, my_vl = 0.f;
    if constexpr (MASK & P_SYM) my_vs = have ? __ldg(p.val_sym + e0 + lane) : 0.f;
    if constexpr (LINW) my_vl = have ? __ldg(p.val_lin + e0 + lane) : 0.f;
    const int steps = (n_here + NG - 1) / NG;
    for (int s = 0; s < steps; s += kGatherUnroll) {
      float x[kGatherUnroll][VEC];
      bool ok[kGatherUnroll];
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u) {
        const int idx = (s + u) * NG + g;
        const int j = __shfl_sync(kFull, my_col, idx & 31);
        ok[u] = active && (s + u) < steps && idx < n_here;
        if (ok[u]) ld_row<VEC>(x[u], src + static_cast<int64_t>(j) * p.BD);
      }
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u) {
        const int idx = (s + u) * NG + g;
        float vs = 0.f, vl = 0.f;
        if constexpr (MASK & P_SYM) vs = __shfl_sync(kFull, my_vs, idx & 31);
        if constexpr (LINW) vl = __shfl_sync(kFull, my_vl, idx & 31);
        if (ok[u]) acc.add(x[u], vs, vl, e0 + idx);
      }
    }
  }
  for (int off = G; off < 32; off <<= 1) acc.merge_xor(off);
}

// value of aggregator `code` for feature k of this lane, from the primitives
template <int MASK, int VEC, bool LINW>
__device__ __forceinline__ float finalize_one(const Acc<MASK, VEC, LINW>& acc, int code, int k, float cntf,
                                              float& mean_out, float& var_out) {
  switch (code) {
    case EGC_AGGR_SUM:
      if constexpr (MASK & P_SUM) return acc.sum[k];
      break;
    case EGC_AGGR_MEAN:
      if constexpr (MASK & P_SUM) return __fdiv_rn(acc.sum[k], cntf);
      break;
    case EGC_AGGR_SYMNORM:
      if constexpr (MASK & P_SYM) return acc.sym[k];
      break;
    case EGC_AGGR_MAX:
      if constexpr (MASK & P_MAX) return acc.amx[k] >= 0 ? acc.mx[k] : 0.f;
      break;
    case EGC_AGGR_MIN:
      if constexpr (MASK & P_MIN) return acc.amn[k] >= 0 ? acc.mn[k] : 0.f;
      break;
    case EGC_AGGR_VAR:
    case EGC_AGGR_STD:
      if constexpr ((MASK & P_SQ) && (MASK & P_SUM)) {
        const float mean = __fdiv_rn(acc.sum[k], cntf);
        const float msq = __fdiv_rn(acc.sq[k], cntf);
        const float var = __fsub_rn(msq, __fmul_rn(mean, mean));   // mean_sq - mean*mean, ref :242 / :271
        mean_out = mean;
        var_out = var;
        if (code == EGC_AGGR_VAR) return var;
        return sqrtf(__fadd_rn(fmaxf(var, 0.f), kStdEps));          // sqrt(relu(var) + 1e-5), ref :244 / :273
      }
      break;
  }
  return 0.f;
}

inline int prim_mask_of(const egc_layer_desc& d) {
  int m = 0;
  for (int a = 0; a < d.n_aggr; ++a) {
    switch (d.aggr[a]) {
      case EGC_AGGR_SUM: case EGC_AGGR_MEAN: m |= P_SUM; break;
      case EGC_AGGR_SYMNORM: m |= P_SYM; break;
      case EGC_AGGR_MAX: m |= P_MAX; break;
      case EGC_AGGR_MIN: m |= P_MIN; break;
      case EGC_AGGR_VAR: case EGC_AGGR_STD: m |= P_SUM | P_SQ; break;
      default: return -1;
    }
  }
  return m;
}

inline int n_slots_of_mask(int m) {
  return ((m & P_SUM) ? 1 : 0) + ((m & P_SYM) ? 1 : 0) + ((m & P_SQ) ? 1 : 0) + ((m & P_MAX) ? 2 : 0) +
         ((m & P_MIN) ? 2 : 0);
}

// host: fill shape / geometry / smem layout fields; returns dynamic smem bytes per CTA
int fill_agg_params(AggParams& p, const egc_layer_desc& d, bool vec4, bool bwd);

// launchers (one translation unit per VEC/BWD combination to keep compile times parallel)
int launch_aggregate_fwd_v4(const AggParams& p, int mask, bool linw, int smem_bytes, cudaStream_t st);
int launch_aggregate_fwd_v1(const AggParams& p, int mask, bool linw, int smem_bytes, cudaStream_t st);
int launch_aggregate_bwd_v4(const AggParams& p, int mask, bool linw, int smem_bytes, cudaStream_t st);
int launch_aggregate_bwd_v1(const AggParams& p, int mask, bool linw, int smem_bytes, cudaStream_t st);

}  // namespace egc
