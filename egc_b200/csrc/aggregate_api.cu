// extern "C" entry points of the fused aggregation forward / backward, plus the two backward
// kernels: the per-target streaming pass (gradient of the combination, target-side streams,
// min/max routing) and the per-source CSC gather pass (atomic-free d_bases).
#include <stdlib.h>

#include <algorithm>

#include "aggregate_fast.cuh"
#include "colsum.cuh"

namespace egc {

int fill_agg_params(AggParams& p, const egc_layer_desc& d, bool vec4) {
  p.H = d.heads; p.B = d.bases; p.D = d.dim; p.A = d.n_aggr;
  p.BD = d.bases * d.dim; p.HD = d.heads * d.dim; p.AB = d.n_aggr * d.bases; p.HAB = d.heads * p.AB;
  int n_arg = 0;
  for (int a = 0; a < EGC_MAX_AGGR; ++a) {
    p.aggr[a] = a < d.n_aggr ? d.aggr[a] : -1;
    p.arg_slot[a] = (a < d.n_aggr && (d.aggr[a] == EGC_AGGR_MAX || d.aggr[a] == EGC_AGGR_MIN)) ? n_arg++ : -1;
  }
  p.n_arg = n_arg;
  p.n_saved = n_saved_slots(d);
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

// =============================================================================================
// backward pass 1: per target node, streaming (no graph traversal)
//   d_w[h,ab]   = sum_d g[h*D+d] * agg[ab*D+d]                       (x sigmoid' when requested)
//   d_agg[a][p] = sum_h w[h*AB + a*B + b(p)] * g[h*D + d(p)]
//   -> target-side streams t_sym / t_lin / t_sq (read by pass 2) + single-winner routing of min/max
// `saved` comes from the forward: per aggregator slot its value (std slots carry the closed relu gate
// in the sign bit), plus one extra slot with the mean when var/std is present; `saved_arg` holds the
// winning nnz position of every min/max slot.
// =============================================================================================
struct CombineBwdParams {
  const int32_t* rowptr;
  const int32_t* col;
  const float* val_lin;
  int n_rows;
  const float* weightings;
  const float* grad_out;
  const float* saved;
  const int32_t* saved_arg;
  float* d_weightings;
  float* tstreams;
  float* d_bases;
  int n_saved, n_arg, n_ts, ts_sym, ts_lin, ts_sq;
  int H, B, D, A, BD, HD, AB, HAB;
  int aggr[EGC_MAX_AGGR];
  int arg_slot[EGC_MAX_AGGR];
  int sigmoid;
  int sm_w, sm_g, sm_saved, sm_arg, sm_per_warp;
  int vec16;
  int64_t ts_row_stride, ts_stream_stride;   // floats between the streams of consecutive rows / between streams of a row
  int ts_slab_w;                             // features per slab (== BD: one slab, the plain interleaved / stream-major layouts)
  int64_t ts_slab_stride;                    // floats between consecutive feature slabs ([slab][row][stream][ts_slab_w])
  float* colsum_part;                        // [grid][HD + HAB] or null
  int skip_route;                            // diagnostics: drop the min/max routing
  float* t_route;                            // [n_rows][n_arg][BD] gradients of the min/max slots, routed by k_route_minmax
};

__device__ __forceinline__ void stage_row(float* dst, const float* src, int n, int lane, bool vec16) {
  if (vec16) {
    for (int t = lane * 4; t < n; t += 128) cp_async_16(dst + t, src + t);
  } else {
    for (int t = lane; t < n; t += 32) cp_async_4(dst + t, src + t);
  }
}

constexpr int kCbColIt = 4;     // per-lane column-sum accumulators: HD <= 32 * EV * kCbColIt, HAB <= 32 * kCbColIt

__device__ __forceinline__ void stg_stream_f4(float* p, const float (&v)[4]) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}

// Persistent warps (one target row per task).  Everything that does not depend on the row - the (head, slot)
// decomposition of a weight index, the (basis, offset) of a feature, the aggregator of a slot - is computed
// once per CTA into shared-memory tables.  The column sums of grad_out (d_bias) and of d_weightings (the
// gradient of the comb-weight bias) ride along in registers and leave as one partial row per CTA
// (deterministic: reduced in CTA order by k_colsum_partials).
template <class Cfg, int EV, bool LINW>
__global__ void __launch_bounds__(kAggThreads, 4) k_combine_bwd(const __grid_constant__ CombineBwdParams p) {
  using GB = GetB<Cfg, CombineBwdParams>;
  extern __shared__ __align__(16) float smem_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = GB::D(p), BD = GB::BD(p), HD = GB::HD(p), AB = GB::AB(p), HAB = GB::HAB(p);
  // CTA-wide tables (ints) live behind the per-warp staging areas
  int* tab_goff = reinterpret_cast<int*>(smem_all + kAggWarps * GB::sm_per_warp(p));   // [HAB] h * D
  int* tab_aoff = tab_goff + HAB;                                                  // [HAB] ab * D
  int* tab_std = tab_aoff + HAB;                                                   // [HAB] slot belongs to a std aggregator
  if constexpr (!Cfg::kStatic) {
    for (int t = threadIdx.x; t < HAB; t += kAggThreads) {
      const int h = t / AB, ab = t - h * AB;
      tab_goff[t] = h * D;
      tab_aoff[t] = ab * D;
      tab_std[t] = p.aggr[ab / GB::B(p)] == EGC_AGGR_STD ? 1 : 0;
    }
    __syncthreads();
  }

  float* sm = smem_all + warp * GB::sm_per_warp(p);
  const bool v16 = p.vec16 != 0;
  const float* w = sm + GB::sm_w(p);
  const float* g = sm + GB::sm_g(p);
  const float* sv = sm + GB::sm_saved(p);
  const int nq = D >> 2;
  const int q0 = nq > 0 ? lane % nq : 0, dd0 = lane % D;
  const int warps_total = gridDim.x * kAggWarps;

  float gsum[kCbColIt][EV], wsum[kCbColIt];
#pragma unroll
  for (int it = 0; it < kCbColIt; ++it) {
    wsum[it] = 0.f;
#pragma unroll
    for (int k = 0; k < EV; ++k) gsum[it][k] = 0.f;
  }

  for (int row = blockIdx.x * kAggWarps + warp; row < p.n_rows; row += warps_total) {
    stage_row(sm + GB::sm_w(p), p.weightings + static_cast<int64_t>(row) * HAB, HAB, lane, v16);
    stage_row(sm + GB::sm_g(p), p.grad_out + static_cast<int64_t>(row) * HD, HD, lane, v16);
    stage_row(sm + GB::sm_saved(p), p.saved + static_cast<int64_t>(row) * GB::n_saved(p) * BD, GB::n_saved(p) * BD, lane, v16);
    const float cntf = static_cast<float>(max(__ldg(p.rowptr + row + 1) - __ldg(p.rowptr + row), 1));
    const float inv_cnt = __frcp_rn(cntf);
    cp_async_wait_all();
    __syncwarp();

    // (0) column sums of grad_out
    if (p.colsum_part != nullptr) {
#pragma unroll
      for (int it = 0; it < kCbColIt; ++it) {
        const int c = lane * EV + 32 * EV * it;
        if (c < HD) {
          float t[EV];
          ld_plain<EV>(t, g + c);
#pragma unroll
          for (int k = 0; k < EV; ++k) gsum[it][k] += t[k];
        }
      }
    }

    // (1) gradient of the combination weights: HAB dot products of length D, skewed start per lane
#pragma unroll
    for (int it = 0; it < kCbColIt; ++it) {
      for (int t = lane + 32 * it; t < HAB; t += 32 * kCbColIt) {
        bool is_std;
        const float* gh;
        const float* aa;
        if constexpr (Cfg::kStatic) {
          const int h = t / AB, ab = t - h * AB;
          is_std = GB::aggr(p, ab / GB::B(p)) == EGC_AGGR_STD;
          gh = g + h * D;
          aa = sv + ab * D;
        } else {
          is_std = tab_std[t] != 0;
          gh = g + tab_goff[t];
          aa = sv + tab_aoff[t];
        }
        float dot = 0.f;
        if constexpr (EV == 4) {
          int q = q0;
          for (int i = 0; i < nq; ++i) {
            const float4 gv = *reinterpret_cast<const float4*>(gh + 4 * q);
            float4 av = *reinterpret_cast<const float4*>(aa + 4 * q);
            if (is_std) { av.x = fabsf(av.x); av.y = fabsf(av.y); av.z = fabsf(av.z); av.w = fabsf(av.w); }
            dot = fmaf(gv.x, av.x, dot); dot = fmaf(gv.y, av.y, dot); dot = fmaf(gv.z, av.z, dot); dot = fmaf(gv.w, av.w, dot);
            q = (q + 1 == nq) ? 0 : q + 1;
          }
        } else {
          int dd = dd0;
          for (int i = 0; i < D; ++i) {
            const float av = is_std ? fabsf(aa[dd]) : aa[dd];
            dot = fmaf(gh[dd], av, dot);
            dd = (dd + 1 == D) ? 0 : dd + 1;
          }
        }
        if (p.sigmoid) { const float s = w[t]; dot *= s * (1.f - s); }
        __stcs(p.d_weightings + static_cast<int64_t>(row) * HAB + t, dot);
        if (t < 32 * kCbColIt) wsum[it] += dot;
      }
    }

    // (2) gradient w.r.t. the aggregates -> target-side streams and min/max routing
    float* ts = p.tstreams + static_cast<int64_t>(row) * p.ts_row_stride;
    for (int p0 = lane * EV; p0 < BD; p0 += 32 * EV) {
      const int b = p0 / D, d = p0 - b * D;
      float t_sym[EV], t_lin[EV], t_sq[EV];
#pragma unroll
      for (int k = 0; k < EV; ++k) { t_sym[k] = 0.f; t_lin[k] = 0.f; t_sq[k] = 0.f; }
#pragma unroll
      for (int a = 0; a < GB::A(p); ++a) {
        float da[EV];
#pragma unroll
        for (int k = 0; k < EV; ++k) da[k] = 0.f;
        const float* wa = w + a * GB::B(p) + b;
#pragma unroll
        for (int h = 0; h < GB::H(p); ++h) {
          const float wv = wa[h * AB];
          float gv[EV];
          ld_plain<EV>(gv, g + h * D + d);
#pragma unroll
          for (int k = 0; k < EV; ++k) da[k] = fmaf(wv, gv[k], da[k]);
        }
        const int code = GB::aggr(p, a);
        if (code == EGC_AGGR_SUM) {
#pragma unroll
          for (int k = 0; k < EV; ++k) t_lin[k] += da[k];
        } else if (code == EGC_AGGR_MEAN) {
#pragma unroll
          for (int k = 0; k < EV; ++k) t_lin[k] = fmaf(da[k], inv_cnt, t_lin[k]);
        } else if (code == EGC_AGGR_SYMNORM) {
#pragma unroll
          for (int k = 0; k < EV; ++k) t_sym[k] += da[k];
        } else if (code == EGC_AGGR_MAX || code == EGC_AGGR_MIN) {
          // routed to the single winning source by k_route_minmax (feature-slab order keeps its atomics in the L2)
          float* tr = p.t_route + (static_cast<int64_t>(row) * GB::n_arg(p) + GB::arg_slot(p, a)) * BD + p0;
          if constexpr (EV == 4) stg_stream_f4(tr, da); else st_row<EV>(tr, da);
        } else {   // VAR / STD
          float sa[EV], mean[EV];
          ld_plain<EV>(sa, sv + a * BD + p0);
          ld_plain<EV>(mean, sv + GB::A(p) * BD + p0);
#pragma unroll
          for (int k = 0; k < EV; ++k) {
            float dv = da[k];
            if (code == EGC_AGGR_STD) dv = sa[k] > 0.f ? __fdividef(dv, 2.f * sa[k]) : 0.f;    // relu gate (sign bit), d sqrt
            const float q = dv * inv_cnt;
            t_sq[k] += q;
            t_lin[k] = fmaf(-2.f * mean[k], q, t_lin[k]);
          }
        }
      }
      // plain (L2 write-back) stores: pass 2 gathers these rows next, whatever part of them survives in the L2 is a hit.
      // Slab layout: feature p0 lives in slab p0 / W at offset p0 % W (an EV-wide piece never straddles slabs).
      const int slab = p0 / p.ts_slab_w;
      float* tsp = ts + static_cast<int64_t>(slab) * p.ts_slab_stride + (p0 - slab * p.ts_slab_w);
      if (GB::ts_sym(p) >= 0) st_row<EV>(tsp + static_cast<int64_t>(GB::ts_sym(p)) * p.ts_stream_stride, t_sym);
      if (GB::ts_lin(p) >= 0) st_row<EV>(tsp + static_cast<int64_t>(GB::ts_lin(p)) * p.ts_stream_stride, t_lin);
      if (GB::ts_sq(p) >= 0) st_row<EV>(tsp + static_cast<int64_t>(GB::ts_sq(p)) * p.ts_stream_stride, t_sq);
    }
    __syncwarp();     // the next row overwrites this warp's staging area
  }

  // ---- column-sum partials of this CTA: [HD] grad_out sums | [HAB] d_weightings sums
  if (p.colsum_part != nullptr) {
    __syncthreads();
    float* red = smem_all;                                   // reuse the staging areas: [kAggWarps][HD + HAB]
    const int width = HD + HAB;
#pragma unroll
    for (int it = 0; it < kCbColIt; ++it) {
      const int c = lane * EV + 32 * EV * it;
      if (c < HD) {
#pragma unroll
        for (int k = 0; k < EV; ++k) red[warp * width + c + k] = gsum[it][k];
      }
      const int t = lane + 32 * it;
      if (t < HAB) red[warp * width + HD + t] = wsum[it];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < width; c += kAggThreads) {
      float t = 0.f;
#pragma unroll
      for (int wv = 0; wv < kAggWarps; ++wv) t += red[wv * width + c];
      p.colsum_part[static_cast<int64_t>(blockIdx.x) * width + c] = t;
    }
  }
}

// out[c] = sum over CTAs of part[cta][c] in CTA order; columns [0, n1) -> out1, [n1, n1 + n2) -> out2
__global__ void k_colsum_partials(const float* __restrict__ part, int n_cta, int n1, int n2, float* __restrict__ out1,
                                  float* __restrict__ out2) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int width = n1 + n2;
  if (c >= width) return;
  float t = 0.f;
  for (int s = lane; s < n_cta; s += 32) t += part[static_cast<int64_t>(s) * width + c];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
  if (lane == 0) {
    if (c < n1) { if (out1 != nullptr) out1[c] = t; }
    else if (out2 != nullptr) out2[c - n1] = t;
  }
}


// =============================================================================================
// min/max gradient routing: d_bases[col[arg[i][s][p]]][p] += t_route[i][s][p]   (x val_lin[arg] when weighted)
// One warp task = 32 consecutive features of kRouteRows consecutive target rows.  The grid walks the
// (feature slab, slot) phases together, so at any moment it adds into one 32-float column slab of d_bases
// (n_src x 128 B - L2-resident) instead of missing to DRAM all over the [n_src, BD] matrix.
// Hub sources (the long columns of the CSC plan: a power-law hub sits in thousands of rows and wins a share of
// the features in each) would serialise tens of thousands of fp32 REDs on ONE 128-byte line per phase - the L2
// atomic unit retires about one lane per cycle per line - so every CTA privatises them: a shared-memory hash
// maps the hub ids to slots, their contributions go to shared-memory accumulators (ATOMS), and each CTA flushes
// one RED per (hub, feature) at the end of the phase.
// =============================================================================================
constexpr int kRouteRows = 8;
constexpr int kRouteThreads = 1024;
constexpr int kRouteMaxHubs = 1024;
constexpr int kRouteHashSize = 2 * kRouteMaxHubs;           // open addressing, load factor <= 0.5

struct RouteParams {
  const int32_t* saved_arg;    // [n_rows][n_arg][BD] winning nnz position (-1: none)
  const float* t_route;        // [n_rows][n_arg][BD]
  const int32_t* col;
  const float* val_lin;        // or null
  float* d_bases;              // [n_src][BD]
  const int32_t* hubs;         // [n_hubs] source ids privatised in shared memory (or null)
  int n_hubs;
  int n_rows, n_arg, BD, n_slabs;
};

__device__ __forceinline__ uint32_t route_hash(int j) { return (static_cast<uint32_t>(j) * 2654435761u) >> (32 - 11); }
static_assert(kRouteHashSize == 1 << 11, "route_hash yields 11 bits");

__global__ void __launch_bounds__(kRouteThreads, 1) k_route_minmax(const __grid_constant__ RouteParams p) {
  extern __shared__ __align__(16) float route_smem[];
  int* hash_key = reinterpret_cast<int*>(route_smem);                       // [kRouteHashSize] source id or -1
  int* hash_slot = hash_key + kRouteHashSize;                               // [kRouteHashSize]
  float* acc = reinterpret_cast<float*>(hash_slot + kRouteHashSize);        // [n_hubs][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_cta = blockDim.x >> 5;
  const int warps_total = gridDim.x * warps_cta;
  const int row_groups = (p.n_rows + kRouteRows - 1) / kRouteRows;
  const int64_t row_stride = static_cast<int64_t>(p.n_arg) * p.BD;
  const bool hubs_on = p.n_hubs > 0;
  if (hubs_on) {
    for (int t = threadIdx.x; t < kRouteHashSize; t += blockDim.x) hash_key[t] = -1;
    for (int t = threadIdx.x; t < p.n_hubs * 32; t += blockDim.x) acc[t] = 0.f;
    __syncthreads();
    for (int t = threadIdx.x; t < p.n_hubs; t += blockDim.x) {
      const int j = __ldg(p.hubs + t);
      uint32_t h = route_hash(j);
      while (atomicCAS(hash_key + h, -1, j) != -1) h = (h + 1) & (kRouteHashSize - 1);   // hub ids are distinct
      hash_slot[h] = t;
    }
    __syncthreads();
  }
  for (int phase = 0; phase < p.n_slabs * p.n_arg; ++phase) {
    const int slab = phase / p.n_arg, slot = phase - slab * p.n_arg;
    const int f = slab * 32 + lane;
    const bool f_ok = f < p.BD;
    const int64_t base = static_cast<int64_t>(slot) * p.BD + f;
    for (int rg = blockIdx.x * warps_cta + warp; rg < row_groups; rg += warps_total) {
      int arg[kRouteRows];
      float v[kRouteRows];
#pragma unroll
      for (int r = 0; r < kRouteRows; ++r) {
        const int row = rg * kRouteRows + r;
        arg[r] = -1;
        v[r] = 0.f;
        if (row < p.n_rows && f_ok) {
          arg[r] = __ldcs(p.saved_arg + row * row_stride + base);
          v[r] = __ldcs(p.t_route + row * row_stride + base);
        }
      }
      int j[kRouteRows];
#pragma unroll
      for (int r = 0; r < kRouteRows; ++r) {
        j[r] = arg[r] >= 0 ? __ldg(p.col + arg[r]) : -1;
        if (p.val_lin != nullptr && arg[r] >= 0) v[r] *= __ldg(p.val_lin + arg[r]);
      }
#pragma unroll
      for (int r = 0; r < kRouteRows; ++r) {
        if (j[r] < 0) continue;
        int hub = -1;
        if (hubs_on) {
          uint32_t h = route_hash(j[r]);
          int k = hash_key[h];
          while (k != -1 && k != j[r]) { h = (h + 1) & (kRouteHashSize - 1); k = hash_key[h]; }
          if (k == j[r]) hub = hash_slot[h];
        }
        if (hub >= 0) atomicAdd(acc + hub * 32 + lane, v[r]);
        else atomicAdd(p.d_bases + static_cast<int64_t>(j[r]) * p.BD + f, v[r]);
      }
    }
    if (hubs_on) {                                           // flush this CTA's hub partials of the phase
      __syncthreads();
      for (int t = threadIdx.x; t < p.n_hubs * 32; t += blockDim.x) {
        const float a = acc[t];
        const int ff = slab * 32 + (t & 31);
        if (a != 0.f && ff < p.BD) atomicAdd(p.d_bases + static_cast<int64_t>(__ldg(p.hubs + (t >> 5))) * p.BD + ff, a);
        acc[t] = 0.f;
      }
      __syncthreads();
    }
  }
}

// =============================================================================================
// backward pass 2: per source column (CSC), gather of the target-side streams, atomic-free
// =============================================================================================
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
  int64_t ts_row_stride;    // floats between the t-streams of consecutive target rows
  int64_t off_sym, off_lin, off_sq;   // float offset of each stream inside a row (interleaved) or of its table (stream-major)
  int BD, nvec, G, n_pass;
  int routed;               // d_bases already holds a partial result (routed min/max gradients, earlier sweeps): accumulate
  int mode;
  int* long_counter;        // [n_long] zero on entry, or null: long columns are merged by a second launch (mode 1)
  // feature-slab layout (k_scatter_slab): tstreams = [n_slabs][n_dst][n_ts][slab_w]; long_counter = [n_slabs][n_long]
  int slab_w, n_slabs;
  int64_t slab_stride;
};

constexpr int kScatterUnroll = 4;

// single-stream instances are lean enough for 4 resident CTAs (<= 64 registers); multi-stream ones get 3
template <int TSMASK, int VEC, bool LINW>
__global__ void __launch_bounds__(kAggThreads, (TSMASK & (TSMASK - 1)) == 0 ? 4 : 3) k_scatter_bwd(const __grid_constant__ ScatterParams p) {
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
  const uint32_t row_stride = static_cast<uint32_t>(p.ts_row_stride);   // n_dst * row_stride < 2^32 (checked by the host)
  const int64_t part_stride = static_cast<int64_t>(p.n_ts) * p.BD;   // chunk partials stay interleaved

  for (int pass = 0; pass < p.n_pass; ++pass) {
    const int piece = pass * 32 + (lane & (G - 1));
    const bool active = piece < p.nvec;
    const int foff = min(piece, p.nvec - 1) * VEC;
    float a_sym[VEC], a_lin[VEC], a_sq[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { a_sym[k] = 0.f; a_lin[k] = 0.f; a_sq[k] = 0.f; }

    if (p.mode == 0) {
      const float* __restrict__ src_sym = p.tstreams + p.off_sym + foff;
      const float* __restrict__ src_lin = p.tstreams + p.off_lin + foff;
      const float* __restrict__ src_sq = p.tstreams + p.off_sq + foff;
      const int last = end - 1;
      for (int e0 = begin + g; e0 < end + g; e0 += kScatterUnroll * NG) {
        int i[kScatterUnroll];
        float m[kScatterUnroll], vs[kScatterUnroll];
#pragma unroll
        for (int u = 0; u < kScatterUnroll; ++u) {
          const int e = e0 + u * NG, ec = min(e, last);
          i[u] = __ldg(p.rowidx + ec);
          m[u] = e <= last ? 1.f : 0.f;
          if constexpr (LINW) m[u] *= __ldg(p.val_lin + ec);
          vs[u] = 0.f;
          if constexpr ((TSMASK & 1) != 0) vs[u] = e <= last ? __ldg(p.val_sym + ec) : 0.f;
        }
        float xs[kScatterUnroll][VEC], xl[kScatterUnroll][VEC], xq[kScatterUnroll][VEC];
#pragma unroll
        for (int u = 0; u < kScatterUnroll; ++u) {
          const size_t r = static_cast<size_t>(static_cast<uint32_t>(i[u]) * row_stride);
          if constexpr ((TSMASK & 1) != 0) ld_row<VEC>(xs[u], src_sym + r);
          if constexpr ((TSMASK & 2) != 0) ld_row<VEC>(xl[u], src_lin + r);
          if constexpr ((TSMASK & 4) != 0) ld_row<VEC>(xq[u], src_sq + r);
        }
#pragma unroll
        for (int u = 0; u < kScatterUnroll; ++u) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            if constexpr ((TSMASK & 1) != 0) a_sym[k] = __fadd_rn(a_sym[k], __fmul_rn(xs[u][k], vs[u]));
            if constexpr ((TSMASK & 2) != 0) a_lin[k] = fmaf(xl[u][k], m[u], a_lin[k]);
            if constexpr ((TSMASK & 4) != 0) a_sq[k] = fmaf(xq[u][k], m[u], a_sq[k]);
          }
        }
      }
      for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          if constexpr ((TSMASK & 1) != 0) a_sym[k] += __shfl_xor_sync(kFull, a_sym[k], off);
          if constexpr ((TSMASK & 2) != 0) a_lin[k] += __shfl_xor_sync(kFull, a_lin[k], off);
          if constexpr ((TSMASK & 4) != 0) a_sq[k] += __shfl_xor_sync(kFull, a_sq[k], off);
        }
      }
    } else {
      const int c0 = p.long_chunk_ptr[long_idx], c1 = p.long_chunk_ptr[long_idx + 1];
#pragma unroll 4
      for (int c = c0; c < c1; ++c) {
        const float* q = p.partials + static_cast<int64_t>(c) * part_stride + foff;
        float t[VEC];
        if constexpr ((TSMASK & 1) != 0) {
          ld_plain<VEC>(t, q + p.ts_sym * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_sym[k] += t[k];
        }
        if constexpr ((TSMASK & 2) != 0) {
          ld_plain<VEC>(t, q + p.ts_lin * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_lin[k] += t[k];
        }
        if constexpr ((TSMASK & 4) != 0) {
          ld_plain<VEC>(t, q + p.ts_sq * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_sq[k] += t[k];
        }
      }
    }

    const bool writer = active && lane < G;
    if (chunk_id >= 0) {
      if (writer) {
        float* q = p.partials + static_cast<int64_t>(chunk_id) * part_stride + foff;
        if constexpr ((TSMASK & 1) != 0) st_row<VEC>(q + p.ts_sym * p.BD, a_sym);
        if constexpr ((TSMASK & 2) != 0) st_row<VEC>(q + p.ts_lin * p.BD, a_lin);
        if constexpr ((TSMASK & 4) != 0) st_row<VEC>(q + p.ts_sq * p.BD, a_sq);
      }
      if (p.long_counter == nullptr) continue;               // separate merge launch (mode 1) sums the partials
      // single-pass rows: the last chunk warp of the column to arrive sums all partials in chunk order
      __threadfence();
      __syncwarp();
      int lo = 0, hi = p.n_long;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p.long_chunk_ptr + mid) <= chunk_id) lo = mid; else hi = mid;
      }
      const int c0 = __ldg(p.long_chunk_ptr + lo), c1 = __ldg(p.long_chunk_ptr + lo + 1);
      int last = 0;
      if (lane == 0) last = atomicAdd(p.long_counter + lo, 1) == c1 - c0 - 1 ? 1 : 0;
      last = __shfl_sync(kFull, last, 0);
      if (!last) continue;
      __threadfence();
      if (lane == 0) p.long_counter[lo] = 0;
#pragma unroll
      for (int k = 0; k < VEC; ++k) { a_sym[k] = 0.f; a_lin[k] = 0.f; a_sq[k] = 0.f; }
      for (int c = c0; c < c1; ++c) {
        const float* q = p.partials + static_cast<int64_t>(c) * part_stride + foff;
        float t[VEC];
        if constexpr ((TSMASK & 1) != 0) {
          ld_cg<VEC>(t, q + p.ts_sym * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_sym[k] += t[k];
        }
        if constexpr ((TSMASK & 2) != 0) {
          ld_cg<VEC>(t, q + p.ts_lin * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_lin[k] += t[k];
        }
        if constexpr ((TSMASK & 4) != 0) {
          ld_cg<VEC>(t, q + p.ts_sq * p.BD);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a_sq[k] += t[k];
        }
      }
      chunk_id = -1;                                         // falls through to the final write of column colj
    }
    if (!writer) continue;
    float* dst = p.d_bases + static_cast<int64_t>(colj) * p.BD + foff;
    float r[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) r[k] = 0.f;
    if (p.routed) ld_plain<VEC>(r, dst);
    if constexpr ((TSMASK & 4) != 0) {
      float xj[VEC];
      ld_row<VEC>(xj, p.bases + static_cast<int64_t>(colj) * p.BD + foff);
#pragma unroll
      for (int k = 0; k < VEC; ++k) r[k] += 2.f * xj[k] * a_sq[k];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if constexpr ((TSMASK & 1) != 0) r[k] += a_sym[k];
      if constexpr ((TSMASK & 2) != 0) r[k] += a_lin[k];
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

// ---------------------------------------------------------------------------------------------
// backward pass 2, column-block variant (the CSC twin of k_aggregate_rows, aggregate_rows.cuh): a task is a block of
// kColsPerTask CONSECUTIVE source columns handed out by an atomic counter; the column pointers are one coalesced
// load, the block's target ids / symnorm weights are one contiguous range staged into the warp's shared memory with
// cp.async in windows of kColWindow entries.  The warp-per-column kernel paid colptr -> rowidx -> gather (three
// dependent global latencies) for every column of ~15 entries.  Chunks of long columns run first (strided), merged by
// the last chunk warp to arrive.  128-bit pieces, unweighted entries, one pass (BD <= 128).
// ---------------------------------------------------------------------------------------------
constexpr int kColsPerTask = 8;
constexpr int kColWindow = 384;
static_assert(kColWindow >= EGC_CHUNK_EDGES, "a normal column must fit the staging window");

template <int TSMASK, int G>
__global__ void __launch_bounds__(kAggThreads, (TSMASK & (TSMASK - 1)) == 0 ? 4 : 3) k_scatter_cols(const __grid_constant__ ScatterParams p, int* __restrict__ task_counter) {
  __shared__ int s_idx_all[kAggWarps][kColWindow];
  __shared__ float s_val_all[kAggWarps][(TSMASK & 1) ? kColWindow : 1];
  // entries in flight per lane group: single-stream instances have the registers for 8
  constexpr int NG = 32 / G, U = (TSMASK & (TSMASK - 1)) == 0 ? 2 * kScatterUnroll : kScatterUnroll, STEP = U * NG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* s_idx = s_idx_all[warp];
  float* s_val = s_val_all[warp];
  const int g = lane / G, piece = lane & (G - 1);
  const bool writer = piece < p.nvec && lane < G;
  const int foff = min(piece, p.nvec - 1) * 4;
  const uint32_t row_stride = static_cast<uint32_t>(p.ts_row_stride);   // n_dst * row_stride < 2^32 (checked by the host)
  const int64_t part_stride = static_cast<int64_t>(p.n_ts) * p.BD;
  const float* __restrict__ src_sym = p.tstreams + p.off_sym + foff;
  const float* __restrict__ src_lin = p.tstreams + p.off_lin + foff;
  const float* __restrict__ src_sq = p.tstreams + p.off_sq + foff;
  float a_sym[4], a_lin[4], a_sq[4];

  auto stage = [&](int wb, int we) {                           // entries [wb, we) -> shared memory (we - wb <= kColWindow)
    for (int i = lane; i < we - wb; i += 32) {
      cp_async_4(s_idx + i, p.rowidx + wb + i);
      if constexpr ((TSMASK & 1) != 0) cp_async_4(s_val + i, p.val_sym + wb + i);
    }
    cp_async_wait_all();
    __syncwarp();
  };
  // sums over the staged entries [b, e) of one column (window base wb), then the xor-merge of the lane groups
  auto accumulate = [&](int b, int e, int wb) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { a_sym[k] = 0.f; a_lin[k] = 0.f; a_sq[k] = 0.f; }
    for (int pos = b; pos < e; pos += STEP) {
      float m[U], vs[U];
      float4 xs[U], xl[U], xq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = pos + u * NG + g, qc = min(q, e - 1);
        const size_t r = static_cast<size_t>(static_cast<uint32_t>(s_idx[qc - wb]) * row_stride);
        m[u] = q < e ? 1.f : 0.f;
        vs[u] = 0.f;
        if constexpr ((TSMASK & 1) != 0) { vs[u] = q < e ? s_val[qc - wb] : 0.f; xs[u] = __ldg(reinterpret_cast<const float4*>(src_sym + r)); }
        if constexpr ((TSMASK & 2) != 0) xl[u] = __ldg(reinterpret_cast<const float4*>(src_lin + r));
        if constexpr ((TSMASK & 4) != 0) xq[u] = __ldg(reinterpret_cast<const float4*>(src_sq + r));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if constexpr ((TSMASK & 1) != 0) {
          a_sym[0] = __fadd_rn(a_sym[0], __fmul_rn(xs[u].x, vs[u])); a_sym[1] = __fadd_rn(a_sym[1], __fmul_rn(xs[u].y, vs[u]));
          a_sym[2] = __fadd_rn(a_sym[2], __fmul_rn(xs[u].z, vs[u])); a_sym[3] = __fadd_rn(a_sym[3], __fmul_rn(xs[u].w, vs[u]));
        }
        if constexpr ((TSMASK & 2) != 0) {
          a_lin[0] = fmaf(xl[u].x, m[u], a_lin[0]); a_lin[1] = fmaf(xl[u].y, m[u], a_lin[1]);
          a_lin[2] = fmaf(xl[u].z, m[u], a_lin[2]); a_lin[3] = fmaf(xl[u].w, m[u], a_lin[3]);
        }
        if constexpr ((TSMASK & 4) != 0) {
          a_sq[0] = fmaf(xq[u].x, m[u], a_sq[0]); a_sq[1] = fmaf(xq[u].y, m[u], a_sq[1]);
          a_sq[2] = fmaf(xq[u].z, m[u], a_sq[2]); a_sq[3] = fmaf(xq[u].w, m[u], a_sq[3]);
        }
      }
    }
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr ((TSMASK & 1) != 0) a_sym[k] += __shfl_xor_sync(kFull, a_sym[k], off);
        if constexpr ((TSMASK & 2) != 0) a_lin[k] += __shfl_xor_sync(kFull, a_lin[k], off);
        if constexpr ((TSMASK & 4) != 0) a_sq[k] += __shfl_xor_sync(kFull, a_sq[k], off);
      }
    }
  };
  auto write_col = [&](int colj) {
    if (!writer) return;
    float* dst = p.d_bases + static_cast<int64_t>(colj) * p.BD + foff;
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.routed) ld_plain<4>(r, dst);
    if constexpr ((TSMASK & 4) != 0) {
      float xj[4];
      ld_row<4>(xj, p.bases + static_cast<int64_t>(colj) * p.BD + foff);
#pragma unroll
      for (int k = 0; k < 4; ++k) r[k] += 2.f * xj[k] * a_sq[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if constexpr ((TSMASK & 1) != 0) r[k] += a_sym[k];
      if constexpr ((TSMASK & 2) != 0) r[k] += a_lin[k];
    }
    st_row<4>(dst, r);
  };

  // =========================== phase 0: chunks of long columns (strided over the grid) ===========================
  {
    const int warps_total = gridDim.x * kAggWarps;
    for (int chunk_id = blockIdx.x * kAggWarps + warp; chunk_id < p.n_chunks; chunk_id += warps_total) {
      const int colj = __ldg(p.chunk_row + chunk_id);
      const int begin = __ldg(p.chunk_begin + chunk_id);
      const int end = min(begin + EGC_CHUNK_EDGES, __ldg(p.colptr + colj + 1));
      stage(begin, end);
      accumulate(begin, end, begin);
      if (writer) {
        float* q = p.partials + static_cast<int64_t>(chunk_id) * part_stride + foff;
        if constexpr ((TSMASK & 1) != 0) st_row<4>(q + p.ts_sym * p.BD, a_sym);
        if constexpr ((TSMASK & 2) != 0) st_row<4>(q + p.ts_lin * p.BD, a_lin);
        if constexpr ((TSMASK & 4) != 0) st_row<4>(q + p.ts_sq * p.BD, a_sq);
      }
      // the last chunk warp of the column to arrive sums all partials in chunk order and writes the column
      __threadfence();
      __syncwarp();
      int lo = 0, hi = p.n_long;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p.long_chunk_ptr + mid) <= chunk_id) lo = mid; else hi = mid;
      }
      const int c0 = __ldg(p.long_chunk_ptr + lo), c1 = __ldg(p.long_chunk_ptr + lo + 1);
      int last = 0;
      if (lane == 0) last = atomicAdd(p.long_counter + lo, 1) == c1 - c0 - 1 ? 1 : 0;
      last = __shfl_sync(kFull, last, 0);
      if (!last) continue;
      __threadfence();
      if (lane == 0) p.long_counter[lo] = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) { a_sym[k] = 0.f; a_lin[k] = 0.f; a_sq[k] = 0.f; }
      for (int c = c0; c < c1; ++c) {
        const float* q = p.partials + static_cast<int64_t>(c) * part_stride + foff;
        float t[4];
        if constexpr ((TSMASK & 1) != 0) {
          ld_cg<4>(t, q + p.ts_sym * p.BD);
#pragma unroll
          for (int k = 0; k < 4; ++k) a_sym[k] += t[k];
        }
        if constexpr ((TSMASK & 2) != 0) {
          ld_cg<4>(t, q + p.ts_lin * p.BD);
#pragma unroll
          for (int k = 0; k < 4; ++k) a_lin[k] += t[k];
        }
        if constexpr ((TSMASK & 4) != 0) {
          ld_cg<4>(t, q + p.ts_sq * p.BD);
#pragma unroll
          for (int k = 0; k < 4; ++k) a_sq[k] += t[k];
        }
      }
      write_col(colj);
    }
  }

  // =========================== phase 1: blocks of consecutive columns (dynamic) ===========================
  const int n_blocks = (p.n_cols + kColsPerTask - 1) / kColsPerTask;
  int task = 0;
  if (lane == 0) task = atomicAdd(task_counter, 1);
  task = __shfl_sync(kFull, task, 0);
  while (task < n_blocks) {
    const int c0 = task * kColsPerTask;
    const int ncols = min(kColsPerTask, p.n_cols - c0);
    const int cp = __ldg(p.colptr + c0 + min(lane, ncols));    // lanes 0..ncols hold the block's column pointers
    int next_task = 0;
    if (lane == 0) next_task = atomicAdd(task_counter, 1);      // consumed at the end of this task
    const int cpn = __shfl_down_sync(kFull, cp, 1);             // lane l < ncols: column l = [cp, cpn)
    const unsigned long_mask = __ballot_sync(kFull, lane < ncols && cpn - cp > EGC_CHUNK_EDGES);
    int ci = 0;
    while (ci < ncols) {
      if ((long_mask >> ci) & 1u) { ++ci; continue; }           // long column: its chunk tasks did it
      const unsigned later_long = long_mask >> ci;
      const int limit = later_long != 0u ? ci + __ffs(later_long) - 1 : ncols;
      const int wb = __shfl_sync(kFull, cp, ci);
      const unsigned fit = __ballot_sync(kFull, lane >= ci && lane < limit && cpn - wb <= kColWindow);
      const int n_fit = __popc(fit);                            // >= 1
      const int we = __shfl_sync(kFull, cpn, ci + n_fit - 1);
      __syncwarp();                                             // every lane is done with the previous window
      stage(wb, we);
      for (int c = ci; c < ci + n_fit; ++c) {
        const int b = __shfl_sync(kFull, cp, c), e = __shfl_sync(kFull, cpn, c);
        accumulate(b, e, wb);
        write_col(c0 + c);
      }
      ci += n_fit;
    }
    task = __shfl_sync(kFull, next_task, 0);
  }
}

template <int G>
static int launch_scatter_cols(const ScatterParams& p, int tsmask, int* task_counter, cudaStream_t st) {
  const int n_blocks = ceil_div(p.n_cols, kColsPerTask);
  const int resident = (tsmask & (tsmask - 1)) == 0 ? 4 : 3;     // CTAs per SM, as the launch bounds
  const int grid = std::max(1, std::min(ceil_div(std::max(n_blocks, p.n_chunks), kAggWarps), sm_count() * resident));
  {
    LaunchScope egc_ls_("k_scatter_bwd", st);
    switch (tsmask) {
      case 1: k_scatter_cols<1, G><<<grid, kAggThreads, 0, st>>>(p, task_counter); break;
      case 2: k_scatter_cols<2, G><<<grid, kAggThreads, 0, st>>>(p, task_counter); break;
      case 3: k_scatter_cols<3, G><<<grid, kAggThreads, 0, st>>>(p, task_counter); break;
      case 4: k_scatter_cols<4, G><<<grid, kAggThreads, 0, st>>>(p, task_counter); break;
      case 5: k_scatter_cols<5, G><<<grid, kAggThreads, 0, st>>>(p, task_counter); break;
      case 6: k_scatter_cols<6, G><<<grid, kAggThreads, 0, st>>>(p, task_counter); break;
      case 7: k_scatter_cols<7, G><<<grid, kAggThreads, 0, st>>>(p, task_counter); break;
      default: set_error("scatter_cols: bad stream mask %d", tsmask); return EGC_ERR_UNSUPPORTED;
    }
  }
  EGC_LAUNCH_CHECK("k_scatter_cols");
  return EGC_OK;
}

static int launch_scatter(const ScatterParams& p, int tsmask, bool vec4, bool linw, cudaStream_t st) {
  if (vec4) return linw ? launch_scatter_mask<4, true>(p, tsmask, st) : launch_scatter_mask<4, false>(p, tsmask, st);
  return linw ? launch_scatter_mask<1, true>(p, tsmask, st) : launch_scatter_mask<1, false>(p, tsmask, st);
}


// =============================================================================================
// backward pass 2, feature-slab variant.  When the target-side streams overflow the L2 (cfg2: 3 x 87 MB),
// pass 1 stores them slab-major - [slab][target][stream][W floats], W = 16 or 32 - and this kernel sweeps
// the CSC once per slab, so every sweep gathers from a table of n_dst x n_ts x W x 4 bytes that stays
// L2-resident (cfg2, W = 16: 32.5 MB) instead of missing to HBM on nearly every entry.
//   * persistent warps; each owns a CONTIGUOUS range of columns (and of the long columns' 256-entry chunks)
//     of equal key mass, key = first entry + column id, found by two 32-ary searches of colptr.  Every warp
//     does the same amount of work per slab, so the grid moves from slab to slab together, and a warp's
//     index reads (colptr, rowidx, val_sym) are sequential;
//   * G = W / 4 lanes cover one entry (all streams: n_ts consecutive 16-byte pieces W floats apart), the
//     32 / G lane groups walk different entries and are merged with xor-shuffles per column;
//   * row ids / symnorm weights of 32 consecutive entries sit in one register per lane (one coalesced load)
//     and are broadcast with shuffles, whatever column boundaries fall inside.
// Requires a plan whose chunks are ordered by first entry (egc_plan_build's order).
// =============================================================================================
constexpr int kSlabMaxSlabs = 16;

// first i in [0, n] with key(i) >= target (key non-decreasing, key(n) = +inf): 32 probes per round
template <class KeyF>
__device__ __forceinline__ int warp_lower_bound(KeyF key, int n, int64_t target, int lane) {
  int lo = 0, hi = n;                                   // the answer lies in [lo, hi]
  while (lo < hi) {
    const int step = (hi - lo + 31) >> 5;
    const int probe = lo + lane * step;
    const bool ge = probe >= hi || key(probe) >= target;
    const unsigned m = __ballot_sync(kFull, ge);
    const int f = m != 0u ? __ffs(m) - 1 : 32;          // first probe at or past the target
    if (f == 0) {
      hi = lo;
    } else {
      const int nlo = lo + (f - 1) * step + 1, nhi = min(lo + f * step, hi);
      lo = nlo;
      hi = nhi;
    }
  }
  return lo;
}

template <int TSMASK, int W>
__global__ void __launch_bounds__(kAggThreads, 4) k_scatter_slab(const __grid_constant__ ScatterParams p) {
  constexpr int G = W / 4, NG = 32 / G;
  constexpr int NS = ((TSMASK & 1) ? 1 : 0) + ((TSMASK & 2) ? 1 : 0) + ((TSMASK & 4) ? 1 : 0);
  constexpr int U = NS >= 2 ? 2 : 4;
  constexpr int BATCH = U * NG;
  static_assert(BATCH <= 32, "one batch must fit the 32-entry index buffer");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kAggWarps + warp, warps_total = gridDim.x * kAggWarps;
  const int g = lane / G, foff = (lane & (G - 1)) * 4;
  const int nnz = __ldg(p.colptr + p.n_cols);

  // ---- this warp's share: keys [k_begin, k_end) of the (first entry + column id) axis
  const int64_t total_key = static_cast<int64_t>(nnz) + p.n_cols;
  const int64_t q = (total_key + warps_total - 1) / warps_total;
  const int64_t k_begin = q * gw, k_end = k_begin + q;
  auto col_key = [&](int c) { return static_cast<int64_t>(__ldg(p.colptr + c)) + c; };
  const int c_lo = warp_lower_bound(col_key, p.n_cols, k_begin, lane);
  const int c_hi = warp_lower_bound(col_key, p.n_cols, k_end, lane);
  int k_lo = 0, k_hi = 0;
  if (p.n_chunks > 0) {
    auto chunk_key = [&](int k) { return static_cast<int64_t>(__ldg(p.chunk_begin + k)) + __ldg(p.chunk_row + k); };
    k_lo = warp_lower_bound(chunk_key, p.n_chunks, k_begin, lane);
    k_hi = warp_lower_bound(chunk_key, p.n_chunks, k_end, lane);
  }
  const int n_my_chunks = k_hi - k_lo, n_my = n_my_chunks + (c_hi - c_lo);
  if (n_my == 0) return;

  const uint64_t pol_keep = l2_policy_keep();
  const uint32_t row_stride = static_cast<uint32_t>(p.n_ts) * W;
  const int64_t part_stride = static_cast<int64_t>(p.n_ts) * p.BD;
  int buf_base = -(1 << 30), my_i = 0;                   // entries [buf_base, buf_base + 32): row ids / symnorm weights
  float my_vs = 0.f;
  int cp_base = -(1 << 30), cp0 = 0, cp1 = 0;            // colptr of columns [cp_base, cp_base + 32] (begin / end)

  for (int slab = 0; slab < p.n_slabs; ++slab) {
    const float* __restrict__ ts = p.tstreams + static_cast<int64_t>(slab) * p.slab_stride + foff;
    const float* __restrict__ src_sym = ts + max(p.ts_sym, 0) * W;
    const float* __restrict__ src_lin = ts + max(p.ts_lin, 0) * W;
    const float* __restrict__ src_sq = ts + max(p.ts_sq, 0) * W;
    const int fcol = slab * W + foff;                    // this lane's first feature inside a [BD] row

    for (int t = 0; t < n_my; ++t) {
      int colj, begin, end, chunk_id = -1;
      if (t < n_my_chunks) {
        chunk_id = k_lo + t;
        colj = __ldg(p.chunk_row + chunk_id);
        begin = __ldg(p.chunk_begin + chunk_id);
        end = min(begin + EGC_CHUNK_EDGES, __ldg(p.colptr + colj + 1));
      } else {
        colj = c_lo + (t - n_my_chunks);
        if (colj < cp_base || colj >= cp_base + 32) {
          cp_base = colj;
          cp0 = __ldg(p.colptr + min(colj + lane, p.n_cols));
          cp1 = __ldg(p.colptr + min(colj + lane + 1, p.n_cols));
        }
        begin = __shfl_sync(kFull, cp0, colj - cp_base);
        end = __shfl_sync(kFull, cp1, colj - cp_base);
        if (end - begin > EGC_CHUNK_EDGES) continue;       // long column: its chunk tasks do it
      }

      float a_sym[4] = {0.f, 0.f, 0.f, 0.f}, a_lin[4] = {0.f, 0.f, 0.f, 0.f}, a_sq[4] = {0.f, 0.f, 0.f, 0.f};
      for (int e0 = begin; e0 < end; e0 += BATCH) {
        if (e0 < buf_base || e0 + BATCH > buf_base + 32) {
          buf_base = e0;
          const int ec = min(e0 + lane, nnz - 1);
          my_i = __ldg(p.rowidx + ec);
          if constexpr ((TSMASK & 1) != 0) my_vs = __ldg(p.val_sym + ec);
        }
        uint32_t iu[U];
        float vsu[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int e = e0 + u * NG + g;
          ok[u] = e < end;
          const int sl = min(e, end - 1) - buf_base;
          iu[u] = static_cast<uint32_t>(__shfl_sync(kFull, my_i, sl));
          vsu[u] = 0.f;
          if constexpr ((TSMASK & 1) != 0) vsu[u] = __shfl_sync(kFull, my_vs, sl);
        }
        float4 xs[U], xl[U], xq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          xs[u] = xl[u] = xq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e0 + u * NG < end) {                         // warp-uniform: some group has a real entry in this slot
            const size_t r = static_cast<size_t>(iu[u] * row_stride);
            if constexpr ((TSMASK & 1) != 0) xs[u] = ldg_f4_hint(src_sym + r, pol_keep);
            if constexpr ((TSMASK & 2) != 0) xl[u] = ldg_f4_hint(src_lin + r, pol_keep);
            if constexpr ((TSMASK & 4) != 0) xq[u] = ldg_f4_hint(src_sq + r, pol_keep);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (ok[u]) {
            if constexpr ((TSMASK & 1) != 0) {
              a_sym[0] = __fadd_rn(a_sym[0], __fmul_rn(xs[u].x, vsu[u])); a_sym[1] = __fadd_rn(a_sym[1], __fmul_rn(xs[u].y, vsu[u]));
              a_sym[2] = __fadd_rn(a_sym[2], __fmul_rn(xs[u].z, vsu[u])); a_sym[3] = __fadd_rn(a_sym[3], __fmul_rn(xs[u].w, vsu[u]));
            }
            if constexpr ((TSMASK & 2) != 0) { a_lin[0] += xl[u].x; a_lin[1] += xl[u].y; a_lin[2] += xl[u].z; a_lin[3] += xl[u].w; }
            if constexpr ((TSMASK & 4) != 0) { a_sq[0] += xq[u].x; a_sq[1] += xq[u].y; a_sq[2] += xq[u].z; a_sq[3] += xq[u].w; }
          }
        }
      }
#pragma unroll
      for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr ((TSMASK & 1) != 0) a_sym[k] += __shfl_xor_sync(kFull, a_sym[k], off);
          if constexpr ((TSMASK & 2) != 0) a_lin[k] += __shfl_xor_sync(kFull, a_lin[k], off);
          if constexpr ((TSMASK & 4) != 0) a_sq[k] += __shfl_xor_sync(kFull, a_sq[k], off);
        }
      }

      const bool writer = lane < G;
      if (chunk_id >= 0) {
        // a chunk of a long column: publish the partial; the LAST chunk warp of (slab, column) to arrive sums all of
        // them in chunk order and writes the column
        if (writer) {
          float* qd = p.partials + static_cast<int64_t>(chunk_id) * part_stride + fcol;
          if constexpr ((TSMASK & 1) != 0) st_row<4>(qd + p.ts_sym * p.BD, a_sym);
          if constexpr ((TSMASK & 2) != 0) st_row<4>(qd + p.ts_lin * p.BD, a_lin);
          if constexpr ((TSMASK & 4) != 0) st_row<4>(qd + p.ts_sq * p.BD, a_sq);
        }
        __threadfence();
        __syncwarp();
        int lo = 0, hi = p.n_long;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(p.long_chunk_ptr + mid) <= chunk_id) lo = mid; else hi = mid;
        }
        const int c0 = __ldg(p.long_chunk_ptr + lo), c1 = __ldg(p.long_chunk_ptr + lo + 1);
        int* counter = p.long_counter + static_cast<int64_t>(slab) * p.n_long + lo;
        int last = 0;
        if (lane == 0) last = atomicAdd(counter, 1) == c1 - c0 - 1 ? 1 : 0;
        last = __shfl_sync(kFull, last, 0);
        if (!last) continue;
        __threadfence();
        if (lane == 0) *counter = 0;                         // ready for the next launch
#pragma unroll
        for (int k = 0; k < 4; ++k) { a_sym[k] = 0.f; a_lin[k] = 0.f; a_sq[k] = 0.f; }
        if (writer) {
          for (int c = c0; c < c1; ++c) {
            const float* qs = p.partials + static_cast<int64_t>(c) * part_stride + fcol;
            float tv[4];
            if constexpr ((TSMASK & 1) != 0) { ld_cg<4>(tv, qs + p.ts_sym * p.BD); for (int k = 0; k < 4; ++k) a_sym[k] += tv[k]; }
            if constexpr ((TSMASK & 2) != 0) { ld_cg<4>(tv, qs + p.ts_lin * p.BD); for (int k = 0; k < 4; ++k) a_lin[k] += tv[k]; }
            if constexpr ((TSMASK & 4) != 0) { ld_cg<4>(tv, qs + p.ts_sq * p.BD); for (int k = 0; k < 4; ++k) a_sq[k] += tv[k]; }
          }
        }
      }
      if (!writer) continue;
      float* dst = p.d_bases + static_cast<int64_t>(colj) * p.BD + fcol;
      float r[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.routed) ld_plain<4>(r, dst);
      if constexpr ((TSMASK & 4) != 0) {
        float xj[4];
        ld_row<4>(xj, p.bases + static_cast<int64_t>(colj) * p.BD + fcol);
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] += 2.f * xj[k] * a_sq[k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr ((TSMASK & 1) != 0) r[k] += a_sym[k];
        if constexpr ((TSMASK & 2) != 0) r[k] += a_lin[k];
      }
      st_row<4>(dst, r);
    }
  }
}

template <int W>
static int launch_scatter_slab(const ScatterParams& p, int tsmask, cudaStream_t st) {
  const int grid = sm_count() * 4;
  LaunchScope egc_ls_("k_scatter_bwd", st);
  switch (tsmask) {
    case 1: k_scatter_slab<1, W><<<grid, kAggThreads, 0, st>>>(p); break;
    case 2: k_scatter_slab<2, W><<<grid, kAggThreads, 0, st>>>(p); break;
    case 3: k_scatter_slab<3, W><<<grid, kAggThreads, 0, st>>>(p); break;
    case 4: k_scatter_slab<4, W><<<grid, kAggThreads, 0, st>>>(p); break;
    case 5: k_scatter_slab<5, W><<<grid, kAggThreads, 0, st>>>(p); break;
    case 6: k_scatter_slab<6, W><<<grid, kAggThreads, 0, st>>>(p); break;
    case 7: k_scatter_slab<7, W><<<grid, kAggThreads, 0, st>>>(p); break;
    default: set_error("scatter_bwd: bad stream mask %d", tsmask); return EGC_ERR_UNSUPPORTED;
  }
  return EGC_OK;
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

static int combine_bwd_grid(int n_rows) { return std::max(1, std::min(ceil_div(n_rows, kAggWarps), sm_count() * 4)); }

struct BwdLayout {
  size_t ts_bytes, csc_part_bytes, colsum_bytes, route_bytes, total;
  int n_ts, ts_sym, ts_lin, ts_sq, tsmask;
  bool has_route;
};

static BwdLayout bwd_layout(const egc_layer_desc& d, const egc_row_plan* csc_plan) {
  BwdLayout L{};
  L.tsmask = stream_mask_of(d, L.has_route);
  int s = 0;
  L.ts_sym = (L.tsmask & 1) ? s++ : -1;
  L.ts_lin = (L.tsmask & 2) ? s++ : -1;
  L.ts_sq = (L.tsmask & 4) ? s++ : -1;
  L.n_ts = s;
  const size_t bd = static_cast<size_t>(d.bases) * d.dim;
  L.ts_bytes = align_up(static_cast<size_t>(d.n_dst) * std::max(L.n_ts, 1) * bd * 4, 256);
  L.csc_part_bytes = align_up(static_cast<size_t>(csc_plan ? csc_plan->n_chunks : 0) * std::max(L.n_ts, 1) * bd * 4, 256) +
                     align_up(static_cast<size_t>(csc_plan ? csc_plan->n_long : 0) * kSlabMaxSlabs * sizeof(int), 256) +
                     256;   // + the task counter of k_scatter_cols
  // per-CTA column-sum partials of the fused pass-1 kernel, or the two-stage colsum scratch when it cannot fuse
  const size_t hd = static_cast<size_t>(d.heads) * d.dim, hab = static_cast<size_t>(d.heads) * d.n_aggr * d.bases;
  const size_t fused = static_cast<size_t>(combine_bwd_grid(d.n_dst)) * (hd + hab) * sizeof(float) + 256;
  L.colsum_bytes = align_up(std::max(fused, std::max(colsum_workspace_bytes(d.n_dst, static_cast<int>(hd)),
                                                     colsum_workspace_bytes(d.n_dst, static_cast<int>(hab)))), 256);
  L.route_bytes = align_up(static_cast<size_t>(d.n_dst) * n_arg_slots(d) * bd * 4, 256);
  L.total = L.ts_bytes + L.csc_part_bytes + L.colsum_bytes + L.route_bytes + 256;
  return L;
}

}  // namespace egc

using namespace egc;

extern "C" {

int32_t egc_saved_slots(const egc_layer_desc* desc) { return desc ? n_saved_slots(*desc) : 0; }
int32_t egc_saved_arg_slots(const egc_layer_desc* desc) { return desc ? n_arg_slots(*desc) : 0; }

size_t egc_aggregate_fwd_workspace_bytes(const egc_layer_desc* desc, const egc_row_plan* plan) {
  if (desc == nullptr) return 0;
  const int mask = prim_mask_of(*desc);
  if (mask <= 0) return 0;
  const size_t bd = static_cast<size_t>(desc->bases) * desc->dim;
  return align_up(static_cast<size_t>(plan ? plan->n_chunks : 0) * n_slots_of_mask(mask) * bd * 4, 256) +
         align_up(static_cast<size_t>(plan ? plan->n_long : 0) * sizeof(int), 256) + 256;
}

int egc_aggregate_fwd(const egc_layer_desc* desc, const int32_t* rowptr, const int32_t* col, const float* val_sym,
                      const float* val_lin, const egc_row_plan* plan, const float* bases, const float* weightings,
                      const float* bias, const int32_t* row_subset, int32_t n_subset, float* out, float* agg_out,
                      int32_t* arg_out, float* saved, int32_t* saved_arg, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (int rc = validate_desc(desc, "egc_aggregate_fwd")) return rc;
  if (int rc = validate_plan(plan, "egc_aggregate_fwd")) return rc;
  EGC_REQUIRE(rowptr && col && bases, "egc_aggregate_fwd: null graph / bases pointer");
  EGC_REQUIRE(out == nullptr || weightings != nullptr, "egc_aggregate_fwd: weightings required to produce out");
  EGC_REQUIRE(out || agg_out || arg_out || saved, "egc_aggregate_fwd: no output requested");
  EGC_REQUIRE(n_subset >= 0 && (n_subset == 0 || row_subset != nullptr), "egc_aggregate_fwd: bad row subset");
  const int mask = prim_mask_of(*desc);
  EGC_REQUIRE(!(mask & P_SYM) || val_sym != nullptr, "egc_aggregate_fwd: symnorm requested without val_sym");
  EGC_REQUIRE(!((mask & P_SYM) && val_lin), "egc_aggregate_fwd: val_lin cannot be combined with symnorm (ref :253-254)");
  EGC_REQUIRE(workspace_bytes >= egc_aggregate_fwd_workspace_bytes(desc, plan) && (workspace || !(plan && plan->n_chunks)),
              "egc_aggregate_fwd: workspace too small");
  const int n_arg = n_arg_slots(*desc);
  EGC_REQUIRE(saved == nullptr || n_arg == 0 || saved_arg != nullptr, "egc_aggregate_fwd: saved_arg required with saved for min/max");
  AggParams p{};
  const int bd = desc->bases * desc->dim;
  const bool vec4 = (bd % 4 == 0) && aligned16(bases) && aligned16(out) && aligned16(agg_out) && aligned16(workspace) &&
                    aligned16(arg_out) && aligned16(saved) && aligned16(saved_arg);
  const int smem = fill_agg_params(p, *desc, vec4);
  EGC_REQUIRE(smem <= 200 * 1024, "egc_aggregate_fwd: layer too wide for the shared-memory staging (%d bytes)", smem);
  p.rowptr = rowptr; p.col = col; p.val_sym = val_sym; p.val_lin = val_lin; p.n_rows = desc->n_dst;
  p.row_map = row_subset;
  p.n_row_tasks = row_subset != nullptr ? n_subset : desc->n_dst;
  set_plan(p, plan);
  p.partials = static_cast<float*>(workspace);
  p.n_slots = n_slots_of_mask(mask);
  p.long_counter = reinterpret_cast<int*>(static_cast<char*>(workspace) +
                                          align_up(static_cast<size_t>(p.n_chunks) * p.n_slots * bd * 4, 256));
  p.bases = bases; p.weightings = weightings; p.bias = bias;
  p.out = out; p.agg_out = agg_out; p.arg_out = arg_out; p.saved = saved; p.saved_arg = saved_arg;
  cudaStream_t st = as_stream(stream);
  const bool want_arg = (arg_out != nullptr || saved_arg != nullptr) && n_arg > 0;
  if (arg_out != nullptr && !want_arg && row_subset == nullptr)   // no min/max slot: every arg is "none"
    EGC_CUDA(cudaMemsetAsync(arg_out, 0xff, static_cast<size_t>(desc->n_dst) * desc->n_aggr * bd * sizeof(int32_t), st));
  auto launch = vec4 ? launch_aggregate_v4 : launch_aggregate_v1;
  p.mode = 0;
  const int hd = desc->heads * desc->dim;
  const bool fast = vec4 && val_lin == nullptr && p.n_pass == 1 && (p.G == 32 || p.G == 16) &&
                    hd <= ((desc->dim % 4 == 0) ? 512 : 128) && static_cast<int64_t>(desc->n_src) * bd < (int64_t{1} << 32) && (desc->dim % 4 != 0 || aligned16(bias));
  const int static_idx = fast ? static_cfg_index(*desc) : -1;
  // row-block kernel: whole-graph calls of the specialised shapes (EGC_FWD_WARP_PER_ROW=1 keeps the warp-per-row kernel)
  static const bool legacy_rows = getenv("EGC_FWD_WARP_PER_ROW") != nullptr;
  // needs the window + 8 rows of weights per warp in shared memory (3 CTAs per SM)
  const bool row_blocks = fast && !legacy_rows && rows_kernel_smem_bytes(p) <= 72 * 1024;
  const size_t counter_off = align_up(static_cast<size_t>(p.n_long) * sizeof(int), 256);   // inside the trailing 256 B
  int* task_counter = reinterpret_cast<int*>(reinterpret_cast<char*>(p.long_counter) + counter_off);
  if (row_blocks) EGC_CUDA(cudaMemsetAsync(p.long_counter, 0, counter_off + sizeof(int), st));
  else if (fast && p.n_long > 0) EGC_CUDA(cudaMemsetAsync(p.long_counter, 0, static_cast<size_t>(p.n_long) * sizeof(int), st));
  if (row_blocks && static_idx >= 0) {
    if (int rc = launch_aggregate_rows_static(static_idx, p, want_arg, task_counter, st)) return rc;
  } else if (row_blocks) {
    if (int rc = (p.G == 32 ? launch_aggregate_rows_g32 : launch_aggregate_rows_g16)(p, mask, want_arg, task_counter, st)) return rc;
  } else if (static_idx >= 0) {
    if (int rc = launch_aggregate_fast_static(static_idx, p, want_arg, smem, st)) return rc;
  } else if (fast) {
    if (int rc = (p.G == 32 ? launch_aggregate_fast_g32 : launch_aggregate_fast_g16)(p, mask, want_arg, smem, st)) return rc;
  } else {
    if (int rc = launch(p, mask, val_lin != nullptr, want_arg, smem, st)) return rc;
  }
  if (p.n_long > 0 && !fast) {                  // the fast kernels merge long rows themselves
    p.mode = 1;
    if (int rc = launch(p, mask, val_lin != nullptr, want_arg, smem, st)) return rc;
  }
  return EGC_OK;
}

size_t egc_aggregate_bwd_workspace_bytes(const egc_layer_desc* desc, const egc_row_plan* csc_plan, int32_t flags) {
  (void)flags;
  if (desc == nullptr || prim_mask_of(*desc) <= 0) return 0;
  return bwd_layout(*desc, csc_plan).total;
}

int egc_aggregate_bwd(const egc_layer_desc* desc, const int32_t* rowptr, const int32_t* col, const float* val_lin,
                      const int32_t* colptr, const int32_t* rowidx, const float* csc_val_sym,
                      const float* csc_val_lin, const egc_row_plan* csc_plan, const float* bases,
                      const float* weightings, const float* saved, const int32_t* saved_arg, const float* grad_out,
                      float* d_weightings, float* d_bases, float* d_bias, float* d_lin_colsum, int32_t flags,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = validate_desc(desc, "egc_aggregate_bwd")) return rc;
  if (int rc = validate_plan(csc_plan, "egc_aggregate_bwd")) return rc;
  EGC_REQUIRE(rowptr && col && colptr && rowidx && bases && weightings && saved && grad_out && d_weightings && d_bases && workspace,
              "egc_aggregate_bwd: null pointer");
  const BwdLayout L = bwd_layout(*desc, csc_plan);
  if ((flags & EGC_BWD_DETERMINISTIC) && L.has_route) {
    set_error("egc_aggregate_bwd: EGC_BWD_DETERMINISTIC routing of min/max gradients is not built yet");
    return EGC_ERR_UNSUPPORTED;
  }
  const int mask = prim_mask_of(*desc);
  const int n_arg = n_arg_slots(*desc);
  EGC_REQUIRE(n_arg == 0 || saved_arg != nullptr, "egc_aggregate_bwd: saved_arg required for min/max");
  EGC_REQUIRE(!(mask & P_SYM) || csc_val_sym, "egc_aggregate_bwd: symnorm requested without csc_val_sym");
  EGC_REQUIRE((val_lin == nullptr) == (csc_val_lin == nullptr), "egc_aggregate_bwd: val_lin and csc_val_lin must come together");
  EGC_REQUIRE(!((mask & P_SYM) && val_lin), "egc_aggregate_bwd: val_lin cannot be combined with symnorm");
  EGC_REQUIRE(workspace_bytes >= L.total, "egc_aggregate_bwd: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  EGC_REQUIRE(L.ts_bytes / 4 < (size_t{1} << 32), "egc_aggregate_bwd: target-side streams exceed 2^32 floats");
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  float* tstreams = reinterpret_cast<float*>(ws);
  float* csc_part = reinterpret_cast<float*>(ws + L.ts_bytes);
  void* colsum_ws = ws + L.ts_bytes + L.csc_part_bytes;
  float* t_route = reinterpret_cast<float*>(ws + L.ts_bytes + L.csc_part_bytes + L.colsum_bytes);

  const int bd = desc->bases * desc->dim, hd = desc->heads * desc->dim;
  const int hab = desc->heads * desc->n_aggr * desc->bases;
  const bool vec4 = (bd % 4 == 0) && aligned16(bases) && aligned16(d_bases) && aligned16(workspace);

  if (L.has_route || L.tsmask == 0)
    EGC_CUDA(cudaMemsetAsync(d_bases, 0, static_cast<size_t>(desc->n_src) * bd * sizeof(float), st));

  // Target-side stream layout.  When the streams together overflow the L2 but one of them fits, they are stored
  // stream-major ([L][N][BD]) and pass 2 sweeps them one at a time, so every sweep gathers from an L2-resident
  // table (the partial d_bases is re-read by the later sweeps); otherwise interleaved ([N][L][BD]), one sweep.
  const size_t one_stream = static_cast<size_t>(desc->n_dst) * bd * sizeof(float);
  const bool stream_major = L.n_ts >= 2 && one_stream * L.n_ts > (size_t{72} << 20) && one_stream <= (size_t{100} << 20) &&
                            (flags & EGC_BWD_STREAM_SWEEPS) != 0;

  // Feature-slab layout ([slab][N][L][W], see k_scatter_slab): an opt-in tuning layout, EGC_BWD_SLAB16 / SLAB32 pick the
  // width; the default (and EGC_BWD_NO_SLABS) is the plain interleaved layout.
  int slab_w = 0;
  {
    const bool eligible = vec4 && val_lin == nullptr && !stream_major && L.tsmask != 0 && (flags & EGC_BWD_NO_SLABS) == 0;
    auto fits = [&](int w) { return bd % w == 0 && bd / w >= 2 && bd / w <= kSlabMaxSlabs; };
    if (eligible) {
      if ((flags & EGC_BWD_SLAB32) && fits(32)) slab_w = 32;
      else if ((flags & EGC_BWD_SLAB16) && fits(16)) slab_w = 16;
      // No automatic choice any more: measured on B200 (profiles/r01e_bwd_layouts.txt) the column-block kernel on the
      // plain interleaved layout beats every slab layout (arxiv shape 0.49 vs 0.60 ms, mag shape 0.39 vs 0.54 ms).
    }
  }

  // ---- pass 1: streaming over target nodes
  bool fuse_colsum = false;
  {
    CombineBwdParams c{};
    c.rowptr = rowptr; c.col = col; c.val_lin = val_lin; c.n_rows = desc->n_dst;
    c.weightings = weightings; c.grad_out = grad_out; c.saved = saved; c.saved_arg = saved_arg;
    c.d_weightings = d_weightings; c.tstreams = tstreams; c.d_bases = d_bases;
    c.n_saved = n_saved_slots(*desc); c.n_arg = n_arg;
    c.n_ts = L.n_ts; c.ts_sym = L.ts_sym; c.ts_lin = L.ts_lin; c.ts_sq = L.ts_sq;
    c.H = desc->heads; c.B = desc->bases; c.D = desc->dim; c.A = desc->n_aggr;
    c.BD = bd; c.HD = hd; c.AB = desc->n_aggr * desc->bases; c.HAB = hab;
    int na = 0;
    for (int a = 0; a < EGC_MAX_AGGR; ++a) {
      c.aggr[a] = a < desc->n_aggr ? desc->aggr[a] : -1;
      c.arg_slot[a] = (a < desc->n_aggr && (desc->aggr[a] == EGC_AGGR_MAX || desc->aggr[a] == EGC_AGGR_MIN)) ? na++ : -1;
    }
    c.sigmoid = desc->sigmoid;
    auto a4 = [](int v) { return (v + 3) & ~3; };
    int off = 0;
    c.sm_w = off; off = a4(off + hab);
    c.sm_g = off; off = a4(off + hd);
    c.sm_saved = off; off = a4(off + c.n_saved * bd);
    c.sm_arg = off; off = a4(off + n_arg * bd);
    c.sm_per_warp = off;
    const int smem = (off * kAggWarps + 3 * hab) * static_cast<int>(sizeof(float));    // staging areas + index tables
    EGC_REQUIRE(smem <= 200 * 1024, "egc_aggregate_bwd: layer too wide for the shared-memory staging (%d bytes)", smem);
    c.ts_row_stride = stream_major ? bd : static_cast<int64_t>(L.n_ts) * bd;
    c.ts_stream_stride = stream_major ? static_cast<int64_t>(desc->n_dst) * bd : bd;
    c.ts_slab_w = bd;
    c.ts_slab_stride = 0;
    if (slab_w > 0) {
      c.ts_row_stride = static_cast<int64_t>(L.n_ts) * slab_w;
      c.ts_stream_stride = slab_w;
      c.ts_slab_w = slab_w;
      c.ts_slab_stride = static_cast<int64_t>(desc->n_dst) * L.n_ts * slab_w;
    }
    c.vec16 = (hab % 4 == 0 && hd % 4 == 0 && bd % 4 == 0 && aligned16(weightings) && aligned16(grad_out) &&
               aligned16(saved) && aligned16(saved_arg)) ? 1 : 0;
    const bool ev4 = (desc->dim % 4 == 0) && aligned16(tstreams);
    const bool linw = val_lin != nullptr;
    const int grid = combine_bwd_grid(desc->n_dst);
    // column sums fused into pass 1 when the per-lane accumulators cover the row widths
    fuse_colsum = (d_bias != nullptr || d_lin_colsum != nullptr) && hab <= 32 * kCbColIt &&
                  hd <= 32 * kCbColIt * (ev4 ? 4 : 1);
    c.colsum_part = fuse_colsum ? static_cast<float*>(colsum_ws) : nullptr;
    c.skip_route = (flags & EGC_BWD_SKIP_ROUTING) ? 1 : 0;
    c.t_route = t_route;
#define EGC_LAUNCH_COMBINE_CFG(CFG, EV, LW)                                                                    \
    {                                                                                                          \
      auto kern = k_combine_bwd<CFG, EV, LW>;                                                                    \
      if (smem > 48 * 1024) EGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      LaunchScope egc_ls_("k_combine_bwd", st);                                                                \
      kern<<<grid, kAggThreads, smem, st>>>(c);                                                                \
    }
    using Dyn = DynCfg<0, 32>;
    const int static_idx = (ev4 && !linw && c.vec16) ? static_cfg_index(*desc) : -1;
    bool launched = false;
#define X(I, ...)                                                                                              \
    if (!launched && static_idx == I) {                                                                        \
      using SC = StaticCfg<__VA_ARGS__>;                                                                       \
      EGC_REQUIRE(SC::bsm_per_warp == c.sm_per_warp && SC::bsm_saved == c.sm_saved && SC::bsm_arg == c.sm_arg && \
                  SC::ts_sym == c.ts_sym && SC::ts_lin == c.ts_lin && SC::ts_sq == c.ts_sq,                    \
                  "egc_aggregate_bwd: static configuration %d disagrees with the host layout", I);             \
      EGC_LAUNCH_COMBINE_CFG(SC, 4, false)                                                                     \
      launched = true;                                                                                         \
    }
    EGC_STATIC_CFGS(X)
#undef X
    if (launched) {
    } else if (ev4 && !linw) EGC_LAUNCH_COMBINE_CFG(Dyn, 4, false)
    else if (ev4 && linw) EGC_LAUNCH_COMBINE_CFG(Dyn, 4, true)
    else if (!linw) EGC_LAUNCH_COMBINE_CFG(Dyn, 1, false)
    else EGC_LAUNCH_COMBINE_CFG(Dyn, 1, true)
#undef EGC_LAUNCH_COMBINE_CFG
    EGC_LAUNCH_CHECK("k_combine_bwd");
    if (fuse_colsum) {
      {
        LaunchScope egc_ls_("k_colsum_partials", st);
        k_colsum_partials<<<ceil_div(hd + hab, 8), 256, 0, st>>>(c.colsum_part, grid, hd, hab, d_bias, d_lin_colsum);
      }
      EGC_LAUNCH_CHECK("k_colsum_partials");
    }
  }

  // ---- min/max routing (feature-slab-major atomics)
  if (L.has_route && !(flags & EGC_BWD_SKIP_ROUTING)) {
    RouteParams r{};
    r.saved_arg = saved_arg; r.t_route = t_route; r.col = col; r.val_lin = val_lin; r.d_bases = d_bases;
    r.n_rows = desc->n_dst; r.n_arg = n_arg; r.BD = bd; r.n_slabs = ceil_div(bd, 32);
    // hubs = the long columns of the CSC plan (more than EGC_CHUNK_EDGES entries), privatised per CTA
    r.n_hubs = (csc_plan != nullptr && !(flags & EGC_BWD_NO_HUB_PRIVATISATION)) ? std::min(csc_plan->n_long, kRouteMaxHubs) : 0;
    r.hubs = r.n_hubs > 0 ? csc_plan->long_rows : nullptr;
    const int row_groups = ceil_div(desc->n_dst, kRouteRows);
    const int grid = std::max(1, std::min(ceil_div(row_groups, kRouteThreads / 32), sm_count()));
    const int smem = r.n_hubs > 0 ? 2 * kRouteHashSize * static_cast<int>(sizeof(int)) + r.n_hubs * 32 * static_cast<int>(sizeof(float)) : 0;
    if (smem > 48 * 1024) EGC_CUDA(cudaFuncSetAttribute(k_route_minmax, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    {
      LaunchScope egc_ls_("k_route_minmax", st);
      k_route_minmax<<<grid, kRouteThreads, smem, st>>>(r);
    }
    EGC_LAUNCH_CHECK("k_route_minmax");
  }

  // ---- pass 2: per source column (CSC), atomic-free
  if (L.tsmask != 0) {
    AggParams geo{};
    fill_agg_params(geo, *desc, vec4);
    ScatterParams s{};
    s.colptr = colptr; s.rowidx = rowidx; s.val_sym = csc_val_sym; s.val_lin = csc_val_lin; s.n_cols = desc->n_src;
    s.n_long = csc_plan ? csc_plan->n_long : 0;
    s.n_chunks = csc_plan ? csc_plan->n_chunks : 0;
    s.long_rows = csc_plan ? csc_plan->long_rows : nullptr;
    s.long_chunk_ptr = csc_plan ? csc_plan->long_chunk_ptr : nullptr;
    s.chunk_row = csc_plan ? csc_plan->chunk_row : nullptr;
    s.chunk_begin = csc_plan ? csc_plan->chunk_begin : nullptr;
    s.partials = csc_part;
    const bool fuse_merge = (geo.n_pass == 1 || slab_w > 0) && s.n_long > 0;
    s.long_counter = fuse_merge ? reinterpret_cast<int*>(reinterpret_cast<char*>(csc_part) +
                                                        align_up(static_cast<size_t>(s.n_chunks) * std::max(L.n_ts, 1) * bd * 4, 256))
                                : nullptr;
    // column-block kernel: 128-bit pieces, unweighted entries, one pass, interleaved streams
    // (EGC_BWD_WARP_PER_COLUMN=1 keeps the warp-per-column kernel)
    static const bool legacy_cols = getenv("EGC_BWD_WARP_PER_COLUMN") != nullptr;
    const bool col_blocks = vec4 && val_lin == nullptr && !stream_major && slab_w == 0 && geo.n_pass == 1 &&
                            (geo.G == 32 || geo.G == 16) && !legacy_cols;
    const size_t counters_bytes = align_up(static_cast<size_t>(s.n_long) * kSlabMaxSlabs * sizeof(int), 256);
    int* task_counter = reinterpret_cast<int*>(reinterpret_cast<char*>(csc_part) +
                                               align_up(static_cast<size_t>(s.n_chunks) * std::max(L.n_ts, 1) * bd * 4, 256) + counters_bytes);
    if (col_blocks && s.long_counter == nullptr)
      s.long_counter = reinterpret_cast<int*>(reinterpret_cast<char*>(task_counter) - counters_bytes);
    if (col_blocks) EGC_CUDA(cudaMemsetAsync(s.long_counter, 0, counters_bytes + sizeof(int), st));
    else if (fuse_merge) EGC_CUDA(cudaMemsetAsync(s.long_counter, 0, static_cast<size_t>(s.n_long) * kSlabMaxSlabs * sizeof(int), st));
    s.tstreams = tstreams; s.bases = bases; s.d_bases = d_bases;
    s.n_ts = L.n_ts; s.ts_sym = L.ts_sym; s.ts_lin = L.ts_lin; s.ts_sq = L.ts_sq;
    s.BD = bd; s.nvec = geo.nvec; s.G = geo.G; s.n_pass = geo.n_pass;
    const int64_t table = static_cast<int64_t>(desc->n_dst) * bd;
    s.ts_row_stride = stream_major ? bd : static_cast<int64_t>(L.n_ts) * bd;
    s.off_sym = L.ts_sym < 0 ? 0 : (stream_major ? L.ts_sym * table : static_cast<int64_t>(L.ts_sym) * bd);
    s.off_lin = L.ts_lin < 0 ? 0 : (stream_major ? L.ts_lin * table : static_cast<int64_t>(L.ts_lin) * bd);
    s.off_sq = L.ts_sq < 0 ? 0 : (stream_major ? L.ts_sq * table : static_cast<int64_t>(L.ts_sq) * bd);
    bool accumulate = L.has_route;
    if (slab_w > 0) {
      s.slab_w = slab_w;
      s.n_slabs = bd / slab_w;
      s.slab_stride = static_cast<int64_t>(desc->n_dst) * L.n_ts * slab_w;
      s.routed = accumulate ? 1 : 0;
      s.mode = 0;
      if (int rc = slab_w == 32 ? launch_scatter_slab<32>(s, L.tsmask, st) : launch_scatter_slab<16>(s, L.tsmask, st)) return rc;
      EGC_LAUNCH_CHECK("k_scatter_slab");
    } else if (col_blocks) {
      s.routed = accumulate ? 1 : 0;
      s.mode = 0;
      if (int rc = geo.G == 32 ? launch_scatter_cols<32>(s, L.tsmask, task_counter, st) : launch_scatter_cols<16>(s, L.tsmask, task_counter, st)) return rc;
    } else
    for (int bit = 1; bit <= 4; bit <<= 1) {
      const int sweep_mask = stream_major ? (L.tsmask & bit) : (bit == 1 ? L.tsmask : 0);
      if (sweep_mask == 0) continue;
      s.routed = accumulate ? 1 : 0;
      s.mode = 0;
      if (int rc = launch_scatter(s, sweep_mask, vec4, val_lin != nullptr, st)) return rc;
      if (s.n_long > 0 && !fuse_merge) {
        s.mode = 1;
        if (int rc = launch_scatter(s, sweep_mask, vec4, val_lin != nullptr, st)) return rc;
      }
      accumulate = true;
    }
  }

  if (!fuse_colsum) {
    if (d_bias != nullptr) {
      if (int rc = colsum_f32(grad_out, desc->n_dst, hd, d_bias, colsum_ws, L.colsum_bytes, st)) return rc;
    }
    if (d_lin_colsum != nullptr) {
      if (int rc = colsum_f32(d_weightings, desc->n_dst, hab, d_lin_colsum, colsum_ws, L.colsum_bytes, st)) return rc;
    }
  }
  return EGC_OK;
}

}  // extern "C"
