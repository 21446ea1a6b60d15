// Kernel template of the fused aggregation (+ combination forward, + target-side backward pass).
// Included by aggregate_{fwd,bwd}_v{4,1}.cu, each of which instantiates one (VEC, BWD) family.
#pragma once

#include "aggregate.cuh"

namespace egc {

// ---------------------------------------------------------------------------------------------
// forward epilogue: out[row, h*D+d] = sum_ab w[h*AB+ab] * agg[ab*D+d] (+ bias)     (ref :195-208)
// agg rows live in this warp's shared memory as [A][BD] == [A*B][D].
// ---------------------------------------------------------------------------------------------
template <int EV>
__device__ __forceinline__ void combine_epilogue(const AggParams& p, const float* sm, int row, int lane) {
  const float* agg = sm + p.sm_agg;
  const float* w = sm + p.sm_w;
  float* out = p.out + static_cast<int64_t>(row) * p.HD;
  for (int o0 = lane * EV; o0 < p.HD; o0 += 32 * EV) {
    const int h = o0 / p.D, d = o0 - h * p.D;
    float acc[EV];
#pragma unroll
    for (int k = 0; k < EV; ++k) acc[k] = 0.f;
    const float* wh = w + h * p.AB;
    for (int ab = 0; ab < p.AB; ++ab) {
      const float wv = wh[ab];
      float a[EV];
      ld_plain<EV>(a, agg + ab * p.D + d);
#pragma unroll
      for (int k = 0; k < EV; ++k) acc[k] = fmaf(wv, a[k], acc[k]);
    }
    if (p.bias != nullptr) {
#pragma unroll
      for (int k = 0; k < EV; ++k) acc[k] += __ldg(p.bias + o0 + k);
    }
    st_row<EV>(out + o0, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// backward epilogue (per target row):
//   d_w[h,ab]   = sum_d g[h*D+d] * agg[ab*D+d]                       (x sigmoid' when requested)
//   d_agg[a][p] = sum_h w[h*AB + a*B + b(p)] * g[h*D + d(p)]
// and from d_agg the target-side streams t_sym / t_lin / t_sq plus min/max routing.
// ---------------------------------------------------------------------------------------------
template <int EV, bool LINW>
__device__ __forceinline__ void backward_epilogue(const AggParams& p, const float* sm, int row, int lane, float cntf) {
  const float* agg = sm + p.sm_agg;
  const float* w = sm + p.sm_w;
  const float* g = sm + p.sm_g;
  const int D = p.D;
  // (1) gradient of the combination weights
  for (int t = lane; t < p.HAB; t += 32) {
    const int h = t / p.AB, ab = t - h * p.AB;
    const float* gh = g + h * D;
    const float* aa = agg + ab * D;
    float dot = 0.f;
    int dd = lane % D;                       // skewed start: spreads lanes over the banks
    for (int i = 0; i < D; ++i) {
      dot = fmaf(gh[dd], aa[dd], dot);
      dd = (dd + 1 == D) ? 0 : dd + 1;
    }
    if (p.sigmoid) { const float s = w[t]; dot *= s * (1.f - s); }
    p.d_weightings[static_cast<int64_t>(row) * p.HAB + t] = dot;
  }
  // (2) gradient w.r.t. the aggregates -> target-side streams
  float* ts = p.tstreams + static_cast<int64_t>(row) * p.n_ts * p.BD;
  const int* amx = reinterpret_cast<const int*>(sm + p.sm_amx);
  const int* amn = reinterpret_cast<const int*>(sm + p.sm_amn);
  for (int p0 = lane * EV; p0 < p.BD; p0 += 32 * EV) {
    const int b = p0 / D, d = p0 - b * D;
    float t_sym[EV], t_lin[EV], t_sq[EV];
#pragma unroll
    for (int k = 0; k < EV; ++k) { t_sym[k] = 0.f; t_lin[k] = 0.f; t_sq[k] = 0.f; }
    for (int a = 0; a < p.A; ++a) {
      float da[EV];
#pragma unroll
      for (int k = 0; k < EV; ++k) da[k] = 0.f;
      const float* wa = w + a * p.B + b;
      for (int h = 0; h < p.H; ++h) {
        const float wv = wa[h * p.AB];
        float gv[EV];
        ld_plain<EV>(gv, g + h * D + d);
#pragma unroll
        for (int k = 0; k < EV; ++k) da[k] = fmaf(wv, gv[k], da[k]);
      }
      const int code = p.aggr[a];
#pragma unroll
      for (int k = 0; k < EV; ++k) {
        const int pp = p0 + k;
        switch (code) {
          case EGC_AGGR_SUM: t_lin[k] += da[k]; break;
          case EGC_AGGR_MEAN: t_lin[k] += __fdiv_rn(da[k], cntf); break;
          case EGC_AGGR_SYMNORM: t_sym[k] += da[k]; break;
          case EGC_AGGR_MAX:
          case EGC_AGGR_MIN: {
            const int arg = (code == EGC_AGGR_MAX) ? amx[pp] : amn[pp];
            if (arg >= 0) {
              float v = da[k];
              if (LINW) v *= __ldg(p.val_lin + arg);
              atomicAdd(p.d_bases + static_cast<int64_t>(__ldg(p.col + arg)) * p.BD + pp, v);
            }
            break;
          }
          case EGC_AGGR_VAR:
          case EGC_AGGR_STD: {
            float dv = da[k];
            if (code == EGC_AGGR_STD) {
              const float var = sm[p.sm_var + pp];
              const float sd = agg[a * p.BD + pp];
              dv = var > 0.f ? dv / (2.f * sd) : 0.f;       // relu gate, d sqrt
            }
            const float q = __fdiv_rn(dv, cntf);
            t_sq[k] += q;
            t_lin[k] -= 2.f * sm[p.sm_mean + pp] * q;
            break;
          }
        }
      }
    }
    if (p.ts_sym >= 0) st_row<EV>(ts + p.ts_sym * p.BD + p0, t_sym);
    if (p.ts_lin >= 0) st_row<EV>(ts + p.ts_lin * p.BD + p0, t_lin);
    if (p.ts_sq >= 0) st_row<EV>(ts + p.ts_sq * p.BD + p0, t_sq);
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int MASK, int VEC, bool LINW, bool BWD>
__global__ void __launch_bounds__(kAggThreads) k_aggregate(const __grid_constant__ AggParams p) {
  extern __shared__ __align__(16) float smem_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kAggWarps + warp;
  float* sm = smem_all + warp * p.sm_per_warp;

  int row, begin, end, chunk_id = -1, long_idx = -1;
  if (p.mode == 0) {
    if (gw < p.n_chunks) {
      chunk_id = gw;
      row = p.chunk_row[gw];
      begin = p.chunk_begin[gw];
      end = min(begin + EGC_CHUNK_EDGES, p.rowptr[row + 1]);
    } else {
      row = gw - p.n_chunks;
      if (row >= p.n_rows) return;
      begin = p.rowptr[row];
      end = p.rowptr[row + 1];
      if (end - begin > EGC_CHUNK_EDGES) return;      // long row: chunks + merge task do it
    }
  } else {
    long_idx = gw;
    if (long_idx >= p.n_long) return;
    row = p.long_rows[long_idx];
    begin = p.rowptr[row];
    end = p.rowptr[row + 1];
  }
  const bool is_chunk = chunk_id >= 0;
  const float cntf = static_cast<float>(max(end - begin, 1));   // mean divides by the nnz count, min 1

  if (!is_chunk) {   // stage this row's combination weights (and upstream gradient) asynchronously
    const float* wsrc = p.weightings + static_cast<int64_t>(row) * p.HAB;
    for (int t = lane; t < p.HAB; t += 32) cp_async_4(sm + p.sm_w + t, wsrc + t);
    if constexpr (BWD) {
      const float* gsrc = p.grad_out + static_cast<int64_t>(row) * p.HD;
      for (int t = lane; t < p.HD; t += 32) cp_async_4(sm + p.sm_g + t, gsrc + t);
    }
  }

  const int G = p.G;
  for (int pass = 0; pass < p.n_pass; ++pass) {
    const int piece = pass * 32 + (lane & (G - 1));
    const bool active = piece < p.nvec;
    const int foff = piece * VEC;
    Acc<MASK, VEC, LINW> acc;
    acc.init();
    if (p.mode == 0) {
      accumulate_range<MASK, VEC, LINW>(acc, p, begin, end, lane, foff, active);
    } else if (active) {
      const int c0 = p.long_chunk_ptr[long_idx], c1 = p.long_chunk_ptr[long_idx + 1];
      for (int c = c0; c < c1; ++c)
        acc.merge_from(p.partials + (static_cast<int64_t>(c) * p.n_slots) * p.BD + foff, p.BD);
    }
    const bool writer = active && lane < G;    // after the xor merge every group holds the full result
    if (is_chunk) {
      if (writer) acc.store(p.partials + (static_cast<int64_t>(chunk_id) * p.n_slots) * p.BD + foff, p.BD);
      continue;
    }
    if (writer) {
      for (int a = 0; a < p.A; ++a) {
        const int code = p.aggr[a];
        float v[VEC];
        float mean[VEC], var[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          mean[k] = 0.f; var[k] = 0.f;
          v[k] = finalize_one<MASK, VEC, LINW>(acc, code, k, cntf, mean[k], var[k]);
        }
        st_row<VEC>(sm + p.sm_agg + a * p.BD + foff, v);
        if constexpr (!BWD) {
          if (p.agg_out != nullptr)
            st_row<VEC>(p.agg_out + (static_cast<int64_t>(row) * p.A + a) * p.BD + foff, v);
          if (p.arg_out != nullptr) {
            int32_t* ao = p.arg_out + (static_cast<int64_t>(row) * p.A + a) * p.BD + foff;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              int arg = -1;
              if constexpr (MASK & P_MAX) { if (code == EGC_AGGR_MAX) arg = acc.amx[k]; }
              if constexpr (MASK & P_MIN) { if (code == EGC_AGGR_MIN) arg = acc.amn[k]; }
              ao[k] = arg;
            }
          }
        } else {
          if constexpr (MASK & P_SQ) {
            if (code == EGC_AGGR_VAR || code == EGC_AGGR_STD) {
              st_row<VEC>(sm + p.sm_mean + foff, mean);
              st_row<VEC>(sm + p.sm_var + foff, var);
            }
          }
        }
      }
      if constexpr (BWD) {
        if constexpr (MASK & P_MAX) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) reinterpret_cast<int*>(sm + p.sm_amx)[foff + k] = acc.amx[k];
        }
        if constexpr (MASK & P_MIN) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) reinterpret_cast<int*>(sm + p.sm_amn)[foff + k] = acc.amn[k];
        }
      }
    }
  }
  if (is_chunk) return;

  cp_async_wait_all();
  __syncwarp();
  const bool ev4 = (p.D % 4 == 0);
  if constexpr (!BWD) {
    if (p.out != nullptr) {
      if (ev4) combine_epilogue<4>(p, sm, row, lane);
      else combine_epilogue<1>(p, sm, row, lane);
    }
  } else {
    if (ev4) backward_epilogue<4, LINW>(p, sm, row, lane, cntf);
    else backward_epilogue<1, LINW>(p, sm, row, lane, cntf);
  }
}

// ---------------------------------------------------------------------------------------------
// dispatch over the primitive mask
// ---------------------------------------------------------------------------------------------
template <int MASK, int VEC, bool LINW, bool BWD>
int launch_one(const AggParams& p, int smem_bytes, cudaStream_t st) {
  auto kern = k_aggregate<MASK, VEC, LINW, BWD>;
  if (smem_bytes > 48 * 1024) {
    EGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  }
  const int64_t tasks = p.mode == 0 ? static_cast<int64_t>(p.n_chunks) + p.n_rows : p.n_long;
  if (tasks <= 0) return EGC_OK;
  const int grid = ceil_div(tasks, kAggWarps);
  {
    LaunchScope egc_ls_(BWD ? (p.mode ? "k_aggregate_bwd_merge" : "k_aggregate_bwd") : (p.mode ? "k_aggregate_fwd_merge" : "k_aggregate_fwd"), st);
    kern<<<grid, kAggThreads, smem_bytes, st>>>(p);
  }
  EGC_LAUNCH_CHECK("k_aggregate");
  return EGC_OK;
}

// valid masks: any non-empty subset of {SUM, SYM, SQ, MAX, MIN} where SQ implies SUM
#define EGC_MASK_CASES(X) \
  X(1) X(2) X(3) X(5) X(7) X(8) X(9) X(10) X(11) X(13) X(15) X(16) X(17) X(18) X(19) X(21) X(23) \
  X(24) X(25) X(26) X(27) X(29) X(31)
// masks without SYM: the only ones that can carry per-nnz linear weights
#define EGC_MASK_CASES_NOSYM(X) X(1) X(5) X(8) X(9) X(13) X(16) X(17) X(21) X(24) X(25) X(29)

template <int VEC, bool BWD>
int launch_family(const AggParams& p, int mask, bool linw, int smem_bytes, cudaStream_t st) {
  if (!linw) {
    switch (mask) {
#define X(M) case M: return launch_one<M, VEC, false, BWD>(p, smem_bytes, st);
      EGC_MASK_CASES(X)
#undef X
    }
  } else {
    switch (mask) {
#define X(M) case M: return launch_one<M, VEC, true, BWD>(p, smem_bytes, st);
      EGC_MASK_CASES_NOSYM(X)
#undef X
    }
  }
  set_error("aggregate: unsupported primitive mask %d (linw=%d)", mask, static_cast<int>(linw));
  return EGC_ERR_UNSUPPORTED;
}

}  // namespace egc
