// tcgen05 parameter-gradient GEMM for sm_100a:  C[F_in, BD + HAB] = x^T . [d_bases | d_lin]
// (dW_b = x^T d_bases, dW_c^T = x^T d_lin; ref: autograd of optimized_layers.py:180-182).
//
// The contraction runs over the NODE dimension, so both operands stream and both are feature-contiguous
// in memory (MN-major).  kind::tf32 MMAs with no-swizzle descriptors only accept K-major operands on
// this part (tools/umma_probe.cu: the MN-major bits yield zeros), so the producers transpose on the
// fly: lane -> (node%4, feature%8) reads 4 x 32 B global sectors per warp instruction and writes 128
// contiguous shared bytes (conflict-free) of the canonical K-major layout (8 rows x 16 B core matrices).
// Each persistent CTA owns a contiguous range of 16-node chunks, accumulates one 128 x N_pad fp32
// tile in TMEM (3-term hi/lo split) and writes its partial tile to the workspace; a small
// deterministic kernel reduces the partials into dW_b / dW_c (transposing the latter).
#include <algorithm>

#include "project.cuh"
#include "tc_common.cuh"

namespace egc {

constexpr int kWgThreads = 288;          // warps 0-3 epilogue, 4 MMA, 5-8 producers
constexpr int kWgChunk = 16;             // nodes per chunk (2 UMMA k-steps)
constexpr int kWgM = 128;                // feature rows of the accumulator tile (F_in padded)
constexpr int kWgMaxSmem = 227 * 1024;
constexpr int kMaxSlots = 40;            // register-prefetched scalars per producer thread and chunk

struct WgParams {
  const float* x; int f_in;              // A^T source: x[n, f_in]
  const float* d1; int n1;               // B source, columns [0, n1)      : d_bases[n, n1]
  const float* d2; int n2;               // B source, columns [n1, n1+n2)  : d_lin[n, n2]
  int n_nodes;
  int n_pad;                             // multiple of 16
  int n_terms;
  int stages;
  int chunks_total, chunks_per_cta;
  float* partial;                        // [grid][kWgM][n_pad]
};

__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_tc(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_half = kWgM * kWgChunk * 4;                       // 8 KB: hi (or lo) of an A chunk
  const uint32_t b_half = static_cast<uint32_t>(p.n_pad) * kWgChunk * 4;
  const uint32_t stage_bytes = 2 * (a_half + b_half);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  // bars: [0,S) full, [S,2S) empty, [2S] accumulator done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (p.stages + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * p.stages);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 128); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int c_begin = blockIdx.x * p.chunks_per_cta;
  const int c_end = min(c_begin + p.chunks_per_cta, p.chunks_total);
  const int n_chunks = max(c_end - c_begin, 0);
  const uint32_t a_lbo = kWgM * 16, b_lbo = static_cast<uint32_t>(p.n_pad) * 16;   // bytes between 4-node K pieces
  constexpr int kASlots = 4 * (kWgM / 8);                   // (K piece, 8-feature group) pairs of A per chunk
  const int b_groups = p.n_pad / 8;
  const int n_slots = kASlots + 4 * b_groups;               // <= kMaxSlots * 4

  if (warp >= 5) {
    // ================= producers: transposing stage-in =================
    const int pw = warp - 5;
    const int kk = lane & 3, mm = lane >> 2;                // node within the 4-piece, feature within the 8-group
    int stage = 0;
    uint32_t phase = 0;
    for (int c = c_begin; c < c_end; ++c) {
      const int node0 = c * kWgChunk;
      float v[kMaxSlots];
#pragma unroll
      for (int i = 0; i < kMaxSlots; ++i) {
        const int slot = pw + 4 * i;
        float t = 0.f;
        if (slot < kASlots) {
          const int piece = slot / (kWgM / 8), grp = slot - piece * (kWgM / 8);
          const int node = node0 + piece * 4 + kk, f = grp * 8 + mm;
          if (node < p.n_nodes && f < p.f_in) t = __ldcs(p.x + static_cast<int64_t>(node) * p.f_in + f);
        } else if (slot < n_slots) {
          const int s2 = slot - kASlots;
          const int piece = s2 / b_groups, grp = s2 - piece * b_groups;
          const int node = node0 + piece * 4 + kk, col = grp * 8 + mm;
          if (node < p.n_nodes) {
            if (col < p.n1) t = __ldcs(p.d1 + static_cast<int64_t>(node) * p.n1 + col);
            else if (col < p.n1 + p.n2) t = __ldcs(p.d2 + static_cast<int64_t>(node) * p.n2 + (col - p.n1));
          }
        }
        v[i] = t;
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      uint8_t* st_base = smem + static_cast<size_t>(stage) * stage_bytes;
#pragma unroll
      for (int i = 0; i < kMaxSlots; ++i) {
        const int slot = pw + 4 * i;
        const float h = tf32_hi(v[i]), l = v[i] - h;
        if (slot < kASlots) {
          const int piece = slot / (kWgM / 8), grp = slot - piece * (kWgM / 8);
          const uint32_t off = piece * a_lbo + (grp * 8 + mm) * 16 + kk * 4;
          *reinterpret_cast<float*>(st_base + off) = h;
          *reinterpret_cast<float*>(st_base + a_half + off) = l;
        } else if (slot < n_slots) {
          const int s2 = slot - kASlots;
          const int piece = s2 / b_groups, grp = s2 - piece * b_groups;
          const uint32_t off = 2 * a_half + piece * b_lbo + (grp * 8 + mm) * 16 + kk * 4;
          *reinterpret_cast<float*>(st_base + off) = h;
          *reinterpret_cast<float*>(st_base + b_half + off) = l;
        }
      }
      fence_proxy_async();
      mbar_arrive(full_bar(stage));
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(p.n_pad >> 3) << 17) |
                             (static_cast<uint32_t>(kWgM >> 4) << 24);
      const uint32_t base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_hi = base + stage * stage_bytes, a_lo = a_hi + a_half;
        const uint32_t b_hi = a_hi + 2 * a_half, b_lo = b_hi + b_half;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const uint64_t da_hi = make_desc(a_hi + s * 2 * a_lbo, a_lbo, 128);
          const uint64_t db_hi = make_desc(b_hi + s * 2 * b_lbo, b_lbo, 128);
          umma_tf32(tmem_base, da_hi, db_hi, idesc, (c | s) != 0 ? 1u : 0u);
          if (p.n_terms == 3) {
            const uint64_t da_lo = make_desc(a_lo + s * 2 * a_lbo, a_lbo, 128);
            const uint64_t db_lo = make_desc(b_lo + s * 2 * b_lbo, b_lbo, 128);
            umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
            umma_tf32(tmem_base, da_lo, db_hi, idesc, 1u);
          }
        }
        umma_commit(empty_bar(stage));
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    // ================= epilogue: this CTA's partial tile -> workspace =================
    float* dst = p.partial + (static_cast<int64_t>(blockIdx.x) * kWgM + tid) * p.n_pad;
    if (n_chunks > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
      for (int col0 = 0; col0 < p.n_pad; col0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + lane_base + static_cast<uint32_t>(col0), r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(dst + col0 + 4 * q) =
              make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                          __uint_as_float(r[4 * q + 3]));
      }
    } else {
      for (int col = 0; col < p.n_pad; col += 4) *reinterpret_cast<float4*>(dst + col) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
}

// dW_b[m][n] = sum_cta partial[cta][m][n] (n < n1);  dW_c[n - n1][m] = sum_cta partial[cta][m][n] (n >= n1)
__global__ void k_wgrad_reduce(const float* __restrict__ partial, int n_cta, int f_in, int n1, int n2, int n_pad,
                               float* __restrict__ d_w_bases, float* __restrict__ d_w_comb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = n1 + n2;
  if (idx >= f_in * N) return;
  const int m = idx / N, n = idx - m * N;
  const float* src = partial + static_cast<int64_t>(m) * n_pad + n;
  const int64_t stride = static_cast<int64_t>(kWgM) * n_pad;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  int c = 0;
  for (; c + 4 <= n_cta; c += 4) {
    t0 += src[c * stride]; t1 += src[(c + 1) * stride]; t2 += src[(c + 2) * stride]; t3 += src[(c + 3) * stride];
  }
  for (; c < n_cta; ++c) t0 += src[c * stride];
  const float t = (t0 + t1) + (t2 + t3);
  if (n < n1) { if (d_w_bases != nullptr) d_w_bases[static_cast<int64_t>(m) * n1 + n] = t; }
  else if (d_w_comb != nullptr) d_w_comb[static_cast<int64_t>(n - n1) * f_in + m] = t;
}

static int round16w(int v) { return (v + 15) / 16 * 16; }

static int wgrad_stages(int n_pad) {
  const size_t stage = 2 * (static_cast<size_t>(kWgM) * kWgChunk * 4 + static_cast<size_t>(n_pad) * kWgChunk * 4);
  const int s = static_cast<int>((kWgMaxSmem - 512) / stage);
  return std::min(s, 6);
}

bool wgrad_tc_supported(int n, int f_in, int bd, int hab) {
  if (n < 1 || f_in % 4 || bd % 4 || hab % 4 || f_in > kWgM) return false;
  const int n_pad = round16w(bd + hab);
  const int slots = 4 * (kWgM / 8) + 4 * (n_pad / 8);      // producer register prefetch: kMaxSlots scalars per thread
  return n_pad <= 256 && wgrad_stages(n_pad) >= 2 && (slots + 3) / 4 <= kMaxSlots;
}

size_t wgrad_tc_workspace(int n, int f_in, int bd, int hab) {
  (void)n; (void)f_in;
  return static_cast<size_t>(sm_count()) * kWgM * round16w(bd + hab) * sizeof(float) + 256;
}

int wgrad_tc(const float* x, const float* d_bases, const float* d_lin, int n, int f_in, int bd, int hab,
             float* d_w_bases, float* d_w_comb, int n_terms, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  EGC_REQUIRE(workspace_bytes >= wgrad_tc_workspace(n, f_in, bd, hab), "wgrad_tc: workspace too small");
  WgParams p{};
  p.x = x; p.f_in = f_in; p.d1 = d_bases; p.n1 = bd; p.d2 = d_lin; p.n2 = hab; p.n_nodes = n;
  p.n_pad = round16w(bd + hab);
  p.n_terms = n_terms;
  p.stages = wgrad_stages(p.n_pad);
  p.chunks_total = ceil_div(n, kWgChunk);
  const int grid = std::min(sm_count(), p.chunks_total);
  p.chunks_per_cta = ceil_div(p.chunks_total, grid);
  p.partial = static_cast<float*>(workspace);
  const size_t stage = 2 * (static_cast<size_t>(kWgM) * kWgChunk * 4 + static_cast<size_t>(p.n_pad) * kWgChunk * 4);
  const size_t smem = p.stages * stage + 512;
  static bool attr_set = false;
  if (!attr_set) {
    EGC_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgMaxSmem));
    attr_set = true;
  }
  {
    LaunchScope ls("k_wgrad_tc", st);
    k_wgrad_tc<<<grid, kWgThreads, smem, st>>>(p);
  }
  EGC_LAUNCH_CHECK("k_wgrad_tc");
  const int total = f_in * (bd + hab);
  {
    LaunchScope ls("k_wgrad_reduce", st);
    k_wgrad_reduce<<<ceil_div(total, 256), 256, 0, st>>>(p.partial, grid, f_in, bd, hab, p.n_pad, d_w_bases, d_w_comb);
  }
  EGC_LAUNCH_CHECK("k_wgrad_reduce");
  return EGC_OK;
}

}  // namespace egc
