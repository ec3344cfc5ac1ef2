// 3x3 (pad 1) / 1x1 convolution as implicit GEMM on tcgen05 tensor cores, NHWC bf16 in, fp32 accumulate in TMEM.
//
// Replaces the cuDNN conv2d call sites K1/K2 of SURVEY.md 2b (reference torchlayers.py:18, models/unet.py:25-29,
// models/phiseg.py:28,32,57-58,91) and, with transposed/flipped packed weights, their dgrad.
//
// GEMM view: M = 128 output pixels (a TN x TH x TW box of the [N,H,W] pixel grid), N = Cout (<= 256), K = taps * Cin.
// One K step = (tap, block of KC input channels).  For every K step the producer issues
//   * one 4-D TMA box load of the input shifted by the tap offset -- out-of-bounds coordinates (the zero padding, and
//     batch overhang of the last tile) are zero-filled by the TMA unit, so no halo logic exists in the kernel;
//   * one 3-D TMA box load of the [Cout x KC] weight slice of that tap.
// Both land in the canonical K-major swizzled layout (swizzle span = KC * 2 bytes) that tcgen05.mma reads directly.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner, warps 2..5 = epilogue.
// Epilogue: TMEM -> registers (tcgen05.ld) -> per-channel affine (+ReLU) -> bf16 -> shared staging -> coalesced 16 B
// stores; optionally per-tile per-channel sum / sum-of-squares of the STORED bf16 values (BatchNorm batch statistics,
// reduced deterministically by uz_bn_finalize).
#include "common.cuh"
#include "unetzoo_b200.h"

namespace uz {
int g_conv_debug_flags = 0;
int conv2_stats_rows(int N, int H, int W, int Cin, int Cout);
int conv2_launch(const void* x, int N, int D, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, void* y, int ldy,
                 const float* scale, const float* shift, int relu, float* stats_partial, const UzConvExtra* ex, void* stream,
                 int* handled);
// conv_small.cu: CUDA-core kernel for maps of at most 4x4 pixels (latency-bound layers)
struct SmallBn {
  float count, eps, momentum;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  int updates, relu;
  float* scale_out; float* shift_out; float* mean_out; float* invstd_out;
  void* a; int lda;
};
int conv_small_supported(int N, int H, int W, int Cin, int Cout);
int conv_small_launch(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, void* y,
                      int ldy, const float* scale, const float* shift, int relu, const SmallBn* fb, void* stream);
}  // namespace uz

namespace {

constexpr int kBlockM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct ConvParams {
  int N, H, W, Cin, Cout;
  int taps;           // 9 or 1
  int TW, TH, TN;     // pixel box of one tile, TW*TH*TN == 128
  int tilesW, tilesH; // tiles per image row / column
  int KC;             // channels per TMA box (16 / 32 / 64)
  int KB;             // channel blocks per pipeline stage (a stage = KB * KC channels of one tap)
  int BN;             // output channels per CTA (multiple of 16, <= 256)
  int stages;
  int ldy;            // output pixel stride (elements)
  int relu;
  uint32_t tmem_cols;
  __nv_bfloat16* y;
  const float* scale;  // [Cout] or nullptr (=1)
  const float* shift;  // [Cout] or nullptr (=0)
  float* stats;        // [2][Cout] accumulators (zero on entry, atomically added to) or nullptr
  int stats_rows;      // 1: stats is [tiles][2][Cout], row = this CTA's tile, plain stores (fixed summation order)
  // UzConvExtra: fused BatchNorm/ReLU backward statistics of the producer layer, residual (see conv_tc2.cu)
  const __nv_bfloat16* bn_y;
  const float* bn_scale;
  const float* bn_shift;
  float* bn_sums;
  int bn_ldy, bn_relu;
  const __nv_bfloat16* res;
  int ld_res;
  float res_sign;
  // fused training-mode BatchNorm + ReLU (uz_conv_bn_act_fused): the CTAs of all pixel tiles form a thread-block cluster
  // and exchange their per-channel sums through distributed shared memory; y keeps the conv output, a the activation
  int fuse_bn;
  float bn_count, bn_eps, bn_momentum;
  const float* bn_gamma;
  const float* bn_beta;
  float* bn_running_mean;
  float* bn_running_var;
  int bn_updates, bn_act_relu;
  float* bn_scale_out;
  float* bn_shift_out;
  float* bn_mean_out;
  float* bn_invstd_out;
  __nv_bfloat16* a_out;
  int lda;
  int dbg;             // profiling knobs (uz_set_debug_flags): 1 = no epilogue body, 2 = no MMA, 4 = no A loads, 8 = no B loads
  unsigned long long* trace;   // profiling build: phase timestamps of CTA 0
};

// sum over the 32 lanes of v[j], j < 16; afterwards lanes l and l + 16 hold the total of column l
__device__ __forceinline__ float transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

__device__ __forceinline__ void load_row16(const __nv_bfloat16* src, float* out) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + j);
    out[j * 8 + 0] = uz::bf16lo(q.x); out[j * 8 + 1] = uz::bf16hi(q.x);
    out[j * 8 + 2] = uz::bf16lo(q.y); out[j * 8 + 3] = uz::bf16hi(q.y);
    out[j * 8 + 4] = uz::bf16lo(q.z); out[j * 8 + 5] = uz::bf16hi(q.z);
    out[j * 8 + 6] = uz::bf16lo(q.w); out[j * 8 + 7] = uz::bf16hi(q.w);
  }
}

template <int KC, int KB>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x A][stages x B] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t swz = KC * 2;                  // swizzle span in bytes = one smem row
  const uint32_t a_box = kBlockM * swz, b_box = p.BN * swz;       // one channel block of activations / weights
  const uint32_t a_bytes = a_box * p.KB;
  const uint32_t b_bytes = b_box * p.KB;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.stages * a_bytes;
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_scale[256];
  __shared__ float s_shift[256];
  __shared__ float s_stats[4][2][256];
  __shared__ float s_scale2[256];
  __shared__ float s_shift2[256];
  __shared__ float s_cta[2][256];     // fused BatchNorm: this CTA's (tile's) per-channel sum / sum of squares

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int tile = blockIdx.x;
  const int n_chunk = blockIdx.y;
  const int tx = tile % p.tilesW;
  const int ty = (tile / p.tilesW) % p.tilesH;
  const int tn = tile / (p.tilesW * p.tilesH);
  const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = tn * p.TN;
  const int c_out0 = n_chunk * p.BN;

  // a pipeline stage holds KB channel blocks of one tap (KB * KC channels): every stage hand-over costs ~0.2 us of
  // barrier latency, which -- not the loads -- bounded these small launches when a stage was a single 64-channel block
  const int kgroups = p.Cin / (KC * KB);
  const int k_iters = p.taps * kgroups;

  UZ_TRACE(p.trace, warp == 0 ? 0 : 15);
  if (warp == 0 && lane == 0) {
    uz::tma_prefetch_desc(&tmap_x);
    uz::tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        uz::mbar_init(&full_bar[s], 1);
        uz::mbar_init(&empty_bar[s], 1);
      }
      uz::mbar_init(&accum_bar, 1);
      uz::fence_barrier_init();
    }
    __syncwarp();
    UZ_TRACE(p.trace, 1);
    uz::tmem_alloc(&tmem_base_slot, p.tmem_cols);
    UZ_TRACE(p.trace, 2);
  }
  uz::pdl_prologue();   // everything above is independent of the previous kernel's output
  for (int c = threadIdx.x; c < p.BN; c += kThreads) {
    s_scale[c] = p.scale ? p.scale[c_out0 + c] : 1.f;
    s_shift[c] = p.shift ? p.shift[c_out0 + c] : 0.f;
    s_scale2[c] = p.bn_y ? p.bn_scale[c_out0 + c] : 1.f;
    s_shift2[c] = p.bn_y ? p.bn_shift[c_out0 + c] : 0.f;
  }
  uz::tc_fence_before();
  __syncthreads();
  uz::tc_fence_after();
  const uint32_t tmem_base = uz::uniform_u32(tmem_base_slot);
  UZ_TRACE(p.trace, warp == 0 ? 3 : 15);

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    {
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % p.stages;
        if (it >= p.stages) uz::mbar_wait(&empty_bar[s], ((it / p.stages) - 1) & 1);
        const int tap = it / kgroups;
        const int kg = it - tap * kgroups;
        int dy = 0, dx = 0;
        if (p.taps == 9) { dx = tap / 3 - 1; dy = tap % 3 - 1; }   // packed taps are dx-major (see uz_pack_conv_weight)
        if (uz::elect_one()) {
          uz::mbar_expect_tx(&full_bar[s], (UZ_DBG(p, 4) ? 0 : a_bytes) + (UZ_DBG(p, 8) ? 0 : b_bytes));
          for (int j = 0; j < KB; ++j) {
            const int c0 = (kg * KB + j) * KC;
            if (!UZ_DBG(p, 4)) uz::tma_load_4d(smem_a + s * a_bytes + j * a_box, &tmap_x, &full_bar[s], c0, x0 + dx, y0 + dy, n0);
            if (!UZ_DBG(p, 8)) uz::tma_load_3d(smem_b + s * b_bytes + j * b_box, &tmap_w, &full_bar[s], c0, c_out0, tap);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Descriptor high words are loop invariant; the KC/16 MMAs of a K step are unrolled with compile-time offsets so every
    // MMA gets its own uniform registers (a runtime loop made ptxas serialise issue behind a waterfall, ~180 cycles/MMA).
    const uint32_t idesc = uz::umma_idesc_bf16(kBlockM, p.BN, 0, 0);
    constexpr uint32_t sbo = 8 * swz;
    const uint64_t desc_hi = uz::umma_desc(0, 16, sbo, swz) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo0 = static_cast<uint32_t>(uz::umma_desc(0, 16, sbo, swz) & 0xFFFFFFFFull);
    for (int it = 0; it < k_iters; ++it) {
      const int s = it % p.stages;
      uz::mbar_wait(&full_bar[s], (it / p.stages) & 1);
      uz::tc_fence_after();
      if (it == 0) UZ_TRACE(p.trace, 4);
      if (it == k_iters - 1) UZ_TRACE(p.trace, 5);
      const uint32_t a_lo = desc_lo0 + (uz::smem_u32(smem_a + s * a_bytes) >> 4);
      const uint32_t b_lo = desc_lo0 + (uz::smem_u32(smem_b + s * b_bytes) >> 4);
      const uint32_t a_step = a_box >> 4, b_step = b_box >> 4;
      if (uz::elect_one()) {                     // ONE election per stage: every extra warp-level instruction in this
        if (!UZ_DBG(p, 2)) {                      // loop costs ~50 ns per iteration on these latency-bound launches
#pragma unroll
          for (int j = 0; j < KB; ++j) {
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              uz::tc_mma_f16(tmem_base, desc_hi | (a_lo + j * a_step + k * 2), desc_hi | (b_lo + j * b_step + k * 2), idesc,
                             (j == 0 && k == 0) ? static_cast<uint32_t>(it != 0) : 1u);
          }
        }
        uz::tc_commit(&empty_bar[s]);           // frees the smem slot once these MMAs retire
        if (it == k_iters - 1) uz::tc_commit(&accum_bar);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // One thread per output pixel (tile row).  No staging: each thread streams its pixel's channel vector to global
    // memory as 16-byte stores (32 contiguous bytes per 16 channels = whole sectors); scale/shift come from shared
    // memory; BatchNorm statistics are reduced across the 32 rows of a warp with a shuffle transpose-reduce.
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int ew = warp - 2;
    const int row = q * 32 + lane;
    const int box_px = p.TW * p.TH;
    const int xx = row % p.TW, yy = (row / p.TW) % p.TH, nn = row / box_px;
    const bool valid = (n0 + nn) < p.N && !UZ_DBG(p, 1);
    const size_t pix = (static_cast<size_t>(n0 + nn) * p.H + (y0 + yy)) * p.W + (x0 + xx);
    __nv_bfloat16* dst = p.y + pix * p.ldy + c_out0;
    const bool bn = p.bn_y != nullptr, rs = p.res != nullptr;
    const __nv_bfloat16* bn_row = bn ? p.bn_y + pix * p.bn_ldy + c_out0 : nullptr;
    const __nv_bfloat16* res_row = rs ? p.res + pix * p.ld_res + c_out0 : nullptr;
    uz::mbar_wait(&accum_bar, 0);
    uz::tc_fence_after();
    UZ_TRACE(p.trace, warp == 2 ? 6 : 15);
    for (int c = 0; c < (UZ_DBG(p, 1) ? 0 : p.BN); c += 16) {
      float yv[16], rv[16];
      if (bn) {
#pragma unroll
        for (int j = 0; j < 16; ++j) yv[j] = 0.f;
        if (valid) load_row16(bn_row + c, yv);
      }
      if (rs) {
#pragma unroll
        for (int j = 0; j < 16; ++j) rv[j] = 0.f;
        if (valid) load_row16(res_row + c, rv);
      }
      uint32_t r[16];
      uz::tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      uz::tmem_ld_wait();
      uint32_t packed[8];
      float v[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v0 = fmaf(__uint_as_float(r[2 * j]), s_scale[c + 2 * j], s_shift[c + 2 * j]);
        float v1 = fmaf(__uint_as_float(r[2 * j + 1]), s_scale[c + 2 * j + 1], s_shift[c + 2 * j + 1]);
        if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        if (rs) { v0 = fmaf(p.res_sign, v0, rv[2 * j]); v1 = fmaf(p.res_sign, v1, rv[2 * j + 1]); }
        if (bn && p.bn_relu) {
          if (!(fmaf(yv[2 * j], s_scale2[c + 2 * j], s_shift2[c + 2 * j]) > 0.f)) v0 = 0.f;
          if (!(fmaf(yv[2 * j + 1], s_scale2[c + 2 * j + 1], s_shift2[c + 2 * j + 1]) > 0.f)) v1 = 0.f;
        }
        packed[j] = uz::pack_bf16x2(v0, v1);
        v[2 * j] = valid ? uz::bf16lo(packed[j]) : 0.f;       // statistics of the stored values, valid rows only
        v[2 * j + 1] = valid ? uz::bf16hi(packed[j]) : 0.f;
      }
      if (valid) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c);
        d4[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        d4[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
      }
      if (p.stats || bn || p.fuse_bn) {
        float sq[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sq[j] = v[j] * (bn ? yv[j] : v[j]);
        const float s1 = transpose_reduce16(v, lane);
        const float s2 = transpose_reduce16(sq, lane);
        if (lane < 16) {
          s_stats[ew][0][c + lane] = s1;
          s_stats[ew][1][c + lane] = s2;
        }
      }
    }
    if (p.stats || p.bn_sums || p.fuse_bn) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = threadIdx.x - 64;
      float* acc = p.stats ? p.stats : p.bn_sums;
      for (int i = et; i < 2 * p.BN; i += 128) {
        const int which = i / p.BN, c = i - which * p.BN;
        if (!UZ_DBG(p, 1)) {
          const float t = (s_stats[0][which][c] + s_stats[1][which][c]) + (s_stats[2][which][c] + s_stats[3][which][c]);
          if (p.fuse_bn) s_cta[which][c] = t;
          else if (p.stats_rows) acc[(static_cast<size_t>(tile) * 2 + which) * p.Cout + c_out0 + c] = t;
          else atomicAdd(acc + which * p.Cout + c_out0 + c, t);
        }
      }
    }
  }

  if (p.fuse_bn) {
    // ===================== BatchNorm statistics across the cluster (all threads of all CTAs take part in the barriers)
    uz::tc_fence_before();
    uz::cluster_sync_all();                      // every tile's s_cta is complete and visible cluster-wide
    if (warp >= 2) {
      const int et = threadIdx.x - 64;
      const uint32_t ntiles = uz::cluster_nctarank();
      for (int c = et; c < p.BN; c += 128) {
        float s1 = 0.f, s2 = 0.f;
        for (uint32_t r = 0; r < ntiles; ++r) {  // fixed rank order: run-to-run deterministic
          s1 += uz::cluster_ld_f32(uz::cluster_map(&s_cta[0][c], r));
          s2 += uz::cluster_ld_f32(uz::cluster_map(&s_cta[1][c], r));
        }
        const int cg = c_out0 + c;
        const float mean = s1 / p.bn_count;
        const float var = fmaxf(s2 / p.bn_count - mean * mean, 0.f);
        const float invstd = rsqrtf(var + p.bn_eps);
        const float sc = (p.bn_gamma ? p.bn_gamma[cg] : 1.f) * invstd;
        const float sh = (p.bn_beta ? p.bn_beta[cg] : 0.f) - mean * sc;
        s_scale2[c] = sc;
        s_shift2[c] = sh;
        if (uz::cluster_ctarank() == 0) {
          p.bn_scale_out[cg] = sc;
          p.bn_shift_out[cg] = sh;
          p.bn_mean_out[cg] = mean;
          p.bn_invstd_out[cg] = invstd;
          if (p.bn_running_mean) {
            const float unbiased = p.bn_count > 1.f ? var * p.bn_count / (p.bn_count - 1.f) : var;
            float rm = p.bn_running_mean[cg], rv = p.bn_running_var[cg];
            for (int u = 0; u < p.bn_updates; ++u) {
              rm = (1.f - p.bn_momentum) * rm + p.bn_momentum * mean;
              rv = (1.f - p.bn_momentum) * rv + p.bn_momentum * unbiased;
            }
            p.bn_running_mean[cg] = rm;
            p.bn_running_var[cg] = rv;
          }
        }
      }
    }
    uz::cluster_sync_all();                      // nobody reads remote shared memory any more; s_scale2 / s_shift2 visible
    if (warp >= 2) {
      // second pass over the accumulators still in TMEM: a = act(BatchNorm(y)) from the bf16-rounded y that was stored
      uz::tc_fence_after();
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int box_px = p.TW * p.TH;
      const int xx = row % p.TW, yy = (row / p.TW) % p.TH, nn = row / box_px;
      const bool valid = (n0 + nn) < p.N;
      const size_t pix = (static_cast<size_t>(n0 + nn) * p.H + (y0 + yy)) * p.W + (x0 + xx);
      __nv_bfloat16* dst = p.a_out + pix * p.lda + c_out0;
      for (int c = 0; c < p.BN; c += 16) {
        uint32_t r[16];
        uz::tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
        uz::tmem_ld_wait();
        uint32_t packed[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t yq = uz::pack_bf16x2(fmaf(__uint_as_float(r[2 * j]), s_scale[c + 2 * j], s_shift[c + 2 * j]),
                                              fmaf(__uint_as_float(r[2 * j + 1]), s_scale[c + 2 * j + 1], s_shift[c + 2 * j + 1]));
          float v0 = fmaf(uz::bf16lo(yq), s_scale2[c + 2 * j], s_shift2[c + 2 * j]);
          float v1 = fmaf(uz::bf16hi(yq), s_scale2[c + 2 * j + 1], s_shift2[c + 2 * j + 1]);
          if (p.bn_act_relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          packed[j] = uz::pack_bf16x2(v0, v1);
        }
        if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + c);
          d4[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          d4[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      }
    }
  }

  UZ_TRACE(p.trace, warp == 2 ? 7 : 15);
  uz::tc_fence_before();
  __syncthreads();
  UZ_TRACE(p.trace, warp == 0 ? 8 : 15);
  if (warp == 1) {
    uz::tc_fence_after();
    uz::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int pow2_div_le(int v, int cap) {
  int t = 1;
  while (t * 2 <= cap && v % (t * 2) == 0) t *= 2;
  return t;
}

}  // namespace

extern "C" int uz_conv_tile_geometry(int N, int H, int W, int* TW, int* TH, int* TN, int* num_tiles) {
  if (N <= 0 || H <= 0 || W <= 0) return UZ_ERR_ARG;
  const int tw = pow2_div_le(W, 16);
  const int th = pow2_div_le(H, kBlockM / tw);
  const int tn = kBlockM / (tw * th);
  if (TW) *TW = tw;
  if (TH) *TH = th;
  if (TN) *TN = tn;
  if (num_tiles) *num_tiles = (W / tw) * (H / th) * ((N + tn - 1) / tn);
  return UZ_OK;
}

extern "C" int uz_conv_fwd(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout,
                           int taps, void* y, int ldy, const float* scale, const float* shift, int relu,
                           float* stats_partial, void* stream) {
  return uz_conv_fwd_ex(x, N, H, W, Cin, ldx, w_packed, Cout, taps, y, ldy, scale, shift, relu, stats_partial, nullptr,
                        stream);
}

extern "C" int uz_conv_stats_rows(int N, int H, int W, int Cin, int Cout, int taps) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return -1;
  if (taps == 9 && !UZ_KNOB(32)) {
    const int rows = uz::conv2_stats_rows(N, H, W, Cin, Cout);
    if (rows > 0) return rows;
  }
  int tiles = 0;
  uz_conv_tile_geometry(N, H, W, nullptr, nullptr, nullptr, &tiles);
  return tiles;
}

namespace {
// training-mode BatchNorm + activation fused into the generic kernel (one thread-block cluster over the pixel tiles)
struct FusedBn {
  float count, eps, momentum;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  int updates, relu;
  float* scale_out;
  float* shift_out;
  float* mean_out;
  float* invstd_out;
  void* a;
  int lda;
};
constexpr int kMaxClusterTiles = 8;     // portable cluster size

int conv_fwd_impl(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, int taps, void* y,
                  int ldy, const float* scale, const float* shift, int relu, float* stats_partial, const UzConvExtra* ex,
                  const FusedBn* fb, void* stream);
}  // namespace

extern "C" int uz_conv_fwd_ex(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout,
                              int taps, void* y, int ldy, const float* scale, const float* shift, int relu,
                              float* stats_partial, const UzConvExtra* ex, void* stream) {
  return conv_fwd_impl(x, N, H, W, Cin, ldx, w_packed, Cout, taps, y, ldy, scale, shift, relu, stats_partial, ex, nullptr,
                       stream);
}

// 1 if uz_conv_bn_act_fused has a plan: a layer of the generic kernel (maps that the persistent kernel does not take)
// whose pixel tiles fit one portable cluster
extern "C" int uz_conv_bn_fused_supported(int N, int H, int W, int Cin, int Cout, int taps) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || Cin % 16 || Cout % 16 || (taps != 9 && taps != 1)) return 0;
  if (taps == 9 && !UZ_KNOB(32) && uz::conv2_stats_rows(N, H, W, Cin, Cout) > 0) return 0;
  int tiles = 0;
  uz_conv_tile_geometry(N, H, W, nullptr, nullptr, nullptr, &tiles);
  int bn = Cout;
  while (bn > 256) { if (bn % 32) return 0; bn /= 2; }
  return tiles <= kMaxClusterTiles ? 1 : 0;
}

// Conv2D of the reference in training mode as ONE launch (torchlayers.py:18-21: conv + bias -> BatchNorm2d(batch
// statistics, eps, momentum) -> ReLU): y = conv(x) + bias (bf16, kept for backward), per-channel statistics of the stored
// y exchanged between the pixel-tile CTAs of a cluster through distributed shared memory (fixed order: deterministic),
// a = act(y * scale + shift) written from the accumulators still in TMEM.  scale / shift / mean / invstd [Cout] are the
// coefficients backward needs; running statistics get `stat_updates` momentum updates (unbiased variance).
extern "C" int uz_conv_bn_act_fused(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout,
                                    int taps, const float* bias, const float* gamma, const float* beta, float eps,
                                    float momentum, float* running_mean, float* running_var, int stat_updates, int relu,
                                    void* y, int ldy, void* a, int lda, float* scale_out, float* shift_out,
                                    float* mean_out, float* invstd_out, void* stream) {
  UZ_CHECK_ARG(uz_conv_bn_fused_supported(N, H, W, Cin, Cout, taps), "uz_conv_bn_act_fused: no plan for this layer");
  UZ_CHECK_ARG(a && scale_out && shift_out && mean_out && invstd_out, "uz_conv_bn_act_fused: null pointer");
  UZ_CHECK_ARG(lda % 8 == 0 && lda >= Cout && (reinterpret_cast<uintptr_t>(a) & 15) == 0, "uz_conv_bn_act_fused: bad a");
  UZ_CHECK_ARG(stat_updates >= 0 && stat_updates <= 4, "uz_conv_bn_act_fused: stat_updates %d", stat_updates);
  FusedBn fb{};
  fb.count = static_cast<float>(static_cast<double>(N) * H * W);
  fb.eps = eps; fb.momentum = momentum; fb.gamma = gamma; fb.beta = beta;
  fb.running_mean = running_mean; fb.running_var = running_mean ? running_var : nullptr;
  fb.updates = stat_updates; fb.relu = relu;
  fb.scale_out = scale_out; fb.shift_out = shift_out; fb.mean_out = mean_out; fb.invstd_out = invstd_out;
  fb.a = a; fb.lda = lda;
  return conv_fwd_impl(x, N, H, W, Cin, ldx, w_packed, Cout, taps, y, ldy, nullptr, bias, 0, nullptr, nullptr, &fb, stream);
}

namespace {
int conv_fwd_impl(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, int taps, void* y,
                  int ldy, const float* scale, const float* shift, int relu, float* stats_partial, const UzConvExtra* ex,
                  const FusedBn* fb, void* stream) {
  UZ_CHECK_ARG(x && w_packed && y, "uz_conv_fwd: null pointer");
  if (ex) {
    UZ_CHECK_ARG(!ex->bn_y || (ex->bn_scale && ex->bn_shift && ex->bn_sums && ex->bn_ldy % 8 == 0 && ex->bn_ldy >= Cout &&
                               (reinterpret_cast<uintptr_t>(ex->bn_y) & 15) == 0 && !stats_partial),
                 "uz_conv_fwd_ex: fused BatchNorm backward needs scale, shift, sums, an aligned y and no forward statistics");
    UZ_CHECK_ARG(!ex->residual || (ex->ld_res % 8 == 0 && ex->ld_res >= Cout &&
                                   (reinterpret_cast<uintptr_t>(ex->residual) & 15) == 0),
                 "uz_conv_fwd_ex: bad residual operand");
    UZ_CHECK_ARG(!ex->stats_rows || stats_partial, "uz_conv_fwd_ex: stats_rows without a statistics buffer");
  }
  UZ_CHECK_ARG(taps == 9 || taps == 1, "uz_conv_fwd: taps must be 9 or 1 (got %d)", taps);
  UZ_CHECK_ARG(Cin % 16 == 0 && Cin > 0, "uz_conv_fwd: Cin must be a positive multiple of 16 (got %d)", Cin);
  UZ_CHECK_ARG(Cout % 16 == 0 && Cout > 0, "uz_conv_fwd: Cout must be a positive multiple of 16 (got %d)", Cout);
  UZ_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= Cin && ldy >= Cout, "uz_conv_fwd: bad pixel strides");
  UZ_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
               "uz_conv_fwd: pointers must be 16-byte aligned");
  if UZ_KNOB(128) return UZ_OK;   // measurement knob: step time without the conv kernels
  if (taps == 9 && !stats_partial && !(ex && (ex->bn_y || ex->residual)) && uz::conv_small_supported(N, H, W, Cin, Cout)) {
    // 2x2 ... 4x4 maps: latency-bound, CUDA-core kernel with the BatchNorm fusion inside one CTA (conv_small.cu)
    uz::SmallBn sb{};
    if (fb) {
      sb.count = fb->count; sb.eps = fb->eps; sb.momentum = fb->momentum; sb.gamma = fb->gamma; sb.beta = fb->beta;
      sb.running_mean = fb->running_mean; sb.running_var = fb->running_var; sb.updates = fb->updates; sb.relu = fb->relu;
      sb.scale_out = fb->scale_out; sb.shift_out = fb->shift_out; sb.mean_out = fb->mean_out;
      sb.invstd_out = fb->invstd_out; sb.a = fb->a; sb.lda = fb->lda;
    }
    return uz::conv_small_launch(x, N, H, W, Cin, ldx, w_packed, Cout, y, ldy, scale, shift, relu, fb ? &sb : nullptr, stream);
  }
  if (taps == 9 && !UZ_KNOB(32) && !fb) {
    int handled = 0;
    int rc = uz::conv2_launch(x, N, 0, H, W, Cin, ldx, w_packed, Cout, y, ldy, scale, shift, relu, stats_partial, ex,
                              stream, &handled);
    if (rc || handled) return rc;
  }
  if UZ_KNOB(2048) return UZ_OK;  // measurement knob: step time without the generic-kernel launches
  ConvParams p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.taps = taps;
  int tiles = 0;
  uz_conv_tile_geometry(N, H, W, &p.TW, &p.TH, &p.TN, &tiles);
  p.tilesW = W / p.TW; p.tilesH = H / p.TH;
  p.KC = (Cin % 64 == 0) ? 64 : ((Cin % 32 == 0) ? 32 : 16);
  // output-channel split: one CTA covers all of Cout unless Cout > 256 or the grid would leave SMs idle
  int bn = Cout;
  int splits = 1;
  while (bn > 256 || (tiles * splits < uz::num_sms() && bn % 32 == 0 && bn >= 64)) {
    if (bn % 32 != 0) break;
    bn /= 2; splits *= 2;
  }
  UZ_CHECK_ARG(bn <= 256 && bn % 16 == 0, "uz_conv_fwd: unsupported Cout %d", Cout);
  p.BN = bn;
  p.ldy = ldy; p.relu = relu;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift; p.stats = stats_partial;
  if (ex) {
    p.stats_rows = ex->stats_rows;
    p.bn_y = static_cast<const __nv_bfloat16*>(ex->bn_y); p.bn_ldy = ex->bn_ldy;
    p.bn_scale = ex->bn_scale; p.bn_shift = ex->bn_shift; p.bn_relu = ex->bn_relu; p.bn_sums = ex->bn_sums;
    p.res = static_cast<const __nv_bfloat16*>(ex->residual); p.ld_res = ex->ld_res;
    p.res_sign = ex->res_sign < 0 ? -1.f : 1.f;
  }
  if (fb) {
    p.fuse_bn = 1;
    p.bn_count = fb->count; p.bn_eps = fb->eps; p.bn_momentum = fb->momentum;
    p.bn_gamma = fb->gamma; p.bn_beta = fb->beta;
    p.bn_running_mean = fb->running_mean; p.bn_running_var = fb->running_var;
    p.bn_updates = fb->updates; p.bn_act_relu = fb->relu;
    p.bn_scale_out = fb->scale_out; p.bn_shift_out = fb->shift_out;
    p.bn_mean_out = fb->mean_out; p.bn_invstd_out = fb->invstd_out;
    p.a_out = static_cast<__nv_bfloat16*>(fb->a); p.lda = fb->lda;
  }
  p.dbg = uz::g_conv_debug_flags;
#ifdef UZ_PROFILE_KNOBS
  p.trace = uz::g_trace;
#endif
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(bn)) cols *= 2;
  p.tmem_cols = cols;
  const uint32_t swz = p.KC * 2;
  // channel blocks per stage: the largest divisor of Cin / KC (<= 4) that still leaves >= 3 stages
  const int kblocks = Cin / p.KC;
  p.KB = 1;
  for (int kb = 4; kb >= 2; --kb) {
    if (kblocks % kb == 0 && (196 * 1024) / (static_cast<size_t>(kBlockM + bn) * swz * kb) >= 2 &&
        !UZ_KNOB(65536)) {
      p.KB = kb;
      break;
    }
  }
  const size_t stage_bytes = static_cast<size_t>(kBlockM + bn) * swz * p.KB;
  const size_t out_bytes = static_cast<size_t>(kBlockM) * (bn * 2 + 16);
  const int k_iters = taps * (kblocks / p.KB);
  int stages = static_cast<int>((196 * 1024) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > k_iters) stages = k_iters;
  if (stages < 1) stages = 1;
  p.stages = stages;
  size_t smem = stages * stage_bytes + 1024;  // + alignment slack
  (void)out_bytes;

  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(W) * ldx * 2,
                           static_cast<uint64_t>(H) * W * ldx * 2};
    uint32_t box[4] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.TW), static_cast<uint32_t>(p.TH),
                       static_cast<uint32_t>(p.TN)};
    int rc = uz::make_tmap_bf16(&tx, x, 4, dims, strides, box, swz);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Cout), static_cast<uint64_t>(taps)};
    uint64_t strides[2] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Cout) * Cin * 2};
    uint32_t box[3] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(bn), 1};
    int rc = uz::make_tmap_bf16(&tw, w_packed, 3, dims, strides, box, swz);
    if (rc) return rc;
  }
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const ConvParams);
  static const KernelFn table[3][4] = {
      {conv_tc_kernel<64, 1>, conv_tc_kernel<64, 2>, conv_tc_kernel<64, 3>, conv_tc_kernel<64, 4>},
      {conv_tc_kernel<32, 1>, conv_tc_kernel<32, 2>, conv_tc_kernel<32, 3>, conv_tc_kernel<32, 4>},
      {conv_tc_kernel<16, 1>, conv_tc_kernel<16, 2>, conv_tc_kernel<16, 3>, conv_tc_kernel<16, 4>}};
  const int kci = p.KC == 64 ? 0 : (p.KC == 32 ? 1 : 2);
  KernelFn kernel = table[kci][p.KB - 1];
  static size_t attr_bytes_kc[3][4] = {};
  size_t& attr_bytes = attr_bytes_kc[kci][p.KB - 1];
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      uz::set_error("uz_conv_fwd: cannot raise dynamic smem limit to %zu: %s", static_cast<size_t>(smem), cudaGetErrorString(e));
      return UZ_ERR_CUDA;
    }
    attr_bytes = smem;
  }
  dim3 grid(tiles, splits, 1);
  if (fb) uz::launch_cluster(kernel, grid, kThreads, smem, static_cast<cudaStream_t>(stream), dim3(tiles, 1, 1), tx, tw, p);
  else uz::launch(kernel, grid, kThreads, smem, static_cast<cudaStream_t>(stream), tx, tw, p);
  UZ_CHECK_LAUNCH("uz_conv_fwd");
  return UZ_OK;
}
}  // namespace

// volumes: x bf16 NDHWC [N,D,H,W,Cin]; taps 27 (3x3x3, pad 1; packed [(kd*3+kw)*3+kh][Cout][Cin]) or 1 (1x1x1)
extern "C" int uz_conv3d_fwd(const void* x, int N, int D, int H, int W, int Cin, int ldx, const void* w_packed, int Cout,
                             int taps, void* y, int ldy, const float* scale, const float* shift, int relu,
                             float* stats_partial, void* stream) {
  return uz_conv3d_fwd_ex(x, N, D, H, W, Cin, ldx, w_packed, Cout, taps, y, ldy, scale, shift, relu, stats_partial, nullptr,
                          stream);
}

extern "C" int uz_conv3d_fwd_ex(const void* x, int N, int D, int H, int W, int Cin, int ldx, const void* w_packed,
                                int Cout, int taps, void* y, int ldy, const float* scale, const float* shift, int relu,
                                float* stats_partial, const UzConvExtra* ex, void* stream) {
  UZ_CHECK_ARG(taps == 27 || taps == 1, "uz_conv3d_fwd: taps must be 27 or 1 (got %d)", taps);
  UZ_CHECK_ARG(D > 0, "uz_conv3d_fwd: D must be positive");
  if (taps == 1) {   // pointwise: a volume is N*D images
    UZ_CHECK_ARG(static_cast<long long>(N) * D < (1ll << 30), "uz_conv3d_fwd: batch too large");
    return uz_conv_fwd_ex(x, N * D, H, W, Cin, ldx, w_packed, Cout, 1, y, ldy, scale, shift, relu, stats_partial, ex,
                          stream);
  }
  UZ_CHECK_ARG(!ex || !ex->stats_rows, "uz_conv3d_fwd_ex: per-CTA statistics rows are a 2-D feature");
  UZ_CHECK_ARG(x && w_packed && y, "uz_conv3d_fwd: null pointer");
  UZ_CHECK_ARG(Cin % 16 == 0 && Cin > 0, "uz_conv3d_fwd: Cin must be a positive multiple of 16 (got %d)", Cin);
  UZ_CHECK_ARG(Cout % 16 == 0 && Cout > 0, "uz_conv3d_fwd: Cout must be a positive multiple of 16 (got %d)", Cout);
  UZ_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= Cin && ldy >= Cout, "uz_conv3d_fwd: bad pixel strides");
  UZ_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
               "uz_conv3d_fwd: pointers must be 16-byte aligned");
  if UZ_KNOB(128) return UZ_OK;
  int handled = 0;
  int rc = uz::conv2_launch(x, N, D, H, W, Cin, ldx, w_packed, Cout, y, ldy, scale, shift, relu, stats_partial, ex,
                            stream, &handled);
  if (rc) return rc;
  UZ_CHECK_ARG(handled, "uz_conv3d_fwd: no kernel plan for Cin=%d Cout=%d", Cin, Cout);
  return UZ_OK;
}

extern "C" int uz_set_debug_flags(int flags) {
#ifndef UZ_PROFILE_KNOBS
  UZ_CHECK_ARG((flags & ~32) == 0, "uz_set_debug_flags: this library was built without -DUZ_PROFILE_KNOBS; only bit 32 "
               "(route 3x3 layers to the generic kernel) is available (got %d)", flags);
#endif
  uz::g_conv_debug_flags = flags;
  return UZ_OK;
}

extern "C" int uz_conv_uses_persistent_kernel(int N, int H, int W, int Cin, int Cout, int taps) {
  if (taps == 9 && !UZ_KNOB(32)) return uz::conv2_stats_rows(N, H, W, Cin, Cout) > 0 ? 1 : 0;
  return 0;
}
