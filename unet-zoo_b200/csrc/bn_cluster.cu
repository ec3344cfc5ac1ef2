// BatchNorm(train) + ReLU backward as ONE launch on a thread-block cluster (replaces the reduce + apply pair of
// elementwise.cu; the host side routes layers of up to 16x16x12 pixels here -- 64 of the 106 Conv2D layers of PHiSeg-7/5,
// reference torchlayers.py:18-21 backward through nn.BatchNorm2d / nn.ReLU; measured in the captured step the pair is
// faster on larger maps, where 2-4 CTAs per SM hide the HBM latency better than one staged slice per SM).
//
// BatchNorm backward needs two per-channel sums over ALL pixels (sum g, sum g*y with g = dout * [ReLU active]) before a
// single output element can be written -- a grid-wide dependency that cost a second launch (3-10 us of the serial dgrad
// chain per layer).  Here a cluster owns a group of 16 channels: its K CTAs split the pixels, stage their slice of
// (g, y) in shared memory while accumulating the sums, exchange the 32 partial sums through distributed shared memory
// (fixed rank order => run-to-run deterministic, no atomics), then write dy from the staged copy.  HBM traffic is the
// algorithmic minimum (read dout, y once; write dy once).
#include <cooperative_groups.h>

#include "common.cuh"
#include "unetzoo_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kRows = kThreads / 2;          // pixel rows per iteration: two threads (8 channels each) per row
constexpr int kMaxStagePix = 3072;           // 3072 px x 16 ch x 2 B x 2 tensors = 192 KB

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = uz::bf16lo(v.x); f[1] = uz::bf16hi(v.x); f[2] = uz::bf16lo(v.y); f[3] = uz::bf16hi(v.y);
  f[4] = uz::bf16lo(v.z); f[5] = uz::bf16hi(v.z); f[6] = uz::bf16lo(v.w); f[7] = uz::bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(uz::pack_bf16x2(f[0], f[1]), uz::pack_bf16x2(f[2], f[3]), uz::pack_bf16x2(f[4], f[5]),
                    uz::pack_bf16x2(f[6], f[7]));
}

struct BnBwdParams {
  const __nv_bfloat16* dout; int ldd;
  const __nv_bfloat16* y; int ldy;
  const float* scale; const float* shift;
  const float* gamma; const float* mean; const float* invstd;
  float* dgamma; float* dbeta;
  __nv_bfloat16* dy; int lddy;
  int relu, npix, pix_per_cta, staged;
  float count;
};

// grid (K, C/16), cluster (K, 1, 1): blockIdx.x = pixel slice (= cluster rank), blockIdx.y = 16-channel group
__global__ void __launch_bounds__(kThreads) bn_bwd_cluster_kernel(const BnBwdParams p) {
  uz::pdl_prologue();
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) uint8_t stage_raw[];      // staged: [2][pix_per_cta][16] bf16 (masked g; y)
  __shared__ float red[kThreads / 32][2][16];               // per warp: [half][sg 0..7, sgy 0..7]
  __shared__ float part[32];                                // this CTA: sg[16], sgy[16]
  __shared__ float tot[32];
  uint4* stage_g = reinterpret_cast<uint4*>(stage_raw);     // consecutive threads -> consecutive 16 B: no bank conflicts
  uint4* stage_y = stage_g + static_cast<size_t>(p.pix_per_cta) * 2;

  const int half = threadIdx.x & 1;
  const int r = threadIdx.x >> 1;
  const int c0 = blockIdx.y * 16 + half * 8;
  const int p0 = blockIdx.x * p.pix_per_cta;
  int pn = p.npix - p0;
  if (pn > p.pix_per_cta) pn = p.pix_per_cta;
  if (pn < 0) pn = 0;

  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = p.scale[c0 + j]; sh[j] = p.shift[c0 + j]; }
  float sg[8], sgy[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sg[j] = 0.f; sgy[j] = 0.f; }

  const __nv_bfloat16* gsrc = p.dout + static_cast<size_t>(p0) * p.ldd + c0;
  const __nv_bfloat16* ysrc = p.y + static_cast<size_t>(p0) * p.ldy + c0;
  auto pass1 = [&](int px, const uint4& vg, const uint4& vy) {
    float g[8], yy[8];
    unpack8(vg, g);
    unpack8(vy, yy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = (!p.relu || fmaf(yy[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
      g[j] = m;
      sg[j] += m;
      sgy[j] = fmaf(m, yy[j], sgy[j]);
    }
    if (p.staged) {
      stage_g[px * 2 + half] = pack8(g);                    // masking a bf16 value is exact
      stage_y[px * 2 + half] = vy;
    }
  };
  int px = r;
  for (; px + 3 * kRows < pn; px += 4 * kRows) {             // eight independent 16-byte loads in flight per thread
    uint4 vg[4], vy[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vg[u] = *reinterpret_cast<const uint4*>(gsrc + static_cast<size_t>(px + u * kRows) * p.ldd);
      vy[u] = *reinterpret_cast<const uint4*>(ysrc + static_cast<size_t>(px + u * kRows) * p.ldy);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) pass1(px + u * kRows, vg[u], vy[u]);
  }
  for (; px < pn; px += kRows)
    pass1(px, *reinterpret_cast<const uint4*>(gsrc + static_cast<size_t>(px) * p.ldd),
          *reinterpret_cast<const uint4*>(ysrc + static_cast<size_t>(px) * p.ldy));

  // rows of a warp: lanes of equal parity (xor 2..16), then the 8 warps through shared memory, all in a fixed order
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = 2; o <= 16; o <<= 1) {
      sg[j] += __shfl_xor_sync(0xffffffffu, sg[j], o);
      sgy[j] += __shfl_xor_sync(0xffffffffu, sgy[j], o);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[warp][lane][j] = sg[j]; red[warp][lane][8 + j] = sgy[j]; }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // part[0..15] = sum g of channel (half*8 + j), part[16..31] = sum g*y
    const int which = threadIdx.x >> 4, ch = threadIdx.x & 15;
    const int hf = ch >> 3, j = ch & 7;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) acc += red[w][hf][which * 8 + j];
    part[threadIdx.x] = acc;
  }
  cluster.sync();                                            // every CTA's part[] is complete and visible cluster-wide
  if (threadIdx.x < 32) {
    float acc = 0.f;
    const unsigned K = cluster.num_blocks();
    for (unsigned k = 0; k < K; ++k) acc += cluster.map_shared_rank(part, k)[threadIdx.x];   // fixed rank order
    tot[threadIdx.x] = acc;
  }
  cluster.sync();                                            // nobody reads remote shared memory after this point

  float ca[8], cb[8], cc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float tsg = tot[half * 8 + j], tsgy = tot[16 + half * 8 + j];
    const float mu = p.mean[c], is = p.invstd[c], gm = p.gamma ? p.gamma[c] : 1.f;
    const float sgx = (tsgy - mu * tsg) * is;
    const float mg = tsg / p.count, mgx = sgx / p.count;
    ca[j] = gm * is;
    cb[j] = -gm * is * is * mgx;
    cc[j] = -gm * is * mg + gm * is * is * mgx * mu;
    if (blockIdx.x == 0 && r == 0) {
      if (p.dgamma) p.dgamma[c] = sgx;
      if (p.dbeta) p.dbeta[c] = tsg;
    }
  }
  __nv_bfloat16* dst = p.dy + static_cast<size_t>(p0) * p.lddy + c0;
  if (p.staged) {
    for (px = r; px < pn; px += kRows) {
      float g[8], yy[8];
      unpack8(stage_g[px * 2 + half], g);
      unpack8(stage_y[px * 2 + half], yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = fmaf(ca[j], g[j], fmaf(cb[j], yy[j], cc[j]));
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(px) * p.lddy) = pack8(g);
    }
  } else {
    auto pass2 = [&](int q, const uint4& vg, const uint4& vy) {
      float g[8], yy[8];
      unpack8(vg, g);
      unpack8(vy, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float m = (!p.relu || fmaf(yy[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
        g[j] = fmaf(ca[j], m, fmaf(cb[j], yy[j], cc[j]));
      }
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(q) * p.lddy) = pack8(g);
    };
    px = r;
    for (; px + 3 * kRows < pn; px += 4 * kRows) {
      uint4 vg[4], vy[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        vg[u] = *reinterpret_cast<const uint4*>(gsrc + static_cast<size_t>(px + u * kRows) * p.ldd);
        vy[u] = *reinterpret_cast<const uint4*>(ysrc + static_cast<size_t>(px + u * kRows) * p.ldy);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) pass2(px + u * kRows, vg[u], vy[u]);
    }
    for (; px < pn; px += kRows)
      pass2(px, *reinterpret_cast<const uint4*>(gsrc + static_cast<size_t>(px) * p.ldd),
            *reinterpret_cast<const uint4*>(ysrc + static_cast<size_t>(px) * p.ldy));
  }
}

// pixel slices per cluster: enough CTAs to keep a slice short, at most the portable cluster size
inline int plan_cluster(long long npix, int* pix_per_cta, int* staged) {
  // <= 512 pixels per CTA (measured: 128 or 512 make no difference up to 16x16x12; one CTA for everything is 2x slower)
  int K = 1;
  while (K < 8 && (npix + K - 1) / K > 512) K *= 2;
  long long ppc = (npix + K - 1) / K;
  ppc = (ppc + 7) / 8 * 8;
  *pix_per_cta = static_cast<int>(ppc);
  *staged = ppc <= kMaxStagePix ? 1 : 0;
  return K;
}

}  // namespace

extern "C" int uz_bn_bwd_fused_supported(long long npix, int C) {
  return (npix > 0 && npix <= 49152 && C > 0 && C % 16 == 0) ? 1 : 0;
}

// dy = BatchNorm(train)+ReLU backward of dout through y (the conv output the statistics were taken of); dgamma, dbeta
// fp32 [C].  One launch; see the file header.  Call only if uz_bn_bwd_fused_supported(npix, C).
extern "C" int uz_bn_bwd_fused(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift,
                               int relu, float count, const float* gamma, const float* mean, const float* invstd,
                               float* dgamma, float* dbeta, void* dy, int lddy, long long npix, int C, void* stream) {
  UZ_CHECK_ARG(dout && y && scale && shift && mean && invstd && dy, "uz_bn_bwd_fused: null pointer");
  UZ_CHECK_ARG(uz_bn_bwd_fused_supported(npix, C), "uz_bn_bwd_fused: unsupported size (npix %lld, C %d)", npix, C);
  UZ_CHECK_ARG(ldd % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0 && ldd >= C && ldy >= C && lddy >= C,
               "uz_bn_bwd_fused: bad pixel strides");
  UZ_CHECK_ARG(((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0,
               "uz_bn_bwd_fused: pointers must be 16-byte aligned");
  BnBwdParams p{};
  p.dout = static_cast<const __nv_bfloat16*>(dout); p.ldd = ldd;
  p.y = static_cast<const __nv_bfloat16*>(y); p.ldy = ldy;
  p.scale = scale; p.shift = shift; p.gamma = gamma; p.mean = mean; p.invstd = invstd;
  p.dgamma = dgamma; p.dbeta = dbeta;
  p.dy = static_cast<__nv_bfloat16*>(dy); p.lddy = lddy;
  p.relu = relu; p.npix = static_cast<int>(npix); p.count = count;
  const int K = plan_cluster(npix, &p.pix_per_cta, &p.staged);
  const size_t smem = p.staged ? static_cast<size_t>(p.pix_per_cta) * 64 : 0;
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(bn_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      uz::set_error("uz_bn_bwd_fused: cannot raise dynamic smem limit to %zu: %s", smem, cudaGetErrorString(e));
      return UZ_ERR_CUDA;
    }
    attr_bytes = smem;
  }
  uz::launch_cluster(bn_bwd_cluster_kernel, dim3(K, C / 16, 1), kThreads, smem, static_cast<cudaStream_t>(stream),
                     dim3(K, 1, 1), p);
  UZ_CHECK_LAUNCH("uz_bn_bwd_fused");
  return UZ_OK;
}


// ---------------------------------------------------------------- large maps: one COOPERATIVE launch
// Maps too large for a cluster (64x64x12 pixels and up) used two launches -- sums, then apply -- each of them a 6-12 us
// kernel that streams 25-50 MB: mostly launch + ramp-up + tail.  Here a single-wave cooperative grid does both passes with
// a grid-wide barrier in between; the second pass re-reads dout / y from L2 (the tensors of all but the 128-channel layers
// fit).  The per-channel sums are accumulated with fp32 atomics like in the two-launch path.
namespace {

struct BnCoopParams {
  const __nv_bfloat16* dout; int ldd;
  const __nv_bfloat16* y; int ldy;
  const float* scale; const float* shift;
  const float* gamma; const float* mean; const float* invstd;
  float* sums;                 // [2][C], zero on entry
  float* dgamma; float* dbeta;
  __nv_bfloat16* dy; int lddy;
  int relu, C;
  long long npix;
  float count;
};

// <= 64 registers: two of these grids (2 blocks per SM each) are co-resident, e.g. the two encoders' backward streams
__global__ void __launch_bounds__(256, 4) bn_bwd_coop_kernel(const BnCoopParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float sm[];                       // pass 1: [rows][2][C] partial sums; pass 2: [3][C] coefficients
  const int C = p.C;
  const int chunks = C / 8;
  const int rows = blockDim.x / chunks;
  const int r = threadIdx.x / chunks;
  const int c0 = (threadIdx.x - r * chunks) * 8;
  const bool active = r < rows;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = active ? p.scale[c0 + j] : 0.f; sh[j] = active ? p.shift[c0 + j] : 0.f; }
  const size_t npix = static_cast<size_t>(p.npix);
  const size_t pstride = static_cast<size_t>(gridDim.x) * rows;
  // ---- pass 1: sum g, sum g*y with g = dout * [ReLU active]
  {
    float sg[8], sgy[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sg[j] = 0.f; sgy[j] = 0.f; }
    if (active) {
      auto one = [&](const uint4& vg, const uint4& vy) {
        float g[8], yy[8];
        unpack8(vg, g);
        unpack8(vy, yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float m = (!p.relu || fmaf(yy[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
          sg[j] += m;
          sgy[j] = fmaf(m, yy[j], sgy[j]);
        }
      };
      size_t pix = static_cast<size_t>(blockIdx.x) * rows + r;
      for (; pix + 3 * pstride < npix; pix += 4 * pstride) {
        uint4 vg[4], vy[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          vg[u] = *reinterpret_cast<const uint4*>(p.dout + (pix + u * pstride) * p.ldd + c0);
          vy[u] = *reinterpret_cast<const uint4*>(p.y + (pix + u * pstride) * p.ldy + c0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) one(vg[u], vy[u]);
      }
      for (; pix < npix; pix += pstride)
        one(*reinterpret_cast<const uint4*>(p.dout + pix * p.ldd + c0), *reinterpret_cast<const uint4*>(p.y + pix * p.ldy + c0));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sm[(r * 2) * C + c0 + j] = sg[j];
        sm[(r * 2 + 1) * C + c0 + j] = sgy[j];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      float acc = 0.f;
      for (int rr = 0; rr < rows; ++rr) acc += sm[rr * 2 * C + i];
      atomicAdd(p.sums + i, acc);
    }
  }
  __threadfence();
  grid.sync();
  // ---- pass 2: dy = A*g + B*y + Cc
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float tsg = __ldcg(p.sums + c), tsgy = __ldcg(p.sums + C + c);
    const float mu = p.mean[c], is = p.invstd[c], gm = p.gamma ? p.gamma[c] : 1.f;
    const float sgx = (tsgy - mu * tsg) * is;
    const float mg = tsg / p.count, mgx = sgx / p.count;
    sm[c] = gm * is;
    sm[C + c] = -gm * is * is * mgx;
    sm[2 * C + c] = -gm * is * mg + gm * is * is * mgx * mu;
    if (blockIdx.x == 0) {
      if (p.dgamma) p.dgamma[c] = sgx;
      if (p.dbeta) p.dbeta[c] = tsg;
    }
  }
  __syncthreads();
  if (active) {
    float ca[8], cb[8], cc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ca[j] = sm[c0 + j]; cb[j] = sm[C + c0 + j]; cc[j] = sm[2 * C + c0 + j]; }
    auto two = [&](const uint4& vg, const uint4& vy, size_t px) {
      float g[8], yy[8];
      unpack8(vg, g);
      unpack8(vy, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float m = (!p.relu || fmaf(yy[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
        g[j] = fmaf(ca[j], m, fmaf(cb[j], yy[j], cc[j]));
      }
      *reinterpret_cast<uint4*>(p.dy + px * p.lddy + c0) = pack8(g);
    };
    // same pixel assignment as pass 1: a block re-reads what it just read (L2, partly L1)
    size_t pix = static_cast<size_t>(blockIdx.x) * rows + r;
    for (; pix + 3 * pstride < npix; pix += 4 * pstride) {
      uint4 vg[4], vy[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        vg[u] = *reinterpret_cast<const uint4*>(p.dout + (pix + u * pstride) * p.ldd + c0);
        vy[u] = *reinterpret_cast<const uint4*>(p.y + (pix + u * pstride) * p.ldy + c0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) two(vg[u], vy[u], pix + u * pstride);
    }
    for (; pix < npix; pix += pstride)
      two(*reinterpret_cast<const uint4*>(p.dout + pix * p.ldd + c0), *reinterpret_cast<const uint4*>(p.y + pix * p.ldy + c0), pix);
  }
}

}  // namespace

// BatchNorm(train)+ReLU backward of a LARGE map in one cooperative launch; sums: [2][C] floats, zero on entry.
extern "C" int uz_bn_bwd_coop(const void* dout, int ldd, const void* y, int ldy, const float* scale, const float* shift,
                              int relu, float* sums, float count, const float* gamma, const float* mean,
                              const float* invstd, float* dgamma, float* dbeta, void* dy, int lddy, long long npix, int C,
                              void* stream) {
  UZ_CHECK_ARG(dout && y && scale && shift && sums && mean && invstd && dy, "uz_bn_bwd_coop: null pointer");
  UZ_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && ldd % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0 && npix > 0,
               "uz_bn_bwd_coop: bad arguments");
  UZ_CHECK_ARG(((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0,
               "uz_bn_bwd_coop: pointers must be 16-byte aligned");
  BnCoopParams p{};
  p.dout = static_cast<const __nv_bfloat16*>(dout); p.ldd = ldd;
  p.y = static_cast<const __nv_bfloat16*>(y); p.ldy = ldy;
  p.scale = scale; p.shift = shift; p.gamma = gamma; p.mean = mean; p.invstd = invstd;
  p.sums = sums; p.dgamma = dgamma; p.dbeta = dbeta;
  p.dy = static_cast<__nv_bfloat16*>(dy); p.lddy = lddy;
  p.relu = relu; p.C = C; p.npix = npix; p.count = count;
  const int chunks = C / 8;
  int threads = 256;
  if (chunks > threads) threads = ((chunks + 31) / 32) * 32;
  const int rows = threads / chunks;
  size_t smem = static_cast<size_t>(rows) * 2 * C * sizeof(float);
  if (smem < static_cast<size_t>(3) * C * sizeof(float)) smem = static_cast<size_t>(3) * C * sizeof(float);
  UZ_CHECK_ARG(smem <= 48 * 1024 && threads <= 256, "uz_bn_bwd_coop: C = %d too wide", C);
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_coop_kernel, threads, smem);
  if (e != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    uz::set_error("uz_bn_bwd_coop: occupancy query failed");
    return UZ_ERR_CUDA;
  }
  if (per_sm > 2) per_sm = 2;
  long long blocks = static_cast<long long>(per_sm) * uz::num_sms();
  const long long need = (npix + rows - 1) / rows;
  if (blocks > need) blocks = need;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(blocks), 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, bn_bwd_coop_kernel, p);
  (void)e;
  UZ_CHECK_LAUNCH("uz_bn_bwd_coop");
  return UZ_OK;
}
