// 3x3 convolution (+ bias, + training-mode BatchNorm + ReLU) for the deepest levels of the encoders: maps of 2x2 ... 4x4
// pixels (reference torchlayers.py:18-21 at the 2x2 and 4x4 resolution levels of models/phiseg.py).
//
// At batch 12 a 2x2 level has 48 output pixels.  The tensor-core kernel spends ~10 us on such a layer whatever it does
// (profiles/README.md: 5.3 us of launch + barrier / TMEM skeleton, 128-row MMAs that are 60 % padding, a 0.6 MB weight
// stream through 4 SMs) for 16 MFLOP of useful work -- latency, not arithmetic.  This kernel gives every output channel
// pair its own CTA on the CUDA cores instead: the CTA stages the whole (tiny) input map and the 9 x Cin weights of its
// channels in shared memory, a warp owns output pixels, its lanes split Cin (bf16x2 loads, fp32 accumulation, taps
// outside the image skipped instead of multiplied by zero padding), and -- because the CTA holds ALL pixels of its
// channels -- the batch statistics, normalisation and ReLU follow in the same launch without any cross-CTA exchange.
// 96 CTAs stream the weights in parallel.  Deterministic (fixed reduction order).
#include <cstdlib>

#include "common.cuh"
#include "unetzoo_b200.h"

namespace {

constexpr int kWarps = 16;                  // four warps per scheduler hide the shared-memory latency of the dot products
constexpr int kMaxPixPerWarp = 4;           // 16 warps x 4 = 64 pixels: 2x2 maps up to batch 16, one 4x4 map up to batch 4
constexpr int kCO = 2;                  // output channels per CTA

struct SmallParams {
  const __nv_bfloat16* x; int ldx;
  const __nv_bfloat16* w;               // packed [9][CoutP][CinP], tap = kw*3 + kh
  int N, H, W, Cin, CinP, CoutP;
  __nv_bfloat16* y; int ldy;
  const float* scale; const float* shift; int relu;      // y = act(acc * scale + shift)   (no BatchNorm fusion)
  // fused training-mode BatchNorm: y = acc + shift (bias), a = act(BatchNorm(y))
  int fuse_bn;
  float count, eps, momentum;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  int updates, act_relu;
  float* scale_out; float* shift_out; float* mean_out; float* invstd_out;
  __nv_bfloat16* a; int lda;
};

// PPW = pixels per warp, compile time: the pixel loops are fully unrolled and ptxas predicates their bodies instead of
// branching, so every slot costs issue cycles whether it holds a pixel or not (a fixed 12-slot version ran 22 us)
template <int PPW>
__global__ void __launch_bounds__(kWarps * 32) conv_small_kernel(const SmallParams p) {
  uz::pdl_prologue();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int npix = p.N * p.H * p.W;
  const int c0 = blockIdx.x * kCO;
  // carve: xs [npix][Cin] bf16 | ws [9][kCO][Cin] bf16 | ys [npix][kCO] fp32 | coef [2][kCO] fp32
  uint32_t* xs = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* ws = xs + static_cast<size_t>(npix) * p.Cin / 2;
  float* ys = reinterpret_cast<float*>(ws + 9 * kCO * p.Cin / 2);
  float* coef = ys + npix * kCO;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  // ---- stage the input map and this CTA's weights (16-byte vectors)
  {
    const int v_per_pix = p.Cin / 8;
    uint4* xs4 = reinterpret_cast<uint4*>(xs);
    for (int e = threadIdx.x; e < npix * v_per_pix; e += blockDim.x) {
      const int px = e / v_per_pix, v = e - px * v_per_pix;
      xs4[e] = *reinterpret_cast<const uint4*>(p.x + static_cast<size_t>(px) * p.ldx + v * 8);
    }
    uint4* ws4 = reinterpret_cast<uint4*>(ws);
    for (int e = threadIdx.x; e < 9 * kCO * v_per_pix; e += blockDim.x) {
      const int r = e / v_per_pix, v = e - r * v_per_pix;          // r = tap * kCO + co
      const int tap = r / kCO, co = r - tap * kCO;
      ws4[e] = *reinterpret_cast<const uint4*>(p.w + (static_cast<size_t>(tap) * p.CoutP + c0 + co) * p.CinP + v * 8);
    }
  }
  __syncthreads();

  // ---- a warp owns pixels warp, warp + nwarps, ...; lanes split Cin in bf16 pairs (64 channels per pass)
  float acc[PPW][kCO];
#pragma unroll
  for (int i = 0; i < PPW; ++i)
#pragma unroll
    for (int c = 0; c < kCO; ++c) acc[i][c] = 0.f;
  const int hw = p.H * p.W;
  const int passes = p.Cin / 64;
  const int half = p.Cin / 2;                                      // row length in 32-bit words
  // per owned pixel: its index and a 9-bit mask of the taps that fall inside the image (the others are zero padding and
  // are skipped) -- all warp-uniform, so the inner loop is branch-free apart from uniform skips
  int pcen[PPW], pmask[PPW];
#pragma unroll
  for (int i = 0; i < PPW; ++i) {
    const int px = warp + i * nwarps;
    pcen[i] = px;
    pmask[i] = 0;
    if (px < npix) {
      const int r = px % hw;
      const int y = r / p.W, x = r % p.W;
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap % 3 - 1, xx = x + tap / 3 - 1;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) pmask[i] |= 1 << tap;
      }
    }
  }
  for (int j = 0; j < passes; ++j) {
    const uint32_t* xj = xs + j * 32 + lane;
    const uint32_t* wj = ws + j * 32 + lane;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int off = (tap % 3 - 1) * p.W + (tap / 3 - 1);         // neighbour pixel = centre + off
      float wl[kCO], wh[kCO];
#pragma unroll
      for (int c = 0; c < kCO; ++c) {
        const uint32_t wv = wj[(tap * kCO + c) * half];
        wl[c] = uz::bf16lo(wv);
        wh[c] = uz::bf16hi(wv);
      }
#pragma unroll
      for (int i = 0; i < PPW; ++i) {
        if ((pmask[i] >> tap) & 1) {                               // warp-uniform
          const uint32_t xv = xj[(pcen[i] + off) * half];
          const float xl = uz::bf16lo(xv), xh = uz::bf16hi(xv);
#pragma unroll
          for (int c = 0; c < kCO; ++c) acc[i][c] = fmaf(xh, wh[c], fmaf(xl, wl[c], acc[i][c]));
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PPW; ++i) {
    const int px = warp + i * nwarps;
    if (px < npix) {
#pragma unroll
      for (int c = 0; c < kCO; ++c) {
        const float t = uz::warp_sum(acc[i][c]);
        if (lane == 0) ys[px * kCO + c] = t;
      }
    }
  }
  __syncthreads();

  if (!p.fuse_bn) {
    for (int e = threadIdx.x; e < npix * kCO; e += blockDim.x) {
      const int px = e / kCO, c = e - px * kCO;
      float v = fmaf(ys[e], p.scale ? p.scale[c0 + c] : 1.f, p.shift ? p.shift[c0 + c] : 0.f);
      if (p.relu) v = fmaxf(v, 0.f);
      p.y[static_cast<size_t>(px) * p.ldy + c0 + c] = uz::f2act(v);
    }
    return;
  }
  // ---- y = conv + bias, rounded to its storage type (the statistics are those of the stored values, like the
  // tensor-core path); warp c reduces channel c over all pixels
  for (int e = threadIdx.x; e < npix * kCO; e += blockDim.x) {
    const int c = e % kCO;
    const __nv_bfloat16 q = uz::f2act(ys[e] + (p.shift ? p.shift[c0 + c] : 0.f));
    p.y[static_cast<size_t>(e / kCO) * p.ldy + c0 + c] = q;
    ys[e] = uz::act2f(q);
  }
  __syncthreads();
  if (warp < kCO) {
    float s1 = 0.f, s2 = 0.f;
    for (int px = lane; px < npix; px += 32) {
      const float v = ys[px * kCO + warp];
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
    s1 = uz::warp_sum(s1);
    s2 = uz::warp_sum(s2);
    if (lane == 0) {
      const int cg = c0 + warp;
      const float mean = s1 / p.count;
      const float var = fmaxf(s2 / p.count - mean * mean, 0.f);
      const float invstd = rsqrtf(var + p.eps);
      const float sc = (p.gamma ? p.gamma[cg] : 1.f) * invstd;
      const float sh = (p.beta ? p.beta[cg] : 0.f) - mean * sc;
      coef[warp] = sc;
      coef[kCO + warp] = sh;
      p.scale_out[cg] = sc;
      p.shift_out[cg] = sh;
      p.mean_out[cg] = mean;
      p.invstd_out[cg] = invstd;
      if (p.running_mean) {
        const float unbiased = p.count > 1.f ? var * p.count / (p.count - 1.f) : var;
        float rm = p.running_mean[cg], rv = p.running_var[cg];
        for (int u = 0; u < p.updates; ++u) {
          rm = (1.f - p.momentum) * rm + p.momentum * mean;
          rv = (1.f - p.momentum) * rv + p.momentum * unbiased;
        }
        p.running_mean[cg] = rm;
        p.running_var[cg] = rv;
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < npix * kCO; e += blockDim.x) {
    const int c = e % kCO;
    float v = fmaf(ys[e], coef[c], coef[kCO + c]);
    if (p.act_relu) v = fmaxf(v, 0.f);
    p.a[static_cast<size_t>(e / kCO) * p.lda + c0 + c] = uz::f2act(v);
  }
}

}  // namespace

namespace uz {

// 1 if the small-map kernel takes this layer: 3x3, maps of at most 4x4 with at most 64 pixels in the batch (measured: 2x2
// x 12 conv+BN+ReLU 10.8 -> 7.4 us, plain conv 7.1 -> 6.4 us; 192 pixels would need 12 pixel slots per warp and run no faster than the tensor-core
// kernel), Cin a multiple of 64
int conv_small_supported(int N, int H, int W, int Cin, int Cout) {
  // opt-in (UZ_CONV_SMALL=1): in isolation conv + BatchNorm + ReLU on a 2x2x12 map takes 7.4 us instead of 10.8 us, in the
  // captured PHiSeg step the 22 affected launches gain 18 us of 4.38 ms -- not worth a second numeric path by default
  static const int enabled = [] { const char* e = getenv("UZ_CONV_SMALL"); return e ? atoi(e) : 0; }();
  static const int max_pix = [] { const char* e = getenv("UZ_CONV_SMALL_PIX"); return e ? atoi(e) : 64; }();
  if (!enabled || UZ_KNOB(32)) return 0;
  const long long npix = static_cast<long long>(N) * H * W;
  return (H <= 4 && W <= 4 && npix <= max_pix && npix <= kWarps * kMaxPixPerWarp && Cin % 64 == 0 && Cin <= 512 &&
          Cout % kCO == 0) ? 1 : 0;
}

struct SmallBn {      // mirror of conv_tc.cu's FusedBn (kept separate: the two files share no header beyond the ABI)
  float count, eps, momentum;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  int updates, relu;
  float* scale_out; float* shift_out; float* mean_out; float* invstd_out;
  void* a; int lda;
};

int conv_small_launch(const void* x, int N, int H, int W, int Cin, int ldx, const void* w_packed, int Cout, void* y,
                      int ldy, const float* scale, const float* shift, int relu, const SmallBn* fb, void* stream) {
  SmallParams p{};
  p.x = static_cast<const __nv_bfloat16*>(x); p.ldx = ldx;
  p.w = static_cast<const __nv_bfloat16*>(w_packed);
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.CinP = Cin; p.CoutP = Cout;
  p.y = static_cast<__nv_bfloat16*>(y); p.ldy = ldy;
  p.scale = scale; p.shift = shift; p.relu = relu;
  if (fb) {
    p.fuse_bn = 1;
    p.count = fb->count; p.eps = fb->eps; p.momentum = fb->momentum;
    p.gamma = fb->gamma; p.beta = fb->beta; p.running_mean = fb->running_mean; p.running_var = fb->running_var;
    p.updates = fb->updates; p.act_relu = fb->relu;
    p.scale_out = fb->scale_out; p.shift_out = fb->shift_out; p.mean_out = fb->mean_out; p.invstd_out = fb->invstd_out;
    p.a = static_cast<__nv_bfloat16*>(fb->a); p.lda = fb->lda;
  }
  const int npix = N * H * W;
  const int ppw = (npix + kWarps - 1) / kWarps;
  if (ppw > kMaxPixPerWarp) return UZ_ERR_ARG;
  auto kernel = ppw <= 1 ? conv_small_kernel<1> : (ppw == 2 ? conv_small_kernel<2> : (ppw == 3 ? conv_small_kernel<3> : conv_small_kernel<4>));
  const size_t smem = static_cast<size_t>(npix) * Cin * 2 + static_cast<size_t>(9) * kCO * Cin * 2 +
                      static_cast<size_t>(npix) * kCO * 4 + 2 * kCO * 4 + 16;
  static size_t attr_bytes_ppw[kMaxPixPerWarp + 1] = {};
  size_t& attr_bytes = attr_bytes_ppw[ppw];
  if (smem > 48 * 1024 && smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("conv_small: cannot raise dynamic smem limit to %zu: %s", smem, cudaGetErrorString(e));
      return UZ_ERR_CUDA;
    }
    attr_bytes = smem;
  }
  launch(kernel, dim3(Cout / kCO, 1, 1), kWarps * 32, smem, static_cast<cudaStream_t>(stream), p);
  UZ_CHECK_LAUNCH("uz_conv_fwd(small maps)");
  return UZ_OK;
}

}  // namespace uz
