// Memory-bound NHWC bf16 kernels around the tensor-core convolutions: weight packing, BatchNorm statistics
// finalisation / apply / backward, 2x2 average pooling, bilinear x2 upsampling (align_corners True and False) written
// straight into channel slices of concat buffers, strided channel copies, input packing with one-hot conditioning.
//
// All activations are [pixels][channels] with a pixel stride `ld` (elements, multiple of 8) so producers can write into
// and consumers can read from channel slices of wider (concat) buffers without copies.  Every thread moves 16-byte
// vectors (8 bf16 channels); consecutive threads touch consecutive 16-byte chunks of a pixel => fully coalesced.
#include <cstdlib>
#include "common.cuh"
#include "unetzoo_b200.h"

namespace {

constexpr int kEwThreads = 256;

// Grid of a grid-stride elementwise kernel.  `per_thread` work items per thread (env UZ_EW_PER_THREAD, default 16; measured single-stream step 8.54 / 8.38 / 8.28 / 7.90 ms for 1 / 4 / 8 / 16): the
// streaming loops keep four 16-byte loads in flight only if a thread HAS four items -- with one item per thread (the
// old sizing: up to 16 blocks per SM) every load was a dependent round trip to HBM and the 128^2-map BatchNorm passes
// ran at 0.2 of the HBM roofline (profiles/r02_ncu_kernels.md).  At least two blocks per SM while there is work for them.
int g_ew_per_thread = [] {
  const char* e = getenv("UZ_EW_PER_THREAD");
  const int v = e ? atoi(e) : 16;
  return v < 1 ? 1 : v;
}();
inline int ew_blocks(size_t work, int per_block = kEwThreads) {
  const size_t sms = static_cast<size_t>(uz::num_sms());
  size_t b1 = (work + per_block - 1) / per_block;                       // one item per thread
  size_t b = (work + static_cast<size_t>(per_block) * g_ew_per_thread - 1) / (static_cast<size_t>(per_block) * g_ew_per_thread);
  if (b < 2 * sms) b = b1 < 2 * sms ? b1 : 2 * sms;
  const size_t cap = sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// block size that is a multiple of the number of 8-channel chunks per pixel (<= 256 when possible)
inline int chunk_aligned_threads(int chunks) {
  if (chunks >= kEwThreads) return chunks > 1024 ? 1024 / 1 : chunks;
  return (kEwThreads / chunks) * chunks;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = uz::bf16lo(v.x); f[1] = uz::bf16hi(v.x); f[2] = uz::bf16lo(v.y); f[3] = uz::bf16hi(v.y);
  f[4] = uz::bf16lo(v.z); f[5] = uz::bf16hi(v.z); f[6] = uz::bf16lo(v.w); f[7] = uz::bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(uz::pack_bf16x2(f[0], f[1]), uz::pack_bf16x2(f[2], f[3]), uz::pack_bf16x2(f[4], f[5]),
                    uz::pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------- weight packing
// w fp32 [Cout][Cin][taps] (PyTorch OIHW flattened over kh,kw; source tap s = kh*3 + kw) ->
//   fwd : bf16 [taps][CoutP][CinP]            wp[t][o][i]  = w[o][i][src(t)]
//   dgrad: bf16 [taps][CinP2][CoutP2]         wd[t][i][o]  = w[o][i][src(taps-1-t)]   (flipped taps, transposed channels)
// Packed taps are dx-major, t = kw*3 + kh, so the three dy taps of one dx are contiguous (one TMA box in conv_tc2).
// 27 taps (OIDHW, source s = (kd*3 + kh)*3 + kw): packed t = (kd*3 + kw)*3 + kh, same idea with the z tap outermost.
__device__ __forceinline__ int src_tap(int t, int taps) {
  if (taps == 9) return (t % 3) * 3 + t / 3;
  if (taps == 27) return ((t / 9) * 3 + t % 3) * 3 + (t / 3) % 3;
  return t;
}
__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, __nv_bfloat16* wp,
                                   int CoutP, int CinP, __nv_bfloat16* wd, int CinP2, int CoutP2) {
  uz::pdl_prologue();
  const size_t n_fwd = static_cast<size_t>(taps) * CoutP * CinP;
  const size_t n_bwd = wd ? static_cast<size_t>(taps) * CinP2 * CoutP2 : 0;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n_fwd + n_bwd;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (idx < n_fwd) {
      const int i = idx % CinP;
      const int o = (idx / CinP) % CoutP;
      const int t = idx / (static_cast<size_t>(CinP) * CoutP);
      float v = (i < Cin && o < Cout) ? w[(static_cast<size_t>(o) * Cin + i) * taps + src_tap(t, taps)] : 0.f;
      wp[idx] = uz::f2act(v);
    } else {
      const size_t j = idx - n_fwd;
      const int o = j % CoutP2;
      const int i = (j / CoutP2) % CinP2;
      const int t = j / (static_cast<size_t>(CoutP2) * CinP2);
      float v = (i < Cin && o < Cout) ? w[(static_cast<size_t>(o) * Cin + i) * taps + src_tap(taps - 1 - t, taps)] : 0.f;
      wd[j] = uz::f2act(v);
    }
  }
}

// second phase of the packing kernels: tile[32 o][32 i x tg taps] (fp32, new weights) -> both packed bf16 copies, two
// channels per thread (4-byte stores; CoutP / CinP are multiples of 16 and tiles start at multiples of 32, so a pair never
// straddles the padded extent)
__device__ __forceinline__ void pack_tile_stores(const float (*tile)[32 * 9 + 1], int tg, int grp, int taps, int o0, int i0,
                                                 int CoutP, int CinP, __nv_bfloat16* wp, __nv_bfloat16* wd) {
  for (int e = threadIdx.x; e < tg * 512; e += kEwThreads) {
    const int tl = e >> 9, a = (e >> 4) & 31, b = (e & 15) * 2;
    const int sl = tg == 9 ? (tl % 3) * 3 + tl / 3 : 0;            // packed tap tl of the group <- source tap sl
    {   // forward copy: a = o, b = i (fastest)
      const int o = o0 + a, i = i0 + b;
      if (o < CoutP && i < CinP)
        *reinterpret_cast<uint32_t*>(wp + (static_cast<size_t>(grp * tg + tl) * CoutP + o) * CinP + i) =
            uz::pack_bf16x2(tile[a][b * tg + sl], tile[a][(b + 1) * tg + sl]);
    }
    if (wd) {   // dgrad copy: wd[taps-1-u][i][o] = w[o][i][src(u)], u = grp*tg + tl; a = i, b = o (fastest)
      const int i = i0 + a, o = o0 + b;
      if (o < CoutP && i < CinP)
        *reinterpret_cast<uint32_t*>(wd + (static_cast<size_t>(taps - 1 - (grp * tg + tl)) * CinP + i) * CoutP + o) =
            uz::pack_bf16x2(tile[b][a * tg + sl], tile[b + 1][a * tg + sl]);
    }
  }
}

// all conv layers of a model in ONE launch: blockIdx.y selects the layer descriptor (device table of UzPackDesc).
// A work item is a 32 (o) x 32 (i) x 9-tap tile staged through shared memory, so that the fp32 reads (contiguous runs
// of the OIHW source), the forward stores (i fastest) and the transposed dgrad stores (o fastest) are all coalesced.
// 27-tap kernels are three 9-tap groups (one per kd): both tap permutations map a group onto itself.
__global__ void __launch_bounds__(kEwThreads) pack_weight_batched_kernel(const UzPackDesc* __restrict__ descs) {
  uz::pdl_prologue();
  __shared__ float tile[32][32 * 9 + 1];
  const UzPackDesc d = descs[blockIdx.y];
  const float* __restrict__ w = static_cast<const float*>(d.w);
  __nv_bfloat16* wp = static_cast<__nv_bfloat16*>(d.w_fwd);
  __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(d.w_dgrad);
  const int tg = d.taps >= 9 ? 9 : 1;             // taps per group (taps is 1, 9 or 27)
  const int groups = d.taps / tg;
  const int ti = (d.CinP + 31) / 32, to = (d.CoutP + 31) / 32;
  const int total = to * ti * groups;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int grp = item % groups;
    const int r = item / groups;
    const int i0 = (r % ti) * 32, o0 = (r / ti) * 32;
    const int row = 32 * tg;
    for (int e = threadIdx.x; e < 32 * row; e += kEwThreads) {
      const int oo = e / row, rem = e - oo * row;
      const int ii = rem / tg, sl = rem - ii * tg;
      const int o = o0 + oo, i = i0 + ii;
      tile[oo][rem] = (o < d.Cout && i < d.Cin) ? w[(static_cast<size_t>(o) * d.Cin + i) * d.taps + grp * tg + sl] : 0.f;
    }
    __syncthreads();
    pack_tile_stores(tile, tg, grp, d.taps, o0, i0, d.CoutP, d.CinP, wp, wd);
    __syncthreads();
  }
}

// ---------------------------------------------------------------- Adam, all parameters in one launch
// torch.optim.Adam(lr, betas, eps, weight_decay as L2 on the gradient) -- the optimizer of the reference
// (train_model.py:49) -- over a device table of (param, grad, exp_avg, exp_avg_sq, numel) and a chunk table that maps
// each block to (tensor, 4096-element chunk).  The step counter lives on the device so the launch is graph-capturable.
__global__ void adam_count_kernel(const UzAdamDesc* __restrict__ descs, int n) {
  uz::pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) descs[i].step[0] += 1.f;
}

constexpr int kAdamChunk = 4096;
__global__ void __launch_bounds__(256) adam_batched_kernel(const UzAdamDesc* __restrict__ descs,
                                                           const int* __restrict__ chunk_table, double lr_d,
                                                           double beta1_d, double beta2_d, float eps,
                                                           float weight_decay) {
  uz::pdl_prologue();
  const int tensor = chunk_table[2 * blockIdx.x], chunk = chunk_table[2 * blockIdx.x + 1];
  const UzAdamDesc d = descs[tensor];
  float* __restrict__ p = static_cast<float*>(d.p);
  const float* __restrict__ g = static_cast<const float*>(d.g);
  float* __restrict__ m = static_cast<float*>(d.m);
  float* __restrict__ v = static_cast<float*>(d.v);
  // bias corrections in double like torch's host-side computation (1 - beta^t cancels badly in fp32 for small t)
  // and 1 - beta in double before rounding to fp32, which is what torch's Python scalars do)
  const double t = static_cast<double>(d.step[0]);
  const double bc1 = 1.0 - pow(beta1_d, t);
  const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(beta2_d, t)));
  const float step_size = static_cast<float>(lr_d / bc1);
  const float beta2 = static_cast<float>(beta2_d);
  const float omb1 = static_cast<float>(1.0 - beta1_d), omb2 = static_cast<float>(1.0 - beta2_d);
  const long long lo = static_cast<long long>(chunk) * kAdamChunk;
  const long long hi = min(lo + kAdamChunk, d.n);
  const bool vec = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0);
  auto update = [&](float& pp, float gg, float& mm, float& vv) {
    gg = fmaf(weight_decay, pp, gg);
    mm = fmaf(omb1, gg - mm, mm);                         // lerp(exp_avg, grad, 1 - beta1)
    vv = fmaf(beta2, vv, omb2 * gg * gg);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * mm / denom;
  };
  if (vec && hi - lo == kAdamChunk) {
#pragma unroll
    for (int k = 0; k < kAdamChunk / (256 * 4); ++k) {
      const long long i = lo + (k * 256 + threadIdx.x) * 4;
      float4 pp = *reinterpret_cast<const float4*>(p + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<const float4*>(m + i);
      float4 vv = *reinterpret_cast<const float4*>(v + i);
      update(pp.x, gg.x, mm.x, vv.x);
      update(pp.y, gg.y, mm.y, vv.y);
      update(pp.z, gg.z, mm.z, vv.z);
      update(pp.w, gg.w, mm.w, vv.w);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
      float pp = p[i], mm = m[i], vv = v[i];
      update(pp, g[i], mm, vv);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  }
}

// ---------------------------------------------------------------- Adam + weight packing in one pass (conv weights)
// The tensor-core kernels read bf16 copies of the conv weights in two tap-major layouts (pack_weight_batched_kernel above).
// Re-packing after every optimizer step was a pass of its own at the head of the next step: 100 MB fp32 read again,
// 100 MB bf16 written, alone on the GPU.  Here the Adam update of a 32 (o) x 32 (i) x 9-tap tile keeps the new fp32
// weights in shared memory and writes both packed copies from there: the parameters are read once per step.
// item_table: int [nitems][2] = (row of the pack table, tile index inside the layer).
template <int MINB>
__global__ void __launch_bounds__(kEwThreads, MINB) adam_pack_kernel(const UzAdamDesc* __restrict__ descs,
                                                               const UzAdamPackDesc* __restrict__ packs,
                                                               const int* __restrict__ item_table, double lr_d,
                                                               double beta1_d, double beta2_d, float eps,
                                                               float weight_decay) {
  uz::pdl_prologue();
  __shared__ float tile[32][32 * 9 + 1];
  const UzAdamPackDesc pk = packs[item_table[2 * blockIdx.x]];
  const int item = item_table[2 * blockIdx.x + 1];
  const UzAdamDesc d = descs[pk.tensor];
  float* __restrict__ p = static_cast<float*>(d.p);
  const float* __restrict__ g = static_cast<const float*>(d.g);
  float* __restrict__ m = static_cast<float*>(d.m);
  float* __restrict__ v = static_cast<float*>(d.v);
  __nv_bfloat16* wp = static_cast<__nv_bfloat16*>(pk.w_fwd);
  __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(pk.w_dgrad);
  const double t = static_cast<double>(d.step[0]);
  const double bc1 = 1.0 - pow(beta1_d, t);
  const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(beta2_d, t)));
  const float step_size = static_cast<float>(lr_d / bc1);
  const float beta2 = static_cast<float>(beta2_d);
  const float omb1 = static_cast<float>(1.0 - beta1_d), omb2 = static_cast<float>(1.0 - beta2_d);
  auto update = [&](float& pp, float gg, float& mm, float& vv) {       // identical to adam_batched_kernel
    gg = fmaf(weight_decay, pp, gg);
    mm = fmaf(omb1, gg - mm, mm);
    vv = fmaf(beta2, vv, omb2 * gg * gg);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * mm / denom;
  };
  const int tg = pk.taps >= 9 ? 9 : 1;
  const int groups = pk.taps / tg;
  const int ti = (pk.CinP + 31) / 32;
  const int grp = item % groups;
  const int r = item / groups;
  const int i0 = (r % ti) * 32, o0 = (r / ti) * 32;
  const int row = 32 * tg;
  // Gradient straight from the weight-gradient kernels' split-K slabs [split][tap][CoutP][CinP] (pk.slab != NULL): the tile
  // of gradients is summed over the splits and transposed into shared memory first -- rows of 32 input channels, 128
  // contiguous bytes per (tap, o) -- so the separate reduction pass (read slabs, write OIHW, read again here) disappears.
  const float* __restrict__ slab = pk.slab;
  if (slab != nullptr) {
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const size_t plane = static_cast<size_t>(pk.CoutP) * pk.CinP;
    const size_t split_stride = plane * pk.taps;
    const int i = i0 + lane;
    constexpr int U = 9;                                           // independent 128-byte row loads in flight per warp
    constexpr int W = kEwThreads / 32;
    for (int rt0 = wrp; rt0 < 32 * tg; rt0 += W * U) {             // (o row, tap of the group) pairs
      float a[U];
      int oo[U], sl[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rt = rt0 + u * W;
        oo[u] = rt / tg;
        sl[u] = rt - oo[u] * tg;
        a[u] = 0.f;
        const int o = o0 + oo[u];
        if (rt < 32 * tg && o < pk.Cout && i < pk.Cin) {
          const float* src = slab + static_cast<size_t>(grp * tg + sl[u]) * plane + static_cast<size_t>(o) * pk.CinP + i;
          float t = src[0];
          for (int sp = 1; sp < pk.splits; ++sp) t += src[static_cast<size_t>(sp) * split_stride];   // fixed order
          a[u] = t;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (rt0 + u * W < 32 * tg) tile[oo[u]][lane * tg + sl[u]] = a[u];
    }
    __syncthreads();
  }
  const bool full = o0 + 32 <= pk.Cout && i0 + 32 <= pk.Cin && tg == 9 && pk.taps == 9 && (pk.Cin & 3) == 0 &&
                    (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0);
  if (full) {
    // a row of the tile = 288 contiguous, 16-byte aligned floats of the OIHW tensor: 72 float4 per row, 2304 per tile,
    // 9 per thread in three batches of 3 x 4 independent 16-byte loads
    constexpr int U = 3;
    for (int e0 = threadIdx.x; e0 < 32 * 72; e0 += kEwThreads * U) {
      size_t idx[U];
      int oo[U], c4[U];
      float4 pp[U], gg[U], mm[U], vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * kEwThreads;
        oo[u] = e / 72;
        c4[u] = e - oo[u] * 72;
        idx[u] = (static_cast<size_t>(o0 + oo[u]) * pk.Cin + i0) * 9 + c4[u] * 4;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pp[u] = *reinterpret_cast<const float4*>(p + idx[u]);
        if (slab != nullptr) {
          const float* trow = &tile[oo[u]][c4[u] * 4];
          gg[u] = make_float4(trow[0], trow[1], trow[2], trow[3]);
        } else {
          gg[u] = *reinterpret_cast<const float4*>(g + idx[u]);
        }
        mm[u] = *reinterpret_cast<const float4*>(m + idx[u]);
        vv[u] = *reinterpret_cast<const float4*>(v + idx[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        update(pp[u].x, gg[u].x, mm[u].x, vv[u].x);
        update(pp[u].y, gg[u].y, mm[u].y, vv[u].y);
        update(pp[u].z, gg[u].z, mm[u].z, vv[u].z);
        update(pp[u].w, gg[u].w, mm[u].w, vv[u].w);
        *reinterpret_cast<float4*>(p + idx[u]) = pp[u];
        *reinterpret_cast<float4*>(m + idx[u]) = mm[u];
        *reinterpret_cast<float4*>(v + idx[u]) = vv[u];
        float* trow = &tile[oo[u]][c4[u] * 4];
        trow[0] = pp[u].x; trow[1] = pp[u].y; trow[2] = pp[u].z; trow[3] = pp[u].w;
      }
    }
  } else {
    constexpr int U = 4;
    for (int e0 = threadIdx.x; e0 < 32 * row; e0 += kEwThreads * U) {
      size_t idx[U];
      bool ok[U];
      int oo[U], rem[U];
      float pp[U], gg[U], mm[U], vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * kEwThreads;
        oo[u] = e / row;
        rem[u] = e - oo[u] * row;
        const int ii = rem[u] / tg, sl = rem[u] - ii * tg;
        const int o = o0 + oo[u], i = i0 + ii;
        ok[u] = e < 32 * row && o < pk.Cout && i < pk.Cin;
        idx[u] = (static_cast<size_t>(o) * pk.Cin + i) * pk.taps + grp * tg + sl;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ok[u]) {
          pp[u] = p[idx[u]];
          gg[u] = slab != nullptr ? tile[oo[u]][rem[u]] : g[idx[u]];
          mm[u] = m[idx[u]];
          vv[u] = v[idx[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ok[u]) {
          update(pp[u], gg[u], mm[u], vv[u]);
          p[idx[u]] = pp[u]; m[idx[u]] = mm[u]; v[idx[u]] = vv[u];
        }
        if (e0 + u * kEwThreads < 32 * row) tile[oo[u]][rem[u]] = ok[u] ? pp[u] : 0.f;
      }
    }
  }
  __syncthreads();
  pack_tile_stores(tile, tg, grp, pk.taps, o0, i0, pk.CoutP, pk.CinP, wp, wd);
}

// ---------------------------------------------------------------- BatchNorm statistics
// partial [tiles][2][C] (sum, sumsq per tile, written by the conv epilogue) -> per-channel scale/shift, saved
// mean/invstd, running-stat update (momentum, unbiased variance) -- reference torchlayers.py:20, SURVEY Appendix A.
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int tiles, int C, float count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                                   float* mean_out, float* invstd_out) {
  uz::pdl_prologue();
  // one warp per channel; lanes stride over tiles; fixed order => deterministic
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= C) return;
  double s = 0.0, q = 0.0;
  for (int t = lane; t < tiles; t += 32) {
    s += static_cast<double>(partial[(static_cast<size_t>(t) * 2) * C + warp]);
    q += static_cast<double>(partial[(static_cast<size_t>(t) * 2 + 1) * C + warp]);
  }
  s = uz::warp_sum_d(s);
  q = uz::warp_sum_d(q);
  if (lane == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float g = gamma ? gamma[warp] : 1.f;
    const float b = beta ? beta[warp] : 0.f;
    const float sc = g * invstd;
    scale[warp] = sc;
    shift[warp] = b - static_cast<float>(mean) * sc;
    if (mean_out) mean_out[warp] = static_cast<float>(mean);
    if (invstd_out) invstd_out[warp] = invstd;
    if (running_mean) {
      const double unbiased = count > 1.f ? var * count / (count - 1.0) : var;
      running_mean[warp] = (1.f - momentum) * running_mean[warp] + momentum * static_cast<float>(mean);
      running_var[warp] = (1.f - momentum) * running_var[warp] + momentum * static_cast<float>(unbiased);
    }
  }
}

// eval-mode fold: y = relu(conv*scale + shift) with running statistics (train_model.py:139 net.eval()).
__global__ void bn_eval_fold_kernel(const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ rm,
                                    const float* __restrict__ rv, float eps, int C, float* scale, float* shift) {
  uz::pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = (gamma ? gamma[c] : 1.f) * rsqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) + ((conv_bias ? conv_bias[c] : 0.f) - rm[c]) * sc;
}

// out = act(y * scale[c] + shift[c])
__global__ void affine_act_kernel(const __nv_bfloat16* __restrict__ y, int ldy, const float* __restrict__ scale,
                                  const float* __restrict__ shift, int relu, __nv_bfloat16* out, int ldo, size_t npix,
                                  int C) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const size_t total = npix * chunks;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = idx / chunks;
    const int c0 = static_cast<int>(idx - pix * chunks) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(y + pix * ldy + c0);
    float f[8];
    unpack8(v, f);
    const float4 s0 = *reinterpret_cast<const float4*>(scale + c0), s1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
    const float4 h0 = *reinterpret_cast<const float4*>(shift + c0), h1 = *reinterpret_cast<const float4*>(shift + c0 + 4);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (relu) f[j] = fmaxf(f[j], 0.f);
    }
    *reinterpret_cast<uint4*>(out + pix * ldo + c0) = pack8(f);
  }
}

// Training-mode BatchNorm normalise + ReLU straight from the conv's statistics accumulators (sum, sumsq per channel):
// every block derives scale/shift for all channels into shared memory (C rsqrt per block), block 0 also publishes them
// (saved for backward), updates the running statistics (momentum, unbiased variance) -- no separate finalize launch.
__global__ void bn_apply_train_kernel(const __nv_bfloat16* __restrict__ y, int ldy, const float* __restrict__ sums,
                                      float count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float eps, float momentum, float* running_mean, float* running_var,
                                      float* scale_out, float* shift_out, float* mean_out, float* invstd_out, int relu,
                                      __nv_bfloat16* out, int ldo, size_t npix, int C,
                                      const __nv_bfloat16* __restrict__ res, int ldr, float res_sign, int updates) {
  // res != nullptr: out = res + res_sign * act(...) (additive coupling of a reversible block written straight into the
  // block output); updates: momentum updates of the running statistics (2 for reversible blocks: their F and G run a
  // second time in backward with the same batch, torchlayers.py:71-78 via revtorch)
  uz::pdl_prologue();
  extern __shared__ float sm[];          // [2][C]: scale, shift
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = sums[c] / count;
    float var = sums[C + c] / count - mean * mean;
    var = fmaxf(var, 0.f);
    const float invstd = rsqrtf(var + eps);
    const float sc = (gamma ? gamma[c] : 1.f) * invstd;
    const float sh = (beta ? beta[c] : 0.f) - mean * sc;
    sm[c] = sc;
    sm[C + c] = sh;
    if (blockIdx.x == 0) {
      scale_out[c] = sc;
      shift_out[c] = sh;
      mean_out[c] = mean;
      invstd_out[c] = invstd;
      if (running_mean) {
        const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
        float rm = running_mean[c], rv = running_var[c];
        for (int u = 0; u < updates; ++u) {
          rm = (1.f - momentum) * rm + momentum * mean;
          rv = (1.f - momentum) * rv + momentum * unbiased;
        }
        running_mean[c] = rm;
        running_var[c] = rv;
      }
    }
  }
  __syncthreads();
  // blockDim.x is a multiple of C/8 (launcher): a thread keeps its 8-channel chunk for the whole loop, so the
  // coefficients live in registers and the loop has no index division
  const int chunks = C / 8;
  const int c0 = (threadIdx.x % chunks) * 8;
  const size_t prow = blockDim.x / chunks;
  const size_t pstride = gridDim.x * prow;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = sm[c0 + j]; sh[j] = sm[C + c0 + j]; }
  // four independent 16-byte loads in flight per thread before the first use (HBM latency x bandwidth needs ~64 KB per SM)
  size_t pix = blockIdx.x * prow + threadIdx.x / chunks;
  if (res == nullptr) {
    for (; pix + 3 * pstride < npix; pix += 4 * pstride) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const uint4*>(y + (pix + u * pstride) * ldy + c0));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaf(f[j], sc[j], sh[j]);
          if (relu) f[j] = fmaxf(f[j], 0.f);
        }
        *reinterpret_cast<uint4*>(out + (pix + u * pstride) * ldo + c0) = pack8(f);
      }
    }
  }
  for (; pix < npix; pix += pstride) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(y + pix * ldy + c0), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (relu) f[j] = fmaxf(f[j], 0.f);
    }
    if (res != nullptr) {
      float r[8];
      unpack8(*reinterpret_cast<const uint4*>(res + pix * ldr + c0), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(res_sign, f[j], r[j]);
    }
    *reinterpret_cast<uint4*>(out + pix * ldo + c0) = pack8(f);
  }
}

// BN+ReLU backward pass 2 straight from the accumulators of pass 1 (sum g, sum g*y): coefficients per block in shared
// memory, block 0 publishes dgamma / dbeta.   dy = A*g + B*y + Cc  (see bn_bwd_finalize_kernel for the algebra)
__global__ void bn_bwd_apply_train_kernel(const __nv_bfloat16* __restrict__ dout, int ldd,
                                          const __nv_bfloat16* __restrict__ y, int ldy, const float* __restrict__ scale,
                                          const float* __restrict__ shift, int relu, const float* __restrict__ sums,
                                          float count, const float* __restrict__ gamma, const float* __restrict__ mean,
                                          const float* __restrict__ invstd, float* dgamma, float* dbeta,
                                          __nv_bfloat16* dy, int lddy, size_t npix, int C) {
  uz::pdl_prologue();
  extern __shared__ float sm[];          // [5][C]: A, B, Cc, scale, shift
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sg = sums[c], sgy = sums[C + c];
    const float mu = mean[c], is = invstd[c], g = gamma ? gamma[c] : 1.f;
    const float sgx = (sgy - mu * sg) * is;
    const float mg = sg / count, mgx = sgx / count;
    sm[c] = g * is;
    sm[C + c] = -g * is * is * mgx;
    sm[2 * C + c] = -g * is * mg + g * is * is * mgx * mu;
    sm[3 * C + c] = scale[c];
    sm[4 * C + c] = shift[c];
    if (blockIdx.x == 0) {
      if (dgamma) dgamma[c] = sgx;
      if (dbeta) dbeta[c] = sg;
    }
  }
  __syncthreads();
  // blockDim.x is a multiple of C/8 (launcher): per-thread coefficients in registers, no index division in the loop
  const int chunks = C / 8;
  const int c0 = (threadIdx.x % chunks) * 8;
  const size_t prow = blockDim.x / chunks;
  const size_t pstride = gridDim.x * prow;
  float ca[8], cb[8], cc[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ca[j] = sm[c0 + j]; cb[j] = sm[C + c0 + j]; cc[j] = sm[2 * C + c0 + j];
    sc[j] = sm[3 * C + c0 + j]; sh[j] = sm[4 * C + c0 + j];
  }
  auto one = [&](const uint4& vg, const uint4& vy, size_t px) {
    float g[8], yy[8];
    unpack8(vg, g);
    unpack8(vy, yy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = (!relu || fmaf(yy[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
      g[j] = fmaf(ca[j], m, fmaf(cb[j], yy[j], cc[j]));
    }
    *reinterpret_cast<uint4*>(dy + px * lddy + c0) = pack8(g);
  };
  size_t pix = blockIdx.x * prow + threadIdx.x / chunks;
  for (; pix + 3 * pstride < npix; pix += 4 * pstride) {        // eight independent 16-byte loads in flight per thread
    uint4 vg[4], vy[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vg[u] = *reinterpret_cast<const uint4*>(dout + (pix + u * pstride) * ldd + c0);
      vy[u] = *reinterpret_cast<const uint4*>(y + (pix + u * pstride) * ldy + c0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(vg[u], vy[u], pix + u * pstride);
  }
  for (; pix < npix; pix += pstride)
    one(*reinterpret_cast<const uint4*>(dout + pix * ldd + c0), *reinterpret_cast<const uint4*>(y + pix * ldy + c0), pix);
}

// BN+ReLU backward, pass 1: per-channel sum(g) and sum(g * y) with g = dout * [y*scale+shift > 0]; partial per block.
// block = (C/8 channel chunks) x rows; grid-stride over pixel rows; block result -> partial[block][2][C].
__global__ void bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, const __nv_bfloat16* __restrict__ y,
                                     int ldy, const float* __restrict__ scale, const float* __restrict__ shift,
                                     int relu, size_t npix, int C, float* partial, int atomic_out,
                                     const __nv_bfloat16* __restrict__ inv_in, int ldi, __nv_bfloat16* inv_out, int ldo) {
  // inv_in != nullptr: also inv_out = inv_in - act(y*scale+shift): the inverse of a reversible block's additive coupling
  // (x2 = y2 - G(y1)) computed in the pass that reads the recomputed y anyway
  uz::pdl_prologue();
  extern __shared__ float red[];  // [rows][2][C]
  const int chunks = C / 8;
  const int rows = blockDim.x / chunks;
  const int r = threadIdx.x / chunks;
  const int c0 = (threadIdx.x - r * chunks) * 8;
  float sg[8], sgy[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sg[j] = 0.f; sgy[j] = 0.f; }
  if (r < rows) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; }
    auto one = [&](const uint4& vg, const uint4& vy) {
      float g[8], yy[8];
      unpack8(vg, g);
      unpack8(vy, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float m = (!relu || fmaf(yy[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
        sg[j] += m;
        sgy[j] = fmaf(m, yy[j], sgy[j]);
      }
    };
    const size_t pstride = static_cast<size_t>(gridDim.x) * rows;
    size_t pix = static_cast<size_t>(blockIdx.x) * rows + r;
    if (inv_in == nullptr) {
      for (; pix + 3 * pstride < npix; pix += 4 * pstride) {      // eight independent 16-byte loads in flight per thread
        uint4 vg[4], vy[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          vg[u] = *reinterpret_cast<const uint4*>(dout + (pix + u * pstride) * ldd + c0);
          vy[u] = *reinterpret_cast<const uint4*>(y + (pix + u * pstride) * ldy + c0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) one(vg[u], vy[u]);
      }
    }
    for (; pix < npix; pix += pstride) {
      const uint4 vy = *reinterpret_cast<const uint4*>(y + pix * ldy + c0);
      one(*reinterpret_cast<const uint4*>(dout + pix * ldd + c0), vy);
      if (inv_in != nullptr) {
        float yy[8], r[8];
        unpack8(vy, yy);
        unpack8(*reinterpret_cast<const uint4*>(inv_in + pix * ldi + c0), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = fmaf(yy[j], sc[j], sh[j]);
          if (relu) a = fmaxf(a, 0.f);
          r[j] -= a;
        }
        *reinterpret_cast<uint4*>(inv_out + pix * ldo + c0) = pack8(r);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[(r * 2) * C + c0 + j] = sg[j];
      red[(r * 2 + 1) * C + c0 + j] = sgy[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int rr = 0; rr < rows; ++rr) acc += red[rr * 2 * C + i];
    if (atomic_out) atomicAdd(partial + i, acc);
    else partial[static_cast<size_t>(blockIdx.x) * 2 * C + i] = acc;
  }
}

// pass 1b: reduce partials -> coefficients of dy = A*g + B*y + Cc, plus dgamma/dbeta (fp32, PyTorch grad layout).
//   xhat = (y - mean) * invstd ; dgamma = sum(g*xhat) ; dbeta = sum(g)
//   dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat))
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, float count,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, float* coefA, float* coefB, float* coefC,
                                       float* dgamma, float* dbeta) {
  uz::pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= C) return;
  double sg = 0.0, sgy = 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    sg += static_cast<double>(partial[static_cast<size_t>(b) * 2 * C + warp]);
    sgy += static_cast<double>(partial[static_cast<size_t>(b) * 2 * C + C + warp]);
  }
  sg = uz::warp_sum_d(sg);
  sgy = uz::warp_sum_d(sgy);
  if (lane == 0) {
    const double mu = mean[warp], is = invstd[warp], g = gamma ? gamma[warp] : 1.0;
    const double sgx = (sgy - mu * sg) * is;  // sum g*xhat
    const double mg = sg / count, mgx = sgx / count;
    coefA[warp] = static_cast<float>(g * is);
    coefB[warp] = static_cast<float>(-g * is * is * mgx);
    coefC[warp] = static_cast<float>(-g * is * mg + g * is * is * mgx * mu);
    if (dgamma) dgamma[warp] = static_cast<float>(sgx);
    if (dbeta) dbeta[warp] = static_cast<float>(sg);
  }
}

// pass 2: dy = A*g + B*y + Cc   (A,B,Cc may be null => plain ReLU backward dy = g)
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, const __nv_bfloat16* __restrict__ y,
                                    int ldy, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                    const float* __restrict__ coefA, const float* __restrict__ coefB,
                                    const float* __restrict__ coefC, __nv_bfloat16* dy, int lddy, size_t npix, int C) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const size_t total = npix * chunks;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = idx / chunks;
    const int c0 = static_cast<int>(idx - pix * chunks) * 8;
    float g[8], yy[8];
    unpack8(*reinterpret_cast<const uint4*>(dout + pix * ldd + c0), g);
    unpack8(*reinterpret_cast<const uint4*>(y + pix * ldy + c0), yy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float m = (!relu || fmaf(yy[j], scale[c], shift[c]) > 0.f) ? g[j] : 0.f;
      g[j] = coefA ? fmaf(coefA[c], m, fmaf(coefB[c], yy[j], coefC[c])) : m;
    }
    *reinterpret_cast<uint4*>(dy + pix * lddy + c0) = pack8(g);
  }
}

// ---------------------------------------------------------------- 2x2 average pooling (AvgPool2d(2,2,ceil_mode), even sizes)
__global__ void avgpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* out, int ldo, int N,
                                    int Ho, int Wo, int C) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * chunks;
  const int W = Wo * 2;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % chunks) * 8;
    const size_t opix = idx / chunks;
    const int xo = opix % Wo;
    const int yo = (opix / Wo) % Ho;
    const size_t n = opix / (static_cast<size_t>(Wo) * Ho);
    const size_t ipix = (n * (Ho * 2) + yo * 2) * W + xo * 2;
    float a[8], b[8], c[8], d[8];
    unpack8(*reinterpret_cast<const uint4*>(x + ipix * ldx + c0), a);
    unpack8(*reinterpret_cast<const uint4*>(x + (ipix + 1) * ldx + c0), b);
    unpack8(*reinterpret_cast<const uint4*>(x + (ipix + W) * ldx + c0), c);
    unpack8(*reinterpret_cast<const uint4*>(x + (ipix + W + 1) * ldx + c0), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.25f * ((a[j] + b[j]) + (c[j] + d[j]));
    *reinterpret_cast<uint4*>(out + opix * ldo + c0) = pack8(a);
  }
}

// dx[n, y, x, c] = 0.25 * dout[n, y/2, x/2, c]  (optionally accumulating into dx: the pooled tensor's source may also
// feed a skip connection whose gradient is already in dx)
__global__ void avgpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, __nv_bfloat16* dx, int ldx, int N,
                                    int Ho, int Wo, int C, int accumulate) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const int H = Ho * 2, W = Wo * 2;
  const size_t total = static_cast<size_t>(N) * H * W * chunks;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % chunks) * 8;
    const size_t ipix = idx / chunks;
    const int xi = ipix % W;
    const int yi = (ipix / W) % H;
    const size_t n = ipix / (static_cast<size_t>(W) * H);
    const size_t opix = (n * Ho + yi / 2) * Wo + xi / 2;
    float g[8];
    unpack8(*reinterpret_cast<const uint4*>(dout + opix * ldd + c0), g);
    if (accumulate) {
      float o[8];
      unpack8(*reinterpret_cast<const uint4*>(dx + ipix * ldx + c0), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = fmaf(0.25f, g[j], o[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= 0.25f;
    }
    *reinterpret_cast<uint4*>(dx + ipix * ldx + c0) = pack8(g);
  }
}

// ---------------------------------------------------------------- bilinear x2
// align_corners=True : src = dst * (in-1)/(out-1)              (reference models/phiseg.py:66,213-216,305-309)
// align_corners=False: src = max((dst+0.5)/2 - 0.5, 0)         (reference models/unet.py:67)
__device__ __forceinline__ void up2_src(int o, int in, int align, int& i0, int& i1, float& w1) {
  float s;
  if (align) {
    s = in > 1 ? o * (static_cast<float>(in - 1) / static_cast<float>(2 * in - 1)) : 0.f;
  } else {
    s = fmaxf((o + 0.5f) * 0.5f - 0.5f, 0.f);
  }
  i0 = static_cast<int>(s);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  w1 = s - i0;
}

// one block per output row (n, yo): the row's source rows / weight are computed once, threads walk (xo, 8-channel
// chunk) with 32-bit arithmetic.  (The first version decoded a flat 64-bit index per element -- five 64-bit divisions
// per 16-byte store: 720 us of a 5.7 ms GED-100 evaluation, 0.14 of the HBM roofline.)
__global__ void up2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* out, int ldo, int N, int h,
                               int w, int C, int align) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const int H = 2 * h, W = 2 * w;
  const int rows = N * H, per_row = W * chunks;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / H, yo = row - n * H;
    int y0, y1;
    float wy;
    up2_src(yo, h, align, y0, y1, wy);
    const __nv_bfloat16* r0 = x + (static_cast<size_t>(n) * h + y0) * w * ldx;
    const __nv_bfloat16* r1 = x + (static_cast<size_t>(n) * h + y1) * w * ldx;
    __nv_bfloat16* o = out + static_cast<size_t>(row) * W * ldo;
    for (int i = threadIdx.x; i < per_row; i += blockDim.x) {
      const int xo = i / chunks, c0 = (i - xo * chunks) * 8;
      int x0, x1;
      float wx;
      up2_src(xo, w, align, x0, x1, wx);
      float a[8], b[8], c[8], d[8];
      unpack8(*reinterpret_cast<const uint4*>(r0 + static_cast<size_t>(x0) * ldx + c0), a);
      unpack8(*reinterpret_cast<const uint4*>(r0 + static_cast<size_t>(x1) * ldx + c0), b);
      unpack8(*reinterpret_cast<const uint4*>(r1 + static_cast<size_t>(x0) * ldx + c0), c);
      unpack8(*reinterpret_cast<const uint4*>(r1 + static_cast<size_t>(x1) * ldx + c0), d);
      const float w00 = (1.f - wy) * (1.f - wx), w01 = (1.f - wy) * wx, w10 = wy * (1.f - wx), w11 = wy * wx;
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = w00 * a[j] + w01 * b[j] + w10 * c[j] + w11 * d[j];
      *reinterpret_cast<uint4*>(o + static_cast<size_t>(xo) * ldo + c0) = pack8(a);
    }
  }
}

// gather form of the transpose: every input pixel collects from the output pixels that read it
__global__ void up2_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, __nv_bfloat16* dx, int ldx, int N, int h,
                               int w, int C, int align) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const int H = 2 * h, W = 2 * w;
  const int rows = N * h, per_row = w * chunks;
  for (int row = blockIdx.x; row < rows; row += gridDim.x)
  for (int i = threadIdx.x; i < per_row; i += blockDim.x) {
    const size_t n = row / h;
    const int yi = row - static_cast<int>(n) * h;
    const int xi = i / chunks, c0 = (i - xi * chunks) * 8;
    const size_t ipix = static_cast<size_t>(row) * w + xi;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int ylo = max(2 * yi - 2, 0), yhi = min(2 * yi + 3, H - 1);
    const int xlo = max(2 * xi - 2, 0), xhi = min(2 * xi + 3, W - 1);
    for (int yo = ylo; yo <= yhi; ++yo) {
      int y0, y1; float wy;
      up2_src(yo, h, align, y0, y1, wy);
      float cy = 0.f;
      if (y0 == yi) cy += 1.f - wy;
      if (y1 == yi) cy += wy;
      if (cy == 0.f) continue;
      for (int xo = xlo; xo <= xhi; ++xo) {
        int x0, x1; float wx;
        up2_src(xo, w, align, x0, x1, wx);
        float cx = 0.f;
        if (x0 == xi) cx += 1.f - wx;
        if (x1 == xi) cx += wx;
        if (cx == 0.f) continue;
        float g[8];
        unpack8(*reinterpret_cast<const uint4*>(dout + ((n * H + yo) * W + xo) * ldd + c0), g);
        const float cw = cy * cx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(cw, g[j], acc[j]);
      }
    }
    *reinterpret_cast<uint4*>(dx + ipix * ldx + c0) = pack8(acc);
  }
}

// ---------------------------------------------------------------- volumes: 2x2x2 average pooling, trilinear x2
// AvgPool3d(2,2,ceil_mode) on even sizes (reference models/phiseg3D.py:100); NDHWC, same vector scheme as the 2-D kernels.
__global__ void avgpool3_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* out, int ldo, int N,
                                    int Do, int Ho, int Wo, int C) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const size_t total = static_cast<size_t>(N) * Do * Ho * Wo * chunks;
  const size_t W = Wo * 2, HW = static_cast<size_t>(Ho * 2) * W;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % chunks) * 8;
    const size_t opix = idx / chunks;
    const int xo = opix % Wo;
    const int yo = (opix / Wo) % Ho;
    const int zo = (opix / (static_cast<size_t>(Wo) * Ho)) % Do;
    const size_t n = opix / (static_cast<size_t>(Wo) * Ho * Do);
    const size_t ipix = ((n * (Do * 2) + zo * 2) * (Ho * 2) + yo * 2) * W + xo * 2;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a[8];
      const size_t off = (k & 1) + ((k >> 1) & 1) * W + (k >> 2) * HW;
      unpack8(*reinterpret_cast<const uint4*>(x + (ipix + off) * ldx + c0), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.125f;
    *reinterpret_cast<uint4*>(out + opix * ldo + c0) = pack8(acc);
  }
}

__global__ void avgpool3_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, __nv_bfloat16* dx, int ldx, int N,
                                    int Do, int Ho, int Wo, int C) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const int D = Do * 2, H = Ho * 2, W = Wo * 2;
  const size_t total = static_cast<size_t>(N) * D * H * W * chunks;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % chunks) * 8;
    const size_t ipix = idx / chunks;
    const int xi = ipix % W;
    const int yi = (ipix / W) % H;
    const int zi = (ipix / (static_cast<size_t>(W) * H)) % D;
    const size_t n = ipix / (static_cast<size_t>(W) * H * D);
    const size_t opix = ((n * Do + zi / 2) * Ho + yi / 2) * Wo + xi / 2;
    float g[8];
    unpack8(*reinterpret_cast<const uint4*>(dout + opix * ldd + c0), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= 0.125f;
    *reinterpret_cast<uint4*>(dx + ipix * ldx + c0) = pack8(g);
  }
}

// trilinear x2, align_corners=True (reference models/phiseg3D.py:143,291-294,382-386)
__global__ void up3_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* out, int ldo, int N, int d,
                               int h, int w, int C) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const size_t total = static_cast<size_t>(N) * D * H * W * chunks;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % chunks) * 8;
    const size_t opix = idx / chunks;
    const int xo = opix % W;
    const int yo = (opix / W) % H;
    const int zo = (opix / (static_cast<size_t>(W) * H)) % D;
    const size_t n = opix / (static_cast<size_t>(W) * H * D);
    int xs[2], ys[2], zs[2];
    float wx, wy, wz;
    up2_src(xo, w, 1, xs[0], xs[1], wx);
    up2_src(yo, h, 1, ys[0], ys[1], wy);
    up2_src(zo, d, 1, zs[0], zs[1], wz);
    const size_t base = n * d * h * w;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int kx = k & 1, ky = (k >> 1) & 1, kz = k >> 2;
      const float cw = (kx ? wx : 1.f - wx) * (ky ? wy : 1.f - wy) * (kz ? wz : 1.f - wz);
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(
                  x + (base + (static_cast<size_t>(zs[kz]) * h + ys[ky]) * w + xs[kx]) * ldx + c0), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(cw, a[j], acc[j]);
    }
    *reinterpret_cast<uint4*>(out + opix * ldo + c0) = pack8(acc);
  }
}

// gather form of the transpose, separable weights
__global__ void up3_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, __nv_bfloat16* dx, int ldx, int N, int d,
                               int h, int w, int C) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const size_t total = static_cast<size_t>(N) * d * h * w * chunks;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % chunks) * 8;
    const size_t ipix = idx / chunks;
    const int xi = ipix % w;
    const int yi = (ipix / w) % h;
    const int zi = (ipix / (static_cast<size_t>(w) * h)) % d;
    const size_t n = ipix / (static_cast<size_t>(w) * h * d);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int zlo = max(2 * zi - 2, 0), zhi = min(2 * zi + 3, D - 1);
    const int ylo = max(2 * yi - 2, 0), yhi = min(2 * yi + 3, H - 1);
    const int xlo = max(2 * xi - 2, 0), xhi = min(2 * xi + 3, W - 1);
    for (int zo = zlo; zo <= zhi; ++zo) {
      int z0, z1; float wz;
      up2_src(zo, d, 1, z0, z1, wz);
      float cz = 0.f;
      if (z0 == zi) cz += 1.f - wz;
      if (z1 == zi) cz += wz;
      if (cz == 0.f) continue;
      for (int yo = ylo; yo <= yhi; ++yo) {
        int y0, y1; float wy;
        up2_src(yo, h, 1, y0, y1, wy);
        float cy = 0.f;
        if (y0 == yi) cy += 1.f - wy;
        if (y1 == yi) cy += wy;
        if (cy == 0.f) continue;
        for (int xo = xlo; xo <= xhi; ++xo) {
          int x0, x1; float wx;
          up2_src(xo, w, 1, x0, x1, wx);
          float cx = 0.f;
          if (x0 == xi) cx += 1.f - wx;
          if (x1 == xi) cx += wx;
          if (cx == 0.f) continue;
          float g[8];
          unpack8(*reinterpret_cast<const uint4*>(dout + (((n * D + zo) * H + yo) * W + xo) * ldd + c0), g);
          const float cw = cz * cy * cx;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(cw, g[j], acc[j]);
        }
      }
    }
    *reinterpret_cast<uint4*>(dx + ipix * ldx + c0) = pack8(acc);
  }
}

// ---------------------------------------------------------------- strided channel copy / add (concat, split, grad sum)
__global__ void copy_channels_kernel(const __nv_bfloat16* __restrict__ src, int lds, __nv_bfloat16* dst, int ldd,
                                     size_t npix, int C, int accumulate, size_t src_period) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const size_t total = npix * chunks;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = idx / chunks;
    const int c0 = static_cast<int>(idx - pix * chunks) * 8;
    const size_t spix = src_period ? pix % src_period : pix;     // broadcast of one image over a batch of copies
    uint4 v = *reinterpret_cast<const uint4*>(src + spix * lds + c0);
    if (accumulate) {          // 1: dst += src, 2: dst -= src
      float a[8], b[8];
      unpack8(v, a);
      unpack8(*reinterpret_cast<const uint4*>(dst + pix * ldd + c0), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = accumulate == 2 ? b[j] - a[j] : b[j] + a[j];
      v = pack8(a);
    }
    *reinterpret_cast<uint4*>(dst + pix * ldd + c0) = v;
  }
}

// out = a + b (sign >= 0) or a - b (sign < 0) on channel slices: the additive coupling of the reversible blocks and its
// inverse in one pass (y1 = x1 + F(x2), x2 = y2 - G(y1)) instead of a copy followed by an accumulate
__global__ void add_channels_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b,
                                    int ldb, __nv_bfloat16* out, int ldo, size_t npix, int C, int sign) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const size_t total = npix * chunks;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = idx / chunks;
    const int c0 = static_cast<int>(idx - pix * chunks) * 8;
    float x[8], y[8];
    unpack8(*reinterpret_cast<const uint4*>(a + pix * lda + c0), x);
    unpack8(*reinterpret_cast<const uint4*>(b + pix * ldb + c0), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = sign < 0 ? x[j] - y[j] : x[j] + y[j];
    *reinterpret_cast<uint4*>(out + pix * ldo + c0) = pack8(x);
  }
}

// ---------------------------------------------------------------- global spatial mean (ProbUNet Gaussian heads)
// out[b][c] = mean over the hw pixels of x[b][:][c]  (torch.mean over H then W, probabilistic_unet.py:114-115)
__global__ void global_mean_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int hw, int C,
                                       __nv_bfloat16* out, int ldo) {
  uz::pdl_prologue();
  // block = (sample, 64-channel group); threads = 32 channel pairs x 8 pixel lanes
  __shared__ float red[8][64];
  const int b = blockIdx.x, cg = blockIdx.y * 64;
  const int cp = threadIdx.x & 31, pl = threadIdx.x >> 5;
  float s0 = 0.f, s1 = 0.f;
  const int c = cg + cp * 2;
  if (c < C) {
    for (int p = pl; p < hw; p += 8) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(x + (static_cast<size_t>(b) * hw + p) * ldx + c);
      s0 += uz::bf16lo(v);
      s1 += uz::bf16hi(v);
    }
  }
  red[pl][cp * 2] = s0;
  red[pl][cp * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64 && cg + threadIdx.x < C) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    out[static_cast<size_t>(b) * ldo + cg + threadIdx.x] = uz::f2act(t / hw);
  }
}
__global__ void global_mean_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int ldd, int hw, int C,
                                       __nv_bfloat16* dx, int ldx, size_t npix) {
  uz::pdl_prologue();
  const int chunks = C / 8;
  const size_t total = npix * chunks;
  const float inv = 1.f / hw;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = idx / chunks;
    const int c0 = static_cast<int>(idx - pix * chunks) * 8;
    float g[8];
    unpack8(*reinterpret_cast<const uint4*>(dout + (pix / hw) * ldd + c0), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= inv;
    *reinterpret_cast<uint4*>(dx + pix * ldx + c0) = pack8(g);
  }
}

// ---------------------------------------------------------------- input packing
// patch fp32 NCHW [B,Cimg,H,W] (+ mask float [B,1,H,W] holding integer labels) -> bf16 NHWC [B,H,W,CP]:
//   channels [0,Cimg) image, [Cimg, Cimg+nlabels) = (mask==k) - 0.5 (reference models/phiseg.py:176-183,
//   utils.py:289-311), rest zero.
__global__ void input_pack_kernel(const float* __restrict__ patch, const float* __restrict__ mask, int B, int Cimg,
                                  int H, int W, int nlabels, __nv_bfloat16* out, int CP) {
  uz::pdl_prologue();
  const size_t hw = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * hw;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = pix / hw, r = pix - b * hw;
    __nv_bfloat16* o = out + pix * CP;
    int c = 0;
    for (; c < Cimg; ++c) o[c] = uz::f2act(patch[(b * Cimg + c) * hw + r]);
    if (mask) {
      const float m = mask[b * hw + r];
      for (int k = 0; k < nlabels; ++k, ++c) o[c] = uz::f2act((m == static_cast<float>(k) ? 1.f : 0.f) - 0.5f);
    }
    for (; c < CP; ++c) o[c] = uz::f2act(0.f);
  }
}

// fp32 NCHW <-> bf16 NHWC (module-boundary conversions; channel padding zero-filled)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int B, int C, size_t hw, __nv_bfloat16* dst, int ld) {
  uz::pdl_prologue();
  const size_t total = static_cast<size_t>(B) * hw * ld;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = idx % ld;
    const size_t pix = idx / ld;
    const size_t b = pix / hw, r = pix - b * hw;
    dst[idx] = uz::f2act(c < C ? src[(b * C + c) * hw + r] : 0.f);
  }
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int ld, int B, int C, size_t hw, float* dst) {
  uz::pdl_prologue();
  const size_t total = static_cast<size_t>(B) * C * hw;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = idx % hw;
    const int c = (idx / hw) % C;
    const size_t b = idx / (hw * C);
    dst[idx] = uz::act2f(src[(b * hw + r) * ld + c]);
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int uz_pack_conv_weight(const float* w, int Cout, int Cin, int taps, void* w_fwd, int CoutP, int CinP,
                                   void* w_dgrad, int CinP2, int CoutP2, void* stream) {
  UZ_CHECK_ARG(w && w_fwd, "uz_pack_conv_weight: null pointer");
  UZ_CHECK_ARG(CoutP >= Cout && CinP >= Cin, "uz_pack_conv_weight: padded dims smaller than logical dims");
  UZ_CHECK_ARG(!w_dgrad || (CinP2 >= Cin && CoutP2 >= Cout), "uz_pack_conv_weight: bad dgrad dims");
  const size_t n = static_cast<size_t>(taps) * CoutP * CinP + (w_dgrad ? static_cast<size_t>(taps) * CinP2 * CoutP2 : 0);
  uz::launch(pack_weight_kernel, ew_blocks(n), kEwThreads, 0, ST(stream), w, Cout, Cin, taps, static_cast<__nv_bfloat16*>(w_fwd),
                                                                 CoutP, CinP, static_cast<__nv_bfloat16*>(w_dgrad),
                                                                 CinP2, CoutP2);
  UZ_CHECK_LAUNCH("uz_pack_conv_weight");
  return UZ_OK;
}

extern "C" int uz_pack_conv_weights_batched(const void* descs_device, int n, int blocks_per_layer, void* stream) {
  UZ_CHECK_ARG(descs_device && n > 0 && blocks_per_layer > 0, "uz_pack_conv_weights_batched: bad arguments");
  dim3 grid(blocks_per_layer, n, 1);
  uz::launch(pack_weight_batched_kernel, grid, kEwThreads, 0, ST(stream), static_cast<const UzPackDesc*>(descs_device));
  UZ_CHECK_LAUNCH("uz_pack_conv_weights_batched");
  return UZ_OK;
}

namespace {
// eval-mode BatchNorm folds of ALL layers of a model in one launch (blockIdx.y = layer): see bn_eval_fold_kernel
__global__ void bn_eval_fold_batched_kernel(const UzFoldDesc* __restrict__ descs, float eps) {
  uz::pdl_prologue();
  const UzFoldDesc d = descs[blockIdx.y];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d.C; c += gridDim.x * blockDim.x) {
    const float sc = (d.gamma ? d.gamma[c] : 1.f) * rsqrtf(d.running_var[c] + eps);
    d.scale[c] = sc;
    d.shift[c] = (d.beta ? d.beta[c] : 0.f) + ((d.conv_bias ? d.conv_bias[c] : 0.f) - d.running_mean[c]) * sc;
  }
}
}  // namespace

extern "C" int uz_bn_eval_fold_batched(const void* descs_device, int n, int max_channels, float eps, void* stream) {
  UZ_CHECK_ARG(descs_device && n > 0 && max_channels > 0, "uz_bn_eval_fold_batched: bad arguments");
  dim3 grid((max_channels + 127) / 128, n, 1);
  uz::launch(bn_eval_fold_batched_kernel, grid, 128, 0, ST(stream), static_cast<const UzFoldDesc*>(descs_device), eps);
  UZ_CHECK_LAUNCH("uz_bn_eval_fold_batched");
  return UZ_OK;
}

extern "C" int uz_adam_chunk_elems(void) { return kAdamChunk; }

extern "C" int uz_adam_step_batched(const void* descs_device, int ntensors, const int* chunk_table_device, int nchunks,
                                    double lr, double beta1, double beta2, double eps, double weight_decay,
                                    void* stream) {
  UZ_CHECK_ARG(descs_device && chunk_table_device && ntensors > 0 && nchunks > 0, "uz_adam_step_batched: bad arguments");
  uz::launch(adam_count_kernel, (ntensors + 255) / 256, 256, 0, ST(stream), static_cast<const UzAdamDesc*>(descs_device),
             ntensors);
  uz::launch(adam_batched_kernel, nchunks, 256, 0, ST(stream), static_cast<const UzAdamDesc*>(descs_device),
             chunk_table_device, lr, beta1, beta2, static_cast<float>(eps), static_cast<float>(weight_decay));
  UZ_CHECK_LAUNCH("uz_adam_step_batched");
  return UZ_OK;
}

extern "C" int uz_adam_pack_items(int CoutP, int CinP, int taps) {
  return ((CoutP + 31) / 32) * ((CinP + 31) / 32) * (taps >= 9 ? taps / 9 : 1);
}

// Adam for all tensors of `descs_device`: the step counters of ALL of them are incremented, tensors covered by the chunk
// table get the plain update (uz_adam_step_batched's kernel), conv weights listed in `packs_device` get the update AND
// their bf16 forward / dgrad copies re-packed from the new values in the same pass.
extern "C" int uz_adam_pack_step(const void* descs_device, int ntensors, const int* chunk_table_device, int nchunks,
                                 const void* packs_device, const int* item_table_device, int nitems, double lr,
                                 double beta1, double beta2, double eps, double weight_decay, void* stream) {
  UZ_CHECK_ARG(descs_device && ntensors > 0 && nchunks >= 0 && nitems >= 0, "uz_adam_pack_step: bad arguments");
  UZ_CHECK_ARG((nchunks == 0 || chunk_table_device) && (nitems == 0 || (packs_device && item_table_device)),
               "uz_adam_pack_step: null table");
  uz::launch(adam_count_kernel, (ntensors + 255) / 256, 256, 0, ST(stream), static_cast<const UzAdamDesc*>(descs_device),
             ntensors);
  UZ_CHECK_LAUNCH("uz_adam_pack_step(count)");
  if (nchunks > 0) {
    uz::launch(adam_batched_kernel, nchunks, 256, 0, ST(stream), static_cast<const UzAdamDesc*>(descs_device),
               chunk_table_device, lr, beta1, beta2, static_cast<float>(eps), static_cast<float>(weight_decay));
    UZ_CHECK_LAUNCH("uz_adam_pack_step(adam)");
  }
  if (nitems > 0) {
    // three resident blocks per SM (80 registers, a few spilled words) instead of two: 4.111 -> 4.098 ms per step
    static const int minb = [] { const char* e = getenv("UZ_ADAM_PACK_MINB"); return e ? atoi(e) : 3; }();
    auto kernel = minb >= 4 ? adam_pack_kernel<4> : (minb == 3 ? adam_pack_kernel<3> : adam_pack_kernel<2>);
    uz::launch(kernel, nitems, kEwThreads, 0, ST(stream), static_cast<const UzAdamDesc*>(descs_device),
               static_cast<const UzAdamPackDesc*>(packs_device), item_table_device, lr, beta1, beta2,
               static_cast<float>(eps), static_cast<float>(weight_decay));
    UZ_CHECK_LAUNCH("uz_adam_pack_step(adam+pack)");
  }
  return UZ_OK;
}

extern "C" int uz_bn_finalize(const float* partial, int tiles, int C, float count, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                              float* scale, float* shift, float* mean_out, float* invstd_out, void* stream) {
  UZ_CHECK_ARG(partial && scale && shift && tiles > 0 && C > 0, "uz_bn_finalize: bad arguments");
  const int threads = 256;
  const int blocks = (C * 32 + threads - 1) / threads;
  uz::launch(bn_finalize_kernel, blocks, threads, 0, ST(stream), partial, tiles, C, count, gamma, beta, eps, momentum,
                                                         running_mean, running_var, scale, shift, mean_out, invstd_out);
  UZ_CHECK_LAUNCH("uz_bn_finalize");
  return UZ_OK;
}

extern "C" int uz_bn_eval_fold(const float* conv_bias, const float* gamma, const float* beta, const float* running_mean,
                               const float* running_var, float eps, int C, float* scale, float* shift, void* stream) {
  UZ_CHECK_ARG(running_mean && running_var && scale && shift && C > 0, "uz_bn_eval_fold: bad arguments");
  uz::launch(bn_eval_fold_kernel, (C + 127) / 128, 128, 0, ST(stream), conv_bias, gamma, beta, running_mean, running_var, eps, C,
                                                              scale, shift);
  UZ_CHECK_LAUNCH("uz_bn_eval_fold");
  return UZ_OK;
}

extern "C" int uz_affine_act(const void* y, int ldy, const float* scale, const float* shift, int relu, void* out,
                             int ldo, long long npix, int C, void* stream) {
  UZ_CHECK_ARG(y && out && scale && shift, "uz_affine_act: null pointer");
  UZ_CHECK_ARG(C % 8 == 0 && ldy % 8 == 0 && ldo % 8 == 0 && aligned16(y) && aligned16(out), "uz_affine_act: alignment");
  if (npix == 0) return UZ_OK;
  uz::launch(affine_act_kernel, ew_blocks(static_cast<size_t>(npix) * (C / 8)), kEwThreads, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(y), ldy, scale, shift, relu, static_cast<__nv_bfloat16*>(out), ldo,
      static_cast<size_t>(npix), C);
  UZ_CHECK_LAUNCH("uz_affine_act");
  return UZ_OK;
}

extern "C" int uz_bn_bwd_num_blocks(long long npix, int C) {
  const int chunks = C / 8;
  int threads = 256;
  if (chunks > threads) threads = ((chunks + 31) / 32) * 32;
  const int rows = threads / chunks;
  long long b = (npix + rows - 1) / rows;
  const long long cap = static_cast<long long>(uz::num_sms()) * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

extern "C" int uz_bn_bwd_reduce(const void* dout, int ldd, const void* y, int ldy, const float* scale,
                                const float* shift, int relu, long long npix, int C, float* partial, void* stream) {
  UZ_CHECK_ARG(dout && y && scale && shift && partial, "uz_bn_bwd_reduce: null pointer");
  UZ_CHECK_ARG(C % 8 == 0 && ldd % 8 == 0 && ldy % 8 == 0, "uz_bn_bwd_reduce: alignment");
  const int chunks = C / 8;
  int threads = 256;
  if (chunks > threads) threads = ((chunks + 31) / 32) * 32;
  const int rows = threads / chunks;
  const int blocks = uz_bn_bwd_num_blocks(npix, C);
  const size_t smem = static_cast<size_t>(rows) * 2 * C * sizeof(float);
  uz::launch(bn_bwd_reduce_kernel, blocks, threads, smem, ST(stream), static_cast<const __nv_bfloat16*>(dout), ldd,
                                                             static_cast<const __nv_bfloat16*>(y), ldy, scale, shift,
                                                             relu, static_cast<size_t>(npix), C, partial, 0,
                                                             static_cast<const __nv_bfloat16*>(nullptr), 0,
                                                             static_cast<__nv_bfloat16*>(nullptr), 0);
  UZ_CHECK_LAUNCH("uz_bn_bwd_reduce");
  return UZ_OK;
}

extern "C" int uz_bn_bwd_reduce_sums(const void* dout, int ldd, const void* y, int ldy, const float* scale,
                                     const float* shift, int relu, long long npix, int C, float* sums, void* stream) {
  return uz_bn_bwd_reduce_sums_ex(dout, ldd, y, ldy, scale, shift, relu, npix, C, sums, nullptr, 0, nullptr, 0, stream);
}

extern "C" int uz_bn_bwd_reduce_sums_ex(const void* dout, int ldd, const void* y, int ldy, const float* scale,
                                        const float* shift, int relu, long long npix, int C, float* sums,
                                        const void* inv_in, int ld_inv_in, void* inv_out, int ld_inv_out, void* stream) {
  UZ_CHECK_ARG(dout && y && scale && shift && sums, "uz_bn_bwd_reduce_sums: null pointer");
  UZ_CHECK_ARG((inv_in == nullptr) == (inv_out == nullptr) &&
                   (!inv_in || (ld_inv_in % 8 == 0 && ld_inv_out % 8 == 0 && aligned16(inv_in) && aligned16(inv_out))),
               "uz_bn_bwd_reduce_sums_ex: bad inverse-coupling operands");
  UZ_CHECK_ARG(C % 8 == 0 && ldd % 8 == 0 && ldy % 8 == 0, "uz_bn_bwd_reduce_sums: alignment");
  const int chunks = C / 8;
  int threads = 256;
  if (chunks > threads) threads = ((chunks + 31) / 32) * 32;
  const int rows = threads / chunks;
  const int blocks = uz_bn_bwd_num_blocks(npix, C);          // <= 4 per SM; one atomic per block per channel
  const size_t smem = static_cast<size_t>(rows) * 2 * C * sizeof(float);
  uz::launch(bn_bwd_reduce_kernel, blocks, threads, smem, ST(stream), static_cast<const __nv_bfloat16*>(dout), ldd,
                                                             static_cast<const __nv_bfloat16*>(y), ldy, scale, shift,
                                                             relu, static_cast<size_t>(npix), C, sums, 1,
                                                             static_cast<const __nv_bfloat16*>(inv_in), ld_inv_in,
                                                             static_cast<__nv_bfloat16*>(inv_out), ld_inv_out);
  UZ_CHECK_LAUNCH("uz_bn_bwd_reduce_sums");
  return UZ_OK;
}

extern "C" int uz_bn_apply_train(const void* y, int ldy, const float* sums, float count, const float* gamma,
                                 const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                 float* scale_out, float* shift_out, float* mean_out, float* invstd_out, int relu,
                                 void* out, int ldo, long long npix, int C, void* stream) {
  return uz_bn_apply_train_ex(y, ldy, sums, count, gamma, beta, eps, momentum, running_mean, running_var, scale_out,
                              shift_out, mean_out, invstd_out, relu, out, ldo, npix, C, nullptr, 0, 1, 1, stream);
}

extern "C" int uz_bn_apply_train_ex(const void* y, int ldy, const float* sums, float count, const float* gamma,
                                    const float* beta, float eps, float momentum, float* running_mean,
                                    float* running_var, float* scale_out, float* shift_out, float* mean_out,
                                    float* invstd_out, int relu, void* out, int ldo, long long npix, int C,
                                    const void* residual, int ld_res, int res_sign, int stat_updates, void* stream) {
  UZ_CHECK_ARG(y && sums && scale_out && shift_out && mean_out && invstd_out && out, "uz_bn_apply_train: null pointer");
  UZ_CHECK_ARG(!residual || (ld_res % 8 == 0 && aligned16(residual)), "uz_bn_apply_train_ex: bad residual operand");
  UZ_CHECK_ARG(stat_updates >= 0 && stat_updates <= 4, "uz_bn_apply_train_ex: stat_updates %d", stat_updates);
  UZ_CHECK_ARG(C % 8 == 0 && C <= 8192 && ldy % 8 == 0 && ldo % 8 == 0 && npix > 0, "uz_bn_apply_train: bad arguments");
  const int threads = chunk_aligned_threads(C / 8);
  uz::launch(bn_apply_train_kernel, ew_blocks(static_cast<size_t>(npix) * (C / 8), threads), threads, 2 * C * sizeof(float), ST(stream), static_cast<const __nv_bfloat16*>(y), ldy, sums, count, gamma, beta, eps,
                                        momentum, running_mean, running_var, scale_out, shift_out, mean_out, invstd_out,
                                        relu, static_cast<__nv_bfloat16*>(out), ldo, static_cast<size_t>(npix), C,
                                        static_cast<const __nv_bfloat16*>(residual), ld_res, res_sign < 0 ? -1.f : 1.f,
                                        stat_updates);
  UZ_CHECK_LAUNCH("uz_bn_apply_train");
  return UZ_OK;
}

extern "C" int uz_bn_bwd_apply_train(const void* dout, int ldd, const void* y, int ldy, const float* scale,
                                     const float* shift, int relu, const float* sums, float count, const float* gamma,
                                     const float* mean, const float* invstd, float* dgamma, float* dbeta, void* dy,
                                     int lddy, long long npix, int C, void* stream) {
  UZ_CHECK_ARG(dout && y && scale && shift && sums && mean && invstd && dy, "uz_bn_bwd_apply_train: null pointer");
  UZ_CHECK_ARG(C % 8 == 0 && C <= 8192 && ldd % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0 && npix > 0,
               "uz_bn_bwd_apply_train: bad arguments");
  const int threads = chunk_aligned_threads(C / 8);
  uz::launch(bn_bwd_apply_train_kernel, ew_blocks(static_cast<size_t>(npix) * (C / 8), threads), threads, 5 * C * sizeof(float), ST(stream), static_cast<const __nv_bfloat16*>(dout), ldd,
                                            static_cast<const __nv_bfloat16*>(y), ldy, scale, shift, relu, sums, count,
                                            gamma, mean, invstd, dgamma, dbeta, static_cast<__nv_bfloat16*>(dy), lddy,
                                            static_cast<size_t>(npix), C);
  UZ_CHECK_LAUNCH("uz_bn_bwd_apply_train");
  return UZ_OK;
}

extern "C" int uz_bn_bwd_finalize(const float* partial, int nblocks, int C, float count, const float* gamma,
                                  const float* mean, const float* invstd, float* coefA, float* coefB, float* coefC,
                                  float* dgamma, float* dbeta, void* stream) {
  UZ_CHECK_ARG(partial && mean && invstd && coefA && coefB && coefC, "uz_bn_bwd_finalize: null pointer");
  const int threads = 256;
  const int blocks = (C * 32 + threads - 1) / threads;
  uz::launch(bn_bwd_finalize_kernel, blocks, threads, 0, ST(stream), partial, nblocks, C, count, gamma, mean, invstd, coefA,
                                                             coefB, coefC, dgamma, dbeta);
  UZ_CHECK_LAUNCH("uz_bn_bwd_finalize");
  return UZ_OK;
}

extern "C" int uz_bn_bwd_apply(const void* dout, int ldd, const void* y, int ldy, const float* scale,
                               const float* shift, int relu, const float* coefA, const float* coefB,
                               const float* coefC, void* dy, int lddy, long long npix, int C, void* stream) {
  UZ_CHECK_ARG(dout && y && dy && scale && shift, "uz_bn_bwd_apply: null pointer");
  UZ_CHECK_ARG(C % 8 == 0 && ldd % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0, "uz_bn_bwd_apply: alignment");
  if (npix == 0) return UZ_OK;
  uz::launch(bn_bwd_apply_kernel, ew_blocks(static_cast<size_t>(npix) * (C / 8)), kEwThreads, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(dout), ldd, static_cast<const __nv_bfloat16*>(y), ldy, scale, shift, relu, coefA,
      coefB, coefC, static_cast<__nv_bfloat16*>(dy), lddy, static_cast<size_t>(npix), C);
  UZ_CHECK_LAUNCH("uz_bn_bwd_apply");
  return UZ_OK;
}

extern "C" int uz_avgpool2_fwd(const void* x, int ldx, void* out, int ldo, int N, int Ho, int Wo, int C, void* stream) {
  UZ_CHECK_ARG(x && out && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "uz_avgpool2_fwd: bad arguments");
  uz::launch(avgpool2_fwd_kernel, ew_blocks(static_cast<size_t>(N) * Ho * Wo * (C / 8)), kEwThreads, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, N, Ho, Wo, C);
  UZ_CHECK_LAUNCH("uz_avgpool2_fwd");
  return UZ_OK;
}

extern "C" int uz_avgpool2_bwd(const void* dout, int ldd, void* dx, int ldx, int N, int Ho, int Wo, int C,
                               int accumulate, void* stream) {
  UZ_CHECK_ARG(dout && dx && C % 8 == 0 && ldx % 8 == 0 && ldd % 8 == 0, "uz_avgpool2_bwd: bad arguments");
  uz::launch(avgpool2_bwd_kernel, ew_blocks(static_cast<size_t>(N) * Ho * Wo * 4 * (C / 8)), kEwThreads, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(dout), ldd, static_cast<__nv_bfloat16*>(dx), ldx, N, Ho, Wo, C, accumulate);
  UZ_CHECK_LAUNCH("uz_avgpool2_bwd");
  return UZ_OK;
}

extern "C" int uz_upsample2x_fwd(const void* x, int ldx, void* out, int ldo, int N, int h, int w, int C,
                                 int align_corners, void* stream) {
  UZ_CHECK_ARG(x && out && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "uz_upsample2x_fwd: bad arguments");
  UZ_CHECK_ARG(static_cast<long long>(N) * 2 * h < (1ll << 31), "uz_upsample2x_fwd: too many rows");
  const int rows_f = N * 2 * h;
  uz::launch(up2_fwd_kernel, rows_f < uz::num_sms() * 16 ? rows_f : uz::num_sms() * 16,
             2 * w * (C / 8) >= 256 ? 256 : ((2 * w * (C / 8) + 31) / 32) * 32, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, N, h, w, C, align_corners);
  UZ_CHECK_LAUNCH("uz_upsample2x_fwd");
  return UZ_OK;
}

extern "C" int uz_upsample2x_bwd(const void* dout, int ldd, void* dx, int ldx, int N, int h, int w, int C,
                                 int align_corners, void* stream) {
  UZ_CHECK_ARG(dout && dx && C % 8 == 0 && ldx % 8 == 0 && ldd % 8 == 0, "uz_upsample2x_bwd: bad arguments");
  UZ_CHECK_ARG(static_cast<long long>(N) * h < (1ll << 31), "uz_upsample2x_bwd: too many rows");
  const int rows_b = N * h;
  uz::launch(up2_bwd_kernel, rows_b < uz::num_sms() * 16 ? rows_b : uz::num_sms() * 16,
             w * (C / 8) >= 256 ? 256 : ((w * (C / 8) + 31) / 32) * 32, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(dout), ldd, static_cast<__nv_bfloat16*>(dx), ldx, N, h, w, C, align_corners);
  UZ_CHECK_LAUNCH("uz_upsample2x_bwd");
  return UZ_OK;
}

extern "C" int uz_avgpool3_fwd(const void* x, int ldx, void* out, int ldo, int N, int Do, int Ho, int Wo, int C,
                               void* stream) {
  UZ_CHECK_ARG(x && out && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "uz_avgpool3_fwd: bad arguments");
  uz::launch(avgpool3_fwd_kernel, ew_blocks(static_cast<size_t>(N) * Do * Ho * Wo * (C / 8)), kEwThreads, 0, ST(stream),
             static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, N, Do, Ho, Wo, C);
  UZ_CHECK_LAUNCH("uz_avgpool3_fwd");
  return UZ_OK;
}

extern "C" int uz_avgpool3_bwd(const void* dout, int ldd, void* dx, int ldx, int N, int Do, int Ho, int Wo, int C,
                               void* stream) {
  UZ_CHECK_ARG(dout && dx && C % 8 == 0 && ldx % 8 == 0 && ldd % 8 == 0, "uz_avgpool3_bwd: bad arguments");
  uz::launch(avgpool3_bwd_kernel, ew_blocks(static_cast<size_t>(N) * Do * Ho * Wo * 8 * (C / 8)), kEwThreads, 0, ST(stream),
             static_cast<const __nv_bfloat16*>(dout), ldd, static_cast<__nv_bfloat16*>(dx), ldx, N, Do, Ho, Wo, C);
  UZ_CHECK_LAUNCH("uz_avgpool3_bwd");
  return UZ_OK;
}

extern "C" int uz_upsample3d_fwd(const void* x, int ldx, void* out, int ldo, int N, int d, int h, int w, int C,
                                 void* stream) {
  UZ_CHECK_ARG(x && out && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "uz_upsample3d_fwd: bad arguments");
  uz::launch(up3_fwd_kernel, ew_blocks(static_cast<size_t>(N) * d * h * w * 8 * (C / 8)), kEwThreads, 0, ST(stream),
             static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, N, d, h, w, C);
  UZ_CHECK_LAUNCH("uz_upsample3d_fwd");
  return UZ_OK;
}

extern "C" int uz_upsample3d_bwd(const void* dout, int ldd, void* dx, int ldx, int N, int d, int h, int w, int C,
                                 void* stream) {
  UZ_CHECK_ARG(dout && dx && C % 8 == 0 && ldx % 8 == 0 && ldd % 8 == 0, "uz_upsample3d_bwd: bad arguments");
  uz::launch(up3_bwd_kernel, ew_blocks(static_cast<size_t>(N) * d * h * w * (C / 8)), kEwThreads, 0, ST(stream),
             static_cast<const __nv_bfloat16*>(dout), ldd, static_cast<__nv_bfloat16*>(dx), ldx, N, d, h, w, C);
  UZ_CHECK_LAUNCH("uz_upsample3d_bwd");
  return UZ_OK;
}

extern "C" int uz_copy_channels(const void* src, int lds, void* dst, int ldd, long long npix, int C, int accumulate,
                                void* stream) {
  UZ_CHECK_ARG(src && dst && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && aligned16(src) && aligned16(dst),
               "uz_copy_channels: bad arguments");
  if (npix == 0) return UZ_OK;
  uz::launch(copy_channels_kernel, ew_blocks(static_cast<size_t>(npix) * (C / 8)), kEwThreads, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(src), lds, static_cast<__nv_bfloat16*>(dst), ldd, static_cast<size_t>(npix), C,
      accumulate, static_cast<size_t>(0));
  UZ_CHECK_LAUNCH("uz_copy_channels");
  return UZ_OK;
}

extern "C" int uz_copy_channels_bcast(const void* src, int lds, long long src_npix, void* dst, int ldd, long long npix,
                                      int C, void* stream) {
  UZ_CHECK_ARG(src && dst && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && aligned16(src) && aligned16(dst) &&
                   src_npix > 0 && npix % src_npix == 0,
               "uz_copy_channels_bcast: bad arguments");
  if (npix == 0) return UZ_OK;
  uz::launch(copy_channels_kernel, ew_blocks(static_cast<size_t>(npix) * (C / 8)), kEwThreads, 0, ST(stream),
      static_cast<const __nv_bfloat16*>(src), lds, static_cast<__nv_bfloat16*>(dst), ldd, static_cast<size_t>(npix), C, 0,
      static_cast<size_t>(src_npix));
  UZ_CHECK_LAUNCH("uz_copy_channels_bcast");
  return UZ_OK;
}

extern "C" int uz_add_channels(const void* a, int lda, const void* b, int ldb, void* out, int ldo, long long npix, int C,
                               int sign, void* stream) {
  UZ_CHECK_ARG(a && b && out && C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0 && aligned16(a) &&
                   aligned16(b) && aligned16(out),
               "uz_add_channels: bad arguments");
  if (npix == 0) return UZ_OK;
  uz::launch(add_channels_kernel, ew_blocks(static_cast<size_t>(npix) * (C / 8)), kEwThreads, 0, ST(stream),
             static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb,
             static_cast<__nv_bfloat16*>(out), ldo, static_cast<size_t>(npix), C, sign);
  UZ_CHECK_LAUNCH("uz_add_channels");
  return UZ_OK;
}

extern "C" int uz_global_mean_fwd(const void* x, int ldx, int B, int hw, int C, void* out, int ldo, void* stream) {
  UZ_CHECK_ARG(x && out && C % 2 == 0 && ldx % 2 == 0, "uz_global_mean_fwd: bad arguments");
  dim3 grid(B, (C + 63) / 64, 1);
  uz::launch(global_mean_fwd_kernel, grid, 256, 0, ST(stream), static_cast<const __nv_bfloat16*>(x), ldx, hw, C,
                                                       static_cast<__nv_bfloat16*>(out), ldo);
  UZ_CHECK_LAUNCH("uz_global_mean_fwd");
  return UZ_OK;
}

extern "C" int uz_global_mean_bwd(const void* dout, int ldd, int B, int hw, int C, void* dx, int ldx, void* stream) {
  UZ_CHECK_ARG(dout && dx && C % 8 == 0 && ldd % 8 == 0 && ldx % 8 == 0, "uz_global_mean_bwd: bad arguments");
  const size_t npix = static_cast<size_t>(B) * hw;
  uz::launch(global_mean_bwd_kernel, ew_blocks(npix * (C / 8)), kEwThreads, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(dout), ldd, hw, C, static_cast<__nv_bfloat16*>(dx), ldx, npix);
  UZ_CHECK_LAUNCH("uz_global_mean_bwd");
  return UZ_OK;
}

extern "C" int uz_input_pack(const float* patch, const float* mask, int B, int Cimg, int H, int W, int nlabels,
                             void* out, int CP, void* stream) {
  UZ_CHECK_ARG(patch && out, "uz_input_pack: null pointer");
  UZ_CHECK_ARG(CP % 8 == 0 && CP >= Cimg + (mask ? nlabels : 0), "uz_input_pack: CP=%d too small", CP);
  uz::launch(input_pack_kernel, ew_blocks(static_cast<size_t>(B) * H * W), kEwThreads, 0, ST(stream), 
      patch, mask, B, Cimg, H, W, nlabels, static_cast<__nv_bfloat16*>(out), CP);
  UZ_CHECK_LAUNCH("uz_input_pack");
  return UZ_OK;
}

extern "C" int uz_nchw_to_nhwc(const float* src, int B, int C, long long hw, void* dst, int ld, void* stream) {
  UZ_CHECK_ARG(src && dst && ld >= C, "uz_nchw_to_nhwc: bad arguments");
  uz::launch(nchw_to_nhwc_kernel, ew_blocks(static_cast<size_t>(B) * hw * ld), kEwThreads, 0, ST(stream), 
      src, B, C, static_cast<size_t>(hw), static_cast<__nv_bfloat16*>(dst), ld);
  UZ_CHECK_LAUNCH("uz_nchw_to_nhwc");
  return UZ_OK;
}

extern "C" int uz_nhwc_to_nchw(const void* src, int ld, int B, int C, long long hw, float* dst, void* stream) {
  UZ_CHECK_ARG(src && dst && ld >= C, "uz_nhwc_to_nchw: bad arguments");
  uz::launch(nhwc_to_nchw_kernel, ew_blocks(static_cast<size_t>(B) * C * hw), kEwThreads, 0, ST(stream), 
      static_cast<const __nv_bfloat16*>(src), ld, B, C, static_cast<size_t>(hw), dst);
  UZ_CHECK_LAUNCH("uz_nhwc_to_nchw");
  return UZ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Synthetic LIDC-shaped batches generated ON THE DEVICE (the reference's BatchProvider is per-image
// host code, data/batch_provider.py:43-67,131-137; this is the data plug-in's device path).  Counter-based RNG: every
// value is a hash of (seed, image, stream, index), so a batch is a pure function of its seed -- reproducible, no state.
//   patch  fp32 [B,1,S,S]   clip(N(0,1) * 0.25, -0.5, 0.5) + 0.2 * mean_m(labels)   (images are [0,1] - 0.5 in the reference)
//   labels uint8 [B,S,S,M]  M annotators: jittered filled ellipses, each empty with probability 1/4
//   mask   fp32 [B,1,S,S]   the labels of one random annotator per image (train_model.py:103-106)
namespace {
__host__ __device__ inline uint32_t uz_mix32(uint32_t h) {      // murmur3 finaliser
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__host__ __device__ inline uint32_t uz_hash4(uint32_t seed, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = uz_mix32(seed ^ 0x9E3779B9u);
  h = uz_mix32(h ^ (a * 0x85EBCA6Bu + 0x27D4EB2Fu));
  h = uz_mix32(h ^ (b * 0xC2B2AE35u + 0x165667B1u));
  h = uz_mix32(h ^ (c * 0x9E3779B1u + 0x85EBCA77u));
  return h;
}
__host__ __device__ inline float uz_u01(uint32_t h) { return (static_cast<float>(h >> 8) + 0.5f) * (1.0f / 16777216.0f); }

struct SynthImage {          // per-image ellipse parameters, derived by thread 0 of the block's image
  float cy[8], cx[8], ry[8], rx[8];
  int empty[8];
  int pick;
};

__device__ inline float uz_normal(uint32_t seed, uint32_t a, uint32_t b, uint32_t c) {
  const float u1 = uz_u01(uz_hash4(seed, a, b, c)), u2 = uz_u01(uz_hash4(seed, a, b, c + 0x40000000u));
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

__global__ void __launch_bounds__(256)
synth_lidc_kernel(uint32_t seed, int B, int S, int M, float* __restrict__ patch, uint8_t* __restrict__ labels,
                  float* __restrict__ mask) {
  uz::pdl_prologue();
  __shared__ SynthImage im;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    // explicit single-rounding operations (no FMA contraction): b200/data.py restates this arithmetic in numpy fp32
    const float fs = static_cast<float>(S);
    auto lin = [](float a, float k, float u) { return __fadd_rn(a, __fmul_rn(k, u)); };
    const float cy = __fmul_rn(lin(0.3f, 0.4f, uz_u01(uz_hash4(seed, b, 1, 0))), fs);
    const float cx = __fmul_rn(lin(0.3f, 0.4f, uz_u01(uz_hash4(seed, b, 1, 1))), fs);
    const float ry = __fmul_rn(lin(0.06f, 0.12f, uz_u01(uz_hash4(seed, b, 1, 2))), fs);
    const float rx = __fmul_rn(lin(0.06f, 0.12f, uz_u01(uz_hash4(seed, b, 1, 3))), fs);
    for (int m = 0; m < M; ++m) {
      im.empty[m] = uz_u01(uz_hash4(seed, b, 2, m)) < 0.25f;
      im.cy[m] = __fadd_rn(cy, __fmul_rn(__fmul_rn(0.02f, fs), uz_normal(seed, b, 3, m)));
      im.cx[m] = __fadd_rn(cx, __fmul_rn(__fmul_rn(0.02f, fs), uz_normal(seed, b, 4, m)));
      im.ry[m] = __fmul_rn(ry, lin(0.8f, 0.45f, uz_u01(uz_hash4(seed, b, 5, m))));
      im.rx[m] = __fmul_rn(rx, lin(0.8f, 0.45f, uz_u01(uz_hash4(seed, b, 6, m))));
    }
    im.pick = static_cast<int>(uz_hash4(seed, b, 7, 0) % static_cast<uint32_t>(M));
  }
  __syncthreads();
  const int hw = S * S;
  for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < hw; px += gridDim.x * blockDim.x) {
    const int y = px / S, x = px - y * S;
    float v = 0.25f * uz_normal(seed, b, 8, static_cast<uint32_t>(px));
    v = fminf(fmaxf(v, -0.5f), 0.5f);
    int count = 0, picked = 0;
    for (int m = 0; m < M; ++m) {
      int inside = 0;
      if (!im.empty[m]) {
        const float dy = __fdiv_rn(__fsub_rn(static_cast<float>(y), im.cy[m]), im.ry[m]);
        const float dx = __fdiv_rn(__fsub_rn(static_cast<float>(x), im.cx[m]), im.rx[m]);
        inside = __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)) <= 1.0f;
      }
      labels[(static_cast<size_t>(b) * hw + px) * M + m] = static_cast<uint8_t>(inside);
      count += inside;
      if (m == im.pick) picked = inside;
    }
    patch[static_cast<size_t>(b) * hw + px] = v + 0.2f * (static_cast<float>(count) / static_cast<float>(M));
    mask[static_cast<size_t>(b) * hw + px] = static_cast<float>(picked);
  }
}
}  // namespace

extern "C" int uz_synth_lidc_batch(unsigned int seed, int B, int size, int annotators, float* patch,
                                   unsigned char* labels, float* mask, void* stream) {
  UZ_CHECK_ARG(patch && labels && mask && B > 0 && size > 0 && annotators >= 1 && annotators <= 8,
               "uz_synth_lidc_batch: bad arguments");
  int bx = (size * size + 255) / 256;
  if (bx > 64) bx = 64;
  uz::launch(synth_lidc_kernel, dim3(bx, B, 1), 256, 0, ST(stream), seed, B, size, annotators, patch, labels, mask);
  UZ_CHECK_LAUNCH("uz_synth_lidc_batch");
  return UZ_OK;
}
