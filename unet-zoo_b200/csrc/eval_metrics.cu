// Sample-evaluation metrics of the reference on the device:
//   generalised_energy_distance   reference utils.py:148-200 (pairwise 1 - mean-label IoU; medpy jc at :166)
//   variance_ncc_dist / ncc       reference utils.py:202-247 / :130-145
// GED: label masks are bit-packed once (ballot), every pair distance is popcount work on 32-bit words, and the three
// sums are formed by ONE thread in the reference's pair order in fp64, so the result is bit-identical to the Python
// loops.  NCC: one thread per pixel walks the N samples (fp32 logs, fp64 means like numpy), then one block per
// annotator forms the normalised cross-correlation in fp64.
#include "common.cuh"
#include "unetzoo_b200.h"

namespace {

constexpr int kMaxLabels = 8;
struct LabelSet {
  int n;
  int v[kMaxLabels];
};

template <typename T>
__global__ void pack_masks_kernel(const T* __restrict__ labels, int count, int hw, int words, LabelSet ls,
                                  uint32_t* bits, int* counts) {
  uz::pdl_prologue();
  // one warp per (mask, word)
  const size_t gw = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= static_cast<size_t>(count) * words) return;
  const int m = gw / words;
  const int w = gw - static_cast<size_t>(m) * words;
  const int px = w * 32 + lane;
  float v = -1.f;
  if (px < hw) v = static_cast<float>(labels[static_cast<size_t>(m) * hw + px]);
  for (int l = 0; l < ls.n; ++l) {
    const uint32_t b = __ballot_sync(0xffffffffu, px < hw && v == static_cast<float>(ls.v[l]));
    if (lane == 0) {
      bits[(static_cast<size_t>(m) * ls.n + l) * words + w] = b;
      if (b) atomicAdd(&counts[m * ls.n + l], __popc(b));
    }
  }
}

// pair p: [0, N*M) = (sample i, gt j); then N*N sample pairs; then M*M gt pairs -- the reference's loop order.
__global__ void pair_distance_kernel(const uint32_t* __restrict__ bits_s, const int* __restrict__ cnt_s, int N,
                                     const uint32_t* __restrict__ bits_y, const int* __restrict__ cnt_y, int M, int nl,
                                     int words, double* pair_d) {
  uz::pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = N * M + N * N + M * M;
  if (warp >= total) return;
  const uint32_t *ba, *bb;
  const int *ca, *cb;
  int i, j;
  if (warp < N * M) {
    i = warp / M; j = warp % M;
    ba = bits_s; ca = cnt_s; bb = bits_y; cb = cnt_y;
  } else if (warp < N * M + N * N) {
    const int q = warp - N * M;
    i = q / N; j = q % N;
    ba = bits_s; ca = cnt_s; bb = bits_s; cb = cnt_s;
  } else {
    const int q = warp - N * M - N * N;
    i = q / M; j = q % M;
    ba = bits_y; ca = cnt_y; bb = bits_y; cb = cnt_y;
  }
  double iou_sum = 0.0;
  for (int l = 0; l < nl; ++l) {
    const int na = ca[i * nl + l], nb = cb[j * nl + l];
    if (na == 0 && nb == 0) {
      iou_sum += 1.0;                       // both empty -> IoU 1 (utils.py:161-162)
    } else if (na == 0 || nb == 0) {
      iou_sum += 0.0;                       // exactly one empty -> IoU 0 (utils.py:163-164)
    } else {
      const uint32_t* pa = ba + (static_cast<size_t>(i) * nl + l) * words;
      const uint32_t* pb = bb + (static_cast<size_t>(j) * nl + l) * words;
      int inter = 0;
      for (int w = lane; w < words; w += 32) inter += __popc(pa[w] & pb[w]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) inter += __shfl_xor_sync(0xffffffffu, inter, o);
      iou_sum += static_cast<double>(inter) / static_cast<double>(na + nb - inter);
    }
  }
  if (lane == 0) pair_d[warp] = __dadd_rn(1.0, -(iou_sum / static_cast<double>(nl)));
}

// Python's builtin sum() over floats: CPython >= 3.12 uses Neumaier compensated summation (Objects/bltinmodule.c),
// which is what the reference's `sum(d_sy)` executes under this image's Python 3.12.  Reproduced step for step so the
// GED is bit-identical; __dadd_rn/__dmul_rn keep the compiler from contracting into FMAs.
__device__ __forceinline__ void py312_step(double x, double& f, double& c) {
  const double t = __dadd_rn(f, x);
  if (fabs(f) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dadd_rn(f, -t), x));
  else c = __dadd_rn(c, __dadd_rn(__dadd_rn(x, -t), f));
  f = t;
}
// sequential by definition (every step depends on the previous running sum); the block stages the terms in shared memory
// in chunks so the one summing thread never waits for global memory
constexpr int kGedChunk = 512;
__device__ double py312_sum_staged(const double* __restrict__ p, int n, double* stage /*[2][kGedChunk]*/, int lane) {
  double f = 0.0, c = 0.0;
  for (int i = lane; i < min(n, kGedChunk); i += 32) stage[i] = p[i];
  __syncwarp();
  for (int base = 0, buf = 0; base < n; base += kGedChunk, buf ^= 1) {
    const int cnt = min(kGedChunk, n - base);
    const int nxt = base + kGedChunk;
    // lanes 1..31 fetch the next chunk while lane 0 walks this one
    if (lane != 0) {
      for (int i = lane - 1; i < min(kGedChunk, n - nxt); i += 31) stage[(buf ^ 1) * kGedChunk + i] = p[nxt + i];
    } else {
      const double* x = stage + buf * kGedChunk;
#pragma unroll 4
      for (int k = 0; k < cnt; ++k) py312_step(x[k], f, c);
    }
    __syncwarp();
  }
  if (c != 0.0 && isfinite(c)) f = __dadd_rn(f, c);
  return f;
}

// out[0] = GED, out[1..3] = sum d_sy, sum d_ss, sum d_yy, in the reference's pair order (utils.py:185-200).
// Three warps: the three Python sums are independent, each is sequential by definition.
__global__ void __launch_bounds__(96) ged_finish_kernel(const double* __restrict__ pair_d, int N, int M, double* out) {
  uz::pdl_prologue();
  __shared__ double sums[3];
  __shared__ double stage[3][2 * kGedChunk];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < 3) {
    const double* base = w == 0 ? pair_d : (w == 1 ? pair_d + N * M : pair_d + N * M + N * N);
    const int n = w == 0 ? N * M : (w == 1 ? N * N : M * M);
    const double r = py312_sum_staged(base, n, stage[w], lane);
    if (lane == 0) sums[w] = r;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double sy = sums[0], ss = sums[1], yy = sums[2];
  out[1] = sy; out[2] = ss; out[3] = yy;
  const double a = __dmul_rn(2.0 / static_cast<double>(N * M), sy);
  const double b = __dmul_rn(1.0 / static_cast<double>(N * N), ss);
  const double c = __dmul_rn(1.0 / static_cast<double>(M * M), yy);
  out[0] = __dadd_rn(__dadd_rn(a, -b), -c);
}

// argmax over classes of fp32 NCHW [N,C,HW] -> uint8 [N,HW] (first maximal index, like torch.argmax)
__global__ void argmax_kernel(const float* __restrict__ x, int N, int C, int hw, uint8_t* out) {
  uz::pdl_prologue();
  const size_t total = static_cast<size_t>(N) * hw;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = idx / hw, r = idx - n * hw;
    float best = x[(n * C) * hw + r];
    int bi = 0;
    for (int c = 1; c < C; ++c) {
      const float v = x[(n * C + c) * hw + r];
      if (v > best) { best = v; bi = c; }
    }
    out[idx] = static_cast<uint8_t>(bi);
  }
}

// ---------------------------------------------------------------- NCC
template <typename GT>
__global__ void ncc_pixel_kernel(const float* __restrict__ probs, const GT* __restrict__ gt, int N, int C, int hw,
                                 int M, double* e_ss, double* e_sy /*[M][hw]*/) {
  uz::pdl_prologue();
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= hw) return;
  constexpr int kC = 8;
  float mean_seg[kC];
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (int i = 0; i < N; ++i) s += probs[(static_cast<size_t>(i) * C + c) * hw + px];
    mean_seg[c] = s / static_cast<float>(N);
  }
  double ss = 0.0;
  double sy[16];
  for (int j = 0; j < M; ++j) sy[j] = 0.0;
  for (int i = 0; i < N; ++i) {
    float lg[kC];
    float inner = 0.f;
    for (int c = 0; c < C; ++c) {
      lg[c] = logf(probs[(static_cast<size_t>(i) * C + c) * hw + px] + 1e-8f);
      inner += mean_seg[c] * lg[c];
    }
    ss += static_cast<double>(-1.0f * inner);
    for (int j = 0; j < M; ++j) {
      double t = 0.0;
      for (int c = 0; c < C; ++c)
        t += static_cast<double>(gt[(static_cast<size_t>(j) * C + c) * hw + px]) * static_cast<double>(lg[c]);
      sy[j] += -1.0 * t;
    }
  }
  e_ss[px] = ss / static_cast<double>(N);
  for (int j = 0; j < M; ++j) e_sy[static_cast<size_t>(j) * hw + px] = sy[j] / static_cast<double>(N);
}

__device__ double block_sum_d(double v, double* red) {
  v = uz::warp_sum_d(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// one block per annotator j: ncc_j = sum((a-abar)/(std_a*len) * (v-vbar)/std_v)
__global__ void ncc_corr_kernel(const double* __restrict__ e_ss, const double* __restrict__ e_sy, int hw, double* ncc_j) {
  uz::pdl_prologue();
  __shared__ double red[32];
  const int j = blockIdx.x;
  const double* v = e_sy + static_cast<size_t>(j) * hw;
  double sa = 0.0, sv = 0.0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) { sa += e_ss[i]; sv += v[i]; }
  const double ma = block_sum_d(sa, red) / hw;
  const double mv = block_sum_d(sv, red) / hw;
  double qa = 0.0, qv = 0.0, cv = 0.0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const double da = e_ss[i] - ma, dv = v[i] - mv;
    qa += da * da; qv += dv * dv; cv += da * dv;
  }
  qa = block_sum_d(qa, red);
  qv = block_sum_d(qv, red);
  cv = block_sum_d(cv, red);
  if (threadIdx.x == 0) {
    const double std_a = sqrt(qa / hw), std_v = sqrt(qv / hw);
    ncc_j[j] = cv / (std_a * hw * std_v);
  }
}

__global__ void ncc_finish_kernel(const double* __restrict__ ncc_j, int M, double* out) {
  uz::pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0;
  for (int j = 0; j < M; ++j) s += ncc_j[j];
  out[0] = (1.0 / M) * s;
}


// ---------------------------------------------------------------------------------------------------------------------
// Fused N-sample evaluation (train_model.py:185-222 after the network): ONE pass over the per-level class logits does
// accumulate_output (phiseg.py:428-434, same fp32 summation order), softmax, argmax -> bit-packed label masks, and the
// two per-pixel sums sum_i p_ic and sum_i log(p_ic + 1e-8) that variance_ncc_dist needs (the
// single-pass form of utils.py:202-247).  The levels may be low-resolution logits (nearest upsampling by factor f,
// phiseg.py:321, is index arithmetic here) so the five full-resolution fp32 tensors are never written.
// Both sums are additive over samples -> over GPUs: one all-reduce of [2][C][HW]; the masks are 2 KB per sample.
constexpr int kEvalMaxLevels = 8;
constexpr int kEvalMaxClasses = 4;
struct EvalLevels {
  const float* s[kEvalMaxLevels];   // fp32 [B][C][(H/f) * (W/f)], B = n * I with batch index b = sample * I + image
  int f[kEvalMaxLevels];
  int L;
};

// grid (ceil(HW/256), G sample groups, I images); thread = pixel, warp = one 32-pixel mask word.
// part: fp32 [I][G][2][C][HW] (sum p, sum log(p+1e-8) over the group's samples, fixed order)
__global__ void __launch_bounds__(256)
eval_sample_stats_kernel(EvalLevels lv, int n, int I, int C, int H, int W, LabelSet ls, uint32_t* bits /*[I][n][nl][words]*/,
                         int* counts /*[I][n][nl], zero on entry*/, float* part) {
  uz::pdl_prologue();
  const int HW = H * W, words = (HW + 31) / 32;
  const int px = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.y, G = gridDim.y, img = blockIdx.z;
  const int per = (n + G - 1) / G;
  const int s0 = g * per, s1 = min(n, s0 + per);
  const bool inside = px < HW;
  const int y = inside ? px / W : 0, x = inside ? px - y * W : 0;
  int off[kEvalMaxLevels], plane[kEvalMaxLevels];
  for (int l = 0; l < lv.L; ++l) {
    const int f = lv.f[l], wl = W / f, hl = H / f;
    off[l] = (y / f) * wl + x / f;
    plane[l] = hl * wl;
  }
  float ps[kEvalMaxClasses], lsum[kEvalMaxClasses];
#pragma unroll
  for (int c = 0; c < kEvalMaxClasses; ++c) { ps[c] = 0.f; lsum[c] = 0.f; }
  for (int smp = s0; smp < s1; ++smp) {
    const size_t b = static_cast<size_t>(smp) * I + img;
    float acc[kEvalMaxClasses];
#pragma unroll
    for (int c = 0; c < kEvalMaxClasses; ++c) {
      acc[c] = 0.f;
      if (c < C && inside) {
        // s_accum = list[-1]; s_accum += list[0], list[1], ... (phiseg.py:429-431): the reference's fp32 order
        float a = lv.s[lv.L - 1][(b * C + c) * plane[lv.L - 1] + off[lv.L - 1]];
        for (int l = 0; l < lv.L - 1; ++l) a += lv.s[l][(b * C + c) * plane[l] + off[l]];
        acc[c] = a;
      }
    }
    float m = acc[0];
#pragma unroll
    for (int c = 1; c < kEvalMaxClasses; ++c) if (c < C) m = fmaxf(m, acc[c]);
    float e[kEvalMaxClasses], den = 0.f;
#pragma unroll
    for (int c = 0; c < kEvalMaxClasses; ++c) { e[c] = c < C ? expf(acc[c] - m) : 0.f; den += e[c]; }
    int arg = 0;
    float best = -1.f;
#pragma unroll
    for (int c = 0; c < kEvalMaxClasses; ++c) {
      if (c < C) {
        const float p = e[c] / den;
        if (p > best) { best = p; arg = c; }              // first maximal index, like torch.argmax
        ps[c] += p;
        lsum[c] += logf(p + 1e-8f);
      }
    }
    for (int l = 0; l < ls.n; ++l) {
      const uint32_t word = __ballot_sync(0xffffffffu, inside && arg == ls.v[l]);
      if (lane == 0 && (px >> 5) < words) {
        const size_t row = (static_cast<size_t>(img) * n + smp) * ls.n + l;
        bits[row * words + (px >> 5)] = word;
        if (word) atomicAdd(&counts[row], __popc(word));    // integer: order independent
      }
    }
  }
  if (inside) {
    float* dst = part + ((static_cast<size_t>(img) * G + g) * 2) * C * HW + px;
    for (int c = 0; c < C; ++c) {
      dst[static_cast<size_t>(c) * HW] = ps[c];
      dst[static_cast<size_t>(C + c) * HW] = lsum[c];
    }
  }
}

// sums [I][2][C][HW] = sum over the G groups in fixed order
__global__ void eval_stats_reduce_kernel(const float* __restrict__ part, int I, int G, size_t per_image, float* sums) {
  uz::pdl_prologue();
  const size_t total = static_cast<size_t>(I) * per_image;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t img = idx / per_image, r = idx - img * per_image;
    float t = 0.f;
    for (int g = 0; g < G; ++g) t += part[(img * G + g) * per_image + r];
    sums[idx] = t;
  }
}

// per pixel from the (all-reduced) sums of ONE image: E_ss = -sum_c (P_c/N)(L_c/N), E_sy[j] = -sum_c [gt_j == c] L_c/N
// (utils.py:219-238), and the Dice counts of argmax_c(mean probs) against one annotator (train_model.py:207-222).
template <typename GT>
__global__ void ncc_from_sums_kernel(const float* __restrict__ sums /*[2][C][HW]*/, const GT* __restrict__ gt /*[M][HW]*/,
                                     int N, int C, int hw, int M, int dice_annotator, double* e_ss, double* e_sy,
                                     int* dice_counts /*[C][3]: |pred|, |gt|, |both|; zero on entry*/) {
  uz::pdl_prologue();
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= hw) return;
  const float inv = 1.f / static_cast<float>(N);
  float lbar[kEvalMaxClasses];
  double ss = 0.0;
  int arg = 0;
  float best = -1.f;
  for (int c = 0; c < C; ++c) {
    const float psum = sums[static_cast<size_t>(c) * hw + px];
    lbar[c] = sums[static_cast<size_t>(C + c) * hw + px] * inv;
    ss -= static_cast<double>(psum * inv) * static_cast<double>(lbar[c]);
    if (psum > best) { best = psum; arg = c; }
  }
  e_ss[px] = ss;
  for (int j = 0; j < M; ++j) {
    const int lbl = static_cast<int>(gt[static_cast<size_t>(j) * hw + px]);
    e_sy[static_cast<size_t>(j) * hw + px] = (lbl >= 0 && lbl < C) ? -static_cast<double>(lbar[lbl]) : 0.0;
  }
  if (dice_counts && dice_annotator >= 0) {
    const int lbl = static_cast<int>(gt[static_cast<size_t>(dice_annotator) * hw + px]);
    atomicAdd(&dice_counts[arg * 3 + 0], 1);
    if (lbl >= 0 && lbl < C) {
      atomicAdd(&dice_counts[lbl * 3 + 1], 1);
      if (lbl == arg) atomicAdd(&dice_counts[arg * 3 + 2], 1);
    }
  }
}

// per label: both empty -> 1, exactly one empty -> 0, else medpy dc = 2|A&B| / (|A| + |B|)  (train_model.py:211-222)
__global__ void dice_finish_kernel(const int* __restrict__ counts, int C, double* dice) {
  uz::pdl_prologue();
  const int c = threadIdx.x;
  if (c >= C) return;
  const int np = counts[c * 3], ng = counts[c * 3 + 1], nb = counts[c * 3 + 2];
  dice[c] = (np == 0 && ng == 0) ? 1.0 : ((np == 0 || ng == 0) ? 0.0 : 2.0 * nb / static_cast<double>(np + ng));
}

}  // namespace

#define ST(s) static_cast<cudaStream_t>(s)

// labels: [count][hw] of dtype (0 = int64, 1 = float32, 2 = uint8).  bits: uint32 [count][nlabels][words],
// words = ceil(hw/32); counts: int32 [count][nlabels] (zeroed here).
extern "C" int uz_ged_pack_masks(const void* labels, int dtype, int count, int hw, const int* label_values,
                                 int nlabels, unsigned int* bits, int* counts, void* stream) {
  UZ_CHECK_ARG(labels && label_values && bits && counts, "uz_ged_pack_masks: null pointer");
  UZ_CHECK_ARG(nlabels >= 1 && nlabels <= kMaxLabels, "uz_ged_pack_masks: nlabels %d unsupported", nlabels);
  LabelSet ls{};
  ls.n = nlabels;
  for (int l = 0; l < nlabels; ++l) ls.v[l] = label_values[l];
  const int words = (hw + 31) / 32;
  cudaMemsetAsync(counts, 0, sizeof(int) * count * nlabels, ST(stream));
  const size_t threads_total = static_cast<size_t>(count) * words * 32;
  const int blocks = static_cast<int>((threads_total + 255) / 256);
  if (dtype == 0)
    uz::launch(pack_masks_kernel<long long>, blocks, 256, 0, ST(stream), static_cast<const long long*>(labels), count, hw, words,
                                                                 ls, bits, counts);
  else if (dtype == 1)
    uz::launch(pack_masks_kernel<float>, blocks, 256, 0, ST(stream), static_cast<const float*>(labels), count, hw, words, ls, bits,
                                                             counts);
  else if (dtype == 2)
    uz::launch(pack_masks_kernel<uint8_t>, blocks, 256, 0, ST(stream), static_cast<const uint8_t*>(labels), count, hw, words, ls,
                                                               bits, counts);
  else
    UZ_CHECK_ARG(false, "uz_ged_pack_masks: dtype %d unsupported", dtype);
  UZ_CHECK_LAUNCH("uz_ged_pack_masks");
  return UZ_OK;
}

// pair_d: double [N*M + N*N + M*M]; out: double [4] = {GED, sum d_sy, sum d_ss, sum d_yy}
extern "C" int uz_ged_pairwise(const unsigned int* bits_s, const int* cnt_s, int N, const unsigned int* bits_y,
                               const int* cnt_y, int M, int nlabels, int hw, double* pair_d, double* out,
                               void* stream) {
  UZ_CHECK_ARG(bits_s && cnt_s && bits_y && cnt_y && pair_d && out && N > 0 && M > 0, "uz_ged_pairwise: bad arguments");
  const int words = (hw + 31) / 32;
  const int total = N * M + N * N + M * M;
  uz::launch(pair_distance_kernel, (total * 32 + 255) / 256, 256, 0, ST(stream), bits_s, cnt_s, N, bits_y, cnt_y, M, nlabels,
                                                                        words, pair_d);
  UZ_CHECK_LAUNCH("uz_ged_pairwise");
  uz::launch(ged_finish_kernel, 1, 96, 0, ST(stream), pair_d, N, M, out);
  UZ_CHECK_LAUNCH("uz_ged_pairwise(finish)");
  return UZ_OK;
}

extern "C" int uz_argmax_classes(const float* x, int N, int C, int hw, unsigned char* out, void* stream) {
  UZ_CHECK_ARG(x && out && C >= 1 && C <= 255, "uz_argmax_classes: bad arguments");
  size_t total = static_cast<size_t>(N) * hw;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > uz::num_sms() * 16) blocks = uz::num_sms() * 16;
  uz::launch(argmax_kernel, blocks, 256, 0, ST(stream), x, N, C, hw, out);
  UZ_CHECK_LAUNCH("uz_argmax_classes");
  return UZ_OK;
}

// probs fp32 [N,C,hw]; gt one-hot [M,C,hw] of gt_dtype (0 = int64, 1 = float32, 2 = uint8);
// work: double [(1 + M) * hw + M]; out: double [1].
extern "C" int uz_variance_ncc(const float* probs, const void* gt, int gt_dtype, int N, int C, int hw, int M,
                               double* work, double* out, void* stream) {
  UZ_CHECK_ARG(probs && gt && work && out, "uz_variance_ncc: null pointer");
  UZ_CHECK_ARG(C >= 1 && C <= 8 && M >= 1 && M <= 16 && N >= 1, "uz_variance_ncc: C=%d M=%d unsupported", C, M);
  double* e_ss = work;
  double* e_sy = work + hw;
  double* ncc_j = work + static_cast<size_t>(1 + M) * hw;
  const int blocks = (hw + 127) / 128;
  if (gt_dtype == 0)
    uz::launch(ncc_pixel_kernel<long long>, blocks, 128, 0, ST(stream), probs, static_cast<const long long*>(gt), N, C, hw, M,
                                                                e_ss, e_sy);
  else if (gt_dtype == 1)
    uz::launch(ncc_pixel_kernel<float>, blocks, 128, 0, ST(stream), probs, static_cast<const float*>(gt), N, C, hw, M, e_ss, e_sy);
  else if (gt_dtype == 2)
    uz::launch(ncc_pixel_kernel<uint8_t>, blocks, 128, 0, ST(stream), probs, static_cast<const uint8_t*>(gt), N, C, hw, M, e_ss,
                                                              e_sy);
  else
    UZ_CHECK_ARG(false, "uz_variance_ncc: gt dtype %d unsupported", gt_dtype);
  UZ_CHECK_LAUNCH("uz_variance_ncc(pixel)");
  uz::launch(ncc_corr_kernel, M, 1024, 0, ST(stream), e_ss, e_sy, hw, ncc_j);
  UZ_CHECK_LAUNCH("uz_variance_ncc(corr)");
  uz::launch(ncc_finish_kernel, 1, 32, 0, ST(stream), ncc_j, M, out);
  UZ_CHECK_LAUNCH("uz_variance_ncc(finish)");
  return UZ_OK;
}


// ---- fused N-sample evaluation -------------------------------------------------------------------------------------
extern "C" int uz_eval_sample_groups(int n, int I, int hw) {
  // ~8 blocks of 256 threads per SM: a thread's samples are a serial chain of dependent-latency loads, the loads in flight
  // come from the number of resident warps (first version: 2 blocks per SM, 59 us = 0.05 of the HBM roofline at N = 100);
  // at least 2 samples per group so the partial sums stay small
  if (n <= 0 || I <= 0 || hw <= 0) return -1;
  const int per_g = (hw + 255) / 256 * I;
  int G = (8 * uz::num_sms() + per_g - 1) / per_g;
  if (G > (n + 1) / 2) G = (n + 1) / 2;
  if (G < 1) G = 1;
  return G;
}

extern "C" int uz_eval_sample_stats(const float* const* levels, const int* factors, int L, int n, int I, int C, int H,
                                    int W, const int* label_values, int nlabels, unsigned int* bits, int* counts,
                                    float* part, float* sums, void* stream) {
  UZ_CHECK_ARG(levels && factors && label_values && bits && counts && part && sums, "uz_eval_sample_stats: null pointer");
  UZ_CHECK_ARG(L >= 1 && L <= kEvalMaxLevels && C >= 1 && C <= kEvalMaxClasses && nlabels >= 1 && nlabels <= kMaxLabels,
               "uz_eval_sample_stats: L=%d C=%d nlabels=%d unsupported", L, C, nlabels);
  EvalLevels lv{};
  lv.L = L;
  for (int l = 0; l < L; ++l) {
    UZ_CHECK_ARG(levels[l] && factors[l] >= 1 && H % factors[l] == 0 && W % factors[l] == 0,
                 "uz_eval_sample_stats: level %d: factor must divide the image size", l);
    lv.s[l] = levels[l];
    lv.f[l] = factors[l];
  }
  LabelSet ls{};
  ls.n = nlabels;
  for (int l = 0; l < nlabels; ++l) ls.v[l] = label_values[l];
  const int hw = H * W;
  const int G = uz_eval_sample_groups(n, I, hw);
  cudaMemsetAsync(counts, 0, sizeof(int) * static_cast<size_t>(I) * n * nlabels, ST(stream));
  uz::launch(eval_sample_stats_kernel, dim3((hw + 255) / 256, G, I), 256, 0, ST(stream), lv, n, I, C, H, W, ls, bits, counts,
             part);
  UZ_CHECK_LAUNCH("uz_eval_sample_stats");
  const size_t per_image = static_cast<size_t>(2) * C * hw;
  int blocks = static_cast<int>((per_image * I + 255) / 256);
  if (blocks > uz::num_sms() * 8) blocks = uz::num_sms() * 8;
  uz::launch(eval_stats_reduce_kernel, blocks, 256, 0, ST(stream), part, I, G, per_image, sums);
  UZ_CHECK_LAUNCH("uz_eval_sample_stats(reduce)");
  return UZ_OK;
}

extern "C" int uz_ncc_dice_from_sums(const float* sums, const void* gt, int gt_dtype, int N, int C, int hw, int M,
                                     int dice_annotator, double* work, int* dice_counts, double* out, void* stream) {
  UZ_CHECK_ARG(sums && gt && work && out, "uz_ncc_dice_from_sums: null pointer");
  UZ_CHECK_ARG(C >= 1 && C <= kEvalMaxClasses && M >= 1 && M <= 16 && N >= 1 && dice_annotator < M,
               "uz_ncc_dice_from_sums: C=%d M=%d unsupported", C, M);
  double* e_ss = work;
  double* e_sy = work + hw;
  double* ncc_j = work + static_cast<size_t>(1 + M) * hw;
  if (dice_counts) cudaMemsetAsync(dice_counts, 0, sizeof(int) * 3 * C, ST(stream));
  const int blocks = (hw + 127) / 128;
  if (gt_dtype == 1)
    uz::launch(ncc_from_sums_kernel<float>, blocks, 128, 0, ST(stream), sums, static_cast<const float*>(gt), N, C, hw, M,
               dice_annotator, e_ss, e_sy, dice_counts);
  else if (gt_dtype == 2)
    uz::launch(ncc_from_sums_kernel<uint8_t>, blocks, 128, 0, ST(stream), sums, static_cast<const uint8_t*>(gt), N, C, hw,
               M, dice_annotator, e_ss, e_sy, dice_counts);
  else if (gt_dtype == 0)
    uz::launch(ncc_from_sums_kernel<long long>, blocks, 128, 0, ST(stream), sums, static_cast<const long long*>(gt), N, C,
               hw, M, dice_annotator, e_ss, e_sy, dice_counts);
  else
    UZ_CHECK_ARG(false, "uz_ncc_dice_from_sums: gt dtype %d unsupported", gt_dtype);
  UZ_CHECK_LAUNCH("uz_ncc_dice_from_sums(pixel)");
  uz::launch(ncc_corr_kernel, M, 1024, 0, ST(stream), e_ss, e_sy, hw, ncc_j);
  UZ_CHECK_LAUNCH("uz_ncc_dice_from_sums(corr)");
  uz::launch(ncc_finish_kernel, 1, 32, 0, ST(stream), ncc_j, M, out);
  if (dice_counts) uz::launch(dice_finish_kernel, 1, 32, 0, ST(stream), dice_counts, C, out + 1);
  UZ_CHECK_LAUNCH("uz_ncc_dice_from_sums(finish)");
  return UZ_OK;
}
