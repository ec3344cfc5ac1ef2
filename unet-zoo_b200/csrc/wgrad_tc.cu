// Weight gradient of the 3x3 / 1x1 convolution on tcgen05 tensor cores.
//
//   dW[tap][co][ci] = sum over pixels p of dy[p][co] * x[p + offset(tap)][ci]
//
// GEMM view per tap: D[M = co][N = ci] += A[M = co][K = pixel] * B[N = ci][K = pixel].  Both operands are NHWC, i.e. the
// M/N index (channel) is contiguous and K (pixel) is strided: the "MN-major" operand mode of tcgen05.mma, fed directly
// by the same TMA boxes the forward kernel uses ([pixels][64 channels], 128-byte swizzle).  The tap offset is applied
// to the x box coordinates; out-of-bounds pixels (zero padding) are zero-filled by TMA.
//
// Work split: grid.x = pixel splits (split-K), grid.y = (tap group) x (128-wide block of co).  A CTA keeps one fp32
// accumulator [128 x Cin] per tap of its group in TMEM (taps_per_cta * Cin <= 512 columns), streams its share of the
// pixel tiles through a TMA/mbarrier pipeline and finally writes a partial [tap][co][ci] slab; uz_wgrad_reduce sums the
// slabs in fixed order (deterministic) into the PyTorch OIHW fp32 gradient.
#include <cstdlib>
#include "common.cuh"
#include "unetzoo_b200.h"

namespace uz {
extern int g_conv_debug_flags;
}

namespace {

constexpr int kThreads = 192;
constexpr int kMaxStages = 6;

struct WgradParams {
  int N, H, W, Cin, Cout;   // Cin / Cout as stored (multiples of 16)
  int taps;                 // 9 or 1
  int PIX;                  // pixels per K tile (64 or 128)
  int TW, TH, TN;
  int tilesW, tilesH, num_tiles;
  int taps_per_cta, tap_groups;
  int a_boxes, b_boxes;     // 64-channel boxes per stage for dy (<=2) and x (of this CTA's input-channel chunk)
  int ci_w, ci_chunks;      // input channels per CTA (Cin, or 64 for tiny layers) and number of such chunks
  int co_blocks;
  int stages;
  uint32_t tmem_cols;
  float* partial;           // [splits][taps][Cout][Cin]
  int accumulate;           // 1: every split ADDS into slab 0 (zero on entry) with bulk reduce-add stores
  unsigned long long* trace;   // profiling build: phase timestamps of CTA 0
};


// Epilogue of the weight-gradient kernels.  The accumulator block [128 rows (TMEM lanes) x ncols] fp32 used to leave one
// row per thread: every 16-byte store instruction of a warp touched 32 different 32-byte sectors, the LSU -- not the
// tensor core, not HBM -- bounded the small launches (phase trace, profiles/r02_phase_trace.md: 11 of 14 us of a
// 192 -> 192 launch were these stores).  Now the block is staged in the (by then idle) pipeline shared memory, rows
// padded by 16 bytes so the row-per-thread 16-byte writes are bank-conflict free, and written out with consecutive
// threads on consecutive 16 bytes: 512 contiguous bytes per warp store.
// accumulate: the rows are ADDED to global memory (cp.reduce.async.bulk ... .add.f32: the reduction happens in L2) -- all
// split-K CTAs of a layer then share ONE [tap][Cout][Cin] slab instead of writing `splits` of them for a later pass
// to read back (0.8 GB per PHiSeg step, a 0.22 ms reduction alone at the end of backward); a CTA without work skips.
__device__ __forceinline__ void store_acc_block(uint32_t taddr, int ncols, bool have_acc, float* stage, float* dst,
                                                size_t ld, int valid_rows, int row, int et, bool accumulate,
                                                unsigned long long* trace = nullptr) {
  if (accumulate && !have_acc) return;
  // fixed pitch (128 columns + 16 bytes: bank-conflict-free 16-byte row writes): a thread's staging row is the same
  // private region for every block, so its own wait_group is all the synchronisation the reuse needs
  constexpr int pitch = 128 + 4;
  float* srow = stage + static_cast<size_t>(row) * pitch;
  int c = 0;
  for (; c + 32 <= ncols; c += 32) {            // two TMEM loads in flight per wait
    uint32_t r[16], r2[16];
    if (have_acc) {
      uz::tmem_ld16(taddr + c, r);
      uz::tmem_ld16(taddr + c + 16, r2);
      uz::tmem_ld_wait();
      if (c == 0) UZ_TRACE(trace, et == 0 ? 9 : 15);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) { r[j] = 0u; r2[j] = 0u; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(srow + c + 4 * j) =
          make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                      __uint_as_float(r[4 * j + 3]));
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(srow + c + 16 + 4 * j) =
          make_float4(__uint_as_float(r2[4 * j]), __uint_as_float(r2[4 * j + 1]), __uint_as_float(r2[4 * j + 2]),
                      __uint_as_float(r2[4 * j + 3]));
  }
  for (; c < ncols; c += 16) {
    uint32_t r[16];
    if (have_acc) {
      uz::tmem_ld16(taddr + c, r);
      uz::tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(srow + c + 4 * j) =
          make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                      __uint_as_float(r[4 * j + 3]));
  }
  UZ_TRACE(trace, et == 0 ? 10 : 15);
  // copy-out: every thread hands the row it just staged to the TMA engine as ONE bulk copy (cp.async.bulk shared ->
  // global, ncols * 4 contiguous bytes on both sides).  No block barrier is needed -- a thread only ever touches its own
  // staging row -- and the stores no longer go through the four epilogue warps' LSU path (row-per-thread st.global ran at
  // ~24 GB/s per SM, a coalesced eight-rows-in-flight loop at ~35 GB/s: profiles/r02_phase_trace.md).
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  UZ_TRACE(trace, et == 0 ? 11 : 15);
  if (row < valid_rows) {
    if (accumulate)
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                       dst + static_cast<size_t>(row) * ld),
                   "r"(uz::smem_u32(srow)), "r"(static_cast<uint32_t>(ncols * 4))
                   : "memory");
    else
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + static_cast<size_t>(row) * ld),
                   "r"(uz::smem_u32(srow)), "r"(static_cast<uint32_t>(ncols * 4))
                   : "memory");
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the staging row may be overwritten again
  UZ_TRACE(trace, et == 0 ? 12 : 15);
}

template <int PIX>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t box_bytes = PIX * 128;                       // [PIX pixels][64 ch] bf16
  const uint32_t a_bytes = p.a_boxes * box_bytes;
  const uint32_t b_bytes = p.b_boxes * box_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x, splits = gridDim.x;
  const int group = blockIdx.y % p.tap_groups;
  const int co0 = ((blockIdx.y / p.tap_groups) % p.co_blocks) * 128;
  const int ci0 = (blockIdx.y / (p.tap_groups * p.co_blocks)) * p.ci_w;     // this CTA's input-channel chunk
  int ncin = p.Cin - ci0; if (ncin > p.ci_w) ncin = p.ci_w;
  const int tap0 = group * p.taps_per_cta;
  int ntaps = p.taps - tap0; if (ntaps > p.taps_per_cta) ntaps = p.taps_per_cta;

  int my_tiles = 0;
  if (split < p.num_tiles) my_tiles = (p.num_tiles - split + splits - 1) / splits;
  const int iters = my_tiles * ntaps;
  UZ_TRACE(p.trace, warp == 0 ? 0 : 15);

  if (warp == 0 && lane == 0) {
    uz::tma_prefetch_desc(&tmap_dy);
    uz::tma_prefetch_desc(&tmap_x);
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
  uz::tc_fence_before();
  __syncthreads();
  uz::tc_fence_after();
  const uint32_t tmem_base = uz::uniform_u32(tmem_base_slot);
  UZ_TRACE(p.trace, warp == 0 ? 3 : 15);

  if (warp == 0) {
    {
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        if (it >= p.stages) uz::mbar_wait(&empty_bar[s], ((it / p.stages) - 1) & 1);
        const int ti = it / ntaps;
        const int tap = tap0 + (it - ti * ntaps);
        const int tile = split + ti * splits;
        const int tx = tile % p.tilesW;
        const int ty = (tile / p.tilesW) % p.tilesH;
        const int tn = tile / (p.tilesW * p.tilesH);
        const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = tn * p.TN;
        int dy = 0, dx = 0;
        if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        if (uz::elect_one()) {
          uz::mbar_expect_tx(&full_bar[s], stage_bytes);
          for (int b = 0; b < p.a_boxes; ++b)
            uz::tma_load_4d(sa + b * box_bytes, &tmap_dy, &full_bar[s], co0 + b * 64, x0, y0, n0);
          for (int b = 0; b < p.b_boxes; ++b)      // boxes past Cin are zero-filled by TMA (the last chunk may be narrower)
            uz::tma_load_4d(sb + b * box_bytes, &tmap_x, &full_bar[s], ci0 + b * 64, x0 + dx, y0 + dy, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // MN-major operands: 64-channel runs LBO = box_bytes apart, 8-pixel groups SBO = 1024 B apart.  High descriptor words
    // are invariant; the PIX/16 K steps are unrolled so every MMA owns its uniform registers (see conv_tc2.cu).
    const uint64_t desc_hi = uz::umma_desc(0, box_bytes, 1024, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo0 = static_cast<uint32_t>(uz::umma_desc(0, box_bytes, 1024, 128) & 0xFFFFFFFFull);
    const int n0 = ncin < 256 ? ncin : 256;            // first N chunk
    const int n1 = ncin - n0;                          // second N chunk (channels in (256, 512]) or 0
    const uint32_t idesc0 = uz::umma_idesc_bf16(128, n0, 1, 1);
    const uint32_t idesc1 = uz::umma_idesc_bf16(128, n1 > 0 ? n1 : 16, 1, 1);
    for (int it = 0; it < iters; ++it) {
      const int s = it % p.stages;
      uz::mbar_wait(&full_bar[s], (it / p.stages) & 1);
      uz::tc_fence_after();
      if (it == 0) UZ_TRACE(p.trace, 4);
      if (it == iters - 1) UZ_TRACE(p.trace, 5);
      const int ti = it / ntaps;
      const int tl = it - ti * ntaps;  // local tap index -> accumulator slot
      const uint32_t a_lo = desc_lo0 + (uz::smem_u32(smem + s * stage_bytes) >> 4);
      const uint32_t b_lo = a_lo + (a_bytes >> 4);
      const uint32_t dcol = tmem_base + tl * p.ci_w;
      if (uz::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < PIX / 16; ++ks) {
          const uint64_t adesc = desc_hi | (a_lo + ks * (2048 >> 4));
          const uint32_t acc = ks == 0 ? static_cast<uint32_t>(ti != 0) : 1u;
          uz::tc_mma_f16(dcol, adesc, desc_hi | (b_lo + ks * (2048 >> 4)), idesc0, acc);
          if (n1 > 0)
            uz::tc_mma_f16(dcol + 256, adesc, desc_hi | (b_lo + 4 * (box_bytes >> 4) + ks * (2048 >> 4)), idesc1, acc);
        }
        uz::tc_commit(&empty_bar[s]);
        if (it == iters - 1) uz::tc_commit(&accum_bar);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;            // co within the block
    if (iters > 0) {
      uz::mbar_wait(&accum_bar, 0);
      uz::tc_fence_after();
    }
    UZ_TRACE(p.trace, warp == 2 ? 6 : 15);
    int valid_rows = p.Cout - co0;
    if (valid_rows > 128) valid_rows = 128;
    float* stage = reinterpret_cast<float*>(smem);          // pipeline buffers: idle once the accumulators are complete
    for (int tl = 0; tl < ntaps; ++tl) {
      float* dst = p.partial + ((static_cast<size_t>(p.accumulate ? 0 : split) * p.taps + tap0 + tl) * p.Cout + co0) * p.Cin + ci0;
      for (int c0 = 0; c0 < ncin; c0 += 128) {             // column blocks of <= 128 (staging: 128 x 132 floats)
        const int ncols = ncin - c0 < 128 ? ncin - c0 : 128;
        store_acc_block(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tl * p.ci_w + c0, ncols, iters > 0, stage,
                        dst + c0, p.Cin, valid_rows, row, threadIdx.x - 64, p.accumulate != 0,
#ifdef UZ_PROFILE_KNOBS
                        (tl == 0 && c0 == 0) ? p.trace : nullptr
#else
                        nullptr
#endif
        );
      }
    }
  }

  if (warp >= 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // bulk stores of the epilogue have landed
  UZ_TRACE(p.trace, warp == 2 ? 7 : 15);
  uz::tc_fence_before();
  __syncthreads();
  UZ_TRACE(p.trace, warp == 0 ? 8 : 15);
  if (warp == 1) {
    uz::tc_fence_after();
    uz::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second-generation weight-gradient kernel for 3x3 convs on feature maps with H % 8 == 0 and W % 16 == 0.
// The generic kernel above re-loads the dy tile and a shifted x tile for every tap: 128 B of operands per tensor-core
// cycle per SM, three times what L2 delivers.  Here a CTA owns (128 output channels, one dx, <= 128 input channels):
// per 8x16-pixel tile it loads the dy tile ONCE and ONE x halo slab (10 rows x 16 pixels, shifted by dx); the three dy
// taps are MMAs on row-offset views of the slab (MN-major operands: K = pixel rows, a dy shift is 16 rows = 2 swizzle
// atoms).  Operand traffic drops to ~45 B per tensor-core cycle.  Three [128 x <=128] fp32 accumulators live in TMEM.
// CTAs a weight-gradient launch aims for.  When the caller runs weight gradients next to the dgrad chain on auxiliary
// streams (b200/ops.py) a launch does not need every SM, and fewer pixel splits mean fewer fp32 partial slabs to write
// and reduce: uz_set_wgrad_sm_percent(25) gave +9 % step throughput on PHiSeg-7/5.  Volumes run one stream: all SMs.
int g_wgrad_sm_percent = 100;       // uz_set_wgrad_sm_percent()
inline int wgrad_target_ctas(bool vol) {
  const int sms = uz::num_sms();
  return vol ? sms : (sms * g_wgrad_sm_percent + 99) / 100;
}

struct Wgrad2Params {
  int N, D, H, W, Cin, Cout;   // D = 1 for 2-D maps
  int nz;                      // z taps: 1 (2-D) or 3 (volumes); tensor maps are always 5-D (C, W, H, D, N)
  int tilesW, tilesH, num_tiles;
  int ci_chunks, co_blocks;
  int ci_w;                    // input channels per CTA: 128, or 64 when the layer has few pixel tiles (output-bound)
  int a_boxes;
  int stages;
  float* partial;           // [splits][9 * nz][Cout][Cin]
  int accumulate;           // 1: every split ADDS into slab 0 (zero on entry) with bulk reduce-add stores
};

constexpr int kW2Pix = 128;                 // 8 rows x 16 pixels
constexpr uint32_t kW2ABox = kW2Pix * 128;  // dy box  [128 px][64 ch]
constexpr uint32_t kW2BBox = 160 * 128;     // x slab  [10 rows x 16 px][64 ch]

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                 const Wgrad2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x, splits = gridDim.x;
  // blockIdx.y -> (co block, (dz, dx), ci chunk)
  const int cc = blockIdx.y % p.ci_chunks;
  const int dxz = (blockIdx.y / p.ci_chunks) % (3 * p.nz);
  const int dx = dxz % 3, dz = dxz / 3;
  const int z_off = p.nz >> 1;
  const int co0 = (blockIdx.y / (p.ci_chunks * 3 * p.nz)) * 128;
  const int ci0 = cc * p.ci_w;
  int nch = p.Cin - ci0; if (nch > p.ci_w) nch = p.ci_w;   // input channels of this CTA (multiple of 16)
  const int b_boxes = (nch + 63) / 64;
  const uint32_t a_bytes = p.a_boxes * kW2ABox;
  const uint32_t stage_bytes = a_bytes + 2 * kW2BBox;     // slab area sized for two boxes
  const uint32_t tx_bytes = a_bytes + b_boxes * kW2BBox;
  int my_tiles = 0;
  if (split < p.num_tiles) my_tiles = (p.num_tiles - split + splits - 1) / splits;

  if (warp == 0 && lane == 0) {
    uz::tma_prefetch_desc(&tmap_dy);
    uz::tma_prefetch_desc(&tmap_x);
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
    uz::tmem_alloc(&tmem_base_slot, 512);
  }
  uz::pdl_prologue();   // everything above is independent of the previous kernel's output
  uz::tc_fence_before();
  __syncthreads();
  uz::tc_fence_after();
  const uint32_t tmem_base = uz::uniform_u32(tmem_base_slot);

  if (warp == 0) {
    int stage = 0, phase = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int tile = split + ti * splits;
      const int x0 = (tile % p.tilesW) * 16;
      const int y0 = ((tile / p.tilesW) % p.tilesH) * 8;
      const int zn = tile / (p.tilesW * p.tilesH);
      const int z0 = zn % p.D;
      const int n = zn / p.D;
      uz::mbar_wait(&empty_bar[stage], phase ^ 1);
      if (uz::elect_one()) {
        uint8_t* sa = smem + stage * stage_bytes;
        uz::mbar_expect_tx(&full_bar[stage], tx_bytes);
        for (int b = 0; b < p.a_boxes; ++b)
          uz::tma_load_5d(sa + b * kW2ABox, &tmap_dy, &full_bar[stage], co0 + b * 64, x0, y0, z0, n);
        for (int b = 0; b < b_boxes; ++b)
          uz::tma_load_5d(sa + a_bytes + b * kW2BBox, &tmap_x, &full_bar[stage], ci0 + b * 64, x0 + dx - 1, y0 - 1,
                          z0 + dz - z_off, n);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint64_t a_hi = uz::umma_desc(0, kW2ABox, 1024, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lo0 = static_cast<uint32_t>(uz::umma_desc(0, kW2ABox, 1024, 128) & 0xFFFFFFFFull);
    const uint64_t b_hi = uz::umma_desc(0, kW2BBox, 1024, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t b_lo0 = static_cast<uint32_t>(uz::umma_desc(0, kW2BBox, 1024, 128) & 0xFFFFFFFFull);
    const uint32_t idesc = uz::umma_idesc_bf16(128, nch, 1, 1);
    int stage = 0, phase = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      uz::mbar_wait(&full_bar[stage], phase);
      uz::tc_fence_after();
      const uint32_t sa = uz::smem_u32(smem + stage * stage_bytes);
      const uint32_t a_lo = a_lo0 + (sa >> 4);
      const uint32_t b_lo = b_lo0 + ((sa + a_bytes) >> 4);
      if (uz::elect_one()) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int ks = 0; ks < kW2Pix / 16; ++ks) {
            const uint32_t acc = ks == 0 ? static_cast<uint32_t>(ti != 0) : 1u;
            uz::tc_mma_f16(tmem_base + dy * 128, a_hi | (a_lo + ks * 128), b_hi | (b_lo + (dy + ks) * 128), idesc, acc);
          }
        }
        uz::tc_commit(&empty_bar[stage]);
        if (ti == my_tiles - 1) uz::tc_commit(&accum_bar);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    if (my_tiles > 0) {
      uz::mbar_wait(&accum_bar, 0);
      uz::tc_fence_after();
    }
    int valid_rows = p.Cout - co0;
    if (valid_rows > 128) valid_rows = 128;
    float* stage = reinterpret_cast<float*>(smem);          // pipeline buffers: idle once the accumulators are complete
    for (int dy = 0; dy < 3; ++dy) {
      const int tap = (dz * 3 + dy) * 3 + dx;      // OI(D)HW order (kd*3 + kh)*3 + kw; dz == 0 in 2-D
      float* dst = p.partial + ((static_cast<size_t>(p.accumulate ? 0 : split) * 9 * p.nz + tap) * p.Cout + co0) * p.Cin + ci0;
      store_acc_block(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + dy * 128, nch, my_tiles > 0, stage, dst, p.Cin,
                      valid_rows, q * 32 + lane, threadIdx.x - 64, p.accumulate != 0);
    }
  }

  if (warp >= 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // bulk stores of the epilogue have landed
  uz::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    uz::tc_fence_after();
    uz::tmem_dealloc(tmem_base, 512);
  }
}

struct Plan2 {
  Wgrad2Params p;
  int splits;
  size_t smem;
};

// D == 0: 2-D map, 9 taps (H % 8 == 0, W % 16 == 0); D >= 1: volume, 27 taps, any H / W (out-of-range dy rows load as zeros)
bool make_plan2(int N, int D, int H, int W, int Cin, int Cout, int taps, Plan2* out) {
  const bool vol = D > 0;
  if (taps != (vol ? 27 : 9) || Cin % 16 || Cout % 16) return false;
  if (!vol && (H % 8 || W % 16)) return false;
  if (!vol && UZ_KNOB(32)) return false;
  Wgrad2Params& p = out->p;
  p = Wgrad2Params{};
  p.N = N; p.D = vol ? D : 1; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.nz = vol ? 3 : 1;
  p.tilesW = (W + 15) / 16; p.tilesH = (H + 7) / 8;
  p.num_tiles = N * p.D * p.tilesW * p.tilesH;
  p.ci_w = (!vol && p.num_tiles <= 32 && Cin > 64) ? 64 : 128;      // few tiles: more, narrower CTAs (see make_plan)
  p.ci_chunks = (Cin + p.ci_w - 1) / p.ci_w;
  p.co_blocks = (Cout + 127) / 128;
  p.a_boxes = Cout > 64 ? 2 : 1;
  const size_t stage_bytes = static_cast<size_t>(p.a_boxes) * kW2ABox + 2 * kW2BBox;
  int stages = static_cast<int>((196 * 1024) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return false;
  p.stages = stages;
  out->smem = stages * stage_bytes + 1024;
  const int per_split = p.co_blocks * 3 * p.nz * p.ci_chunks;
  int splits = wgrad_target_ctas(vol) / per_split;
  if (splits > p.num_tiles) splits = p.num_tiles;
  if (splits < 1) splits = 1;
  out->splits = splits;
  return true;
}

// dw[o][i][t] = sum_s partial[s][t][o][i], deterministic.  Block = (256 / G) (o, i) pairs x G split groups, blockIdx.y =
// tap: consecutive threads of a group read consecutive floats of a partial slab (coalesced); the G groups walk the
// splits in parallel (layers with few channels have up to 49 splits and too few pairs to hide the load latency
// otherwise) and are combined through shared memory in a fixed order.  G = 1, 2, 4 or 8 so that a thread sums >= ~6 terms.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps,
                                                           int CoutP, int CinP, int Cout, int Cin, float* dw, int G) {
  uz::pdl_prologue();
  __shared__ float red[256 + 8];
  const int ppb = 256 / G;                       // pairs per block
  const int lane = threadIdx.x % ppb, grp = threadIdx.x / ppb;
  const size_t pairs = static_cast<size_t>(Cout) * Cin;
  const size_t slab = static_cast<size_t>(CoutP) * CinP;
  const int t = blockIdx.y;
  for (size_t base = static_cast<size_t>(blockIdx.x) * ppb; base < pairs; base += static_cast<size_t>(gridDim.x) * ppb) {
    const size_t idx = base + lane;
    float acc = 0.f;
    if (idx < pairs) {
      const int i = static_cast<int>(idx % Cin);
      const int o = static_cast<int>(idx / Cin);
      const float* src = partial + static_cast<size_t>(t) * slab + static_cast<size_t>(o) * CinP + i;
#pragma unroll 4
      for (int s = grp; s < splits; s += G) acc += src[static_cast<size_t>(s) * taps * slab];
    }
    if (G == 1) {
      if (idx < pairs) dw[idx * taps + t] = acc;
      continue;
    }
    red[grp * (ppb + 1) + lane] = acc;
    __syncthreads();
    if (grp == 0 && idx < pairs) {
      float tot = red[lane];
      for (int g = 1; g < G; ++g) tot += red[g * (ppb + 1) + lane];
      dw[idx * taps + t] = tot;
    }
    __syncthreads();
  }
}

// The same reduction for MANY layers in one launch (uz_wgrad_reduce_batched): a work unit is one output channel o and a
// block of 64 input channels; the block's 64 lanes x 4 tap groups read the [split][tap][o][i] partials (256-byte rows,
// fully coalesced), sum the splits in a fixed order, transpose through shared memory and write the unit's contiguous
// run dw[o][i0 .. i0+63][0 .. taps-1] coalesced (the per-layer kernel above writes every float with a stride of `taps`).
// The descriptor rows travel as a kernel PARAMETER (<= 64 rows = 3 KB): a captured CUDA graph keeps them by value, no
// device table to upload or keep alive.
struct ReduceBatch {
  UzWgradReduceDesc d[UZ_WGRAD_REDUCE_MAX_ROWS];
  int n;
};

// rows (output channels) per work unit: enough bytes per block to amortise its start-up (a unit of one row moved 2.3 KB
// and the 43 712 blocks of a PHiSeg step ran at 1.3 TB/s), bounded by the shared-memory tile for 27-tap volume layers
__host__ __device__ inline int reduce_rows_per_unit(int taps) { return taps <= 9 ? 8 : 2; }

__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(const __grid_constant__ ReduceBatch b) {
  uz::pdl_prologue();
  __shared__ float tile[8 * 64 * 9];             // [rows][64 ci][taps]; 2 rows x 27 taps fits as well
  int lo = 0, hi = b.n - 1;                      // last descriptor whose first unit is <= blockIdx.x (block-uniform)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.d[mid].unit_begin <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const UzWgradReduceDesc d = b.d[lo];
  const int R = reduce_rows_per_unit(d.taps);
  const int unit = blockIdx.x - d.unit_begin;
  const int chunks = (d.Cin + 63) / 64;
  const int o0 = (unit / chunks) * R, i0 = (unit % chunks) * 64;
  const int lane = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int i = i0 + lane;
  const size_t slab = static_cast<size_t>(d.CoutP) * d.CinP;
  const size_t split_stride = static_cast<size_t>(d.taps) * slab;
  const int nrows = d.Cout - o0 < R ? d.Cout - o0 : R;
  const int ni = d.Cin - i0 < 64 ? d.Cin - i0 : 64;
  if (i < d.Cin) {
    constexpr int U = 6;                         // (row, tap) pairs in flight per thread: the loads are 600 ns apiece
    const int total = nrows * d.taps;
    for (int rt0 = grp; rt0 < total; rt0 += 4 * U) {
      float acc[U];
      int dstidx[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rt = rt0 + 4 * u;
        acc[u] = 0.f;
        dstidx[u] = -1;
        if (rt < total) {
          const int r = rt / d.taps, t = rt - r * d.taps;
          const float* src = d.partial + static_cast<size_t>(t) * slab + static_cast<size_t>(o0 + r) * d.CinP + i;
          dstidx[u] = (r * ni + lane) * d.taps + t;
          float a = src[0];                      // splits summed strictly in order
          for (int s = 1; s < d.splits; ++s) a += src[static_cast<size_t>(s) * split_stride];
          acc[u] = a;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (dstidx[u] >= 0) tile[dstidx[u]] = acc[u];
    }
  }
  __syncthreads();
  // row r of the unit is the contiguous run dw[o0 + r][i0 .. i0 + ni - 1][0 .. taps - 1]
  const int run = ni * d.taps;
  for (int e = threadIdx.x; e < nrows * run; e += 256) {
    const int r = e / run, k = e - r * run;
    d.dw[(static_cast<size_t>(o0 + r) * d.Cin + i0) * d.taps + k] = tile[e];
  }
}

inline int reduce_groups(int splits) { return splits >= 32 ? 8 : (splits >= 16 ? 4 : (splits >= 8 ? 2 : 1)); }

int pow2_div_le(int v, int cap) {
  int t = 1;
  while (t * 2 <= cap && v % (t * 2) == 0) t *= 2;
  return t;
}

struct Plan {
  WgradParams p;
  int splits, co_blocks;
  size_t smem;
};

int make_plan(int N, int H, int W, int Cin, int Cout, int taps, Plan* out) {
  WgradParams& p = out->p;
  p = WgradParams{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.taps = taps;
  p.PIX = Cin <= 128 ? 128 : 64;
  p.TW = pow2_div_le(W, 16);
  p.TH = pow2_div_le(H, p.PIX / p.TW);
  p.TN = p.PIX / (p.TW * p.TH);
  p.tilesW = W / p.TW; p.tilesH = H / p.TH;
  p.num_tiles = p.tilesW * p.tilesH * ((N + p.TN - 1) / p.TN);
  // Tiny layers (a handful of pixel tiles) are bound by WRITING their [tap][co][ci] result through one SM's store path
  // (~35 GB/s per SM measured, profiles/r02_phase_trace.md), not by the MMAs: they get one tap and 64 input channels
  // per CTA (9 x more, 3 x narrower CTAs: every CTA writes 32 KB instead of 196 KB) and no pixel split.
  const bool tiny = taps == 9 && p.num_tiles <= 8 && Cin > 64;
  p.ci_w = tiny ? 64 : Cin;
  p.ci_chunks = (Cin + p.ci_w - 1) / p.ci_w;
  int tpc = tiny ? 1 : 512 / Cin; if (tpc < 1) return UZ_ERR_ARG;
  if (tpc > taps) tpc = taps;
  if (taps == 9 && tpc >= 3 && tpc < 9) tpc = 3;   // balanced groups of three
  p.taps_per_cta = tpc;
  p.tap_groups = (taps + tpc - 1) / tpc;
  p.a_boxes = Cout > 64 ? 2 : 1;
  p.b_boxes = (p.ci_w + 63) / 64;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(tpc * p.ci_w)) cols *= 2;
  p.tmem_cols = cols;
  const size_t stage_bytes = static_cast<size_t>(p.a_boxes + p.b_boxes) * p.PIX * 128;
  int stages = static_cast<int>((196 * 1024) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 1) return UZ_ERR_ARG;
  p.stages = stages;
  out->smem = stages * stage_bytes + 1024;
  out->co_blocks = (Cout + 127) / 128;
  p.co_blocks = out->co_blocks;
  const int per_split = p.tap_groups * out->co_blocks * p.ci_chunks;
  int splits = tiny ? 1 : (wgrad_target_ctas(false) + per_split - 1) / per_split;
  if (splits > p.num_tiles) splits = p.num_tiles;
  if (splits < 1) splits = 1;
  out->splits = splits;
  return UZ_OK;
}

}  // namespace

extern "C" int uz_set_wgrad_sm_percent(int percent) {
  UZ_CHECK_ARG(percent >= 5 && percent <= 100, "uz_set_wgrad_sm_percent: %d outside [5, 100]", percent);
  g_wgrad_sm_percent = percent;
  return UZ_OK;
}

extern "C" long long uz_wgrad3d_workspace_floats(int N, int D, int H, int W, int Cin, int Cout) {
  Plan2 pl2;
  if (D > 0 && Cin > 0 && Cout > 0 && make_plan2(N, D, H, W, Cin, Cout, 27, &pl2))
    return static_cast<long long>(pl2.splits) * 27 * Cout * Cin;
  return -1;
}

extern "C" long long uz_wgrad_workspace_floats(int N, int H, int W, int Cin, int Cout, int taps) {
  Plan2 pl2;
  if (Cin > 0 && Cout > 0 && make_plan2(N, 0, H, W, Cin, Cout, taps, &pl2))
    return static_cast<long long>(pl2.splits) * taps * Cout * Cin;
  Plan pl;
  if (Cin % 16 || Cout % 16 || Cin <= 0 || Cout <= 0 || Cin > 512 || make_plan(N, H, W, Cin, Cout, taps, &pl)) return -1;
  return static_cast<long long>(pl.splits) * taps * Cout * Cin;
}

namespace {
int wgrad_impl(const void* x, int ldx, const void* dy, int lddy, int N, int D, int H, int W, int Cin, int Cout, int taps,
               int Cin_logical, int Cout_logical, float* workspace, float* dw, void* stream, int* splits_out = nullptr,
               int accumulate = 0);
}

// The tensor-core part alone: writes the split-K partial slabs [splits][taps][Cout][Cin] fp32 into `workspace`
// (uz_wgrad_workspace_floats) and reports the number of splits.  The slabs of many layers are then reduced, transposed to
// OIHW and written to their gradient tensors by ONE uz_wgrad_reduce_batched launch.  D == 0: images, D > 0: volumes (27 taps).
extern "C" int uz_conv_wgrad_partial(const void* x, int ldx, const void* dy, int lddy, int N, int D, int H, int W,
                                     int Cin, int Cout, int taps, float* workspace, int accumulate, int* splits,
                                     void* stream) {
  UZ_CHECK_ARG(splits, "uz_conv_wgrad_partial: null pointer");
  UZ_CHECK_ARG(D == 0 ? (taps == 9 || taps == 1) : taps == 27, "uz_conv_wgrad_partial: taps %d with D %d", taps, D);
  return wgrad_impl(x, ldx, dy, lddy, N, D, H, W, Cin, Cout, taps, Cin, Cout, workspace, nullptr, stream, splits,
                    accumulate ? 1 : 0);
}

extern "C" int uz_wgrad_reduce_units(int Cout_logical, int Cin_logical, int taps) {
  const int R = reduce_rows_per_unit(taps);
  return ((Cout_logical + R - 1) / R) * ((Cin_logical + 63) / 64);
}

// descs: n <= UZ_WGRAD_REDUCE_MAX_ROWS rows in HOST memory (copied into the launch parameters); unit_begin is filled in
// here.  Deterministic (fixed split order).
extern "C" int uz_wgrad_reduce_batched(const UzWgradReduceDesc* descs, int n, void* stream) {
  UZ_CHECK_ARG(descs && n > 0 && n <= UZ_WGRAD_REDUCE_MAX_ROWS, "uz_wgrad_reduce_batched: 1..%d rows (got %d)",
               UZ_WGRAD_REDUCE_MAX_ROWS, n);
  if UZ_KNOB(4096) return UZ_OK;
  ReduceBatch b;
  b.n = n;
  int units = 0;
  for (int k = 0; k < n; ++k) {
    b.d[k] = descs[k];
    UZ_CHECK_ARG(b.d[k].partial && b.d[k].dw && b.d[k].splits > 0 && b.d[k].taps > 0 && b.d[k].taps <= 27 &&
                     b.d[k].Cout > 0 && b.d[k].Cin > 0 && b.d[k].Cout <= b.d[k].CoutP && b.d[k].Cin <= b.d[k].CinP,
                 "uz_wgrad_reduce_batched: bad row %d", k);
    b.d[k].unit_begin = units;
    units += uz_wgrad_reduce_units(b.d[k].Cout, b.d[k].Cin, b.d[k].taps);
  }
  uz::launch(wgrad_reduce_batched_kernel, dim3(units, 1, 1), 256, 0, static_cast<cudaStream_t>(stream), b);
  UZ_CHECK_LAUNCH("uz_wgrad_reduce_batched");
  return UZ_OK;
}



// x: bf16 NHWC [N,H,W,Cin] (ldx), dy: bf16 NHWC [N,H,W,Cout] (lddy); dw: fp32 [Cout_logical][Cin_logical][taps].
extern "C" int uz_conv_wgrad(const void* x, int ldx, const void* dy, int lddy, int N, int H, int W, int Cin, int Cout,
                             int taps, int Cin_logical, int Cout_logical, float* workspace, float* dw, void* stream) {
  UZ_CHECK_ARG(taps == 9 || taps == 1, "uz_conv_wgrad: taps must be 9 or 1");
  return wgrad_impl(x, ldx, dy, lddy, N, 0, H, W, Cin, Cout, taps, Cin_logical, Cout_logical, workspace, dw, stream);
}

// volumes: x bf16 NDHWC [N,D,H,W,Cin], dy [N,D,H,W,Cout]; dw fp32 [Cout_logical][Cin_logical][27] (OIDHW)
extern "C" int uz_conv3d_wgrad(const void* x, int ldx, const void* dy, int lddy, int N, int D, int H, int W, int Cin,
                               int Cout, int Cin_logical, int Cout_logical, float* workspace, float* dw, void* stream) {
  UZ_CHECK_ARG(D > 0, "uz_conv3d_wgrad: D must be positive");
  return wgrad_impl(x, ldx, dy, lddy, N, D, H, W, Cin, Cout, 27, Cin_logical, Cout_logical, workspace, dw, stream);
}

namespace {
// splits_out != nullptr: only the tensor-core kernel runs; the partial slabs stay in `workspace` for uz_wgrad_reduce_batched
int wgrad_impl(const void* x, int ldx, const void* dy, int lddy, int N, int D, int H, int W, int Cin, int Cout, int taps,
               int Cin_logical, int Cout_logical, float* workspace, float* dw, void* stream, int* splits_out,
               int accumulate) {
  UZ_CHECK_ARG(x && dy && workspace && (dw || splits_out), "uz_conv_wgrad: null pointer");
  UZ_CHECK_ARG(!accumulate || splits_out, "uz_conv_wgrad: accumulation needs the deferred reduction");
  UZ_CHECK_ARG(Cin % 16 == 0 && Cout % 16 == 0 && Cin > 0 && Cout > 0 && Cin <= 512,
               "uz_conv_wgrad: channels must be multiples of 16, Cin <= 512 (got %d, %d)", Cin, Cout);
  UZ_CHECK_ARG(ldx % 8 == 0 && lddy % 8 == 0 && ldx >= Cin && lddy >= Cout, "uz_conv_wgrad: bad pixel strides");
  UZ_CHECK_ARG(Cin_logical <= Cin && Cout_logical <= Cout, "uz_conv_wgrad: logical dims exceed stored dims");
  if UZ_KNOB(256) return UZ_OK;   // measurement knob: step time without the wgrad kernels
  Plan2 pl2;
  if (make_plan2(N, D, H, W, Cin, Cout, taps, &pl2)) {
    pl2.p.partial = workspace;
    pl2.p.accumulate = accumulate;
    const uint64_t Dd = static_cast<uint64_t>(pl2.p.D);
    CUtensorMap tdy2, tx2;
    {
      uint64_t dims[5] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(W), static_cast<uint64_t>(H), Dd,
                          static_cast<uint64_t>(N)};
      uint64_t strides[4] = {static_cast<uint64_t>(lddy) * 2, static_cast<uint64_t>(W) * lddy * 2,
                             static_cast<uint64_t>(H) * W * lddy * 2, Dd * H * W * lddy * 2};
      uint32_t box[5] = {64, 16, 8, 1, 1};
      int rc2 = uz::make_tmap_bf16(&tdy2, dy, 5, dims, strides, box, 128);
      if (rc2) return rc2;
    }
    {
      uint64_t dims[5] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H), Dd,
                          static_cast<uint64_t>(N)};
      uint64_t strides[4] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(W) * ldx * 2,
                             static_cast<uint64_t>(H) * W * ldx * 2, Dd * H * W * ldx * 2};
      uint32_t box[5] = {64, 16, 10, 1, 1};
      int rc2 = uz::make_tmap_bf16(&tx2, x, 5, dims, strides, box, 128);
      if (rc2) return rc2;
    }
    static size_t attr2 = 0;
    if (pl2.smem > attr2) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(pl2.smem));
      if (e != cudaSuccess) {
        (void)cudaGetLastError();
        uz::set_error("uz_conv_wgrad(v2): cannot raise dynamic smem limit: %s", cudaGetErrorString(e));
        return UZ_ERR_CUDA;
      }
      attr2 = pl2.smem;
    }
    dim3 grid2(pl2.splits, pl2.p.co_blocks * 3 * pl2.p.nz * pl2.p.ci_chunks, 1);
    uz::launch(wgrad_tc2_kernel, grid2, kThreads, pl2.smem, static_cast<cudaStream_t>(stream), tdy2, tx2, pl2.p);
    UZ_CHECK_LAUNCH("uz_conv_wgrad(v2)");
    if (splits_out) { *splits_out = accumulate ? 1 : pl2.splits; return UZ_OK; }
    const size_t total2 = static_cast<size_t>(Cout_logical) * Cin_logical;     // one thread per (o, i) pair
    const int G2 = reduce_groups(pl2.splits);
    int blocks2 = static_cast<int>((total2 + 256 / G2 - 1) / (256 / G2));
    if (blocks2 > uz::num_sms() * 8) blocks2 = uz::num_sms() * 8;
    if (!UZ_KNOB(4096)) uz::launch(wgrad_reduce_kernel, dim3(blocks2, taps, 1), 256, 0, static_cast<cudaStream_t>(stream), workspace, pl2.splits, taps, Cout, Cin,
                                                                               Cout_logical, Cin_logical, dw, G2);
    UZ_CHECK_LAUNCH("uz_conv_wgrad(v2 reduce)");
    return UZ_OK;
  }
  UZ_CHECK_ARG(D == 0, "uz_conv3d_wgrad: unsupported shape Cin=%d Cout=%d", Cin, Cout);
  Plan pl;
  int rc = make_plan(N, H, W, Cin, Cout, taps, &pl);
  UZ_CHECK_ARG(rc == UZ_OK, "uz_conv_wgrad: no plan for Cin=%d Cout=%d", Cin, Cout);
  pl.p.partial = workspace;
  pl.p.accumulate = accumulate;
#ifdef UZ_PROFILE_KNOBS
  pl.p.trace = uz::g_trace;
#endif
  CUtensorMap tdy, tx;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(lddy) * 2, static_cast<uint64_t>(W) * lddy * 2,
                           static_cast<uint64_t>(H) * W * lddy * 2};
    uint32_t box[4] = {64, static_cast<uint32_t>(pl.p.TW), static_cast<uint32_t>(pl.p.TH),
                       static_cast<uint32_t>(pl.p.TN)};
    rc = uz::make_tmap_bf16(&tdy, dy, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(W) * ldx * 2,
                           static_cast<uint64_t>(H) * W * ldx * 2};
    uint32_t box[4] = {64, static_cast<uint32_t>(pl.p.TW), static_cast<uint32_t>(pl.p.TH),
                       static_cast<uint32_t>(pl.p.TN)};
    rc = uz::make_tmap_bf16(&tx, x, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  auto kernel = pl.p.PIX == 128 ? wgrad_tc_kernel<128> : wgrad_tc_kernel<64>;
  static size_t attr_bytes_pix[2] = {0, 0};
  size_t& attr_bytes = attr_bytes_pix[pl.p.PIX == 128 ? 0 : 1];
  if (pl.smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      uz::set_error("uz_conv_wgrad: cannot raise dynamic smem limit to %zu: %s", static_cast<size_t>(pl.smem), cudaGetErrorString(e));
      return UZ_ERR_CUDA;
    }
    attr_bytes = pl.smem;
  }
  dim3 grid(pl.splits, pl.p.tap_groups * pl.co_blocks * pl.p.ci_chunks, 1);
  uz::launch(kernel, grid, kThreads, pl.smem, static_cast<cudaStream_t>(stream), tdy, tx, pl.p);
  UZ_CHECK_LAUNCH("uz_conv_wgrad");
  if (splits_out) { *splits_out = accumulate ? 1 : pl.splits; return UZ_OK; }
  const size_t total = static_cast<size_t>(Cout_logical) * Cin_logical;       // one thread per (o, i) pair
  const int G1 = reduce_groups(pl.splits);
  int blocks = static_cast<int>((total + 256 / G1 - 1) / (256 / G1));
  if (blocks > uz::num_sms() * 8) blocks = uz::num_sms() * 8;
  if (!UZ_KNOB(4096)) uz::launch(wgrad_reduce_kernel, dim3(blocks, taps, 1), 256, 0, static_cast<cudaStream_t>(stream), workspace, pl.splits, taps, Cout, Cin,
                                                                            Cout_logical, Cin_logical, dw, G1);
  UZ_CHECK_LAUNCH("uz_conv_wgrad(reduce)");
  return UZ_OK;
}
}  // namespace
