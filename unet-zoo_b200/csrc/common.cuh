// Shared device helpers for the sm_100a kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 / TMEM
// wrappers in inline PTX, and small utilities.  No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

// kind::f16 operand format of the tcgen05 instruction descriptor: 1 = BF16 (default build), 0 = F16 (-DUZ_ACT_FP16)
#ifdef UZ_ACT_FP16
#define UZ_MMA_OPERAND_FORMAT 0u
#else
#define UZ_MMA_OPERAND_FORMAT 1u
#endif

#define UZ_OK 0
#define UZ_ERR_ARG 1
#define UZ_ERR_CUDA 2
#define UZ_ERR_DRIVER 3

// Profiling knobs (uz_set_debug_flags): bits that elide loads / MMAs / epilogues / whole launches produce garbage and
// exist only in builds with -DUZ_PROFILE_KNOBS (make PROFILE=1).  The product library keeps bit 32 alone (route 3x3 layers
// to the generic kernel: correct results, used by the kernel-vs-kernel parity test).
#ifdef UZ_PROFILE_KNOBS
#define UZ_KNOB(bits) (uz::g_conv_debug_flags & (bits))
#define UZ_DBG(p, bits) ((p).dbg & (bits))
#else
#define UZ_KNOB(bits) (uz::g_conv_debug_flags & (bits) & 32)
#define UZ_DBG(p, bits) 0
#endif

namespace uz {
extern int g_conv_debug_flags;
#ifdef UZ_PROFILE_KNOBS
// phase timestamps (globaltimer, ns) of CTA 0 of the traced kernels: uz_set_trace_buffer(ptr to 16 x uint64, device)
extern unsigned long long* g_trace;
#define UZ_TRACE(ptr, slot)                                                                   \
  do {                                                                                        \
    if ((ptr) && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0) {             \
      unsigned long long t__;                                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                 \
      (ptr)[slot] = t__;                                                                      \
    }                                                                                         \
  } while (0)
#else
#define UZ_TRACE(ptr, slot) do { } while (0)
#endif

// last error text, readable through uz_last_error()
void set_error(const char* fmt, ...);
void count_launch();

#define UZ_CHECK_ARG(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      uz::set_error(__VA_ARGS__);              \
      return UZ_ERR_ARG;                       \
    }                                          \
  } while (0)

#define UZ_CHECK_LAUNCH(name)                                                   \
  do {                                                                          \
    uz::count_launch();                                                         \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      uz::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return UZ_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of this library can be launched with cudaLaunchAttributeProgrammaticStreamSerialization (uz::launch
// below, opt-in) and starts with pdl_prologue(): `griddepcontrol.wait` blocks until the preceding kernel in the stream has completed
// and its memory is visible, `griddepcontrol.launch_dependents` lets the next kernel's CTAs be scheduled early.  The
// launch latency and the input-independent prologue (barrier init, TMEM allocation, descriptor prefetch) of kernel N+1
// thereby overlap the tail of kernel N.  Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_wait();
  pdl_trigger();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp; keeps the surrounding code warp-uniform so descriptors stay in uniform registers
// (a divergent `if (lane == 0)` makes ptxas wrap every UTCHMMA / UTMALDG in an R2UR waterfall loop: ~180 cycles/MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- thread-block clusters / distributed shared memory
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// address of the same shared-memory variable in CTA `rank` of the cluster, and a load through it
__device__ __forceinline__ uint32_t cluster_map(const void* smem_ptr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(smem_ptr)), "r"(rank));
  return out;
}
__device__ __forceinline__ float cluster_ld_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers bf16 inputs with fp32 accumulation.
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i of the warp = lane base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100 format, version 1).
//   swizzle_bytes in {32, 64, 128}; start address may be advanced by multiples of 16 B inside a swizzle row.
//   K-major : rows of swizzle_bytes, 8-row groups SBO apart (LBO unused).
//   MN-major: 64-element (swizzle_bytes) runs along M/N, K rows swizzle_bytes apart, 8-K-row groups SBO apart,
//             successive M/N runs LBO apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t swizzle_bytes) {
  uint64_t layout = swizzle_bytes == 128 ? 2ull : (swizzle_bytes == 64 ? 4ull : (swizzle_bytes == 32 ? 6ull : 0ull));
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}
// instruction descriptor, kind::f16, bf16 x bf16 -> fp32
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;   // D format fp32
  d |= UZ_MMA_OPERAND_FORMAT << 7;   // A: bf16, or f16 in the UZ_ACT_FP16 build
  d |= UZ_MMA_OPERAND_FORMAT << 10;  // B likewise
  d |= (a_mn_major & 1u) << 15;
  d |= (b_mn_major & 1u) << 16;
  d |= (N >> 3) << 17;
  d |= (M >> 4) << 24;
  return d;
}

// Storage type of activations and packed weights: bf16 (default build), or -- build with -DUZ_ACT_FP16 ->
// libunetzoo_b200_fp16.so -- IEEE half: 10 mantissa bits like TF32 (8x finer than bf16) for the tolerance-matched parity
// mode.  Both are 2 bytes, so every layout, tensor map and kernel is shared; only these conversions and the operand
// format field of the tcgen05 instruction descriptor differ.  The C++ element type stays __nv_bfloat16 as an opaque
// 2-byte container; the names below keep "bf16" for the default build.
#ifdef UZ_ACT_FP16
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo(uint32_t v) {
  return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xFFFFu)));
}
__device__ __forceinline__ float bf16hi(uint32_t v) {
  return __half2float(__ushort_as_half(static_cast<unsigned short>(v >> 16)));
}
__device__ __forceinline__ __nv_bfloat16 f2act(float v) {
  const unsigned short b = __half_as_ushort(__float2half_rn(v));
  return *reinterpret_cast<const __nv_bfloat16*>(&b);
}
__device__ __forceinline__ float act2f(__nv_bfloat16 x) {
  return __half2float(__ushort_as_half(*reinterpret_cast<const unsigned short*>(&x)));
}
#else
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ __nv_bfloat16 f2act(float v) { return __float2bfloat16(v); }
__device__ __forceinline__ float act2f(__nv_bfloat16 x) { return __bfloat162float(x); }
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace uz

// ---------------------------------------------------------------- host side: launches and tensor maps
namespace uz {
extern int g_pdl;   // uz_set_pdl(); default from the environment variable UZ_PDL (0)
void prepare_kernel(const void* fn);   // once per kernel: shared-memory carveout preference (uz_set_smem_carveout)

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args&&... args) {
  prepare_kernel(reinterpret_cast<const void*>(kernel));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // g_pdl: 1 = every launch, 2 = only light kernels (<= 48 KB dynamic shared memory: the elementwise / reduction
  // kernels), 3 = only the tensor-core kernels (their barrier / TMEM / descriptor prologue overlaps the predecessor's tail)
  cfg.numAttrs = (g_pdl == 1 || (g_pdl == 2 && smem <= 48 * 1024) || (g_pdl == 3 && smem > 48 * 1024)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// same, as a thread-block cluster of `cluster` CTAs (grid dimensions must be multiples of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  dim3 cluster, Args&&... args) {
  prepare_kernel(reinterpret_cast<const void*>(kernel));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster.x;
  attr[0].val.clusterDim.y = cluster.y;
  attr[0].val.clusterDim.z = cluster.z;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_pdl == 1 || (g_pdl == 2 && smem <= 48 * 1024) || (g_pdl == 3 && smem > 48 * 1024)) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Encodes a tiled bf16 tensor map; dims/strides innermost first, strides in BYTES for dims 1..rank-1.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, uint32_t swizzle_bytes);
int num_sms();
}  // namespace uz
