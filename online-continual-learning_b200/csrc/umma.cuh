// umma.cuh -- thin inline-PTX layer over the Blackwell 5th-generation tensor core path
// (tcgen05.mma with TMEM accumulators) for sm_100a, as used by the tensor-core convolution.
//
// Operand layout used throughout: K-major tiles with the 128-byte swizzle.  A tile is
// [rows][32 fp32] = rows x 128 B; row r, 16-byte chunk c lives at
//     (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16)
// from a 1024-byte aligned base (the hardware applies the same XOR to address bits 4..6 from
// bits 7..9).  One tcgen05.mma.kind::tf32 consumes K = 8 fp32 (32 bytes) per row; consecutive
// K steps inside the 128-byte row advance the descriptor start address by 32 bytes.
#pragma once
#include <stdint.h>

namespace b200ocl {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- swizzled K-major tile addressing (in floats from the tile base)
__device__ __forceinline__ int sw128_offset_f32(int row, int chunk16) {
  return (row >> 3) * 256 + (row & 7) * 32 + ((chunk16 ^ (row & 7)) << 2);
}

// ---- descriptors
// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;          // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                               // layout type: SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4)                       // D format: F32
         | (2u << 7)                     // A format: TF32
         | (2u << 10)                    // B format: TF32
         | ((uint32_t)(N >> 3) << 17)    // N / 8
         | ((uint32_t)(M >> 4) << 24);   // M / 16
}

// ---- MN-major operands (the contraction index K is the STRIDED one).  Measured on B200 with one-hot probes
// (tests/test_gpu_umma.py keeps the result reproducible): for kind::tf32 an MN-major operand must use layout type 1, SWIZZLE_128B_BASE32B -- with layout
// type 2 (the 16-byte-base swizzle of the K-major tiles above) the instruction completes and writes ZEROS.  Element
// (mn, k) of such an operand lives at
//     start + (mn / 32) * LBO + (k / 4) * SBO + (k % 4) * 128 + (((mn % 32) / 8) ^ row) * 32 + (mn % 8) * 4   bytes,
// row = bits 7..8 of the address: 32 consecutive M/N elements per 128-byte row, the four 32-byte chunks of a row XOR-ed
// with the row index mod 4, one row per k, 4-row k groups SBO bytes apart, 32-element M/N blocks LBO bytes apart; one
// kind::tf32 instruction (K = 8) reads two k groups.  With SBO = 512 the 8 k of an instruction are 8 consecutive rows:
// exactly a "strip" (one 128-byte row of 32 channel slots per pixel position).  A weight gradient contracts over
// positions, so activations and output gradients are both MN-major operands read in place -- and LBO = 128 makes M
// block j the same strip shifted by j rows.  (CUTLASS: mma_sm100_desc.hpp LayoutType::SWIZZLE_128B_BASE32B,
// Swizzle<2,5,2> o ((T,8,m),(4,k)).)
// float offset of (row, 16-byte chunk c16) inside an array of 128-byte rows stored for such operands
__device__ __forceinline__ int sw128b32_offset_f32(int row, int c16) {
  return row * 32 + ((((c16 >> 1) ^ (row & 3)) << 3) | ((c16 & 1) << 2));
}
__device__ __forceinline__ uint64_t make_smem_desc_mn_b32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes = 512) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);           // start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading byte offset: next 32-element M/N block
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // stride byte offset: next 4-row K group
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)1 << 61;                               // layout type: SWIZZLE_128B_BASE32B
  return d;
}
// kind::tf32 instruction descriptor with A and / or B MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc_tf32_major(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM allocation (whole warp)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)),
               "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols));
}

// ---- fences
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::); }
// Bounded wait: returns false if the phase did not complete within max_spins polls (a wrong
// descriptor must not hang the GPU).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t max_spins = (1u << 24)) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1   // ptxas otherwise unrolls the poll 64x at every call site (160 KB of SASS in the conv kernels)
  for (uint32_t i = 0; i < max_spins; ++i) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMA 1-D bulk copy global -> shared (cp.async.bulk, UBLKCP in SASS); completion counted in bytes on an mbarrier.
// dst / src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- one elected lane of a converged warp.  Unlike `lane == 0`, the elect.sync predicate keeps the region
// warp-uniform for ptxas: tcgen05 / bulk-copy instructions inside are emitted once, without the per-active-lane
// ELECT / BRA.U.ANY loops a divergent branch needs around every uniform-datapath instruction.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- MMA issue (one thread) and completion tracking
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes, 20 (16 + 4) or 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- TF32 split: x = hi + lo with hi = x ROUNDED to TF32 (cvt.rna: one instruction) and lo the exact
// fp32 remainder.  Rounding matters: a truncated hi leaves a same-signed remainder whose own truncation
// by the tensor core then accumulates linearly over K (measured 1e-5 relative at K = 736 on ReLU inputs).
// With a rounded hi the remainder is sign-symmetric, so the hardware's truncation of lo to 10 mantissa
// bits (|error| <= 2^-22 |x|) is unbiased and needs no second rounding.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = round_tf32(x);
  lo = x - hi;
}

}  // namespace umma
}  // namespace b200ocl
