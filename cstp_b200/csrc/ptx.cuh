// Thin inline-PTX wrappers for the sm_100a features the CSTP hot path uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and fences.
// Everything here is device-side only and compiled for sm_100a exclusively.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace cstp {

#ifndef CSTP_SPIN_LIMIT
// Bounded mbarrier spin: a pipeline bug must surface as a trap (launch failure), never as a hung GPU.
#define CSTP_SPIN_LIMIT (1u << 26)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > CSTP_SPIN_LIMIT) __trap();
  }
}

// The same for warps that wait for a whole main loop (split-K epilogues): sleep between polls instead of competing with
// the producer / issuer / transform warps for issue slots (ncu: four spinning epilogue warps executed 19 % of all
// instructions of a weight-gradient kernel).
__device__ __forceinline__ void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(512);
    if (++spins > CSTP_SPIN_LIMIT) __trap();
  }
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// The multicast form: the box lands at the same shared-memory offset of every CTA in `mask` (bits = ranks in the cluster)
// and completes `bytes` on the mbarrier at the same offset in each of them.  One L2 read feeds two SMs.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// ---------------------------------------------------------------- thread-block clusters
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
// Every thread of every CTA of the cluster (a cluster of one when the kernel was launched without the attribute).
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
// Allocation is warp-collective (.sync.aligned): call from one full warp.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same two instructions without the "memory" clobber, for tight issue loops: nothing the issuing warp reads or writes
// through ordinary memory operations is touched by the MMA (operands are written by TMA, read by the tensor core), and
// volatile asm statements keep their order among themselves -- but a clobber makes the compiler re-load every kernel
// parameter from the constant bank (LDCU + scoreboard wait) in front of each instruction.
__device__ __forceinline__ void umma_bf16_nc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_bf16_acc_nc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc));
}
// Make all previously issued MMAs arrive on an mbarrier when they retire. Issued by ONE thread.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// The same arrival on the mbarrier at this offset in EVERY CTA of `mask`: a stage that was filled by multicast loads may only
// be overwritten once the MMAs of all the CTAs that received it have read it.
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i <-> lane base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (sm_100 "version 1"):
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4   [46,48) version = 1
//   [49,52) base offset               [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// The tcgen05 issuer is ONE thread: every ALU instruction between two MMAs is exposed latency (a 128x64x16 MMA keeps
// the tensor pipe busy for only 32 cycles).  The constant upper bits of a descriptor are therefore built once and the
// 14-bit start-address field is advanced by plain adds: +2 per 32-byte K step (K-major), +128 per 2048-byte K step
// (MN-major).  Shared-memory addresses are < 256 KB, so the field never carries into its neighbours.
__device__ __forceinline__ uint64_t umma_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// K-major operand whose rows are `row_bytes` (128 / 64 / 32) wide: layout type 2 / 4 / 6, 8-row atoms of 8*row_bytes.
__device__ __forceinline__ uint64_t umma_desc_hi_kmajor(uint32_t row_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>(1) << 16;                                           // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(((8u * row_bytes) >> 4) & 0x3FFF) << 32;             // SBO: next 8-row atom
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6)) << 61;
  return d;
}
// Same, with an explicit stride between 8-row atoms (rows of a tap inside a staged box with a halo along w).
__device__ __forceinline__ uint64_t umma_desc_hi_kmajor_sbo(uint32_t row_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6)) << 61;
  return d;
}
__device__ __forceinline__ uint64_t umma_desc_at(uint64_t hi, uint32_t saddr) {
  return hi | static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
}
// D[tmem] += A * B (accumulate unconditionally).
__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn_major,
                                                            uint32_t b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (a_mn_major & 1u) << 15;
  d |= (b_mn_major & 1u) << 16;
  d |= ((n >> 3) & 0x3Fu) << 17;
  d |= ((m >> 4) & 0x1Fu) << 24;
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// One 32-byte global store (STG.256, sm_100): a full sector per thread.  `p` must be 32-byte aligned.
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
// 16 fp32 accumulator columns -> 16 bf16 -> one 32-byte store.
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* dst, const uint32_t (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
  st_global_256(dst, w);
}
__device__ __forceinline__ void ld_global_256(const void* p, uint32_t (&w)[8]) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p)
               : "memory");
}
// 16 fp32 accumulator columns + 16 bf16 already in memory -> 16 bf16 -> one 32-byte store.
__device__ __forceinline__ void store_bf16x16_acc(__nv_bfloat16* dst, const uint32_t (&v)[16], const uint32_t (&o)[8]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    w[i] = pack_bf16x2(__uint_as_float(v[2 * i]) + __uint_as_float(o[i] << 16),
                       __uint_as_float(v[2 * i + 1]) + __uint_as_float(o[i] & 0xFFFF0000u));
  st_global_256(dst, w);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// Epilogue of one accumulator row (thread = TMEM lane) for a plain bf16 output: `ncols` (a multiple of 16) fp32 columns
// at `taddr` -> dst[0..ncols), optionally added to what dst holds.  Software-pipelined: the TMEM load (and, when
// accumulating, the 32-byte global load) of the next 16 columns is in flight while the current 16 are packed and written
// with ONE 32-byte store.  dst must be 32-byte aligned.  Every lane of the warp must call it (tcgen05.ld is .aligned);
// `valid` only gates the global accesses.
template <bool kAcc>
__device__ __forceinline__ void epilogue_row_bf16(uint32_t taddr, int ncols, __nv_bfloat16* dst, bool valid) {
  uint32_t va[16], vb[16], oa[8], ob[8];
  tmem_ld16(taddr, va);
  if (kAcc && valid) ld_global_256(dst, oa);
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    tmem_ld_wait();
    const bool more = c0 + 16 < ncols;
    if (more) {
      tmem_ld16(taddr + c0 + 16, vb);
      if (kAcc && valid) ld_global_256(dst + c0 + 16, ob);
    }
    if (valid) {
      if (kAcc) store_bf16x16_acc(dst + c0, va, oa);
      else store_bf16x16(dst + c0, va);
    }
    if (more) {
      tmem_ld_wait();
      if (c0 + 32 < ncols) {
        tmem_ld16(taddr + c0 + 32, va);
        if (kAcc && valid) ld_global_256(dst + c0 + 32, oa);
      }
      if (valid) {
        if (kAcc) store_bf16x16_acc(dst + c0 + 16, vb, ob);
        else store_bf16x16(dst + c0 + 16, vb);
      }
    }
  }
}

// The fp32 flavour (split-K partials of the weight-gradient kernels): 16 columns = two 32-byte stores.
__device__ __forceinline__ void epilogue_row_f32(uint32_t taddr, int ncols, float* dst, bool valid) {
  uint32_t va[16], vb[16];
  tmem_ld16(taddr, va);
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    tmem_ld_wait();
    const bool more = c0 + 16 < ncols;
    if (more) tmem_ld16(taddr + c0 + 16, vb);
    if (valid) {
      const uint32_t lo[8] = {va[0], va[1], va[2], va[3], va[4], va[5], va[6], va[7]};
      const uint32_t hi[8] = {va[8], va[9], va[10], va[11], va[12], va[13], va[14], va[15]};
      st_global_256(dst + c0, lo);
      st_global_256(dst + c0 + 8, hi);
    }
    if (more) {
      tmem_ld_wait();
      if (c0 + 32 < ncols) tmem_ld16(taddr + c0 + 32, va);
      if (valid) {
        const uint32_t lo[8] = {vb[0], vb[1], vb[2], vb[3], vb[4], vb[5], vb[6], vb[7]};
        const uint32_t hi[8] = {vb[8], vb[9], vb[10], vb[11], vb[12], vb[13], vb[14], vb[15]};
        st_global_256(dst + c0 + 16, lo);
        st_global_256(dst + c0 + 24, hi);
      }
    }
  }
}


// 2^x on the special-function unit, one instruction (results below 2^-126 flush to zero; relative error 2^-22).
// exp2f() without -use_fast_math wraps the same MUFU.EX2 in a denormal-range check (FSETP + 2 FMUL + predicate logic).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------- warpgroup register re-allocation
// All four warps of a warpgroup must execute these (setmaxnreg is .sync.aligned over the warpgroup).
template <uint32_t kRegs>
__device__ __forceinline__ void warpgroup_reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <uint32_t kRegs>
__device__ __forceinline__ void warpgroup_reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

// ---------------------------------------------------------------- operand prologue: BatchNorm affine + ReLU in place
// The consumer of a conv+BN unit reads the producer's RAW output and applies y = max(x*scale[c] + shift[c], 0) to the
// TMA-staged operand box in shared memory (then bf16 again: the very value cstp_bn_apply would have written to HBM and
// the consumer would have loaded), so the normalised activation never exists in HBM (models/pace/r21d_byol.py:94-97:
// conv -> bn -> relu -> conv).  Convolution zero padding must stay exactly 0 behind the affine map: the tensor map of a
// transformed operand fills out-of-range elements with NaN (CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) and
// fmaxf(NaN, 0) == 0 -- no coordinate predicates.  Padded channels carry scale == shift == 0.
//
// Swizzled boxes (128 / 64 / 32-byte rows): the 16-byte unit at shared address a holds the channels
// 8 * (((a >> 4) ^ (a >> 7)) & (units_per_row - 1)) .. + 8 of its row's chunk.  A thread that walks units with a stride
// that is a multiple of 1024 bytes therefore keeps ONE channel vector: its 16 coefficients stay in registers.
struct XformCoef {
  float s[8], b[8];
};

// Coefficients of the 8 channels starting at `ch` (a multiple of 8) of a [Cp] fp32 row; zeros beyond Cp.
__device__ __forceinline__ void xform_load(XformCoef& k, const float* __restrict__ scale, const float* __restrict__ shift,
                                           int ch, int Cp) {
  if (ch < Cp) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + ch)), s1 = __ldg(reinterpret_cast<const float4*>(scale + ch) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + ch)), b1 = __ldg(reinterpret_cast<const float4*>(shift + ch) + 1);
    k.s[0] = s0.x; k.s[1] = s0.y; k.s[2] = s0.z; k.s[3] = s0.w; k.s[4] = s1.x; k.s[5] = s1.y; k.s[6] = s1.z; k.s[7] = s1.w;
    k.b[0] = b0.x; k.b[1] = b0.y; k.b[2] = b0.z; k.b[3] = b0.w; k.b[4] = b1.x; k.b[5] = b1.y; k.b[6] = b1.z; k.b[7] = b1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) k.s[i] = k.b[i] = 0.f;
  }
}
// First channel (within its 64-channel chunk) of the 16-byte unit at shared address `a`; units_mask = units per row - 1.
__device__ __forceinline__ int xform_unit_channel(uint32_t a, uint32_t units_mask) {
  return static_cast<int>((((a >> 4) ^ (a >> 7)) & units_mask) * 8u);
}
__device__ __forceinline__ uint32_t xform_pair(uint32_t v, float s0, float b0, float s1, float b1) {
  return pack_bf16x2(fmaxf(fmaf(bf16_lo(v), s0, b0), 0.f), fmaxf(fmaf(bf16_hi(v), s1, b1), 0.f));
}
__device__ __forceinline__ uint4 xform_unit(const uint4& v, const XformCoef& k) {
  uint4 o;
  o.x = xform_pair(v.x, k.s[0], k.b[0], k.s[1], k.b[1]);
  o.y = xform_pair(v.y, k.s[2], k.b[2], k.s[3], k.b[3]);
  o.z = xform_pair(v.z, k.s[4], k.b[4], k.s[5], k.b[5]);
  o.w = xform_pair(v.w, k.s[6], k.b[6], k.s[7], k.b[7]);
  return o;
}
// The same map when nothing downstream reads out-of-range elements (no NaN -> 0 rescue needed): ReLU and the bf16 rounding
// are ONE conversion (cvt.rn.relu.bf16x2.f32: negative results clamp to +0) and the low half is widened on the FMA pipe
// (IMAD) -- 2 ALU-pipe + 3 FMA-pipe instructions per pair instead of 5 + 2; the ALU pipe (one warp instruction per two
// cycles and scheduler) is what bounds the transform warps.  Same bits as xform_pair for finite inputs.
__device__ __forceinline__ uint32_t xform_pair_relu(uint32_t v, float s0, float b0, float s1, float b1) {
  const float lo = __uint_as_float(v * 65536u), hi = __uint_as_float(v & 0xffff0000u);
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(fmaf(hi, s1, b1)), "f"(fmaf(lo, s0, b0)));
  return r;
}
__device__ __forceinline__ uint4 xform_unit_relu(const uint4& v, const XformCoef& k) {
  uint4 o;
  o.x = xform_pair_relu(v.x, k.s[0], k.b[0], k.s[1], k.b[1]);
  o.y = xform_pair_relu(v.y, k.s[2], k.b[2], k.s[3], k.b[3]);
  o.z = xform_pair_relu(v.z, k.s[4], k.b[4], k.s[5], k.b[5]);
  o.w = xform_pair_relu(v.w, k.s[6], k.b[6], k.s[7], k.b[7]);
  return o;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Transforms the units [first, end) of the box at shared address `base` that belong to this thread: first, first + kStep,
// ...  (kStep = participating threads; kStep * 16 must be a multiple of 1024).  Every load of a batch of (up to) four units
// is issued before the first result is needed -- the loop is latency bound, not throughput bound (ncu: the transform
// warps of the first version sat on the LDS scoreboard while the issue slots were 35 % busy).
// kRelu: the box has no out-of-range element (interior tile, full channel chunk): xform_unit_relu applies.
template <int kStep, bool kRelu = false>
__device__ __forceinline__ void xform_span(uint32_t base, uint32_t first, uint32_t end, const XformCoef& k) {
  static_assert((kStep * 16) % 1024 == 0, "thread stride must keep the swizzle phase");
  constexpr uint32_t kB = kStep * 16u;
  auto f = [&k](const uint4& v) { return kRelu ? xform_unit_relu(v, k) : xform_unit(v, k); };
  uint32_t u = first;
  for (; u + 3 * kStep < end; u += 4 * kStep) {
    const uint32_t a = base + u * 16u;
    const uint4 v0 = lds128(a), v1 = lds128(a + kB), v2 = lds128(a + 2u * kB), v3 = lds128(a + 3u * kB);
    sts128(a, f(v0));
    sts128(a + kB, f(v1));
    sts128(a + 2u * kB, f(v2));
    sts128(a + 3u * kB, f(v3));
  }
  if (u < end) {                    // one to three units left: same shape, predicated
    const uint32_t a = base + u * 16u;
    const bool p1 = u + kStep < end, p2 = u + 2 * kStep < end;
    uint4 v0 = lds128(a), v1 = v0, v2 = v0;
    if (p1) v1 = lds128(a + kB);
    if (p2) v2 = lds128(a + 2u * kB);
    sts128(a, f(v0));
    if (p1) sts128(a + kB, f(v1));
    if (p2) sts128(a + 2u * kB, f(v2));
  }
}

// Coefficient table in shared memory, filled once per CTA by the transform warps: for every statistics group, 64-channel
// chunk c and 8-channel vector j one record {scale[8], shift[8]} (zeros beyond Cp), so that a thread fetches the
// coefficients of a staged box with four LDS.128 instead of four L2 round trips per stage.  Records are 80 bytes apart:
// the eight records of a chunk (the eight vectors the lanes of a warp ask for) then start in eight different 16-byte
// bank groups and every LDS.128 of the fetch is ONE conflict-free broadcast wavefront (with 64-byte records it was a
// 4-way conflict, as much shared-memory time as the data of the box itself).
constexpr uint32_t kXformRec = 80;                                  // bytes per record
__host__ __device__ __forceinline__ uint32_t xform_table_bytes(int groups, int Cp) {
  return static_cast<uint32_t>(groups * ((Cp + 63) / 64) * 8) * kXformRec;
}
template <int kThreads>
__device__ __forceinline__ void xform_table_fill(float* tbl, const float* __restrict__ scale, const float* __restrict__ shift,
                                                 int groups, int Cp, uint32_t tid) {
  const int chunks = (Cp + 63) / 64;
  const int total = groups * chunks * 64;
  for (int i = static_cast<int>(tid); i < total; i += kThreads) {
    const int g = i / (chunks * 64), ch = i % (chunks * 64);
    const bool in = ch < Cp;
    float* rec = tbl + (static_cast<size_t>(g * chunks + ch / 64) * 8 + (ch % 64) / 8) * (kXformRec / 4) + (ch % 8);
    rec[0] = in ? __ldg(scale + g * Cp + ch) : 0.f;
    rec[8] = in ? __ldg(shift + g * Cp + ch) : 0.f;
  }
}
// Record of (group g, chunk c, channel offset cj = 8 * j inside the chunk); tbl_addr is the table's shared address.
__device__ __forceinline__ void xform_load_smem(XformCoef& k, uint32_t tbl_addr, int chunks, int g, int c, int cj) {
  const uint32_t a = tbl_addr + static_cast<uint32_t>((g * chunks + c) * 8 + (cj >> 3)) * kXformRec;
  const uint4 s0 = lds128(a), s1 = lds128(a + 16u), b0 = lds128(a + 32u), b1 = lds128(a + 48u);
  k.s[0] = __uint_as_float(s0.x); k.s[1] = __uint_as_float(s0.y); k.s[2] = __uint_as_float(s0.z); k.s[3] = __uint_as_float(s0.w);
  k.s[4] = __uint_as_float(s1.x); k.s[5] = __uint_as_float(s1.y); k.s[6] = __uint_as_float(s1.z); k.s[7] = __uint_as_float(s1.w);
  k.b[0] = __uint_as_float(b0.x); k.b[1] = __uint_as_float(b0.y); k.b[2] = __uint_as_float(b0.z); k.b[3] = __uint_as_float(b0.w);
  k.b[4] = __uint_as_float(b1.x); k.b[5] = __uint_as_float(b1.y); k.b[6] = __uint_as_float(b1.z); k.b[7] = __uint_as_float(b1.w);
}
// Named barrier among the transform warps only (id 2; the statistics epilogue uses id 1).
template <int kThreads>
__device__ __forceinline__ void xform_bar_sync() {
  asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");
}
// Smallest u >= lo with u == tid (mod kStep).
template <int kStep>
__device__ __forceinline__ uint32_t xform_first(uint32_t lo, uint32_t tid) {
  return lo + (tid + kStep - lo % kStep) % kStep;
}

}  // namespace cstp
