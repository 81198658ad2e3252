// World-synchronised BatchNorm statistics over NVLink peer memory: the [groups][2][Cp] fp32 row every BatchNorm call
// exchanges (SyncBN forward sums / backward sums) is pushed straight into every peer's receive buffer and summed locally,
// in ONE single-CTA kernel -- in place of an NCCL all-reduce whose ~25-30 us latency, paid ~140 times per step, was 4 ms of
// an 18.9 ms step at 8 x batch 16 (profiles/README.md).  The reference has no counterpart (its --sync_bn group holds one
// rank, models/model.py:95-96); the north star asks for SyncBN over the data-parallel group.
//
// Buffer of rank p (peer-mapped on every rank, torch symmetric memory):
//   float    data[slots][world][row_max]     row of rank q for call `seq` lands in data[seq % slots][q]
//   uint32_t flag[slots][world]              flag[seq % slots][q] == seq  <=>  that row is complete
// Call `seq` (1, 2, ... identical on every rank: all ranks run the same BatchNorm call sequence per stream):
//   1. every thread stores its part of the local row into data[slot][rank] of EVERY rank, then fence.sys
//   2. after a CTA barrier, thread q publishes flag[slot][rank] = seq in rank q's buffer (st.release.sys)
//   3. thread q spins (ld.acquire.sys, bounded) until flag[slot][q] == seq in the LOCAL buffer
//   4. the rows are summed in rank order (the same order on every rank: bit-identical statistics everywhere), read with
//      ld.cg so that no stale L1 line of the slot's previous use is seen.
// Slot reuse is safe with slots >= 2: a rank can only be `slots` calls ahead of a peer after that peer has published the
// calls in between, i.e. after it finished reading this slot.
#include "common.h"

namespace cstp {

constexpr int kSyncThreads = 1024;
constexpr long long kSpinCycles = 20LL * 1000 * 1000 * 1000;   // ~10 s of SM clock: report instead of hanging the GPU

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kSyncThreads) bn_sync_exchange_kernel(const float* __restrict__ row, int n,
                                                                       const uint64_t* __restrict__ peers, int world, int rank,
                                                                       int slots, int row_max, uint32_t seq,
                                                                       uint32_t* __restrict__ seq_dev,
                                                                       float* __restrict__ out, int* __restrict__ err) {
  if (seq_dev != nullptr) {      // call counter in device memory (one per buffer set): replayable from a CUDA graph
    __shared__ uint32_t s_seq;
    if (threadIdx.x == 0) {
      s_seq = *seq_dev + 1u;
      *seq_dev = s_seq;
    }
    __syncthreads();
    seq = s_seq;
  }
  const int slot = static_cast<int>(seq % static_cast<uint32_t>(slots));
  const size_t flag_off = static_cast<size_t>(slots) * world * row_max * sizeof(float);
  const int n4 = n >> 2;
  const float4* row4 = reinterpret_cast<const float4*>(row);
  for (int p = 0; p < world; ++p) {
    const int peer = (rank + p) % world;                     // start with the own buffer, spread the NVLink targets
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(peers[peer]) +
                                            (static_cast<size_t>(slot) * world + rank) * row_max);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = row4[i];
  }
  __threadfence_system();
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < world) {
    const int q = threadIdx.x;
    uint32_t* theirs = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peers[q]) + flag_off) + slot * world + rank;
    st_release_sys(theirs, seq);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(peers[rank]) + flag_off) +
                           slot * world + q;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) != seq) {
      if (clock64() - t0 > kSpinCycles) {
        atomicExch(err, 1 + q);
        break;
      }
    }
  }
  __syncthreads();
  const float4* mine4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(peers[rank]) +
                                                        static_cast<size_t>(slot) * world * row_max);
  const int stride4 = row_max >> 2;
  float4* out4 = reinterpret_cast<float4*>(out);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 s = __ldcg(mine4 + i);
    for (int q = 1; q < world; ++q) {
      const float4 v = __ldcg(mine4 + static_cast<size_t>(q) * stride4 + i);
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
    out4[i] = s;
  }
}

}  // namespace cstp

using namespace cstp;

extern "C" long long cstp_bn_sync_buffer_bytes(int world, int slots, int row_max) {
  if (world < 1 || slots < 2 || row_max < 4) return CSTP_EINVAL;
  return static_cast<long long>(slots) * world * row_max * 4 + static_cast<long long>(slots) * world * 4;
}

extern "C" int cstp_bn_sync_exchange(const float* row, int n, const uint64_t* peer_buffers, int world, int rank, int slots,
                                     int row_max, uint32_t seq, uint32_t* seq_dev, float* out, int* err_flag, void* stream) {
  CSTP_REQUIRE(row != nullptr && out != nullptr && peer_buffers != nullptr && err_flag != nullptr);
  CSTP_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && slots >= 2 && ((seq != 0) != (seq_dev != nullptr)));
  CSTP_REQUIRE(n > 0 && n % 4 == 0 && n <= row_max && row_max % 4 == 0);
  CSTP_REQUIRE(reinterpret_cast<uintptr_t>(row) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0);
  bn_sync_exchange_kernel<<<1, kSyncThreads, 0, static_cast<cudaStream_t>(stream)>>>(row, n, peer_buffers, world, rank, slots,
                                                                                    row_max, seq, seq_dev, out, err_flag);
  CSTP_LAUNCHED();
  return CSTP_OK;
}
