// NT-Xent on the tensor cores (loss/NTXent.py:46-62 in closed form).
//
// The SIMT version (loss_optim.cu) walks the rows x rows similarity matrix with one CTA per 64 rows and fp32 FMAs: time
// grows linearly in rows at ~7 TFLOP/s (6.7 ms at rows = 8192).  Here the matrix is produced 128 x 128 tiles at a time by
// tcgen05.mma from bf16 copies of the normalised embeddings (S = Zn Zn^T, fp32 accumulate in TMEM, double buffered):
//   forward   the epilogue warps keep a running (max, sum-exp) per row and pick out the positive logit; partials per
//             (row, column split) are merged by a small kernel -> lse_i, loss;
//   backward  the epilogue turns each S tile into W_ij = (P_ij + P_ji - 2[j = pos(i)]) / (tau * rows) with
//             P_ij = exp(s_ij - lse_i) and stores it as bf16; dZn = W Zn is then one call of the implicit-GEMM kernel
//             of conv_gemm.cu (a 1-tap "linear" plan), followed by the cosine-normalisation Jacobian.
// The index map is the reference's: pos(i) = (i + N) mod 2N over cat(zjs, zis), masked set exactly {i, pos(i)}.
#include "common.h"
#include "ptx.cuh"

namespace cstp {

constexpr int kNtThreads = 256;
constexpr int kNtThreads2 = 384;                  // forward / fused backward: eight epilogue warps (two per TMEM lane quarter,
                                                 // each taking half of the 128 columns): one warp per scheduler cannot keep
                                                 // its MUFU unit busy through the TMEM-load / exp2 / store latency chain
constexpr int kNtMaxStages = 6;
constexpr int kNtSmemLimit = 232448;
constexpr uint32_t kNtBoxBytes = 128 * 64 * 2;   // 128 rows x 64 bf16 channels
constexpr uint32_t kNtWPitch = 272;              // bytes per staged W row (256 + 16)
constexpr uint32_t kNtWSlab = 32 * kNtWPitch;    // one warp's 32 rows (8704 B, a multiple of 16)

struct NtxKParams {
  CUtensorMap zmap;          // bf16 [rows_pad][d], box 64 x 128
  int rows, rows_pad, half, nkc, tiles, tiles_per_split;
  int stages;
  uint32_t idesc;
  float scale_log2;          // log2(e) / tau : logits in base-2 units
  float wscale;              // 1 / (tau * rows)
  // forward outputs: partials [nsplit][rows_pad] of running max (base-2 units) and sum, and the positive logit
  float* pm;
  float* pl;
  float* ppos;
  // backward inputs / outputs
  const float* lse2;         // [rows_pad] log-sum-exp in base-2 units
  __nv_bfloat16* W;          // [rows_pad][rows_pad]
};

template <bool kBackward>
__global__ void __launch_bounds__(kNtThreads2, 1) ntxent_s_kernel(const __grid_constant__ NtxKParams p) {
  constexpr int kEpiWarps = kBackward ? 4 : 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t tile_bytes = static_cast<uint32_t>(p.nkc) * kNtBoxBytes;      // one 128 x d operand tile
  uint8_t* stage0 = smem + tile_bytes;                                        // [0, tile_bytes): the resident row tile
  // backward only: per-epilogue-warp staging of its 32 x 128 bf16 slab of W (row pitch 272 B against bank conflicts)
  uint8_t* wstage = stage0 + static_cast<size_t>(p.stages) * tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wstage + (kBackward ? 4 * kNtWSlab : 0));
  uint64_t* full = bars;
  uint64_t* empty = bars + kNtMaxStages;
  uint64_t* tfull = bars + 2 * kNtMaxStages;
  uint64_t* tempty = bars + 2 * kNtMaxStages + 2;
  uint64_t* afull = bars + 2 * kNtMaxStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNtMaxStages + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rt = blockIdx.x, split = blockIdx.y;
  const int jt0 = split * p.tiles_per_split;
  const int jt1 = min(p.tiles, jt0 + p.tiles_per_split);

  if (warp == 0 && lane == 0) tma_prefetch_desc(&p.zmap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps);
    }
    mbar_init(afull, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(afull, tile_bytes);
      for (int kc = 0; kc < p.nkc; ++kc) tma_load_2d(smem + kc * kNtBoxBytes, &p.zmap, afull, kc * 64, rt * 128);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int jt = jt0; jt < jt1; ++jt) {
      mbar_wait(&empty[stage], phase ^ 1u);
      if (leader) {
        uint8_t* st = stage0 + static_cast<size_t>(stage) * tile_bytes;
        mbar_expect_tx(&full[stage], tile_bytes);
        for (int kc = 0; kc < p.nkc; ++kc) tma_load_2d(st + kc * kNtBoxBytes, &p.zmap, &full[stage], kc * 64, jt * 128);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: S tile = Z_i (128 x d) . Z_j^T
    const bool leader = elect_one();
    mbar_wait(afull, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(smem);
    const uint32_t s_addr0 = smem_u32(stage0);
    const uint64_t dhi = umma_desc_hi(16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int jt = jt0; jt < jt1; ++jt, ++it) {
      const int as = it & 1;
      mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1u);
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (leader) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * 128);
        const uint32_t b_addr = s_addr0 + static_cast<uint32_t>(stage) * tile_bytes;
        for (int kc = 0; kc < p.nkc; ++kc) {
          const uint64_t da = umma_desc_at(dhi, a_addr + kc * kNtBoxBytes);
          const uint64_t db = umma_desc_at(dhi, b_addr + kc * kNtBoxBytes);
          umma_bf16(d_tmem, da, db, p.idesc, kc != 0 ? 1u : 0u);
          umma_bf16_acc(d_tmem, da + 2, db + 2, p.idesc);
          umma_bf16_acc(d_tmem, da + 4, db + 4, p.idesc);
          umma_bf16_acc(d_tmem, da + 6, db + 6, p.idesc);
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[as]);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ------------------------------------------------------------ epilogue: lane owns row i of the row tile
    const int q = (warp - 4) & 3;
    const int g = (warp - 4) >> 2;                   // column half of the tile (forward only; 0 otherwise)
    const int c_lo = kBackward ? 0 : g * 64, c_hi = kBackward ? 128 : g * 64 + 64;
    const int i = rt * 128 + q * 32 + lane;
    const int pos = i < p.half ? i + p.half : i - p.half;
    const bool row_ok = i < p.rows;
    float m = -INFINITY, l = 0.f, spos = 0.f;
    float lse_i = 0.f;
    const float scale_f = p.scale_log2;
    if (kBackward) lse_i = row_ok ? p.lse2[i] : 0.f;
    int it = 0;
    for (int jt = jt0; jt < jt1; ++jt, ++it) {
      const int as = it & 1;
      mbar_wait(&tfull[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * 128);
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        const int j0 = jt * 128 + c0;
        if (!kBackward) {
          float s[32];
          float tm = -INFINITY;
          // (plain chunk: no diagonal, no positive, inside the matrix -- for this warp's 32 rows)
          const bool special = (i >= j0 && i < j0 + 32) || (pos >= j0 && pos < j0 + 32) || j0 + 32 > p.rows;
          if (!__any_sync(0xffffffffu, special)) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              s[c] = __uint_as_float(v[c]) * scale_f;
              tm = fmaxf(tm, s[c]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int j = j0 + c;
              float x = __uint_as_float(v[c]) * scale_f;
              if (j == pos) spos = x;
              x = (j == i || j >= p.rows) ? -INFINITY : x;
              s[c] = x;
              tm = fmaxf(tm, x);
            }
          }
          const float mn = fmaxf(m, tm);
          if (mn > -INFINITY) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) acc += ex2_approx(s[c] - mn);
            l = l * ex2_approx(m - mn) + acc;
            m = mn;
          }
        } else {
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float w2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = j0 + c + u;
              const float x = __uint_as_float(v[c + u]) * p.scale_log2;
              float wv = 0.f;
              if (row_ok && j < p.rows && j != i) {
                wv = exp2f(x - lse_i) + exp2f(x - __ldg(p.lse2 + j));
                if (j == pos) wv -= 2.f;
              }
              w2[u] = wv * p.wscale;
            }
            pk[c >> 1] = pack_bf16x2(w2[0], w2[1]);
          }
          // stage this lane's 64 bytes (32 columns) in the warp's slab; written out coalesced once the tile is complete
          uint4* srow = reinterpret_cast<uint4*>(wstage + static_cast<size_t>(q) * kNtWSlab + lane * kNtWPitch + c0 * 2);
#pragma unroll
          for (int u = 0; u < 4; ++u) srow[u] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      }
      if (kBackward) {
        // Each lane used to store its own row of W straight to global memory: 32 different 128-byte lines per store
        // instruction, and the LSU serialised the whole backward (0.52 ms at rows = 8192).  Now the warp writes its
        // 32 x 256-byte slab with 16 fully coalesced instructions (two rows of 256 contiguous bytes each).
        __syncwarp();
        const uint8_t* slab = wstage + static_cast<size_t>(q) * kNtWSlab;
        const long long row0 = static_cast<long long>(rt) * 128 + q * 32;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
          const int r = 2 * k + (lane >> 4), piece = lane & 15;
          const uint4 val = *reinterpret_cast<const uint4*>(slab + r * kNtWPitch + piece * 16);
          *reinterpret_cast<uint4*>(p.W + (row0 + r) * p.rows_pad + jt * 128 + piece * 8) = val;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
    if (!kBackward) {
      // one (max, sum) partial per (column split, column half): the merge kernel sees 2 x nsplit of them
      const long long o = static_cast<long long>(split * 2 + g) * p.rows_pad + i;
      p.pm[o] = m;
      p.pl[o] = l;
      if (row_ok && pos >= jt0 * 128 && pos < jt1 * 128 && ((pos & 127) >> 6) == g) p.ppos[i] = spos;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// ---------------------------------------------------------------------------------------------- fused backward
// dZn_i = sum_j W_ij Zn_j without ever writing W: for every 128 x 128 tile the S tile is recomputed by tcgen05 into TMEM,
// the epilogue warps turn it into W (bf16) and store it as a K-major, 128B-swizzled A operand in shared memory, and a second
// tcgen05.mma accumulates W_ij . Zn_j into a 128 x d accumulator that stays in TMEM for the whole row tile.  The Zn_j tile
// staged for the first MMA (K-major: rows j, channels contiguous) IS the B operand of the second one read MN-major
// (N = channels contiguous, K = rows j) -- one TMA load serves both.  The materialised-W version wrote and re-read
// rows^2 bf16 (134 MB at rows = 8192) and ran a separate GEMM; this one is bound by the 2 x 128 x 128 exp2 per tile (MUFU).
// d <= 128 (TMEM: 2 x 128 columns of S + d columns of dZ; shared memory: Z_i + stages x Z_j + 2 x 32 KB of W).
struct NtxBwdParams {
  CUtensorMap zmap;          // bf16 [rows_pad][d], box 64 x 128
  int rows, rows_pad, half, nkc, tiles, tiles_per_split, stages, d;
  uint32_t idesc_s, idesc_d;
  float scale_log2, wscale;
  const float* lse2;         // [rows_pad] log-sum-exp in base-2 units
  float* dzn_part;           // [nsplit][rows_pad][d] partial dZn per column split
};

constexpr uint32_t kNtWTile = 128 * 128 * 2;      // one W tile: two 128 x 64 K-major boxes

__global__ void __launch_bounds__(kNtThreads2, 1) ntxent_bwd_fused_kernel(const __grid_constant__ NtxBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t tile_bytes = static_cast<uint32_t>(p.nkc) * kNtBoxBytes;
  uint8_t* stage0 = smem + tile_bytes;
  uint8_t* wbuf = stage0 + static_cast<size_t>(p.stages) * tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbuf + 2 * kNtWTile);
  uint64_t* full = bars;
  uint64_t* empty = bars + kNtMaxStages;
  uint64_t* tfull = bars + 2 * kNtMaxStages;          // S accumulator ready              (issuer -> epilogue)
  uint64_t* tempty = tfull + 2;                       // S accumulator drained            (epilogue -> issuer)
  uint64_t* wfull = tempty + 2;                       // W tile written to shared memory  (epilogue -> issuer)
  uint64_t* wempty = wfull + 2;                       // W tile consumed by the 2nd MMA   (issuer -> epilogue)
  uint64_t* afull = wempty + 2;
  uint64_t* dfull = afull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rt = blockIdx.x, split = blockIdx.y;
  const int jt0 = split * p.tiles_per_split;
  const int jt1 = min(p.tiles, jt0 + p.tiles_per_split);

  if (warp == 0 && lane == 0) tma_prefetch_desc(&p.zmap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);
      mbar_init(&wfull[a], 8);
      mbar_init(&wempty[a], 1);
    }
    mbar_init(afull, 1);
    mbar_init(dfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_dz = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(afull, tile_bytes);
      for (int kc = 0; kc < p.nkc; ++kc) tma_load_2d(smem + kc * kNtBoxBytes, &p.zmap, afull, kc * 64, rt * 128);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int jt = jt0; jt < jt1; ++jt) {
      mbar_wait(&empty[stage], phase ^ 1u);
      if (leader) {
        uint8_t* st = stage0 + static_cast<size_t>(stage) * tile_bytes;
        mbar_expect_tx(&full[stage], tile_bytes);
        for (int kc = 0; kc < p.nkc; ++kc) tma_load_2d(st + kc * kNtBoxBytes, &p.zmap, &full[stage], kc * 64, jt * 128);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: S(it) first, then W(it - 1) . Zn(it - 1)
    const bool leader = elect_one();
    mbar_wait(afull, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(smem);
    const uint32_t s_addr0 = smem_u32(stage0);
    const uint32_t w_addr0 = smem_u32(wbuf);
    const uint64_t dhi_k = umma_desc_hi(16, 1024);                 // K-major operands (Z_i, Z_j for S; W for dZ)
    const uint64_t dhi_mn = umma_desc_hi(kNtBoxBytes, 1024);       // Z_j read MN-major: LBO = next 64-channel box
    const int nkc = p.nkc, stages = p.stages;
    uint32_t idesc_s, idesc_d;
    asm volatile("mov.u32 %0, %1;" : "=r"(idesc_s) : "r"(p.idesc_s));
    asm volatile("mov.u32 %0, %1;" : "=r"(idesc_d) : "r"(p.idesc_d));
    int stage = 0, stage2 = 0;
    uint32_t phase = 0;
    auto second = [&](int t) {           // dZ += W(t) . Zn(t); frees the W buffer and the Zn stage of tile t
      const int bs = t & 1;
      mbar_wait(&wfull[bs], (t >> 1) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t da0 = umma_desc_at(dhi_k, w_addr0 + static_cast<uint32_t>(bs) * kNtWTile);
        const uint64_t db0 = umma_desc_at(dhi_mn, s_addr0 + static_cast<uint32_t>(stage2) * tile_bytes);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // A: box kb (16 KB apart), 32 bytes per K step; B: 16 rows j (2048 bytes) per K step
            const uint64_t da = da0 + static_cast<uint64_t>(kb) * (kNtBoxBytes >> 4) + static_cast<uint64_t>(ks) * 2;
            const uint64_t db = db0 + static_cast<uint64_t>(kb * 4 + ks) * 128;
            umma_bf16_nc(tmem_dz, da, db, idesc_d, (t != 0 || kb != 0 || ks != 0) ? 1u : 0u);
          }
        }
        umma_commit(&wempty[bs]);
        umma_commit(&empty[stage2]);
      }
      __syncwarp();
      if (++stage2 == stages) stage2 = 0;
    };
    int it = 0;
    for (int jt = jt0; jt < jt1; ++jt, ++it) {
      const int as = it & 1;
      mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1u);
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (leader) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * 128);
        const uint32_t b_addr = s_addr0 + static_cast<uint32_t>(stage) * tile_bytes;
        for (int kc = 0; kc < nkc; ++kc) {
          const uint64_t da = umma_desc_at(dhi_k, a_addr + kc * kNtBoxBytes);
          const uint64_t db = umma_desc_at(dhi_k, b_addr + kc * kNtBoxBytes);
          umma_bf16_nc(d_tmem, da, db, idesc_s, kc != 0 ? 1u : 0u);
          umma_bf16_acc_nc(d_tmem, da + 2, db + 2, idesc_s);
          umma_bf16_acc_nc(d_tmem, da + 4, db + 4, idesc_s);
          umma_bf16_acc_nc(d_tmem, da + 6, db + 6, idesc_s);
        }
        umma_commit(&tfull[as]);
      }
      __syncwarp();
      if (++stage == stages) {
        stage = 0;
        phase ^= 1u;
      }
      if (it > 0) second(it - 1);
    }
    if (it > 0) second(it - 1);
    if (leader) umma_commit(dfull);
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: lane owns row i; S tile -> W tile (smem)
    const int q = (warp - 4) & 3;
    const int g = (warp - 4) >> 2;                    // column half of the tile this warp converts
    const int r = q * 32 + lane;                      // row inside the tile
    const int i = rt * 128 + r;
    const int pos = i < p.half ? i + p.half : i - p.half;
    const bool row_ok = i < p.rows;
    const float lse_i = row_ok ? p.lse2[i] : 0.f;
    const uint32_t w_addr0 = smem_u32(wbuf);
    const float* __restrict__ lse2 = p.lse2;
    const float scale = p.scale_log2, wscale = p.wscale;
    int it = 0;
    for (int jt = jt0; jt < jt1; ++jt, ++it) {
      const int as = it & 1;
      mbar_wait(&tfull[as], (it >> 1) & 1);
      mbar_wait(&wempty[as], ((it >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * 128);
      const uint32_t wrow = w_addr0 + static_cast<uint32_t>(as) * kNtWTile + static_cast<uint32_t>(r) * 128u;
      // A tile is "plain" for this warp when none of its rows has its diagonal or its positive in it and it lies inside
      // the matrix: then w = (2^(x - lse_i) + 2^(x - lse_j)) / (tau rows) for all 32 x 128 elements -- two FFMA, two MUFU,
      // an add, a multiply and half a pack per element, the lse_j row fetched as float4 (the generic form below spends
      // ~40 instructions per element on index tests, per-element loads and exp2f's denormal handling: 18k cycles per tile).
      const int jbase = jt * 128;
      const bool special = jt == rt || (pos >= jbase && pos < jbase + 128) || jbase + 128 > p.rows || !row_ok;
      const bool plain = !__any_sync(0xffffffffu, special);
      for (int c0 = g * 64; c0 < g * 64 + 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        const int j0 = jbase + c0;
        uint32_t pk[16];
        if (plain) {
          const float4* lj = reinterpret_cast<const float4*>(lse2 + j0);
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 l4 = __ldg(lj + (c >> 2));
            const float x0 = __uint_as_float(v[c]) * scale, x1 = __uint_as_float(v[c + 1]) * scale;
            const float x2 = __uint_as_float(v[c + 2]) * scale, x3 = __uint_as_float(v[c + 3]) * scale;
            const float w0 = (ex2_approx(x0 - lse_i) + ex2_approx(x0 - l4.x)) * wscale;
            const float w1 = (ex2_approx(x1 - lse_i) + ex2_approx(x1 - l4.y)) * wscale;
            const float w2_ = (ex2_approx(x2 - lse_i) + ex2_approx(x2 - l4.z)) * wscale;
            const float w3 = (ex2_approx(x3 - lse_i) + ex2_approx(x3 - l4.w)) * wscale;
            pk[c >> 1] = pack_bf16x2(w0, w1);
            pk[(c >> 1) + 1] = pack_bf16x2(w2_, w3);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float w2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = j0 + c + u;
              const float x = __uint_as_float(v[c + u]) * scale;
              float wv = 0.f;
              if (row_ok && j < p.rows && j != i) {
                wv = ex2_approx(x - lse_i) + ex2_approx(x - __ldg(lse2 + j));
                if (j == pos) wv -= 2.f;
              }
              w2[u] = wv * wscale;
            }
            pk[c >> 1] = pack_bf16x2(w2[0], w2[1]);
          }
        }
        // columns c0 .. c0 + 31 = four 16-byte units of box c0 / 64; unit u of row r sits at ((u ^ (r & 7)) * 16)
        const uint32_t box = wrow + static_cast<uint32_t>(c0 >> 6) * kNtBoxBytes;
        const uint32_t u0 = static_cast<uint32_t>((c0 & 63) >> 3);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          sts128(box + (((u0 + u) ^ static_cast<uint32_t>(r & 7)) << 4),
                 make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]));
      }
      tc_fence_before();
      fence_proxy_async();                 // the W tile is read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tempty[as]);
        mbar_arrive(&wfull[as]);
      }
    }
    // ---- the row tile's dZn partial: TMEM -> fp32 rows
    mbar_wait(dfull, 0);
    tc_fence_after();
    const int dh = p.d >> 1;                         // each column-half warp writes half of the d columns
    float* dst = p.dzn_part + (static_cast<long long>(split) * p.rows_pad + i) * p.d + g * dh;
    epilogue_row_f32(tmem_dz + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * dh), dh, dst, jt1 > jt0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dzn[i][c] = sum over the column splits, in split order (deterministic).  n is a multiple of 4 (rows_pad x d).
__global__ void ntxent_sum_splits_kernel(const float4* __restrict__ part, int nsplit, long long n4, float4* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 acc = part[i];
    for (int s = 1; s < nsplit; ++s) {
      const float4 v = part[static_cast<long long>(s) * n4 + i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[i] = acc;
  }
}

// zn fp32 [rows][d] -> bf16 [rows_pad][d] (zero rows beyond `rows`) and its transpose bf16 [d][rows_pad].
__global__ void ntxent_cast_kernel(const float* __restrict__ zn, int rows, int rows_pad, int d,
                                   __nv_bfloat16* __restrict__ zb, __nv_bfloat16* __restrict__ zt) {
  const long long total = static_cast<long long>(rows_pad) * d;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / d), c = static_cast<int>(i % d);
    const __nv_bfloat16 v = __float2bfloat16_rn(r < rows ? zn[i] : 0.f);
    zb[i] = v;
    if (zt != nullptr) zt[static_cast<long long>(c) * rows_pad + r] = v;      // (only the materialised-W backward reads it)
  }
}

// Merges the per-split (max, sum) partials in fixed order: lse (base-2 units, for the backward), row losses (natural).
__global__ void ntxent_merge_kernel(const float* __restrict__ pm, const float* __restrict__ pl,
                                    const float* __restrict__ ppos, int nsplit, int rows, int rows_pad,
                                    float* __restrict__ lse2, float* __restrict__ row_loss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_pad) return;
  if (i >= rows) {
    lse2[i] = 0.f;
    return;
  }
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, pm[static_cast<long long>(s) * rows_pad + i]);
  float L = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = pm[static_cast<long long>(s) * rows_pad + i];
    if (ms > -INFINITY) L += pl[static_cast<long long>(s) * rows_pad + i] * exp2f(ms - M);
  }
  const float v = M + log2f(L);
  lse2[i] = v;
  row_loss[i] = (v - ppos[i]) * 0.6931471805599453f;
}

}  // namespace cstp

using namespace cstp;

static int ntx_splits(int tiles) {
  // One CTA per SM (TMEM / shared memory): the kernel takes ceil(row tiles x splits / SMs) waves of `tiles / splits`
  // column tiles each.  Pick the split count with the shortest critical path; ties go to fewer splits (fewer partials).
  // (rows = 8192: 64 row tiles -> 2 splits, 128 CTAs, 32 tile-times; the old rule "two waves" gave 5 splits = 39.)
  const int sms = num_sms();
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= tiles; ++s) {
    const int per = (tiles + s - 1) / s;
    const int eff = (tiles + per - 1) / per;
    const long long waves = (1LL * tiles * eff + sms - 1) / sms;
    // every CTA also pays ~2 tile-times of fixed work (resident row tile, pipeline fill, dZn write-back), and the merge /
    // sum kernels read `eff` partials (measured: 16 splits of 4 tiles ran 2.4x slower per tile than 2 splits of 32)
    const long long cost = waves * (per + 2) * 16 + eff * 4;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

extern "C" long long cstp_ntxent_workspace_floats(int rows, int d) {
  if (rows < 2 || d < 1) return 0;
  const long long rp = (rows + 127) / 128 * 128;
  const long long tiles = rp / 128;
  const long long nsplit = ntx_splits(static_cast<int>(tiles));
  // fp32: norms, lse, row_loss, zn | ppos, lse2, pm, pl, dzn | bf16 (2 per float): zb, zt, W
  long long n = 3LL * rows + static_cast<long long>(rows) * d;
  n += 2 * rp + 4 * nsplit * rp + rp * d;            // ppos, lse2 | pm, pl (two column halves per split) | dzn
  const long long w_or_parts = d <= 128 ? nsplit * rp * d : (rp * rp + 1) / 2;       // fused backward: split partials of dZn
  n += (rp * d + 1) / 2 * 2 + w_or_parts;
  return n + 64;
}

namespace cstp {
// Tensor-core path of cstp_ntxent (called from loss_optim.cu once zn / norms are computed).  `ws` points behind the
// SIMT path's workspace prefix (3*rows + rows*d floats).
int ntxent_tensor_path(const float* zn, int rows, int d, float temperature, bool backward, float* row_loss, float* ws,
                       float** dzn_out, cudaStream_t stream);
}

int cstp::ntxent_tensor_path(const float* zn, int rows, int d, float temperature, bool backward, float* row_loss,
                             float* ws, float** dzn_out, cudaStream_t stream) {
  const int rp = (rows + 127) / 128 * 128;
  const int tiles = rp / 128;
  const int nsplit = ntx_splits(tiles);
  float* ppos = ws;
  float* lse2 = ppos + rp;
  float* pm = lse2 + rp;
  float* pl = pm + 2LL * nsplit * rp;
  float* dzn = pl + 2LL * nsplit * rp;
  __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(dzn + static_cast<long long>(rp) * d);
  __nv_bfloat16* zt = zb + static_cast<long long>(rp) * d;
  __nv_bfloat16* W = zt + static_cast<long long>(rp) * d;
  if ((reinterpret_cast<uintptr_t>(zb) % 16) != 0 || (reinterpret_cast<uintptr_t>(W) % 16) != 0)
    return fail_inval("ntxent workspace must be 16-byte aligned");

  {
    const long long total = static_cast<long long>(rp) * d;
    int blocks = ceil_div(total, 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    ntxent_cast_kernel<<<blocks, 256, 0, stream>>>(zn, rows, rp, d, zb, (backward && d > 128) ? zt : nullptr);
    CSTP_LAUNCHED();
  }
  NtxKParams k;
  memset(&k, 0, sizeof(k));
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)rp};
    const uint64_t strides[1] = {(uint64_t)d * 2};
    const uint32_t box[2] = {64u, 128u};
    int rc = encode_tmap_bf16(&k.zmap, zb, 2, dims, strides, box);
    if (rc != CSTP_OK) return rc;
  }
  k.rows = rows;
  k.rows_pad = rp;
  k.half = rows / 2;
  k.nkc = d / 64;
  k.tiles = tiles;
  k.tiles_per_split = ceil_div(tiles, nsplit);
  const int nsplit_eff = ceil_div(tiles, k.tiles_per_split);
  const uint32_t tile_bytes = static_cast<uint32_t>(k.nkc) * kNtBoxBytes;
  int stages = (kNtSmemLimit - 1024 - 256 - 4 * static_cast<int>(kNtWSlab) - static_cast<int>(tile_bytes)) /
               static_cast<int>(tile_bytes);
  if (stages > kNtMaxStages) stages = kNtMaxStages;
  if (stages < 2) return fail_inval("embedding dimension too large for the tensor-core NT-Xent path");
  k.stages = stages;
  k.idesc = umma_idesc_bf16(128, 128, 0, 0);
  k.scale_log2 = 1.4426950408889634f / temperature;
  k.wscale = 1.f / (temperature * static_cast<float>(rows));
  k.pm = pm;
  k.pl = pl;
  k.ppos = ppos;
  k.lse2 = lse2;
  k.W = W;
  const int smem = 1024 + static_cast<int>(tile_bytes) * (stages + 1) + 4 * static_cast<int>(kNtWSlab) + 256;
  static bool attr_set = false;
  if (!attr_set) {
    CSTP_CUDA(cudaFuncSetAttribute(ntxent_s_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNtSmemLimit));
    CSTP_CUDA(cudaFuncSetAttribute(ntxent_s_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNtSmemLimit));
    attr_set = true;
  }
  const dim3 grid(tiles, nsplit_eff);
  const int smem_launch = smem < 120 * 1024 ? 120 * 1024 : smem;     // one CTA per SM (TMEM ownership)
  ntxent_s_kernel<false><<<grid, kNtThreads2, smem_launch, stream>>>(k);
  CSTP_LAUNCHED();
  ntxent_merge_kernel<<<ceil_div(rp, 256), 256, 0, stream>>>(pm, pl, ppos, 2 * nsplit_eff, rows, rp, lse2, row_loss);
  CSTP_LAUNCHED();
  *dzn_out = dzn;
  if (!backward) return CSTP_OK;
  if (d <= 128) {
    // fused backward: no W in memory.  The region behind zt holds the per-split partials of dZn.
    NtxBwdParams b;
    memset(&b, 0, sizeof(b));
    b.zmap = k.zmap;
    b.rows = rows; b.rows_pad = rp; b.half = rows / 2; b.nkc = k.nkc; b.tiles = tiles; b.tiles_per_split = k.tiles_per_split;
    b.d = d;
    int bstages = (kNtSmemLimit - 1024 - 256 - 2 * static_cast<int>(kNtWTile) - static_cast<int>(tile_bytes)) /
                  static_cast<int>(tile_bytes);
    if (bstages > 4) bstages = 4;
    if (bstages < 2) return fail_inval("embedding dimension too large for the fused NT-Xent backward");
    b.stages = bstages;
    b.idesc_s = k.idesc;
    b.idesc_d = umma_idesc_bf16(128, static_cast<uint32_t>(d), 0, 1);
    b.scale_log2 = k.scale_log2;
    b.wscale = k.wscale;
    b.lse2 = lse2;
    b.dzn_part = reinterpret_cast<float*>(W);
    if ((reinterpret_cast<uintptr_t>(b.dzn_part) % 32) != 0) return fail_inval("ntxent workspace: dZn partials must be 32-byte aligned");
    static bool battr = false;
    if (!battr) {
      CSTP_CUDA(cudaFuncSetAttribute(ntxent_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNtSmemLimit));
      battr = true;
    }
    const int bsmem = 1024 + static_cast<int>(tile_bytes) * (bstages + 1) + 2 * static_cast<int>(kNtWTile) + 256;
    ntxent_bwd_fused_kernel<<<grid, kNtThreads2, bsmem < 120 * 1024 ? 120 * 1024 : bsmem, stream>>>(b);
    CSTP_LAUNCHED();
    const long long n4 = static_cast<long long>(rp) * d / 4;
    int blocks = ceil_div(n4, 256);
    if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
    ntxent_sum_splits_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(b.dzn_part), nsplit_eff, n4,
                                                        reinterpret_cast<float4*>(dzn));
    CSTP_LAUNCHED();
    return CSTP_OK;
  }
  ntxent_s_kernel<true><<<grid, kNtThreads, smem_launch, stream>>>(k);
  CSTP_LAUNCHED();
  // dZn [rows_pad][d] = W [rows_pad][rows_pad] . Zn : a 1-tap implicit GEMM over W with the transposed embeddings as the
  // packed weight matrix [d][rows_pad]
  cstp_conv_desc cd;
  memset(&cd, 0, sizeof(cd));
  cd.n_amaps = 1;
  cd.amap[0].ptr = W;
  cd.amap[0].dims[0] = rp;
  cd.amap[0].dims[1] = rp;
  cd.amap[0].dims[2] = 1;
  cd.amap[0].dims[3] = 1;
  cd.amap[0].dims[4] = 1;
  for (int i = 0; i < 4; ++i) cd.amap[0].strides[i] = static_cast<int64_t>(rp) * 2 * (i == 0 ? 1 : rp);
  cd.a_channels = rp;
  cd.n_taps = 1;
  cd.taps[0].map_id = 0;
  cd.w_packed = zt;
  cd.Np = d;
  cd.Ktot = rp;
  cd.n_tile = d;
  cd.Wt = rp;
  cd.Ht = cd.Tt = cd.Nt = 1;
  cd.bw = 128;
  cd.bh = cd.bt = cd.bn = 1;
  cd.out_f32 = dzn;
  cd.osw = d;
  cd.osh = cd.ost = cd.osn = static_cast<int64_t>(rp) * d;
  cstp_conv_plan* plan = nullptr;
  int rc = cstp_conv_plan_create(&cd, &plan);
  if (rc != CSTP_OK) return rc;
  rc = cstp_conv_plan_run(plan, stream);
  cstp_conv_plan_destroy(plan);
  return rc;        // the caller applies the cosine-normalisation Jacobian to dzn
}
