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
__global__ void __launch_bounds__(kNtThreads, 1) ntxent_s_kernel(const __grid_constant__ NtxKParams p) {
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
      mbar_init(&tempty[a], 4);
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
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: lane owns row i of the row tile
    const int q = warp - 4;
    const int i = rt * 128 + q * 32 + lane;
    const int pos = i < p.half ? i + p.half : i - p.half;
    const bool row_ok = i < p.rows;
    float m = -INFINITY, l = 0.f, spos = 0.f;
    float lse_i = 0.f;
    if (kBackward) lse_i = row_ok ? p.lse2[i] : 0.f;
    int it = 0;
    for (int jt = jt0; jt < jt1; ++jt, ++it) {
      const int as = it & 1;
      mbar_wait(&tfull[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * 128);
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        const int j0 = jt * 128 + c0;
        if (!kBackward) {
          float s[32];
          float tm = -INFINITY;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int j = j0 + c;
            float x = __uint_as_float(v[c]) * p.scale_log2;
            if (j == pos) spos = x;
            x = (j == i || j >= p.rows) ? -INFINITY : x;
            s[c] = x;
            tm = fmaxf(tm, x);
          }
          const float mn = fmaxf(m, tm);
          if (mn > -INFINITY) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) acc += exp2f(s[c] - mn);
            l = l * exp2f(m - mn) + acc;
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
      const long long o = static_cast<long long>(split) * p.rows_pad + i;
      p.pm[o] = m;
      p.pl[o] = l;
      if (row_ok && pos >= jt0 * 128 && pos < jt1 * 128) p.ppos[i] = spos;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
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
    zt[static_cast<long long>(c) * rows_pad + r] = v;
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
  // enough CTAs to fill the machine: row tiles x column splits >= ~2 waves where the problem allows
  int want = (2 * num_sms() + tiles - 1) / tiles;
  if (want > tiles) want = tiles;
  if (want < 1) want = 1;
  return want;
}

extern "C" long long cstp_ntxent_workspace_floats(int rows, int d) {
  if (rows < 2 || d < 1) return 0;
  const long long rp = (rows + 127) / 128 * 128;
  const long long tiles = rp / 128;
  const long long nsplit = ntx_splits(static_cast<int>(tiles));
  // fp32: norms, lse, row_loss, zn | ppos, lse2, pm, pl, dzn | bf16 (2 per float): zb, zt, W
  long long n = 3LL * rows + static_cast<long long>(rows) * d;
  n += 2 * rp + 2 * nsplit * rp + rp * d;
  n += (rp * d + 1) / 2 * 2 + (rp * rp + 1) / 2;
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
  float* pl = pm + static_cast<long long>(nsplit) * rp;
  float* dzn = pl + static_cast<long long>(nsplit) * rp;
  __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(dzn + static_cast<long long>(rp) * d);
  __nv_bfloat16* zt = zb + static_cast<long long>(rp) * d;
  __nv_bfloat16* W = zt + static_cast<long long>(rp) * d;
  if ((reinterpret_cast<uintptr_t>(zb) % 16) != 0 || (reinterpret_cast<uintptr_t>(W) % 16) != 0)
    return fail_inval("ntxent workspace must be 16-byte aligned");

  {
    const long long total = static_cast<long long>(rp) * d;
    int blocks = ceil_div(total, 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    ntxent_cast_kernel<<<blocks, 256, 0, stream>>>(zn, rows, rp, d, zb, zt);
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
  ntxent_s_kernel<false><<<grid, kNtThreads, smem_launch, stream>>>(k);
  CSTP_LAUNCHED();
  ntxent_merge_kernel<<<ceil_div(rp, 256), 256, 0, stream>>>(pm, pl, ppos, nsplit_eff, rows, rp, lse2, row_loss);
  CSTP_LAUNCHED();
  *dzn_out = dzn;
  if (!backward) return CSTP_OK;
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
