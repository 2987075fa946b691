// Causal flash-attention BACKWARD on tcgen05 / TMEM / TMA for sm_100a, head_dim 64, bf16, MHA + GQA.
//
//   P  = exp(S*scale - lse)          S  = Q K^T
//   dV = P^T dO                      dP = dO V^T
//   dS = P * (dP - D) * scale        D  = rowsum(dO * O)        (pre-pass kernel)
//   dK = dS^T Q                      dQ = dS K                  (dQ: fp32 TMA reduce-add across key tiles, post-pass -> bf16)
//
// One CTA per (batch, kv-head, 128-key tile); it keeps K and V in shared memory and dK/dV accumulators in tensor memory
// and walks over the query heads of its GQA group and the causal range of 128-query tiles.  Five GEMMs per tile pair, all
// tcgen05.mma from shared-memory descriptors:
//     S  = Q  K^T    A = Q  (K-major)   B = K  (K-major)    N = 128
//     dP = dO V^T    A = dO (K-major)   B = V  (K-major)    N = 128
//     dV += P^T  dO  A = P  (MN-major)  B = dO (MN-major)   N = 64      P, dS: bf16 tiles written by the softmax warps
//     dK += dS^T Q   A = dS (MN-major)  B = Q  (MN-major)   N = 64
//     dQ  = dS   K   A = dS (K-major)   B = K  (MN-major)   N = 64
// The same [128 x 64] swizzled tiles serve as K-major and as MN-major operands - only the descriptor changes.
// TMEM: S [0,128) dP [128,256) dV [256,320) dK [320,384) dQ [384,448).   512 threads in four warpgroups with their own
// register budgets (setmaxnreg): warp 0 TMA + warp 1 MMA (96 regs), warps 4-11 element-wise (two threads per query row,
// 64 keys each; 160 regs) + dK/dV epilogue, warps 12-15 dQ drain (TMEM -> staging -> TMA reduce-add; 96 regs).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "attn_math.cuh"

using namespace odb;
using namespace sm100;
using namespace attn_math;

namespace attn_bwd {

constexpr int BQ = 128, BKV = 128, D = 64;
constexpr int TILE = 128 * 128;                  // bytes of a [128 x 64] bf16 tile
constexpr int THREADS = 512;                   // warp 0 TMA, 1 MMA, 2-3 idle | 4-11 element-wise | 12-15 dQ drain
constexpr int EW_THREADS = 256, DRAIN_THREADS = 128;
constexpr int EW_WARP0 = 4, DRAIN_WARP0 = 12;
// K, V, 2 x (Q, dO), P (2 chunks), dS (2 chunks), dQ staging fp32 [128 x 64] = 2 chunks of 128 B rows
constexpr int SMEM_BYTES = 2 * TILE + 4 * TILE + 2 * TILE + 2 * TILE + 2 * TILE + 1024 + 256;
constexpr uint32_t TMEM_COLS = 512;

struct Params {
  int B, S, Hq, Hkv;
  float scale, scale_log2;
  const float* lse;      // [B, Hq, S] natural log (forward)
  const float* dsum;     // [B, Hq, S] rowsum(dO * O)
  long long* dbg;        // optional timeline of CTA (0,0,0): [iteration][16] clock64 stamps (nullptr = off)
};

#define BWD_STAMP(slot) do { if (dbg_on) p.dbg[it * 16 + (slot)] = clock64(); } while (0)

template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dk,
                const __grid_constant__ CUtensorMap tmap_dv, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE;
  uint8_t* sQ = sV + TILE;                 // 2 stages
  uint8_t* sdO = sQ + 2 * TILE;            // 2 stages
  uint8_t* sP = sdO + 2 * TILE;            // [2 key-chunks][128 q rows][128 B]
  uint8_t* sdS = sP + 2 * TILE;
  uint8_t* sdQ = sdS + 2 * TILE;           // fp32 staging: [2 column-chunks][128 rows][128 B]
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(sdQ + 2 * TILE);
  uint64_t* kv_empty = kv_full + 1;        // the key tile's last MMAs retired: K / V may be replaced
  uint64_t* q_full = kv_empty + 1;         // [2]
  uint64_t* q_empty = q_full + 2;          // [2]
  uint64_t* sdp_full = q_empty + 2;        // S and dP ready
  uint64_t* sdp_free = sdp_full + 1;       // S and dP live in registers: the next S/dP MMAs may overwrite the columns (8 warps)
  uint64_t* pds_full = sdp_free + 1;       // P and dS written to smem (8 warp arrivals)
  uint64_t* mma_done = pds_full + 1;       // dV/dK/dQ MMAs of this iteration retired (P/dS buffers + dQ accumulator ready)
  uint64_t* dq_empty = mma_done + 1;       // dQ accumulator drained (4 warp arrivals)
  uint64_t* dkv_empty = dq_empty + 1;      // dK/dV accumulators read by the epilogue (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dkv_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = p.S / BQ;
  // Each CTA owns TWO key tiles of one (batch, kv head): tile x (nq-x query tiles to visit) and tile nq-1-x (x+1), so every
  // CTA runs group*(nq+1) iterations - a balanced grid - and pays the fixed costs once.  `g` is the CTA-wide iteration
  // index all per-iteration barrier parities derive from.
  const int kb_first = (int)blockIdx.x, kb_second = nq - 1 - (int)blockIdx.x;
  const int nitems = kb_first != kb_second ? 2 : 1;
  const int hk = blockIdx.y, b = blockIdx.z;
  const int group = p.Hq / p.Hkv;
  const int col_k = (p.Hq + hk) * D, col_v = (p.Hq + p.Hkv + hk) * D;
  const bool dbg_cta = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 8);
    mbar_init(pds_full, 8);
    mbar_init(mma_done, 1);
    mbar_init(dq_empty, 4);
    mbar_init(dkv_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = *tmem_slot;
  const uint32_t tS = tb, tdP = tb + 128, tdV = tb + 256, tdK = tb + 320, tdQ = tb + 384;

  // register file split per warpgroup: 4 x 128 threads x {96, 160, 160, 96} = 64K registers
  // (each setmaxnreg sits inside its role branch and the branches only re-join at the final barrier, so ptxas allocates
  // every role's code against its own budget)
  if (warp < EW_WARP0) {
    reg_dealloc<96>();
    if (warp == 0 && lane == 0) {
      {
      for (int w = 0, g = 0; w < nitems; ++w) {
        const int kb = w ? kb_second : kb_first;          // causal: query tiles kb..nq-1
        mbar_wait(kv_empty, (w & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full, 2 * TILE);
        tma_load_2d(sK, &tmap_qkv, kv_full, col_k, b * p.S + kb * BKV);
        tma_load_2d(sV, &tmap_qkv, kv_full, col_v, b * p.S + kb * BKV);
        for (int hh = 0; hh < group; ++hh) {
          const int h = hk * group + hh;
          for (int qi = kb; qi < nq; ++qi, ++g) {
            const int st = g & 1;
            mbar_wait(&q_empty[st], ((g >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&q_full[st], 2 * TILE);
            tma_load_2d(sQ + st * TILE, &tmap_qkv, &q_full[st], h * D, b * p.S + qi * BQ);
            tma_load_2d(sdO + st * TILE, &tmap_do, &q_full[st], h * D, b * p.S + qi * BQ);
          }
        }
      }
    }
    } else if (warp == 1) {
    constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);     // S, dP
    constexpr uint32_t id_t = make_idesc_bf16(128, 64, 1, 1);      // dV, dK  (both operands MN-major)
    constexpr uint32_t id_q = make_idesc_bf16(128, 64, 0, 1);      // dQ      (A K-major, B MN-major)
    const bool leader = elect_one();
    const uint64_t kd = make_smem_desc_sw128(smem_u32(sK), 16, 1024);        // K as K-major B (S) / MN-major B (dQ)
    const uint64_t vd = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
    const uint64_t pd_mn = make_smem_desc_sw128(smem_u32(sP), TILE, 1024);   // P^T: 2 MN chunks (64 keys) TILE bytes apart
    const uint64_t dsd_mn = make_smem_desc_sw128(smem_u32(sdS), TILE, 1024);
    const uint64_t dsd_k0 = make_smem_desc_sw128(smem_u32(sdS), 16, 1024);   // dS as K-major A: key chunk 0 / 1
    const uint64_t dsd_k1 = make_smem_desc_sw128(smem_u32(sdS + TILE), 16, 1024);
    const bool dbg_on = dbg_cta && lane == 0;
    // Software pipeline (the tensor pipe executes in issue order):
    //   S/dP(g+1) is issued as soon as the element-wise warps hold S/dP(g) in registers, so it runs under their math;
    //   dV/dK(g) follow once P/dS(g) are in shared memory, dQ(g) once the drain warps emptied the previous dQ tile.
    auto issue_sdp = [&](int it) {
      const int st = it & 1;
      const uint64_t qd = make_smem_desc_sw128(smem_u32(sQ + st * TILE), 16, 1024);
      const uint64_t dod = make_smem_desc_sw128(smem_u32(sdO + st * TILE), 16, 1024);
      mbar_wait(&q_full[st], (it >> 1) & 1);
      tc_fence_after_sync();
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS, qd + 2 * k, kd + 2 * k, id_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tdP, dod + 2 * k, vd + 2 * k, id_s, k > 0);
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    for (int w = 0, g = 0; w < nitems; ++w) {
      const int iters = group * (nq - (w ? kb_second : kb_first));
      mbar_wait(kv_full, w & 1);
      if (g > 0) mbar_wait(sdp_free, (g - 1) & 1);       // the previous key tile's last S/dP have been read
      issue_sdp(g);
      for (int i = 0; i < iters; ++i) {
        const int it = g + i;
        const int st = it & 1;
        const uint64_t qd = make_smem_desc_sw128(smem_u32(sQ + st * TILE), 16, 1024);
        const uint64_t dod = make_smem_desc_sw128(smem_u32(sdO + st * TILE), 16, 1024);
        BWD_STAMP(9);
        if (i + 1 < iters) {
          mbar_wait(sdp_free, it & 1);
          issue_sdp(it + 1);
        }
        BWD_STAMP(10);
        mbar_wait(pds_full, it & 1);                     // P, dS in shared memory
        if (i == 0 && w > 0) mbar_wait(dkv_empty, (w - 1) & 1);   // the epilogue read the previous key tile's dK / dV
        tc_fence_after_sync();
        BWD_STAMP(11);
        const bool first = (i == 0);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tdV, pd_mn + 128 * k, dod + 128 * k, id_t, (first && k == 0) ? 0u : 1u);
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tdK, dsd_mn + 128 * k, qd + 128 * k, id_t, (first && k == 0) ? 0u : 1u);
        }
        __syncwarp();
        if (it > 0) mbar_wait(dq_empty, (it - 1) & 1);   // previous dQ tile drained
        tc_fence_after_sync();
        BWD_STAMP(12);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tdQ, ((k >> 2) ? dsd_k1 : dsd_k0) + 2 * (k & 3), kd + 128 * k, id_q, k > 0);
          umma_commit(&q_empty[st]);
          umma_commit(mma_done);
          if (i + 1 == iters) umma_commit(kv_empty);
        }
        __syncwarp();
      }
      g += iters;
    }
    }
  } else if (warp >= DRAIN_WARP0) {
    reg_dealloc<96>();
    // ------------------------------------------------------------------ dQ drain warps
    // dQ tile of iteration `it`: TMEM -> fp32 staging -> TMA reduce-add into the fp32 accumulator, off the critical path
    // of the element-wise warps.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool elected = threadIdx.x == DRAIN_WARP0 * 32;
    const bool dbg_on = dbg_cta && elected;
    for (int w = 0, it = 0; w < nitems; ++w) {
      const int kb = w ? kb_second : kb_first;
      for (int hh = 0; hh < group; ++hh) {
        const int h = hk * group + hh;
        for (int qi = kb; qi < nq; ++qi, ++it) {
          mbar_wait(mma_done, it & 1);
          tc_fence_after_sync();
          uint32_t dq[64];
          tmem_ld_32x32b_x32(tdQ + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&dq[0]));
          tmem_ld_32x32b_x32(tdQ + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&dq[32]));
          tmem_ld_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(dq_empty);
          BWD_STAMP(6);
          if (elected) tma_store_wait_read<0>();         // previous reduce finished reading the staging tile
          named_bar_sync(3, DRAIN_THREADS);
          BWD_STAMP(7);
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<uint4*>(sdQ + c * TILE + row * 128 + ((u ^ (row & 7)) * 16)) =
                  make_uint4(dq[c * 32 + u * 4 + 0], dq[c * 32 + u * 4 + 1], dq[c * 32 + u * 4 + 2], dq[c * 32 + u * 4 + 3]);
          fence_proxy_async_smem();
          named_bar_sync(4, DRAIN_THREADS);
          if (elected) {
            tma_reduce_add_2d(&tmap_dq, sdQ, h * D, b * p.S + qi * BQ);
            tma_reduce_add_2d(&tmap_dq, sdQ + TILE, h * D + 32, b * p.S + qi * BQ);
            tma_store_commit();
          }
          BWD_STAMP(8);
        }
      }
    }
    if (elected) tma_store_wait<0>();
  } else {
    // ------------------------------------------------------------------ element-wise warps
    reg_alloc<160>();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = (warp - EW_WARP0) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c = p.scale_log2;
    const uint64_t c2 = pack2(c, c), sc2 = pack2(p.scale, p.scale);
    const bool elected = threadIdx.x == EW_WARP0 * 32;
    const bool dbg_on = dbg_cta && elected;
    for (int w = 0, it = 0; w < nitems; ++w) {
      const int kb = w ? kb_second : kb_first;
      for (int hh = 0; hh < group; ++hh) {
        const int h = hk * group + hh;
        for (int qi = kb; qi < nq; ++qi, ++it) {
          const size_t stat = ((size_t)b * p.Hq + h) * p.S + qi * BQ + row;
          const float lse2 = p.lse[stat] * 1.4426950408889634f;
          const float dsum = p.dsum[stat];
          const bool diag = (qi == kb);
          BWD_STAMP(0);
          mbar_wait(sdp_full, it & 1);
          tc_fence_after_sync();
          BWD_STAMP(1);
          uint32_t pp[32], ds[32];
          {
            uint32_t sr[64], dr[64];
            tmem_ld_32x32b_x32(tS + lane_off + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
            tmem_ld_32x32b_x32(tS + lane_off + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
            tmem_ld_32x32b_x32(tdP + lane_off + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&dr[0]));
            tmem_ld_32x32b_x32(tdP + lane_off + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&dr[32]));
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_free);        // S/dP(it+1) may now be computed into the same columns
            BWD_STAMP(2);
            if (diag) {
#pragma unroll
              for (int i = 0; i < 64; ++i)
                if (half * 64 + i > row) sr[i] = 0xff800000u;     // -inf -> P = 0
            }
            // P = exp2(S*c - lse2), dS = P * (dP*scale - dsum*scale): packed fp32x2 math; of every four exponentials two
            // go through the MUFU and two through the FMA-pipe polynomial (see attn_sm100.cu)
            const uint64_t nl2 = pack2(-lse2, -lse2), nd2 = pack2(-dsum * p.scale, -dsum * p.scale);
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
              const uint64_t ya = ffma2(pack2u(sr[i], sr[i + 1]), c2, nl2);
              const uint64_t yb = ffma2(pack2u(sr[i + 2], sr[i + 3]), c2, nl2);
              float a0, a1;
              unpack2(ya, a0, a1);
              const uint64_t pa = pack2(fast_exp2(a0), fast_exp2(a1));
              const uint64_t pb = poly_exp2x2(yb);
              const uint64_t da = fmul2(pa, ffma2(pack2u(dr[i], dr[i + 1]), sc2, nd2));
              const uint64_t db = fmul2(pb, ffma2(pack2u(dr[i + 2], dr[i + 3]), sc2, nd2));
              pp[i / 2] = cvt_bf16x2(pa);
              pp[i / 2 + 1] = cvt_bf16x2(pb);
              ds[i / 2] = cvt_bf16x2(da);
              ds[i / 2 + 1] = cvt_bf16x2(db);
            }
          }
          BWD_STAMP(3);
          if (it > 0) {
            mbar_wait(mma_done, (it - 1) & 1);           // the previous tile's dV/dK/dQ MMAs no longer read P / dS
            tc_fence_after_sync();
          }
          if (w > 0 && hh == 0 && qi == kb) {
            // P / dS doubled as the dK / dV store staging of the previous key tile: its TMA store must have read them
            if (elected) tma_store_wait_read<0>();
            named_bar_sync(1, EW_THREADS);
          }
          BWD_STAMP(4);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int off = half * TILE + row * 128 + ((u ^ (row & 7)) * 16);
            *reinterpret_cast<uint4*>(sP + off) = make_uint4(pp[u * 4 + 0], pp[u * 4 + 1], pp[u * 4 + 2], pp[u * 4 + 3]);
            *reinterpret_cast<uint4*>(sdS + off) = make_uint4(ds[u * 4 + 0], ds[u * 4 + 1], ds[u * 4 + 2], ds[u * 4 + 3]);
          }
          fence_proxy_async_smem();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(pds_full);
          BWD_STAMP(5);
        }
      }
      // ---- dK / dV epilogue of this key tile: this thread's 32 columns -> bf16 -> staging (the P buffer) -> TMA store
      mbar_wait(mma_done, (it - 1) & 1);
      tc_fence_after_sync();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tdK + lane_off + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[u * 8 + i]);
        *reinterpret_cast<uint4*>(sP + row * 128 + (((half * 4 + u) ^ (row & 7)) * 16)) = pack8(f);
      }
      tmem_ld_32x32b_x32(tdV + lane_off + half * 32, r);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkv_empty);             // the next key tile may start accumulating
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[u * 8 + i]);
        *reinterpret_cast<uint4*>(sP + TILE + row * 128 + (((half * 4 + u) ^ (row & 7)) * 16)) = pack8(f);
      }
      fence_proxy_async_smem();
      named_bar_sync(2, EW_THREADS);
      if (elected) {
        tma_store_2d(&tmap_dk, sP, hk * D, b * p.S + kb * BKV);
        tma_store_2d(&tmap_dv, sP + TILE, hk * D, b * p.S + kb * BKV);
        tma_store_commit();
      }
    }
    if (elected) tma_store_wait<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, TMEM_COLS);
}

// D[b,h,s] = sum_d dO * O   (one warp per (token, head): 64 elements)
__global__ void __launch_bounds__(256) attn_dsum_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                                                        float* __restrict__ dsum, int B, int S, int Hq, long long ld) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)B * S * Hq;
  if (w >= total) return;
  const long long t = w / Hq;
  const int h = (int)(w % Hq);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(dout + t * ld + h * 64 + lane * 2);
  const __nv_bfloat162 o = *reinterpret_cast<const __nv_bfloat162*>(out + t * ld + h * 64 + lane * 2);
  float s = __bfloat162float(a.x) * __bfloat162float(o.x) + __bfloat162float(a.y) * __bfloat162float(o.y);
  s = warp_sum(s);
  if (lane == 0) dsum[((t / S) * Hq + h) * S + (t % S)] = s;
}

}  // namespace attn_bwd

static long long* g_bwd_dbg = nullptr;
// debugging aid: device buffer [iterations][16] that receives the clock64 timeline of CTA (0,0,0); nullptr switches it off
ODB_EXPORT int odb_attn_bwd_set_dbg(void* buf) {
  g_bwd_dbg = (long long*)buf;
  return 0;
}

// dq_acc: fp32 [T, Hq*64] ZERO-INITIALISED by the caller (accumulated with TMA reduce-add); dk, dv: bf16 [T, Hkv*64].
ODB_EXPORT int odb_attn_bwd(const void* qkv, const void* out, const void* dout, const void* lse, void* dsum, void* dq_acc,
                            void* dk, void* dv, int B, int S, int Hq, int Hkv, long long ld_qkv, long long ld_out,
                            float softmax_scale, cudaStream_t st) {
  using namespace attn_bwd;
  if (S % BQ || Hq % Hkv || ld_qkv % 8 || ld_out % 8) return -1;
  const long long T = (long long)B * S;
  {
    const long long warps = T * Hq;
    attn_dsum_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)out, (float*)dsum,
                                                                  B, S, Hq, ld_out);
  }
  CUtensorMap tq, tdo, tdq, tdk, tdv;
  int rc;
  if ((rc = make_tmap_2d(&tq, qkv, T, (Hq + 2 * Hkv) * D, ld_qkv * 2, 128, 64, 2))) return rc;
  if ((rc = make_tmap_2d(&tdo, dout, T, Hq * D, ld_out * 2, 128, 64, 2))) return rc;
  if ((rc = make_tmap_2d(&tdq, dq_acc, T, Hq * D, (long long)Hq * D * 4, 128, 32, 4))) return rc;
  if ((rc = make_tmap_2d(&tdk, dk, T, Hkv * D, (long long)Hkv * D * 2, 128, 64, 2))) return rc;
  if ((rc = make_tmap_2d(&tdv, dv, T, Hkv * D, (long long)Hkv * D * 2, 128, 64, 2))) return rc;
  Params p{};
  p.B = B; p.S = S; p.Hq = Hq; p.Hkv = Hkv;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.lse = (const float*)lse;
  p.dsum = (const float*)dsum;
  p.dbg = g_bwd_dbg;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((S / BKV + 1) / 2, Hkv, B);
  attn_bwd_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tq, tdo, tdq, tdk, tdv, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
