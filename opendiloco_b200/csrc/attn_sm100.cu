// Causal flash-attention forward on tcgen05 / TMEM / TMA for sm_100a, head_dim 64, bf16, MHA + GQA.
//
// Reads q/k/v straight out of the packed projection buffer qkv[T, (Hq + 2*Hkv) * 64] (token-major rows: no transposes)
// through ONE tensor map, writes out[T, Hq*64] token-major (the o-proj GEMM's A operand) and lse[B, Hq, S].
//
// One CTA per (batch, q-head, 128-query block); two CTAs co-reside per SM (112 KB smem, 256 TMEM columns each) so one
// CTA's softmax overlaps the other's MMAs.  320 threads:
//   warp 0     TMA: Q once, then K_j / V_j 128-key tiles through a 2-stage mbarrier ring
//   warp 1     MMA: S = Q K_j^T  (tcgen05.mma M128 N128 K16 x4, both operands K-major, accumulator S in TMEM)
//                   O += P V_j   (M128 N64 K16 x8; A = P from shared memory K-major, B = V_j tile MN-major)
//   warps 2-9  softmax: TWO threads per query row (TMEM lane == row; warps 2-5 take keys 0-63 of the tile, warps 6-9 keys
//              64-127, row max exchanged through shared memory): tcgen05.ld S, online max/sum with exp2, causal mask
//              on the diagonal tile, rescale O in TMEM (tcgen05.ld/st), write P (bf16) into the swizzled smem operand;
//              epilogue: O / l -> bf16 -> swizzled staging tile -> TMA store
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "packed_math.cuh"

using namespace odb;
using namespace sm100;
using namespace packed_math;

namespace attn {

constexpr int BQ = 128, BKV = 128, D = 64;
constexpr int TILE_BYTES = 128 * 128;           // a [128 x 64] bf16 tile (128-byte rows)
constexpr int THREADS = 320;
constexpr int SOFTMAX_THREADS = 256;
// 7 tiles + barriers = 114,816 B: two CTAs per SM need 2 x (smem + 1 KB reserved) <= 227 KB, so there is no slack for a
// manual 1024-byte round-up - the dynamic window is declared 1024-aligned instead (it starts at the CTA's smem base)
constexpr int SMEM_BYTES = TILE_BYTES /*Q*/ + 2 * TILE_BYTES /*K x2*/ + TILE_BYTES /*V x1*/ + TILE_BYTES /*out staging*/ + 128 /*barriers*/;   // 82 KB + 2 KB static
// the row-max / row-sum exchange between the two column halves lives in the (otherwise unused) tail of the barrier block
// plus a small static array
constexpr uint32_t TMEM_COLS = 256;             // S: [0,128)  O: [128,192)  P (bf16, 2 keys per column): [192,256)

struct Params {
  int B, S, Hq, Hkv;
  float scale_log2;        // softmax_scale * log2(e)
  float* lse;              // [B, Hq, S], natural log
  long long* dbg;          // optional timeline of CTA (0,0,0): [tile][16] clock64 stamps (nullptr = off)
};

#define ATT_STAMP(slot) do { if (dbg_on) p.dbg[j * 16 + (slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out, Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ float red[2 * 256];                 // row max / row sum exchange between the two column halves (double-buffered)
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();     // swizzled tiles need 1024-byte alignment
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + 2 * TILE_BYTES;            // single V stage: V_{j+1} is only needed after softmax_{j+1}
  uint8_t* sOut = sV + TILE_BYTES;              // output staging tile (P lives in tensor memory)
  uint64_t* q_full = reinterpret_cast<uint64_t*>(sOut + TILE_BYTES);
  uint64_t* k_full = q_full + 1;                // [2]
  uint64_t* k_empty = k_full + 2;               // [2]
  uint64_t* v_full = k_empty + 2;
  uint64_t* v_empty = v_full + 1;
  uint64_t* s_full = v_empty + 1;
  uint64_t* p_full = s_full + 1;
  uint64_t* s_free = p_full + 1;                // S_j is in registers: the next QK may overwrite the S columns
  uint64_t* o_done = s_free + 1;
  uint64_t* q_empty = o_done + 1;               // every QK^T of the current query tile retired: Q may be replaced
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = p.S / BQ;
  // Each CTA owns TWO query tiles of one (batch, head): tile nq-1-x (long) and tile x (short).  Every CTA therefore
  // walks nq+1 key tiles - a perfectly balanced grid - and the fixed cost (barrier init, TMEM allocation, pipeline
  // fill) is paid once per ~nq+1 tiles instead of once per ~nq/2.  `tg` below is the CTA-wide running tile index that
  // all barrier parities are derived from.
  const int qb[2] = {nq - 1 - (int)blockIdx.x, (int)blockIdx.x};
  const int nitems = qb[0] != qb[1] ? 2 : 1;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int col_q = h * D, col_k = (p.Hq + hk) * D, col_v = (p.Hq + p.Hkv + hk) * D;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_out);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);      // one arrival per softmax warp
    mbar_init(s_free, 8);
    mbar_init(o_done, 1);
    mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 192;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // K runs one tile ahead of V: K_{j+1} is requested as soon as its stage is free (QK_{j-1} retired), V_j after
      // PV_{j-1} retired - V_j is not needed before softmax_j is done, K_{j+1} is needed right after softmax_j's S load.
      auto load_k = [&](int j, int tg) {
        const int st = tg & 1;
        mbar_wait(&k_empty[st], ((tg >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_2d(sK + st * TILE_BYTES, &tmap_qkv, &k_full[st], col_k, b * p.S + j * BKV);
      };
      for (int w = 0, t0 = 0; w < nitems; ++w) {
        const int nkv = qb[w] + 1;               // causal: key tiles 0..qblk
        mbar_wait(q_empty, (w & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, TILE_BYTES);
        tma_load_2d(sQ, &tmap_qkv, q_full, col_q, b * p.S + qb[w] * BQ);
        load_k(0, t0);
        for (int j = 0; j < nkv; ++j) {
          const int tg = t0 + j;
          if (j + 1 < nkv) load_k(j + 1, tg + 1);
          mbar_wait(v_empty, (tg & 1) ^ 1);
          mbar_arrive_expect_tx(v_full, TILE_BYTES);
          tma_load_2d(sV, &tmap_qkv, v_full, col_v, b * p.S + j * BKV);
        }
        t0 += nkv;
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);   // S[128,128] = Q (K-major) x K^T (K-major)
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, D, 0, 1);     // O[128,64] = P (TMEM, K-major) x V (smem, MN-major)
    const bool leader = elect_one();             // the same lane issues every MMA / commit of this CTA
    const uint64_t qd = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t vd = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
    const uint64_t kd0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t kd1 = make_smem_desc_sw128(smem_u32(sK + TILE_BYTES), 16, 1024);
    const bool dbg_cta = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
    auto issue_qk = [&](int tg, bool last) {
      const int st = tg & 1;
      mbar_wait(&k_full[st], (tg >> 1) & 1);
      tc_fence_after_sync();
      if (leader) {
        const uint64_t kd = st ? kd1 : kd0;
        umma_ss(tS, qd, kd, idesc_qk, 0u);
        umma_ss(tS, qd + 2, kd + 2, idesc_qk, 1u);
        umma_ss(tS, qd + 4, kd + 4, idesc_qk, 1u);
        umma_ss(tS, qd + 6, kd + 6, idesc_qk, 1u);
        umma_commit(s_full);
        umma_commit(&k_empty[st]);               // K stage reusable once this QK retired
        if (last) umma_commit(q_empty);          // ... and so is Q after the tile's last QK
      }
      __syncwarp();
    };
    for (int w = 0, t0 = 0; w < nitems; ++w) {
    const int nkv = qb[w] + 1;
    const bool dbg_on = dbg_cta && w == 0;
    mbar_wait(q_full, w & 1);
    if (w > 0) mbar_wait(s_free, (t0 - 1) & 1);  // the previous query tile's last S has been read
    issue_qk(t0, nkv == 1);                      // S_0 = Q K_0^T
    for (int j = 0; j < nkv; ++j) {
      const int tg = t0 + j;
      ATT_STAMP(8);
      // As soon as the softmax warps hold S_j in registers, the next QK^T goes into the same TMEM columns - it runs while
      // they compute the probabilities, so S_{j+1} is waiting for them when they come back.
      if (j + 1 < nkv) {
        mbar_wait(s_free, tg & 1);
        issue_qk(tg + 1, j + 2 == nkv);
      }
      ATT_STAMP(9);
      mbar_wait(p_full, tg & 1);                 // P_j in tensor memory, O rescaled
      ATT_STAMP(10);
      mbar_wait(v_full, tg & 1);
      tc_fence_after_sync();
      if (leader) {
        // A operand straight from tensor memory: 16 keys of P = 8 columns per K step, no shared-memory read for A
        umma_ts(tO, tP, vd, idesc_pv, j > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 1; k < BKV / 16; ++k) umma_ts(tO, tP + 8 * k, vd + 128 * k, idesc_pv, 1u);
        umma_commit(v_empty);                    // V stage reusable
        umma_commit(o_done);                     // O (and P) stable
      }
      __syncwarp();
      ATT_STAMP(11);
    }
    t0 += nkv;
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;               // query row inside the tile == TMEM lane
    const int half = (warp - 2) >> 2;            // 0: keys [0,64) of every tile and O columns [0,32) ; 1: the other halves
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c = p.scale_log2;
    const bool dbg_cta = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64;
    for (int w = 0, t0 = 0; w < nitems; ++w) {
    const int qblk = qb[w], nkv = qblk + 1;
    const int row0 = b * p.S + qblk * BQ;        // first token row of this query tile
    const bool dbg_on = dbg_cta && w == 0;
    float m = -INFINITY, l = 0.f;                // l: this thread's half of the row sum
    for (int j = 0; j < nkv; ++j) {
      const int tg = t0 + j;
      ATT_STAMP(0);
      mbar_wait(s_full, tg & 1);
      tc_fence_after_sync();
      ATT_STAMP(1);
      const bool diag = (j == qblk);
      uint32_t sr[64];
      tmem_ld_32x32b_x32(tS + lane_off + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld_32x32b_x32(tS + lane_off + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);        // S_j now lives in registers
      ATT_STAMP(2);
      if (diag) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (half * 64 + i > row) sr[i] = 0xff800000u;      // -inf
      }
      float mx4[4] = {m, m, m, m};
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mx4[0] = fmaxf(mx4[0], __uint_as_float(sr[i]));
        mx4[1] = fmaxf(mx4[1], __uint_as_float(sr[i + 1]));
        mx4[2] = fmaxf(mx4[2], __uint_as_float(sr[i + 2]));
        mx4[3] = fmaxf(mx4[3], __uint_as_float(sr[i + 3]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // combine with the thread that owns the other 64 keys of this row
      red[(tg & 1) * 256 + half * 128 + row] = mx;
      named_bar_sync(2, SOFTMAX_THREADS);
      ATT_STAMP(3);
      mx = fmaxf(mx, red[(tg & 1) * 256 + (half ^ 1) * 128 + row]);
      // lazy rescaling: the reference point m only moves when the row maximum grew by more than 8 in the exp2 domain;
      // until then probabilities may exceed 1 (<= 2^8), which fp32 sums and bf16 P represent without trouble, and O / l
      // need no correction.  Both threads of a row see the same mx, so they take the same decision.
      if ((mx - m) * c <= 8.f) mx = m;
      const float alpha = fast_exp2((m - mx) * c);   // m = -inf on the first tile -> 0
      const float mc = mx * c;
      // p = exp2(s*c - m*c) with packed fp32x2 arithmetic (FFMA2 / FADD2).  Of every four elements two go through the
      // MUFU (ex2.approx) and two through a Cody-Waite + degree-3 polynomial on the FMA pipe, so both pipes work in
      // parallel (16384 exponentials per tile would otherwise keep the 16-lane MUFU busy for 1024 cycles per SM).
      uint32_t pk[32];
      const uint64_t c2 = pack2(c, c), nmc2 = pack2(-mc, -mc);
      uint64_t sum2a = pack2(0.f, 0.f), sum2b = pack2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const uint64_t ya = ffma2(pack2u(sr[i], sr[i + 1]), c2, nmc2);
        const uint64_t yb = ffma2(pack2u(sr[i + 2], sr[i + 3]), c2, nmc2);
        float a0, a1;
        unpack2(ya, a0, a1);
        const uint64_t pa = pack2(fast_exp2(a0), fast_exp2(a1));      // MUFU pair (exp2(-inf) = 0 handles the mask)
        const uint64_t pb = poly_exp2x2(yb);                          // FMA-pipe pair (clamped at 2^-126)
        sum2a = fadd2(sum2a, pa);
        sum2b = fadd2(sum2b, pb);
        pk[i / 2] = cvt_bf16x2(pa);
        pk[i / 2 + 1] = cvt_bf16x2(pb);
      }
      {
        float s0, s1, s2, s3;
        unpack2(sum2a, s0, s1);
        unpack2(sum2b, s2, s3);
        l = l * alpha + ((s0 + s1) + (s2 + s3));
      }
      m = mx;
      ATT_STAMP(4);
      if (j > 0) {
        mbar_wait(o_done, (tg - 1) & 1);         // PV_{j-1} retired: O is stable, P may be overwritten
        tc_fence_after_sync();
        // rescale this thread's 32 columns of the running output - skipped when no row of the warp moved its reference max
        if (!__all_sync(0xffffffffu, alpha == 1.f)) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tO + lane_off + half * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_32x32b_x32(tO + lane_off + half * 32, r);
        }
      }
      ATT_STAMP(5);
      // P -> tensor memory: row == lane, this thread's 64 keys = 32 packed columns (the MMA reads A from TMEM)
      tmem_st_32x32b_x32(tP + lane_off + half * 32, pk);
      tmem_st_wait();
      ATT_STAMP(6);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      ATT_STAMP(7);
    }
    t0 += nkv;
    // ---- epilogue: total row sum = both halves
    named_bar_sync(2, SOFTMAX_THREADS);          // everyone is done reading the last tile's max exchange
    red[half * 128 + row] = l;
    named_bar_sync(2, SOFTMAX_THREADS);
    l += red[(half ^ 1) * 128 + row];
    mbar_wait(o_done, (t0 - 1) & 1);
    tc_fence_after_sync();
    const float inv_l = 1.f / l;
    uint32_t r[32];
    tmem_ld_32x32b_x32(tO + lane_off + half * 32, r);
    tmem_ld_wait();
    uint8_t* stg = sOut;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[u * 8 + i]) * inv_l;
      *reinterpret_cast<uint4*>(stg + row * 128 + (((half * 4 + u) ^ (row & 7)) * 16)) = pack8(f);
    }
    if (half == 0) p.lse[((size_t)b * p.Hq + h) * p.S + qblk * BQ + row] = (m * c + log2f(l)) * 0.6931471805599453f;
    fence_proxy_async_smem();
    named_bar_sync(1, SOFTMAX_THREADS);
    if (threadIdx.x == 64) {
      tma_store_2d(&tmap_out, stg, col_q, row0);
      tma_store_commit();
      tma_store_wait<0>();
    }
    }  // query tiles of this CTA
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace attn

// qkv: [B*S, (Hq+2*Hkv)*64] bf16 (row stride ld_qkv elements); out: [B*S, Hq*64] bf16; lse: [B, Hq, S] fp32
ODB_EXPORT int odb_attn_fwd(const void* qkv, void* out, void* lse, int B, int S, int Hq, int Hkv, long long ld_qkv,
                            long long ld_out, float softmax_scale, void* dbg, cudaStream_t st) {
  using namespace attn;
  if (S % BQ || Hq % Hkv || ld_qkv % 8 || ld_out % 8) return -1;
  CUtensorMap tq, to;
  int rc;
  const long long T = (long long)B * S;
  if ((rc = make_tmap_2d(&tq, qkv, T, (Hq + 2 * Hkv) * D, ld_qkv * 2, 128, 64, 2))) return rc;
  if ((rc = make_tmap_2d(&to, out, T, Hq * D, ld_out * 2, 128, 64, 2))) return rc;
  Params p{};
  p.B = B; p.S = S; p.Hq = Hq; p.Hkv = Hkv;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.lse = (float*)lse;
  p.dbg = (long long*)dbg;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((S / BQ + 1) / 2, Hq, B);
  launch_pdl(attn_fwd_kernel, grid, dim3(THREADS), SMEM_BYTES, st, tq, to, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
