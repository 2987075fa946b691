// 2-CTA (cta_group::2) version of the tcgen05 GEMM in gemm_sm100.cu: a CTA PAIR on one TPC computes a 256x256 tile.
// Each CTA stages only its own 128 rows of A and its own 128-row HALF of B per k-block (32 KB instead of 48 KB), the
// leader CTA issues tcgen05.mma.cta_group::2 (M=256) which reads both halves, and every CTA drains its own 128x256
// accumulator.  Halving the B traffic per SM is what lifts the kernel off the L2->SM feed limit of the 1-CTA version.
//
// (original header follows)
// tcgen05 / TMEM / TMA GEMM for sm_100a with fused Llama epilogues.
//
//   C[M,N] = A[M,K] * B[N,K]^T      bf16 operands (both K-major), fp32 accumulation in TMEM, bf16 output
//
// Persistent, warp-specialised kernel: one CTA per SM, 192 threads
//   warp 0      TMA producer: A tile 128x64, B tile 256x64 per stage (128-byte swizzle), mbarrier full/empty ring
//   warp 1      MMA issuer: one elected lane issues 4 x tcgen05.mma (M128 N256 K16) per stage into one of TWO TMEM
//               accumulators (2 x 256 columns = all 512), commits free the smem stage / publish the accumulator
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32b x 32 cols) -> registers -> fused math -> swizzled smem tile ->
//               TMA store; overlaps with the MMAs of the next tile (double-buffered accumulator)
//
// Epilogues:
//   kStore   : plain bf16 store
//   kSwiGLU  : B rows are [gate rows n0..n0+127 | up rows I+n0..I+n0+127]; writes gate|up (needed by backward) AND
//              act = silu(gate) * up, so the SwiGLU kernel and its re-read of the 2I-wide tensor disappear
//   kSwiGLUBwd: the GEMM result is d(act) of the down projection; the epilogue TMA-loads the matching gate / up tiles of
//              C = gu [M, 2I], turns them IN PLACE into d(gate) / d(up) and stores them back - d(act) never reaches memory
//   kRoPE    : C is the packed qkv buffer; 64-column chunks are heads; chunks below `rope_cols` are rotated with the
//              HF rotate-half convention using fp32 cos/sin tables (position = row % S)
//   kLCEFwd  : fused linear-cross-entropy forward.  The accumulator holds a tile of LM-head logits z;
//              the epilogue turns it into shifted exponentials e = exp(z - c_row) (c_row = the row's label logit, so
//              sum_v e >= 1 and loss_row = log sum_v e), accumulates the row sums into per-tile partial planes and
//              (training only) stores bf16(e).  Neither logits nor a log-softmax pass ever exist; in evaluation mode
//              (store = 0) nothing of size [T, V] is written at all.
//   kLCEdX   : dgrad of the LM head on the shifted exponentials: out = rowscale * acc - g * W[label]  (softmax - onehot
//              is never formed: the per-row normaliser is applied to the fp32 accumulator, the one-hot term is a gather)
//
// B_MN = true: the B operand is stored [K, N] row-major (N contiguous, "MN-major"): C = A * B.  This is the dgrad form
// dX = dY * W with W in its forward [out, in] layout, so no transposed weight copy is needed (the 64 x 64 boxes and the
// LBO / SBO descriptor fields are the ones the weight-gradient kernel uses for both of its operands).
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace odb;
using namespace sm100;

namespace gemm2 {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2;       // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2; // 16 KB: this CTA's half of the 256-row B tile
constexpr int EPI_CHUNK = 64;              // output columns per epilogue store (128 B of bf16)
constexpr int EPI_BYTES = BM * EPI_CHUNK * 2;  // 16 KB
constexpr int THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue group 0, warps 6-9 epilogue group 1
constexpr int EPI_THREADS = 128;      // per group
constexpr int EPI_GROUPS = 2;
constexpr int GROUP_M = 8;            // rasterisation: 8 m-blocks x all n-blocks per super-block (L2 reuse of A and B); measured
                                      // 4 / 8 / 16 / 32 on the 150M layer shapes: 979 / 979 / 986 / 1018 us per layer (profiles/r2_gemm_raster.txt)
constexpr uint32_t TMEM_COLS = 512;

enum Epi { kStore = 0, kSwiGLU = 1, kRoPE = 2, kSwiGLUBwd = 3, kLCEFwd = 4, kLCEdX = 5 };
constexpr int BOX_BYTES = 64 * 64 * 2;    // MN-major B: one 64(k) x 64(n) box

template <int EPI> struct Cfg { static constexpr int STAGES = 5, NBUF = 4; };     // NBUF: staging tiles (2 per epilogue group)
template <> struct Cfg<kSwiGLU> { static constexpr int STAGES = 4, NBUF = 6; };   // 3 per epilogue group
template <> struct Cfg<kSwiGLUBwd> { static constexpr int STAGES = 3, NBUF = 8; };   // per group: 2 chunks x (gate, up) in/out tiles

template <int EPI>
constexpr int smem_bytes() { return Cfg<EPI>::STAGES * (A_BYTES + B_BYTES) + Cfg<EPI>::NBUF * EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/; }

// group_m > 0: groups of group_m m-blocks x all n-blocks, m fastest inside a group (A and B tiles of a group stay in L2).
// group_m == 0: n fastest over the whole problem - for a few n-blocks and a very deep K (the LM-head dX: 4 n-blocks,
// K = vocabulary) the pairs that share an A row-block then run side by side and the A operand is fetched from HBM once.
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
  if (group_m == 0) {
    m_blk = tile / num_n;
    n_blk = tile - m_blk * num_n;
    return;
  }
  const int group_size = group_m * num_n;
  const int group_id = tile / group_size;
  const int first_m = group_id * group_m;
  const int gsz = (num_m - first_m) < group_m ? (num_m - first_m) : group_m;
  const int r = tile - group_id * group_size;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
}

struct Params {
  int M, N, K;
  int num_m, num_n;          // tile counts
  int group_m;               // rasterisation (see tile_coords)
  int I;                     // kSwiGLU: intermediate size (row offset of the up projection inside B, column offset in C)
  int S, rope_cols;          // kRoPE: sequence length, number of leading columns that get rotated
  const float* cos_t;        // [S, 32] fp32 (head_dim 64)
  const float* sin_t;
  int ks1, ks2;              // A operand split along K into up to three tensors: k-blocks [0,ks1) from A, [ks1,ks2) from A1,
                             // [ks2,..) from A2 (0/0 = single tensor).  Lets dq|dk|dv feed one dgrad GEMM without packing.
  // ---- fused linear-cross-entropy
  const float* row_shift;    // kLCEFwd: c_row [M] (label logit; 0 for ignored rows)
  float* partials;           // kLCEFwd: [2 * num_n (+ 2 * num_n when want_sumsq)][M] partial row sums of exp(z - c)
  int want_sumsq;            // kLCEFwd: also accumulate sum z^2 (lm_head activation norm) into the second half of partials
  int store;                 // kLCEFwd: 1 = write bf16(e) to C (training) ; 0 = statistics only (evaluation)
  const float* rowscale;     // kLCEdX: s_row [M] = gscale / sum_v e  (0 for ignored rows)
  const long long* labels;   // kLCEdX: [M], < 0 = ignored
  const float* gscale;       // kLCEdX: device scalar loss_scale / n_valid
  const __nv_bfloat16* wlab; // kLCEdX: the LM-head weight [V, N] (row stride ldwlab) for the one-hot gather
  long long ldwlab;
};

// write one 64-column chunk of this thread's row into the swizzled staging tile (row r, 8 x 16-byte chunks)
__device__ __forceinline__ void stage_row_bf16(uint8_t* buf, int row, const float (&v)[64]) {
  uint8_t* rbase = buf + row * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float f[8] = {v[j * 8 + 0], v[j * 8 + 1], v[j * 8 + 2], v[j * 8 + 3], v[j * 8 + 4], v[j * 8 + 5], v[j * 8 + 6], v[j * 8 + 7]};
    const uint4 p = pack8(f);
    const int phys = j ^ (row & 7);
    *reinterpret_cast<uint4*>(rbase + phys * 16) = p;
  }
}

template <int EPI, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux,
                    const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_a2, Params p) {
  constexpr int STAGES = Cfg<EPI>::STAGES;
  constexpr int NBUF = Cfg<EPI>::NBUF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_epi = smem_b + STAGES * B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + NBUF * EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* gu_full = tmem_empty + 2;                    // kSwiGLUBwd: [group][chunk] gate/up tiles landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gu_full + 4);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();          // 0 = leader of the pair
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = p.num_m * p.num_n;               // tiles of 256 x 256 (one per CTA pair)
  const int num_k = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 16); }   // 8 epilogue warps x 2 CTAs
    for (int i = 0; i < 4; ++i) mbar_init(&gu_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync();          // barrier inits + TMEM allocation visible to the peer CTA
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();              // everything above touched only this CTA's shared / tensor memory

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_blk, n_blk;
        tile_coords(tile, p.num_m, p.num_n, p.group_m, m_blk, n_blk);
        const int m_idx = m_blk * (2 * BM) + (int)cta_rank * BM;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);   // the pair's "full" barrier lives in CTA 0
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_BYTES + B_BYTES));
          if (p.ks1 == 0 || kb < p.ks1) tma_load_2d_2cta(smem_a + stage * A_BYTES, &tmap_a, leader_full, kb * BK, m_idx);
          else if (kb < p.ks2) tma_load_2d_2cta(smem_a + stage * A_BYTES, &tmap_a1, leader_full, (kb - p.ks1) * BK, m_idx);
          else tma_load_2d_2cta(smem_a + stage * A_BYTES, &tmap_a2, leader_full, (kb - p.ks2) * BK, m_idx);
          int b_row;
          if constexpr (EPI == kSwiGLU) b_row = (cta_rank == 0) ? n_blk * (BN / 2) : p.I + n_blk * (BN / 2);   // gate half | up half
          else b_row = n_blk * BN + (int)cta_rank * (BN / 2);
          if constexpr (B_MN) {      // B is [K, N] row-major: two 64(k) x 64(n) boxes cover this CTA's 128 n-values
            tma_load_2d_2cta(smem_b + stage * B_BYTES, &tmap_b, leader_full, b_row, kb * BK);
            tma_load_2d_2cta(smem_b + stage * B_BYTES + BOX_BYTES, &tmap_b, leader_full, b_row + 64, kb * BK);
          } else {
            tma_load_2d_2cta(smem_b + stage * B_BYTES, &tmap_b, leader_full, kb * BK, b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_BYTES), 16, 1024);
          // MN-major B: 64-element n-chunks are BOX_BYTES apart (LBO), 8-row k-groups 1024 B apart (SBO)
          const uint64_t bdesc = B_MN ? make_smem_desc_sw128(smem_u32(smem_b + stage * B_BYTES), BOX_BYTES, 1024)
                                      : make_smem_desc_sw128(smem_u32(smem_b + stage * B_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in descriptor units
            // MN-major: 16 k further = 16 rows of 128 B = 2048 B = +128 units
            umma_ss_2cta(d_tmem, adesc + 2 * k, bdesc + (B_MN ? 128 : 2) * k, idesc, (kb | k) != 0);
          }
          umma_commit_2cta(&empty_bar[stage]);                       // frees this stage in BOTH CTAs
          if (kb == num_k - 1) umma_commit_2cta(&tmem_full[acc]);    // accumulator ready in BOTH CTAs
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ epilogue (warps 2..9, two groups of 4)
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;              // row inside the 128-row tile
    const int eg = (warp - 2) >> 2;             // epilogue group: handles the 64-column chunks c with c % 2 == eg
    const int epi_tid = threadIdx.x - 64 - eg * EPI_THREADS;       // 0..127 inside the group
    const bool store_thread = (epi_tid == 0);
    const int bar_a = 1 + 2 * eg, bar_b = 2 + 2 * eg;               // named barriers private to the group
    uint8_t* my_epi = smem_epi + eg * (NBUF / EPI_GROUPS) * EPI_BYTES;
    int it = 0;
    int buf_i = 0;
    uint32_t gu_phase = 0;                        // kSwiGLUBwd: phase bit per chunk slot (a ragged tile skips a slot)
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int m_blk, n_blk;
      tile_coords(tile, p.num_m, p.num_n, p.group_m, m_blk, n_blk);
      const int m_idx = m_blk * (2 * BM) + (int)cta_rank * BM;
      const uint32_t leader_tmem_empty = mapa_shared(smem_u32(&tmem_empty[acc]), 0);
      if constexpr (EPI == kSwiGLUBwd) {
        // fetch this group's gate / up tiles while the accumulator is still being computed.  The staging tiles are the
        // ones the previous tile's stores read from, so those must have drained first.
        if (store_thread) {
          tma_store_wait_read<0>();
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int col = n_blk * BN + (eg + 2 * k) * 64;
            if (col >= p.I) continue;                    // ragged last tile (I % 256 != 0): chunk lies beyond the gate block
            uint64_t* bar = &gu_full[eg * 2 + k];
            mbar_arrive_expect_tx(bar, 2 * EPI_BYTES);
            tma_load_2d(my_epi + (2 * k) * EPI_BYTES, &tmap_c, bar, col, m_idx);
            tma_load_2d(my_epi + (2 * k + 1) * EPI_BYTES, &tmap_c, bar, p.I + col, m_idx);
          }
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);

      if constexpr (EPI == kSwiGLU) {
        const int n0 = n_blk * (BN / 2);
        {
          const int c = eg;
          float g[64], u[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(t_row + c * 64, r0);
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { g[j] = __uint_as_float(r0[j]); g[32 + j] = __uint_as_float(r1[j]); }
            tmem_ld_32x32b_x32(t_row + 128 + c * 64, r0);
            tmem_ld_32x32b_x32(t_row + 128 + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { u[j] = __uint_as_float(r0[j]); u[32 + j] = __uint_as_float(r1[j]); }
          }
          {   // this warp's share of the accumulator is in registers: hand it back to the MMA warp
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
          }
          if (store_thread) tma_store_wait_read<0>();
          named_bar_sync(bar_a, EPI_THREADS);
          uint8_t* bg = my_epi;
          uint8_t* bu = my_epi + EPI_BYTES;
          uint8_t* ba = my_epi + 2 * EPI_BYTES;
          stage_row_bf16(bg, row, g);
          stage_row_bf16(bu, row, u);
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const float gj = bf16_round(g[j]);     // the activation is computed from the bf16 values the backward will see
            const float uj = bf16_round(u[j]);
            g[j] = gj * fast_sigmoid(gj) * uj;
          }
          stage_row_bf16(ba, row, g);
          fence_proxy_async_smem();
          named_bar_sync(bar_b, EPI_THREADS);
          if (store_thread) {
            tma_store_2d(&tmap_c, bg, n0 + c * 64, m_idx);
            tma_store_2d(&tmap_c, bu, p.I + n0 + c * 64, m_idx);
            tma_store_2d(&tmap_aux, ba, n0 + c * 64, m_idx);
            tma_store_commit();
          }
        }
      } else if constexpr (EPI == kSwiGLUBwd) {
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
          const int c = eg + 2 * k;
          float d[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(t_row + c * 64, r0);
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { d[j] = __uint_as_float(r0[j]); d[32 + j] = __uint_as_float(r1[j]); }
          }
          if (k == 1) {                                  // this warp's last chunk of the tile
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
          }
          if (n_blk * BN + c * 64 >= p.I) continue;      // nothing was loaded for this chunk, nothing to store
          mbar_wait(&gu_full[eg * 2 + k], (gu_phase >> k) & 1);
          gu_phase ^= 1u << k;
          uint8_t* bg = my_epi + (2 * k) * EPI_BYTES + row * 128;
          uint8_t* bu = my_epi + (2 * k + 1) * EPI_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int phys = (j ^ (row & 7)) * 16;
            float g[8], u[8], og[8], ou[8];
            unpack8(*reinterpret_cast<const uint4*>(bg + phys), g);
            unpack8(*reinterpret_cast<const uint4*>(bu + phys), u);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float dj = bf16_round(d[j * 8 + e]);           // the un-fused path rounds d(act) to bf16 first
              const float sg = fast_sigmoid(g[e]);
              og[e] = dj * u[e] * sg * (1.f + g[e] * (1.f - sg));
              ou[e] = dj * g[e] * sg;
            }
            *reinterpret_cast<uint4*>(bg + phys) = pack8(og);
            *reinterpret_cast<uint4*>(bu + phys) = pack8(ou);
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_b, EPI_THREADS);
          if (store_thread) {
            const int col = n_blk * BN + c * 64;
            tma_store_2d(&tmap_c, my_epi + (2 * k) * EPI_BYTES, col, m_idx);
            tma_store_2d(&tmap_c, my_epi + (2 * k + 1) * EPI_BYTES, p.I + col, m_idx);
            tma_store_commit();
          }
        }
      } else if constexpr (EPI == kLCEFwd) {
        // logits tile -> shifted exponentials + partial row sums.  Thread = one token row, this group's 2 x 64 columns.
        constexpr float L2E = 1.4426950408889634f;
        const int grow = m_idx + row;
        const bool row_ok = grow < p.M;
        const float shift = row_ok ? p.row_shift[grow] * L2E : 0.f;
        float psum = 0.f, psq = 0.f;
#pragma unroll 1
        for (int c = eg; c < BN / EPI_CHUNK; c += EPI_GROUPS) {
          float v[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(t_row + c * 64, r0);
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
          }
          if (c >= BN / EPI_CHUNK - EPI_GROUPS) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
          }
          const int col0 = n_blk * BN + c * 64;
          const int ncols = p.N - col0;                     // < 64 only in a ragged last tile; <= 0: chunk is outside
          if (p.want_sumsq) {
#pragma unroll
            for (int j = 0; j < 64; ++j) psq += (j < ncols) ? v[j] * v[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            // clamp: a row whose label is > 69 nats below its best logit would overflow the fp32 row sum otherwise
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(fmaf(v[j], L2E, -shift), 100.f)));
            v[j] = e;
          }
          if (ncols < 64) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = (j < ncols) ? v[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 64; ++j) psum += v[j];
          if (p.store && ncols > 0) {
            uint8_t* buf = my_epi + (buf_i & 1) * EPI_BYTES;
            if (store_thread) tma_store_wait_read<1>();
            named_bar_sync(bar_a, EPI_THREADS);
            stage_row_bf16(buf, row, v);
            fence_proxy_async_smem();
            named_bar_sync(bar_b, EPI_THREADS);
            if (store_thread) {
              tma_store_2d(&tmap_c, buf, col0, m_idx);
              tma_store_commit();
            }
            ++buf_i;
          }
        }
        if (row_ok) {
          const size_t plane = (size_t)(n_blk * EPI_GROUPS + eg);
          p.partials[plane * p.M + grow] = psum;
          if (p.want_sumsq) p.partials[((size_t)p.num_n * EPI_GROUPS + plane) * p.M + grow] = psq;
        }
      } else if constexpr (EPI == kLCEdX) {
        const int grow = m_idx + row;
        float rs = 0.f, g = 0.f;
        const __nv_bfloat16* wl = nullptr;
        if (grow < p.M) {
          rs = p.rowscale[grow];
          const long long y = p.labels[grow];
          if (y >= 0) { g = *p.gscale; wl = p.wlab + (size_t)y * p.ldwlab; }
        }
#pragma unroll 1
        for (int c = eg; c < BN / EPI_CHUNK; c += EPI_GROUPS) {
          const int col0 = n_blk * BN + c * 64;
          uint4 wv[8];
          const bool have_w = (wl != nullptr) && (col0 + 64 <= p.N);
          if (have_w) {
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = ld_nc_v4(wl + col0 + j * 8);      // in flight while the accumulator is read
          }
          float v[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(t_row + c * 64, r0);
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
          }
          if (c >= BN / EPI_CHUNK - EPI_GROUPS) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
          }
          if (have_w) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float w8[8];
              unpack8(wv[j], w8);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[j * 8 + e] = fmaf(rs, v[j * 8 + e], -g * w8[e]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] *= rs;
          }
          uint8_t* buf = my_epi + (buf_i & 1) * EPI_BYTES;
          if (store_thread) tma_store_wait_read<1>();
          named_bar_sync(bar_a, EPI_THREADS);
          stage_row_bf16(buf, row, v);
          fence_proxy_async_smem();
          named_bar_sync(bar_b, EPI_THREADS);
          if (store_thread) {
            tma_store_2d(&tmap_c, buf, col0, m_idx);
            tma_store_commit();
          }
          ++buf_i;
        }
      } else {
        float cs[32], sn[32];
        bool rope_tile = false;
        if constexpr (EPI == kRoPE) {
          rope_tile = (n_blk * BN) < p.rope_cols;
          if (rope_tile) {
            const int grow = m_idx + row;
            const int pos = (grow < p.M ? grow : 0) % p.S;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 c4 = ld_f4(p.cos_t + (size_t)pos * 32 + j * 4);
              const float4 s4 = ld_f4(p.sin_t + (size_t)pos * 32 + j * 4);
              cs[j * 4 + 0] = c4.x; cs[j * 4 + 1] = c4.y; cs[j * 4 + 2] = c4.z; cs[j * 4 + 3] = c4.w;
              sn[j * 4 + 0] = s4.x; sn[j * 4 + 1] = s4.y; sn[j * 4 + 2] = s4.z; sn[j * 4 + 3] = s4.w;
            }
          }
        }
#pragma unroll 1
        for (int c = eg; c < BN / EPI_CHUNK; c += EPI_GROUPS) {
          float v[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(t_row + c * 64, r0);
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
          }
          if (c >= BN / EPI_CHUNK - EPI_GROUPS) {      // this warp's last chunk of the tile
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
          }
          if constexpr (EPI == kRoPE) {
            if (rope_tile && (n_blk * BN + c * 64) < p.rope_cols) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float a = bf16_round(v[j]), b = bf16_round(v[j + 32]);   // match the unfused path: rotate the bf16 projection
                v[j] = a * cs[j] - b * sn[j];
                v[j + 32] = b * cs[j] + a * sn[j];
              }
            }
          }
          uint8_t* buf = my_epi + (buf_i & 1) * EPI_BYTES;
          if (store_thread) tma_store_wait_read<1>();      // the store that last read this buffer has drained
          named_bar_sync(bar_a, EPI_THREADS);
          stage_row_bf16(buf, row, v);
          fence_proxy_async_smem();
          named_bar_sync(bar_b, EPI_THREADS);
          if (store_thread) {
            tma_store_2d(&tmap_c, buf, n_blk * BN + c * 64, m_idx);
            tma_store_commit();
          }
          ++buf_i;
        }
      }
    }
    if (store_thread) tma_store_wait<0>();
  }

  tc_fence_before_sync();
  cluster_sync();          // nobody leaves (or frees TMEM) while the peer may still touch our smem / barriers / TMEM
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

struct ASplit {
  const void* A1 = nullptr; const void* A2 = nullptr;
  long long lda1 = 0, lda2 = 0;
  int K0 = 0, K1 = 0, K2 = 0;      // K extents of the three A tensors (multiples of 64); K0 == 0 => no split
};

struct LceArgs {
  const float* row_shift = nullptr; float* partials = nullptr; int want_sumsq = 0; int store = 1;
  const float* rowscale = nullptr; const long long* labels = nullptr; const float* gscale = nullptr;
  const void* wlab = nullptr; long long ldwlab = 0;
};

template <int EPI, bool B_MN = false>
static int launch(const void* A, const void* B, void* C, void* aux, int M, int N, int K, long long lda, long long ldb,
                  long long ldc, long long ldaux, int I, int S, int rope_cols, const float* cos_t, const float* sin_t,
                  cudaStream_t st, const ASplit& sp = ASplit(), const LceArgs& lce = LceArgs()) {
  CUtensorMap ta, tb, tc, tx, ta1, ta2;
  int rc;
  const int n_out = (EPI == kSwiGLU || EPI == kSwiGLUBwd) ? 2 * I : N;
  if ((rc = make_tmap_2d(&ta, A, M, sp.K0 ? sp.K0 : K, lda * 2, BM, BK, 2))) return rc;
  ta1 = ta; ta2 = ta;
  if (sp.K0) {
    if ((rc = make_tmap_2d(&ta1, sp.A1, M, sp.K1, sp.lda1 * 2, BM, BK, 2))) return rc;
    if ((rc = make_tmap_2d(&ta2, sp.A2, M, sp.K2, sp.lda2 * 2, BM, BK, 2))) return rc;
  }
  if (B_MN) {
    if ((rc = make_tmap_2d(&tb, B, K, N, ldb * 2, 64, 64, 2))) return rc;            // [K rows, N cols], 64 x 64 boxes
  } else {
    if ((rc = make_tmap_2d(&tb, B, (EPI == kSwiGLU) ? 2 * I : N, K, ldb * 2, BN / 2, BK, 2))) return rc;
  }
  if (EPI == kLCEFwd && !lce.store) tc = ta;       // statistics only: nothing is stored, any valid map will do
  else if ((rc = make_tmap_2d(&tc, C, M, n_out, ldc * 2, BM, EPI_CHUNK, 2))) return rc;
  if (EPI == kSwiGLU) {
    if ((rc = make_tmap_2d(&tx, aux, M, I, ldaux * 2, BM, EPI_CHUNK, 2))) return rc;
  } else {
    tx = tc;
  }
  Params p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m = ceil_div(M, 2 * BM);
  p.num_n = (EPI == kSwiGLU) ? ceil_div(I, BN / 2) : ceil_div(N, BN);
  {
    static int env_group = -2, env_dx = -2;
    if (env_group == -2) { const char* e = getenv("ODB_GEMM_GROUP_M"); env_group = e ? atoi(e) : -1; }
    if (env_dx == -2) { const char* e = getenv("ODB_LCE_DX_GROUP_M"); env_dx = e ? atoi(e) : -1; }
    p.group_m = env_group >= 0 ? env_group : GROUP_M;
    if (EPI == kLCEdX) p.group_m = env_dx >= 0 ? env_dx : 0;      // n fastest: 1491 -> 1415 us at T = 32768, V = 32000
  }
  p.I = I; p.S = S > 0 ? S : 1; p.rope_cols = rope_cols; p.cos_t = cos_t; p.sin_t = sin_t;
  p.ks1 = sp.K0 ? sp.K0 / BK : 0;
  p.ks2 = sp.K0 ? (sp.K0 + sp.K1) / BK : 0;
  p.row_shift = lce.row_shift; p.partials = lce.partials; p.want_sumsq = lce.want_sumsq; p.store = lce.store;
  p.rowscale = lce.rowscale; p.labels = lce.labels; p.gscale = lce.gscale;
  p.wlab = (const __nv_bfloat16*)lce.wlab; p.ldwlab = lce.ldwlab;
  static bool attr_set = false;
  constexpr int smem = smem_bytes<EPI>();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_tn_kernel<EPI, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = p.num_m * p.num_n;
  const int max_clusters = sm_count() / 2;
  const int grid = 2 * (tiles < max_clusters ? tiles : max_clusters);
  launch_pdl(gemm2_bf16_tn_kernel<EPI, B_MN>, dim3(grid), dim3(THREADS), smem, st, ta, tb, tc, tx, ta1, ta2, p);    // cluster dims are compiled in (__cluster_dims__)
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace gemm2

// C[M,N] (bf16, row stride ldc) = A[M,K] (row stride lda) * B[N,K]^T (row stride ldb); strides in elements.
ODB_EXPORT int odb_gemm2_bf16_tn(const void* A, const void* B, void* C, int M, int N, int K, long long lda, long long ldb,
                                long long ldc, cudaStream_t st) {
  if (K % 8 || lda % 8 || ldb % 8 || ldc % 8) return -1;
  return gemm2::launch<gemm2::kStore>(A, B, C, nullptr, M, N, K, lda, ldb, ldc, 0, 0, 0, 0, nullptr, nullptr, st);
}

// gu[M,2I] = x[M,K] * Wgu[2I,K]^T and act[M,I] = silu(gu[:, :I]) * gu[:, I:]   (I % 64 == 0)
ODB_EXPORT int odb_gemm2_swiglu(const void* X, const void* Wgu, void* gu, void* act, int M, int I, int K, cudaStream_t st) {
  if (K % 8 || I % 64) return -1;
  return gemm2::launch<gemm2::kSwiGLU>(X, Wgu, gu, act, M, 2 * I, K, K, K, 2 * I, I, I, 0, 0, nullptr, nullptr, st);
}

// Backward of the SwiGLU MLP's middle: gu[M,2I] (gate|up, bf16) is replaced IN PLACE by d(gate)|d(up), where
// d(act)[M,I] = dY[M,K] * WdT[I,K]^T is formed in tensor memory only (I % 64 == 0).
ODB_EXPORT int odb_gemm2_swiglu_bwd(const void* dY, const void* WdT, void* gu, int M, int I, int K, long long lda, long long ldb,
                                    cudaStream_t st) {
  if (K % 8 || I % 64 || lda % 8 || ldb % 8) return -1;
  return gemm2::launch<gemm2::kSwiGLUBwd>(dY, WdT, gu, nullptr, M, I, K, lda, ldb, 2 * I, 0, I, 0, 0, nullptr, nullptr, st);
}

// qkv[M,N] = x[M,K] * Wqkv[N,K]^T with RoPE applied to the first rope_cols columns (head_dim 64, position = row % S)
ODB_EXPORT int odb_gemm2_qkv_rope(const void* X, const void* W, void* qkv, int M, int N, int K, int S, int rope_cols,
                                 const void* cos_t, const void* sin_t, cudaStream_t st) {
  if (K % 8 || N % 64 || rope_cols % 64) return -1;
  return gemm2::launch<gemm2::kRoPE>(X, W, qkv, nullptr, M, N, K, K, K, N, 0, 0, S, rope_cols, (const float*)cos_t,
                                   (const float*)sin_t, st);
}

// C[M,N] = [A0 | A1 | A2][M, K0+K1+K2] * B[N, K0+K1+K2]^T without materialising the concatenation (K segments % 64 == 0)
ODB_EXPORT int odb_gemm2_bf16_tn_a3(const void* A0, const void* A1, const void* A2, long long lda0, long long lda1, long long lda2,
                                    int K0, int K1, int K2, const void* B, void* C, int M, int N, long long ldb, long long ldc,
                                    cudaStream_t st) {
  if (K0 % 64 || K1 % 64 || K2 % 64 || K0 <= 0 || K1 <= 0 || K2 <= 0) return -1;
  gemm2::ASplit sp;
  sp.A1 = A1; sp.A2 = A2; sp.lda1 = lda1; sp.lda2 = lda2; sp.K0 = K0; sp.K1 = K1; sp.K2 = K2;
  return gemm2::launch<gemm2::kStore>(A0, B, C, nullptr, M, N, K0 + K1 + K2, lda0, ldb, ldc, 0, 0, 0, 0, nullptr, nullptr, st, sp);
}

// ---------------------------------------------------------------------------------------------------- MN-major B (dgrad)
// C[M,N] (bf16) = A[M,K] (row stride lda) * B[K,N] (row stride ldb, N contiguous).  dX = dY * W with W as stored.
ODB_EXPORT int odb_gemm2_bf16_nn(const void* A, const void* B, void* C, int M, int N, int K, long long lda, long long ldb,
                                long long ldc, cudaStream_t st) {
  if (K % 8 || N % 8 || lda % 8 || ldb % 8 || ldc % 8) return -1;
  return gemm2::launch<gemm2::kStore, true>(A, B, C, nullptr, M, N, K, lda, ldb, ldc, 0, 0, 0, 0, nullptr, nullptr, st);
}

// as odb_gemm2_bf16_tn_a3 with B = W[K0+K1+K2, N] in its forward layout (dgrad of the fused QKV projection)
ODB_EXPORT int odb_gemm2_bf16_nn_a3(const void* A0, const void* A1, const void* A2, long long lda0, long long lda1, long long lda2,
                                    int K0, int K1, int K2, const void* B, void* C, int M, int N, long long ldb, long long ldc,
                                    cudaStream_t st) {
  if (K0 % 64 || K1 % 64 || K2 % 64 || K0 <= 0 || K1 <= 0 || K2 <= 0 || N % 8 || ldb % 8) return -1;
  gemm2::ASplit sp;
  sp.A1 = A1; sp.A2 = A2; sp.lda1 = lda1; sp.lda2 = lda2; sp.K0 = K0; sp.K1 = K1; sp.K2 = K2;
  return gemm2::launch<gemm2::kStore, true>(A0, B, C, nullptr, M, N, K0 + K1 + K2, lda0, ldb, ldc, 0, 0, 0, 0, nullptr, nullptr, st, sp);
}

// SwiGLU backward fused into the down-proj dgrad, reading W_down [K = hidden, I] in its forward layout
ODB_EXPORT int odb_gemm2_swiglu_bwd_nn(const void* dY, const void* Wd, void* gu, int M, int I, int K, long long lda, long long ldb,
                                       cudaStream_t st) {
  if (K % 8 || I % 64 || lda % 8 || ldb % 8) return -1;
  return gemm2::launch<gemm2::kSwiGLUBwd, true>(dY, Wd, gu, nullptr, M, I, K, lda, ldb, 2 * I, 0, I, 0, 0, nullptr, nullptr, st);
}

// ---------------------------------------------------------------------------------------------------- fused linear-cross-entropy
// E[M,V] (bf16, optional) = exp(X[M,K] W[V,K]^T - shift[M]) ; partials[(2*ceil(V/256)) (x2 with sumsq)][M] = partial row sums
ODB_EXPORT int odb_lce_fwd(const void* X, const void* W, void* E, int M, int V, int K, long long ldx, long long ldw, long long lde,
                           const void* row_shift, void* partials, int want_sumsq, int store, cudaStream_t st) {
  if (K % 8 || ldx % 8 || ldw % 8 || (store && (lde % 8 || !E))) return -1;
  gemm2::LceArgs a;
  a.row_shift = (const float*)row_shift; a.partials = (float*)partials; a.want_sumsq = want_sumsq; a.store = store;
  return gemm2::launch<gemm2::kLCEFwd, false>(X, W, E, nullptr, M, V, K, ldx, ldw, lde, 0, 0, 0, 0, nullptr, nullptr, st,
                                               gemm2::ASplit(), a);
}
// number of partial planes odb_lce_fwd writes per statistic
ODB_EXPORT int odb_lce_planes(int V) { return 2 * ((V + gemm2::BN - 1) / gemm2::BN); }

// dX[M,N] (bf16) = rowscale[M] * (E[M,V] * W[V,N]) - gscale * [label >= 0] * W[label]     (N % 64 == 0)
ODB_EXPORT int odb_lce_dx(const void* E, const void* W, void* dX, int M, int N, int V, long long lde, long long ldw, long long ldd,
                          const void* rowscale, const void* labels, const void* gscale, cudaStream_t st) {
  if (V % 8 || N % 64 || lde % 8 || ldw % 8 || ldd % 8) return -1;
  gemm2::LceArgs a;
  a.rowscale = (const float*)rowscale; a.labels = (const long long*)labels; a.gscale = (const float*)gscale;
  a.wlab = W; a.ldwlab = ldw;
  return gemm2::launch<gemm2::kLCEdX, true>(E, W, dX, nullptr, M, N, V, lde, ldw, ldd, 0, 0, 0, 0, nullptr, nullptr, st,
                                             gemm2::ASplit(), a);
}
