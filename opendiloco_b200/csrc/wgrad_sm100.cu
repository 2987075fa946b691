// Weight-gradient GEMM on tcgen05 (SURVEY.md §2.5 K3/K6 backward):
//
//   dW[N_out, K_out] (fp32) += dY[T, N_out]^T * X[T, K_out]          bf16 operands, reduction over the T tokens
//
// Both operands are "MN-major" for the tensor core (the reduction index t is the slow memory dimension), which the UMMA
// shared-memory descriptor expresses directly - no transposed copies of activations are ever made.  The output is
// ACCUMULATED in global memory with TMA reduce-add (cp.reduce.async.bulk.tensor ... .add, fp32), which gives three
// things at once: gradient accumulation across micro-batches, split-K (each CTA pair owns a slice of T, so the 16..84
// output tiles of a Llama weight still fill 74 CTA pairs), and no read-modify-write through the SMs.
//
// Structure = gemm2_sm100.cu (CTA pair, 256x256 tile, tcgen05.mma.cta_group::2, double-buffered TMEM accumulator, two
// epilogue warp-groups) with MN-major tiles: per k-block (64 tokens) each CTA loads
//     A: dY[t0..t0+64, m0..m0+128]  as two 64x64 boxes   (8 KB each, 128-byte swizzle)
//     B: X [t0..t0+64, n0..n0+128]  as two 64x64 boxes
// A may be split along M into three source tensors (dq | dk | dv of the fused QKV projection).
#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace odb;
using namespace sm100;

namespace wgrad {

constexpr int BM = 128, BN = 256, BK = 64;       // per CTA: 128 rows of the 256-row pair tile; BN is the pair-tile width
constexpr int BOX_BYTES = 64 * 64 * 2;           // one 64(t) x 64(mn) bf16 box
constexpr int A_BYTES = 2 * BOX_BYTES;           // 128 m-values
constexpr int B_BYTES = 2 * BOX_BYTES;           // this CTA's 128 n-values
constexpr int STAGES = 5;
constexpr int EPI_COLS = 32;                     // fp32 columns per staged chunk (128 B rows)
constexpr int EPI_BYTES = BM * EPI_COLS * 4;     // 16 KB
constexpr int NBUF = 4;                          // 2 per epilogue group
constexpr int THREADS = 320;
constexpr int EPI_THREADS = 128;
constexpr uint32_t TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + NBUF * EPI_BYTES + 1024 + 256;

struct Params {
  int num_m, num_n;      // 256x256 output tiles
  int splits;            // split-K factor over the token dimension
  int kb_total;          // number of 64-token blocks
  int m_split1, m_split2;  // tile-row indices where A switches to the 2nd / 3rd source tensor (num_m = no switch)
};

__device__ __forceinline__ void stage_row_f32(uint8_t* buf, int row, const uint32_t (&v)[32]) {
  uint8_t* rbase = buf + row * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int phys = j ^ (row & 7);
    *reinterpret_cast<uint4*>(rbase + phys * 16) = make_uint4(v[j * 4 + 0], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
             const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ CUtensorMap tmap_c, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_epi = smem_b + STAGES * B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + NBUF * EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = p.num_m * p.num_n;
  const int num_work = num_tiles * p.splits;       // work item = (output tile, token slice)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a0);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 16); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // work item -> (token slice, tile): pairs running at the same time work on DIFFERENT output tiles of the SAME token
  // slice, so they share dY / X tiles in L2 and never reduce-add into the same addresses at the same moment
  auto decode = [&](int w, int& m_blk, int& n_blk, int& kb0, int& kb1) {
    const int sp = w / num_tiles, tile = w - sp * num_tiles;
    // n fastest: the (few) X-column tiles that share one dY column block run side by side, so a dY slice is fetched from
    // HBM once even when a token slice has more tiles than the machine has CTA pairs (lm_head: 125 x 4 tiles - with m
    // fastest its 524 MB dlogits chunk was streamed four times)
    n_blk = tile % p.num_n;
    m_blk = tile / p.num_n;
    const int per = (p.kb_total + p.splits - 1) / p.splits;
    kb0 = sp * per;
    kb1 = kb0 + per < p.kb_total ? kb0 + per : p.kb_total;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        int m_blk, n_blk, kb0, kb1;
        decode(w, m_blk, n_blk, kb0, kb1);
        const CUtensorMap* ma = (m_blk < p.m_split1) ? &tmap_a0 : (m_blk < p.m_split2 ? &tmap_a1 : &tmap_a2);
        const int m_base = (m_blk < p.m_split1) ? m_blk : (m_blk < p.m_split2 ? m_blk - p.m_split1 : m_blk - p.m_split2);
        const int m_idx = m_base * (2 * BM) + (int)cta_rank * BM;           // column of dY (inside its source tensor)
        const int n_idx = n_blk * BN + (int)cta_rank * (BN / 2);            // column of X
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_BYTES + B_BYTES));
          uint8_t* sa = smem_a + stage * A_BYTES;
          uint8_t* sb = smem_b + stage * B_BYTES;
          tma_load_2d_2cta(sa, ma, leader_full, m_idx, kb * BK);
          tma_load_2d_2cta(sa + BOX_BYTES, ma, leader_full, m_idx + 64, kb * BK);
          tma_load_2d_2cta(sb, &tmap_b, leader_full, n_idx, kb * BK);
          tma_load_2d_2cta(sb + BOX_BYTES, &tmap_b, leader_full, n_idx + 64, kb * BK);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, /*a MN-major*/ 1, /*b MN-major*/ 1);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int w = cluster_id; w < num_work; w += num_clusters, ++it) {
      int m_blk, n_blk, kb0, kb1;
      decode(w, m_blk, n_blk, kb0, kb1);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          // MN-major, 128-byte swizzle: 64-element MN chunks are BOX_BYTES apart (LBO), 8-token groups 1024 B apart (SBO)
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_BYTES), BOX_BYTES, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_BYTES), BOX_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // 16 tokens further along K = 16 rows of 128 B = 2048 B = 128 descriptor units
            umma_ss_2cta(d_tmem, adesc + 128 * k, bdesc + 128 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2cta(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit_2cta(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int eg = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64 - eg * EPI_THREADS;
    const bool store_thread = (epi_tid == 0);
    const int bar_a = 1 + 2 * eg, bar_b = 2 + 2 * eg;
    uint8_t* my_epi = smem_epi + eg * (NBUF / 2) * EPI_BYTES;
    int it = 0, buf_i = 0;
    for (int w = cluster_id; w < num_work; w += num_clusters, ++it) {
      int m_blk, n_blk, kb0, kb1;
      decode(w, m_blk, n_blk, kb0, kb1);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int out_row = m_blk * (2 * BM) + (int)cta_rank * BM;      // row of dW (global, across the A source tensors)
      const uint32_t leader_tmem_empty = mapa_shared(smem_u32(&tmem_empty[acc]), 0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      constexpr int NCHUNK = BN / EPI_COLS;       // 8 chunks of 32 fp32 columns
#pragma unroll 1
      for (int c = eg; c < NCHUNK; c += 2) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c * EPI_COLS, v);
        tmem_ld_wait();
        if (c >= NCHUNK - 2) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
        }
        uint8_t* buf = my_epi + (buf_i & 1) * EPI_BYTES;
        if (store_thread) tma_store_wait_read<1>();
        named_bar_sync(bar_a, EPI_THREADS);
        stage_row_f32(buf, row, v);
        fence_proxy_async_smem();
        named_bar_sync(bar_b, EPI_THREADS);
        if (store_thread) {
          tma_reduce_add_2d(&tmap_c, buf, n_blk * BN + c * EPI_COLS, out_row);
          tma_store_commit();
        }
        ++buf_i;
      }
    }
    if (store_thread) tma_store_wait<0>();
  }

  tc_fence_before_sync();
  cluster_sync();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

}  // namespace wgrad

// dW[N_out, K_out] (fp32, row stride ldw) += [dY0 | dY1 | dY2][T, N_out]^T * X[T, K_out]
// n1, n2: column counts of the 2nd / 3rd dY tensors (0 = unused; then dY0 has all N_out columns); when used, every
// piece must be a multiple of 256 columns.  T, N_out, K_out need no alignment beyond 8 elements (TMA clips).
ODB_EXPORT int odb_wgrad_bf16(const void* dY0, const void* dY1, const void* dY2, long long ld0, long long ld1, long long ld2,
                              int n0, int n1, int n2, const void* X, long long ldx, void* dW, long long ldw, int T, int K_out,
                              cudaStream_t st) {
  using namespace wgrad;
  const int N_out = n0 + n1 + n2;
  if ((n1 || n2) && (n0 % 256 || n1 % 256 || n2 % 256)) return -1;
  if (ld0 % 8 || ldx % 8 || ldw % 4 || T <= 0) return -1;
  CUtensorMap ta0, ta1, ta2, tb, tc;
  int rc;
  // operands are described to TMA as [T rows, cols] row-major with 64x64 boxes (inner = MN index)
  if ((rc = make_tmap_2d(&ta0, dY0, T, n0, ld0 * 2, 64, 64, 2))) return rc;
  ta1 = ta0; ta2 = ta0;
  if (n1 && (rc = make_tmap_2d(&ta1, dY1, T, n1, ld1 * 2, 64, 64, 2))) return rc;
  if (n2 && (rc = make_tmap_2d(&ta2, dY2, T, n2, ld2 * 2, 64, 64, 2))) return rc;
  if ((rc = make_tmap_2d(&tb, X, T, K_out, ldx * 2, 64, 64, 2))) return rc;
  if ((rc = make_tmap_2d(&tc, dW, N_out, K_out, ldw * 4, BM, EPI_COLS, 4))) return rc;
  Params p{};
  p.num_m = ceil_div(N_out, 2 * BM);
  p.num_n = ceil_div(K_out, BN);
  p.kb_total = ceil_div(T, BK);
  p.m_split1 = n1 ? n0 / 256 : p.num_m;
  p.m_split2 = n2 ? (n0 + n1) / 256 : p.num_m;
  const int tiles = p.num_m * p.num_n;
  const int pairs = sm_count() / 2;
  // split-K so that tiles * splits is close to a whole number of waves, with at least ~32 k-blocks per work item
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 64 && p.kb_total / s >= 16; ++s) {
    const int work = tiles * s;
    const int waves = ceil_div(work, pairs);
    const double eff = (double)work / (double)(waves * pairs);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  {   // no empty slices: every work item must issue at least one MMA (the epilogue waits for its commit)
    const int per = ceil_div(p.kb_total, best);
    best = ceil_div(p.kb_total, per);
  }
  p.splits = best;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int work = tiles * p.splits;
  const int grid = 2 * (work < pairs ? work : pairs);
  launch_pdl(wgrad_kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, st, ta0, ta1, ta2, tb, tc, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
