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
//   kRoPE    : C is the packed qkv buffer; 64-column chunks are heads; chunks below `rope_cols` are rotated with the
//              HF rotate-half convention using fp32 cos/sin tables (position = row % S)
#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace odb;
using namespace sm100;

namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2;       // 16 KB
constexpr int B_BYTES = BN * BK * 2;       // 32 KB
constexpr int EPI_CHUNK = 64;              // output columns per epilogue store (128 B of bf16)
constexpr int EPI_BYTES = BM * EPI_CHUNK * 2;  // 16 KB
constexpr int THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr uint32_t TMEM_COLS = 512;

enum Epi { kStore = 0, kSwiGLU = 1, kRoPE = 2 };

template <int EPI> struct Cfg { static constexpr int STAGES = 4, NBUF = 2; };
template <> struct Cfg<kSwiGLU> { static constexpr int STAGES = 3, NBUF = 3; };

template <int EPI>
constexpr int smem_bytes() { return Cfg<EPI>::STAGES * (A_BYTES + B_BYTES) + Cfg<EPI>::NBUF * EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/; }

struct Params {
  int M, N, K;
  int num_m, num_n;          // tile counts
  int I;                     // kSwiGLU: intermediate size (row offset of the up projection inside B, column offset in C)
  int S, rope_cols;          // kRoPE: sequence length, number of leading columns that get rotated
  const float* cos_t;        // [S, 32] fp32 (head_dim 64)
  const float* sin_t;
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

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux, Params p) {
  pdl_launch_dependents();
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m * p.num_n;
  const int num_k = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_idx = (tile % p.num_m) * BM;
        const int n_blk = tile / p.num_m;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
          tma_load_2d(smem_a + stage * A_BYTES, &tmap_a, &full_bar[stage], kb * BK, m_idx);
          if constexpr (EPI == kSwiGLU) {
            const int n0 = n_blk * (BN / 2);
            tma_load_2d(smem_b + stage * B_BYTES, &tmap_b, &full_bar[stage], kb * BK, n0);
            tma_load_2d(smem_b + stage * B_BYTES + B_BYTES / 2, &tmap_b, &full_bar[stage], kb * BK, p.I + n0);
          } else {
            tma_load_2d(smem_b + stage * B_BYTES, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in descriptor units
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;              // row inside the 128-row tile
    const int epi_tid = threadIdx.x - 64;       // 0..127
    const bool store_thread = (epi_tid == 0);
    int it = 0;
    int buf_i = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_idx = (tile % p.num_m) * BM;
      const int n_blk = tile / p.num_m;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);

      if constexpr (EPI == kSwiGLU) {
        const int n0 = n_blk * (BN / 2);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
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
          if (c == 1) {   // accumulator fully drained into registers: hand it back to the MMA warp
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          if (store_thread) tma_store_wait_read<0>();
          named_bar_sync(1, EPI_THREADS);
          uint8_t* bg = smem_epi;
          uint8_t* bu = smem_epi + EPI_BYTES;
          uint8_t* ba = smem_epi + 2 * EPI_BYTES;
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
          named_bar_sync(2, EPI_THREADS);
          if (store_thread) {
            tma_store_2d(&tmap_c, bg, n0 + c * 64, m_idx);
            tma_store_2d(&tmap_c, bu, p.I + n0 + c * 64, m_idx);
            tma_store_2d(&tmap_aux, ba, n0 + c * 64, m_idx);
            tma_store_commit();
          }
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
        for (int c = 0; c < BN / EPI_CHUNK; ++c) {
          float v[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(t_row + c * 64, r0);
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
          }
          if (c == BN / EPI_CHUNK - 1) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
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
          uint8_t* buf = smem_epi + (buf_i & 1) * EPI_BYTES;
          if (store_thread) tma_store_wait_read<1>();      // the store that last read this buffer has drained
          named_bar_sync(1, EPI_THREADS);
          stage_row_bf16(buf, row, v);
          fence_proxy_async_smem();
          named_bar_sync(2, EPI_THREADS);
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
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int EPI>
static int launch(const void* A, const void* B, void* C, void* aux, int M, int N, int K, long long lda, long long ldb,
                  long long ldc, long long ldaux, int I, int S, int rope_cols, const float* cos_t, const float* sin_t,
                  cudaStream_t st) {
  CUtensorMap ta, tb, tc, tx;
  int rc;
  const int n_out = (EPI == kSwiGLU) ? 2 * I : N;
  if ((rc = make_tmap_2d(&ta, A, M, K, lda * 2, BM, BK, 2))) return rc;
  if ((rc = make_tmap_2d(&tb, B, (EPI == kSwiGLU) ? 2 * I : N, K, ldb * 2, (EPI == kSwiGLU) ? BN / 2 : BN, BK, 2))) return rc;
  if ((rc = make_tmap_2d(&tc, C, M, n_out, ldc * 2, BM, EPI_CHUNK, 2))) return rc;
  if (EPI == kSwiGLU) {
    if ((rc = make_tmap_2d(&tx, aux, M, I, ldaux * 2, BM, EPI_CHUNK, 2))) return rc;
  } else {
    tx = tc;
  }
  Params p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m = ceil_div(M, BM);
  p.num_n = (EPI == kSwiGLU) ? ceil_div(I, BN / 2) : ceil_div(N, BN);
  p.I = I; p.S = S > 0 ? S : 1; p.rope_cols = rope_cols; p.cos_t = cos_t; p.sin_t = sin_t;
  static bool attr_set = false;
  constexpr int smem = smem_bytes<EPI>();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = p.num_m * p.num_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  launch_pdl(gemm_bf16_tn_kernel<EPI>, dim3(grid), dim3(THREADS), smem, st, ta, tb, tc, tx, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace gemm

// C[M,N] (bf16, row stride ldc) = A[M,K] (row stride lda) * B[N,K]^T (row stride ldb); strides in elements.
ODB_EXPORT int odb_gemm_bf16_tn(const void* A, const void* B, void* C, int M, int N, int K, long long lda, long long ldb,
                                long long ldc, cudaStream_t st) {
  if (K % 8 || lda % 8 || ldb % 8 || ldc % 8) return -1;
  return gemm::launch<gemm::kStore>(A, B, C, nullptr, M, N, K, lda, ldb, ldc, 0, 0, 0, 0, nullptr, nullptr, st);
}

// gu[M,2I] = x[M,K] * Wgu[2I,K]^T and act[M,I] = silu(gu[:, :I]) * gu[:, I:]   (I % 64 == 0)
ODB_EXPORT int odb_gemm_swiglu(const void* X, const void* Wgu, void* gu, void* act, int M, int I, int K, cudaStream_t st) {
  if (K % 8 || I % 64) return -1;
  return gemm::launch<gemm::kSwiGLU>(X, Wgu, gu, act, M, 2 * I, K, K, K, 2 * I, I, I, 0, 0, nullptr, nullptr, st);
}

// qkv[M,N] = x[M,K] * Wqkv[N,K]^T with RoPE applied to the first rope_cols columns (head_dim 64, position = row % S)
ODB_EXPORT int odb_gemm_qkv_rope(const void* X, const void* W, void* qkv, int M, int N, int K, int S, int rope_cols,
                                 const void* cos_t, const void* sin_t, cudaStream_t st) {
  if (K % 8 || N % 64 || rope_cols % 64) return -1;
  return gemm::launch<gemm::kRoPE>(X, W, qkv, nullptr, M, N, K, K, K, N, 0, 0, S, rope_cols, (const float*)cos_t,
                                   (const float*)sin_t, st);
}
