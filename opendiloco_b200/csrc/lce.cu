// Cross-entropy over a logits tile for the chunked linear-cross-entropy path.
// The LM-head GEMM writes a [Tc, V] bf16 tile; this kernel turns it IN PLACE into d(loss)/d(logits) and
// accumulates the un-normalised loss sum, so the fp32 [T, V] logits tensor of the reference
// (modeling_llama.py:487-491 -> loss_utils.py:45-67) never exists.
//   row r (label y, -100 = ignore):  lse = log sum_j exp(x_j) ;  loss_r = lse - x_y
//   dlogits_j = (softmax_j - [j == y]) * gscale        gscale = loss_scale / n_valid  (read from device memory)
// One CTA per row; the whole row is cached in registers (V <= 256*8*NV) so HBM sees one read and one write.
#include "common.cuh"
#include "packed_math.cuh"

using namespace odb;

// max of the two bf16 halves of each word, packed (HMNMX2.BF16): the row maximum needs no unpacking
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// One CTA per row; the row stays packed (bf16) in registers across the three passes and the arithmetic is packed fp32x2
// (FFMA2 / FADD2 / FMUL2) with branch-free loops (out-of-range vectors are padded with -inf, the label element is read
// and patched by one thread): ~7 instructions per logit instead of ~19, which is what bounded this kernel (32000 logits
// per row, 55 rows per SM per 8192-row chunk), leaving the 16-lane MUFU (one ex2 per logit) and HBM as the limits.
template <int NV, bool kSumSq>
__global__ void __launch_bounds__(256) ce_fwd_bwd_kernel(__nv_bfloat16* __restrict__ logits,
                                                         const long long* __restrict__ labels, int V, long long row_stride,
                                                         const float* __restrict__ gscale_ptr, float* __restrict__ loss_sum,
                                                         float* __restrict__ lse_out, float* __restrict__ sumsq) {
  using namespace packed_math;
  __shared__ float sm[33];
  const int row = blockIdx.x;
  const long long y = labels[row];
  __nv_bfloat16* lrow = logits + (size_t)row * row_stride;
  uint4* xr = reinterpret_cast<uint4*>(lrow);
  const int nvec = V / 8;
  const bool ignore = (y < 0);
  if (ignore && !kSumSq) {
    // masked token: gradient is zero, no loss contribution
    for (int c = threadIdx.x; c < nvec; c += 256) st_na_v4(xr + c, make_uint4(0, 0, 0, 0));
    if (lse_out && threadIdx.x == 0) lse_out[row] = 0.f;
    return;
  }
  const float x_label = (!ignore && threadIdx.x == 0) ? __bfloat162float(lrow[y]) : 0.f;   // read before the row is overwritten
  uint32_t raw[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {                         // all loads of the row in flight at once
    const int c = threadIdx.x + k * 256;
    uint4 v = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);      // -inf padding: exp -> 0
    if (c < nvec) v = ld_nc_v4(xr + c);
    raw[k][0] = v.x; raw[k][1] = v.y; raw[k][2] = v.z; raw[k][3] = v.w;
  }
  uint32_t m2 = 0xff80ff80u;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    m2 = max_bf16x2(max_bf16x2(m2, raw[k][0]), max_bf16x2(raw[k][1], max_bf16x2(raw[k][2], raw[k][3])));
    if constexpr (kSumSq) {
      if (threadIdx.x + k * 256 < nvec) {
        float f[8];
        unpack8(make_uint4(raw[k][0], raw[k][1], raw[k][2], raw[k][3]), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sq += f[j] * f[j];
      }
    }
  }
  float mx = fmaxf(__uint_as_float(m2 << 16), __uint_as_float(m2 & 0xffff0000u));
  mx = block_max(mx, sm);
  // exp() once per element: the fp32 values feed the row sum, their bf16 roundings replace the logits in registers and
  // become the softmax numerators of the last pass (the output is bf16 anyway).
  const float L2E = 1.4426950408889634f;
  const uint64_t l2 = pack2(L2E, L2E), nm2 = pack2(-mx * L2E, -mx * L2E);
  uint64_t s2 = pack2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w = raw[k][i];
      float e0, e1;
      unpack2(ffma2(pack2u(w << 16, w & 0xffff0000u), l2, nm2), e0, e1);
      const uint64_t e = pack2(fast_exp2(e0), fast_exp2(e1));
      s2 = fadd2(s2, e);
      raw[k][i] = cvt_bf16x2(e);
    }
  }
  float s;
  {
    float sa, sb;
    unpack2(s2, sa, sb);
    s = block_sum(sa + sb, sm);
  }
  if constexpr (kSumSq) {
    sq = block_sum(sq, sm);
    if (threadIdx.x == 0) atomicAdd(sumsq, sq);
  }
  const float lse = mx + __logf(s);
  const float gscale = ignore ? 0.f : *gscale_ptr;
  const float inv_s = gscale / s;
  const uint64_t is2 = pack2(inv_s, inv_s);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + k * 256;
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w = raw[k][i];
      o[i] = cvt_bf16x2(fmul2(pack2u(w << 16, w & 0xffff0000u), is2));
    }
    if (c < nvec) st_na_v4(xr + c, make_uint4(o[0], o[1], o[2], o[3]));
  }
  if (!ignore) {
    __syncthreads();                                     // the vector store that covers the label element is ordered first
    if (threadIdx.x == 0) {
      // softmax - onehot at the label, from the same bf16-rounded numerator the vector pass used
      const float e_lab = bf16_round(fast_exp2(fmaf(x_label, L2E, -mx * L2E)));
      lrow[y] = __float2bfloat16(e_lab * inv_s - gscale);
      atomicAdd(loss_sum, lse - x_label);
    }
  }
  if (lse_out && threadIdx.x == 0) lse_out[row] = ignore ? 0.f : lse;
}

// forward only (evaluation): loss sum, no gradient written
template <int NV>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits,
                                                     const long long* __restrict__ labels, int V, long long row_stride,
                                                     float* __restrict__ loss_sum) {
  __shared__ float sm[33];
  const int row = blockIdx.x;
  const long long y = labels[row];
  if (y < 0) return;
  const uint4* xr = reinterpret_cast<const uint4*>(logits + (size_t)row * row_stride);
  const int nvec = V / 8;
  uint4 raw[NV];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < nvec) {
      raw[k] = ld_nc_v4(xr + c);
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
    }
  }
  mx = block_max(mx, sm);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < nvec) {
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += __expf(f[j] - mx);
    }
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) {
    const float xy = __bfloat162float(logits[(size_t)row * row_stride + y]);
    atomicAdd(loss_sum, mx + __logf(s) - xy);
  }
}

#define ODB_DISPATCH_NV(V, ...)                                 \
  do {                                                          \
    const int _n = ceil_div((V), 2048);                         \
    if (_n <= 1) { constexpr int NV = 1; __VA_ARGS__; }         \
    else if (_n <= 2) { constexpr int NV = 2; __VA_ARGS__; }    \
    else if (_n <= 4) { constexpr int NV = 4; __VA_ARGS__; }    \
    else if (_n <= 8) { constexpr int NV = 8; __VA_ARGS__; }    \
    else if (_n <= 16) { constexpr int NV = 16; __VA_ARGS__; }  \
    else if (_n <= 32) { constexpr int NV = 32; __VA_ARGS__; }  \
    else return -2;                                             \
  } while (0)

ODB_EXPORT int odb_ce_fwd_bwd(void* logits, const void* labels, int rows, int V, long long row_stride,
                              const void* gscale_ptr, void* loss_sum, void* lse_out, void* sumsq, cudaStream_t st) {
  if (V % 8 || row_stride % 8) return -1;
  if (rows <= 0) return 0;
  if (sumsq) {
    ODB_DISPATCH_NV(V, (ce_fwd_bwd_kernel<NV, true><<<rows, 256, 0, st>>>((__nv_bfloat16*)logits, (const long long*)labels, V,
                                                                          row_stride, (const float*)gscale_ptr,
                                                                          (float*)loss_sum, (float*)lse_out, (float*)sumsq)));
  } else {
    ODB_DISPATCH_NV(V, (ce_fwd_bwd_kernel<NV, false><<<rows, 256, 0, st>>>((__nv_bfloat16*)logits, (const long long*)labels, V,
                                                                           row_stride, (const float*)gscale_ptr,
                                                                           (float*)loss_sum, (float*)lse_out, (float*)sumsq)));
  }
  ODB_CHECK_LAST();
  return 0;
}

ODB_EXPORT int odb_ce_fwd(const void* logits, const void* labels, int rows, int V, long long row_stride, void* loss_sum,
                          cudaStream_t st) {
  if (V % 8 || row_stride % 8) return -1;
  if (rows <= 0) return 0;
  ODB_DISPATCH_NV(V, (ce_fwd_kernel<NV><<<rows, 256, 0, st>>>((const __nv_bfloat16*)logits, (const long long*)labels, V,
                                                              row_stride, (float*)loss_sum)));
  ODB_CHECK_LAST();
  return 0;
}
