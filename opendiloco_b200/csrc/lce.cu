// Cross-entropy over a logits tile for the chunked linear-cross-entropy path.
// The LM-head GEMM writes a [Tc, V] bf16 tile; this kernel turns it IN PLACE into d(loss)/d(logits) and
// accumulates the un-normalised loss sum, so the fp32 [T, V] logits tensor of the reference
// (modeling_llama.py:487-491 -> loss_utils.py:45-67) never exists.
//   row r (label y, -100 = ignore):  lse = log sum_j exp(x_j) ;  loss_r = lse - x_y
//   dlogits_j = (softmax_j - [j == y]) * gscale        gscale = loss_scale / n_valid  (read from device memory)
// One CTA per row; the whole row is cached in registers (V <= 256*8*NV) so HBM sees one read and one write.
#include "common.cuh"

using namespace odb;

template <int NV>
__global__ void __launch_bounds__(256) ce_fwd_bwd_kernel(__nv_bfloat16* __restrict__ logits,
                                                         const long long* __restrict__ labels, int V, long long row_stride,
                                                         const float* __restrict__ gscale_ptr, float* __restrict__ loss_sum,
                                                         float* __restrict__ lse_out, float* __restrict__ sumsq) {
  __shared__ float sm[33];
  const int row = blockIdx.x;
  const long long y = labels[row];
  uint4* xr = reinterpret_cast<uint4*>(logits + (size_t)row * row_stride);
  const int nvec = V / 8;
  const bool ignore = (y < 0);
  if (ignore && sumsq == nullptr) {
    // masked token: gradient is zero, no loss contribution
    for (int c = threadIdx.x; c < nvec; c += 256) st_na_v4(xr + c, make_uint4(0, 0, 0, 0));
    if (lse_out && threadIdx.x == 0) lse_out[row] = 0.f;
    return;
  }
  uint4 raw[NV];
  float mx = -INFINITY, sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < nvec) {
      raw[k] = ld_nc_v4(xr + c);
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { mx = fmaxf(mx, f[j]); sq += f[j] * f[j]; }
    }
  }
  mx = block_max(mx, sm);
  // exp() once per element: the fp32 values feed the row sum, their bf16 roundings replace the logits in registers and
  // become the softmax numerators of the second pass (the output is bf16 anyway).  The label logit is picked up here.
  float s = 0.f;
  float x_label = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < nvec) {
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if ((long long)(c * 8 + j) == y) x_label = f[j];
        f[j] = __expf(f[j] - mx);
        s += f[j];
      }
      raw[k] = pack8(f);
    }
  }
  s = block_sum(s, sm);
  x_label = block_sum(x_label, sm);     // exactly one thread holds a non-zero value
  if (sumsq) {
    sq = block_sum(sq, sm);
    if (threadIdx.x == 0) atomicAdd(sumsq, sq);
  }
  const float lse = mx + __logf(s);
  const float gscale = ignore ? 0.f : *gscale_ptr;
  const float inv_s = gscale / s;
  if (!ignore && threadIdx.x == 0) atomicAdd(loss_sum, lse - x_label);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < nvec) {
      float f[8], o[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = f[j] * inv_s - ((long long)(c * 8 + j) == y ? gscale : 0.f);
      st_na_v4(xr + c, pack8(o));
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
  ODB_DISPATCH_NV(V, (ce_fwd_bwd_kernel<NV><<<rows, 256, 0, st>>>((__nv_bfloat16*)logits, (const long long*)labels, V,
                                                                  row_stride, (const float*)gscale_ptr, (float*)loss_sum,
                                                                  (float*)lse_out, (float*)sumsq)));
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
