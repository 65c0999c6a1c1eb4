// HBM-bound row kernels of the denoiser block path (no tensor cores: these are bandwidth work).
//
// osb_ln_modulate: LayerNorm without affine (fp32 two-pass statistics on a register-resident row)
// fused with the adaLN modulate (1 + scale) * x + shift; one warp per row, 16-byte vector access,
// algorithmic traffic = read x + write y = 4 bytes per element.
// Replaces: opensora/models/mmdit/layers.py:205-206,223-224,248,252,312,400 and upstream v1.2
// t2i_modulate(norm(x), shift, scale).
#include "common.cuh"

namespace osb {

constexpr int kLnWarpsPerBlock = 8;

template <int NCH>  // 16-byte chunks per lane (row has C/8 chunks, lane handles chunk lane + 32*i)
__global__ void __launch_bounds__(kLnWarpsPerBlock * 32)
ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ shift,
                   const float* __restrict__ scale, __nv_bfloat16* __restrict__ y, int64_t rows, int C,
                   int64_t group_rows, const int32_t* __restrict__ mod_index, int64_t mod_stride, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kLnWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nchunks = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * C);

  float v[NCH][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      uint4 t;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(xr + c));
      const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(tw[e]);
        v[i][2 * e] = f.x;
        v[i][2 * e + 1] = f.y;
        sum += f.x + f.y;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);

  int64_t g = row / group_rows;
  if (mod_index) g = mod_index[g];
  const float* sh = shift + g * mod_stride;
  const float* sc = scale + g * mod_stride;
  uint4* yr = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc + c * 8));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(sc + c * 8 + 4));
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(sh + c * 8));
      const float4 h1 = __ldg(reinterpret_cast<const float4*>(sh + c * 8 + 4));
      const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * (1.0f + s[e]) + hh[e];
      uint4 t;
      t.x = pack_bf16x2(o[0], o[1]);
      t.y = pack_bf16x2(o[2], o[3]);
      t.z = pack_bf16x2(o[4], o[5]);
      t.w = pack_bf16x2(o[6], o[7]);
      yr[c] = t;
    }
  }
}

}  // namespace osb

extern "C" int osb_ln_modulate(const void* x, const float* shift, const float* scale, void* y,
                               int64_t rows, int C, int64_t group_rows, const int32_t* mod_index,
                               int64_t mod_stride, float eps, void* stream) {
  using namespace osb;
  if (!initialised()) { set_error("osb_init() has not been called"); return OSB_ERR_NOT_INIT; }
  OSB_REQUIRE(x && y && shift && scale, "osb_ln_modulate: null tensor");
  OSB_REQUIRE(rows > 0, "osb_ln_modulate: rows must be positive");
  OSB_REQUIRE(C > 0 && C % 8 == 0 && C <= 8192, "osb_ln_modulate: C must be a multiple of 8 and <= 8192 (got %d)", C);
  OSB_REQUIRE(mod_stride % 4 == 0, "osb_ln_modulate: mod_stride must be a multiple of 4");
  OSB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                reinterpret_cast<uintptr_t>(shift) | reinterpret_cast<uintptr_t>(scale)) & 15) == 0,
              "osb_ln_modulate: tensors must be 16-byte aligned");
  if (group_rows <= 0) group_rows = rows;
  const int nch = (C / 8 + 31) / 32;
  const unsigned blocks = (unsigned)((rows + kLnWarpsPerBlock - 1) / kLnWarpsPerBlock);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
#define OSB_LN_CASE(N)                                                                               \
  if (nch <= N) {                                                                                    \
    ln_modulate_kernel<N><<<blocks, kLnWarpsPerBlock * 32, 0, s>>>(xb, shift, scale, yb, rows, C,    \
                                                                   group_rows, mod_index, mod_stride, eps); \
    OSB_CHECK_CUDA(cudaGetLastError());                                                              \
    count_launch();                                                                                  \
    return OSB_OK;                                                                                   \
  }
  OSB_LN_CASE(1) OSB_LN_CASE(2) OSB_LN_CASE(3) OSB_LN_CASE(5) OSB_LN_CASE(8) OSB_LN_CASE(12)
  OSB_LN_CASE(16) OSB_LN_CASE(32)
#undef OSB_LN_CASE
  set_error("osb_ln_modulate: unsupported C %d", C);
  return OSB_ERR_UNSUPPORTED;
}
