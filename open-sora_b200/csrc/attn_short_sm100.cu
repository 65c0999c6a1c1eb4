// Attention for short key sets on sm_100a (tcgen05 + TMEM), one CTA per (128 query rows, head).
//
// The whole key/value set of a sequence (<= 320 keys, e.g. STDiT3 spatial S=256, temporal T=64,
// T5 cross-attention 300) is resident in shared memory, so softmax is a single exact pass over
// the S = Q K^T row held in TMEM - no online rescaling.  Short sequences (Lq < 128) are packed
// G = 128/Lq per tile with a block-diagonal mask.  Per-head RMSNorm of q,k and interleaved-pair
// RoPE are applied in fp32 while staging operands into the 128B-swizzled K-major smem tiles that
// feed tcgen05.mma, so q/k/v are read exactly once from the projection GEMM's output.
//
// Replaces: opensora/models/mmdit/math.py:22-36 (attention), layers.py:102-135 (QK RMSNorm) and the
// upstream-v1.2 STDiT3 Attention / MultiHeadCrossAttention restated in SURVEY.md App. A.
#include "common.cuh"

namespace osb {

constexpr int kAttnThreads = 128;
constexpr int kMaxKeys = 320;        // padded keys per tile
constexpr int kSCols = 320;          // TMEM columns reserved for S
constexpr int kChunkRowBytes = 128;  // one swizzle row = 64 bf16

struct AttnParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v; __nv_bfloat16* out;
  int64_t q_ld, k_ld, v_ld, out_ld;
  int64_t num_seqs, seqs_per_batch;
  int64_t q_bs, q_ss, q_ts, k_bs, k_ss, k_ts;
  int32_t Lq, Lk;
  const int32_t* kv_lens;
  int32_t H;
  const __nv_bfloat16* qw; const __nv_bfloat16* kw;
  float eps;
  const float* cos; const float* sin;
  float scale_log2;   // softmax_scale * log2(e)
  int32_t G;          // sequences packed per tile
  int32_t tiles_per_seq;
  int32_t NK, NKP;    // keys per tile (G*Lk) and padded to 16
};

template <int D>
struct AttnCfg {
  static constexpr int DP = (D + 15) / 16 * 16;          // padded head dim (MMA K of QK^T, N of PV)
  static constexpr int KC = (DP + 63) / 64;              // 64-wide K chunks of Q / K tiles
  static constexpr int Q_CHUNK = 128 * kChunkRowBytes;   // 16 KB
  static constexpr int K_CHUNK = kMaxKeys * kChunkRowBytes;  // 40 KB
  static constexpr int P_CHUNKS = kMaxKeys / 64;         // 5
  static constexpr int QK_BYTES = KC * (Q_CHUNK + K_CHUNK);
  static constexpr int P_BYTES = P_CHUNKS * Q_CHUNK;
  static constexpr int R1_BYTES = QK_BYTES > P_BYTES ? QK_BYTES : P_BYTES;  // P overlays Q,K
  static constexpr int VT_CHUNK = DP * kChunkRowBytes;
  static constexpr int VT_BYTES = P_CHUNKS * VT_CHUNK;
  static constexpr int SMEM_BYTES = R1_BYTES + VT_BYTES + 64 + 1024;
  static_assert(kSCols + DP <= 512, "S and O must fit TMEM");
};

// byte offset of 16-byte unit `u` (0..7) of row `r` inside a [rows x 64] bf16 SW128 K-major chunk
__device__ __forceinline__ uint32_t sw128_off(int r, int u) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4));
}

// load one head row (D bf16), optional RMSNorm (fp32 stats, weight) and interleaved-pair RoPE
template <int D>
__device__ __forceinline__ void load_head_row(const __nv_bfloat16* src, const __nv_bfloat16* w,
                                              float eps, const float* cosr, const float* sinr,
                                              float (&x)[D]) {
#pragma unroll
  for (int u = 0; u < D / 8; ++u) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(src) + u);
    const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(tw[e]);
      x[u * 8 + 2 * e] = f.x;
      x[u * 8 + 2 * e + 1] = f.y;
    }
  }
  if (w != nullptr) {
    float ss = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) ss += x[d] * x[d];
    const float r = rsqrtf(ss * (1.0f / D) + eps);
#pragma unroll
    for (int u = 0; u < D / 8; ++u) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(w) + u);
      const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(tw[e]);
        x[u * 8 + 2 * e] *= r * f.x;
        x[u * 8 + 2 * e + 1] *= r * f.y;
      }
    }
  }
  if (cosr != nullptr) {
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
      const float c = __ldg(cosr + i), s = __ldg(sinr + i);
      const float a = x[2 * i], b = x[2 * i + 1];
      x[2 * i] = a * c - b * s;
      x[2 * i + 1] = b * c + a * s;
    }
  }
}

// store a row of D floats (padded with zeros to DP) as bf16 into K-major SW128 chunks
template <int D, int DP>
__device__ __forceinline__ void store_row_kmajor(uint8_t* base, int chunk_bytes, int r, const float (&x)[D]) {
#pragma unroll
  for (int u = 0; u < DP / 8; ++u) {
    uint4 o;
    if (u * 8 < D) {
      o.x = pack_bf16x2(x[u * 8 + 0], x[u * 8 + 1]);
      o.y = pack_bf16x2(x[u * 8 + 2], x[u * 8 + 3]);
      o.z = pack_bf16x2(x[u * 8 + 4], x[u * 8 + 5]);
      o.w = pack_bf16x2(x[u * 8 + 6], x[u * 8 + 7]);
    } else {
      o = make_uint4(0, 0, 0, 0);
    }
    *reinterpret_cast<uint4*>(base + (u >> 3) * chunk_bytes + sw128_off(r, u & 7)) = o;
  }
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_short_kernel(const AttnParams p) {
  using Cfg = AttnCfg<D>;
  constexpr int DP = Cfg::DP;
  static_assert(D % 8 == 0, "head_dim must be a multiple of 8");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // KC chunks of [128 x 64]
  uint8_t* sK = smem + Cfg::KC * Cfg::Q_CHUNK;          // KC chunks of [320 x 64]
  uint8_t* sP = smem;                                   // overlays Q,K after S is complete
  uint8_t* sVt = smem + Cfg::R1_BYTES;                  // 5 chunks of [DP x 64]  (V transposed)
  const uint32_t bar_s = smem_u32(smem + Cfg::R1_BYTES + Cfg::VT_BYTES);
  const uint32_t bar_o = bar_s + 8;
  const uint32_t tmem_slot = bar_s + 16;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int tile = blockIdx.x;
  const int h = blockIdx.y;

  if (warp == 0) {
    if (tid == 0) {
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
  }

  // ---- which sequences / rows does this tile cover -------------------------------------
  int64_t seq0;
  int tok0;
  if (p.G > 1) { seq0 = (int64_t)tile * p.G; tok0 = 0; }
  else { seq0 = tile / p.tiles_per_seq; tok0 = (tile % p.tiles_per_seq) * 128; }

  // query row of this thread
  const int r = tid;
  const int g = (p.G > 1) ? r / p.Lq : 0;
  const int qtok = (p.G > 1) ? r % p.Lq : tok0 + r;
  const int64_t qseq = seq0 + g;
  const bool q_valid = (g < p.G) && (qseq < p.num_seqs) && (qtok < p.Lq);
  int64_t q_row = 0;
  if (q_valid) {
    const int64_t b = qseq / p.seqs_per_batch, j = qseq % p.seqs_per_batch;
    q_row = b * p.q_bs + j * p.q_ss + (int64_t)qtok * p.q_ts;
  }

  // ---- stage Q ---------------------------------------------------------------------------
  {
    float x[D];
    if (q_valid) {
      load_head_row<D>(p.q + q_row * p.q_ld + (int64_t)h * D, p.qw, p.eps,
                       p.cos ? p.cos + (int64_t)qtok * (D / 2) : nullptr,
                       p.sin ? p.sin + (int64_t)qtok * (D / 2) : nullptr, x);
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = 0.f;
    }
    store_row_kmajor<D, DP>(sQ, Cfg::Q_CHUNK, r, x);
  }
  // ---- stage K and V^T -------------------------------------------------------------------
  for (int slot = tid; slot < p.NKP; slot += kAttnThreads) {
    const int kg = slot / p.Lk, ktok = slot % p.Lk;
    const int64_t kseq = seq0 + kg;
    const bool k_valid = (slot < p.NK) && (kseq < p.num_seqs);
    int64_t k_row = 0;
    if (k_valid) {
      const int64_t b = kseq / p.seqs_per_batch, j = kseq % p.seqs_per_batch;
      k_row = b * p.k_bs + j * p.k_ss + (int64_t)ktok * p.k_ts;
    }
    float x[D];
    if (k_valid) {
      load_head_row<D>(p.k + k_row * p.k_ld + (int64_t)h * D, p.kw, p.eps,
                       p.cos ? p.cos + (int64_t)ktok * (D / 2) : nullptr,
                       p.sin ? p.sin + (int64_t)ktok * (D / 2) : nullptr, x);
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = 0.f;
    }
    store_row_kmajor<D, DP>(sK, Cfg::K_CHUNK, slot, x);

    if (k_valid) {
      load_head_row<D>(p.v + k_row * p.v_ld + (int64_t)h * D, nullptr, 0.f, nullptr, nullptr, x);
    }  // else x is already zero
    uint8_t* vt = sVt + (slot >> 6) * Cfg::VT_CHUNK;
    const int kk = slot & 63;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      *reinterpret_cast<__nv_bfloat16*>(vt + sw128_off(d, kk >> 3) + (kk & 7) * 2) = __float2bfloat16_rn(x[d]);
    }
  }

  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // ---- S = Q K^T -------------------------------------------------------------------------
  if (tid == 0) {
    for (int n0 = 0; n0 < p.NKP; n0 += 256) {
      const int n = (p.NKP - n0) < 256 ? (p.NKP - n0) : 256;
      const uint32_t idesc = make_idesc_bf16_f32(128, n);
      int step = 0;
#pragma unroll
      for (int kc = 0; kc < Cfg::KC; ++kc) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (kc * 64 + ks * 16 >= DP) break;
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(sQ) + kc * Cfg::Q_CHUNK + ks * 32);
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(sK) + kc * Cfg::K_CHUNK + n0 * kChunkRowBytes + ks * 32);
          umma_bf16<1>(tmem_base + n0, da, db, idesc, step > 0 ? 1u : 0u);
          ++step;
        }
      }
    }
    umma_commit<1>(bar_s);
  }
  mbar_wait(bar_s, 0);
  tc_fence_after();

  // ---- exact softmax over the row held in TMEM lane `r` ----------------------------------
  int key_lo = 0, key_hi = 0;
  if (q_valid) {
    const int len = p.kv_lens ? p.kv_lens[qseq] : p.Lk;
    key_lo = g * p.Lk;
    key_hi = key_lo + (len < p.Lk ? len : p.Lk);
  }
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  float mx = -INFINITY;
  for (int c0 = 0; c0 < p.NKP; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int key = c0 + j;
      const float s = __uint_as_float(v[j]);
      if (key >= key_lo && key < key_hi) mx = fmaxf(mx, s);
    }
  }
  const float mscaled = (mx == -INFINITY) ? 0.f : mx * p.scale_log2;
  float sum = 0.f;
  for (int c0 = 0; c0 < p.NKP; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + c0, v);
    tmem_ld_wait();
    float pr[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int key = c0 + j;
      const float s = __uint_as_float(v[j]);
      const float e = exp2f(s * p.scale_log2 - mscaled);
      pr[j] = (key >= key_lo && key < key_hi) ? e : 0.f;
      sum += pr[j];
    }
#pragma unroll
    for (int u4 = 0; u4 < 4; ++u4) {
      const int u = (c0 >> 3) + u4;  // 16-byte unit index along keys
      uint4 o;
      o.x = pack_bf16x2(pr[u4 * 8 + 0], pr[u4 * 8 + 1]);
      o.y = pack_bf16x2(pr[u4 * 8 + 2], pr[u4 * 8 + 3]);
      o.z = pack_bf16x2(pr[u4 * 8 + 4], pr[u4 * 8 + 5]);
      o.w = pack_bf16x2(pr[u4 * 8 + 6], pr[u4 * 8 + 7]);
      *reinterpret_cast<uint4*>(sP + (u >> 3) * Cfg::Q_CHUNK + sw128_off(r, u & 7)) = o;
    }
  }
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;

  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- O = P V ---------------------------------------------------------------------------
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(128, DP);
    const int steps = p.NKP / 16;
    for (int s = 0; s < steps; ++s) {
      const int c = s >> 2, ks = s & 3;
      const uint64_t da = make_sw128_kmajor_desc(smem_u32(sP) + c * Cfg::Q_CHUNK + ks * 32);
      const uint64_t db = make_sw128_kmajor_desc(smem_u32(sVt) + c * Cfg::VT_CHUNK + ks * 32);
      umma_bf16<1>(tmem_base + kSCols, da, db, idesc, s > 0 ? 1u : 0u);
    }
    umma_commit<1>(bar_o);
  }
  mbar_wait(bar_o, 0);
  tc_fence_after();

  // ---- epilogue: normalise, round once to bf16, store -------------------------------------
  __nv_bfloat16* orow = p.out + q_row * p.out_ld + (int64_t)h * D;
#pragma unroll 1
  for (int c0 = 0; c0 < D; c0 += 8) {
    uint32_t v[8];
    tmem_ld_32x32b_x8(t_row + kSCols + c0, v);
    tmem_ld_wait();
    if (q_valid) {
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
      o.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
      o.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
      o.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
      *reinterpret_cast<uint4*>(orow + c0) = o;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int D>
static int attn_launch(const AttnParams& p, int tiles, cudaStream_t stream) {
  dim3 grid((unsigned)tiles, (unsigned)p.H);
  attn_short_kernel<D><<<grid, kAttnThreads, AttnCfg<D>::SMEM_BYTES, stream>>>(p);
  OSB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OSB_OK;
}

int attn_init() {
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_short_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      AttnCfg<64>::SMEM_BYTES));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_short_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      AttnCfg<72>::SMEM_BYTES));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_short_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      AttnCfg<128>::SMEM_BYTES));
  return OSB_OK;
}

}  // namespace osb

extern "C" int osb_attn_short(const osb_attn_short_args* a, void* stream) {
  using namespace osb;
  if (!initialised()) { set_error("osb_init() has not been called"); return OSB_ERR_NOT_INIT; }
  OSB_REQUIRE(a != nullptr, "osb_attn_short: null args");
  OSB_REQUIRE(a->q && a->k && a->v && a->out, "osb_attn_short: null tensor");
  OSB_REQUIRE(a->Lq > 0 && a->Lk > 0 && a->num_seqs > 0 && a->num_heads > 0,
              "osb_attn_short: empty problem");
  OSB_REQUIRE(a->seqs_per_batch > 0, "osb_attn_short: seqs_per_batch must be positive");
  const int D = a->head_dim;
  OSB_REQUIRE(D == 64 || D == 72 || D == 128, "osb_attn_short: head_dim %d not built (64, 72, 128)", D);
  OSB_REQUIRE((a->q_ld % 8) == 0 && (a->k_ld % 8) == 0 && (a->v_ld % 8) == 0 && (a->out_ld % 8) == 0,
              "osb_attn_short: leading dimensions must be multiples of 8 elements");
  OSB_REQUIRE(((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
                reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->out)) & 15) == 0,
              "osb_attn_short: tensors must be 16-byte aligned");
  OSB_REQUIRE((a->q_norm_w == nullptr) == (a->k_norm_w == nullptr), "osb_attn_short: q/k norm weights must come together");
  OSB_REQUIRE((a->rope_cos == nullptr) == (a->rope_sin == nullptr), "osb_attn_short: rope cos/sin must come together");

  AttnParams p;
  p.q = static_cast<const __nv_bfloat16*>(a->q);
  p.k = static_cast<const __nv_bfloat16*>(a->k);
  p.v = static_cast<const __nv_bfloat16*>(a->v);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.q_ld = a->q_ld; p.k_ld = a->k_ld; p.v_ld = a->v_ld; p.out_ld = a->out_ld;
  p.num_seqs = a->num_seqs; p.seqs_per_batch = a->seqs_per_batch;
  p.q_bs = a->q_batch_stride; p.q_ss = a->q_seq_stride; p.q_ts = a->q_tok_stride;
  p.k_bs = a->k_batch_stride; p.k_ss = a->k_seq_stride; p.k_ts = a->k_tok_stride;
  p.Lq = a->Lq; p.Lk = a->Lk; p.kv_lens = a->kv_lens; p.H = a->num_heads;
  p.qw = static_cast<const __nv_bfloat16*>(a->q_norm_w);
  p.kw = static_cast<const __nv_bfloat16*>(a->k_norm_w);
  p.eps = a->norm_eps;
  p.cos = a->rope_cos; p.sin = a->rope_sin;
  p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
  int64_t tiles;
  if (a->Lq >= 128) {
    p.G = 1;
    p.tiles_per_seq = (a->Lq + 127) / 128;
    tiles = a->num_seqs * p.tiles_per_seq;
  } else {
    p.G = 128 / a->Lq;
    p.tiles_per_seq = 1;
    tiles = (a->num_seqs + p.G - 1) / p.G;
  }
  // packing more sequences than the key budget allows would overflow the S tile: shrink G
  while (p.G > 1 && (int64_t)p.G * a->Lk > kMaxKeys) {
    --p.G;
    tiles = (a->num_seqs + p.G - 1) / p.G;
  }
  p.NK = p.G * a->Lk;
  OSB_REQUIRE(p.NK <= kMaxKeys, "osb_attn_short: %d keys per tile exceed the %d-key budget (use the streaming kernel)",
              p.NK, kMaxKeys);
  p.NKP = (p.NK + 15) / 16 * 16;
  OSB_REQUIRE(tiles <= 0x7fffffff && a->num_heads <= 65535, "osb_attn_short: grid too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (D == 64) return attn_launch<64>(p, (int)tiles, s);
  if (D == 72) return attn_launch<72>(p, (int)tiles, s);
  return attn_launch<128>(p, (int)tiles, s);
}
