// Attention for short key sets on sm_100a (tcgen05 + TMEM).
//
// One CTA (8 warps) owns one key/value set - a sequence and head (STDiT3 spatial S=256, T5
// cross-attention 300 keys) or G = 128/Lq packed short sequences with a block-diagonal mask
// (temporal T=64) - stages K and V^T into shared memory ONCE, then loops over up to QT 128-row query
// tiles against it.  Because the whole key set is resident, softmax is a single exact pass over the
// S = Q K^T row held in TMEM (no online rescaling): S via tcgen05.mma, row max / exp2 / sum by the
// two threads that share a row (one per half of the key range), P written as bf16 into a swizzled
// K-major tile, O = P V via tcgen05.mma, normalised in fp32 and rounded once to bf16.
// Per-head RMSNorm of q,k and interleaved-pair RoPE are applied in fp32 while staging the operands,
// so q/k/v are read exactly once, straight from the projection GEMM's output.
//
// Operand tiles are K-major: 64-wide chunks in the 128B-swizzle layout; the head-dim tail of D=72
// (columns 64..79, zero padded) uses the no-swizzle core-matrix layout (8 rows x 16 B) so that K, Q
// tiles cost 160 B per row instead of 256 B - this is what lets K + V^T + P + Q fit for 304 keys.
//
// Replaces: opensora/models/mmdit/math.py:22-36 (attention), layers.py:102-135 (QK RMSNorm) and the
// upstream-v1.2 STDiT3 Attention / MultiHeadCrossAttention restated in SURVEY.md App. A.
#include <stdlib.h>

#include "common.cuh"
#include "tiles.cuh"

namespace osb {

constexpr int kAttnThreads = 256;
constexpr int kMaxKeys = 320;  // padded keys per tile (resident kernel)

struct AttnParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v; __nv_bfloat16* out;
  int64_t q_ld, k_ld, v_ld, out_ld;
  int64_t num_seqs, seqs_per_batch;
  int64_t q_bs, q_ss, q_ts, k_bs, k_ss, k_ts;
  int32_t Lq, Lk;
  const int32_t* kv_lens;
  const __nv_bfloat16* qw; const __nv_bfloat16* kw;
  const __nv_bfloat16* qw2; const __nv_bfloat16* kw2;  // weights for tokens >= norm_split (joint txt|img sequences)
  int32_t norm_split;
  float eps;
  const float* cos; const float* sin;
  int32_t rope_half;     // 1: rotate-half pairing (i, i + D/2) (HF / Liger layout), 0: interleaved pairs (2i, 2i+1)
  float scale_log2;      // softmax_scale * log2(e)
  int32_t G;             // sequences packed per tile (1 when Lq >= 128)
  int32_t tiles_per_seq; // q-tiles per sequence (G == 1)
  int32_t QT;            // q-tiles handled per CTA (G == 1)
  int32_t groups_per_seq;
  int32_t NK, NKP;       // keys per CTA (G*Lk) and padded to 16
  int32_t tmem_cols;     // 256 or 512
  int32_t o_col;         // TMEM column of the O accumulator
  // shared memory carve-up (bytes from the 1024-aligned base)
  int32_t off_qt, off_k, off_kt, off_vt, off_p, off_misc;
};

template <int D>
struct AttnCfg {
  static constexpr int DP = (D + 15) / 16 * 16;   // padded head dim (MMA K of QK^T, N of PV)
  static constexpr int MAIN = D / 64;             // full 64-wide swizzled chunks of the Q / K tiles
  static constexpr int TAIL = DP - MAIN * 64;     // 0 or 16: head-dim tail in the no-swizzle layout
  static constexpr int U = D / 8;                 // 16-byte units per head row
  static constexpr int UP = DP / 8;
  static constexpr int U0 = (U + 1) / 2;          // units handled by the first thread of a row pair
  static_assert(TAIL == 0 || TAIL == 16, "head_dim tail must be one MMA K step");
  static_assert(D % 8 == 0, "head_dim must be a multiple of 8");
};

__device__ __forceinline__ void unpack8(const uint4& t, float* x) {
  const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_bf16x2(tw[e]);
    x[2 * e] = f.x;
    x[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* x) {
  uint4 o;
  o.x = pack_bf16x2(x[0], x[1]);
  o.y = pack_bf16x2(x[2], x[3]);
  o.z = pack_bf16x2(x[4], x[5]);
  o.w = pack_bf16x2(x[6], x[7]);
  return o;
}

__device__ __forceinline__ float sumsq8(const uint4& t) {
  float x[8];
  unpack8(t, x);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += x[e] * x[e];
  return s;
}

// Units [UB, UE) of one head row, given as raw bf16x8 registers: optional RMSNorm scale r*w and RoPE in
// fp32 -> bf16 -> K-major operand tile (main chunks swizzled, tail no-swizzle); units >= U are zero pad.
template <int D, int UB, int UE>
__device__ __forceinline__ float finish_and_store_units(const uint4* raw, float r, const __nv_bfloat16* w,
                                                        const float* cosr, const float* sinr, uint8_t* main_base,
                                                        int main_chunk_bytes, uint8_t* tail_base, int row,
                                                        bool want_norm = false) {
  using Cfg = AttnCfg<D>;
  float nrm2 = 0.f;  // squared length of the staged vector part (RoPE is a rotation: measured before it)
#pragma unroll
  for (int u = UB; u < UE; ++u) {
    uint4 o;
    if (u < Cfg::U) {
      o = raw[u - UB];
      if (w == nullptr && cosr == nullptr && want_norm) nrm2 += sumsq8(o);
      if (w != nullptr || cosr != nullptr) {
        float xu[8];
        unpack8(o, xu);
        if (w != nullptr) {
          float wf[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(w) + u), wf);
#pragma unroll
          for (int e = 0; e < 8; ++e) xu[e] *= r * wf[e];
        }
        if (want_norm) {
#pragma unroll
          for (int e = 0; e < 8; ++e) nrm2 += xu[e] * xu[e];
        }
        if (cosr != nullptr) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(cosr) + u);
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(sinr) + u);
          const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = xu[2 * i], b = xu[2 * i + 1];
            xu[2 * i] = a * cc[i] - b * ss[i];
            xu[2 * i + 1] = b * cc[i] + a * ss[i];
          }
        }
        o = pack8(xu);
      }
    } else {
      o = make_uint4(0, 0, 0, 0);
    }
    if (u < Cfg::MAIN * 8)
      *reinterpret_cast<uint4*>(main_base + (u >> 3) * main_chunk_bytes + sw128_off(row, u & 7)) = o;
    else
      *reinterpret_cast<uint4*>(tail_base + tail_off(row, u - Cfg::MAIN * 8)) = o;
  }
  return nrm2;
}

// Full head row with rotate-half RoPE (LigerRopeFunction, math.py:27): element i pairs with i + D/2, so units u and
// u + U/2 are processed together: x1' = x1 cos_i - x2 sin_i, x2' = x2 cos_i + x1 sin_i.
template <int D>
__device__ __forceinline__ void finish_and_store_row_half(const uint4* raw, float r, const __nv_bfloat16* w, const float* cosr,
                                                          const float* sinr, uint8_t* main_base, int main_chunk_bytes,
                                                          uint8_t* tail_base, int row) {
  using Cfg = AttnCfg<D>;
  constexpr int U = Cfg::U, HU = U / 2;
  static_assert(U % 2 == 0 || true, "rotate-half needs an even number of 16-byte units");
#pragma unroll
  for (int u = 0; u < HU; ++u) {
    float x1[8], x2[8];
    unpack8(raw[u], x1);
    unpack8(raw[u + HU], x2);
    if (w != nullptr) {
      float w1[8], w2[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + u), w1);
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + u + HU), w2);
#pragma unroll
      for (int e = 0; e < 8; ++e) { x1[e] *= r * w1[e]; x2[e] *= r * w2[e]; }
    }
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(cosr) + 2 * u), c1 = __ldg(reinterpret_cast<const float4*>(cosr) + 2 * u + 1);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sinr) + 2 * u), s1 = __ldg(reinterpret_cast<const float4*>(sinr) + 2 * u + 1);
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = x1[e], b = x2[e];
      x1[e] = a * cc[e] - b * ss[e];
      x2[e] = b * cc[e] + a * ss[e];
    }
    const uint4 o1 = pack8(x1), o2 = pack8(x2);
    const int ua = u, ub = u + HU;
    if (ua < Cfg::MAIN * 8) *reinterpret_cast<uint4*>(main_base + (ua >> 3) * main_chunk_bytes + sw128_off(row, ua & 7)) = o1;
    else *reinterpret_cast<uint4*>(tail_base + tail_off(row, ua - Cfg::MAIN * 8)) = o1;
    if (ub < Cfg::MAIN * 8) *reinterpret_cast<uint4*>(main_base + (ub >> 3) * main_chunk_bytes + sw128_off(row, ub & 7)) = o2;
    else *reinterpret_cast<uint4*>(tail_base + tail_off(row, ub - Cfg::MAIN * 8)) = o2;
  }
#pragma unroll
  for (int u = U; u < Cfg::UP; ++u)
    *reinterpret_cast<uint4*>(tail_base + tail_off(row, u - Cfg::MAIN * 8)) = make_uint4(0, 0, 0, 0);
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads, (D <= 72) ? 2 : 1) attn_short_kernel(const AttnParams p) {
  using Cfg = AttnCfg<D>;
  constexpr int DP = Cfg::DP, U = Cfg::U, UP = Cfg::UP, U0 = Cfg::U0;

  extern __shared__ __align__(1024) uint8_t smem_raw[];  // 1024-byte alignment: SWIZZLE_128B atoms
  uint8_t* smem = smem_raw;
  uint8_t* sQ = smem;                  // MAIN chunks of [128 x 64]
  uint8_t* sQt = smem + p.off_qt;      // [128 x 16] tail
  uint8_t* sK = smem + p.off_k;        // MAIN chunks of [NKP x 64]
  uint8_t* sKt = smem + p.off_kt;      // [NKP x 16] tail
  uint8_t* sVt = smem + p.off_vt;      // ceil(NKP/64) chunks of [DP x 64]   (V transposed)
  uint8_t* sP = smem + p.off_p;        // ceil(NKP/64) chunks of [128 x 64]
  float* xch = reinterpret_cast<float*>(smem + p.off_misc);          // [2][128] pair exchange (ss / max)
  float* xsum = xch + 256;                                           // [2][128] partial row sums
  float* xqn = xsum;   // [2][128] partial |q|^2 of the staged rows: read before the max-exchange barrier, xsum written after it
  uint32_t* kmax2 = reinterpret_cast<uint32_t*>(xch + 512);          // max |k|^2 over the staged keys (float bits)
  const uint32_t bar_s = smem_u32(smem + p.off_misc + 2056);
  const uint32_t bar_o = bar_s + 8;
  const uint32_t tmem_slot = bar_s + 16;
  const int k_chunk_bytes = p.NKP * 128;
  const int vt_chunk_bytes = DP * 128;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int r = tid & 127;      // query row inside the tile
  const int part = tid >> 7;    // 0/1: which member of the row pair this thread is
  const int unit = blockIdx.x;
  const int h = blockIdx.y;

  if (warp == 0) {
    if (tid == 0) {
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      fence_barrier_init();
      *kmax2 = 0u;
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, (uint32_t)p.tmem_cols);
  }
  __syncthreads();  // kmax2 is zeroed before any key's atomicMax

  // ---- which sequences / q-tiles does this CTA cover --------------------------------------
  int64_t seq0;
  int qt_begin, qt_end;
  if (p.G > 1) { seq0 = (int64_t)unit * p.G; qt_begin = 0; qt_end = 1; }
  else {
    seq0 = unit / p.groups_per_seq;
    qt_begin = (unit % p.groups_per_seq) * p.QT;
    qt_end = qt_begin + p.QT < p.tiles_per_seq ? qt_begin + p.QT : p.tiles_per_seq;
  }

  pdl_wait();  // barrier init / TMEM allocation above overlapped the previous kernel; q, k, v are visible from here
  pdl_launch_dependents();  // the next kernel may set up on SMs as our CTAs retire (its own pdl_wait orders the data)
  // ---- Q rows are prefetched into registers one q-tile ahead: the first tile's loads fly during the K / V staging,
  //      the next tile's during the current tile's softmax (saves one exposed DRAM round trip per tile) ----------
  const int pf_ub = part == 0 ? 0 : U0;
  const int pf_ue = part == 0 ? U0 : U;
  uint4 tq[U0];
  auto prefetch_q = [&](int qt) {
    const int g_ = (p.G > 1) ? r / p.Lq : 0;
    const int tok_ = (p.G > 1) ? r - g_ * p.Lq : qt * 128 + r;
    const int64_t seq_ = seq0 + g_;
#pragma unroll
    for (int i = 0; i < U0; ++i) tq[i] = make_uint4(0, 0, 0, 0);
    if ((g_ < p.G) && (seq_ < p.num_seqs) && (tok_ < p.Lq)) {
      const int64_t b = seq_ / p.seqs_per_batch, j = seq_ % p.seqs_per_batch;
      const int64_t row_ = b * p.q_bs + j * p.q_ss + (int64_t)tok_ * p.q_ts;
      const uint4* qs = reinterpret_cast<const uint4*>(p.q + row_ * p.q_ld + (int64_t)h * D);
#pragma unroll
      for (int i = 0; i < U0; ++i)
        if (pf_ub + i < pf_ue) tq[i] = __ldg(qs + pf_ub + i);
    }
  };
  prefetch_q(qt_begin);

  // ---- stage K and V^T once -----------------------------------------------------------------
  for (int slot = tid; slot < p.NKP; slot += kAttnThreads) {
    const int kg = slot / p.Lk, ktok = slot - kg * p.Lk;
    const int64_t kseq = seq0 + kg;
    const bool k_valid = (slot < p.NK) && (kseq < p.num_seqs);
    int64_t k_row = 0;
    if (k_valid) {
      const int64_t b = kseq / p.seqs_per_batch, j = kseq % p.seqs_per_batch;
      k_row = b * p.k_bs + j * p.k_ss + (int64_t)ktok * p.k_ts;
    }
    uint4 tk[UP], tv[U];
    if (k_valid) {  // issue every load of the slot before touching any of them
      const uint4* ks = reinterpret_cast<const uint4*>(p.k + k_row * p.k_ld + (int64_t)h * D);
      const uint4* vs = reinterpret_cast<const uint4*>(p.v + k_row * p.v_ld + (int64_t)h * D);
#pragma unroll
      for (int u = 0; u < U; ++u) tk[u] = __ldg(ks + u);
#pragma unroll
      for (int u = 0; u < U; ++u) tv[u] = __ldg(vs + u);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) { tk[u] = make_uint4(0, 0, 0, 0); tv[u] = make_uint4(0, 0, 0, 0); }
    }
    float rk = 1.f;
    if (p.kw != nullptr) {
      float ss = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) ss += sumsq8(tk[u]);
      rk = rsqrtf(ss * (1.0f / D) + p.eps);
    }
    const __nv_bfloat16* kwt = (p.kw2 != nullptr && ktok >= p.norm_split) ? p.kw2 : p.kw;
    const float kn2 = finish_and_store_units<D, 0, UP>(tk, rk, kwt, p.cos ? p.cos + (int64_t)ktok * (D / 2) : nullptr,
                                                       p.sin ? p.sin + (int64_t)ktok * (D / 2) : nullptr, sK, k_chunk_bytes,
                                                       sKt, slot, true);
    atomicMax(kmax2, __float_as_uint(kn2));  // non-negative floats order like their bit patterns
    // V^T: raw bf16 halves go straight to their transposed position (no fp32 round trip)
    uint8_t* vt = sVt + (slot >> 6) * vt_chunk_bytes + (slot & 7) * 2;
    const int ku = (slot & 63) >> 3;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t tw[4] = {tv[u].x, tv[u].y, tv[u].z, tv[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        *reinterpret_cast<uint16_t*>(vt + sw128_off(u * 8 + 2 * e, ku)) = (uint16_t)(tw[e] & 0xffffu);
        *reinterpret_cast<uint16_t*>(vt + sw128_off(u * 8 + 2 * e + 1, ku)) = (uint16_t)(tw[e] >> 16);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's TMEM lane quarter

  // key chunks (32 wide) of the S row handled by this member of the row pair
  const int n32 = (p.NKP + 31) >> 5;
  const int c_begin = part == 0 ? 0 : (n32 + 1) / 2;
  const int c_end = part == 0 ? (n32 + 1) / 2 : n32;

  for (int qt = qt_begin; qt < qt_end; ++qt) {
    const uint32_t par = (uint32_t)(qt - qt_begin) & 1u;
    // ---- query row of this thread pair ---------------------------------------------------------
    const int g = (p.G > 1) ? r / p.Lq : 0;
    const int qtok = (p.G > 1) ? r - g * p.Lq : qt * 128 + r;
    const int64_t qseq = seq0 + g;
    const bool q_valid = (g < p.G) && (qseq < p.num_seqs) && (qtok < p.Lq);
    int64_t q_row = 0;
    if (q_valid) {
      const int64_t b = qseq / p.seqs_per_batch, j = qseq % p.seqs_per_batch;
      q_row = b * p.q_bs + j * p.q_ss + (int64_t)qtok * p.q_ts;
    }
    // ---- stage Q: the pair splits the row's units [0,U0) / [U0,UP) ------------------------------
    {
      uint4 t[U0];
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < U0; ++i) t[i] = tq[i];   // prefetched one tile ahead
      float rq = 1.f;
      if (p.qw != nullptr) {
#pragma unroll
        for (int i = 0; i < U0; ++i) ss += sumsq8(t[i]);
        xch[part * 128 + r] = ss;
        __syncthreads();
        rq = rsqrtf((xch[r] + xch[128 + r]) * (1.0f / D) + p.eps);
      }
      const float* cq = p.cos ? p.cos + (int64_t)qtok * (D / 2) : nullptr;
      const float* sq = p.sin ? p.sin + (int64_t)qtok * (D / 2) : nullptr;
      const __nv_bfloat16* qwt = (p.qw2 != nullptr && qtok >= p.norm_split) ? p.qw2 : p.qw;
      float qn2;
      if (part == 0) qn2 = finish_and_store_units<D, 0, U0>(t, rq, qwt, cq, sq, sQ, 128 * 128, sQt, r, true);
      else qn2 = finish_and_store_units<D, U0, UP>(t, rq, qwt, cq, sq, sQ, 128 * 128, sQt, r, true);
      xqn[part * 128 + r] = qn2;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- S = Q K^T ----------------------------------------------------------------------------
    // (warp 0 computes the descriptors warp-uniformly and one elected lane issues: operands stay in uniform registers
    // and the MMAs go out back to back instead of one uniformisation loop per instruction)
    if (warp == 0) {
      const uint64_t qd0 = make_sw128_kmajor_desc(smem_u32(sQ)), kd0 = make_sw128_kmajor_desc(smem_u32(sK));
      const uint64_t qtd = make_noswz_kmajor_desc(smem_u32(sQt)), ktd = make_noswz_kmajor_desc(smem_u32(sKt));
      const uint32_t kchunk16 = (uint32_t)k_chunk_bytes >> 4;
      if (elect_one()) {
        for (int n0 = 0; n0 < p.NKP; n0 += 256) {
          const int n = (p.NKP - n0) < 256 ? (p.NKP - n0) : 256;
          const uint32_t idesc = make_idesc_bf16_f32(128, n);
          uint32_t acc = 0;
#pragma unroll
          for (int kc = 0; kc < Cfg::MAIN; ++kc) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_bf16<1>(tmem_base + n0, qd0 + (uint64_t)(kc * ((128 * 128) >> 4) + ks * 2),
                           kd0 + (uint64_t)(kc * kchunk16 + n0 * (128 >> 4) + ks * 2), idesc, acc);
              acc = 1;
            }
          }
          if (Cfg::TAIL) umma_bf16<1>(tmem_base + n0, qtd, ktd + (uint64_t)((n0 >> 3) * (256 >> 4)), idesc, acc);
        }
        umma_commit<1>(bar_s);
      }
      __syncwarp();
    }
    if (qt + 1 < qt_end) prefetch_q(qt + 1);
    mbar_wait(bar_s, par);
    tc_fence_after();

    // ---- exact softmax: each member of the pair owns half of the key chunks ----------------------
    int key_lo = 0, key_hi = 0;
    if (q_valid) {
      const int len = p.kv_lens ? p.kv_lens[qseq] : p.Lk;
      key_lo = g * p.Lk;
      key_hi = key_lo + (len < p.Lk ? len : p.Lk);
    }
    // tcgen05.ld is warp-collective: skip decisions use the union of the warp's key ranges
    const int w_lo = __reduce_min_sync(0xffffffffu, q_valid ? key_lo : 0x7fffffff);
    const int w_hi = __reduce_max_sync(0xffffffffu, q_valid ? key_hi : 0);
    // Softmax is shift invariant: any shift >= the row maximum that does not underflow everything is exact.
    // Cauchy-Schwarz gives one for free, |q| max|k| >= max_j q.k_j; when it is within 2^64 of fp32 range (it is for
    // RMS-normed q,k) the separate row-maximum pass over S (one TMEM read + compare per element) is skipped.
    const float bound = sqrtf((xqn[r] + xqn[128 + r]) * __uint_as_float(*kmax2)) * fabsf(p.scale_log2);
    const bool one_pass = __all_sync(0xffffffffu, bound < 64.f);
    float mx = -INFINITY;
    if (!one_pass) {
      for (int c = c_begin; c < c_end; ++c) {
        const int c0 = c * 32;
        if (c0 >= w_hi || c0 + 32 <= w_lo) continue;
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c0, v);
        tmem_ld_wait();
        if (c0 >= key_lo && c0 + 32 <= key_hi) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j >= key_lo && c0 + j < key_hi) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
    }
    xch[part * 128 + r] = mx;
    __syncthreads();   // (also orders the pair's reads of xqn/kmax2 before the next q-tile's writes)
    mx = fmaxf(xch[r], xch[128 + r]);
    const float mscaled = one_pass ? bound : ((mx == -INFINITY) ? 0.f : mx * p.scale_log2);
    float sum = 0.f;
    for (int c = c_begin; c < c_end; ++c) {
      const int c0 = c * 32;
      float pr[32];
      if (c0 >= w_hi || c0 + 32 <= w_lo) {
#pragma unroll
        for (int j = 0; j < 32; ++j) pr[j] = 0.f;
      } else {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c0, v);
        tmem_ld_wait();
        if (c0 >= key_lo && c0 + 32 <= key_hi) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            pr[j] = fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - mscaled);
            sum += pr[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - mscaled);
            pr[j] = (c0 + j >= key_lo && c0 + j < key_hi) ? e : 0.f;
            sum += pr[j];
          }
        }
      }
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const int u = (c0 >> 3) + u4;  // 16-byte unit index along keys
        *reinterpret_cast<uint4*>(sP + (u >> 3) * (128 * 128) + sw128_off(r, u & 7)) = pack8(pr + u4 * 8);
      }
    }
    xsum[part * 128 + r] = sum;

    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- O = P V ---------------------------------------------------------------------------------
    if (warp == 0) {
      const uint64_t pd0 = make_sw128_kmajor_desc(smem_u32(sP)), vd0 = make_sw128_kmajor_desc(smem_u32(sVt));
      const uint32_t vchunk16 = (uint32_t)vt_chunk_bytes >> 4;
      const uint32_t idesc = make_idesc_bf16_f32(128, DP);
      const int steps = p.NKP / 16;
      if (elect_one()) {
        for (int s = 0; s < steps; ++s) {
          const int c = s >> 2, ks = s & 3;
          umma_bf16<1>(tmem_base + p.o_col, pd0 + (uint64_t)(c * ((128 * 128) >> 4) + ks * 2),
                       vd0 + (uint64_t)(c * vchunk16 + ks * 2), idesc, s > 0 ? 1u : 0u);
        }
        umma_commit<1>(bar_o);
      }
      __syncwarp();
    }
    const float tot = xsum[r] + xsum[128 + r];
    const float inv = tot > 0.f ? 1.0f / tot : 0.f;
    mbar_wait(bar_o, par);
    tc_fence_after();

    // ---- epilogue: normalise, round once to bf16, store (the pair splits the head columns) ----------
    __nv_bfloat16* orow = p.out + q_row * p.out_ld + (int64_t)h * D;
    const int ub = part == 0 ? 0 : U0, ue = part == 0 ? U0 : U;
#pragma unroll 1
    for (int u = ub; u < ue; ++u) {
      uint32_t v[8];
      tmem_ld_32x32b_x8(t_row + p.o_col + u * 8, v);
      tmem_ld_wait();
      if (q_valid) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[e]) * inv;
        *reinterpret_cast<uint4*>(orow + u * 8) = pack8(o);
      }
    }
    tc_fence_before();
    __syncthreads();  // S, O, sQ, sP and the exchange buffers are reused by the next q-tile
    tc_fence_after();
  }

  if (warp == 0) tmem_dealloc<1>(tmem_base, (uint32_t)p.tmem_cols);
}


// ==========================================================================================================
// Flash variant: warp-specialised, key blocks of BK <= 128, online softmax.  Roles (288 threads): warps 0-3 =
// softmax / O-correction / epilogue (one thread per query row), warps 4-7 = loaders (RMSNorm + RoPE in fp32 while
// staging Q, K and V into tensor-core operand tiles), warp 8 = tcgen05 issuer.  All hand-offs are mbarriers.
//  * S (fp32) and P (bf16) share one TMEM region: P is written in place over S (kPTmem) and fed to the PV MMA as
//    a tensor-memory A operand, or goes through a swizzled smem tile; O accumulates in a second TMEM region and is
//    rescaled there only when a row maximum actually grows.
//  * V is NOT transposed: it is staged row-major ([key][d], the same tile layout as K) and consumed as an
//    MN-major B operand (64-wide main chunk 128B-swizzled, the 72->80 head-dim tail as a second N=16 MMA over a
//    no-swizzle tile), so staging V costs ten 16-byte stores per key instead of 72 two-byte ones.
//  * Key sets that fit the ring (<= 3 blocks: STDiT3 spatial / temporal / T5 cross) stay RESIDENT: they are staged
//    once per CTA and reused by the QT query tiles of the work item; longer key sets (MMDiT's 8828-token joint
//    sequence) stream through a 2-stage ring.
//  208-240 TMEM columns and <= ~100 KB smem at D=72 -> two CTAs per SM: one CTA's softmax overlaps the other's
//  staging and MMAs.
// ==========================================================================================================
constexpr int kFlashThreads = 288;

struct FlashGeom {
  int32_t BK, NKB, NST;      // keys per block (multiple of 16, <= 128), blocks per key set, ring stages
  int32_t resident;          // 1: NST == NKB, key blocks staged once per work item
  int32_t QT, groups_per_seq; // q-tiles per work item (G == 1)
  int32_t kv_stage_bytes, k_bytes, kt_bytes, v_bytes;  // per-stage smem sizes
  int32_t off_qt, off_kv, off_p, off_bar;
  int32_t o_col, tmem_cols;
  int64_t units;             // work units per head
  int64_t items;             // units * heads
};

template <int D, bool kPTmem>
__global__ void __launch_bounds__(kFlashThreads, (D <= 72) ? 2 : 1) attn_flash_kernel(const AttnParams p, const FlashGeom g) {
  using Cfg = AttnCfg<D>;
  constexpr int U = Cfg::U, UP = Cfg::UP;
  constexpr int NMAIN = Cfg::MAIN * 64;   // O columns produced by the swizzled main chunk(s)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sQt = smem + g.off_qt;
  auto sK = [&](int st) { return smem + g.off_kv + st * g.kv_stage_bytes; };
  auto sKt = [&](int st) { return smem + g.off_kv + st * g.kv_stage_bytes + g.k_bytes; };
  auto sV = [&](int st) { return smem + g.off_kv + st * g.kv_stage_bytes + g.k_bytes + g.kt_bytes; };
  auto sVt = [&](int st) { return smem + g.off_kv + st * g.kv_stage_bytes + g.k_bytes + g.kt_bytes + g.v_bytes; };
  uint8_t* sP = smem + g.off_p;
  const uint32_t bar0 = smem_u32(smem + g.off_bar);
  const uint32_t q_full = bar0, q_empty = bar0 + 8, s_full = bar0 + 16, p_full = bar0 + 24, o_full = bar0 + 32;
  auto kv_full = [&](int st) { return bar0 + 40 + 8 * st; };
  auto kv_empty = [&](int st) { return bar0 + 64 + 8 * st; };
  const uint32_t tmem_slot = bar0 + 88;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 8) {
    if ((tid & 31) == 0) {
      mbar_init(q_full, 128); mbar_init(q_empty, 1); mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
      for (int st = 0; st < 3; ++st) { mbar_init(kv_full(st), 128); mbar_init(kv_empty(st), 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, (uint32_t)g.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  pdl_wait();
  pdl_launch_dependents();
  // work item -> (sequence(s), q-tile range, head); consecutive items share a head (K/V of a head stay in L2)
  auto decode = [&](int64_t item, int64_t& seq0, int& qt0, int& qt1, int& h) {
    h = (int)(item / g.units);
    const int64_t unit = item % g.units;
    if (p.G > 1) { seq0 = unit * p.G; qt0 = 0; qt1 = 1; }
    else {
      seq0 = unit / g.groups_per_seq;
      qt0 = (int)(unit % g.groups_per_seq) * g.QT;
      qt1 = qt0 + g.QT < p.tiles_per_seq ? qt0 + g.QT : p.tiles_per_seq;
    }
  };

  if (warp < 4) {
    // ============================ softmax / correction / epilogue ============================
    const int r = tid;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t n_s = 0, n_o = 0;   // completed waits on s_full / o_full (parity = count & 1)
    for (int64_t item = blockIdx.x; item < g.items; item += gridDim.x) {
      int64_t seq0; int qt0, qt1, h;
      decode(item, seq0, qt0, qt1, h);
      for (int qt = qt0; qt < qt1; ++qt) {
        const int grp = (p.G > 1) ? r / p.Lq : 0;
        const int qtok = (p.G > 1) ? r - grp * p.Lq : qt * 128 + r;
        const int64_t qseq = seq0 + grp;
        const bool q_valid = (grp < p.G) && (qseq < p.num_seqs) && (qtok < p.Lq);
        int64_t q_row = 0;
        int key_lo = 0x7fffffff, key_hi = 0;
        if (q_valid) {
          const int64_t b = qseq / p.seqs_per_batch, j = qseq % p.seqs_per_batch;
          q_row = b * p.q_bs + j * p.q_ss + (int64_t)qtok * p.q_ts;
          const int len = p.kv_lens ? p.kv_lens[qseq] : p.Lk;
          key_lo = grp * p.Lk;
          key_hi = key_lo + (len < p.Lk ? len : p.Lk);
        }
        float m = -INFINITY, l = 0.f;
        for (int jb = 0; jb < g.NKB; ++jb) {
          const int k0 = jb * g.BK;  // first key slot of this block
          mbar_wait(s_full, n_s & 1); ++n_s;
          tc_fence_after();
          // ---- block maximum over this row's valid keys ----
          float mb = -INFINITY;
          for (int c0 = 0; c0 < g.BK; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + c0, v);
            tmem_ld_wait();
            const int ka = k0 + c0;
            if (ka >= key_lo && ka + 32 <= key_hi && c0 + 32 <= g.BK) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mb = fmaxf(mb, __uint_as_float(v[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (ka + j >= key_lo && ka + j < key_hi && c0 + j < g.BK) mb = fmaxf(mb, __uint_as_float(v[j]));
            }
          }
          const float m_new = fmaxf(m, mb);
          const float ms = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
          // ---- rescale the running O / l when this row's maximum grew (needs the previous PV finished) ----
          if (jb > 0) {
            mbar_wait(o_full, n_o & 1); ++n_o;
            tc_fence_after();
            const float alpha = (m == -INFINITY) ? 1.f : fast_exp2(m * p.scale_log2 - ms);
            if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
              for (int c = 0; c < Cfg::DP; c += 8) {
                uint32_t o[8];
                tmem_ld_32x32b_x8(t_row + g.o_col + c, o);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                tmem_st_32x32b_x8(t_row + g.o_col + c, o);
              }
              tmem_st_wait();
            }
            l *= alpha;
          }
          // ---- P = exp2(S*scale - max), row sum, store P (bf16) ----
          for (int c0 = 0; c0 < g.BK; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + c0, v);
            tmem_ld_wait();
            float pr[32];
            const int ka = k0 + c0;
            if (ka >= key_lo && ka + 32 <= key_hi && c0 + 32 <= g.BK) {
              float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;  // four chains: the adds do not serialise
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                pr[j] = fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - ms); l0 += pr[j];
                pr[j + 1] = fast_exp2(__uint_as_float(v[j + 1]) * p.scale_log2 - ms); l1 += pr[j + 1];
                pr[j + 2] = fast_exp2(__uint_as_float(v[j + 2]) * p.scale_log2 - ms); l2 += pr[j + 2];
                pr[j + 3] = fast_exp2(__uint_as_float(v[j + 3]) * p.scale_log2 - ms); l3 += pr[j + 3];
              }
              l += (l0 + l1) + (l2 + l3);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const bool ok = ka + j >= key_lo && ka + j < key_hi && c0 + j < g.BK;
                pr[j] = ok ? fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - ms) : 0.f;
                l += pr[j];
              }
            }
            if constexpr (kPTmem) {
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(pr[2 * j], pr[2 * j + 1]);
              tmem_st_32x32b_x16(t_row + (c0 >> 1), pk);   // in place: P columns trail the S columns already read
            } else {
#pragma unroll
              for (int u4 = 0; u4 < 4; ++u4) {
                const int u = (c0 >> 3) + u4;
                *reinterpret_cast<uint4*>(sP + (u >> 3) * (128 * 128) + sw128_off(r, u & 7)) = pack8(pr + u4 * 8);
              }
            }
          }
          m = m_new;
          if constexpr (kPTmem) tmem_st_wait(); else fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(p_full);
        }
        // ---- epilogue: wait for the last PV, normalise, store ----
        mbar_wait(o_full, n_o & 1); ++n_o;
        tc_fence_after();
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        __nv_bfloat16* orow = p.out + q_row * p.out_ld + (int64_t)h * D;
#pragma unroll 1
        for (int u = 0; u < U; ++u) {
          uint32_t v[8];
          tmem_ld_32x32b_x8(t_row + g.o_col + u * 8, v);
          tmem_ld_wait();
          if (q_valid) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[e]) * inv;
            *reinterpret_cast<uint4*>(orow + u * 8) = pack8(o);
          }
        }
        tc_fence_before();  // O / S regions are overwritten by later MMAs only after our next p_full arrivals
      }
    }
  } else if (warp < 8) {
    // ============================ loaders: Q per q-tile, K / V per key block (once per item when resident) ======
    const int lt = tid - 128;
    uint32_t n_q = 0, n_kv[3] = {0, 0, 0};
    for (int64_t item = blockIdx.x; item < g.items; item += gridDim.x) {
      int64_t seq0; int qt0, qt1, h;
      decode(item, seq0, qt0, qt1, h);
      for (int qt = qt0; qt < qt1; ++qt) {
        {  // ---- Q row `lt` of q-tile qt ----
          const int grp = (p.G > 1) ? lt / p.Lq : 0;
          const int qtok = (p.G > 1) ? lt - grp * p.Lq : qt * 128 + lt;
          const int64_t qseq = seq0 + grp;
          const bool q_valid = (grp < p.G) && (qseq < p.num_seqs) && (qtok < p.Lq);
          uint4 t[UP];
#pragma unroll
          for (int u = 0; u < U; ++u) t[u] = make_uint4(0, 0, 0, 0);
          if (q_valid) {
            const int64_t b = qseq / p.seqs_per_batch, j = qseq % p.seqs_per_batch;
            const int64_t q_row = b * p.q_bs + j * p.q_ss + (int64_t)qtok * p.q_ts;
            const uint4* qs = reinterpret_cast<const uint4*>(p.q + q_row * p.q_ld + (int64_t)h * D);
#pragma unroll
            for (int u = 0; u < U; ++u) t[u] = __ldg(qs + u);
          }
          float rq = 1.f;
          if (p.qw != nullptr) {
            float ss = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) ss += sumsq8(t[u]);
            rq = rsqrtf(ss * (1.0f / D) + p.eps);
          }
          const __nv_bfloat16* qwt = (p.qw2 != nullptr && qtok >= p.norm_split) ? p.qw2 : p.qw;
          mbar_wait(q_empty, (n_q & 1) ^ 1);   // the previous q-tile's S MMAs are done with sQ
          if ((U % 2 == 0) && p.rope_half)
            finish_and_store_row_half<D>(t, rq, qwt, p.cos + (int64_t)qtok * (D / 2), p.sin + (int64_t)qtok * (D / 2), sQ,
                                         128 * 128, sQt, lt);
          else
            finish_and_store_units<D, 0, UP>(t, rq, qwt, p.cos ? p.cos + (int64_t)qtok * (D / 2) : nullptr,
                                             p.sin ? p.sin + (int64_t)qtok * (D / 2) : nullptr, sQ, 128 * 128, sQt, lt);
          fence_proxy_async_smem();
          mbar_arrive(q_full);
          ++n_q;
        }
        if (g.resident && qt != qt0) continue;   // key blocks of this item are already in the ring
        for (int jb = 0; jb < g.NKB; ++jb) {
          const int st = g.resident ? jb : (jb & 1);
          // K/V slot handled by this thread inside block jb (BK <= 128 keys: one slot per loader thread)
          const int slot = jb * g.BK + lt;
          const bool in_blk = lt < g.BK;
          const int kg = slot / p.Lk, ktok = slot - kg * p.Lk;
          const int64_t kseq = seq0 + kg;
          const bool k_valid = in_blk && (slot < p.NK) && (kseq < p.num_seqs);
          uint4 tk[UP], tv[UP];
#pragma unroll
          for (int u = 0; u < U; ++u) { tk[u] = make_uint4(0, 0, 0, 0); tv[u] = make_uint4(0, 0, 0, 0); }
          if (k_valid) {
            const int64_t b = kseq / p.seqs_per_batch, j = kseq % p.seqs_per_batch;
            const int64_t k_row = b * p.k_bs + j * p.k_ss + (int64_t)ktok * p.k_ts;
            const uint4* ks = reinterpret_cast<const uint4*>(p.k + k_row * p.k_ld + (int64_t)h * D);
            const uint4* vs = reinterpret_cast<const uint4*>(p.v + k_row * p.v_ld + (int64_t)h * D);
#pragma unroll
            for (int u = 0; u < U; ++u) tk[u] = __ldg(ks + u);
#pragma unroll
            for (int u = 0; u < U; ++u) tv[u] = __ldg(vs + u);
          }
          float rk = 1.f;
          if (p.kw != nullptr) {
            float ss = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) ss += sumsq8(tk[u]);
            rk = rsqrtf(ss * (1.0f / D) + p.eps);
          }
          mbar_wait(kv_empty(st), (n_kv[st] & 1) ^ 1);   // every MMA that read this stage before has completed
          if (in_blk) {
            const __nv_bfloat16* kwt = (p.kw2 != nullptr && ktok >= p.norm_split) ? p.kw2 : p.kw;
            if ((U % 2 == 0) && p.rope_half)
              finish_and_store_row_half<D>(tk, rk, kwt, p.cos + (int64_t)ktok * (D / 2), p.sin + (int64_t)ktok * (D / 2),
                                           sK(st), g.BK * 128, sKt(st), lt);
            else
              finish_and_store_units<D, 0, UP>(tk, rk, kwt, p.cos ? p.cos + (int64_t)ktok * (D / 2) : nullptr,
                                               p.sin ? p.sin + (int64_t)ktok * (D / 2) : nullptr, sK(st), g.BK * 128, sKt(st), lt);
            // V keeps its [key][d] orientation (MN-major B operand): same tile layout as K, no arithmetic
            finish_and_store_units<D, 0, UP>(tv, 1.f, nullptr, nullptr, nullptr, sV(st), g.BK * 128, sVt(st), lt);
          }
          fence_proxy_async_smem();
          mbar_arrive(kv_full(st));
          ++n_kv[st];
        }
      }
    }
  } else {
    // ============================ tcgen05 issuer ============================
    // the whole warp runs the control flow (waits / descriptor arithmetic warp-uniform: operands in uniform registers,
    // MMAs issue back to back); one elected lane issues
    {
      const bool leader = elect_one();
      uint32_t n_q = 0, n_p = 0, n_kv[3] = {0, 0, 0};
      const uint32_t idesc_s = make_idesc_bf16_f32(128, g.BK);
      const uint32_t idesc_om = make_idesc_bf16_f32_bmn(128, NMAIN);
      const uint32_t idesc_ot = make_idesc_bf16_f32_bmn(128, 16);
      const uint32_t kchunk16 = (uint32_t)(g.BK * 128) >> 4;
      const int steps = g.BK / 16;
      for (int64_t item = blockIdx.x; item < g.items; item += gridDim.x) {
        int64_t seq0; int qt0, qt1, h;
        decode(item, seq0, qt0, qt1, h);
        for (int qt = qt0; qt < qt1; ++qt) {
          mbar_wait(q_full, n_q & 1); ++n_q;
          for (int jb = 0; jb < g.NKB; ++jb) {
            const int st = g.resident ? jb : (jb & 1);
            if (!g.resident || qt == qt0) { mbar_wait(kv_full(st), n_kv[st] & 1); ++n_kv[st]; }
            tc_fence_after();
            // S = Q K^T  (overwrites the S/P region: ordered after the previous PV by the in-order MMA pipe)
            const uint64_t qd0 = make_sw128_kmajor_desc(smem_u32(sQ)), kd0 = make_sw128_kmajor_desc(smem_u32(sK(st)));
            const uint64_t qt0d = make_noswz_kmajor_desc(smem_u32(sQt)), kt0d = make_noswz_kmajor_desc(smem_u32(sKt(st)));
            if (leader) {
              uint32_t acc = 0;
#pragma unroll
              for (int kc = 0; kc < Cfg::MAIN; ++kc) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  umma_bf16<1>(tmem_base, qd0 + (uint64_t)(kc * ((128 * 128) >> 4) + ks * 2), kd0 + (uint64_t)(kc * kchunk16 + ks * 2),
                               idesc_s, acc);
                  acc = 1;
                }
              }
              if (Cfg::TAIL) umma_bf16<1>(tmem_base, qt0d, kt0d, idesc_s, acc);
              umma_commit<1>(s_full);
              if (jb == g.NKB - 1) umma_commit<1>(q_empty);   // sQ may be restaged for the next q-tile
            }
            __syncwarp();
            mbar_wait(p_full, n_p & 1); ++n_p;
            tc_fence_after();
            // O (+)= P V : per 16 keys, one MMA over the swizzled main chunk(s) (N = 64 / 128) and one over the
            // no-swizzle head-dim tail (N = 16); V is read as an MN-major operand (no transposition anywhere)
            const uint64_t vd0 = make_sw128_mnmajor_desc(smem_u32(sV(st)), (uint32_t)(g.BK * 128));
            const uint64_t vt0 = make_noswz_mnmajor_desc(smem_u32(sVt(st)));
            const uint64_t pd0 = make_sw128_kmajor_desc(smem_u32(sP));
            if (leader) {
              for (int s = 0; s < steps; ++s) {
                const uint32_t accu = (jb > 0 || s > 0) ? 1u : 0u;
                const uint64_t dbm = vd0 + (uint64_t)(s * (2048 >> 4));
                const uint64_t dbt = vt0 + (uint64_t)(s * (512 >> 4));
                if constexpr (kPTmem) {
                  umma_bf16_ts(tmem_base + g.o_col, tmem_base + s * 8, dbm, idesc_om, accu);
                  if (Cfg::TAIL) umma_bf16_ts(tmem_base + g.o_col + NMAIN, tmem_base + s * 8, dbt, idesc_ot, accu);
                } else {
                  const uint64_t da = pd0 + (uint64_t)((s >> 2) * ((128 * 128) >> 4) + (s & 3) * 2);
                  umma_bf16<1>(tmem_base + g.o_col, da, dbm, idesc_om, accu);
                  if (Cfg::TAIL) umma_bf16<1>(tmem_base + g.o_col + NMAIN, da, dbt, idesc_ot, accu);
                }
              }
              if (!g.resident || qt == qt1 - 1) umma_commit<1>(kv_empty(st));   // stage free once its last reader is done
              umma_commit<1>(o_full);
            }
            __syncwarp();
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, (uint32_t)g.tmem_cols);
  }
}

template <int D, bool kPTmem>
static int attn_flash_launch(AttnParams& p, int H, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  FlashGeom g = {};
  auto up1k = [](int x) { return (x + 1023) / 1024 * 1024; };
  g.NKB = (p.NK + 127) / 128;
  g.BK = (((p.NK + g.NKB - 1) / g.NKB) + 15) / 16 * 16;
  const int nkc = (g.BK + 63) / 64;
  g.k_bytes = up1k(Cfg::MAIN * g.BK * 128);
  g.kt_bytes = up1k(Cfg::TAIL ? g.BK * 32 : 0);
  g.v_bytes = g.k_bytes;
  g.kv_stage_bytes = 2 * (g.k_bytes + g.kt_bytes);
  g.resident = (g.NKB <= 3) ? 1 : 0;
  g.NST = g.resident ? g.NKB : 2;
  const int p_bytes = kPTmem ? 0 : nkc * 128 * 128;
  int off = Cfg::MAIN * 128 * 128;
  g.off_qt = off; off += up1k(Cfg::TAIL ? 128 * 32 : 0);
  g.off_kv = off; off += g.NST * g.kv_stage_bytes;
  g.off_p = off; off += p_bytes;
  g.off_bar = off; off += 128;
  const int smem = off;
  const int s_cols = (g.BK + 31) / 32 * 32;
  g.o_col = s_cols;
  g.tmem_cols = (s_cols + Cfg::DP <= 256) ? 256 : 512;
  if (smem > 227 * 1024) { set_error("osb_attn_short(flash): %d B smem", smem); return OSB_ERR_UNSUPPORTED; }
  const int per_sm = (D <= 72 && smem <= 112 * 1024 && g.tmem_cols == 256) ? 2 : 1;
  const int64_t slots = (int64_t)sm_count() * per_sm;
  // work units: q-tiles grouped QT at a time when the key set is resident (amortises its staging)
  g.QT = 1;
  g.groups_per_seq = p.tiles_per_seq;
  if (p.G > 1) {
    g.units = (p.num_seqs + p.G - 1) / p.G;
  } else {
    if (g.resident) {
      const double staging = 0.7 * g.NKB;  // staging one key block ~ 0.7 q-tile-blocks of softmax work
      double best = 1e30;
      for (int cand = 1; cand <= p.tiles_per_seq && cand <= 64; ++cand) {
        const int64_t groups = (p.tiles_per_seq + cand - 1) / cand;
        const int64_t ctas = p.num_seqs * groups * H;
        const int64_t waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * ((double)((p.tiles_per_seq + groups - 1) / groups) * g.NKB + staging);
        if (cost < best - 1e-9) { best = cost; g.QT = cand; }
      }
      g.groups_per_seq = (p.tiles_per_seq + g.QT - 1) / g.QT;
    }
    g.units = p.num_seqs * g.groups_per_seq;
  }
  g.items = g.units * H;
  int64_t grid = slots < g.items ? slots : g.items;
  cudaLaunchAttribute attr[2];
  cudaLaunchConfig_t cfg = launch_config(dim3((unsigned)grid), dim3(kFlashThreads), smem, stream, attr);
  OSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_flash_kernel<D, kPTmem>, p, g));
  count_launch();
  return OSB_OK;
}

// One head row that cp.async has already placed in its K-major operand tile: optional RMSNorm scale and interleaved
// RoPE, in place.  Two sweeps over shared memory (sum of squares, then one 16-byte unit at a time) instead of holding
// the row in registers: the loader warps run with a small register budget so that many of them fit next to the
// softmax warpgroups.  Returns the squared length of the finished vector (before bf16 rounding).
template <int D>
__device__ __forceinline__ float finish_row_inplace(uint8_t* main_base, int main_chunk_bytes, uint8_t* tail_base, int row,
                                                    const __nv_bfloat16* w, float eps, const float* cosr, const float* sinr) {
  using Cfg = AttnCfg<D>;
  auto unit_ptr = [&](int u) -> uint4* {
    return (u < Cfg::MAIN * 8) ? reinterpret_cast<uint4*>(main_base + (u >> 3) * main_chunk_bytes + sw128_off(row, u & 7))
                               : reinterpret_cast<uint4*>(tail_base + tail_off(row, u - Cfg::MAIN * 8));
  };
  float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
  for (int u = 0; u < Cfg::U; ++u) { const float q = sumsq8(*unit_ptr(u)); if (u & 1) ss1 += q; else ss0 += q; }
  const float ss = ss0 + ss1;
  if (w == nullptr && cosr == nullptr) return ss;
  const float r = (w != nullptr) ? rsqrtf(ss * (1.0f / D) + eps) : 1.f;
  float n0 = 0.f, n1 = 0.f;
#pragma unroll
  for (int u = 0; u < Cfg::U; ++u) {
    uint4* ptr = unit_ptr(u);
    float xu[8];
    unpack8(*ptr, xu);
    if (w != nullptr) {
      float wf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + u), wf);
#pragma unroll
      for (int e = 0; e < 8; ++e) xu[e] *= r * wf[e];
    }
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) q += xu[e] * xu[e];
    if (u & 1) n1 += q; else n0 += q;
    if (cosr != nullptr) {
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(cosr) + u);
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(sinr) + u);
      const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = xu[2 * i], b = xu[2 * i + 1];
        xu[2 * i] = a * cc[i] - b * sn[i];
        xu[2 * i + 1] = b * cc[i] + a * sn[i];
      }
    }
    *ptr = pack8(xu);
  }
  return n0 + n1;
}

// ------------------------------------------------------------------------------------------------------------------
// Ping-pong kernel for resident key sets (<= 2 blocks of <= 160 keys): STDiT3 spatial / temporal / T5 cross attention.
//
// One CTA per SM walks a contiguous range of "jobs" (one 128-row q-tile against its key set).  Consecutive jobs
// alternate between two slots; each slot has its own Q tile in shared memory, its own S/P and O accumulators in TMEM
// and its own softmax warpgroup, so the tensor pipe works on one slot's S = Q K^T / O += P V while the other slot's
// warpgroup is in its exp2 pass (the MUFU-bound part).  Key sets (K, V blocks) live in a ring of stages shared by both
// slots: the two q-tiles of a spatial sequence and the 128 q-tiles of a cross-attention head reuse one staged set.
//
// Loads are register-free: the loader warps (one group for Q tiles, one for K/V blocks) issue cp.async (16 B, zero-fill for padding rows) straight into the final
// swizzled operand position, one task ahead of the one being finished; rows that need RMSNorm / RoPE are then read
// back by their owner thread, transformed in fp32 and written in place.  V (and q/k without norm) are pure copies.
// ------------------------------------------------------------------------------------------------------------------
#ifdef OSB_PP_TRACE
// debug build only (tests/pp_trace.py): CTA 0 records (tag, clock64) per role: 0/1 softmax slots, 2 loader, 3 issuer
__device__ unsigned long long g_pp_trace[4][512];
__device__ int g_pp_trace_n[4];
#define PP_TR(role, tag)                                                                              \
  do {                                                                                                \
    if (blockIdx.x == 0) {                                                                            \
      const int n_ = g_pp_trace_n[role];                                                              \
      if (n_ < 256) { g_pp_trace[role][2 * n_] = (tag); g_pp_trace[role][2 * n_ + 1] = clock64(); g_pp_trace_n[role] = n_ + 1; } \
    }                                                                                                 \
  } while (0)
#else
#define PP_TR(role, tag) do { } while (0)
#endif
constexpr int kPPThreads = 800;   // warps 0-3 / 4-7 softmax of slot 0 / 1, 8-11 / 12-15 Q loaders of slot 0 / 1,
constexpr int kPPIssuerWarp = 24;   // 16-23 K/V loaders, 24 tcgen05 issuer (25 warps: 80 registers per thread; the
constexpr int kPPKvThreads = 256;   // row transforms are latency bound, so they get thread-level parallelism)
constexpr int kPPKvWarps = kPPKvThreads / 32;
constexpr int kPPMaxStages = 4;

struct PPGeom {
  int32_t BK, NKB, NSETS;        // keys per block (multiple of 16, <= 160), blocks per key set, key sets in the ring
  int32_t q_bytes, qt_off;       // per-slot Q buffer size; offset of its head-dim tail tile
  int32_t k_bytes, kt_bytes;     // per stage: K main, K tail (V main / tail have the same sizes)
  int32_t stage_bytes, off_kv, off_bar;
  int32_t s_cols, o_col, o_cols; // TMEM: slot s has S/P at s * s_cols and O at o_col + s * o_cols
  int32_t tps;                   // jobs (q-tiles) per key set
  int64_t units;                 // key sets per head
  int64_t num_sets;              // units * heads
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;   // src-size 0: the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D>
__global__ void __launch_bounds__(kPPThreads, 1) attn_pp_kernel(const AttnParams p, const PPGeom g) {
  using Cfg = AttnCfg<D>;
  constexpr int U = Cfg::U;
  constexpr int NMAIN = Cfg::MAIN * 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  auto sQ = [&](int s) { return smem + s * g.q_bytes; };
  auto sQt = [&](int s) { return smem + s * g.q_bytes + g.qt_off; };
  auto sK = [&](int st) { return smem + g.off_kv + st * g.stage_bytes; };
  auto sKt = [&](int st) { return smem + g.off_kv + st * g.stage_bytes + g.k_bytes; };
  auto sV = [&](int st) { return smem + g.off_kv + st * g.stage_bytes + g.k_bytes + g.kt_bytes; };
  auto sVt = [&](int st) { return smem + g.off_kv + st * g.stage_bytes + 2 * g.k_bytes + g.kt_bytes; };
  const uint32_t bar0 = smem_u32(smem + g.off_bar);
  auto q_full = [&](int s) { return bar0 + 8u * s; };
  auto q_empty = [&](int s) { return bar0 + 16u + 8u * s; };
  auto s_full = [&](int s) { return bar0 + 32u + 8u * s; };
  auto p_full = [&](int s) { return bar0 + 48u + 8u * s; };
  auto o_full = [&](int s) { return bar0 + 64u + 8u * s; };
  auto kv_full = [&](int st) { return bar0 + 80u + 8u * st; };
  auto kv_empty = [&](int st) { return bar0 + 112u + 8u * st; };
  const uint32_t tmem_slot = bar0 + 144u;
  // squared lengths of the staged (normalised) vectors: |q_row|^2 per slot and row, max |k_row|^2 per stage and loader
  // warp.  |q||k| bounds every logit of the row (Cauchy-Schwarz), which lets the softmax skip its max pass.
  float* qn2 = reinterpret_cast<float*>(smem + g.off_bar + 256);       // [2 slots][2 (job parity)][128]
  float* kmax2 = qn2 + 512;                                              // [kPPMaxStages][kPPKvWarps]

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == kPPIssuerWarp) {
    if ((tid & 31) == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(q_full(s), 128); mbar_init(q_empty(s), 1); mbar_init(s_full(s), 1); mbar_init(p_full(s), 128);
        mbar_init(o_full(s), 1);
      }
      for (int st = 0; st < kPPMaxStages; ++st) { mbar_init(kv_full(st), kPPKvThreads); mbar_init(kv_empty(st), 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512u);
  }
  if (Cfg::TAIL) {
    // head-dim tail tiles: the zero-pad unit (columns D..DP-1) is written once here; copies only ever fill unit 0
    const int nst = g.NSETS * g.NKB;
    for (int i = tid; i < 2 * 256; i += kPPThreads)   // two Q tail tiles of 128 rows x 32 B
      *reinterpret_cast<uint4*>(sQt(i >> 8) + (i & 255) * 16) = make_uint4(0, 0, 0, 0);
    const int units_t = g.kt_bytes / 16;
    for (int i = tid; i < nst * 2 * units_t; i += kPPThreads) {
      const int st = i / (2 * units_t), rem = i - st * 2 * units_t;
      uint8_t* base = rem < units_t ? sKt(st) : sVt(st);
      *reinterpret_cast<uint4*>(base + (rem % units_t) * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  pdl_wait();
  pdl_launch_dependents();

  // ---- this CTA's contiguous job range; jobs of one key set are never split when a set has few q-tiles ----------
  int64_t J0, J1;
  if (g.tps <= 4) {
    J0 = ((int64_t)blockIdx.x * g.num_sets / gridDim.x) * g.tps;
    J1 = ((int64_t)(blockIdx.x + 1) * g.num_sets / gridDim.x) * g.tps;
  } else {
    const int64_t total = g.num_sets * g.tps;
    J0 = (int64_t)blockIdx.x * total / gridDim.x;
    J1 = (int64_t)(blockIdx.x + 1) * total / gridDim.x;
  }
  const int nj = (int)(J1 - J0);
  auto decode = [&](int64_t J, int64_t& seq0, int& qt, int& h) {
    const int64_t set = J / g.tps;
    qt = (int)(J - set * g.tps);
    h = (int)(set / g.units);
    const int64_t unit = set - (int64_t)h * g.units;
    seq0 = (p.G > 1) ? unit * p.G : unit;
  };
  auto first_of_set = [&](int i) { return i == 0 || ((J0 + i) % g.tps) == 0; };
  auto last_of_set = [&](int i) { return i == nj - 1 || ((J0 + i + 1) % g.tps) == 0; };

  if (warp < 8) {
    // ============================ softmax / correction / epilogue of one slot ============================
    const int slot = warp >> 2;
    const int r = tid & 127;
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_s = t_row + (uint32_t)(slot * g.s_cols);
    const uint32_t t_o = t_row + (uint32_t)(g.o_col + slot * g.o_cols);
    const int nch = g.BK >> 4;   // 16-column chunks (BK is a multiple of 16)
    uint32_t n_s = 0, n_o = 0;
    int64_t set_ord = -1;
    for (int i = 0; i < nj; ++i) {
      if (first_of_set(i)) ++set_ord;
      if ((i & 1) != slot) continue;
      int64_t seq0; int qt, h;
      decode(J0 + i, seq0, qt, h);
      const int grp = (p.G > 1) ? r / p.Lq : 0;
      const int qtok = (p.G > 1) ? r - grp * p.Lq : qt * 128 + r;
      const int64_t qseq = seq0 + grp;
      const bool q_valid = (grp < p.G) && (qseq < p.num_seqs) && (qtok < p.Lq);
      int64_t q_row = 0;
      int key_lo = 0x7fffffff, key_hi = 0;
      if (q_valid) {
        const int64_t b = qseq / p.seqs_per_batch, j = qseq % p.seqs_per_batch;
        q_row = b * p.q_bs + j * p.q_ss + (int64_t)qtok * p.q_ts;
        const int len = p.kv_lens ? p.kv_lens[qseq] : p.Lk;
        key_lo = grp * p.Lk;
        key_hi = key_lo + (len < p.Lk ? len : p.Lk);
      }
      // warp-wide key range: it guards the .sync.aligned TMEM loads (rows of one warp can belong to different packed
      // sequences), the per-lane range only guards arithmetic
      const int wkey_lo = __reduce_min_sync(0xffffffffu, key_lo), wkey_hi = __reduce_max_sync(0xffffffffu, key_hi);
      float m = -INFINITY, l = 0.f, shift = 0.f;
      bool onepass = false;
      for (int jb = 0; jb < g.NKB; ++jb) {
        const int k0 = jb * g.BK;
        mbar_wait(s_full(slot), n_s & 1); ++n_s;
        tc_fence_after();
        if (r == 0) PP_TR(slot, 10);
        if (jb == 0) {
          // logit bound |q_r| * max|k| * scale (log2 domain): small enough -> it replaces the row maximum, exactly
          // (softmax is shift invariant; 2^-2B stays a normal fp32 / bf16 number for B <= 60)
          float km = 0.f;
          const int st0 = (int)(set_ord % g.NSETS) * g.NKB;
          for (int b2 = 0; b2 < g.NKB; ++b2) {
#pragma unroll
            for (int w4 = 0; w4 < kPPKvWarps / 4; ++w4) {
              const float4 k4 = *reinterpret_cast<const float4*>(kmax2 + (st0 + b2) * kPPKvWarps + w4 * 4);
              km = fmaxf(km, fmaxf(fmaxf(k4.x, k4.y), fmaxf(k4.z, k4.w)));
            }
          }
          const float bound = sqrtf(qn2[(slot * 2 + ((i >> 1) & 1)) * 128 + r] * km) * fabsf(p.scale_log2);
          onepass = !__any_sync(0xffffffffu, !(bound <= 60.f));
          shift = bound;
        }
        if (onepass) {
          // ---- single pass: P = exp2(S * scale - bound) and its row sum, 16 columns per TMEM load ----
          const float ms = shift;
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
          for (int c0 = 0; c0 < g.BK; c0 += 16) {
            const int ka = k0 + c0;
            uint32_t pk[8];
            if (ka + 16 <= wkey_lo || ka >= wkey_hi) {
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = 0u;
            } else {
              uint32_t v[16];
              tmem_ld_32x32b_x16(t_s + c0, v);
              tmem_ld_wait();
              if (ka >= key_lo && ka + 16 <= key_hi) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const float e0 = fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - ms);
                  const float e1 = fast_exp2(__uint_as_float(v[j + 1]) * p.scale_log2 - ms);
                  const float e2 = fast_exp2(__uint_as_float(v[j + 2]) * p.scale_log2 - ms);
                  const float e3 = fast_exp2(__uint_as_float(v[j + 3]) * p.scale_log2 - ms);
                  l0 += e0; l1 += e1; l2 += e2; l3 += e3;
                  pk[j >> 1] = pack_bf16x2(e0, e1);
                  pk[(j >> 1) + 1] = pack_bf16x2(e2, e3);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  const bool ok0 = ka + j >= key_lo && ka + j < key_hi;
                  const bool ok1 = ka + j + 1 >= key_lo && ka + j + 1 < key_hi;
                  const float e0 = ok0 ? fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - ms) : 0.f;
                  const float e1 = ok1 ? fast_exp2(__uint_as_float(v[j + 1]) * p.scale_log2 - ms) : 0.f;
                  l0 += e0; l1 += e1;
                  pk[j >> 1] = pack_bf16x2(e0, e1);
                }
              }
            }
            tmem_st_32x32b_x8(t_s + (c0 >> 1), pk);
          }
          l += (l0 + l1) + (l2 + l3);
          // stay in lockstep with the P V of the previous block (no rescale needed): a waiter may lag an mbarrier by at
          // most one phase, or its parity test aliases to the phase after next
          if (jb > 0) { mbar_wait(o_full(slot), n_o & 1); ++n_o; }
        } else {
          // ---- two passes with an online maximum (logit bound too large to be used as the shift) ----
          auto active = [&](int c) { const int ka = k0 + c * 16; return !(ka + 16 <= wkey_lo || ka >= wkey_hi); };
          auto full = [&](int c) { const int ka = k0 + c * 16; return ka >= key_lo && ka + 16 <= key_hi; };
          uint32_t va[16], vb[16];
          float mb0 = -INFINITY, mb1 = -INFINITY, mb2 = -INFINITY, mb3 = -INFINITY;
          auto max_chunk = [&](const uint32_t (&v)[16], int c) {
            if (full(c)) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                mb0 = fmaxf(mb0, __uint_as_float(v[j])); mb1 = fmaxf(mb1, __uint_as_float(v[j + 1]));
                mb2 = fmaxf(mb2, __uint_as_float(v[j + 2])); mb3 = fmaxf(mb3, __uint_as_float(v[j + 3]));
              }
            } else {
              const int ka = k0 + c * 16;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (ka + j >= key_lo && ka + j < key_hi) mb0 = fmaxf(mb0, __uint_as_float(v[j]));
            }
          };
          if (active(0)) tmem_ld_32x32b_x16(t_s, va);
          tmem_ld_wait();
          for (int c = 0; c < nch; c += 2) {
            if (c + 1 < nch && active(c + 1)) tmem_ld_32x32b_x16(t_s + (c + 1) * 16, vb);
            if (active(c)) max_chunk(va, c);
            tmem_ld_wait();
            if (c + 1 < nch) {
              if (c + 2 < nch && active(c + 2)) tmem_ld_32x32b_x16(t_s + (c + 2) * 16, va);
              if (active(c + 1)) max_chunk(vb, c + 1);
              tmem_ld_wait();
            }
          }
          const float m_new = fmaxf(m, fmaxf(fmaxf(mb0, mb1), fmaxf(mb2, mb3)));
          const float ms = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
          if (jb > 0) {   // rescale the running O / l when the maximum grew (the previous P V must have landed)
            mbar_wait(o_full(slot), n_o & 1); ++n_o;
            tc_fence_after();
            const float alpha = (m == -INFINITY) ? 1.f : fast_exp2(m * p.scale_log2 - ms);
            if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
              for (int c = 0; c < Cfg::DP; c += 16) {
                uint32_t o[16];
                tmem_ld_32x32b_x16(t_o + c, o);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                tmem_st_32x32b_x16(t_o + c, o);
              }
              tmem_st_wait();
            }
            l *= alpha;
          }
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
          auto exp_chunk = [&](const uint32_t (&v)[16], int c) {
            uint32_t pk[8];
            if (!active(c)) {
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = 0u;
            } else if (full(c)) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float e0 = fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - ms);
                const float e1 = fast_exp2(__uint_as_float(v[j + 1]) * p.scale_log2 - ms);
                const float e2 = fast_exp2(__uint_as_float(v[j + 2]) * p.scale_log2 - ms);
                const float e3 = fast_exp2(__uint_as_float(v[j + 3]) * p.scale_log2 - ms);
                l0 += e0; l1 += e1; l2 += e2; l3 += e3;
                pk[j >> 1] = pack_bf16x2(e0, e1);
                pk[(j >> 1) + 1] = pack_bf16x2(e2, e3);
              }
            } else {
              const int ka = k0 + c * 16;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const bool ok0 = ka + j >= key_lo && ka + j < key_hi;
                const bool ok1 = ka + j + 1 >= key_lo && ka + j + 1 < key_hi;
                const float e0 = ok0 ? fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - ms) : 0.f;
                const float e1 = ok1 ? fast_exp2(__uint_as_float(v[j + 1]) * p.scale_log2 - ms) : 0.f;
                l0 += e0; l1 += e1;
                pk[j >> 1] = pack_bf16x2(e0, e1);
              }
            }
            tmem_st_32x32b_x8(t_s + c * 8, pk);
          };
          if (active(0)) tmem_ld_32x32b_x16(t_s, va);
          tmem_ld_wait();
          for (int c = 0; c < nch; c += 2) {
            if (c + 1 < nch && active(c + 1)) tmem_ld_32x32b_x16(t_s + (c + 1) * 16, vb);
            exp_chunk(va, c);
            tmem_ld_wait();
            if (c + 1 < nch) {
              if (c + 2 < nch && active(c + 2)) tmem_ld_32x32b_x16(t_s + (c + 2) * 16, va);
              exp_chunk(vb, c + 1);
              tmem_ld_wait();
            }
          }
          l += (l0 + l1) + (l2 + l3);
          m = m_new;
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(slot));
        if (r == 0) PP_TR(slot, 13);
      }
      // ---- epilogue: wait for the last P V ----
      mbar_wait(o_full(slot), n_o & 1); ++n_o;
      tc_fence_after();
      if (r == 0) PP_TR(slot, 14);
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      __nv_bfloat16* orow = p.out + q_row * p.out_ld + (int64_t)h * D;
      {
        // O row: MAIN*64 + tail columns, read in 16-column pieces
        uint32_t v[16];
#pragma unroll
        for (int c = 0; c < Cfg::MAIN * 64; c += 16) {
          tmem_ld_32x32b_x16(t_o + c, v);
          tmem_ld_wait();
          if (q_valid) {
#pragma unroll
            for (int u2 = 0; u2 < 2; ++u2) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[u2 * 8 + e]) * inv;
              *reinterpret_cast<uint4*>(orow + c + u2 * 8) = pack8(o);
            }
          }
        }
        if (Cfg::TAIL) {
          uint32_t w8[8];
          tmem_ld_32x32b_x8(t_o + Cfg::MAIN * 64, w8);
          tmem_ld_wait();
          if (q_valid) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(w8[e]) * inv;
            *reinterpret_cast<uint4*>(orow + Cfg::MAIN * 64) = pack8(o);
          }
        }
      }
      if (r == 0) PP_TR(slot, 15);
      tc_fence_before();  // this slot's O / S are overwritten by MMAs issued only after our next p_full arrival
    }
  } else if (warp < 16) {
    // ============================ Q loaders: warps 8-11 stage slot 0's tiles, warps 12-15 slot 1's ============================
    const int slot = (warp - 8) >> 2;
    const int lt = tid - 256 - slot * 128;   // tile row
    uint32_t n_fill = 0;
    for (int i = slot; i < nj; i += 2) {
      int64_t seq0; int qt, h;
      decode(J0 + i, seq0, qt, h);
      const int grp = (p.G > 1) ? lt / p.Lq : 0;
      const int qtok = (p.G > 1) ? lt - grp * p.Lq : qt * 128 + lt;
      const int64_t qseq = seq0 + grp;
      const bool ok = (grp < p.G) && (qseq < p.num_seqs) && (qtok < p.Lq);
      const __nv_bfloat16* src = p.q;
      if (ok) {
        const int64_t b = qseq / p.seqs_per_batch, j = qseq % p.seqs_per_batch;
        src = p.q + (b * p.q_bs + j * p.q_ss + (int64_t)qtok * p.q_ts) * p.q_ld + (int64_t)h * D;
      }
      mbar_wait(q_empty(slot), (n_fill & 1) ^ 1); ++n_fill;   // the previous tile's S MMAs are done with this buffer
      const uint32_t mb = smem_u32(sQ(slot)), tb = smem_u32(sQt(slot));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t dst = (u < Cfg::MAIN * 8) ? mb + (u >> 3) * (128 * 128) + sw128_off(lt, u & 7)
                                                 : tb + tail_off(lt, u - Cfg::MAIN * 8);
        cp_async16(dst, src + (ok ? u * 8 : 0), ok);
      }
      cp_async_commit();
      cp_async_wait<0>();
      if (lt == 0) PP_TR(2, 23);
      const __nv_bfloat16* qwt = (p.qw2 != nullptr && qtok >= p.norm_split) ? p.qw2 : p.qw;
      const int ctok = ok ? qtok : 0;   // padding rows: any valid table row (their values are zero)
      const float nrm2 = finish_row_inplace<D>(sQ(slot), 128 * 128, sQt(slot), lt, qwt, p.eps,
                                               p.cos ? p.cos + (int64_t)ctok * (D / 2) : nullptr,
                                               p.sin ? p.sin + (int64_t)ctok * (D / 2) : nullptr);
      // (double-buffered: with one key block per job the slot's next tile may be staged while this one's softmax starts)
      qn2[(slot * 2 + ((i >> 1) & 1)) * 128 + lt] = nrm2 * 1.02f;   // margin for the bf16 rounding of the staged values
      fence_proxy_async_smem();
      mbar_arrive(q_full(slot));
      if (lt == 0) PP_TR(2, 24);
    }
  } else if (warp < kPPIssuerWarp) {
    // ============================ K / V loaders: one task per key set (all its blocks) ============================
    const int lt = tid - 512;
    struct KTask { int ring, h; int64_t seq0; };
    int ki = 0;
    int64_t kord = -1;
    auto next_task = [&](KTask& t) -> bool {
      while (ki < nj && !first_of_set(ki)) ++ki;
      if (ki >= nj) return false;
      ++kord;
      int qt;
      decode(J0 + ki, t.seq0, qt, t.h);
      t.ring = (int)(kord % g.NSETS);
      ++ki;
      return true;
    };
    uint32_t n_fill[kPPMaxStages] = {0, 0, 0, 0};
    const int rows_set = g.NKB * g.BK;
    auto issue = [&](const KTask& t, bool blocking) -> bool {
      for (int jb = 0; jb < g.NKB; ++jb) {
        const int st = t.ring * g.NKB + jb;
        const uint32_t par = (n_fill[st] & 1) ^ 1;
        if (blocking) mbar_wait(kv_empty(st), par);
        else if (!mbar_test_wait(kv_empty(st), par)) return false;
      }
      for (int jb = 0; jb < g.NKB; ++jb) ++n_fill[t.ring * g.NKB + jb];
      for (int rs = lt; rs < rows_set; rs += kPPKvThreads) {
        const int jb = rs / g.BK, row = rs - jb * g.BK;
        const int st = t.ring * g.NKB + jb;
        const uint32_t kb = smem_u32(sK(st)), ktb = smem_u32(sKt(st));
        const uint32_t vb = smem_u32(sV(st)), vtb = smem_u32(sVt(st));
        const int kg = rs / p.Lk, ktok = rs - kg * p.Lk;
        const int64_t kseq = t.seq0 + kg;
        const bool ok = (rs < p.NK) && (kseq < p.num_seqs);
        const __nv_bfloat16 *ks = p.k, *vs = p.v;
        if (ok) {
          const int64_t b = kseq / p.seqs_per_batch, j = kseq % p.seqs_per_batch;
          const int64_t k_row = b * p.k_bs + j * p.k_ss + (int64_t)ktok * p.k_ts;
          ks = p.k + k_row * p.k_ld + (int64_t)t.h * D;
          vs = p.v + k_row * p.v_ld + (int64_t)t.h * D;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool main = u < Cfg::MAIN * 8;
          const uint32_t offm = (u >> 3) * (uint32_t)(g.BK * 128) + sw128_off(row, u & 7);
          const uint32_t offt = tail_off(row, u - Cfg::MAIN * 8);
          cp_async16(main ? kb + offm : ktb + offt, ks + (ok ? u * 8 : 0), ok);
          cp_async16(main ? vb + offm : vtb + offt, vs + (ok ? u * 8 : 0), ok);
        }
      }
      cp_async_commit();
      return true;
    };
    auto finish = [&](const KTask& t) {
      float kmx[2] = {0.f, 0.f};   // this thread's max |k_row|^2 per block
      for (int rs = lt; rs < rows_set; rs += kPPKvThreads) {
        const int jb = rs / g.BK, row = rs - jb * g.BK;
        const int st = t.ring * g.NKB + jb;
        const int kg = rs / p.Lk;
        const int ktok = rs - kg * p.Lk;
        const __nv_bfloat16* kwt = (p.kw2 != nullptr && ktok >= p.norm_split) ? p.kw2 : p.kw;
        const float nrm2 = finish_row_inplace<D>(sK(st), g.BK * 128, sKt(st), row, kwt, p.eps,
                                                 p.cos ? p.cos + (int64_t)ktok * (D / 2) : nullptr,
                                                 p.sin ? p.sin + (int64_t)ktok * (D / 2) : nullptr);
        if (jb == 0) kmx[0] = fmaxf(kmx[0], nrm2); else kmx[1] = fmaxf(kmx[1], nrm2);
      }
      for (int jb = 0; jb < g.NKB; ++jb) {
        float v = (jb == 0 ? kmx[0] : kmx[1]) * 1.02f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((lt & 31) == 0) kmax2[(t.ring * g.NKB + jb) * kPPKvWarps + (lt >> 5)] = v;
      }
      fence_proxy_async_smem();
      for (int jb = 0; jb < g.NKB; ++jb) mbar_arrive(kv_full(t.ring * g.NKB + jb));
    };
    KTask cur, nxt;
    bool have_cur = next_task(cur);
    if (have_cur) issue(cur, true);
    while (have_cur) {
      const bool have_nxt = next_task(nxt);
      const bool early = have_nxt && issue(nxt, false);   // per lane: a lane that could not start early issues after finish()
      if (early) cp_async_wait<1>(); else cp_async_wait<0>();
      if (lt == 0) PP_TR(2, 31);
      finish(cur);
      if (lt == 0) PP_TR(2, 32);
      if (have_nxt && !early) issue(nxt, true);
      cur = nxt;
      have_cur = have_nxt;
    }
  } else {
    // ============================ tcgen05 issuer ============================
    // The whole warp runs the control flow (barrier waits, descriptor arithmetic stay warp-uniform, so operands live
    // in uniform registers); one elected lane issues the tcgen05 instructions.
    const bool leader = elect_one();
    uint32_t n_q[2] = {0, 0}, n_p[2] = {0, 0}, n_kv[kPPMaxStages] = {0, 0, 0, 0};
    const uint32_t idesc_s = make_idesc_bf16_f32(128, g.BK);
    const uint32_t idesc_om = make_idesc_bf16_f32_bmn(128, NMAIN);
    const uint32_t idesc_ot = make_idesc_bf16_f32_bmn(128, 16);
    const uint32_t kchunk16 = (uint32_t)(g.BK * 128) >> 4;   // K main chunk stride in descriptor (16 B) units
    const int steps = g.BK / 16;
    int64_t set_ord = -1;
    int set_slot[2] = {0, 0};     // ring position of the key set each slot's current job uses
    bool set_first[2] = {false, false}, set_last[2] = {false, false};
    auto issue_s = [&](int s, int jb) {
      const int st = set_slot[s] * g.NKB + jb;
      if (set_first[s]) { mbar_wait(kv_full(st), n_kv[st] & 1); ++n_kv[st]; }
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(s * g.s_cols);
      // descriptors advance by constant address steps: built once per S, bumped per MMA
      const uint64_t qd0 = make_sw128_kmajor_desc(smem_u32(sQ(s)));
      const uint64_t kd0 = make_sw128_kmajor_desc(smem_u32(sK(st)));
      const uint64_t qt0 = make_noswz_kmajor_desc(smem_u32(sQt(s)));
      const uint64_t kt0 = make_noswz_kmajor_desc(smem_u32(sKt(st)));
      if (leader) {
        uint32_t acc = 0;
#pragma unroll
        for (int kc = 0; kc < Cfg::MAIN; ++kc) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_bf16<1>(d, qd0 + (uint64_t)(kc * ((128 * 128) >> 4) + ks * 2), kd0 + (uint64_t)(kc * kchunk16 + ks * 2), idesc_s, acc);
            acc = 1;
          }
        }
        if (Cfg::TAIL) umma_bf16<1>(d, qt0, kt0, idesc_s, acc);
        umma_commit<1>(s_full(s));
        if (jb == g.NKB - 1) umma_commit<1>(q_empty(s));   // this slot's Q tile may be overwritten
      }
      __syncwarp();
      if (leader) PP_TR(3, 44 + s);
    };
    auto start = [&](int s, int i) {   // bind job i to slot s and issue its first S
      set_first[s] = first_of_set(i);
      set_last[s] = last_of_set(i);
      if (set_first[s]) ++set_ord;
      set_slot[s] = (int)(set_ord % g.NSETS);
      mbar_wait(q_full(s), n_q[s] & 1); ++n_q[s];
      if (leader) PP_TR(3, 40 + s);
      issue_s(s, 0);
    };
    auto step = [&](int s, int jb) {   // O (+)= P V of key block jb, then the S of the next block
      const int st = set_slot[s] * g.NKB + jb;
      mbar_wait(p_full(s), n_p[s] & 1); ++n_p[s];
      tc_fence_after();
      if (leader) PP_TR(3, 46 + s);
      const uint32_t d_o = tmem_base + (uint32_t)(g.o_col + s * g.o_cols);
      const uint32_t a_p = tmem_base + (uint32_t)(s * g.s_cols);
      const uint64_t vd0 = make_sw128_mnmajor_desc(smem_u32(sV(st)), (uint32_t)(g.BK * 128));
      const uint64_t vt0 = make_noswz_mnmajor_desc(smem_u32(sVt(st)));
      const uint32_t acc0 = jb > 0 ? 1u : 0u;
      if (leader) {
        // per 16 keys: one MMA over the swizzled main chunk(s) (N = 64 / 128), one over the head-dim tail (N = 16)
        umma_bf16_ts(d_o, a_p, vd0, idesc_om, acc0);
        if (Cfg::TAIL) umma_bf16_ts(d_o + NMAIN, a_p, vt0, idesc_ot, acc0);
#pragma unroll 3
        for (int k = 1; k < steps; ++k) {
          umma_bf16_ts(d_o, a_p + k * 8, vd0 + (uint64_t)(k * (2048 >> 4)), idesc_om, 1u);
          if (Cfg::TAIL) umma_bf16_ts(d_o + NMAIN, a_p + k * 8, vt0 + (uint64_t)(k * (512 >> 4)), idesc_ot, 1u);
        }
        if (set_last[s] && jb == g.NKB - 1)   // every reader of this key set has been issued: free its stages
          for (int b2 = 0; b2 < g.NKB; ++b2) umma_commit<1>(kv_empty(set_slot[s] * g.NKB + b2));
        umma_commit<1>(o_full(s));
      }
      __syncwarp();
      if (leader) PP_TR(3, 48 + s);
      if (jb + 1 < g.NKB) issue_s(s, jb + 1);
    };
    for (int jp = 0; jp < nj; jp += 2) {
      const int ns = (nj - jp) < 2 ? (nj - jp) : 2;
      // With a single ring position, a new key set can only be staged after the previous set's last P V: a pair that
      // straddles the boundary runs its two jobs one after the other (happens once per head in cross attention).
      const bool serial = ns == 2 && g.NSETS == 1 && first_of_set(jp + 1);
      if (serial) {
        start(0, jp);
        for (int jb = 0; jb < g.NKB; ++jb) step(0, jb);
        start(1, jp + 1);
        for (int jb = 0; jb < g.NKB; ++jb) step(1, jb);
      } else {
        for (int s = 0; s < ns; ++s) start(s, jp + s);
        for (int jb = 0; jb < g.NKB; ++jb)
          for (int s = 0; s < ns; ++s) step(s, jb);
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == kPPIssuerWarp) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512u);
  }
}

// eligibility + geometry of the ping-pong kernel; returns OSB_ERR_UNSUPPORTED (without setting an error) when the
// shape belongs to the other kernels
template <int D>
static int attn_pp_launch(AttnParams& p, int H, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  PPGeom g = {};
  auto up1k = [](int x) { return (x + 1023) / 1024 * 1024; };
  g.NKB = p.NK > 160 ? 2 : 1;
  g.BK = (((p.NK + g.NKB - 1) / g.NKB) + 15) / 16 * 16;
  if (p.NK > 320 || g.BK > 160 || p.rope_half) return OSB_ERR_UNSUPPORTED;
  g.qt_off = Cfg::MAIN * 128 * 128;
  g.q_bytes = g.qt_off + up1k(Cfg::TAIL ? 128 * 32 : 0);
  g.k_bytes = up1k(Cfg::MAIN * g.BK * 128);
  g.kt_bytes = up1k(Cfg::TAIL ? g.BK * 32 : 0);
  g.stage_bytes = 2 * (g.k_bytes + g.kt_bytes);
  g.off_kv = 2 * g.q_bytes;
  const int budget = 227 * 1024 - g.off_kv - 256 - 4 * 128 * 4 - kPPMaxStages * kPPKvWarps * 4;
  int nst = budget / g.stage_bytes;
  if (nst > kPPMaxStages) nst = kPPMaxStages;
  g.NSETS = nst / g.NKB;
  if (g.NSETS < 1) return OSB_ERR_UNSUPPORTED;
  g.off_bar = g.off_kv + g.NSETS * g.NKB * g.stage_bytes;
  const int smem = g.off_bar + 256 + 4 * 128 * 4 + kPPMaxStages * kPPKvWarps * 4;
  g.s_cols = (g.BK + 31) / 32 * 32;
  g.o_cols = Cfg::DP;
  g.o_col = 2 * g.s_cols;
  if (g.o_col + 2 * g.o_cols > 512) return OSB_ERR_UNSUPPORTED;
  g.tps = (p.G > 1) ? 1 : p.tiles_per_seq;
  g.units = (p.G > 1) ? (p.num_seqs + p.G - 1) / p.G : p.num_seqs;
  g.num_sets = g.units * H;
  const int64_t work = (g.tps <= 4) ? g.num_sets : g.num_sets * g.tps;
  const int64_t grid = work < sm_count() ? work : sm_count();
  cudaLaunchAttribute attr[2];
  cudaLaunchConfig_t cfg = launch_config(dim3((unsigned)grid), dim3(kPPThreads), smem, stream, attr);
  OSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_pp_kernel<D>, p, g));
  count_launch();
  return OSB_OK;
}

template <int D>
static int attn_launch(AttnParams& p, int64_t units, int H, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  // shared memory carve-up (all tile bases 1024-byte aligned for the 128B swizzle)
  auto up1k = [](int x) { return (x + 1023) / 1024 * 1024; };
  const int nkc = (p.NKP + 63) / 64;
  int off = Cfg::MAIN * 128 * 128;                 // sQ main
  p.off_qt = off; off += up1k(Cfg::TAIL ? 128 * 32 : 0);
  p.off_k = off; off += up1k(Cfg::MAIN * p.NKP * 128);
  p.off_kt = off; off += up1k(Cfg::TAIL ? p.NKP * 32 : 0);
  p.off_vt = off; off += up1k(nkc * Cfg::DP * 128);
  p.off_p = off; off += nkc * 128 * 128;
  p.off_misc = off; off += 2056 + 32;
  const int smem = off;
  const int s_cols = (p.NKP + 31) / 32 * 32;
  p.o_col = s_cols;
  p.tmem_cols = (s_cols + Cfg::DP <= 256) ? 256 : 512;
  if (s_cols + Cfg::DP > 512 || smem > 227 * 1024) {
    set_error("osb_attn_short: %d keys x head_dim %d need %d B smem / %d TMEM columns", p.NKP, D, smem, s_cols + Cfg::DP);
    return OSB_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)units, (unsigned)H);
  cudaLaunchAttribute attr[2];
  cudaLaunchConfig_t cfg = launch_config(grid, dim3(kAttnThreads), smem, stream, attr);
  OSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_short_kernel<D>, p));
  count_launch();
  return OSB_OK;
}

// OSB_ATTN_PP=0/1: let the default dispatch use the ping-pong kernel for two-block resident key sets
bool pp_auto() {
  static const bool on = [] { const char* e = getenv("OSB_ATTN_PP"); return e != nullptr && e[0] == '1'; }();
  return on;
}

int attn_init() {
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_short_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_short_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_short_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_kernel<72, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_kernel<72, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_pp_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_pp_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return OSB_OK;
}

}  // namespace osb

#ifdef OSB_PP_TRACE
extern "C" int osb_debug_pp_trace(unsigned long long* dst, int* counts) {
  int zero[4] = {0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(dst, osb::g_pp_trace, sizeof(unsigned long long) * 4 * 512) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(counts, osb::g_pp_trace_n, sizeof(int) * 4) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(osb::g_pp_trace_n, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif

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
  OSB_REQUIRE((a->q_norm_w2 == nullptr) == (a->k_norm_w2 == nullptr) && (a->q_norm_w2 == nullptr || a->q_norm_w != nullptr),
              "osb_attn_short: the second norm weight pair needs the first");
  OSB_REQUIRE(a->rope_cos == nullptr || ((reinterpret_cast<uintptr_t>(a->rope_cos) | reinterpret_cast<uintptr_t>(a->rope_sin)) & 15) == 0,
              "osb_attn_short: rope tables must be 16-byte aligned");

  AttnParams p;
  p.q = static_cast<const __nv_bfloat16*>(a->q);
  p.k = static_cast<const __nv_bfloat16*>(a->k);
  p.v = static_cast<const __nv_bfloat16*>(a->v);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.q_ld = a->q_ld; p.k_ld = a->k_ld; p.v_ld = a->v_ld; p.out_ld = a->out_ld;
  p.num_seqs = a->num_seqs; p.seqs_per_batch = a->seqs_per_batch;
  p.q_bs = a->q_batch_stride; p.q_ss = a->q_seq_stride; p.q_ts = a->q_tok_stride;
  p.k_bs = a->k_batch_stride; p.k_ss = a->k_seq_stride; p.k_ts = a->k_tok_stride;
  p.Lq = a->Lq; p.Lk = a->Lk; p.kv_lens = a->kv_lens;
  p.qw = static_cast<const __nv_bfloat16*>(a->q_norm_w);
  p.kw = static_cast<const __nv_bfloat16*>(a->k_norm_w);
  p.qw2 = static_cast<const __nv_bfloat16*>(a->q_norm_w2);
  p.kw2 = static_cast<const __nv_bfloat16*>(a->k_norm_w2);
  p.norm_split = a->norm_split;
  p.eps = a->norm_eps;
  p.cos = a->rope_cos; p.sin = a->rope_sin;
  p.rope_half = a->rope_cos != nullptr && a->rope_half ? 1 : 0;
  OSB_REQUIRE(!p.rope_half || (D % 16 == 0), "osb_attn_short: rotate-half RoPE needs head_dim %% 16 == 0");
  p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
  int64_t units;
  if (a->Lq >= 128) {
    p.G = 1;
    p.tiles_per_seq = (a->Lq + 127) / 128;
    // q-tiles per CTA: amortise the K/V staging against whole waves of CTAs (1 CTA per SM at these smem sizes).
    // cost(QT) = waves x (QT + staging), staging ~ 0.9 q-tile-equivalents per 256 keys (measured, profiles/).
    const double staging = 0.9 * (double)a->Lk / 256.0;
    int qt = 1;
    double best = 1e30;
    for (int cand = 1; cand <= p.tiles_per_seq && cand <= 64; ++cand) {
      const int64_t groups = (p.tiles_per_seq + cand - 1) / cand;
      const int64_t ctas = a->num_seqs * groups * a->num_heads;
      const int64_t waves = (ctas + sm_count() - 1) / sm_count();
      const double per_cta = (double)((p.tiles_per_seq + groups - 1) / groups) + staging;
      const double cost = (double)waves * per_cta;
      if (cost < best - 1e-9) { best = cost; qt = cand; }
    }
    p.QT = qt;
    p.groups_per_seq = (p.tiles_per_seq + qt - 1) / qt;
    units = a->num_seqs * p.groups_per_seq;
  } else {
    p.G = 128 / a->Lq;
    // packing more sequences than the key budget allows would overflow the S tile: shrink G
    while (p.G > 1 && (int64_t)p.G * a->Lk > kMaxKeys) --p.G;
    p.tiles_per_seq = 1;
    p.QT = 1;
    p.groups_per_seq = 1;
    units = (a->num_seqs + p.G - 1) / p.G;
  }
  p.NK = p.G * a->Lk;
  p.NKP = (p.NK + 15) / 16 * 16;
  // implementation: reserved = 1 resident-key kernel, 2 flash (P through smem), 3 flash (P in TMEM), 0 = default
  int impl = a->reserved;
  if (impl == 0 && pp_auto() && D <= 72 && !p.rope_half && p.NK > 160 && p.NK <= 320) impl = 4;   // two resident key blocks
  if (impl == 0) {
    // measured on B200 (profiles/r01_attn_v4.log): the flash kernel with P in TMEM wins when two key blocks are
    // resident and two CTAs fit an SM (STDiT3 spatial: 128 vs 153 us); the one-pass resident kernel wins for a
    // single block (temporal) and for 3 resident blocks at one CTA per SM (T5 cross: 117 vs 209 us)
    const int nkb = (p.NK + 127) / 128;
    impl = (p.NK > kMaxKeys) ? 3 : ((nkb == 2 && D <= 72) ? 3 : 1);
  }
  if (impl == 4) {   // ping-pong kernel (resident key sets, head_dim <= 72); other shapes keep the default choice
    int rc = OSB_ERR_UNSUPPORTED;
    if (D == 64) rc = attn_pp_launch<64>(p, a->num_heads, static_cast<cudaStream_t>(stream));
    else if (D == 72) rc = attn_pp_launch<72>(p, a->num_heads, static_cast<cudaStream_t>(stream));
    if (rc != OSB_ERR_UNSUPPORTED) return rc;
    const int nkb = (p.NK + 127) / 128;
    impl = (p.NK > kMaxKeys) ? 3 : ((nkb == 2 && D <= 72) ? 3 : 1);
  }
  if ((p.NK > kMaxKeys || p.rope_half) && impl == 1) impl = 3;  // the resident kernel splits q rows between two threads
  if (impl == 2 || impl == 3) {
    cudaStream_t fs = static_cast<cudaStream_t>(stream);
    const int H = a->num_heads;
    if (impl == 3) {
      if (D == 64) return attn_flash_launch<64, true>(p, H, fs);
      if (D == 72) return attn_flash_launch<72, true>(p, H, fs);
      return attn_flash_launch<128, true>(p, H, fs);
    }
    if (D == 64) return attn_flash_launch<64, false>(p, H, fs);
    if (D == 72) return attn_flash_launch<72, false>(p, H, fs);
    return attn_flash_launch<128, false>(p, H, fs);
  }
  OSB_REQUIRE(p.NK <= kMaxKeys, "osb_attn_short: %d keys per tile exceed the %d-key budget of the resident kernel", p.NK, kMaxKeys);
  OSB_REQUIRE(units <= 0x7fffffff && a->num_heads <= 65535, "osb_attn_short: grid too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (D == 64) return attn_launch<64>(p, units, a->num_heads, s);
  if (D == 72) return attn_launch<72>(p, units, a->num_heads, s);
  return attn_launch<128>(p, units, a->num_heads, s);
}
