// Attention over head tiles (tiles.cuh) on sm_100a: bulk-copy fed, persistent, two query tiles in flight per CTA.
//
// q / k / v arrive as operand-tile images written by the projection GEMM's head-tile epilogue (bias, per-head
// RMSNorm and RoPE already applied), so every load is ONE cp.async.bulk of a whole tile issued by a single
// thread - no register-path loads, no transforms, no loader warps.  One CTA per SM walks a contiguous range of
// "pairs" of query tiles:
//   warps 0-3 / 4-7   softmax warpgroup of slot 0 / 1: one thread per query row; S is read from TMEM ONCE into
//                     registers (TMEM reads are the scarce resource), exact running max with lazy rescale (O is
//                     only touched when a row maximum grows by more than 2^8), P (bf16) written in place over S;
//   warp 8            loader: bulk copies of Q tiles (one buffer per slot) and K/V tiles (ring of stages);
//   warp 9            tcgen05 issuer: S = Q K^T (K-major tiles), O += P V with P as a tensor-memory A operand and V
//                     as an MN-major B operand (the K tile layout, no transposition anywhere).
// The issuer interleaves the two slots (S_a, S_b, PV_a, S_a', PV_b, S_b', ...), so the tensor pipe serves one slot
// while the other is in its exp2 pass.  Both slots work on query tiles of the SAME key set when a sequence has
// several query tiles (STDiT3 spatial S = 256, T5 cross-attention: the key tiles are loaded once per pair and stay
// resident across pairs of one (sequence, head)), or on two different packed tiles (temporal T = 64: two sequences
// per tile, block-diagonal mask).
//
// Replaces: opensora/models/mmdit/math.py:22-36 (attention) for the STDiT3 Attention / MultiHeadCrossAttention
// restated in SURVEY.md App. A; QK-RMSNorm + RoPE (layers.py:102-135, math.py:60-65) moved into the GEMM epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "tiles.cuh"

namespace osb {

constexpr int kTAThreads = 384;   // 3 warpgroups: softmax 0, softmax 1, {loader, issuer, 2 output store warps}
constexpr int kTALoaderWarp = 8;
constexpr int kTAIssuerWarp = 9;
constexpr int kTAMaxStages = 4;
constexpr float kTARescaleThreshold = 8.0f;   // log2 units: P stays below 2^8 without touching O

struct TileAttnParams {
  const uint8_t* q; const uint8_t* k; const uint8_t* v;   // first tile of head 0
  int64_t q_head_stride, kv_head_stride;                  // bytes
  TileMap qmap;                                           // q tiles <-> rows of `out`
  int32_t q_tile_bytes, q_slot_bytes;                     // TRq * ROW_BYTES, rounded up to 1024
  int32_t kv_tile_bytes, kv_slot_bytes;
  int32_t BK, nkb, Lk;                                    // key-tile rows, key tiles per set, keys per sequence
  int64_t num_seqs, num_sets;
  const int32_t* kv_lens;
  int32_t H;
  __nv_bfloat16* out;
  int64_t out_ld;
  float scale_log2;
  int32_t shared_mode;      // 1: both slots take query tiles of one key set; 0: two different sets
  int32_t ppg;              // pairs per group: shared: ceil(tps / 2) per (set, head); split: ceil(num_sets / 2) per head
  int64_t num_pairs;
  int32_t nst;              // K/V ring stages
  int32_t resident;         // shared mode and nkb <= nst: key tiles stay across the pairs of one (set, head)
  int32_t off_kv, off_out, off_bar;  // shared memory carve-up
  RowScatter out_sc;        // sequence parallel: output rows go straight to the consuming rank's buffer
  int32_t exp_flags;        // OSB_TA_EXP (timing experiments only, results become wrong): 1 no PV tail MMA, 2 no PV MMAs,
                            // 4 no S tail MMA, 8 no exp2 (P = 0), 16 no output epilogue, 32 no S MMAs (any value: generic PV issue loop)
};

#ifdef OSB_TA_TRACE
// debug build only (tests/ta_trace.py): CTA 0 records (tag, clock64) per role: 0/1 softmax slots, 2 loader, 3 issuer.
// The event count lives in a register of the tracing lane (a global counter would add an L2 round trip per event).
__device__ unsigned long long g_ta_trace[4][1024];
__device__ int g_ta_trace_n[4];
#define TA_TR_DECL int ta_n_ = 0;
#define TA_TR(role, tag)                                                                                   \
  do {                                                                                                     \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ta_n_ < 512) {                                       \
      g_ta_trace[role][2 * ta_n_] = (tag); g_ta_trace[role][2 * ta_n_ + 1] = clock64(); ++ta_n_;          \
      g_ta_trace_n[role] = ta_n_;                                                                          \
    }                                                                                                      \
  } while (0)
#else
#define TA_TR_DECL
#define TA_TR(role, tag) do { } while (0)
#endif

struct PairJob {
  int head;
  int set[2];
  int qt[2];
  bool act[2];
  int keys[2];   // valid key slots of the set (packed tiles: all G * Lk)
  int nkb[2];    // key tiles that hold valid keys (>= 1)
  bool load, release;
};

template <int D>
__global__ void __launch_bounds__(kTAThreads, 1) attn_tiles_kernel(const TileAttnParams p) {
  using Cfg = HeadTileCfg<D>;
  constexpr int NMAIN = Cfg::MAIN * 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem0 = smem_u32(smem);
  auto sQ = [&](int slot) { return smem0 + (uint32_t)(slot * p.q_slot_bytes); };
  auto sK = [&](int st) { return smem0 + (uint32_t)(p.off_kv + st * 2 * p.kv_slot_bytes); };
  auto sV = [&](int st) { return smem0 + (uint32_t)(p.off_kv + st * 2 * p.kv_slot_bytes + p.kv_slot_bytes); };
  const uint32_t bar0 = smem0 + (uint32_t)p.off_bar;
  auto q_full = [&](int s) { return bar0 + 8u * s; };
  auto q_empty = [&](int s) { return bar0 + 16u + 8u * s; };
  auto s_full = [&](int s) { return bar0 + 32u + 8u * s; };
  auto p_full = [&](int s) { return bar0 + 48u + 8u * s; };
  auto o_full = [&](int s) { return bar0 + 64u + 8u * s; };
  auto kv_full = [&](int st) { return bar0 + 80u + 8u * st; };
  auto kv_empty = [&](int st) { return bar0 + 80u + 8u * kTAMaxStages + 8u * st; };
  auto out_full = [&](int s) { return bar0 + 80u + 16u * kTAMaxStages + 8u * s; };
  auto out_empty = [&](int s) { return bar0 + 96u + 16u * kTAMaxStages + 8u * s; };
  const uint32_t tmem_slot = bar0 + 112u + 16u * kTAMaxStages;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == kTAIssuerWarp) {
    if ((tid & 31) == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); mbar_init(s_full(s), 1);
        mbar_init(p_full(s), 128); mbar_init(o_full(s), 1);
        mbar_init(out_full(s), 128); mbar_init(out_empty(s), 1);
      }
      for (int st = 0; st < kTAMaxStages; ++st) { mbar_init(kv_full(st), 1); mbar_init(kv_empty(st), 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  pdl_wait();

  // this CTA's contiguous pair range (host: num_pairs < 2^31)
  const int pair_lo = (int)((int64_t)blockIdx.x * p.num_pairs / gridDim.x);
  const int pair_hi = (int)((int64_t)(blockIdx.x + 1) * p.num_pairs / gridDim.x);
  const uint32_t ppg = (uint32_t)p.ppg, nH = (uint32_t)p.H;

  auto decode = [&](int pair, PairJob& j) {
    if (p.shared_mode) {
      const uint32_t grp = (uint32_t)pair / ppg;
      const uint32_t jp = (uint32_t)pair - grp * ppg;
      const uint32_t set = grp / nH;
      j.head = (int)(grp - set * nH);
      j.set[0] = j.set[1] = (int)set;
      j.qt[0] = 2 * (int)jp; j.qt[1] = 2 * (int)jp + 1;
      j.act[0] = true; j.act[1] = j.qt[1] < p.qmap.tps;
      int keys = p.Lk;
      if (p.kv_lens) { const int l = __ldg(p.kv_lens + set); keys = l < keys ? (l < 0 ? 0 : l) : keys; }
      j.keys[0] = j.keys[1] = keys;
      const int nk = keys > 0 ? (int)((uint32_t)(keys + p.BK - 1) / (uint32_t)p.BK) : 1;
      j.nkb[0] = j.nkb[1] = nk;
      const bool same_prev = pair > pair_lo && jp > 0;                       // the previous pair belongs to this group
      const bool same_next = pair + 1 < pair_hi && jp + 1 < ppg;
      j.load = !(p.resident && same_prev);
      j.release = !(p.resident && same_next);
    } else {
      j.head = (int)((uint32_t)pair / ppg);
      const int ip = pair - j.head * (int)ppg;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        j.set[s] = 2 * ip + s;
        j.qt[s] = 0;
        j.act[s] = j.set[s] < (int)p.num_sets;
        int keys = p.qmap.G > 1 ? p.qmap.G * p.Lk : p.Lk;
        if (p.qmap.G == 1 && p.kv_lens && j.act[s]) { const int l = __ldg(p.kv_lens + j.set[s]); keys = l < keys ? (l < 0 ? 0 : l) : keys; }
        j.keys[s] = keys;
        j.nkb[s] = keys > 0 ? (int)((uint32_t)(keys + p.BK - 1) / (uint32_t)p.BK) : 1;
      }
      j.load = true;
      j.release = true;
    }
  };
  // valid key columns of key tile kb, rounded up to one MMA K step
  auto ncol_of = [&](int keys, int kb) {
    int n = keys - kb * p.BK;
    n = n < 0 ? 0 : (n > p.BK ? p.BK : n);
    return (n + 15) & ~15;
  };

  // register file: the softmax warpgroups hold a whole 128-column S row per thread; the loader / issuer need few.
  // 216 * 256 + 72 * 128 = 168 * 384: exactly what the launch allocated (asking for more blocks forever)
  if (warp < 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;\n" ::);
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n" ::);

  if (warp < 8) {
    // =========================================== softmax warpgroups ===========================================
    const int slot = warp >> 2;
    const int r = tid & 127;
    const uint32_t t_s = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(slot * 256);
    const uint32_t t_o = t_s + 128;
    const float sc = p.scale_log2;
    uint32_t n_s = 0, n_o = 0, n_out = 0;
    TA_TR_DECL
    for (int pair = pair_lo; pair < pair_hi; ++pair) {
      PairJob j;
      decode(pair, j);
      // this slot's job as scalars (a runtime index into the struct would push it to local memory)
      const bool j_act = slot ? j.act[1] : j.act[0];
      const int j_set = slot ? j.set[1] : j.set[0];
      const int j_qt = slot ? j.qt[1] : j.qt[0];
      const int j_keys = slot ? j.keys[1] : j.keys[0];
      const int nkb = slot ? j.nkb[1] : j.nkb[0];
      const int j_head = j.head;
      if (!j_act) continue;
      // my query row: sequence, position, valid key range [lo, hi) in the set's key slots
      int64_t seq;
      int pos, lo = 0, hi = 0;
      bool valid;
      if (p.qmap.G > 1) {
        const int g = r / p.qmap.L;
        pos = r - g * p.qmap.L;
        seq = (int64_t)j_set * p.qmap.G + g;
        valid = g < p.qmap.G && seq < p.num_seqs;
        if (valid) { lo = g * p.Lk; hi = lo + p.Lk; }
      } else {
        pos = j_qt * p.qmap.TR + r;
        seq = j_set;
        valid = r < p.qmap.TR && pos < p.qmap.L;
        if (valid) hi = j_keys;
      }
      float m = -INFINITY, l = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int ncol = ncol_of(j_keys, kb);
        int blo = lo - kb * p.BK, bhi = hi - kb * p.BK;
        blo = blo < 0 ? 0 : blo;
        bhi = bhi > ncol ? ncol : bhi;
        const bool some = blo < bhi;
        const int wlo = __reduce_min_sync(0xffffffffu, some ? blo : 0x7fffffff);
        const int whi = __reduce_max_sync(0xffffffffu, some ? bhi : 0);
        mbar_wait(s_full(slot), n_s & 1); ++n_s;
        tc_fence_after();
        if ((warp & 3) == 0) TA_TR(slot, 10);
        uint32_t s[4][32];
        bool live[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          live[c] = (c * 32 < whi) && (c * 32 + 32 > wlo);     // warp-uniform
          if (live[c]) tmem_ld_32x32b_x32(t_s + c * 32, s[c]);
        }
        tmem_ld_wait();
        if ((warp & 3) == 0) TA_TR(slot, 11);
        // ---- block maximum over this row's valid keys ----
        float mb = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (!live[c]) continue;
          if (blo <= c * 32 && bhi >= c * 32 + 32) {
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              m0 = fmaxf(m0, __uint_as_float(s[c][e])); m1 = fmaxf(m1, __uint_as_float(s[c][e + 1]));
              m2 = fmaxf(m2, __uint_as_float(s[c][e + 2])); m3 = fmaxf(m3, __uint_as_float(s[c][e + 3]));
            }
            mb = fmaxf(mb, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c * 32 + e >= blo && c * 32 + e < bhi) mb = fmaxf(mb, __uint_as_float(s[c][e]));
          }
        }
        // ---- running maximum: lazy rescale of O (previous PV is complete: s_full was committed after it) ----
        const bool grow = (m != -INFINITY) && (mb * sc > m * sc + kTARescaleThreshold);
        if (kb > 0 && __any_sync(0xffffffffu, grow)) {
          float alpha = 1.f;
          if (grow) { alpha = fast_exp2((m - mb) * sc); m = mb; }
#pragma unroll 1
          for (int c = 0; c < Cfg::DP; c += 8) {
            uint32_t o[8];
            tmem_ld_32x32b_x8(t_o + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st_32x32b_x8(t_o + c, o);
          }
          l *= alpha;
        }
        if (m == -INFINITY) m = mb;
        const float ms = (m == -INFINITY) ? 0.f : m * sc;
        if ((warp & 3) == 0) TA_TR(slot, 12);
        // ---- P = exp2(S * scale - max), row sum, P (bf16) in place over S ----
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 >= ncol) continue;   // warp-uniform: the PV MMA does not read these columns
          uint32_t pk[16];
          if (!live[c] || !some || (p.exp_flags & 8)) {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = 0u;
          } else if (blo <= c * 32 && bhi >= c * 32 + 32) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float p0 = fast_exp2(fmaf(__uint_as_float(s[c][e]), sc, -ms));
              const float p1 = fast_exp2(fmaf(__uint_as_float(s[c][e + 1]), sc, -ms));
              const float p2 = fast_exp2(fmaf(__uint_as_float(s[c][e + 2]), sc, -ms));
              const float p3 = fast_exp2(fmaf(__uint_as_float(s[c][e + 3]), sc, -ms));
              l0 += p0; l1 += p1; l2 += p2; l3 += p3;
              pk[e >> 1] = pack_bf16x2(p0, p1);
              pk[(e >> 1) + 1] = pack_bf16x2(p2, p3);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const bool ok0 = c * 32 + e >= blo && c * 32 + e < bhi;
              const bool ok1 = c * 32 + e + 1 >= blo && c * 32 + e + 1 < bhi;
              const float p0 = ok0 ? fast_exp2(fmaf(__uint_as_float(s[c][e]), sc, -ms)) : 0.f;
              const float p1 = ok1 ? fast_exp2(fmaf(__uint_as_float(s[c][e + 1]), sc, -ms)) : 0.f;
              l0 += p0; l1 += p1;
              pk[e >> 1] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st_32x32b_x16(t_s + c * 16, pk);
        }
        l += (l0 + l1) + (l2 + l3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(slot));
        if ((warp & 3) == 0) TA_TR(slot, 13);
      }
      // ---- epilogue: last PV done -> normalise, round once, store ----
      mbar_wait(o_full(slot), n_o & 1); ++n_o;
      tc_fence_after();
      if ((warp & 3) == 0) TA_TR(slot, 14);
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      if (p.exp_flags & 16) continue;
      uint32_t o[Cfg::U * 8];
#pragma unroll
      for (int c = 0; c + 32 <= D; c += 32) tmem_ld_32x32b_x32(t_o + c, reinterpret_cast<uint32_t(&)[32]>(o[c]));
      if constexpr (D % 32 == 16) tmem_ld_32x32b_x16(t_o + D / 32 * 32, reinterpret_cast<uint32_t(&)[16]>(o[D / 32 * 32]));
      if constexpr (D % 32 == 8) tmem_ld_32x32b_x8(t_o + D / 32 * 32, reinterpret_cast<uint32_t(&)[8]>(o[D / 32 * 32]));
      if constexpr (D % 32 == 24) {
        tmem_ld_32x32b_x16(t_o + D / 32 * 32, reinterpret_cast<uint32_t(&)[16]>(o[D / 32 * 32]));
        tmem_ld_32x32b_x8(t_o + D / 32 * 32 + 16, reinterpret_cast<uint32_t(&)[8]>(o[D / 32 * 32 + 16]));
      }
      tmem_ld_wait();
      tc_fence_before();
      if ((warp & 3) == 0) TA_TR(slot, 15);
      // Output: the thread's row goes through a shared-memory staging tile so that the global stores are issued
      // row-contiguously (U consecutive lanes write one row's D*2 bytes; a thread-per-row store would touch 32 lines per
      // instruction and costs ~70 LSU cycles each).  Row pointers are resolved once per row (sequence-parallel rows may
      // live in a peer's buffer).
      uint8_t* stg = smem + p.off_out + slot * (128 * D * 2 + 128 * 8);
      unsigned long long* rowptr = reinterpret_cast<unsigned long long*>(stg + 128 * D * 2);
      unsigned long long my_ptr = 0ull;
      if (valid) {
        int64_t orow_i = row_of_token(p.qmap, seq, pos);
        __nv_bfloat16* obase = p.out;
        if (p.out_sc.mode != 0) {
          int peer;
          scatter_row(p.out_sc, orow_i, peer, orow_i);
          obase = static_cast<__nv_bfloat16*>(scatter_base(p.out_sc, peer));
        }
        my_ptr = reinterpret_cast<unsigned long long>(obase + orow_i * p.out_ld + (int64_t)j_head * D);
      }
      if (p.off_out < 0) {   // no room for a staging tile (large head_dim): each thread stores its own row
        if (my_ptr != 0ull) {
#pragma unroll
          for (int u = 0; u < Cfg::U; ++u) {
            uint4 w = make_uint4(0, 0, 0, 0);
            if (l > 0.f) {
              w.x = pack_bf16x2(__uint_as_float(o[8 * u]) * inv, __uint_as_float(o[8 * u + 1]) * inv);
              w.y = pack_bf16x2(__uint_as_float(o[8 * u + 2]) * inv, __uint_as_float(o[8 * u + 3]) * inv);
              w.z = pack_bf16x2(__uint_as_float(o[8 * u + 4]) * inv, __uint_as_float(o[8 * u + 5]) * inv);
              w.w = pack_bf16x2(__uint_as_float(o[8 * u + 6]) * inv, __uint_as_float(o[8 * u + 7]) * inv);
            }
            *reinterpret_cast<uint4*>(my_ptr + (unsigned long long)(u * 16)) = w;
          }
        }
        continue;
      }
      mbar_wait(out_empty(slot), (n_out & 1) ^ 1);   // the store warp is done reading the previous job's tile
      ++n_out;
      if ((warp & 3) == 0) TA_TR(slot, 17);
      rowptr[r] = my_ptr;
#pragma unroll
      for (int u = 0; u < Cfg::U; ++u) {
        uint4 w = make_uint4(0, 0, 0, 0);   // no valid key: zeros, never 0 * garbage
        if (l > 0.f) {
          w.x = pack_bf16x2(__uint_as_float(o[8 * u]) * inv, __uint_as_float(o[8 * u + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[8 * u + 2]) * inv, __uint_as_float(o[8 * u + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[8 * u + 4]) * inv, __uint_as_float(o[8 * u + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[8 * u + 6]) * inv, __uint_as_float(o[8 * u + 7]) * inv);
        }
        *reinterpret_cast<uint4*>(stg + r * (D * 2) + u * 16) = w;
      }
      mbar_arrive(out_full(slot));   // release: the staged tile and the row pointers are visible to the store warp
      if ((warp & 3) == 0) TA_TR(slot, 16);
    }
  } else if (warp >= 10) {
    // =========================================== output store warps (one per slot) ===========================================
    // The softmax warpgroup stages a finished [128 x D] bf16 output tile in shared memory and goes on with its next job;
    // this warp writes the tile out row-contiguously (U consecutive lanes cover one row's D*2 bytes) - to this GPU's
    // `out`, or (sequence parallel) straight into the consuming rank's buffer over NVLink.
    const int slot = warp - 10;
    const int lane = tid & 31;
    const uint8_t* stg = smem + p.off_out + slot * (128 * D * 2 + 128 * 8);
    const unsigned long long* rowptr = reinterpret_cast<const unsigned long long*>(stg + 128 * D * 2);
    uint32_t n_out = 0;
    for (int pair = pair_lo; pair < pair_hi; ++pair) {
      PairJob j;
      decode(pair, j);
      if (!(slot ? j.act[1] : j.act[0])) continue;
      if ((p.exp_flags & 16) || p.off_out < 0) continue;
      mbar_wait(out_full(slot), n_out & 1); ++n_out;
#pragma unroll 4
      for (int c = lane; c < 128 * Cfg::U; c += 32) {
        const int row = c / Cfg::U, u = c - row * Cfg::U;
        const unsigned long long dst = rowptr[row];
        if (dst != 0ull) *reinterpret_cast<uint4*>(dst + (unsigned long long)(u * 16)) = *reinterpret_cast<const uint4*>(stg + c * 16);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(out_empty(slot));
    }
  } else if (warp == kTALoaderWarp) {
    // =========================================== loader: bulk copies ===========================================
    const bool leader = elect_one();
    uint32_t n_q[2] = {0, 0};
    uint32_t ring = 0;
    TA_TR_DECL
    for (int pair = pair_lo; pair < pair_hi; ++pair) {
      PairJob j;
      decode(pair, j);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (!j.act[s]) continue;
        mbar_wait(q_empty(s), (n_q[s] & 1) ^ 1);   // the previous job's S MMAs are done with this buffer
        ++n_q[s];
        TA_TR(2, 20 + s);
        if (leader) {
          const int64_t qtile = (int64_t)j.set[s] * p.qmap.tps + j.qt[s];
          mbar_expect_tx(q_full(s), (uint32_t)p.q_tile_bytes);
          bulk_load_1d(sQ(s), p.q + (int64_t)j.head * p.q_head_stride + qtile * p.q_tile_bytes, (uint32_t)p.q_tile_bytes, q_full(s));
        }
        __syncwarp();
      }
      if (!j.load) continue;
      const int nkmax = j.act[1] && !p.shared_mode && j.nkb[1] > j.nkb[0] ? j.nkb[1] : j.nkb[0];
      for (int kb = 0; kb < nkmax; ++kb) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if ((s == 1 && p.shared_mode) || !j.act[s] || kb >= j.nkb[s]) continue;
          const int st = (int)(ring % (uint32_t)p.nst);
          mbar_wait(kv_empty(st), ((ring / (uint32_t)p.nst) & 1) ^ 1);
          ++ring;
          TA_TR(2, 22);
          if (leader) {
            const int64_t ktile = (int64_t)j.set[s] * p.nkb + kb;
            const int64_t off = (int64_t)j.head * p.kv_head_stride + ktile * p.kv_tile_bytes;
            mbar_expect_tx(kv_full(st), 2u * (uint32_t)p.kv_tile_bytes);
            bulk_load_1d(sK(st), p.k + off, (uint32_t)p.kv_tile_bytes, kv_full(st));
            bulk_load_1d(sV(st), p.v + off, (uint32_t)p.kv_tile_bytes, kv_full(st));
          }
          __syncwarp();
        }
      }
    }
    pdl_launch_dependents();
  } else if (warp == kTAIssuerWarp) {
    // =========================================== tcgen05 issuer ===========================================
    // the whole warp runs the control flow (waits / descriptor arithmetic warp-uniform), one elected lane issues
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16_f32(128, (uint32_t)p.BK);
    const uint32_t idesc_om = make_idesc_bf16_f32_bmn(128, NMAIN);
    const uint32_t idesc_ot = make_idesc_bf16_f32_bmn(128, 16);
    // shared-memory descriptors as (low word with the 16-byte-unit address, constant high word): stepping an operand
    // is one 32-bit add, and everything stays in uniform registers
    constexpr uint32_t kHiSw128 = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));   // SBO 1024, version 1, SWIZZLE_128B
    constexpr uint32_t kHiTailK = (uint32_t)((256u >> 4) | (1u << 14));                  // K-major tail: SBO 256
    constexpr uint32_t kHiTailMN = (uint32_t)((128u >> 4) | (1u << 14));                 // MN-major tail: SBO 128
    auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    auto a16 = [](uint32_t addr) { return (addr & 0x3FFFFu) >> 4; };
    const uint32_t q_chunk16 = (uint32_t)(p.qmap.TR * 128) >> 4, k_chunk16 = (uint32_t)(p.BK * 128) >> 4;
    const uint32_t lbo_tail_k = (128u >> 4) << 16, lbo_tail_mn = (256u >> 4) << 16, lbo_v = (k_chunk16 & 0x3FFFu) << 16;
    const uint32_t kv_slot16 = (uint32_t)p.kv_slot_bytes >> 4, stage16 = 2 * kv_slot16;
    const uint32_t q16[2] = {a16(sQ(0)), a16(sQ(1))};
    const uint32_t kv16 = a16(sK(0));
    const uint32_t nst = (uint32_t)p.nst;
    uint32_t n_q[2] = {0, 0}, n_p[2] = {0, 0};
    uint32_t ring = 0, set_ring0 = 0;
    TA_TR_DECL

    for (int pair = pair_lo; pair < pair_hi; ++pair) {
      PairJob j;
      decode(pair, j);
      if (p.shared_mode && j.load) { set_ring0 = ring; ring += (uint32_t)j.nkb[0]; }
      // split mode: slot s, tile kb sits at ring index r0 + (tiles of both slots issued before it), kb-major
      const uint32_t r0 = ring;
      auto ring_of = [&](int s, int kb) -> uint32_t {
        if (p.shared_mode) return set_ring0 + (uint32_t)kb;
        // tiles of slot 0 with index <= kb (s == 1) or < kb (s == 0), tiles of slot 1 with index < kb
        const int n0 = j.nkb[0] < (s == 1 ? kb + 1 : kb) ? j.nkb[0] : (s == 1 ? kb + 1 : kb);
        const int n1 = !j.act[1] ? 0 : (j.nkb[1] < kb ? j.nkb[1] : kb);
        return r0 + (uint32_t)(n0 + n1);
      };
      auto issue_S = [&](int s, int kb) {
        const uint32_t ri = ring_of(s, kb);
        const uint32_t st = ri % nst;
        if (kb == 0) { mbar_wait(q_full(s), n_q[s] & 1); ++n_q[s]; TA_TR(3, 30 + s); }
        if (p.shared_mode ? (j.load && s == 0) : true) { mbar_wait(kv_full((int)st), (ri / nst) & 1); TA_TR(3, 32); }
        tc_fence_after();
        // everything but the tcgen05 instructions themselves runs on the whole warp, so the operands stay in uniform
        // registers and the MMAs issue back to back (a leader-only region makes them thread-private: ~30 extra
        // instructions per MMA)
        const uint32_t tS = tmem_base + (uint32_t)(s * 256);
        const uint32_t qlo = q16[s], klo = kv16 + st * stage16;
#pragma unroll
        for (int kc = 0; kc < Cfg::MAIN; ++kc) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = desc(qlo + kc * q_chunk16 + ks * 2, kHiSw128), db = desc(klo + kc * k_chunk16 + ks * 2, kHiSw128);
            if (leader && !(p.exp_flags & 32)) umma_bf16<1>(tS, da, db, idesc_s, (kc | ks) ? 1u : 0u);
          }
        }
        if (Cfg::TAIL) {
          const uint64_t da = desc((qlo + Cfg::MAIN * q_chunk16) | lbo_tail_k, kHiTailK);
          const uint64_t db = desc((klo + Cfg::MAIN * k_chunk16) | lbo_tail_k, kHiTailK);
          if (leader && !(p.exp_flags & (4 | 32))) umma_bf16<1>(tS, da, db, idesc_s, 1u);
        }
        const bool last_s = kb == j.nkb[s] - 1;
        if (leader) {
          umma_commit<1>(s_full(s));
          if (last_s) umma_commit<1>(q_empty(s));   // the Q buffer may be refilled for the next job
        }
        __syncwarp();
        TA_TR(3, 34 + s);
      };
      auto issue_PV = [&](int s, int kb) {
        const uint32_t ri = ring_of(s, kb);
        const uint32_t st = ri % nst;
        const int steps = ncol_of(j.keys[s], kb) >> 4;
        mbar_wait(p_full(s), n_p[s] & 1); ++n_p[s];
        tc_fence_after();
        TA_TR(3, 36 + s);
        const uint32_t tS = tmem_base + (uint32_t)(s * 256), tO = tS + 128;
        const uint32_t vlo = (kv16 + st * stage16 + kv_slot16) | lbo_v;
        const uint32_t vtl = (kv16 + st * stage16 + kv_slot16 + Cfg::MAIN * k_chunk16) | lbo_tail_mn;
        const uint32_t acc0 = kb > 0 ? 1u : 0u;
        if (steps == 8 && p.exp_flags == 0) {
          // full key tile (the common case): straight-line issue, every operand a compile-time offset from a uniform
          // base - the tensor pipe, not this thread's instruction stream, must pace the PV product
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint64_t dm = desc(vlo + i * (2048 >> 4), kHiSw128), dt = desc(vtl + i * (512 >> 4), kHiTailMN);
            if (leader) {
              umma_bf16_ts(tO, tS + (uint32_t)(i * 8), dm, idesc_om, i == 0 ? acc0 : 1u);
              if (Cfg::TAIL) umma_bf16_ts(tO + NMAIN, tS + (uint32_t)(i * 8), dt, idesc_ot, i == 0 ? acc0 : 1u);
            }
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < steps; ++i) {
            const uint32_t accu = (kb > 0 || i > 0) ? 1u : 0u;
            const uint64_t dm = desc(vlo + i * (2048 >> 4), kHiSw128), dt = desc(vtl + i * (512 >> 4), kHiTailMN);
            if (leader && !(p.exp_flags & 2)) {
              umma_bf16_ts(tO, tS + (uint32_t)(i * 8), dm, idesc_om, accu);
              if (Cfg::TAIL && !(p.exp_flags & 1)) umma_bf16_ts(tO + NMAIN, tS + (uint32_t)(i * 8), dt, idesc_ot, accu);
            }
          }
        }
        // the stage is free once its last reader is done: slot 1 (or slot 0 alone) in shared mode, the slot itself otherwise
        const bool last_reader = p.shared_mode ? (j.release && (s == 1 || !j.act[1])) : true;
        const bool last_pv = kb == j.nkb[s] - 1;
        if (leader) {
          if (last_reader) umma_commit<1>(kv_empty((int)st));
          if (last_pv) umma_commit<1>(o_full(s));
        }
        __syncwarp();
        TA_TR(3, 38 + s);
      };
      const int nkmax = j.act[1] && j.nkb[1] > j.nkb[0] ? j.nkb[1] : j.nkb[0];
      issue_S(0, 0);
      if (j.act[1]) issue_S(1, 0);
      for (int kb = 0; kb < nkmax; ++kb) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!j.act[s] || kb >= j.nkb[s]) continue;
          issue_PV(s, kb);
          if (kb + 1 < j.nkb[s]) issue_S(s, kb + 1);
        }
      }
      if (!p.shared_mode) ring = r0 + (uint32_t)j.nkb[0] + (j.act[1] ? (uint32_t)j.nkb[1] : 0u);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == kTAIssuerWarp) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

int attn_tiles_init() {
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_tiles_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_tiles_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  OSB_CHECK_CUDA(cudaFuncSetAttribute(attn_tiles_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return OSB_OK;
}

template <int D>
static int attn_tiles_launch(TileAttnParams& p, cudaStream_t stream) {
  using Cfg = HeadTileCfg<D>;
  auto up1k = [](int x) { return (x + 1023) / 1024 * 1024; };
  p.q_tile_bytes = p.qmap.TR * Cfg::ROW_BYTES;
  p.q_slot_bytes = up1k(128 * Cfg::ROW_BYTES);      // the S MMA reads 128 rows whatever TR is
  p.kv_tile_bytes = p.BK * Cfg::ROW_BYTES;
  p.kv_slot_bytes = up1k(p.kv_tile_bytes);
  p.off_kv = 2 * p.q_slot_bytes;
  int out_bytes = 2 * (128 * D * 2 + 128 * 8);   // output staging tile + row pointers, per slot
  if ((227 * 1024 - p.off_kv - out_bytes - 1024) / (2 * p.kv_slot_bytes) < 2) out_bytes = 0;   // large head_dim: direct stores
  const int budget = 227 * 1024 - p.off_kv - out_bytes - 1024;
  int nst = budget / (2 * p.kv_slot_bytes);
  if (nst > kTAMaxStages) nst = kTAMaxStages;
  if (nst < 2) { set_error("osb_attn_tiles: key tiles of %d rows do not fit twice in shared memory", p.BK); return OSB_ERR_UNSUPPORTED; }
  p.nst = nst;
  p.off_out = out_bytes ? p.off_kv + nst * 2 * p.kv_slot_bytes : -1;
  p.off_bar = p.off_kv + nst * 2 * p.kv_slot_bytes + out_bytes;
  const int smem = p.off_bar + 256;
  p.shared_mode = (p.qmap.G == 1 && p.qmap.tps >= 2) ? 1 : 0;
  if (p.shared_mode) {
    p.ppg = (p.qmap.tps + 1) / 2;
    p.num_pairs = p.num_sets * p.H * p.ppg;
    p.resident = (p.nkb <= nst && p.ppg > 1) ? 1 : 0;
  } else {
    p.ppg = (int)((p.num_sets + 1) / 2);
    p.num_pairs = (int64_t)p.ppg * p.H;
    p.resident = 0;
  }
  if (p.num_pairs >= (1ll << 31) || p.num_sets >= (1ll << 31)) { set_error("osb_attn_tiles: problem too large (%lld query tile pairs)", (long long)p.num_pairs); return OSB_ERR_UNSUPPORTED; }
  const int64_t grid = p.num_pairs < sm_count() ? p.num_pairs : sm_count();
  cudaLaunchAttribute attr[2];
  cudaLaunchConfig_t cfg = launch_config(dim3((unsigned)grid), dim3(kTAThreads), smem, stream, attr);
  OSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_tiles_kernel<D>, p));
  count_launch();
  return OSB_OK;
}

}  // namespace osb

#ifdef OSB_TA_TRACE
extern "C" int osb_debug_ta_trace(unsigned long long* dst, int* counts) {
  int zero[4] = {0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(dst, osb::g_ta_trace, sizeof(unsigned long long) * 4 * 1024) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(counts, osb::g_ta_trace_n, sizeof(int) * 4) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(osb::g_ta_trace_n, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int64_t osb_head_tiles_per_head(const osb_tile_map* m, int64_t rows) {
  if (m == nullptr || m->L <= 0 || m->tile_rows <= 0) return -1;
  const int64_t seqs = rows / m->L;
  if (m->G > 1) return (seqs + m->G - 1) / m->G;
  return seqs * m->tps;
}

extern "C" int osb_attn_tiles(const osb_attn_tiles_args* a, void* stream) {
  using namespace osb;
  if (!initialised()) { set_error("osb_init() has not been called"); return OSB_ERR_NOT_INIT; }
  OSB_REQUIRE(a != nullptr, "osb_attn_tiles: null args");
  OSB_REQUIRE(a->q_tiles && a->k_tiles && a->v_tiles && (a->out || a->out_scatter), "osb_attn_tiles: null tensor");
  const int D = a->head_dim;
  OSB_REQUIRE(D == 64 || D == 72 || D == 128, "osb_attn_tiles: head_dim %d not built (64, 72, 128)", D);
  const osb_tile_map& m = a->q_map;
  OSB_REQUIRE(m.mode == 0 || m.mode == 1, "osb_attn_tiles: unknown tile map mode %d", m.mode);
  OSB_REQUIRE(m.L > 0 && m.G >= 1 && m.tile_rows > 0 && m.tile_rows <= 128 && m.tile_rows % 8 == 0,
              "osb_attn_tiles: bad q tile map (L %d G %d rows %d)", m.L, m.G, m.tile_rows);
  OSB_REQUIRE(m.G == 1 ? (m.tps == (m.L + m.tile_rows - 1) / m.tile_rows) : (m.G * m.L <= m.tile_rows && m.tps == 1),
              "osb_attn_tiles: q tile map inconsistent (L %d G %d tps %d rows %d)", m.L, m.G, m.tps, m.tile_rows);
  OSB_REQUIRE(m.mode == 0 || (m.S > 0 && m.T == m.L), "osb_attn_tiles: temporal map needs S > 0 and T == L");
  OSB_REQUIRE(a->kv_tile_rows >= 16 && a->kv_tile_rows <= 128 && a->kv_tile_rows % 16 == 0,
              "osb_attn_tiles: key tiles must have 16..128 rows in multiples of 16, got %d", a->kv_tile_rows);
  OSB_REQUIRE(a->Lk > 0 && a->kv_tiles_per_set >= 1 && (int64_t)a->kv_tiles_per_set * a->kv_tile_rows >= (int64_t)(m.G > 1 ? m.G : 1) * a->Lk,
              "osb_attn_tiles: %d key tiles of %d rows cannot hold %d keys", a->kv_tiles_per_set, a->kv_tile_rows, a->Lk);
  OSB_REQUIRE(m.G == 1 || (a->kv_tiles_per_set == 1), "osb_attn_tiles: packed sequences use one key tile per set");
  OSB_REQUIRE(a->num_seqs > 0 && a->num_heads > 0, "osb_attn_tiles: empty problem");
  OSB_REQUIRE(a->out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "osb_attn_tiles: out must be 16-byte aligned");
  OSB_REQUIRE(((reinterpret_cast<uintptr_t>(a->q_tiles) | reinterpret_cast<uintptr_t>(a->k_tiles) | reinterpret_cast<uintptr_t>(a->v_tiles)) & 15) == 0 &&
              a->q_head_stride % 16 == 0 && a->kv_head_stride % 16 == 0, "osb_attn_tiles: tile buffers must be 16-byte aligned");

  TileAttnParams p = {};
  p.q = static_cast<const uint8_t*>(a->q_tiles);
  p.k = static_cast<const uint8_t*>(a->k_tiles);
  p.v = static_cast<const uint8_t*>(a->v_tiles);
  p.q_head_stride = a->q_head_stride; p.kv_head_stride = a->kv_head_stride;
  p.qmap.mode = m.mode; p.qmap.L = m.L; p.qmap.S = m.S; p.qmap.T = m.T; p.qmap.G = m.G; p.qmap.tps = m.tps; p.qmap.TR = m.tile_rows;
  p.BK = a->kv_tile_rows; p.nkb = a->kv_tiles_per_set; p.Lk = a->Lk;
  p.num_seqs = a->num_seqs;
  p.num_sets = m.G > 1 ? (a->num_seqs + m.G - 1) / m.G : a->num_seqs;
  p.kv_lens = a->kv_lens;
  p.H = a->num_heads;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.out_ld = a->out_ld;
  p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
  { static const int exp_flags = [] { const char* e = getenv("OSB_TA_EXP"); return e ? atoi(e) : 0; }(); p.exp_flags = exp_flags; }
  {
    const int64_t out_rows = a->num_seqs * m.L;
    const int rc = make_row_scatter(&p.out_sc, a->out_scatter, out_rows, "osb_attn_tiles");
    if (rc) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (D == 64) return attn_tiles_launch<64>(p, s);
  if (D == 72) return attn_tiles_launch<72>(p, s);
  return attn_tiles_launch<128>(p, s);
}
