// bf16 GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] * W[N,K]^T + bias)
//
// Persistent, warp-specialised kernel: one TMA producer warp, one tcgen05.mma issuer warp, four
// epilogue warps.  Operands are staged by TMA into 128-byte-swizzled shared memory tiles
// (BLOCK_K = 64 bf16 = one swizzle row), the fp32 accumulator lives in TMEM and is double
// buffered so the epilogue of tile i overlaps the main loop of tile i+1.  kCta == 2 pairs two
// SMs on one 256 x BLOCK_N tile (tcgen05.mma.cta_group::2): each CTA stages its own 128 rows of A
// and half of the W tile, the leader CTA issues the MMAs for both.
//
// Replaces every nn.Linear on the denoiser block path of the reference
// (opensora/models/mmdit/layers.py:209-214,247-252,277-281,314-334,401) and the fused epilogues
// replace the separate bias / GELU(tanh) / gate*x+residual elementwise kernels.
#include <stdlib.h>

#include "common.cuh"
#include "tiles.cuh"

namespace osb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kEpiWarps = 8;                          // two warps per TMEM lane quarter, 32 columns each
constexpr int kFirstEpiWarp = 3;                      // warp 0 TMA(A,W), warp 1 MMA, warp 2 TMA(residual)
constexpr int kNumThreads = (kFirstEpiWarp + kEpiWarps) * 32;
constexpr int kStagePitch = 36;                       // floats per staged row: 32 columns + 4 pad (bank spread)
constexpr int kEpiStageBytes = kEpiWarps * 32 * kStagePitch * 4;  // one [32 x 32] fp32 tile per epilogue warp
constexpr int kResChunkBytes = kBlockM * 64 * 2;      // residual ring slot: [128 rows x 64 cols] bf16, SW128
constexpr int kBarBytes = 256;
constexpr int kRopeSmemBytes = 18 * 1024;             // head-tile epilogue: RoPE table of up to 64 positions x 72 dims in smem
constexpr int kSmemTotal = 227 * 1024;

struct GemmEpilogueParams {
  const __nv_bfloat16* bias;
  __nv_bfloat16* D;
  const float* gate;
  const int32_t* mod_index;
  int64_t M, N, K;
  int64_t ldd;
  int64_t group_rows;
  int64_t gate_stride;
  int32_t epilogue;
};

// Head-tile epilogue (kHT): every output row is split into heads of D = BLOCK_N / 2 columns; per head: bias, optional
// RMSNorm (fp32 statistics over the fp32 accumulator), optional interleaved-pair RoPE by token position, one rounding
// to bf16, stored at the row's place inside the head's operand tile (tiles.cuh) - the projection output never exists
// in token layout.  Column group kidx = col / C (C = heads * D) selects the kind (q / k / v) = kidx % nkinds.
struct HeadTileParams {
  uint8_t* base;
  int64_t kind_stride, head_stride;
  TileMap map;
  int32_t tile_bytes;
  int32_t heads, nkinds;
  uint32_t norm_mask, rope_mask;
  const __nv_bfloat16* norm_w[4];
  float eps;
  const float* cos;
  const float* sin;
  // aligned fast path: every CTA's 128 accumulator rows ARE one head tile (row r of the CTA = row r of the tile), so the
  // epilogue builds the tile image in shared memory and writes it with one bulk store per head.
  //   1: contiguous sequences with tile_rows == 128 and L % 128 == 0, or G = 128 / L short sequences packed per tile
  //      (tile = CTA's row block; the temporal case when the producer wrote the rows transposed to [B, S, T]);
  //   2: temporal view: the A operand is loaded through a strided TMA view [k][t][s][b] whose box is (64, T, G):
  //      the M-tile of (batch b, sequence group sg) holds rows g*T + t - exactly the packed temporal attention tile.
  int32_t fast;
  int32_t a_rows;         // accumulator rows that carry data (128, or G*T in mode 2)
  int32_t tiles_total;    // head tiles per head
  int32_t groups_per_batch;   // mode 2: S / G
};

// Implicit-GEMM view of a causal 3D convolution over a (replicate-)padded NDHWC activation tensor: one
// CTA owns a Tt x Ht x Wt box of output positions (128 rows); CTA pairs stack two boxes along H.  The K loop
// walks (kt, kh, kw) taps x 64-channel chunks; every k-block is ONE 5-D TMA box load at the tap's offset.
struct ConvGeom {
  int32_t wt_log2, ht_log2;              // box: Wt = 1 << wt_log2, Ht = 1 << ht_log2, Tt = 128 / (Wt * Ht)
  int32_t tiles_w, tiles_h, tiles_t, nb; // (pair-)tiles per dimension, batch
  int32_t w_out, h_out, t_out;
  int32_t sw, sh, st;                    // convolution strides
  int32_t kw_n, kh_n;                    // taps along w, h (taps along t = k-blocks / (kw_n * kh_n * cin_chunks))
  int32_t cin_chunks;                    // 64-wide channel chunks per tap
};

template <int BLOCK_N, int kCta, bool kRes, bool kHT = false>
struct GemmCfg {
  static constexpr int LOAD_N = BLOCK_N / kCta;
  static constexpr int A_BYTES = kBlockM * kBlockK * 2;
  static constexpr int B_BYTES = LOAD_N * kBlockK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RES_STAGES = kRes ? (BLOCK_N <= 128 ? 4 : 3) : 0;
  static constexpr int RES_BYTES = RES_STAGES * kResChunkBytes;
  // epilogue staging: per-warp fp32 tiles, or (head tiles) the bf16 images of the two head tiles a CTA produces
  static constexpr int ROPE_BYTES = kHT ? kRopeSmemBytes : 0;   // cos / sin rows of a short sequence, staged once per CTA
  static constexpr int EPI_BYTES = kHT ? 2 * kBlockM * (((BLOCK_N / 2) + 15) / 16 * 16) * 2 + ROPE_BYTES : kEpiStageBytes;
  static constexpr int FIXED_BYTES = RES_BYTES + kBarBytes + EPI_BYTES + 1024;  // +1024 alignment slack
  static constexpr int STAGES_RAW = (kSmemTotal - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
  static constexpr uint32_t TMEM_COLS = 512;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(2 * BLOCK_N <= 512, "two accumulator stages must fit TMEM");
  static_assert(B_BYTES % 1024 == 0, "W tile must keep 1024-byte swizzle-atom alignment");
  static_assert(kHT || BLOCK_N % 64 == 0, "epilogue works in 64-column chunks");
  static_assert(BLOCK_N % 16 == 0 && LOAD_N % 8 == 0, "UMMA N / swizzle-atom granularity");
  static_assert(2 * STAGES + 4 + 2 * 4 + 1 <= kBarBytes / 8, "barrier area");
};

template <int BLOCK_N, int kCta, bool kRes, bool kConv, bool kHT = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_r, const GemmEpilogueParams p, const ConvGeom cg,
                 const HeadTileParams ht) {
  using Cfg = GemmCfg<BLOCK_N, kCta, kRes, kHT>;
  constexpr int kStages = Cfg::STAGES;
  constexpr int kResStages = Cfg::RES_STAGES;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles (the offset is identical in both CTAs of a pair)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t res_base = smem_base + kStages * Cfg::STAGE_BYTES;            // residual ring (1024-aligned)
  const uint32_t bar_base = res_base + Cfg::RES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  auto res_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 4 + s); };
  auto res_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 8 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 12);
  const uint32_t stage_base = bar_base + kBarBytes;  // epilogue staging tiles (16-byte aligned)
  auto smem_a = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCta == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;

  const int64_t tile_m = (int64_t)kBlockM * kCta;
  const int64_t num_m_blocks = kConv ? (int64_t)cg.nb * cg.tiles_t * cg.tiles_h * cg.tiles_w
                               : ((kHT && ht.fast == 2) ? ((int64_t)ht.tiles_total + kCta - 1) / kCta : (p.M + tile_m - 1) / tile_m);
  // conv: m_blk -> (batch, t-tile, h-tile, w-tile); this CTA's box origin in OUTPUT coordinates
  auto conv_origin = [&](int64_t m_blk, int& n_i, int& t0, int& h0, int& w0) {
    const int tw = (int)(m_blk % cg.tiles_w);
    const int th = (int)((m_blk / cg.tiles_w) % cg.tiles_h);
    const int tt = (int)((m_blk / ((int64_t)cg.tiles_w * cg.tiles_h)) % cg.tiles_t);
    n_i = (int)(m_blk / ((int64_t)cg.tiles_w * cg.tiles_h * cg.tiles_t));
    w0 = tw << cg.wt_log2;
    h0 = (th * kCta + (int)cta_rank) << cg.ht_log2;
    t0 = tt * (128 >> (cg.wt_log2 + cg.ht_log2));
  };
  const int64_t num_n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int64_t num_tiles = num_m_blocks * num_n_blocks;
  const int64_t num_k_blocks = (p.K + kBlockK - 1) / kBlockK;
  const int64_t cluster_id = blockIdx.x / kCta;
  const int64_t num_clusters = gridDim.x / kCta;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    if constexpr (kRes) tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(tmem_full_bar(s), 1);
        mbar_init(tmem_empty_bar(s), kEpiWarps * kCta);  // one lane per epilogue warp, from every CTA of the pair
      }
      for (int s = 0; s < kResStages; ++s) {
        mbar_init(res_full_bar(s), 1);
        mbar_init(res_empty_bar(s), kEpiWarps);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<kCta>(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    // ===================== TMA producer (A and W tiles) =====================
    if (lane == 0) {
      uint32_t leader_full[kStages];
      if constexpr (kCta == 2) {
        for (int s = 0; s < kStages; ++s)
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_full[s]) : "r"(full_bar(s)), "r"(0));
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int64_t m_blk = tile / num_n_blocks, n_blk = tile % num_n_blocks;
        const int32_t a_row = (int32_t)(m_blk * tile_m + cta_rank * kBlockM);
        const int32_t w_row = (int32_t)(n_blk * BLOCK_N + cta_rank * Cfg::LOAD_N);
        int n_i = 0, t0 = 0, h0 = 0, w0 = 0;
        if constexpr (kConv) conv_origin(m_blk, n_i, t0, h0, w0);
        for (int64_t kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const int32_t k0 = (int32_t)(kb * kBlockK);
          int32_t ac = 0, aw = 0, ah = 0, at = 0;
          if constexpr (kConv) {  // tap offsets in the padded input, output origin scaled by the stride
            const int tap = (int)(kb / cg.cin_chunks);
            ac = (int32_t)(kb - (int64_t)tap * cg.cin_chunks) * kBlockK;
            aw = w0 * cg.sw + tap % cg.kw_n;
            ah = h0 * cg.sh + (tap / cg.kw_n) % cg.kh_n;
            at = t0 * cg.st + tap / (cg.kw_n * cg.kh_n);
          }
          if constexpr (kHT) {
            if (ht.fast == 2) {   // temporal view: box (64 k, T frames, G sequences) of batch b, sequence group sg
              int64_t ti = m_blk * kCta + cta_rank;
              if (ti >= ht.tiles_total) ti = ht.tiles_total - 1;   // odd tile count: the pair's second CTA repeats the last tile
              at = (int32_t)(ti / ht.groups_per_batch);                          // batch
              ah = (int32_t)(ti - (int64_t)at * ht.groups_per_batch) * ht.map.G;  // first sequence (s) of the group
            }
          }
          if constexpr (kCta == 1) {
            mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            if constexpr (kConv) tma_load_5d(&tmap_a, full_bar(stage), smem_a(stage), ac, aw, ah, at, n_i);
            else tma_load_2d(&tmap_a, full_bar(stage), smem_a(stage), k0, a_row);
            tma_load_2d(&tmap_w, full_bar(stage), smem_b(stage), k0, w_row);
          } else {
            uint32_t tx = Cfg::STAGE_BYTES * 2;
            if constexpr (kHT) { if (ht.fast == 2) tx = 2u * (uint32_t)(ht.a_rows * kBlockK * 2 + Cfg::B_BYTES); }
            if (is_leader) mbar_expect_tx(full_bar(stage), tx);
            if constexpr (kConv) tma_load_5d_cg2(&tmap_a, leader_full[stage], smem_a(stage), ac, aw, ah, at, n_i);
            else if (kHT && ht.fast == 2) tma_load_5d_cg2(&tmap_a, leader_full[stage], smem_a(stage), k0, 0, ah, at, 0);
            else tma_load_2d_cg2(&tmap_a, leader_full[stage], smem_a(stage), k0, a_row);
            tma_load_2d_cg2(&tmap_w, leader_full[stage], smem_b(stage), k0, w_row);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      pdl_launch_dependents();  // this CTA has issued all its loads: dependents may start filling vacated SMs
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    // The whole warp runs the loop (waits and descriptor arithmetic stay warp-uniform, so the operands sit in uniform
    // registers and consecutive tcgen05.mma issue back to back); one elected lane issues.  With a single-lane region
    // the compiler wraps every MMA in a ~20-instruction uniformisation loop, which is longer than a 128 x 192 MMA runs.
    if (is_leader) {
      const bool issuer = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16_f32(kBlockM * kCta, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        if constexpr (kCta == 2) mbar_wait_cluster(tmem_empty_bar(as), aphase ^ 1);
        else mbar_wait(tmem_empty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int64_t kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t da0 = make_sw128_kmajor_desc(smem_a(stage));
          const uint64_t db0 = make_sw128_kmajor_desc(smem_b(stage));
          const uint32_t acc0 = kb > 0 ? 1u : 0u;
          if (issuer) {
            umma_bf16<kCta>(d_tmem, da0, db0, idesc, acc0);
#pragma unroll
            for (int k = 1; k < kBlockK / 16; ++k)   // +32 bytes along K per step = +2 in the descriptor's 16-byte units
              umma_bf16<kCta>(d_tmem, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, 1u);
            umma_commit<kCta>(empty_bar(stage));  // smem slot reusable once these MMAs retire
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (issuer) umma_commit<kCta>(tmem_full_bar(as));   // accumulator complete -> epilogue
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================== TMA producer (residual chunks) =====================
    // Streams the residual operand R through a ring of [128 x 64] chunks in exactly the order the
    // epilogue consumes them, a few chunks ahead, so the epilogue never waits on an HBM round trip.
    if constexpr (kRes) {
      if (lane == 0) {
        int rs = 0;
        uint32_t rphase = 0;
        for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
          const int64_t m_blk = tile / num_n_blocks, n_blk = tile % num_n_blocks;
          const int32_t r_row = (int32_t)(m_blk * tile_m + cta_rank * kBlockM);
          int n_i = 0, t0 = 0, h0 = 0, w0 = 0;
          if constexpr (kConv) conv_origin(m_blk, n_i, t0, h0, w0);
          for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {
            const int64_t n0 = n_blk * BLOCK_N + c0;
            if (n0 >= p.N) break;
            mbar_wait(res_empty_bar(rs), rphase ^ 1);
            mbar_expect_tx(res_full_bar(rs), kResChunkBytes);
            if constexpr (kConv) tma_load_5d(&tmap_r, res_full_bar(rs), res_base + rs * kResChunkBytes, (int32_t)n0, w0, h0, t0, n_i);
            else tma_load_2d(&tmap_r, res_full_bar(rs), res_base + rs * kResChunkBytes, (int32_t)n0, r_row);
            if (++rs == kResStages) { rs = 0; rphase ^= 1; }
          }
        }
      }
    }
  } else if constexpr (kHT) {
    // ===================== epilogue warps: head tiles =====================
    // thread = accumulator row; warps 0-3 / 4-7 of the epilogue take head 0 / 1 of the tile's two heads.  The whole
    // head row (D fp32 values) sits in registers, so RMSNorm and RoPE need no exchange between threads.
    constexpr int D = BLOCK_N / 2;
    using HT = HeadTileCfg<D>;
    const int e = warp - kFirstEpiWarp;
    const int q = warp & 3;
    const int hh = e >> 2;
    const int C = ht.heads * D;
    const int r_loc = q * 32 + lane;                       // row inside the CTA's 128-row block
    uint8_t* stage_h = smem_raw + (stage_base - smem_u32(smem_raw)) + hh * (kBlockM * HT::ROW_BYTES);
    const bool store_leader = (q == 0 && lane == 0);       // issues this head's bulk stores
    const int chunk_bytes = ht.map.TR * 128;
    // RoPE tables: in the aligned temporal path a warp's 32 rows are 32 DIFFERENT positions, so table reads from global
    // memory would touch 32 lines per instruction (18 instructions per row): short tables are staged in shared memory once
    // (row pitch D*2 bytes: conflict-free 16-byte reads)
    float* rope_s = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw)) + 2 * (kBlockM * HT::ROW_BYTES));
    const bool rope_smem = ht.rope_mask != 0 && ht.map.L * D * 4 <= kRopeSmemBytes;
    if (rope_smem) {
      const int n = ht.map.L * (D / 2);
      for (int i = (e * 32 + lane); i < n; i += kEpiWarps * 32) { rope_s[i] = __ldg(ht.cos + i); rope_s[n + i] = __ldg(ht.sin + i); }
      asm volatile("bar.sync 3, %0;" ::"r"(kEpiWarps * 32) : "memory");
    }
    int as = 0;
    uint32_t aphase = 0;
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int64_t m_blk = tile / num_n_blocks, n_blk = tile % num_n_blocks;
      const int64_t row = m_blk * tile_m + cta_rank * kBlockM + r_loc;
      mbar_wait(tmem_full_bar(as), aphase);
      tc_fence_after();
      uint32_t raw[D];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + hh * D);
#pragma unroll
      for (int c = 0; c + 32 <= D; c += 32) tmem_ld_32x32b_x32(taddr + c, reinterpret_cast<uint32_t(&)[32]>(raw[c]));
      if constexpr (D % 32 == 8) tmem_ld_32x32b_x8(taddr + D / 32 * 32, reinterpret_cast<uint32_t(&)[8]>(raw[D / 32 * 32]));
      if constexpr (D % 32 == 16) tmem_ld_32x32b_x16(taddr + D / 32 * 32, reinterpret_cast<uint32_t(&)[16]>(raw[D / 32 * 32]));
      tmem_ld_wait();
      // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // relaxed: the payload is TMEM (ordered by the fences above); a release here would wait for the previous
        // tile's global stores to drain while the MMA warp is waiting for this accumulator stage
        if constexpr (kCta == 2) mbar_arrive_cluster_relaxed(tmem_empty_bar(as), 0);
        else mbar_arrive(tmem_empty_bar(as));
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
      const int64_t col0 = n_blk * BLOCK_N + hh * D;
      if (col0 >= p.N) continue;                            // uniform over the four warps of this head
      const int kidx = (int)((uint32_t)col0 / (uint32_t)C);
      const int head = (int)((uint32_t)col0 - (uint32_t)kidx * (uint32_t)C) / D;
      const int kind = kidx % ht.nkinds;
      // token row -> (position in its sequence, tile, row in tile)
      uint32_t pos, tile_i;
      int r;
      bool row_ok;
      if (ht.fast != 0) {
        tile_i = (uint32_t)(m_blk * kCta + cta_rank);
        r = r_loc;
        row_ok = r_loc < ht.a_rows && tile_i < (uint32_t)ht.tiles_total;
        pos = ht.map.G > 1 ? (uint32_t)r_loc % (uint32_t)ht.map.L
                           : (tile_i % (uint32_t)ht.map.tps) * (uint32_t)ht.map.TR + (uint32_t)r_loc;
      } else {   // (32-bit arithmetic: the host checks M < 2^31)
        uint32_t seq;
        const uint32_t row32 = (uint32_t)row;
        row_ok = row < p.M;
        if (ht.map.mode == 0) {
          seq = row32 / (uint32_t)ht.map.L;
          pos = row32 - seq * (uint32_t)ht.map.L;
        } else {
          const uint32_t ts = (uint32_t)ht.map.T * (uint32_t)ht.map.S;
          const uint32_t b = row32 / ts;
          const uint32_t rem = row32 - b * ts;
          pos = rem / (uint32_t)ht.map.S;
          seq = b * (uint32_t)ht.map.S + (rem - pos * (uint32_t)ht.map.S);
        }
        if (ht.map.G > 1) {
          tile_i = seq / (uint32_t)ht.map.G;
          r = (int)((seq - tile_i * (uint32_t)ht.map.G) * (uint32_t)ht.map.L + pos);
        } else {
          const uint32_t jt = pos / (uint32_t)ht.map.TR;
          tile_i = seq * (uint32_t)ht.map.tps + jt;
          r = (int)(pos - jt * (uint32_t)ht.map.TR);
        }
      }
      float x[D];
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = __uint_as_float(raw[i]);
      if (row_ok) {
        if (p.bias) {
#pragma unroll
          for (int u = 0; u < HT::U; ++u) {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col0) + u);
            const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = unpack_bf16x2(bw[k]);
              x[8 * u + 2 * k] += f.x;
              x[8 * u + 2 * k + 1] += f.y;
            }
          }
        }
        if ((ht.norm_mask >> kind) & 1u) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int i = 0; i < D; i += 4) { s0 += x[i] * x[i]; s1 += x[i + 1] * x[i + 1]; s2 += x[i + 2] * x[i + 2]; s3 += x[i + 3] * x[i + 3]; }
          const float rs = rsqrtf(((s0 + s1) + (s2 + s3)) * (1.0f / D) + ht.eps);
          // (a runtime index into the parameter struct would move the whole struct to local memory)
          const __nv_bfloat16* w = kind == 0 ? ht.norm_w[0] : (kind == 1 ? ht.norm_w[1] : (kind == 2 ? ht.norm_w[2] : ht.norm_w[3]));
#pragma unroll
          for (int u = 0; u < HT::U; ++u) {
            const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w) + u);
            const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = unpack_bf16x2(ww[k]);
              x[8 * u + 2 * k] *= rs * f.x;
              x[8 * u + 2 * k + 1] *= rs * f.y;
            }
          }
        }
        if ((ht.rope_mask >> kind) & 1u) {
          const float4* cr = rope_smem ? reinterpret_cast<const float4*>(rope_s + pos * (D / 2))
                                       : reinterpret_cast<const float4*>(ht.cos + (int64_t)pos * (D / 2));
          const float4* sr = rope_smem ? reinterpret_cast<const float4*>(rope_s + ht.map.L * (D / 2) + pos * (D / 2))
                                       : reinterpret_cast<const float4*>(ht.sin + (int64_t)pos * (D / 2));
#pragma unroll
          for (int u = 0; u < HT::U; ++u) {
            const float4 c4 = cr[u], s4 = sr[u];
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a = x[8 * u + 2 * i], b = x[8 * u + 2 * i + 1];
              x[8 * u + 2 * i] = a * cc[i] - b * sn[i];
              x[8 * u + 2 * i + 1] = b * cc[i] + a * sn[i];
            }
          }
        }
      }
      uint8_t* gdst = ht.base + (int64_t)kidx * ht.kind_stride + (int64_t)head * ht.head_stride + (int64_t)tile_i * ht.tile_bytes;
      if (ht.fast != 0) {
        // the previous bulk store of this head must have finished READING the staging image before it is overwritten
        if (store_leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + hh) : "memory");
      }
      uint8_t* dst = ht.fast != 0 ? stage_h : gdst;
      if (ht.fast != 0 || row_ok) {
#pragma unroll
        for (int u = 0; u < HT::UP; ++u) {
          uint4 o = make_uint4(0, 0, 0, 0);
          if (u < HT::U && row_ok) {
            o.x = pack_bf16x2(x[8 * u], x[8 * u + 1]);
            o.y = pack_bf16x2(x[8 * u + 2], x[8 * u + 3]);
            o.z = pack_bf16x2(x[8 * u + 4], x[8 * u + 5]);
            o.w = pack_bf16x2(x[8 * u + 6], x[8 * u + 7]);
          }
          if (ht.fast != 0 && r >= ht.map.TR) continue;   // (tile_rows < 128: rows past the tile are not part of its image)
          if (u < HT::MAIN * 8) *reinterpret_cast<uint4*>(dst + (u >> 3) * chunk_bytes + sw128_off(r, u & 7)) = o;
          else *reinterpret_cast<uint4*>(dst + HT::MAIN * chunk_bytes + tail_off(r, u - HT::MAIN * 8)) = o;
        }
      }
      if (ht.fast != 0) {
        fence_proxy_async_smem();                          // generic-proxy writes -> visible to the bulk-copy engine
        asm volatile("bar.sync %0, 128;" ::"r"(1 + hh) : "memory");
        if (store_leader && tile_i < (uint32_t)ht.tiles_total) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
                       "r"(smem_u32(stage_h)), "r"((uint32_t)ht.tile_bytes)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (ht.fast != 0 && store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // smem must outlive the stores
  } else {
    // ===================== epilogue warps =====================
    // TMEM -> registers (thread = accumulator row) -> per-warp fp32 staging tile in smem -> re-read with
    // a row-contiguous mapping so output stores are coalesced (8 rows x 64 B per warp instruction);
    // the residual arrives through the TMA ring above (swizzled smem, conflict-free reads).
    const int e = warp - kFirstEpiWarp;
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int half = e >> 2;       // which 32-column half of each 64-column chunk
    float* stage_w = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw))) + e * (32 * kStagePitch);
    const uint8_t* res_ring = smem_raw + (res_base - smem_u32(smem_raw));
    int as = 0;
    uint32_t aphase = 0;
    int rs = 0;
    uint32_t rphase = 0;
    const uint32_t group_rows32 = (uint32_t)(p.group_rows > 0xffffffffll ? 0xffffffffu : p.group_rows);
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int64_t m_blk = tile / num_n_blocks, n_blk = tile % num_n_blocks;
      const int64_t row0 = m_blk * tile_m + cta_rank * kBlockM + q * 32;  // first row of this warp
      int n_i = 0, t0 = 0, h0 = 0, w0 = 0;
      if constexpr (kConv) conv_origin(m_blk, n_i, t0, h0, w0);
      mbar_wait(tmem_full_bar(as), aphase);
      tc_fence_after();

#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {
        // bias / gate of this lane's 8 columns: the same for the 4 row groups below, fetched once per chunk and before the
        // TMEM load so their L1 latency is off the per-row path (the gate row only changes at a modulation-group boundary)
        const int64_t nl = n_blk * BLOCK_N + c0 + half * 32 + (lane & 3) * 8;
        const bool nl_ok = nl < p.N;
        uint4 bias_u = make_uint4(0, 0, 0, 0);
        if (p.bias && nl_ok) bias_u = __ldg(reinterpret_cast<const uint4*>(p.bias + nl));
        float4 gate0 = make_float4(1.f, 1.f, 1.f, 1.f), gate1 = gate0;
        bool gate_uniform = false;
        if (!kConv && p.epilogue == OSB_EPI_BIAS_GATE_RES && p.gate != nullptr && nl_ok && row0 < p.M) {
          // rows past M (ragged last tile) carry no data: clamp so the prefetch never indexes past the last group
          const int64_t row_last = row0 + 31 < p.M ? row0 + 31 : p.M - 1;
          const uint32_t g_first = (uint32_t)row0 / group_rows32, g_last = (uint32_t)row_last / group_rows32;
          if (g_first == g_last) {
            gate_uniform = true;
            int64_t gi = g_first;
            if (p.mod_index) gi = p.mod_index[gi];
            const float* gate_row = p.gate + gi * p.gate_stride + nl;
            gate0 = __ldg(reinterpret_cast<const float4*>(gate_row));
            gate1 = __ldg(reinterpret_cast<const float4*>(gate_row + 4));
          }
        }
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + c0 + half * 32);
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        if (c0 + 64 >= BLOCK_N) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCta == 2) mbar_arrive_cluster(tmem_empty_bar(as), 0);
            else mbar_arrive(tmem_empty_bar(as));
          }
        }
        const int64_t n0 = n_blk * BLOCK_N + c0;
        if (n0 >= p.N) continue;  // CTA-uniform: such chunks are not in the residual ring either
        float4* my = reinterpret_cast<float4*>(stage_w + lane * kStagePitch);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          my[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        const uint8_t* res_slot = nullptr;
        if constexpr (kRes) {
          mbar_wait(res_full_bar(rs), rphase);
          res_slot = res_ring + rs * kResChunkBytes;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int i = it * 32 + lane;
          const int rr = i >> 2, g4 = i & 3;
          int64_t row = row0 + rr;
          bool row_ok = row < p.M;
          if constexpr (kConv) {  // box-local row -> (t, h, w) output position -> flattened NDHWC row
            const int li = q * 32 + rr;
            const int w = w0 + (li & ((1 << cg.wt_log2) - 1));
            const int h = h0 + ((li >> cg.wt_log2) & ((1 << cg.ht_log2) - 1));
            const int t = t0 + (li >> (cg.wt_log2 + cg.ht_log2));
            row_ok = w < cg.w_out && h < cg.h_out && t < cg.t_out;
            row = (((int64_t)n_i * cg.t_out + t) * cg.h_out + h) * cg.w_out + w;
          }
          const int64_t n = n0 + half * 32 + g4 * 8;
          if (row_ok && n < p.N) {
            const float4 a0 = *reinterpret_cast<const float4*>(stage_w + rr * kStagePitch + g4 * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(stage_w + rr * kStagePitch + g4 * 8 + 4);
            float acc[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            if (p.bias) {
              const uint32_t bw[4] = {bias_u.x, bias_u.y, bias_u.z, bias_u.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = unpack_bf16x2(bw[k]);
                acc[2 * k] += f.x;
                acc[2 * k + 1] += f.y;
              }
            }
            if (p.epilogue == OSB_EPI_BIAS_GELU_TANH) {
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] = gelu_tanh(acc[k]);
            } else if (p.epilogue == OSB_EPI_BIAS_GATE_RES) {
              if (p.gate != nullptr) {
                float4 g0 = gate0, g1 = gate1;
                if (!gate_uniform) {   // this warp's 32 rows straddle two modulation groups (or conv row order)
                  int64_t gi = (uint32_t)row / group_rows32;
                  if (p.mod_index) gi = p.mod_index[gi];
                  const float* gate_row = p.gate + gi * p.gate_stride + n;
                  g0 = __ldg(reinterpret_cast<const float4*>(gate_row));
                  g1 = __ldg(reinterpret_cast<const float4*>(gate_row + 4));
                }
                acc[0] *= g0.x; acc[1] *= g0.y; acc[2] *= g0.z; acc[3] *= g0.w;
                acc[4] *= g1.x; acc[5] *= g1.y; acc[6] *= g1.z; acc[7] *= g1.w;
              }
              if constexpr (kRes) {
                const int lr = q * 32 + rr;            // row inside the [128 x 64] residual chunk
                const int u = half * 4 + g4;           // 16-byte unit inside the 128-byte row
                const uint4 ru = *reinterpret_cast<const uint4*>(res_slot + (lr >> 3) * 1024 + (lr & 7) * 128 +
                                                                 ((u ^ (lr & 7)) << 4));
                const uint32_t rw[4] = {ru.x, ru.y, ru.z, ru.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = unpack_bf16x2(rw[k]);
                  acc[2 * k] += f.x;
                  acc[2 * k + 1] += f.y;
                }
              }
            }
            uint4 o;
            o.x = pack_bf16x2(acc[0], acc[1]);
            o.y = pack_bf16x2(acc[2], acc[3]);
            o.z = pack_bf16x2(acc[4], acc[5]);
            o.w = pack_bf16x2(acc[6], acc[7]);
            *reinterpret_cast<uint4*>(p.D + row * p.ldd + n) = o;
          }
        }
        __syncwarp();  // staging tile (and the residual slot) are free again
        if constexpr (kRes) {
          if (lane == 0) mbar_arrive(res_empty_bar(rs));
          if (++rs == kResStages) { rs = 0; rphase ^= 1; }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  // ===================== teardown =====================
  __syncwarp();
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCta>(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int BLOCK_N, int kCta, bool kRes, bool kConv, bool kHT = false>
static int launch_kernel(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tr, const GemmEpilogueParams& p,
                         const ConvGeom& cg, int64_t tiles, cudaStream_t stream, const HeadTileParams& ht = HeadTileParams()) {
  using Cfg = GemmCfg<BLOCK_N, kCta, kRes, kHT>;
  int64_t clusters = sm_count() / kCta;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchAttribute attr[2];
  cudaLaunchConfig_t cfg = launch_config(dim3((unsigned)(clusters * kCta)), dim3(kNumThreads), Cfg::SMEM_BYTES, stream, attr, kCta);
  OSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BLOCK_N, kCta, kRes, kConv, kHT>, ta, tw, tr, p, cg, ht));
  count_launch();
  return OSB_OK;
}

// head-tile GEMM: BLOCK_N = 2 heads, CTA pairs
// OSB_HT_FAST=0 forces the general (per-row scattered store) head-tile epilogue: A/B measurements and tests
static bool ht_fast_enabled() {
  static const bool on = [] { const char* e = getenv("OSB_HT_FAST"); return !(e && e[0] == '0'); }();
  return on;
}

template <int D>
static int launch_gemm_ht(const osb_gemm_args& a, HeadTileParams ht, bool allow_fast, cudaStream_t stream) {
  constexpr int BLOCK_N = 2 * D, kCta = 2;
  using Cfg = GemmCfg<BLOCK_N, kCta, false, true>;
  CUtensorMap ta, tw;
  int rc;
  const TileMap& m = ht.map;
  ht.fast = 0;
  ht.a_rows = kBlockM;
  if (allow_fast && ht_fast_enabled()) {
    if (m.mode == 0 && m.TR == kBlockM && (m.G == 1 ? m.L % kBlockM == 0 : m.G * m.L == kBlockM)) ht.fast = 1;
    else if (m.mode == 1 && m.tps == 1 && m.S % m.G == 0 && m.G * m.L <= m.TR && (int64_t)m.L * m.G <= kBlockM) ht.fast = 2;
  }
  int64_t m_rows = a.M;
  if (ht.fast == 2) {
    // A viewed as [k][t][s][b] (strides lda, S*lda, lda... in elements): one box = (64 k, T frames, G sequences) = the rows
    // g*T + t of one packed temporal tile, in exactly that order.  Out-of-range coordinates never occur (S % G == 0).
    const uint64_t B = (uint64_t)(a.M / ((int64_t)m.S * m.T));
    const uint64_t dims[5] = {(uint64_t)a.K, (uint64_t)m.T, (uint64_t)m.S, B, 1};
    const uint64_t row_b = (uint64_t)a.lda * 2;
    const uint64_t str[4] = {(uint64_t)m.S * row_b, row_b, (uint64_t)m.T * m.S * row_b, (uint64_t)m.T * m.S * row_b * B};
    const uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)m.T, (uint32_t)m.G, 1, 1};
    const uint32_t es[5] = {1, 1, 1, 1, 1};
    rc = make_tmap_5d_bf16(&ta, a.A, dims, str, box, es);
    ht.a_rows = m.G * m.L;
    ht.groups_per_batch = m.S / m.G;
    m_rows = (int64_t)ht.tiles_total * kBlockM;   // virtual rows: one 128-row accumulator block per head tile
  } else {
    rc = make_tmap_2d_bf16(&ta, a.A, a.M, a.K, a.lda, kBlockM, kBlockK);
  }
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tw, a.W, a.N, a.K, a.ldw, Cfg::LOAD_N, kBlockK);
  if (rc) return rc;
  GemmEpilogueParams p = {};
  p.bias = static_cast<const __nv_bfloat16*>(a.bias);
  p.M = m_rows; p.N = a.N; p.K = a.K;
  p.group_rows = m_rows;
  p.epilogue = OSB_EPI_BIAS;
  const int64_t tile_m = (int64_t)kBlockM * kCta;
  const int64_t tiles = ((m_rows + tile_m - 1) / tile_m) * ((a.N + BLOCK_N - 1) / BLOCK_N);
  ConvGeom cg = {};
  return launch_kernel<BLOCK_N, kCta, false, false, true>(ta, tw, ta, p, cg, tiles, stream, ht);
}

template <int BLOCK_N, int kCta, bool kRes>
static int launch_gemm(const osb_gemm_args& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, kCta, kRes>;
  CUtensorMap ta, tw, tr;
  int rc = make_tmap_2d_bf16(&ta, a.A, a.M, a.K, a.lda, kBlockM, kBlockK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tw, a.W, a.N, a.K, a.ldw, Cfg::LOAD_N, kBlockK);
  if (rc) return rc;
  if (kRes) {
    rc = make_tmap_2d_bf16(&tr, a.R, a.M, a.N, a.ldr, kBlockM, 64);
    if (rc) return rc;
  } else {
    tr = ta;
  }

  GemmEpilogueParams p;
  p.bias = static_cast<const __nv_bfloat16*>(a.bias);
  p.D = static_cast<__nv_bfloat16*>(a.D);
  p.gate = a.gate;
  p.mod_index = a.mod_index;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.ldd = a.ldd;
  p.group_rows = a.group_rows > 0 ? a.group_rows : a.M;
  p.gate_stride = a.gate_stride;
  p.epilogue = a.epilogue;

  const int64_t tile_m = (int64_t)kBlockM * kCta;
  const int64_t tiles = ((a.M + tile_m - 1) / tile_m) * ((a.N + BLOCK_N - 1) / BLOCK_N);
  ConvGeom cg = {};
  return launch_kernel<BLOCK_N, kCta, kRes, false>(ta, tw, tr, p, cg, tiles, stream);
}

template <int BLOCK_N, int kCta, bool kRes>
static int init_one() {
  OSB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BLOCK_N, kCta, kRes, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GemmCfg<BLOCK_N, kCta, kRes>::SMEM_BYTES));
  if (kCta == 2) {  // the convolution path always pairs CTAs
    OSB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BLOCK_N, kCta, kRes, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        GemmCfg<BLOCK_N, kCta, kRes>::SMEM_BYTES));
  }
  return OSB_OK;
}

template <int BLOCK_N>
static int init_bn() {
  int rc = 0;
  if ((rc = init_one<BLOCK_N, 1, false>())) return rc;
  if ((rc = init_one<BLOCK_N, 1, true>())) return rc;
  if ((rc = init_one<BLOCK_N, 2, false>())) return rc;
  if ((rc = init_one<BLOCK_N, 2, true>())) return rc;
  return OSB_OK;
}

template <int D>
static int init_ht() {
  OSB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<2 * D, 2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GemmCfg<2 * D, 2, false, true>::SMEM_BYTES));
  return OSB_OK;
}

int gemm_init() {
  int rc = 0;
  if ((rc = init_ht<64>())) return rc;
  if ((rc = init_ht<72>())) return rc;
  if ((rc = init_ht<128>())) return rc;
  if ((rc = init_bn<64>())) return rc;
  if ((rc = init_bn<128>())) return rc;
  if ((rc = init_bn<192>())) return rc;
  if ((rc = init_bn<256>())) return rc;
  return OSB_OK;
}

// Tile width: minimise (waves of tiles over the clusters) x (tile width) / (relative efficiency of that width).
// Wider tiles amortise the per-tile epilogue and re-read A less often; the efficiencies are the measured per-FLOP rates
// of full waves on the STDiT3 shapes (profiles/r01_gemm_tune_v5.log): e.g. N = 1152 runs 5.19 -> 6 waves of 192 columns
// (1152 column-times) against 4.32 -> 5 waves of 256 (1280), and 192 wins; N = 4608 divides evenly by 256, which wins.
static int pick_block_n(int64_t M, int64_t N, int cta, bool has_res) {
  if (N <= 64) return 64;
  const int cands[3] = {256, 192, 128};
  // measured per-wave time of a 192-wide tile relative to 0.75 x a 256-wide one: 0.92-0.94 with the residual ring in the
  // epilogue (it is the epilogue that paces those tiles), 0.86-0.89 without
  const double eff[3] = {1.0, has_res ? 0.93 : 0.87, 0.78};
  const int64_t tile_m = (int64_t)kBlockM * cta;
  const int64_t clusters = sm_count() / cta;
  int best = 256;
  double best_cost = 1e300;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const int64_t tiles = ((M + tile_m - 1) / tile_m) * ((N + bn - 1) / bn);
    const int64_t waves = (tiles + clusters - 1) / clusters;
    const double cost = (double)waves * bn / eff[i];
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// ------------------------------------------------------------------------------------------
// causal 3D convolution as implicit GEMM (NDHWC, input already replicate-padded by osb_vae_prep)
// ------------------------------------------------------------------------------------------
template <int BLOCK_N, bool kRes>
static int launch_conv(const osb_conv3d_args& a, const ConvGeom& cg, const uint32_t box[5], cudaStream_t stream) {
  constexpr int kCta = 2;
  using Cfg = GemmCfg<BLOCK_N, kCta, kRes>;
  CUtensorMap ta, tw, tr;
  const uint64_t cp = a.cp;
  int rc;
  if (a.narrow) {
    // overlapping windows: 64 contiguous elements starting at every w (stride Cp elements): (kw, c) folded
    const uint64_t dims[5] = {64, (uint64_t)a.wp, (uint64_t)a.hp, (uint64_t)a.tp, (uint64_t)a.nb};
    const uint64_t str[4] = {cp * 2, (uint64_t)a.wp * cp * 2, (uint64_t)a.hp * a.wp * cp * 2,
                             (uint64_t)a.tp * a.hp * a.wp * cp * 2};
    const uint32_t es[5] = {1, (uint32_t)a.sw, (uint32_t)a.sh, (uint32_t)a.st, 1};
    rc = make_tmap_5d_bf16(&ta, a.x_pad, dims, str, box, es);
  } else {
    const uint64_t dims[5] = {cp, (uint64_t)a.wp, (uint64_t)a.hp, (uint64_t)a.tp, (uint64_t)a.nb};
    const uint64_t str[4] = {cp * 2, (uint64_t)a.wp * cp * 2, (uint64_t)a.hp * a.wp * cp * 2,
                             (uint64_t)a.tp * a.hp * a.wp * cp * 2};
    const uint32_t es[5] = {1, (uint32_t)a.sw, (uint32_t)a.sh, (uint32_t)a.st, 1};
    rc = make_tmap_5d_bf16(&ta, a.x_pad, dims, str, box, es);
  }
  if (rc) return rc;
  const int64_t K = (int64_t)(a.narrow ? a.kt * a.kh : a.kt * a.kh * a.kw * (a.cp / 64)) * 64;
  rc = make_tmap_2d_bf16(&tw, a.w, a.cout, K, K, Cfg::LOAD_N, kBlockK);
  if (rc) return rc;
  if (kRes) {
    const uint64_t dims[5] = {(uint64_t)a.cout, (uint64_t)a.w_out, (uint64_t)a.h_out, (uint64_t)a.t_out, (uint64_t)a.nb};
    const uint64_t c2 = (uint64_t)a.cout * 2;
    const uint64_t str[4] = {c2, a.w_out * c2, (uint64_t)a.h_out * a.w_out * c2, (uint64_t)a.t_out * a.h_out * a.w_out * c2};
    const uint32_t rbox[5] = {64, 1u << cg.wt_log2, 1u << cg.ht_log2, 128u >> (cg.wt_log2 + cg.ht_log2), 1};
    const uint32_t es[5] = {1, 1, 1, 1, 1};
    rc = make_tmap_5d_bf16(&tr, a.residual, dims, str, rbox, es);
    if (rc) return rc;
  } else {
    tr = tw;
  }
  GemmEpilogueParams p = {};
  p.bias = static_cast<const __nv_bfloat16*>(a.bias);
  p.D = static_cast<__nv_bfloat16*>(a.y);
  p.M = (int64_t)a.nb * a.t_out * a.h_out * a.w_out;
  p.N = a.cout;
  p.K = K;
  p.ldd = a.cout;
  p.group_rows = p.M;
  p.epilogue = kRes ? OSB_EPI_BIAS_GATE_RES : OSB_EPI_BIAS;
  const int64_t tiles = (int64_t)cg.nb * cg.tiles_t * cg.tiles_h * cg.tiles_w * ((a.cout + BLOCK_N - 1) / BLOCK_N);
  return launch_kernel<BLOCK_N, kCta, kRes, true>(ta, tw, tr, p, cg, tiles, stream);
}

}  // namespace osb

extern "C" int osb_gemm_bf16(const osb_gemm_args* args, void* stream) {
  using namespace osb;
  if (!initialised()) { set_error("osb_init() has not been called"); return OSB_ERR_NOT_INIT; }
  OSB_REQUIRE(args != nullptr, "osb_gemm_bf16: null args");
  const osb_gemm_args& a = *args;
  OSB_REQUIRE(a.A && a.W && a.D, "osb_gemm_bf16: null operand");
  OSB_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "osb_gemm_bf16: empty problem (M %lld N %lld K %lld)",
              (long long)a.M, (long long)a.N, (long long)a.K);
  OSB_REQUIRE(a.K % 8 == 0 && a.N % 8 == 0, "osb_gemm_bf16: K and N must be multiples of 8 (K %lld N %lld)",
              (long long)a.K, (long long)a.N);
  OSB_REQUIRE(a.ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(a.D) & 15) == 0,
              "osb_gemm_bf16: D must be 16-byte aligned with ldd %% 8 == 0");
  OSB_REQUIRE(a.epilogue >= OSB_EPI_BIAS && a.epilogue <= OSB_EPI_BIAS_GATE_RES,
              "osb_gemm_bf16: unknown epilogue %d", a.epilogue);
  if (a.epilogue == OSB_EPI_BIAS_GATE_RES) {
    OSB_REQUIRE(a.R == nullptr || (a.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(a.R) & 15) == 0),
                "osb_gemm_bf16: R must be 16-byte aligned with ldr %% 8 == 0");
    OSB_REQUIRE(a.gate == nullptr || (a.gate_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(a.gate) & 15) == 0),
                "osb_gemm_bf16: gate must be 16-byte aligned with gate_stride %% 4 == 0");
  }
  OSB_REQUIRE(a.bias == nullptr || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0,
              "osb_gemm_bf16: bias must be 16-byte aligned");
  int cta = a.cta_group ? a.cta_group : 2;
  OSB_REQUIRE(cta == 1 || cta == 2, "osb_gemm_bf16: cta_group must be 0, 1 or 2");
  const bool has_res = (a.epilogue == OSB_EPI_BIAS_GATE_RES) && a.R != nullptr;
  int bn = a.block_n ? a.block_n : pick_block_n(a.M, a.N, cta, has_res);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define OSB_GEMM_CASE(BN, CG)                                                   \
  if (bn == BN && cta == CG)                                                    \
    return has_res ? launch_gemm<BN, CG, true>(a, s) : launch_gemm<BN, CG, false>(a, s);
  OSB_GEMM_CASE(64, 1) OSB_GEMM_CASE(128, 1) OSB_GEMM_CASE(192, 1) OSB_GEMM_CASE(256, 1)
  OSB_GEMM_CASE(64, 2) OSB_GEMM_CASE(128, 2) OSB_GEMM_CASE(192, 2) OSB_GEMM_CASE(256, 2)
#undef OSB_GEMM_CASE
  set_error("osb_gemm_bf16: unsupported block_n %d", bn);
  return OSB_ERR_UNSUPPORTED;
}

extern "C" int osb_conv3d_ndhwc(const osb_conv3d_args* args, void* stream) {
  using namespace osb;
  if (!initialised()) { set_error("osb_init() has not been called"); return OSB_ERR_NOT_INIT; }
  OSB_REQUIRE(args != nullptr, "osb_conv3d_ndhwc: null args");
  const osb_conv3d_args& a = *args;
  OSB_REQUIRE(a.x_pad && a.w && a.y, "osb_conv3d_ndhwc: null tensor");
  OSB_REQUIRE(a.nb > 0 && a.t_out > 0 && a.h_out > 0 && a.w_out > 0 && a.cout > 0, "osb_conv3d_ndhwc: empty output");
  OSB_REQUIRE(a.cout % 8 == 0, "osb_conv3d_ndhwc: Cout must be a multiple of 8 (pad the weights), got %d", a.cout);
  OSB_REQUIRE(a.st >= 1 && a.st <= 2 && a.sh >= 1 && a.sh <= 2 && a.sw >= 1 && a.sw <= 2, "osb_conv3d_ndhwc: strides must be 1 or 2");
  OSB_REQUIRE(a.kt >= 1 && a.kh >= 1 && a.kw >= 1 && a.kt <= 3 && a.kh <= 3 && a.kw <= 3, "osb_conv3d_ndhwc: taps must be 1..3");
  if (a.narrow) {
    OSB_REQUIRE((a.cp == 8 || a.cp == 16) && a.kw * a.cp <= 64,
                "osb_conv3d_ndhwc: narrow mode needs Cp in {8,16} with kw*Cp <= 64 (Cp %d kw %d)", a.cp, a.kw);
  } else {
    OSB_REQUIRE(a.cp % 64 == 0, "osb_conv3d_ndhwc: Cp must be a multiple of 64 (or use narrow mode), got %d", a.cp);
  }
  OSB_REQUIRE((a.t_out - 1) * a.st + a.kt <= a.tp && (a.h_out - 1) * a.sh + a.kh <= a.hp && (a.w_out - 1) * a.sw + a.kw <= a.wp,
              "osb_conv3d_ndhwc: padded input [%d,%d,%d] too small for output [%d,%d,%d]", a.tp, a.hp, a.wp, a.t_out,
              a.h_out, a.w_out);
  OSB_REQUIRE(((reinterpret_cast<uintptr_t>(a.x_pad) | reinterpret_cast<uintptr_t>(a.w) | reinterpret_cast<uintptr_t>(a.y) |
                reinterpret_cast<uintptr_t>(a.bias) | reinterpret_cast<uintptr_t>(a.residual)) & 15) == 0,
              "osb_conv3d_ndhwc: tensors must be 16-byte aligned");

  // box of 128 output positions per CTA (pairs stack along H): minimise padded work, prefer wide W
  ConvGeom cg = {};
  double best = 1e30;
  for (int wl = 7; wl >= 3; --wl) {
    for (int hl = 0; wl + hl <= 7; ++hl) {
      const int Wt = 1 << wl, Ht = 1 << hl, Tt = 128 >> (wl + hl);
      const int64_t tw = (a.w_out + Wt - 1) / Wt, th = (a.h_out + 2 * Ht - 1) / (2 * Ht), tt = (a.t_out + Tt - 1) / Tt;
      const double work = (double)(tw * Wt) * (double)(th * 2 * Ht) * (double)(tt * Tt);
      if (work < best * 0.999) {
        best = work;
        cg.wt_log2 = wl; cg.ht_log2 = hl;
        cg.tiles_w = (int)tw; cg.tiles_h = (int)th; cg.tiles_t = (int)tt;
      }
    }
  }
  cg.nb = a.nb;
  cg.w_out = a.w_out; cg.h_out = a.h_out; cg.t_out = a.t_out;
  cg.sw = a.sw; cg.sh = a.sh; cg.st = a.st;
  cg.kw_n = a.narrow ? 1 : a.kw; cg.kh_n = a.kh;
  cg.cin_chunks = a.narrow ? 1 : a.cp / 64;
  const uint32_t box[5] = {64, (uint32_t)((1 << cg.wt_log2) * a.sw), (uint32_t)((1 << cg.ht_log2) * a.sh),
                           (uint32_t)((128 >> (cg.wt_log2 + cg.ht_log2)) * a.st), 1};
  int bn = a.block_n;
  if (bn == 0) bn = a.cout <= 64 ? 64 : (a.cout <= 128 ? 128 : (a.cout % 256 == 0 ? 256 : (a.cout % 192 == 0 ? 192 : 256)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool has_res = a.residual != nullptr;
#define OSB_CONV_CASE(BN) \
  if (bn == BN) return has_res ? launch_conv<BN, true>(a, cg, box, s) : launch_conv<BN, false>(a, cg, box, s);
  OSB_CONV_CASE(64) OSB_CONV_CASE(128) OSB_CONV_CASE(192) OSB_CONV_CASE(256)
#undef OSB_CONV_CASE
  set_error("osb_conv3d_ndhwc: unsupported block_n %d", bn);
  return OSB_ERR_UNSUPPORTED;
}

extern "C" int osb_gemm_head_tiles(const osb_gemm_args* args, const osb_head_tiles_args* targs, void* stream) {
  using namespace osb;
  if (!initialised()) { set_error("osb_init() has not been called"); return OSB_ERR_NOT_INIT; }
  OSB_REQUIRE(args != nullptr && targs != nullptr, "osb_gemm_head_tiles: null args");
  const osb_gemm_args& a = *args;
  const osb_head_tiles_args& t = *targs;
  OSB_REQUIRE(a.A && a.W && t.tiles, "osb_gemm_head_tiles: null operand");
  OSB_REQUIRE(a.M > 0 && a.M < (1ll << 31) && a.N > 0 && a.N < (1ll << 31) && a.K > 0 && a.K % 8 == 0, "osb_gemm_head_tiles: bad problem (M %lld N %lld K %lld)",
              (long long)a.M, (long long)a.N, (long long)a.K);
  const int D = t.head_dim;
  OSB_REQUIRE(D == 64 || D == 72 || D == 128, "osb_gemm_head_tiles: head_dim %d not built (64, 72, 128)", D);
  OSB_REQUIRE(t.num_heads > 0 && t.num_heads % 2 == 0 && a.N % ((int64_t)t.num_heads * D) == 0,
              "osb_gemm_head_tiles: N (%lld) must be a multiple of num_heads*head_dim with an even head count (%d x %d)",
              (long long)a.N, t.num_heads, D);
  OSB_REQUIRE(t.nkinds >= 1 && t.nkinds <= 4, "osb_gemm_head_tiles: nkinds must be 1..4");
  const osb_tile_map& m = t.map;
  OSB_REQUIRE(m.mode == 0 || m.mode == 1, "osb_gemm_head_tiles: unknown tile map mode %d", m.mode);
  OSB_REQUIRE(m.L > 0 && m.G >= 1 && m.tile_rows > 0 && m.tile_rows <= 128 && m.tile_rows % 8 == 0,
              "osb_gemm_head_tiles: bad tile map (L %d G %d rows %d)", m.L, m.G, m.tile_rows);
  OSB_REQUIRE(m.G == 1 ? (m.tps == (m.L + m.tile_rows - 1) / m.tile_rows) : (m.G * m.L <= m.tile_rows && m.tps == 1),
              "osb_gemm_head_tiles: tile map inconsistent (L %d G %d tps %d rows %d)", m.L, m.G, m.tps, m.tile_rows);
  OSB_REQUIRE(m.mode == 0 ? (a.M % m.L == 0) : (m.S > 0 && m.T == m.L && a.M % ((int64_t)m.S * m.T) == 0),
              "osb_gemm_head_tiles: M (%lld) is not a whole number of sequences", (long long)a.M);
  OSB_REQUIRE((reinterpret_cast<uintptr_t>(t.tiles) & 15) == 0 && t.kind_stride % 16 == 0 && t.head_stride % 16 == 0,
              "osb_gemm_head_tiles: tile buffer must be 16-byte aligned");
  OSB_REQUIRE(a.bias == nullptr || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "osb_gemm_head_tiles: bias must be 16-byte aligned");
  const int dp = (D + 15) / 16 * 16;
  const int64_t tile_bytes = (int64_t)m.tile_rows * dp * 2;
  const int64_t tph = osb_head_tiles_per_head(&m, a.M);
  OSB_REQUIRE(t.head_stride >= tph * tile_bytes, "osb_gemm_head_tiles: head_stride %lld < %lld tiles of %lld bytes",
              (long long)t.head_stride, (long long)tph, (long long)tile_bytes);
  HeadTileParams ht = {};
  ht.base = static_cast<uint8_t*>(t.tiles);
  ht.kind_stride = t.kind_stride; ht.head_stride = t.head_stride;
  ht.map.mode = m.mode; ht.map.L = m.L; ht.map.S = m.S; ht.map.T = m.T; ht.map.G = m.G; ht.map.tps = m.tps; ht.map.TR = m.tile_rows;
  ht.tile_bytes = (int32_t)tile_bytes;
  ht.tiles_total = (int32_t)tph;
  OSB_REQUIRE(tph < (1ll << 31), "osb_gemm_head_tiles: too many tiles");
  ht.heads = t.num_heads; ht.nkinds = t.nkinds;
  ht.norm_mask = t.norm_mask; ht.rope_mask = t.rope_mask;
  for (int k = 0; k < 4; ++k) {
    ht.norm_w[k] = static_cast<const __nv_bfloat16*>(t.norm_w[k]);
    OSB_REQUIRE(!((t.norm_mask >> k) & 1u) || (k < t.nkinds && t.norm_w[k] != nullptr && (reinterpret_cast<uintptr_t>(t.norm_w[k]) & 15) == 0),
                "osb_gemm_head_tiles: kind %d has RMSNorm enabled but no (16-byte aligned) weight", k);
  }
  ht.eps = t.norm_eps;
  ht.cos = t.rope_cos; ht.sin = t.rope_sin;
  OSB_REQUIRE(t.rope_mask == 0 || (t.rope_cos && t.rope_sin && ((reinterpret_cast<uintptr_t>(t.rope_cos) | reinterpret_cast<uintptr_t>(t.rope_sin)) & 15) == 0),
              "osb_gemm_head_tiles: RoPE enabled but the cos / sin tables are missing or not 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool allow_fast = (t.reserved & 1) == 0;   // bit 0 of `reserved`: force the general (per-row store) epilogue
  if (D == 64) return launch_gemm_ht<64>(a, ht, allow_fast, s);
  if (D == 72) return launch_gemm_ht<72>(a, ht, allow_fast, s);
  return launch_gemm_ht<128>(a, ht, allow_fast, s);
}
