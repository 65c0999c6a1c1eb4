// HBM-bound helpers of the causal 3D VAE path (channels-last NDHWC, bf16 storage, fp32 math).
//
//  osb_group_stats : GroupNorm statistics, one read of the tensor (fp32 block partials, fp64 across blocks)
//  osb_vae_prep    : GroupNorm-apply + SiLU + nearest upsample (first-frame rule) + replicate padding in ONE
//                    pass that writes the padded convolution input (never materialising the intermediate
//                    normalised / activated / upsampled / padded tensors the reference creates one by one)
//
// Replaces: torch.nn.GroupNorm + SiLU (unet_causal_3d_blocks.py:246-250; vae.py:115,146,229,234),
// chunk_nearest_interpolate / UpsampleCausal3D (:41-49,136-150) and F.pad(mode="replicate") (:95).
#include "common.cuh"

namespace osb {

constexpr int kStatsThreads = 256;
constexpr int kStatsPositionsPerBlock = 2048;

// grid (position chunks, nb); each thread owns one 8-channel vector index (256 % (C/8) == 0).
// Deterministic: per-thread partials -> shared memory -> fixed-order sum per group -> one fp32 partial per
// (chunk, n, group); the finalize kernel adds the chunk partials in a fixed order in fp64.  No atomics.
__global__ void __launch_bounds__(kStatsThreads)
group_stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t positions, int C, int groups, float* __restrict__ partial) {
  __shared__ float sp[kStatsThreads][17];
  const int n = blockIdx.y;
  const int vecs = C >> 3;
  const int cg = C / groups;
  const int64_t p0 = (int64_t)blockIdx.x * kStatsPositionsPerBlock;
  const int64_t p1 = p0 + kStatsPositionsPerBlock < positions ? p0 + kStatsPositionsPerBlock : positions;
  const uint4* base = reinterpret_cast<const uint4*>(x + (int64_t)n * positions * C);
  float s[8], q[8], k[8];
  // sums are taken relative to a per-(n, group) shift K = the group's first element: E[(x-K)^2] - E[x-K]^2 cancels on
  // the scale of |mean - K| ~ std instead of |mean| (a group with |mean| >> std would lose its variance in fp32)
  const int v0 = threadIdx.x % vecs;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s[e] = 0.f; q[e] = 0.f;
    k[e] = __bfloat162float(x[(int64_t)n * positions * C + ((v0 * 8 + e) / cg) * cg]);
  }
  for (int64_t i = p0 * vecs + threadIdx.x; i < p1 * vecs; i += kStatsThreads) {
    uint4 t;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(base + i));
    const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(tw[e]);
      const float d0 = f.x - k[2 * e], d1 = f.y - k[2 * e + 1];
      s[2 * e] += d0; q[2 * e] += d0 * d0;
      s[2 * e + 1] += d1; q[2 * e + 1] += d1 * d1;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { sp[threadIdx.x][e] = s[e]; sp[threadIdx.x][8 + e] = q[e]; }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += kStatsThreads) {
    float ss = 0.f, qq = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      const int v = c >> 3, e = c & 7;
      for (int t = v; t < kStatsThreads; t += vecs) { ss += sp[t][e]; qq += sp[t][8 + e]; }
    }
    float* out = partial + (((int64_t)blockIdx.x * gridDim.y + n) * groups + g) * 2;
    out[0] = ss;
    out[1] = qq;
  }
}

// one block per (n, group): fixed-order strided sums + tree in fp64
__global__ void __launch_bounds__(256)
group_stats_finalize_kernel(const float* __restrict__ partial, float* __restrict__ mean_rstd, int64_t chunks, int ng,
                            double count, float eps, const __nv_bfloat16* __restrict__ x, int64_t positions, int C, int groups) {
  __shared__ double rs[256], rq[256];
  const int i = blockIdx.x;  // n * groups + g
  double s = 0.0, q = 0.0;
  for (int64_t c = threadIdx.x; c < chunks; c += 256) {
    const float2 v = *reinterpret_cast<const float2*>(partial + (c * ng + i) * 2);
    s += (double)v.x;
    q += (double)v.y;
  }
  rs[threadIdx.x] = s; rq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { rs[threadIdx.x] += rs[threadIdx.x + o]; rq[threadIdx.x] += rq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int n = i / groups, g = i - n * groups;
    const double shift = (double)__bfloat162float(x[(int64_t)n * positions * C + g * (C / groups)]);   // the kernel's K
    const double dm = rs[0] / count;
    double var = rq[0] / count - dm * dm;
    if (var < 0.0) var = 0.0;
    const double mean = shift + dm;
    mean_rstd[2 * i] = (float)mean;
    mean_rstd[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

struct PrepParams {
  const __nv_bfloat16* x; __nv_bfloat16* y;
  const float* mean_rstd; const __nv_bfloat16* gamma; const __nv_bfloat16* beta;
  int nb, t, h, w, c, groups, silu, ft, fh, fw, pad_t, pad_h, pad_w, cp;
  int tu, hu, wu;   // upsampled dims
  int tp, hp, wp;   // padded output dims
  int64_t total;    // nb*tp*hp*wp*(cp/8)
};

__global__ void __launch_bounds__(256) vae_prep_kernel(const PrepParams p) {
  const int vp = p.cp >> 3;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vp);
    int64_t r = idx / vp;
    const int wi = (int)(r % p.wp); r /= p.wp;
    const int hi = (int)(r % p.hp); r /= p.hp;
    const int ti = (int)(r % p.tp);
    const int n = (int)(r / p.tp);
    uint4 o = make_uint4(0, 0, 0, 0);
    if (v * 8 < p.c) {
      // replicate padding = clamp in the (upsampled) source; T is padded in front only (causal)
      int tu = ti - p.pad_t; tu = tu < 0 ? 0 : (tu >= p.tu ? p.tu - 1 : tu);
      int hu = hi - p.pad_h; hu = hu < 0 ? 0 : (hu >= p.hu ? p.hu - 1 : hu);
      int wu = wi - p.pad_w; wu = wu < 0 ? 0 : (wu >= p.wu ? p.wu - 1 : wu);
      // nearest upsample; frame 0 maps to source frame 0, frames 1.. to 1 + (t-1)/ft  (T' = 1 + ft*(T-1))
      const int ts = (p.ft == 1 || tu == 0) ? tu : 1 + (tu - 1) / p.ft;
      const int hs = hu / p.fh, ws = wu / p.fw;
      const int64_t src = ((((int64_t)n * p.t + ts) * p.h + hs) * p.w + ws) * p.c + v * 8;
      o = __ldg(reinterpret_cast<const uint4*>(p.x + src));
      if (p.mean_rstd != nullptr || p.silu) {
        const uint32_t tw[4] = {o.x, o.y, o.z, o.w};
        float f[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 u = unpack_bf16x2(tw[e]);
          f[2 * e] = u.x; f[2 * e + 1] = u.y;
        }
        if (p.mean_rstd != nullptr) {
          const int cg = p.c / p.groups;
          const uint4 gu = __ldg(reinterpret_cast<const uint4*>(p.gamma + v * 8));
          const uint4 bu = __ldg(reinterpret_cast<const uint4*>(p.beta + v * 8));
          const uint32_t gw[4] = {gu.x, gu.y, gu.z, gu.w}, bw[4] = {bu.x, bu.y, bu.z, bu.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int g = (v * 8 + e) / cg;
            const float2 mr = __ldg(reinterpret_cast<const float2*>(p.mean_rstd) + (int64_t)n * p.groups + g);
            const float2 ga = unpack_bf16x2(gw[e >> 1]), be = unpack_bf16x2(bw[e >> 1]);
            const float gam = (e & 1) ? ga.y : ga.x, bet = (e & 1) ? be.y : be.x;
            f[e] = (f[e] - mr.x) * mr.y * gam + bet;
          }
        }
        if (p.silu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __fdividef(f[e], 1.0f + __expf(-f[e]));
        }
        o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
      }
    }
    reinterpret_cast<uint4*>(p.y)[idx] = o;
  }
}

}  // namespace osb

extern "C" int64_t osb_group_stats_workspace_bytes(int64_t nb, int64_t positions, int32_t groups) {
  const int64_t chunks = (positions + osb::kStatsPositionsPerBlock - 1) / osb::kStatsPositionsPerBlock;
  return chunks * nb * groups * 2 * (int64_t)sizeof(float);
}

extern "C" int osb_group_stats(const void* x, int64_t nb, int64_t positions, int32_t C, int32_t groups, float eps,
                               void* workspace, int64_t workspace_bytes, float* mean_rstd, void* stream) {
  using namespace osb;
  if (!initialised()) { set_error("osb_init() has not been called"); return OSB_ERR_NOT_INIT; }
  OSB_REQUIRE(x && workspace && mean_rstd, "osb_group_stats: null tensor");
  OSB_REQUIRE(nb > 0 && positions > 0 && C > 0 && groups > 0 && C % groups == 0, "osb_group_stats: bad shape");
  OSB_REQUIRE(C % 8 == 0 && kStatsThreads % (C / 8) == 0, "osb_group_stats: C/8 must divide %d (C = %d)", kStatsThreads, C);
  OSB_REQUIRE(nb <= 65535 && groups <= 1024, "osb_group_stats: batch / groups too large");
  OSB_REQUIRE(workspace_bytes >= osb_group_stats_workspace_bytes(nb, positions, groups),
              "osb_group_stats: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
              (long long)osb_group_stats_workspace_bytes(nb, positions, groups));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int ng = (int)(nb * groups);
  const int64_t chunks = (positions + kStatsPositionsPerBlock - 1) / kStatsPositionsPerBlock;
  group_stats_kernel<<<dim3((unsigned)chunks, (unsigned)nb), kStatsThreads, 0, s>>>(
      static_cast<const __nv_bfloat16*>(x), positions, C, groups, static_cast<float*>(workspace));
  group_stats_finalize_kernel<<<ng, 256, 0, s>>>(static_cast<const float*>(workspace), mean_rstd, chunks, ng,
                                                 (double)positions * (C / groups), eps,
                                                 static_cast<const __nv_bfloat16*>(x), positions, C, groups);
  OSB_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return OSB_OK;
}

extern "C" int osb_vae_prep(const osb_vae_prep_args* a, void* stream) {
  using namespace osb;
  if (!initialised()) { set_error("osb_init() has not been called"); return OSB_ERR_NOT_INIT; }
  OSB_REQUIRE(a && a->x && a->y, "osb_vae_prep: null tensor");
  OSB_REQUIRE(a->c % 8 == 0 && a->cp % 8 == 0 && a->cp >= a->c, "osb_vae_prep: channels must be multiples of 8 (c %d cp %d)", a->c, a->cp);
  OSB_REQUIRE(a->ft >= 1 && a->ft <= 2 && a->fh >= 1 && a->fh <= 2 && a->fw >= 1 && a->fw <= 2, "osb_vae_prep: upsample factors must be 1 or 2");
  OSB_REQUIRE(a->mean_rstd == nullptr || (a->gamma && a->beta && a->groups > 0 && a->c % a->groups == 0),
              "osb_vae_prep: GroupNorm needs gamma, beta and a valid group count");
  PrepParams p;
  p.x = static_cast<const __nv_bfloat16*>(a->x); p.y = static_cast<__nv_bfloat16*>(a->y);
  p.mean_rstd = a->mean_rstd;
  p.gamma = static_cast<const __nv_bfloat16*>(a->gamma); p.beta = static_cast<const __nv_bfloat16*>(a->beta);
  p.nb = a->nb; p.t = a->t; p.h = a->h; p.w = a->w; p.c = a->c; p.groups = a->groups > 0 ? a->groups : 1; p.silu = a->silu;
  p.ft = a->ft; p.fh = a->fh; p.fw = a->fw; p.pad_t = a->pad_t; p.pad_h = a->pad_h; p.pad_w = a->pad_w; p.cp = a->cp;
  p.tu = a->ft == 1 ? a->t : 1 + a->ft * (a->t - 1);
  p.hu = a->h * a->fh; p.wu = a->w * a->fw;
  p.tp = p.tu + a->pad_t; p.hp = p.hu + 2 * a->pad_h; p.wp = p.wu + 2 * a->pad_w;
  p.total = (int64_t)p.nb * p.tp * p.hp * p.wp * (p.cp / 8);
  int64_t blocks = (p.total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  vae_prep_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  OSB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OSB_OK;
}
