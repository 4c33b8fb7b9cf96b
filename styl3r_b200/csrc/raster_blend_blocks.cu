// Rasterizer stage 5, warp-granular version: per-tile front-to-back alpha blend (forward) over the per-block survivor
// lists the tile sort wrote (raster_sort.cu:build_block_lists).
//
// Work unit = one 8x4-pixel block of one (view, tile): a persistent WARP fetches units from the device-side queue
// (tiles in longest-processing-time-first order, eight blocks each), gathers its own survivors' 48-byte records from
// the tile's sorted range with cp.async (32 records per group, double-buffered in the warp's private shared memory)
// and composites them four at a time with the same arithmetic, guard bands and exact re-decisions as the tile-granular
// kernel (raster_blend.cu).  Compared with that kernel there is no cull, no shared TMA ring and no CTA-wide barrier:
// warps of the same tile are independent, so a busy block no longer holds the other seven warps (round-2 profile of the
// tile-granular kernel: 13 % of all warp samples sat in the end-of-tile barrier, 8 % waited on the shared ring, 11 % of
// the instructions were the cull).
//
// Semantics: oracle/raster_oracle.c:s3r_oracle_render; results are bit-identical to
// raster_blend.cu (same operations in the same order per pixel).
#include "s3r_common.cuh"

#define BB_ALPHA_MIN (1.0f / 255.0f)
#define BB_ALPHA_LO (BB_ALPHA_MIN * (1.0f - S3R_ALPHA_BAND))
#define BB_ALPHA_HI (BB_ALPHA_MIN * (1.0f + S3R_ALPHA_BAND))
#define BBK_WARPS 8                 // warps per CTA (all of them consumers)
#define BBK_THREADS (32 * BBK_WARPS)
#ifndef BBK_RPL
#define BBK_RPL 1                   // records gathered per lane and group (2 measured the same: 0.112 vs 0.113 pipelined)
#endif
#define BBK_GROUP (32 * BBK_RPL)    // records per group
#define BBK_U 4                     // survivors per batch
#ifndef BBK_MINB
#define BBK_MINB 3
#endif

__device__ __forceinline__ uint32_t k_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t k_pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void k_unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t k_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t k_mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t k_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float k_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 k_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 k_lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void k_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void k_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void k_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __noinline__ bool k_exact_decide(const float4 co, float dx, float dy, float* alpha_out) {
  const float q = __fadd_rn(__fmul_rn(__fmul_rn(co.x, dx), dx), __fmul_rn(__fmul_rn(co.z, dy), dy));
  const float power = __fsub_rn(__fmul_rn(-0.5f, q), __fmul_rn(__fmul_rn(co.y, dx), dy));
  if (power > 0.0f) return false;
  const float a = fminf(0.99f, __fmul_rn(co.w, (float)exp((double)power)));
  *alpha_out = a;
  return a >= BB_ALPHA_MIN;
}

struct __align__(16) BlendWarpSmem {
  float4 rec[2][BBK_GROUP * 3];  // two groups in flight; a partial last group is padded with dummy records in place
};

template <bool kHasNT>
__global__ void __launch_bounds__(BBK_THREADS, BBK_MINB) s3r_blend_blocks_fwd_kernel(
    int W, int H, int P, int tiles_x, int tiles, int only_tile, uint32_t n_units, const uint32_t* __restrict__ work_order,
    unsigned* __restrict__ counters, const uint2* __restrict__ ranges, const float4* __restrict__ records,
    const uint32_t* __restrict__ blists, const uint32_t* __restrict__ bcounts, const uint32_t* __restrict__ point_list,
    const float4* __restrict__ conic_opacity, const float* __restrict__ background, float* __restrict__ out_color,
    float* __restrict__ out_depth, float* __restrict__ out_opacity, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ n_contrib_blk, int32_t* __restrict__ n_touched) {
  __shared__ BlendWarpSmem sm[BBK_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t buf0 = k_smem_u32(&sm[w].rec[0][0]);
  constexpr uint32_t kBufBytes = BBK_GROUP * S3R_REC_BYTES;
  s3r_grid_dependency_sync();

  const uint32_t n_warps = gridDim.x * BBK_WARPS;
  for (bool first = true;; first = false) {
    // first unit: by the warp's own index (no atomic round trip in front of a single view's only unit), strided so that
    // the eight warps of a CTA start on blocks of eight different tiles across the weight-ordered queue (the busiest
    // blocks then share their SM with light ones), afterwards the queue
    uint32_t unit = w * gridDim.x + blockIdx.x;
    if (!first) {
      if (lane == 0) unit = n_warps + atomicAdd(&counters[1], 1u);
      unit = __shfl_sync(0xffffffffu, unit, 0);
    }
    if (unit >= n_units) break;
    const uint32_t vt = only_tile >= 0 ? (uint32_t)only_tile : work_order[unit >> 3];
    const int blk = unit & 7;
    const int view = vt / tiles, tile = vt % tiles;
    const uint2 rg = ranges[vt];
    const uint32_t n = rg.y - rg.x;
    const uint32_t cnt = bcounts[(size_t)vt * 8 + blk];
    const uint32_t* list = blists + (size_t)rg.x * 8 + (size_t)blk * n;
    const float4* src = records + (size_t)rg.x * 3;

    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int X0 = tx * S3R_TILE + (blk & 1) * 8, Y0 = ty * S3R_TILE + (blk >> 1) * 4;
    const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;

    float T = 1.0f;
    uint64_t Crg = k_pack2(0.f, 0.f), Cbd = k_pack2(0.f, 0.f);
    uint64_t negpix = k_pack2(inside ? -pxf : -__int_as_float(0x7f800000), -pyf);
    bool alive = inside;
    int lastk = -1;  // position in the block list of the last composited splat
    bool warp_done = !__any_sync(0xffffffffu, inside);
    const uint32_t ngroups = warp_done ? 0u : (cnt + BBK_GROUP - 1) / BBK_GROUP;

    // gather of one group: lane l fetches survivor g*32 + l (three 16-byte cp.async); lanes past the list's end store
    // the dummy record (alpha = 0 for every pixel), which pads the last batch.  The list index is loaded one group
    // earlier than the records it addresses (a warp issues in order: a dependent address would stall the whole warp
    // for an L2 round trip per group).
    struct Idx { uint32_t v[BBK_RPL]; };
    auto load_idx = [&](uint32_t g) -> Idx {
      Idx r;
#pragma unroll
      for (int e = 0; e < BBK_RPL; e++) {
        const uint32_t j = g * BBK_GROUP + e * 32 + lane;
        r.v[e] = j < cnt ? list[j] : 0u;
      }
      return r;
    };
    auto gather = [&](uint32_t g, const Idx& idx) {
#pragma unroll
      for (int e = 0; e < BBK_RPL; e++) {
        const uint32_t slot = e * 32 + lane;
        const uint32_t j = g * BBK_GROUP + slot;
        const uint32_t dst = buf0 + (g & 1u) * kBufBytes + slot * S3R_REC_BYTES;
        if (j < cnt) {
          const float4* r = src + (size_t)idx.v[e] * 3;
          k_cp_async16(dst, r);
          k_cp_async16(dst + 16, r + 1);
          k_cp_async16(dst + 32, r + 2);
        } else if (j < ((cnt + BBK_U - 1) / BBK_U) * BBK_U) {  // only the last batch's padding is ever read
          float4* d = &sm[w].rec[g & 1u][slot * 3];
          d[0] = make_float4(3e9f, 3e9f, 0.0f, -1.0f);
          d[1] = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
          d[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
      }
      k_cp_commit();
    };
    Idx idx_next;
#pragma unroll
    for (int e = 0; e < BBK_RPL; e++) idx_next.v[e] = 0u;
    if (ngroups) {
      gather(0, load_idx(0));
      idx_next = load_idx(1);
    }

    for (uint32_t g = 0; g < ngroups; g++) {
      if (g + 1 < ngroups) {
        gather(g + 1, idx_next);
        idx_next = load_idx(g + 2);
        k_cp_wait<1>();
      } else {
        k_cp_wait<0>();
      }
      __syncwarp();
      const uint32_t gcnt = min((uint32_t)BBK_GROUP, cnt - g * BBK_GROUP);
      const uint32_t gbase = buf0 + (g & 1u) * kBufBytes;
#pragma unroll 1
      for (uint32_t k = 0; k < gcnt; k += BBK_U) {
        const uint32_t a0 = gbase + k * S3R_REC_BYTES;
        float alpha[BBK_U];
        uint64_t crg[BBK_U], cbd[BBK_U];
        float t[BBK_U + 1];
        const int lastk_in = lastk;
        const int kpos = (int)(g * BBK_GROUP + k);
        bool near_thr = false;
#pragma unroll
        for (int u = 0; u < BBK_U; u++) {
          const float4 r0 = k_lds128(a0 + u * S3R_REC_BYTES);
          const float4 r1 = k_lds128(a0 + u * S3R_REC_BYTES + 16);
          const float2 r2 = k_lds64(a0 + u * S3R_REC_BYTES + 32);
          float dx, dy;
          k_unpack2(k_add2(k_pack2(r0.x, r0.y), negpix), dx, dy);
          float bdy, cdy;
          k_unpack2(k_mul2(k_pack2(r0.z, r0.w), k_pack2(dy, dy)), bdy, cdy);
          const float l2g = fmaf(dx, fmaf(r1.x, dx, bdy), cdy * dy);
          const float a = fminf(0.99f, r1.y * k_exp2(l2g));
          const bool keep = a >= BB_ALPHA_HI;
          near_thr = near_thr || ((a >= BB_ALPHA_LO) && !keep) || (l2g > -S3R_PZERO_BAND);
          alpha[u] = keep ? a : 0.0f;
          lastk = keep ? kpos + u : lastk;
          crg[u] = k_pack2(r1.z, r1.w);
          cbd[u] = k_pack2(r2.x, r2.y);
        }
        const uint64_t a01 = k_pack2(alpha[0], alpha[1]), a23 = k_pack2(alpha[2], alpha[3]);
        float om[BBK_U];
        k_unpack2(k_fma2(a01, k_pack2(-1.0f, -1.0f), k_pack2(1.0f, 1.0f)), om[0], om[1]);
        k_unpack2(k_fma2(a23, k_pack2(-1.0f, -1.0f), k_pack2(1.0f, 1.0f)), om[2], om[3]);
        t[0] = T;
#pragma unroll
        for (int u = 0; u < BBK_U; u++) t[u + 1] = t[u] * om[u];
        auto composite_all = [&](uint64_t p01, uint64_t p23) {
          float wgt[BBK_U];
          k_unpack2(k_mul2(p01, k_pack2(t[0], t[1])), wgt[0], wgt[1]);
          k_unpack2(k_mul2(p23, k_pack2(t[2], t[3])), wgt[2], wgt[3]);
#pragma unroll
          for (int u = 0; u < BBK_U; u++) {
            const uint64_t w2 = k_pack2(wgt[u], wgt[u]);
            Crg = k_fma2(crg[u], w2, Crg);
            Cbd = k_fma2(cbd[u], w2, Cbd);
            if (kHasNT) {
              if (alpha[u] > 0.0f && t[u + 1] > 0.5f)
                atomicAdd(&n_touched[(size_t)view * P + point_list[(size_t)rg.x + list[kpos + u]]], 1);
            }
          }
          T = t[BBK_U];
        };
        if (!__any_sync(0xffffffffu, near_thr || t[BBK_U] < 0.0001f)) {
          composite_all(a01, a23);
        } else {
          if (__any_sync(0xffffffffu, near_thr)) {
            lastk = lastk_in;
#pragma unroll
            for (int u = 0; u < BBK_U; u++) {
              if (k + u < gcnt) {
                const float4 r0 = k_lds128(a0 + u * S3R_REC_BYTES);
                const float4 r1 = k_lds128(a0 + u * S3R_REC_BYTES + 16);
                float dx, dy;
                k_unpack2(k_add2(k_pack2(r0.x, r0.y), negpix), dx, dy);
                const float l2g = fmaf(dx, fmaf(r1.x, dx, r0.z * dy), (r0.w * dy) * dy);
                const float a = fminf(0.99f, r1.y * k_exp2(l2g));
                if ((a >= BB_ALPHA_LO && a < BB_ALPHA_HI) || l2g > -S3R_PZERO_BAND) {
                  const uint32_t gid = point_list[(size_t)rg.x + list[kpos + u]];
                  float ax = a;
                  const bool kx = k_exact_decide(conic_opacity[(size_t)view * P + gid], dx, dy, &ax);
                  alpha[u] = (kx && alive) ? ax : 0.0f;
                }
              }
              lastk = alpha[u] > 0.0f ? kpos + u : lastk;
            }
#pragma unroll
            for (int u = 0; u < BBK_U; u++) t[u + 1] = t[u] * (1.0f - alpha[u]);
          }
          if (!__any_sync(0xffffffffu, t[BBK_U] < 0.0001f)) {
            composite_all(k_pack2(alpha[0], alpha[1]), k_pack2(alpha[2], alpha[3]));
          } else {
            lastk = lastk_in;
#pragma unroll
            for (int u = 0; u < BBK_U; u++) {
              const bool stop = t[u + 1] < 0.0001f;
              const bool acc = !stop && alpha[u] > 0.0f;
              const float wgt = stop ? 0.0f : alpha[u] * t[u];
              const uint64_t w2 = k_pack2(wgt, wgt);
              Crg = k_fma2(crg[u], w2, Crg);
              Cbd = k_fma2(cbd[u], w2, Cbd);
              if (kHasNT) {
                if (acc && t[u + 1] > 0.5f)
                  atomicAdd(&n_touched[(size_t)view * P + point_list[(size_t)rg.x + list[kpos + u]]], 1);
              }
              T = stop ? T : t[u + 1];
              lastk = acc ? kpos + u : lastk;
            }
            if (t[BBK_U] < 0.0001f) {
              alive = false;
              negpix = k_pack2(-__int_as_float(0x7f800000), -pyf);
            }
            if (!__any_sync(0xffffffffu, alive)) {
              warp_done = true;
              break;
            }
          }
        }
      }
      __syncwarp();  // the group's buffer may be refilled two iterations from now
      if (warp_done) break;
    }
    k_cp_wait<0>();  // a gather issued ahead of an early exit must land before the buffers are reused
    __syncwarp();
    if (inside) {
      const size_t HW = (size_t)H * W;
      const size_t pix = (size_t)py * W + px;
      const float* bg = background + view * 3;
      float* oc = out_color + (size_t)view * 3 * HW;
      float Cr, Cg, Cb, D;
      k_unpack2(Crg, Cr, Cg);
      k_unpack2(Cbd, Cb, D);
      oc[pix] = Cr + T * bg[0];
      oc[HW + pix] = Cg + T * bg[1];
      oc[2 * HW + pix] = Cb + T * bg[2];
      out_depth[(size_t)view * HW + pix] = D;
      out_opacity[(size_t)view * HW + pix] = 1.0f - T;
      final_T[(size_t)view * HW + pix] = T;
      n_contrib[(size_t)view * HW + pix] = lastk >= 0 ? list[lastk] + 1u : 0u;
      n_contrib_blk[(size_t)view * HW + pix] = (uint32_t)(lastk + 1);  // where the backward walk of this pixel starts
    }
  }
  // the last CTA to leave rewinds the queue for the next launch on this state buffer
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&counters[2], 1u) == gridDim.x - 1) {
      counters[1] = 0u;
      counters[2] = 0u;
      counters[7] = 1u;  // this state was rendered by the warp-granular kernel (its backward twin checks)
    }
  }
}

int s3r_launch_blend_blocks(const s3r_raster_params& p, const s3r_raster_outputs& o, const s3r_raster_layout& L,
                            char* state, cudaStream_t st, int only_tile) {
  auto kern = o.n_touched ? s3r_blend_blocks_fwd_kernel<true> : s3r_blend_blocks_fwd_kernel<false>;
  static int slots[64] = {};
  int dev = 0;
  S3R_CUDA_CHECK(cudaGetDevice(&dev));
  dev &= 63;
  if (slots[dev] == 0) {
    int sms = 0, per_sm = 0;
    S3R_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    S3R_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s3r_blend_blocks_fwd_kernel<false>, BBK_THREADS, 0));
    if (per_sm > BBK_MINB) per_sm = BBK_MINB;
    slots[dev] = sms * (per_sm > 0 ? per_sm : 1);
  }
  uint32_t n_units = (uint32_t)L.tiles * (uint32_t)p.n_views * 8u;
  if (only_tile >= 0) n_units = 8u;
  const uint32_t ctas_needed = (n_units + BBK_WARPS - 1) / BBK_WARPS;
  dim3 grid(ctas_needed < (uint32_t)slots[dev] ? ctas_needed : (uint32_t)slots[dev]);
  S3R_CUDA_CHECK(s3r_launch_pdl(kern, grid, dim3(BBK_THREADS), 0, st, (s3r_raster_pdl_mask() >> 4) & 1, p.width, p.height,
                                p.P, L.tiles_x, L.tiles, only_tile, n_units, (const uint32_t*)(state + L.work_order),
                                (unsigned*)(state + L.counters), (const uint2*)(state + L.ranges),
                                (const float4*)(state + L.records), (const uint32_t*)(state + L.blists),
                                (const uint32_t*)(state + L.bcounts), (const uint32_t*)(state + L.point_list),
                                (const float4*)(state + L.conic_opacity), p.background, o.color, o.depth, o.opacity,
                                (float*)(state + L.final_T), (uint32_t*)(state + L.n_contrib),
                                (uint32_t*)(state + L.n_contrib_blk), o.n_touched));
  return S3R_OK;
}
