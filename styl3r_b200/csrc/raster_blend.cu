// Rasterizer stage 5: per-tile front-to-back alpha blend (forward).
//
// One CTA per (view, 16x16 tile); warp w owns the 8x4 pixel block (w&1, w>>1) of the tile, one pixel per
// lane.  The tile's sorted-gathered 48-byte records are streamed through a 2-stage shared-memory ring by
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx; SASS: UBLKCP / SYNCS), 256 records per stage.
// Per stage every thread tests one record's alpha>=1/255 bounding box against the eight pixel blocks and
// the warps exchange 8 ballots, so that each warp then walks only the records that can touch its block
// (warp-uniform compaction — skipping a record that cannot reach 1/255 is result-neutral).
//
// Semantics follow upstream renderCUDA + the "-w-pose" fork (blended depth, opacity, n_touched) as restated by
// oracle/raster_oracle.c:s3r_oracle_render:  power is evaluated with the
// oracle's exact unfused fp32 operation order; exp uses MUFU.EX2, and the rare evaluations whose alpha
// lands within 2e-5 (relative) of the 1/255 threshold are re-evaluated with a correctly rounded exp so
// that the keep/skip decision matches the oracle.
#include "s3r_common.cuh"

#define BLEND_THREADS 256
#define BLEND_CHUNK 256
#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_LO (ALPHA_MIN * (1.0f - 2e-5f))
#define ALPHA_HI (ALPHA_MIN * (1.0f + 2e-5f))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// exact (oracle-identical) alpha for threshold-band evaluations
__device__ __noinline__ float exact_alpha(float power, float opacity) {
  const float e = (float)exp((double)power);
  return fminf(0.99f, __fmul_rn(opacity, e));
}

struct __align__(16) BlendSmem {
  float4 rec[2][BLEND_CHUNK * 3];
  uint32_t mask[8][8];
  uint64_t full[2];
};

__global__ void __launch_bounds__(BLEND_THREADS) s3r_blend_fwd_kernel(
    int W, int H, int P, int tiles_x, int tiles, const uint2* __restrict__ ranges, const float4* __restrict__ records,
    const uint32_t* __restrict__ point_list, const float* __restrict__ background, float* __restrict__ out_color,
    float* __restrict__ out_depth, float* __restrict__ out_opacity, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, int32_t* __restrict__ n_touched) {
  __shared__ BlendSmem sm;
  const int view = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, w = tid >> 5;
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int X0 = tx * S3R_TILE, Y0 = ty * S3R_TILE;
  const int px = X0 + (w & 1) * 8 + (lane & 7), py = Y0 + (w >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const uint2 rg = ranges[(size_t)view * tiles + tile];
  const uint32_t n = rg.y - rg.x;
  const uint32_t nchunks = (n + BLEND_CHUNK - 1) / BLEND_CHUNK;
  const float4* src = records + (size_t)rg.x * 3;

  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t c = 0; c < 2 && c < nchunks; c++) {
      const uint32_t cnt = min((uint32_t)BLEND_CHUNK, n - c * BLEND_CHUNK);
      mbar_expect_tx(&sm.full[c], cnt * S3R_REC_BYTES);
      bulk_g2s(sm.rec[c], src + (size_t)c * BLEND_CHUNK * 3, cnt * S3R_REC_BYTES, &sm.full[c]);
    }
  }

  float T = 1.0f, Cr = 0.f, Cg = 0.f, Cb = 0.f, D = 0.f;
  uint32_t last = 0;
  bool done = !inside;
  // block pixel bounds for the cull test
  const float bx0 = (float)X0, by0 = (float)Y0;

  for (uint32_t c = 0; c < nchunks; c++) {
    const int s = c & 1;
    const uint32_t cnt = min((uint32_t)BLEND_CHUNK, n - c * BLEND_CHUNK);
    mbar_wait(&sm.full[s], (c >> 1) & 1);
    // ---- cull: record `tid` against the 8 pixel blocks
    {
      bool hx0 = false, hx1 = false, hy[4] = {false, false, false, false};
      if ((uint32_t)tid < cnt) {
        const float4 r0 = sm.rec[s][tid * 3];
        const float4 r2 = sm.rec[s][tid * 3 + 2];
        const float xl = r0.x - r2.z, xh = r0.x + r2.z, yl = r0.y - r2.w, yh = r0.y + r2.w;
        const bool ok = r2.z >= 0.f;
        hx0 = ok && xh >= bx0 && xl <= bx0 + 7.f;
        hx1 = ok && xh >= bx0 + 8.f && xl <= bx0 + 15.f;
#pragma unroll
        for (int k = 0; k < 4; k++) hy[k] = yh >= by0 + 4.f * k && yl <= by0 + 4.f * k + 3.f;
      }
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const uint32_t m = __ballot_sync(0xffffffffu, ((b & 1) ? hx1 : hx0) && hy[b >> 1]);
        if (lane == 0) sm.mask[b][w] = m;
      }
    }
    __syncthreads();
    // ---- blend: warp w walks the survivors of its block in list order
    if (__any_sync(0xffffffffu, !done)) {
      const uint32_t base_idx = c * BLEND_CHUNK;
#pragma unroll 1
      for (int j = 0; j < 8; j++) {
        uint32_t m = sm.mask[w][j];
        while (m) {
          const int i = j * 32 + __ffs(m) - 1;
          m &= m - 1;
          const float4 r0 = sm.rec[s][i * 3];
          const float4 r1 = sm.rec[s][i * 3 + 1];
          const float4 r2 = sm.rec[s][i * 3 + 2];
          if (!done) {
            const float dx = r0.x - pxf, dy = r0.y - pyf;
            const float q = __fadd_rn(__fmul_rn(__fmul_rn(r0.z, dx), dx), __fmul_rn(__fmul_rn(r1.x, dy), dy));
            const float power = __fsub_rn(__fmul_rn(-0.5f, q), __fmul_rn(__fmul_rn(r0.w, dx), dy));
            if (!(power > 0.0f)) {
              float alpha = fminf(0.99f, r1.y * __expf(power));
              bool keep = true;
              if (alpha < ALPHA_HI) {
                keep = false;
                if (alpha >= ALPHA_LO) {
                  alpha = exact_alpha(power, r1.y);
                  keep = alpha >= ALPHA_MIN;
                }
              }
              if (keep) {
                const float test_T = T * (1.0f - alpha);
                if (test_T < 0.0001f) {
                  done = true;
                } else {
                  const float wgt = alpha * T;
                  Cr += r1.z * wgt;
                  Cg += r1.w * wgt;
                  Cb += r2.x * wgt;
                  D += r2.y * wgt;
                  if (n_touched != nullptr && test_T > 0.5f)
                    atomicAdd(&n_touched[(size_t)view * P + point_list[(size_t)rg.x + base_idx + i]], 1);
                  T = test_T;
                  last = base_idx + i + 1;
                }
              }
            }
          }
        }
        if (!__any_sync(0xffffffffu, !done)) break;
      }
    }
    const int ndone = __syncthreads_count(done);
    if (ndone == BLEND_THREADS) {
      // a bulk copy for chunk c+1 may still be in flight into the other stage: drain it before the CTA retires
      if (tid == 0 && c + 1 < nchunks) mbar_wait(&sm.full[s ^ 1], ((c + 1) >> 1) & 1);
      break;
    }
    if (tid == 0 && c + 2 < nchunks) {
      const uint32_t c2 = c + 2;
      const uint32_t cnt2 = min((uint32_t)BLEND_CHUNK, n - c2 * BLEND_CHUNK);
      mbar_expect_tx(&sm.full[s], cnt2 * S3R_REC_BYTES);
      bulk_g2s(sm.rec[s], src + (size_t)c2 * BLEND_CHUNK * 3, cnt2 * S3R_REC_BYTES, &sm.full[s]);
    }
  }
  if (inside) {
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    const float* bg = background + view * 3;
    float* oc = out_color + (size_t)view * 3 * HW;
    oc[pix] = Cr + T * bg[0];
    oc[HW + pix] = Cg + T * bg[1];
    oc[2 * HW + pix] = Cb + T * bg[2];
    out_depth[(size_t)view * HW + pix] = D;
    out_opacity[(size_t)view * HW + pix] = 1.0f - T;
    final_T[(size_t)view * HW + pix] = T;
    n_contrib[(size_t)view * HW + pix] = last;
  }
}

int s3r_launch_blend(const s3r_raster_params& p, const s3r_raster_outputs& o, const s3r_raster_layout& L,
                     char* state, cudaStream_t st) {
  dim3 grid(L.tiles, p.n_views);
  s3r_blend_fwd_kernel<<<grid, BLEND_THREADS, 0, st>>>(
      p.width, p.height, p.P, L.tiles_x, L.tiles, (const uint2*)(state + L.ranges),
      (const float4*)(state + L.records), (const uint32_t*)(state + L.point_list), p.background, o.color, o.depth,
      o.opacity, (float*)(state + L.final_T), (uint32_t*)(state + L.n_contrib), o.n_touched);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
