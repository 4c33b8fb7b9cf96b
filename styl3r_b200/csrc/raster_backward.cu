// Rasterizer backward: blend backward (per tile, back to front) + preprocess backward (per Gaussian),
// including the camera-pose gradient dL/dtau of the "-w-pose" fork that Styl3R consumes through
// cam_trans_delta -> rho / cam_rot_delta -> theta (src/model/decoder/cuda_splatting.py:127-128,
// src/misc/cam_utils.py:103-137: w2c' = SE3_exp([rho, theta]) @ w2c).
//
// Semantics: oracle/raster_oracle.c:s3r_oracle_render_backward / s3r_oracle_preprocess_backward (SURVEY.md
// Appendix B "Backward"), which are validated against torch autograd (oracle/torch_mirror.py).
//
// blend backward: same warp-specialised layout as the forward (TMA producer warp streaming the tile's
// 48-byte records through an mbarrier ring, 8 decoupled consumer warps each owning an 8x4 pixel block), but
// the ring runs from the last contributing chunk to the first.  Per-record gradients are summed over the
// 32 pixels of a warp with a multi-value butterfly (30 values / 3 records: 31 shuffles instead of 150) and
// one atomicAdd per value goes to a per-(view, Gaussian) accumulator [V, P, 12]:
//   0-1 dL/dmean2D (NDC)  2-4 dL/dconic (xx, xy(half), yy)  5 dL/dopacity  6-8 dL/drgb  9 dL/ddepth
#include "s3r_common.cuh"

#define BB_THREADS 288
#define BB_CHUNK 128
#define BB_STAGES 4
#define BB_CWARPS 8
#define BB_U 3  // records per reduction batch: 3 x 10 gradient values fill a 32-lane butterfly
#define ACC_STRIDE 12
#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_LO (ALPHA_MIN * (1.0f - S3R_ALPHA_BAND))
#define ALPHA_HI (ALPHA_MIN * (1.0f + S3R_ALPHA_BAND))

__device__ __forceinline__ uint32_t bsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bmbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bsmem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bmbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bsmem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bmbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bsmem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool bmbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bsmem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bmbar_wait(uint64_t* bar, uint32_t parity) {
  while (!bmbar_try(bar, parity)) {
  }
}
__device__ __forceinline__ void bbulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   bsmem_u32(dst)),
               "l"(src), "r"(bytes), "r"(bsmem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float bfast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __noinline__ float bexact_alpha(float power, float opacity) {
  const float e = (float)exp((double)power);
  return fminf(0.99f, __fmul_rn(opacity, e));
}

// Sum v[j] over the 32 lanes for j < 32; afterwards lane L holds the total of v[L] (returned).
__device__ __forceinline__ float warp_multi_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; j++) {
      const float send = up ? v[j] : v[j + off];
      const float keep = up ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ uint64_t bpack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void bunpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t bmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t bfma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float bfast_rcp(float x) {  // MUFU.RCP, 1 ulp; x = 1 - alpha >= 0.01
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct __align__(128) BwdSmem {
  float4 rec[BB_STAGES][BB_CHUNK * 3];
  uint8_t list[BB_CWARPS][BB_CHUNK + 16];
  uint64_t full[BB_STAGES];
  uint64_t empty[BB_STAGES];
  unsigned max_last;
};

// Which per-splat gradients a launch accumulates (chosen on the host from the NULL-ness of the requested outputs):
//   ALL   10 values per splat (mean2D 2, conic 3, opacity 1, colour 3, depth 1), 3 splats per reduction batch
//   GEOM   6 values (mean2D, conic, depth): the pose-align loop (dL/dtau only, colours independent of the pose), 5 per batch
//   COLOR  3 values: stage-2 training (only dL/dSH is needed, model_wrapper_style.py:854-868), 10 per batch; the
//          dL/dalpha recurrence is not evaluated at all
// Every batch fills a 32-lane butterfly (30 values), so the reduction costs 124 / 3, 124 / 5 or 124 / 10 instructions
// per splat.
enum { BWD_ALL = 0, BWD_GEOM = 1, BWD_COLOR = 2 };
template <int MODE> struct BwdMode;
template <> struct BwdMode<BWD_ALL> { static constexpr int NV = 10, U = 3; };
template <> struct BwdMode<BWD_GEOM> { static constexpr int NV = 6, U = 5; };
template <> struct BwdMode<BWD_COLOR> { static constexpr int NV = 3, U = 10; };
__device__ __forceinline__ int bwd_slot(int mode, int j) {  // accumulator column of value j
  return mode == BWD_ALL ? j : (mode == BWD_GEOM ? (j < 5 ? j : 9) : 6 + j);
}

template <int MODE>
__global__ void __launch_bounds__(BB_THREADS, 2) s3r_blend_bwd_kernel(
    int W, int H, int P, int tiles_x, int tiles, const uint32_t* __restrict__ work_order, const uint2* __restrict__ ranges,
    const float4* __restrict__ records, const uint32_t* __restrict__ point_list, const float4* __restrict__ conic_opacity,
    const float* __restrict__ background, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
    float* __restrict__ acc) {
  constexpr int NV = BwdMode<MODE>::NV, U = BwdMode<MODE>::U;
  __shared__ BwdSmem sm;
  // (view, tile) in the order of the forward pass's blend queue (raster_sort.cu: heaviest tiles first, the first waves laid
  // out so that an SM's second CTA gets a light tile)
  const uint32_t vt = work_order[blockIdx.x];
  const int view = vt / tiles, tile = vt % tiles, tid = threadIdx.x;
  const int lane = tid & 31, w = tid >> 5;
  const uint2 rg = ranges[(size_t)view * tiles + tile];
  const float4* src = records + (size_t)rg.x * 3;
  const size_t HW = (size_t)H * W;

  // consumer-side pixel state (producer warp computes dummies)
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int wc = w < BB_CWARPS ? w : 0;
  const int X0 = tx * S3R_TILE + (wc & 1) * 8, Y0 = ty * S3R_TILE + (wc >> 1) * 4;
  const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
  const bool inside = (w < BB_CWARPS) && px < W && py < H;
  const size_t pix = (size_t)py * W + px;
  const uint32_t last = inside ? n_contrib[(size_t)view * HW + pix] : 0u;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < BB_STAGES; s++) {
      bmbar_init(&sm.full[s], 1);
      bmbar_init(&sm.empty[s], BB_CWARPS);
    }
    sm.max_last = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t wmax = last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0 && wmax) atomicMax(&sm.max_last, wmax);
  __syncthreads();
  const uint32_t max_last = sm.max_last;               // records [0, max_last) can matter for this tile
  const uint32_t nch = (max_last + BB_CHUNK - 1) / BB_CHUNK;
  if (nch == 0) return;

  if (w == BB_CWARPS) {
    if (lane == 0) {
      for (uint32_t q = 0; q < nch; q++) {
        const uint32_t c = nch - 1 - q;
        const int s = q % BB_STAGES;
        if (q >= BB_STAGES) bmbar_wait(&sm.empty[s], ((q / BB_STAGES) - 1) & 1);
        const uint32_t cnt = min((uint32_t)BB_CHUNK, max_last - c * BB_CHUNK);
        bmbar_expect_tx(&sm.full[s], cnt * S3R_REC_BYTES);
        bbulk_g2s(sm.rec[s], src + (size_t)c * BB_CHUNK * 3, cnt * S3R_REC_BYTES, &sm.full[s]);
      }
      const uint32_t first = nch > BB_STAGES ? nch - BB_STAGES : 0;
      for (uint32_t q = first; q < nch; q++) bmbar_wait(&sm.full[q % BB_STAGES], (q / BB_STAGES) & 1);
    }
    return;
  }

  // ===== consumers
  const float pxf = (float)px, pyf = (float)py;
  const uint32_t cellbits = 3u << (4 * (w >> 1) + 2 * (w & 1));
  const uint32_t lt = (1u << lane) - 1u;
  uint8_t* list = sm.list[w];
  const float T_final = inside ? final_T[(size_t)view * HW + pix] : 0.f;
  float T = T_final;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLd = 0.f;
  if (inside) {
    const float* g = dL_dcolor + (size_t)view * 3 * HW;
    dLp0 = g[pix];
    dLp1 = g[HW + pix];
    dLp2 = g[2 * HW + pix];
    if (dL_ddepth) dLd = dL_ddepth[(size_t)view * HW + pix];
  }
  const float* bg = background + view * 3;
  const float nbg = -T_final * (bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2);  // d(T_final * bg) / dalpha = nbg / (1 - alpha)
  const float fW = (float)W, fH = (float)H;  // 2 * ddelx_dx, 2 * ddely_dy
  const uint64_t dLp01 = bpack2(dLp0, dLp1), dLp2d = bpack2(dLp2, dLd);
  uint64_t acc01 = bpack2(0.f, 0.f), acc2d = bpack2(0.f, 0.f);  // accum_rec / accum_depth_rec (colours behind the splat)
  uint64_t lc01 = bpack2(0.f, 0.f), lc2d = bpack2(0.f, 0.f);    // colour / depth of the last kept splat
  float last_alpha = 0.f;
  float* accv = acc + (size_t)view * P * ACC_STRIDE;

  for (uint32_t q = 0; q < nch; q++) {
    const uint32_t c = nch - 1 - q;
    const int s = q % BB_STAGES;
    bmbar_wait(&sm.full[s], (q / BB_STAGES) & 1);
    const uint32_t base_idx = c * BB_CHUNK;
    if (base_idx < wmax) {
      const uint32_t cnt = min((uint32_t)BB_CHUNK, max_last - base_idx);
      int count = 0;
#pragma unroll
      for (int j = 0; j < BB_CHUNK / 32; j++) {
        const int i = j * 32 + lane;
        // the instance's cell mask (sort epilogue): bits of the two 4x4 cells that make up this warp's 8x4 block
        const bool hit = (uint32_t)i < cnt && (__float_as_uint(sm.rec[s][i * 3 + 2].z) & cellbits) != 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (hit) list[count + __popc(m & lt)] = (uint8_t)i;
        count += __popc(m);
      }
      __syncwarp();
      // walk the survivors back to front, U at a time
#pragma unroll 1
      for (int k = count; k > 0; k -= U) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 0.f;
        bool any_keep = false;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int kk = k - 1 - u;  // list position, descending
          const int i = kk >= 0 ? (int)list[kk] : 0;
          const float4 r0 = sm.rec[s][i * 3];
          const float4 r1 = sm.rec[s][i * 3 + 1];
          const float dx = r0.x - pxf, dy = r0.y - pyf;
          // same fast log2-domain evaluation and guard bands as the forward kernel (raster_blend.cu), so that the
          // set of contributing Gaussians is the one the forward pass composited
          const float l2g = fmaf(dx, fmaf(r1.x, dx, r0.z * dy), (r0.w * dy) * dy);
          float G = bfast_exp2(l2g);
          float alpha = fminf(0.99f, r1.y * G);
          bool kp = kk >= 0 && (base_idx + (uint32_t)i) < last;
          // the true conic for the gradient formulas (the record holds B' = KB*B, C' = KA*C | A' = KA*A)
          float cA = r1.x * S3R_INV_KA, cB = r0.z * S3R_INV_KB, cC = r0.w * S3R_INV_KA;
          if (kp && ((alpha >= ALPHA_LO && alpha < ALPHA_HI) || fabsf(l2g) < S3R_PZERO_BAND)) {
            // rare: decided like the forward pass, from the exact conic
            const uint32_t gid = __ldg(point_list + (size_t)rg.x + base_idx + i);
            const float4 co = __ldg(conic_opacity + (size_t)view * P + gid);
            const float qf = __fadd_rn(__fmul_rn(__fmul_rn(co.x, dx), dx), __fmul_rn(__fmul_rn(co.z, dy), dy));
            const float power = __fsub_rn(__fmul_rn(-0.5f, qf), __fmul_rn(__fmul_rn(co.y, dx), dy));
            kp = !(power > 0.0f);
            if (kp) {
              alpha = bexact_alpha(power, co.w);
              kp = alpha >= ALPHA_MIN;
              G = __expf(power);
              cA = co.x, cB = co.y, cC = co.z;
            }
          } else {
            kp = kp && !(l2g > 0.0f) && alpha >= ALPHA_HI;
          }
          // branch-free from here: a splat that is not kept runs with alpha = 0, G = 0 - T, the colour recurrence and
          // every gradient value come out exactly as if it had been skipped
          any_keep = any_keep || kp;
          alpha = kp ? alpha : 0.f;
          G = kp ? G : 0.f;
          const float inv = bfast_rcp(1.f - alpha);
          T *= inv;
          const float dch = alpha * T;  // d channel / d colour
          if (MODE != BWD_GEOM) {
            const int o = MODE == BWD_ALL ? u * NV + 6 : u * NV;
            float c0, c1;
            bunpack2(bmul2(dLp01, bpack2(dch, dch)), c0, c1);
            v[o] = c0;
            v[o + 1] = c1;
            v[o + 2] = dch * dLp2;
            if (MODE == BWD_ALL) v[o + 3] = dch * dLd;
          }
          if (MODE != BWD_COLOR) {
            const float2 r2 = *reinterpret_cast<const float2*>(&sm.rec[s][i * 3 + 2]);
            // colour behind this splat: acc = last_alpha * last_colour + (1 - last_alpha) * acc
            const float oml = 1.f - last_alpha;
            acc01 = bfma2(bpack2(last_alpha, last_alpha), lc01, bmul2(bpack2(oml, oml), acc01));
            acc2d = bfma2(bpack2(last_alpha, last_alpha), lc2d, bmul2(bpack2(oml, oml), acc2d));
            lc01 = bpack2(r1.z, r1.w);
            lc2d = bpack2(r2.x, r2.y);
            last_alpha = alpha;
            // dL/dalpha = T * sum_c (colour_c - acc_c) * dL/dpixel_c  -  T_final / (1 - alpha) * (bg . dL/dpixel)
            float s0, s1;
            bunpack2(bfma2(bfma2(acc2d, bpack2(-1.f, -1.f), lc2d), dLp2d,
                           bmul2(bfma2(acc01, bpack2(-1.f, -1.f), lc01), dLp01)), s0, s1);
            const float dL_dalpha = fmaf(s0 + s1, T, nbg * inv);
            const float h = -0.5f * (r1.y * dL_dalpha);  // -0.5 * dL/dG
            const float hgx = h * (G * dx), hgy = h * (G * dy);
            const int o = u * NV;
            v[o + 0] = fW * fmaf(hgx, cA, hgy * cB);  // dL/dG * dG/ddelx * ddelx/dx (ddelx/dx = W / 2)
            v[o + 1] = fH * fmaf(hgy, cC, hgx * cB);
            v[o + 2] = hgx * dx;
            v[o + 3] = hgx * dy;
            v[o + 4] = hgy * dy;
            v[o + 5] = MODE == BWD_ALL ? G * dL_dalpha : dch * dLd;  // opacity (ALL) / depth (GEOM) column
          }
        }
        if (__any_sync(0xffffffffu, any_keep)) {
          // lane L < U * NV will own value L = (splat L / NV, component L % NV): fetch that splat's Gaussian id early
          const int u_l = lane / NV, comp = lane - u_l * NV;
          const bool owner = lane < U * NV && (k - 1 - u_l) >= 0;
          uint32_t gid = 0;
          if (owner) gid = __ldg(point_list + (size_t)rg.x + base_idx + list[k - 1 - u_l]);
          const float total = warp_multi_reduce32(v, lane);
          if (owner && total != 0.f) atomicAdd(accv + (size_t)gid * ACC_STRIDE + bwd_slot(MODE, comp), total);
        }
      }
    }
    __syncwarp();
    if (lane == 0) bmbar_arrive(&sm.empty[s]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// blend backward, warp-granular: the backward twin of raster_blend_blocks.cu.  One warp per (view, tile, 8x4 block)
// walks the block's survivor list (tile sort epilogue) back to front, from the list position the forward pass
// recorded per pixel (`n_contrib_blk` = entries up to and including the last contributor).  Records are gathered
// with cp.async, 32 per group, double-buffered in the warp's private shared memory: no cull, no shared ring, warps
// of a tile independent.  Per-splat arithmetic, guard bands and reduction as in s3r_blend_bwd_kernel above.
// ------------------------------------------------------------------------------------------------------------
#define BWB_WARPS 8
#define BWB_THREADS (32 * BWB_WARPS)
#define BWB_GROUP 32
#define BWB_RED_STRIDE 36
#define BWB_SMEM_BYTES (sizeof(float4) * BWB_WARPS * 2 * BWB_GROUP * 3 + sizeof(float) * BWB_WARPS * 30 * BWB_RED_STRIDE)

__device__ __forceinline__ void bcp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bcp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bcp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(BWB_THREADS, 2) s3r_blend_bwd_blocks_kernel(
    int W, int H, int P, int tiles_x, int tiles, uint32_t n_units, const unsigned* __restrict__ counters,
    const uint32_t* __restrict__ work_order,
    const uint2* __restrict__ ranges, const float4* __restrict__ records, const uint32_t* __restrict__ blists,
    const uint32_t* __restrict__ point_list, const float4* __restrict__ conic_opacity,
    const float* __restrict__ background, const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib_blk,
    const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth, float* __restrict__ acc) {
  constexpr int NV = BwdMode<MODE>::NV, U = BwdMode<MODE>::U;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  float4 (*s_rec)[2][BWB_GROUP * 3] = reinterpret_cast<float4 (*)[2][BWB_GROUP * 3]>(s_dyn);  // [BWB_WARPS][2][96]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // per-warp transpose buffer of the gradient reduction: value j of lane l at [j * BWB_RED_STRIDE + l]
  float* red = reinterpret_cast<float*>(s_dyn + sizeof(float4) * BWB_WARPS * 2 * BWB_GROUP * 3) + w * (30 * BWB_RED_STRIDE);
  // strided like the forward kernel's first units: a CTA's warps work on blocks of eight different tiles across the
  // weight-ordered queue (n_units is a multiple of 8, the grid is n_units / 8)
  // the walk starts from n_contrib_blk, which only the warp-granular forward kernel writes: fail loudly on a state that
  // the other forward kernel rendered (S3R_TUNE_BLEND_KERNEL toggled between a forward and its backward)
  if (counters[7] != 1u) __trap();
  const uint32_t unit = w * gridDim.x + blockIdx.x;
  if (unit >= n_units) return;
  const uint32_t vt = work_order[unit >> 3];
  const int blk = unit & 7;
  const int view = vt / tiles, tile = vt % tiles;
  const uint2 rg = ranges[vt];
  const uint32_t n = rg.y - rg.x;
  const uint32_t* list = blists + (size_t)rg.x * 8 + (size_t)blk * n;
  const float4* src = records + (size_t)rg.x * 3;
  const size_t HW = (size_t)H * W;
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int X0 = tx * S3R_TILE + (blk & 1) * 8, Y0 = ty * S3R_TILE + (blk >> 1) * 4;
  const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const size_t pix = (size_t)py * W + px;
  const uint32_t klast = inside ? n_contrib_blk[(size_t)view * HW + pix] : 0u;  // list entries [0, klast) matter
  uint32_t kmax = klast;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  if (kmax == 0) return;
  const uint32_t ngroups = (kmax + BWB_GROUP - 1) / BWB_GROUP;

  const float pxf = (float)px, pyf = (float)py;
  const float T_final = inside ? final_T[(size_t)view * HW + pix] : 0.f;
  float T = T_final;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLd = 0.f;
  if (inside) {
    const float* g = dL_dcolor + (size_t)view * 3 * HW;
    dLp0 = g[pix];
    dLp1 = g[HW + pix];
    dLp2 = g[2 * HW + pix];
    if (dL_ddepth) dLd = dL_ddepth[(size_t)view * HW + pix];
  }
  const float* bg = background + view * 3;
  const float nbg = -T_final * (bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2);
  const float fW = (float)W, fH = (float)H;
  const uint64_t dLp01 = bpack2(dLp0, dLp1), dLp2d = bpack2(dLp2, dLd);
  uint64_t acc01 = bpack2(0.f, 0.f), acc2d = bpack2(0.f, 0.f);
  uint64_t lc01 = bpack2(0.f, 0.f), lc2d = bpack2(0.f, 0.f);
  float last_alpha = 0.f;
  float* accv = acc + (size_t)view * P * ACC_STRIDE;
  const uint32_t buf0 = bsmem_u32(&s_rec[w][0][0]);
  constexpr uint32_t kBufBytes = BWB_GROUP * S3R_REC_BYTES;

  // groups are walked from the last one down; `it` counts iterations (buffer parity), g = ngroups - 1 - it
  auto load_idx = [&](uint32_t it) -> uint32_t {
    if (it >= ngroups) return 0u;
    const uint32_t j = (ngroups - 1 - it) * BWB_GROUP + lane;
    return j < kmax ? list[j] : 0u;
  };
  auto gather = [&](uint32_t it, uint32_t idx) {
    const uint32_t j = (ngroups - 1 - it) * BWB_GROUP + lane;
    const uint32_t dst = buf0 + (it & 1u) * kBufBytes + lane * S3R_REC_BYTES;
    if (j < kmax) {
      const float4* r = src + (size_t)idx * 3;
      bcp_async16(dst, r);
      bcp_async16(dst + 16, r + 1);
      bcp_async16(dst + 32, r + 2);
    } else {  // past the range: a finite dummy (never kept)
      float4* d = &s_rec[w][it & 1u][lane * 3];
      d[0] = make_float4(3e9f, 3e9f, 0.0f, -1.0f);
      d[1] = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
      d[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    bcp_commit();
  };
  uint32_t idx_cur = load_idx(0);
  gather(0, idx_cur);
  uint32_t idx_gath = load_idx(1);  // indices of the group whose gather is issued next

  for (uint32_t it = 0; it < ngroups; it++) {
    const uint32_t g = ngroups - 1 - it;
    uint32_t idx_pref = 0;
    if (it + 1 < ngroups) {
      gather(it + 1, idx_gath);
      idx_pref = load_idx(it + 2);
      bcp_wait<1>();
    } else {
      bcp_wait<0>();
    }
    __syncwarp();
    const float4* rec = &s_rec[w][it & 1u][0];
    const int gcnt = (int)min((uint32_t)BWB_GROUP, kmax - g * BWB_GROUP);
#pragma unroll 1
    for (int k = gcnt; k > 0; k -= U) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = 0.f;
      bool any_keep = false;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int kk = k - 1 - u;  // slot in the group, descending
        const int i = kk >= 0 ? kk : 0;
        const float4 r0 = rec[i * 3];
        const float4 r1 = rec[i * 3 + 1];
        const float dx = r0.x - pxf, dy = r0.y - pyf;
        const float l2g = fmaf(dx, fmaf(r1.x, dx, r0.z * dy), (r0.w * dy) * dy);
        float G = bfast_exp2(l2g);
        float alpha = fminf(0.99f, r1.y * G);
        bool kp = kk >= 0 && (g * BWB_GROUP + (uint32_t)kk) < klast;
        float cA = r1.x * S3R_INV_KA, cB = r0.z * S3R_INV_KB, cC = r0.w * S3R_INV_KA;
        if (kp && ((alpha >= ALPHA_LO && alpha < ALPHA_HI) || fabsf(l2g) < S3R_PZERO_BAND)) {
          const uint32_t gid = __ldg(point_list + (size_t)rg.x + __ldg(list + g * BWB_GROUP + kk));
          const float4 co = __ldg(conic_opacity + (size_t)view * P + gid);
          const float qf = __fadd_rn(__fmul_rn(__fmul_rn(co.x, dx), dx), __fmul_rn(__fmul_rn(co.z, dy), dy));
          const float power = __fsub_rn(__fmul_rn(-0.5f, qf), __fmul_rn(__fmul_rn(co.y, dx), dy));
          kp = !(power > 0.0f);
          if (kp) {
            alpha = bexact_alpha(power, co.w);
            kp = alpha >= ALPHA_MIN;
            G = __expf(power);
            cA = co.x, cB = co.y, cC = co.z;
          }
        } else {
          kp = kp && !(l2g > 0.0f) && alpha >= ALPHA_HI;
        }
        any_keep = any_keep || kp;
        alpha = kp ? alpha : 0.f;
        G = kp ? G : 0.f;
        const float inv = bfast_rcp(1.f - alpha);
        T *= inv;
        const float dch = alpha * T;
        if (MODE != BWD_GEOM) {
          const int o = MODE == BWD_ALL ? u * NV + 6 : u * NV;
          float c0, c1;
          bunpack2(bmul2(dLp01, bpack2(dch, dch)), c0, c1);
          v[o] = c0;
          v[o + 1] = c1;
          v[o + 2] = dch * dLp2;
          if (MODE == BWD_ALL) v[o + 3] = dch * dLd;
        }
        if (MODE != BWD_COLOR) {
          const float2 r2 = *reinterpret_cast<const float2*>(&rec[i * 3 + 2]);
          const float oml = 1.f - last_alpha;
          acc01 = bfma2(bpack2(last_alpha, last_alpha), lc01, bmul2(bpack2(oml, oml), acc01));
          acc2d = bfma2(bpack2(last_alpha, last_alpha), lc2d, bmul2(bpack2(oml, oml), acc2d));
          lc01 = bpack2(r1.z, r1.w);
          lc2d = bpack2(r2.x, r2.y);
          last_alpha = alpha;
          float s0, s1;
          bunpack2(bfma2(bfma2(acc2d, bpack2(-1.f, -1.f), lc2d), dLp2d,
                         bmul2(bfma2(acc01, bpack2(-1.f, -1.f), lc01), dLp01)), s0, s1);
          const float dL_dalpha = fmaf(s0 + s1, T, nbg * inv);
          const float h = -0.5f * (r1.y * dL_dalpha);
          const float hgx = h * (G * dx), hgy = h * (G * dy);
          const int o = u * NV;
          v[o + 0] = fW * fmaf(hgx, cA, hgy * cB);
          v[o + 1] = fH * fmaf(hgy, cC, hgx * cB);
          v[o + 2] = hgx * dx;
          v[o + 3] = hgx * dy;
          v[o + 4] = hgy * dy;
          v[o + 5] = MODE == BWD_ALL ? G * dL_dalpha : dch * dLd;
        }
      }
      if (__any_sync(0xffffffffu, any_keep)) {
        // lane L < U * NV owns value L = (splat L / NV, component L % NV); its Gaussian id comes from the group's list
        // indices, which lane `slot` still holds (idx_cur)
        const int u_l = lane / NV, comp = lane - u_l * NV;
        const int slot = k - 1 - u_l;
        const bool owner = lane < U * NV && slot >= 0;
        const uint32_t pos = __shfl_sync(0xffffffffu, idx_cur, slot >= 0 ? slot : 0);
        uint32_t gid = 0;
        if (owner) gid = __ldg(point_list + (size_t)rg.x + pos);
        // sum of every value over the 32 pixels through a shared-memory transpose: 30 conflict-free stores, then lane L
        // adds up row L with 8 LDS.128 (row stride 36 floats: a quarter warp covers all banks once) - 70 instructions
        // instead of the 124 of the select / shuffle butterfly
#pragma unroll
        for (int j = 0; j < U * NV; j++) red[j * BWB_RED_STRIDE + lane] = v[j];
        __syncwarp();
        float total = 0.f;
        if (lane < U * NV) {
          const float4* row = reinterpret_cast<const float4*>(red + lane * BWB_RED_STRIDE);
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float4 x = row[q];
            total += (x.x + x.y) + (x.z + x.w);
          }
        }
        __syncwarp();  // the next batch overwrites the buffer
        if (owner && total != 0.f) atomicAdd(accv + (size_t)gid * ACC_STRIDE + bwd_slot(MODE, comp), total);
      }
    }
    __syncwarp();  // the buffer is refilled two iterations from now
    idx_cur = idx_gath;
    idx_gath = idx_pref;
  }
}

// ------------------------------------------------------------------------------------------------------------
// preprocess backward: one thread per (view, Gaussian).  fp32 restatement of the oracle's double-precision
// chain (conic -> cov2D -> (Sigma, t, R);  NDC mean -> t;  depth -> t.z;  SH;  t, R -> tau).
// ------------------------------------------------------------------------------------------------------------
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
__device__ __constant__ float bSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                           -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float bSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                           0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                           -0.5900435899266435f};

__global__ void __launch_bounds__(256) s3r_preprocess_bwd_kernel(s3r_raster_params prm, s3r_raster_grads gr,
                                                                 const uint32_t* __restrict__ rect,
                                                                 const float4* __restrict__ rgbm,
                                                                 const float* __restrict__ acc) {
  __shared__ float s_tau[8][6];
  const int view = blockIdx.y, tid = threadIdx.x;
  const int g = blockIdx.x * blockDim.x + tid;
  const int P = prm.P, W = prm.width, H = prm.height;
  float tau[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const size_t vi = (size_t)view * P + g;
  const bool live = g < P && rect[g < P ? vi : 0] != 0u;
  if (live) {
    const int set = prm.view_set ? prm.view_set[view] : view;
    const size_t gi = (size_t)set * P + g;
    const float s = prm.scales ? prm.scales[view] : 1.0f, s2 = s * s;
    const float* vm = prm.viewmatrix + view * 16;
    const float* pr = prm.projmatrix_raw + view * 16;
    const float tanx = prm.tanfov[view * 2], tany = prm.tanfov[view * 2 + 1];
    const float fx = W / (2.0f * tanx), fy = H / (2.0f * tany);
    const float* a = acc + vi * ACC_STRIDE;
    const float* mp = prm.means3D + gi * 3;
    const float p[3] = {mp[0] * s, mp[1] * s, mp[2] * s};
    float R[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) R[r][c] = vm[4 * c + r];
    float t[3];
#pragma unroll
    for (int r = 0; r < 3; r++) t[r] = R[r][0] * p[0] + R[r][1] * p[1] + R[r][2] * p[2] + vm[12 + r];
    float dL_dt[3] = {0.f, 0.f, 0.f};
    float dL_dR[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float dL_dp_direct[3] = {0.f, 0.f, 0.f};
    // ---- (1) conic -> cov2D -> Sigma, t, R
    {
      const float* cp = prm.cov3D + gi * prm.cov_stride;
      float S[3][3];
      if (prm.cov_stride == 9) {
        S[0][0] = cp[0] * s2; S[0][1] = cp[1] * s2; S[0][2] = cp[2] * s2;
        S[1][1] = cp[4] * s2; S[1][2] = cp[5] * s2; S[2][2] = cp[8] * s2;
      } else {
        S[0][0] = cp[0] * s2; S[0][1] = cp[1] * s2; S[0][2] = cp[2] * s2;
        S[1][1] = cp[3] * s2; S[1][2] = cp[4] * s2; S[2][2] = cp[5] * s2;
      }
      S[1][0] = S[0][1]; S[2][0] = S[0][2]; S[2][1] = S[1][2];
      const float limx = 1.3f * tanx, limy = 1.3f * tany;
      const float txtz = t[0] / t[2], tytz = t[1] / t[2];
      const float cx = fminf(limx, fmaxf(-limx, txtz)), cy = fminf(limy, fmaxf(-limy, tytz));
      const float xg = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
      const float yg = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
      const float tz = t[2], tcx = cx * tz, tcy = cy * tz;
      const float Jc[2][3] = {{fx / tz, 0.f, -fx * tcx / (tz * tz)}, {0.f, fy / tz, -fy * tcy / (tz * tz)}};
      float M[2][3], MS[2][3];
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) M[r][c] = Jc[r][0] * R[0][c] + Jc[r][1] * R[1][c] + Jc[r][2] * R[2][c];
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) MS[r][c] = M[r][0] * S[0][c] + M[r][1] * S[1][c] + M[r][2] * S[2][c];
      const float ca = MS[0][0] * M[0][0] + MS[0][1] * M[0][1] + MS[0][2] * M[0][2] + 0.3f;
      const float cb = MS[0][0] * M[1][0] + MS[0][1] * M[1][1] + MS[0][2] * M[1][2];
      const float cc = MS[1][0] * M[1][0] + MS[1][1] * M[1][1] + MS[1][2] * M[1][2] + 0.3f;
      const float det = ca * cc - cb * cb;
      const float gA = a[2], gB = a[3], gC = a[4];
      const float d2i = 1.0f / (det * det + 1e-30f);
      const float dL_da = (-cc * cc * gA + 2.f * cb * cc * gB + (det - ca * cc) * gC) * d2i;
      const float dL_dc = (-ca * ca * gC + 2.f * ca * cb * gB + (det - ca * cc) * gA) * d2i;
      const float dL_db = 2.f * (cb * cc * gA - (det + 2.f * cb * cb) * gB + ca * cb * gC) * d2i;
      const float G2[2][2] = {{dL_da, 0.5f * dL_db}, {0.5f * dL_db, dL_dc}};
      float GM[2][3];
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) GM[r][c] = G2[r][0] * M[0][c] + G2[r][1] * M[1][c];
      if (gr.dL_dcov3D) {
        float dS[3][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++) dS[r][c] = M[0][r] * GM[0][c] + M[1][r] * GM[1][c];
        float* o = gr.dL_dcov3D + gi * prm.cov_stride;
        // packed symmetric parameters (off-diagonals appear twice); the in-kernel scale s2 chains through
        const float g6[6] = {dS[0][0], dS[0][1] + dS[1][0], dS[0][2] + dS[2][0], dS[1][1], dS[1][2] + dS[2][1], dS[2][2]};
        if (prm.cov_stride == 9) {
          const int slot[6] = {0, 1, 2, 4, 5, 8};
#pragma unroll
          for (int k = 0; k < 6; k++) atomicAdd(o + slot[k], g6[k] * s2);
        } else {
#pragma unroll
          for (int k = 0; k < 6; k++) atomicAdd(o + k, g6[k] * s2);
        }
      }
      float dM[2][3];
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) dM[r][c] = 2.f * (GM[r][0] * S[0][c] + GM[r][1] * S[1][c] + GM[r][2] * S[2][c]);
      float dJ[2][3];
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) dJ[r][c] = dM[r][0] * R[c][0] + dM[r][1] * R[c][1] + dM[r][2] * R[c][2];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) dL_dR[r][c] += Jc[0][r] * dM[0][c] + Jc[1][r] * dM[1][c];
      const float tz2 = tz * tz, tz3 = tz2 * tz;
      const float dL_dtcx = -fx / tz2 * dJ[0][2];
      const float dL_dtcy = -fy / tz2 * dJ[1][2];
      float dL_dtz = -fx / tz2 * dJ[0][0] - fy / tz2 * dJ[1][1] + 2.f * fx * tcx / tz3 * dJ[0][2] +
                     2.f * fy * tcy / tz3 * dJ[1][2];
      dL_dt[0] += xg * dL_dtcx;
      dL_dt[1] += yg * dL_dtcy;
      dL_dtz += (1.f - xg) * cx * dL_dtcx + (1.f - yg) * cy * dL_dtcy;
      dL_dt[2] += dL_dtz;
    }
    // ---- (2) NDC mean through the raw projection
    {
      float h[4];
#pragma unroll
      for (int r = 0; r < 4; r++) h[r] = pr[r] * t[0] + pr[4 + r] * t[1] + pr[8 + r] * t[2] + pr[12 + r];
      const float wv = 1.0f / (h[3] + 1e-7f);
      const float gx_ = a[0], gy_ = a[1];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float dh0 = pr[4 * k + 0], dh1 = pr[4 * k + 1], dh3 = pr[4 * k + 3];
        dL_dt[k] += gx_ * (dh0 * wv - h[0] * wv * wv * dh3) + gy_ * (dh1 * wv - h[1] * wv * wv * dh3);
      }
      if (gr.dL_dmeans2D) {
        float* o = gr.dL_dmeans2D + vi * 3;
        o[0] = gx_; o[1] = gy_; o[2] = 0.f;
      }
    }
    // ---- (3) depth
    dL_dt[2] += a[9];
    // ---- (4) opacity
    if (gr.dL_dopacities) atomicAdd(gr.dL_dopacities + gi, a[5]);
    // ---- (5) colour
    const float4 cm = rgbm[vi];
    const unsigned clampmask = __float_as_uint(cm.w);
    float gc[3] = {(clampmask & 1u) ? 0.f : a[6], (clampmask & 2u) ? 0.f : a[7], (clampmask & 4u) ? 0.f : a[8]};
    if (prm.colors_precomp) {
      if (gr.dL_dcolors) {
#pragma unroll
        for (int k = 0; k < 3; k++) atomicAdd(gr.dL_dcolors + gi * 3 + k, a[6 + k]);
      }
    } else {
      const int M = prm.sh_coeffs, deg = prm.sh_degree;
      const float* sh = prm.shs + gi * M * 3;
      float* dsh = gr.dL_dshs ? gr.dL_dshs + gi * M * 3 : nullptr;
      const float* cam = prm.campos + view * 3;
      float dir[3] = {p[0] - (prm.campos ? cam[0] : 0.f), p[1] - (prm.campos ? cam[1] : 0.f),
                      p[2] - (prm.campos ? cam[2] : 0.f)};
      const float len2 = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
      const float il = rsqrtf(len2);
      const float x = dir[0] * il, y = dir[1] * il, z = dir[2] * il;
      float dRdx[3] = {0, 0, 0}, dRdy[3] = {0, 0, 0}, dRdz[3] = {0, 0, 0};
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float gg = gc[c];
        if (dsh) atomicAdd(dsh + c, SH_C0 * gg);
        if (deg > 0) {
          if (dsh) {
            atomicAdd(dsh + 3 + c, -SH_C1 * y * gg);
            atomicAdd(dsh + 6 + c, SH_C1 * z * gg);
            atomicAdd(dsh + 9 + c, -SH_C1 * x * gg);
          }
          dRdx[c] = -SH_C1 * sh[9 + c];
          dRdy[c] = -SH_C1 * sh[3 + c];
          dRdz[c] = SH_C1 * sh[6 + c];
          if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            if (dsh) {
              atomicAdd(dsh + 12 + c, bSH_C2[0] * xy * gg);
              atomicAdd(dsh + 15 + c, bSH_C2[1] * yz * gg);
              atomicAdd(dsh + 18 + c, bSH_C2[2] * (2.f * zz - xx - yy) * gg);
              atomicAdd(dsh + 21 + c, bSH_C2[3] * xz * gg);
              atomicAdd(dsh + 24 + c, bSH_C2[4] * (xx - yy) * gg);
            }
            dRdx[c] += bSH_C2[0] * y * sh[12 + c] + bSH_C2[2] * 2.f * -x * sh[18 + c] + bSH_C2[3] * z * sh[21 + c] +
                       bSH_C2[4] * 2.f * x * sh[24 + c];
            dRdy[c] += bSH_C2[0] * x * sh[12 + c] + bSH_C2[1] * z * sh[15 + c] + bSH_C2[2] * 2.f * -y * sh[18 + c] +
                       bSH_C2[4] * 2.f * -y * sh[24 + c];
            dRdz[c] += bSH_C2[1] * y * sh[15 + c] + bSH_C2[2] * 4.f * z * sh[18 + c] + bSH_C2[3] * x * sh[21 + c];
            if (deg > 2) {
              if (dsh) {
                atomicAdd(dsh + 27 + c, bSH_C3[0] * y * (3.f * xx - yy) * gg);
                atomicAdd(dsh + 30 + c, bSH_C3[1] * xy * z * gg);
                atomicAdd(dsh + 33 + c, bSH_C3[2] * y * (4.f * zz - xx - yy) * gg);
                atomicAdd(dsh + 36 + c, bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * gg);
                atomicAdd(dsh + 39 + c, bSH_C3[4] * x * (4.f * zz - xx - yy) * gg);
                atomicAdd(dsh + 42 + c, bSH_C3[5] * z * (xx - yy) * gg);
                atomicAdd(dsh + 45 + c, bSH_C3[6] * x * (xx - 3.f * yy) * gg);
              }
              dRdx[c] += bSH_C3[0] * sh[27 + c] * 6.f * xy + bSH_C3[1] * sh[30 + c] * yz +
                         bSH_C3[2] * sh[33 + c] * -2.f * xy + bSH_C3[3] * sh[36 + c] * -6.f * xz +
                         bSH_C3[4] * sh[39 + c] * (-3.f * xx + 4.f * zz - yy) + bSH_C3[5] * sh[42 + c] * 2.f * xz +
                         bSH_C3[6] * sh[45 + c] * 3.f * (xx - yy);
              dRdy[c] += bSH_C3[0] * sh[27 + c] * 3.f * (xx - yy) + bSH_C3[1] * sh[30 + c] * xz +
                         bSH_C3[2] * sh[33 + c] * (-3.f * yy + 4.f * zz - xx) + bSH_C3[3] * sh[36 + c] * -6.f * yz +
                         bSH_C3[4] * sh[39 + c] * -2.f * xy + bSH_C3[5] * sh[42 + c] * -2.f * yz +
                         bSH_C3[6] * sh[45 + c] * -6.f * xy;
              dRdz[c] += bSH_C3[1] * sh[30 + c] * xy + bSH_C3[2] * sh[33 + c] * 8.f * yz +
                         bSH_C3[3] * sh[36 + c] * 3.f * (2.f * zz - xx - yy) + bSH_C3[4] * sh[39 + c] * 8.f * xz +
                         bSH_C3[5] * sh[42 + c] * (xx - yy);
            }
          }
        }
      }
      if (deg > 0) {
        float dLd[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; c++) {
          dLd[0] += dRdx[c] * gc[c];
          dLd[1] += dRdy[c] * gc[c];
          dLd[2] += dRdz[c] * gc[c];
        }
        const float il3 = il * il * il;
        const float dot = dir[0] * dLd[0] + dir[1] * dLd[1] + dir[2] * dLd[2];
#pragma unroll
        for (int k = 0; k < 3; k++) dL_dp_direct[k] = (len2 * dLd[k] - dir[k] * dot) * il3;
      }
    }
    // ---- world-space mean (the in-kernel scale s chains through)
    if (gr.dL_dmeans3D) {
#pragma unroll
      for (int k = 0; k < 3; k++)
        atomicAdd(gr.dL_dmeans3D + gi * 3 + k,
                  (R[0][k] * dL_dt[0] + R[1][k] * dL_dt[1] + R[2][k] * dL_dt[2] + dL_dp_direct[k]) * s);
    }
    // ---- pose gradient at tau = 0 (w2c' = exp(tau) w2c)
#pragma unroll
    for (int k = 0; k < 3; k++)
      tau[k] = dL_dt[k] + R[k][0] * dL_dp_direct[0] + R[k][1] * dL_dp_direct[1] + R[k][2] * dL_dp_direct[2];
    tau[3] = t[1] * dL_dt[2] - t[2] * dL_dt[1];
    tau[4] = t[2] * dL_dt[0] - t[0] * dL_dt[2];
    tau[5] = t[0] * dL_dt[1] - t[1] * dL_dt[0];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int a1 = (k + 1) % 3, a2 = (k + 2) % 3;
      float sacc = 0.f;
#pragma unroll
      for (int c = 0; c < 3; c++) sacc += dL_dR[a2][c] * R[a1][c] - dL_dR[a1][c] * R[a2][c];
      tau[3 + k] += sacc;
    }
  }
  if (gr.dL_dtau) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      float x = tau[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((tid & 31) == 0) s_tau[tid >> 5][k] = x;
    }
    __syncthreads();
    if (tid < 6) {
      float x = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; wq++) x += s_tau[wq][tid];
      if (x != 0.f) atomicAdd(gr.dL_dtau + view * 6 + tid, x);
    }
  }
}

int& s3r_bwd_mode_override() {  // S3R_TUNE_BWD_ALL: != 0 forces the 10-value kernel (tests compare the modes against it)
  static int v = 0;
  return v;
}

extern "C" size_t s3r_raster_backward_scratch_bytes(int32_t n_views, int32_t P) {
  return (size_t)n_views * P * ACC_STRIDE * sizeof(float);
}

extern "C" int s3r_raster_backward(const s3r_raster_params* params, const void* state, size_t state_bytes,
                                   int64_t capacity, const s3r_raster_grads* grads, void* stream) {
  if (!params || !state || !grads || !grads->dL_dcolor || !grads->scratch) return S3R_ERR_INVALID_ARG;
  if (!params->projmatrix_raw) return S3R_ERR_INVALID_ARG;
  s3r_raster_layout L;
  int rc = s3r_raster_layout_query(params->n_views, params->P, params->width, params->height, capacity, &L);
  if (rc != S3R_OK) return rc;
  if ((int64_t)state_bytes < L.total_bytes) return S3R_ERR_STATE_TOO_SMALL;
  const size_t need = s3r_raster_backward_scratch_bytes(params->n_views, params->P);
  if (grads->scratch_bytes < need) return S3R_ERR_STATE_TOO_SMALL;
  cudaStream_t st = (cudaStream_t)stream;
  const char* s = (const char*)state;
  float* acc = (float*)grads->scratch;
  S3R_CUDA_CHECK(cudaMemsetAsync(acc, 0, need, st));
  dim3 g1(L.tiles * params->n_views);
  // which per-splat gradients are needed (see BwdMode): colours enter the geometry only through the view direction of
  // SH degree > 0
  const bool geom = grads->dL_dmeans3D || grads->dL_dcov3D || grads->dL_dtau || grads->dL_dmeans2D;
  const bool colour = grads->dL_dshs || grads->dL_dcolors || (geom && !params->colors_precomp && params->sh_degree > 0);
  const bool opac = grads->dL_dopacities != nullptr;
  int mode = BWD_ALL;
  if (s3r_bwd_mode_override() == 0) {
    if (geom && !colour && !opac) mode = BWD_GEOM;
    else if (colour && !geom && !opac) mode = BWD_COLOR;
  }
  if (s3r_blend_kernel_choice() == 0) {
    // warp-granular twin of the default forward kernel (the state must come from that kernel: n_contrib_blk)
    const uint32_t n_units = (uint32_t)L.tiles * (uint32_t)params->n_views * 8u;
    auto kern = mode == BWD_GEOM ? s3r_blend_bwd_blocks_kernel<BWD_GEOM>
                                 : (mode == BWD_COLOR ? s3r_blend_bwd_blocks_kernel<BWD_COLOR> : s3r_blend_bwd_blocks_kernel<BWD_ALL>);
    static size_t configured[3][64] = {};
    rc = s3r_ensure_dynamic_smem(kern, BWB_SMEM_BYTES, configured[mode]);
    if (rc != S3R_OK) return rc;
    kern<<<(n_units + BWB_WARPS - 1) / BWB_WARPS, BWB_THREADS, BWB_SMEM_BYTES, st>>>(
        params->width, params->height, params->P, L.tiles_x, L.tiles, n_units, (const unsigned*)(s + L.counters),
        (const uint32_t*)(s + L.work_order), (const uint2*)(s + L.ranges), (const float4*)(s + L.records),
        (const uint32_t*)(s + L.blists),
        (const uint32_t*)(s + L.point_list), (const float4*)(s + L.conic_opacity), params->background,
        (const float*)(s + L.final_T), (const uint32_t*)(s + L.n_contrib_blk), grads->dL_dcolor, grads->dL_ddepth, acc);
  } else {
    auto kern = mode == BWD_GEOM ? s3r_blend_bwd_kernel<BWD_GEOM>
                                 : (mode == BWD_COLOR ? s3r_blend_bwd_kernel<BWD_COLOR> : s3r_blend_bwd_kernel<BWD_ALL>);
    kern<<<g1, BB_THREADS, 0, st>>>(
        params->width, params->height, params->P, L.tiles_x, L.tiles, (const uint32_t*)(s + L.work_order),
        (const uint2*)(s + L.ranges), (const float4*)(s + L.records), (const uint32_t*)(s + L.point_list),
        (const float4*)(s + L.conic_opacity), params->background, (const float*)(s + L.final_T),
        (const uint32_t*)(s + L.n_contrib), grads->dL_dcolor, grads->dL_ddepth, acc);
  }
  S3R_CUDA_CHECK(cudaGetLastError());
  dim3 g2((params->P + 255) / 256, params->n_views);
  s3r_preprocess_bwd_kernel<<<g2, 256, 0, st>>>(*params, *grads, (const uint32_t*)(s + L.rect),
                                                (const float4*)(s + L.rgb), acc);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
