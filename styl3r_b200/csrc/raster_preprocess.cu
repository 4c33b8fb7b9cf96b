// Rasterizer stage 1: per-Gaussian projection, EWA covariance, tile rect, SH colour, and the per-CTA
// tile incidence histogram that replaces the upstream scan over tiles_touched.
//
// COMPILED WITH -fmad=false: depth, pixel position, radius and tile rect must be bit-identical to the
// oracle's unfused fp32 arithmetic (oracle/raster_oracle.c:s3r_oracle_preprocess), which restates
// upstream preprocessCUDA/computeCov2D as called from
// src/model/decoder/cuda_splatting.py:101-129.  The scale-invariant rescale of cuda_splatting.py:65-72
// (mean*s, cov*(s*s)) and the 3x3 -> upper-triangle gather of :118,126 are fused in here.
#include "s3r_common.cuh"

#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
__device__ __constant__ float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                           -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                           0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                           -0.5900435899266435f};

__device__ __forceinline__ float3 sh_to_rgb(int deg, int M, const float* __restrict__ sh, float3 p, const float* campos,
                                            unsigned& clampmask) {
  float x = 0.f, y = 0.f, z = 0.f;
  if (deg > 0) {  // the view direction only enters the degree >= 1 bands (Styl3R ships degree 0)
    float dx = p.x - campos[0], dy = p.y - campos[1], dz = p.z - campos[2];
    float len = sqrtf(dx * dx + dy * dy + dz * dz);
    x = dx / len, y = dy / len, z = dz / len;
  }
  float out[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float r = SH_C0 * __ldg(sh + c);
    if (deg > 0) {
      r = r - SH_C1 * y * __ldg(sh + 3 + c) + SH_C1 * z * __ldg(sh + 6 + c) - SH_C1 * x * __ldg(sh + 9 + c);
      if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + kSH_C2[0] * xy * __ldg(sh + 12 + c) + kSH_C2[1] * yz * __ldg(sh + 15 + c) +
            kSH_C2[2] * (2.0f * zz - xx - yy) * __ldg(sh + 18 + c) + kSH_C2[3] * xz * __ldg(sh + 21 + c) +
            kSH_C2[4] * (xx - yy) * __ldg(sh + 24 + c);
        if (deg > 2) {
          r = r + kSH_C3[0] * y * (3.0f * xx - yy) * __ldg(sh + 27 + c) + kSH_C3[1] * xy * z * __ldg(sh + 30 + c) +
              kSH_C3[2] * y * (4.0f * zz - xx - yy) * __ldg(sh + 33 + c) +
              kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * __ldg(sh + 36 + c) +
              kSH_C3[4] * x * (4.0f * zz - xx - yy) * __ldg(sh + 39 + c) +
              kSH_C3[5] * z * (xx - yy) * __ldg(sh + 42 + c) + kSH_C3[6] * x * (xx - 3.0f * yy) * __ldg(sh + 45 + c);
        }
      }
    }
    r += 0.5f;
    if (r < 0.0f) {
      clampmask |= (1u << c);
      r = 0.0f;
    }
    out[c] = r;
  }
  (void)M;
  return make_float3(out[0], out[1], out[2]);
}

// grid (chunks, n_views), 256 threads. dynamic smem: tiles * 8 words.
// 4 CTAs per SM (<= 64 registers): the 512 chunks of a 131 072-Gaussian view fit one wave of the 148 SMs.
__global__ void __launch_bounds__(S3R_CHUNK, 4) s3r_preprocess_kernel(
    s3r_raster_params prm, int tiles_x, int tiles_y, int chunks, float* __restrict__ depths,
    float2* __restrict__ xy_out, float4* __restrict__ conic_opacity, float4* __restrict__ rgb_out,
    uint32_t* __restrict__ rect_out, uint16_t* __restrict__ chunk_hist, int32_t* __restrict__ radii,
    float4* __restrict__ grecords, long long* __restrict__ status, unsigned* __restrict__ counters) {
  extern __shared__ uint32_t s_mask[];  // [tiles][8]
  __shared__ S3rViewConst vc;
  const int view = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int tiles = tiles_x * tiles_y;
  const int P = prm.P, W = prm.width, H = prm.height;

  for (int i = tid; i < tiles * 8; i += S3R_CHUNK) s_mask[i] = 0u;
  s3r_grid_dependency_sync();  // everything above touches only parameters / shared memory
  if (tid < 16) {
    vc.vm[tid] = prm.viewmatrix[view * 16 + tid];
    vc.pm[tid] = prm.projmatrix[view * 16 + tid];
  } else if (tid < 19) {
    vc.campos[tid - 16] = prm.campos ? prm.campos[view * 3 + tid - 16] : 0.f;
  } else if (tid == 19) {
    vc.tanx = prm.tanfov[view * 2];
    vc.tany = prm.tanfov[view * 2 + 1];
    float s = prm.scales ? prm.scales[view] : 1.0f;
    vc.scale = s;
    vc.scale2 = s * s;
    vc.set = prm.view_set ? prm.view_set[view] : view;
    // per-view constants of computeCov2D, evaluated once (same single-rounding ops as per Gaussian)
    vc.fx = W / (2.0f * vc.tanx);
    vc.fy = H / (2.0f * vc.tany);
    vc.limx = 1.3f * vc.tanx;
    vc.limy = 1.3f * vc.tany;
  }
  if (view == 0 && chunk == 0 && tid < 4) {
    status[tid] = 0;  // R_total, overflow, max_tile_count, reserved — rewritten by the bin stage
    counters[tid] = 0u;  // [0] bin_scan ticket, [1] blend work queue, [2] blend exit ticket
  }
  __syncthreads();

  const int g = chunk * S3R_CHUNK + tid;
  uint32_t rect = 0u;
  if (g < P) {
    const size_t gi = (size_t)vc.set * P + g;   // index into the Gaussian set
    const size_t vi = (size_t)view * P + g;     // index into per-view arrays
    const float s = vc.scale, s2 = vc.scale2;
    const float* mp = prm.means3D + gi * 3;
    // all inputs of this Gaussian are requested up front: one DRAM round trip instead of three dependent ones
    // (mean -> cull test -> covariance -> rect test -> colour / opacity)
    const float m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2);
    const float* cp = prm.cov3D + gi * prm.cov_stride;
    const bool full33 = prm.cov_stride == 9;
    const float r0 = __ldg(cp), r1 = __ldg(cp + 1), r2 = __ldg(cp + 2), r3 = __ldg(cp + (full33 ? 4 : 3)),
                r4 = __ldg(cp + (full33 ? 5 : 4)), r5 = __ldg(cp + (full33 ? 8 : 5));
    const float opac_in = __ldg(prm.opacities + gi);
    float sh0[3] = {0.f, 0.f, 0.f};
    if (prm.colors_precomp) {
      const float* q = prm.colors_precomp + gi * 3;
      sh0[0] = __ldg(q), sh0[1] = __ldg(q + 1), sh0[2] = __ldg(q + 2);
    } else if (prm.sh_degree == 0) {
      const float* q = prm.shs + gi * prm.sh_coeffs * 3;
      sh0[0] = __ldg(q), sh0[1] = __ldg(q + 1), sh0[2] = __ldg(q + 2);
    }
    float3 p = make_float3(m0 * s, m1 * s, m2 * s);
    const float* vm = vc.vm;
    const float* pm = vc.pm;
    float3 t;
    t.x = vm[0] * p.x + vm[4] * p.y + vm[8] * p.z + vm[12];
    t.y = vm[1] * p.x + vm[5] * p.y + vm[9] * p.z + vm[13];
    t.z = vm[2] * p.x + vm[6] * p.y + vm[10] * p.z + vm[14];
    int radius = 0;
    float depth = 0.f;
    float2 pix = make_float2(0.f, 0.f);
    float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t.z > 0.2f) {
      float hx = pm[0] * p.x + pm[4] * p.y + pm[8] * p.z + pm[12];
      float hy = pm[1] * p.x + pm[5] * p.y + pm[9] * p.z + pm[13];
      float hw = pm[3] * p.x + pm[7] * p.y + pm[11] * p.z + pm[15];
      float pw = 1.0f / (hw + 0.0000001f);
      float ndcx = hx * pw, ndcy = hy * pw;
      // packed symmetric covariance (xx,xy,xz,yy,yz,zz), scaled
      const float c0 = r0 * s2, c1 = r1 * s2, c2 = r2 * s2, c3 = r3 * s2, c4 = r4 * s2, c5 = r5 * s2;
      // --- computeCov2D, glm op order (see oracle cov2d) with the structurally-zero terms dropped
      const float fx = vc.fx, fy = vc.fy, limx = vc.limx, limy = vc.limy;
      const float txtz = t.x / t.z, tytz = t.y / t.z;
      const float tx = fminf(limx, fmaxf(-limx, txtz)) * t.z;
      const float ty = fminf(limy, fmaxf(-limy, tytz)) * t.z;
      const float J00 = fx / t.z, J02 = -(fx * tx) / (t.z * t.z);
      const float J11 = fy / t.z, J12 = -(fy * ty) / (t.z * t.z);
      // W[k][r]: W[0]=(vm0,vm4,vm8) W[1]=(vm1,vm5,vm9) W[2]=(vm2,vm6,vm10);  T[c][r], c in {0,1}
      float T0[3], T1[3];
      T0[0] = vm[0] * J00 + vm[2] * J02;  T0[1] = vm[4] * J00 + vm[6] * J02;  T0[2] = vm[8] * J00 + vm[10] * J02;
      T1[0] = vm[1] * J11 + vm[2] * J12;  T1[1] = vm[5] * J11 + vm[6] * J12;  T1[2] = vm[9] * J11 + vm[10] * J12;
      // A[k][r] = T[r][0]*V(0,k) + T[r][1]*V(1,k) + T[r][2]*V(2,k)
      const float A00 = T0[0] * c0 + T0[1] * c1 + T0[2] * c2;
      const float A10 = T0[0] * c1 + T0[1] * c3 + T0[2] * c4;
      const float A20 = T0[0] * c2 + T0[1] * c4 + T0[2] * c5;
      const float A01 = T1[0] * c0 + T1[1] * c1 + T1[2] * c2;
      const float A11 = T1[0] * c1 + T1[1] * c3 + T1[2] * c4;
      const float A21 = T1[0] * c2 + T1[1] * c4 + T1[2] * c5;
      const float cxx = (A00 * T0[0] + A10 * T0[1] + A20 * T0[2]) + 0.3f;
      const float cxy = A01 * T0[0] + A11 * T0[1] + A21 * T0[2];
      const float cyy = (A01 * T1[0] + A11 * T1[1] + A21 * T1[2]) + 0.3f;
      const float det = cxx * cyy - cxy * cxy;
      if (det != 0.0f) {
        const float det_inv = 1.f / det;
        const float mid = 0.5f * (cxx + cyy);
        const float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        const float l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        const float my_radius = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
        const float px = (float)((((double)ndcx + 1.0) * (double)W - 1.0) * 0.5);
        const float py = (float)((((double)ndcy + 1.0) * (double)H - 1.0) * 0.5);
        const int r = (int)my_radius;
        const int xmin = min(tiles_x, max(0, (int)((px - r) / S3R_TILE)));
        const int ymin = min(tiles_y, max(0, (int)((py - r) / S3R_TILE)));
        const int xmax = min(tiles_x, max(0, (int)((px + r + S3R_TILE - 1) / S3R_TILE)));
        const int ymax = min(tiles_y, max(0, (int)((py + r + S3R_TILE - 1) / S3R_TILE)));
        if ((xmax - xmin) * (ymax - ymin) != 0) {
          unsigned clampmask = 0u;
          float3 c;
          if (prm.colors_precomp) {
            c = make_float3(sh0[0], sh0[1], sh0[2]);
          } else if (prm.sh_degree == 0) {  // DC band already in registers: rgb = C0 * sh + 0.5, clamped at 0
            float o3[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
              float r = SH_C0 * sh0[ch];
              r += 0.5f;
              if (r < 0.0f) {
                clampmask |= (1u << ch);
                r = 0.0f;
              }
              o3[ch] = r;
            }
            c = make_float3(o3[0], o3[1], o3[2]);
          } else {
            c = sh_to_rgb(prm.sh_degree, prm.sh_coeffs, prm.shs + gi * prm.sh_coeffs * 3, p, vc.campos, clampmask);
          }
          radius = r;
          depth = t.z;
          pix = make_float2(px, py);
          co = make_float4(cyy * det_inv, -cxy * det_inv, cxx * det_inv, opac_in);
          col = make_float4(c.x, c.y, c.z, __uint_as_float(clampmask));
          rect = (uint32_t)xmin | ((uint32_t)ymin << 8) | ((uint32_t)xmax << 16) | ((uint32_t)ymax << 24);
          const uint32_t bit = 1u << (tid & 31);
          const int w = tid >> 5;
          for (int y = ymin; y < ymax; y++)
            for (int x = xmin; x < xmax; x++) atomicOr(&s_mask[(y * tiles_x + x) * 8 + w], bit);
        }
      }
    }
    depths[vi] = depth;
    radii[vi] = radius;
    xy_out[vi] = pix;
    conic_opacity[vi] = co;
    rgb_out[vi] = col;
    rect_out[vi] = rect;
    if (rect != 0u) {
      // 48-byte blend record of this (view, Gaussian), gathered into sorted order by the tile sort:
      //   (x, y, B', C' | A', opacity, r, g | b, depth, ex, ey)   A' = -0.5*log2(e)*A ... (s3r_common.cuh); the sort
      //   epilogue replaces (ex, ey) by the instance's cell mask
      // (ex, ey) = half-extent in pixels of { alpha >= 1/255 }: quadratic form q <= 2 ln(255 o);
      // |dx| <= sqrt(q C / det), |dy| <= sqrt(q A / det).  Not a parity quantity (a conservative cull box).
      float ex = -1.f, ey = -1.f;
      const float q = 2.0f * __logf(255.0f * co.w);
      const float det = co.x * co.z - co.y * co.y;
      if (q >= 0.f && det > 0.f) {
        const float qi = q * 1.0001f / det;
        ex = sqrtf(qi * co.z) + 0.01f;
        ey = sqrtf(qi * co.x) + 0.01f;
      } else if (!(det > 0.f) && q >= 0.f) {
        ex = ey = 1e30f;  // degenerate conic: never cull
      }
      float4* r = grecords + vi * 3;
      r[0] = make_float4(pix.x, pix.y, co.y * S3R_KB, co.z * S3R_KA);
      r[1] = make_float4(co.x * S3R_KA, co.w, col.x, col.y);
      r[2] = make_float4(col.z, depth, ex, ey);
    }
  }
  __syncthreads();
  uint16_t* hist = chunk_hist + ((size_t)view * chunks + chunk) * tiles;
  for (int t = tid; t < tiles; t += S3R_CHUNK) {
    const uint4 a = *reinterpret_cast<const uint4*>(&s_mask[t * 8]);
    const uint4 b = *reinterpret_cast<const uint4*>(&s_mask[t * 8 + 4]);
    hist[t] = (uint16_t)(__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) +
                         __popc(b.z) + __popc(b.w));
  }
}

int s3r_launch_preprocess(const s3r_raster_params& p, const s3r_raster_layout& L, char* state, int32_t* radii,
                          cudaStream_t st) {
  const size_t smem = (size_t)L.tiles * 8 * sizeof(uint32_t);
  static size_t configured[64] = {};
  int rc = s3r_ensure_dynamic_smem(s3r_preprocess_kernel, smem, configured);
  if (rc != S3R_OK) return rc;
  dim3 grid(L.chunks, p.n_views);
  S3R_CUDA_CHECK(s3r_launch_pdl(
      s3r_preprocess_kernel, grid, dim3(S3R_CHUNK), smem, st, (s3r_raster_pdl_mask() >> 0) & 1, p, L.tiles_x, L.tiles_y, L.chunks,
      (float*)(state + L.depths), (float2*)(state + L.xy), (float4*)(state + L.conic_opacity),
      (float4*)(state + L.rgb), (uint32_t*)(state + L.rect), (uint16_t*)(state + L.chunk_hist), radii,
      (float4*)(state + L.grecords), (long long*)(state + L.status), (unsigned*)(state + L.counters)));
  return S3R_OK;
}
