/*
 * raster_oracle.c — CPU restatement of the Gaussian-splatting rasterizer Styl3R calls.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under styl3r_b200/ may import, link or call this
 * file; it is the checker used by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py.
 *
 * PARITY UNPINNED: the algorithm lives in the third-party dependency
 * `diff-gaussian-rasterization-w-pose` (github.com/rmurai0610/…, requirements.txt:17 of the
 * reference, *no pinned commit*), which is absent from /root/reference and cannot be fetched.
 * This file restates the published 3DGS tile rasterizer (Kerbl et al. 2023: preprocessCUDA /
 * duplicateWithKeys / radix sort / identifyTileRanges / renderCUDA, forward and backward) plus the
 * MonoGS "-w-pose" additions (blended depth, opacity, n_touched, camera-pose gradient) as
 * summarised in SURVEY.md Appendix B, and is anchored on the reference's own call site
 * src/model/decoder/cuda_splatting.py:101-129 (argument layout, matrix conventions, cov packing).
 * Gradients are validated against torch autograd of oracle/torch_mirror.py.
 *
 * Arithmetic: fp32, one rounding per operation, no FMA contraction (compile with
 * -ffp-contract=off).  exp() in the blend is the correctly rounded fp32 exponential
 * ((float)exp((double)x)) so that the alpha >= 1/255 decision is well defined.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* matrices: m[4*col+row] (column-major as glm / as the reference passes them) */
static void xform4x3(const float* m, const float* p, float* o) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void xform4x4(const float* m, const float* p, float* o) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
  o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}
/* glm-style 3x3 (m[col][row]) product r = a*b, terms summed left to right */
static void mat3_mul(const float a[3][3], const float b[3][3], float r[3][3]) {
  for (int c = 0; c < 3; c++)
    for (int rr = 0; rr < 3; rr++)
      r[c][rr] = a[0][rr] * b[c][0] + a[1][rr] * b[c][1] + a[2][rr] * b[c][2];
}
static void mat3_t(const float a[3][3], float r[3][3]) {
  for (int c = 0; c < 3; c++)
    for (int rr = 0; rr < 3; rr++) r[c][rr] = a[rr][c];
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* EWA projection of the 3D covariance (forward.cu computeCov2D) */
static void cov2d(const float* mean, float fx, float fy, float tanx, float tany, const float* c6,
                  const float* vm, float out[3]) {
  float t[3];
  xform4x3(vm, mean, t);
  const float limx = 1.3f * tanx, limy = 1.3f * tany;
  const float txtz = t[0] / t[2], tytz = t[1] / t[2];
  t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
  t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
  float J[3][3] = {{fx / t[2], 0.0f, -(fx * t[0]) / (t[2] * t[2])},
                   {0.0f, fy / t[2], -(fy * t[1]) / (t[2] * t[2])},
                   {0.0f, 0.0f, 0.0f}};
  float Wm[3][3] = {{vm[0], vm[4], vm[8]}, {vm[1], vm[5], vm[9]}, {vm[2], vm[6], vm[10]}};
  float T[3][3], Tt[3][3], Vrk[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
  float Vt[3][3], A[3][3], C[3][3];
  mat3_mul(Wm, J, T);
  mat3_t(T, Tt);
  mat3_t(Vrk, Vt);
  mat3_mul(Tt, Vt, A);
  mat3_mul(A, T, C);
  C[0][0] += 0.3f;
  C[1][1] += 0.3f;
  out[0] = C[0][0];
  out[1] = C[0][1];
  out[2] = C[1][1];
}

static void color_from_sh(int deg, int M, const float* mean, const float* campos, const float* sh,
                          float rgb[3], uint8_t clamped[3]) {
  float dir[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
  float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  float x = dir[0] / len, y = dir[1] / len, z = dir[2] / len;
  (void)M;
  for (int c = 0; c < 3; c++) {
    float r = SH_C0 * sh[0 * 3 + c];
    if (deg > 0) {
      r = r - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
      if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
            SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
            SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
        if (deg > 2) {
          r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
              SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
              SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
              SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] + SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] +
              SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
        }
      }
    }
    r += 0.5f;
    clamped[c] = (r < 0.0f);
    rgb[c] = r < 0.0f ? 0.0f : r;
  }
}

/* ---- stage 1: preprocess (forward.cu preprocessCUDA, cov3D_precomp path) ----
 * rects[P,4] = (xmin, ymin, xmax, ymax) in tiles; everything zero for culled Gaussians. */
void s3r_oracle_preprocess(int P, int deg, int M, const float* means, const float* cov6, const float* shs,
                           const float* colors_precomp, const float* opac, const float* vm, const float* pm,
                           const float* campos, int W, int H, float tanx, float tany, int32_t* radii, float* xy,
                           float* depths, float* conic_opacity, float* rgb, uint8_t* clamped,
                           uint32_t* tiles_touched, int32_t* rects) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const float fx = W / (2.0f * tanx), fy = H / (2.0f * tany);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    xy[2 * i] = xy[2 * i + 1] = 0.f;
    depths[i] = 0.f;
    for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0.f, rects[4 * i + k] = 0;
    for (int k = 0; k < 3; k++) rgb[3 * i + k] = 0.f, clamped[3 * i + k] = 0;
    const float* p = means + 3 * i;
    float pv[3], ph[4];
    xform4x3(vm, p, pv);
    if (pv[2] <= 0.2f) continue; /* in_frustum near cull */
    xform4x4(pm, p, ph);
    const float pw = 1.0f / (ph[3] + 0.0000001f);
    const float pproj[2] = {ph[0] * pw, ph[1] * pw};
    float cov[3];
    cov2d(p, fx, fy, tanx, tany, cov6 + 6 * i, vm, cov);
    const float det = cov[0] * cov[2] - cov[1] * cov[1];
    if (det == 0.0f) continue;
    const float det_inv = 1.f / det;
    const float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
    const float mid = 0.5f * (cov[0] + cov[2]);
    const float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    const float l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    const float my_radius = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
    const float px = (float)((((double)pproj[0] + 1.0) * (double)W - 1.0) * 0.5);
    const float py = (float)((((double)pproj[1] + 1.0) * (double)H - 1.0) * 0.5);
    const int r = (int)my_radius;
    const int xmin = imin(gx, imax(0, (int)((px - r) / TILE)));
    const int ymin = imin(gy, imax(0, (int)((py - r) / TILE)));
    const int xmax = imin(gx, imax(0, (int)((px + r + TILE - 1) / TILE)));
    const int ymax = imin(gy, imax(0, (int)((py + r + TILE - 1) / TILE)));
    if ((xmax - xmin) * (ymax - ymin) == 0) continue;
    if (colors_precomp) {
      for (int k = 0; k < 3; k++) rgb[3 * i + k] = colors_precomp[3 * i + k];
    } else {
      color_from_sh(deg, M, p, campos, shs + (size_t)i * M * 3, rgb + 3 * i, clamped + 3 * i);
    }
    depths[i] = pv[2];
    radii[i] = r;
    xy[2 * i] = px;
    xy[2 * i + 1] = py;
    conic_opacity[4 * i + 0] = conic[0];
    conic_opacity[4 * i + 1] = conic[1];
    conic_opacity[4 * i + 2] = conic[2];
    conic_opacity[4 * i + 3] = opac[i];
    tiles_touched[i] = (uint32_t)((ymax - ymin) * (xmax - xmin));
    rects[4 * i + 0] = xmin;
    rects[4 * i + 1] = ymin;
    rects[4 * i + 2] = xmax;
    rects[4 * i + 3] = ymax;
  }
}

/* ---- stage 2-5: inclusive scan, duplicateWithKeys, stable radix sort, identifyTileRanges ----
 * keys/vals must hold R = sum(tiles_touched) entries; ranges is [tiles,2]. Returns R. */
int64_t s3r_oracle_bin_sort(int P, int W, int H, const int32_t* radii, const float* depths,
                            const uint32_t* tiles_touched, const int32_t* rects, uint64_t* keys_unsorted,
                            uint32_t* vals_unsorted, uint64_t* keys, uint32_t* vals, uint32_t* ranges) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  int64_t R = 0;
  for (int i = 0; i < P; i++) R += tiles_touched[i];
  if (!keys) return R;
  int64_t off = 0;
  for (int i = 0; i < P; i++) {
    if (radii[i] > 0) {
      uint32_t dbits;
      memcpy(&dbits, depths + i, 4);
      for (int y = rects[4 * i + 1]; y < rects[4 * i + 3]; y++)
        for (int x = rects[4 * i + 0]; x < rects[4 * i + 2]; x++) {
          uint64_t key = (uint64_t)(y * gx + x);
          key <<= 32;
          key |= dbits;
          keys_unsorted[off] = key;
          vals_unsorted[off] = (uint32_t)i;
          off++;
        }
    }
  }
  /* stable LSD radix sort, 8 passes of 8 bits (bit range is result-neutral) */
  uint64_t* kb = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(R ? R : 1));
  uint32_t* vb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(R ? R : 1));
  memcpy(keys, keys_unsorted, sizeof(uint64_t) * (size_t)R);
  memcpy(vals, vals_unsorted, sizeof(uint32_t) * (size_t)R);
  uint64_t *ks = keys, *kd = kb;
  uint32_t *vs = vals, *vd = vb;
  for (int pass = 0; pass < 8 && R > 0; pass++) {
    size_t hist[257];
    memset(hist, 0, sizeof(hist));
    const int sh = 8 * pass;
    for (int64_t j = 0; j < R; j++) hist[((ks[j] >> sh) & 255) + 1]++;
    if (hist[((ks[0] >> sh) & 255) + 1] == (size_t)R) continue; /* all equal: skip */
    for (int b = 0; b < 256; b++) hist[b + 1] += hist[b];
    for (int64_t j = 0; j < R; j++) {
      size_t d = hist[(ks[j] >> sh) & 255]++;
      kd[d] = ks[j];
      vd[d] = vs[j];
    }
    uint64_t* tk = ks; ks = kd; kd = tk;
    uint32_t* tv = vs; vs = vd; vd = tv;
  }
  if (ks != keys) {
    memcpy(keys, ks, sizeof(uint64_t) * (size_t)R);
    memcpy(vals, vs, sizeof(uint32_t) * (size_t)R);
  }
  free(kb);
  free(vb);
  memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)(gx * gy));
  for (int64_t j = 0; j < R; j++) {
    uint32_t t = (uint32_t)(keys[j] >> 32);
    if (j == 0) ranges[2 * t] = 0;
    else {
      uint32_t pt = (uint32_t)(keys[j - 1] >> 32);
      if (t != pt) { ranges[2 * pt + 1] = (uint32_t)j; ranges[2 * t] = (uint32_t)j; }
    }
    if (j == R - 1) ranges[2 * t + 1] = (uint32_t)R;
  }
  return R;
}

static inline float exp_cr(float x) { return (float)exp((double)x); }

/* ---- stage 6: renderCUDA forward (per-pixel sequential emulation of the tile kernel) ----
 * sens[H*W] (optional) counts decisions within a relative 1e-5 band of a threshold
 * (used by tests to list pixels where a fast-math exp could legitimately flip a branch). */
void s3r_oracle_render(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                       const float* rgb, const float* depths, const float* conic_opacity, const float* bg,
                       float* out_color, float* out_depth, float* out_opacity, float* final_T,
                       uint32_t* n_contrib, int32_t* n_touched, uint32_t* sens) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        const int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        const float pixf[2] = {(float)px, (float)py};
        float T = 1.0f, C[3] = {0, 0, 0}, D = 0.f;
        uint32_t contributor = 0, last = 0, ns = 0;
        for (uint32_t j = r0; j < r1; j++) {
          contributor++;
          const uint32_t id = point_list[j];
          const float dx = xy[2 * id] - pixf[0], dy = xy[2 * id + 1] - pixf[1];
          const float* co = conic_opacity + 4 * id;
          const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0.0f) continue;
          const float alpha = fminf(0.99f, co[3] * exp_cr(power));
          if (fabsf(alpha * 255.0f - 1.0f) < 1e-5f) ns++;
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = T * (1 - alpha);
          if (fabsf(test_T - 0.0001f) < 1e-9f) ns++;
          if (test_T < 0.0001f) break; /* done */
          for (int ch = 0; ch < 3; ch++) C[ch] += rgb[3 * id + ch] * alpha * T;
          D += depths[id] * alpha * T;
          if (test_T > 0.5f && n_touched) {
#pragma omp atomic
            n_touched[id]++;
          }
          T = test_T;
          last = contributor;
        }
        const int pix = py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        for (int ch = 0; ch < 3; ch++) out_color[ch * H * W + pix] = C[ch] + T * bg[ch];
        out_depth[pix] = D;
        out_opacity[pix] = 1.0f - T;
        if (sens) sens[pix] = ns;
      }
  }
}

/* ---- backward: renderCUDA (back-to-front) ----
 * Accumulates into dL_dmean2D[P,2] (NDC units), dL_dconic[P,3] (xx, xy, yy — the xy entry
 * holds the gradient of the single off-diagonal parameter b), dL_dopacity[P], dL_dcolor[P,3],
 * dL_ddepthg[P].  All must be zero-filled by the caller. Serial (deterministic) on purpose. */
void s3r_oracle_render_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                                const float* xy, const float* rgb, const float* depths,
                                const float* conic_opacity, const float* bg, const float* final_T,
                                const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dpixdepth,
                                float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                                float* dL_ddepthg) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  for (int tile = 0; tile < gx * gy; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        const int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        const int pix = py * W + px;
        const float pixf[2] = {(float)px, (float)py};
        const float T_final = final_T[pix];
        float T = T_final;
        const uint32_t last_contributor = n_contrib[pix];
        float accum_rec[3] = {0, 0, 0}, dLp[3], accum_d = 0.f, last_alpha = 0.f, last_color[3] = {0, 0, 0},
              last_depth = 0.f;
        for (int ch = 0; ch < 3; ch++) dLp[ch] = dL_dpix[ch * H * W + pix];
        const float dLd = dL_dpixdepth ? dL_dpixdepth[pix] : 0.f;
        uint32_t contributor = r1 - r0;
        for (uint32_t jj = r1; jj > r0; jj--) {
          const uint32_t j = jj - 1;
          contributor--;
          if (contributor >= last_contributor) continue;
          const uint32_t id = point_list[j];
          const float dx = xy[2 * id] - pixf[0], dy = xy[2 * id + 1] - pixf[1];
          const float* co = conic_opacity + 4 * id;
          const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0.0f) continue;
          const float G = exp_cr(power);
          const float alpha = fminf(0.99f, co[3] * G);
          if (alpha < 1.0f / 255.0f) continue;
          T = T / (1.f - alpha);
          const float dchannel_dcolor = alpha * T;
          float dL_dalpha = 0.0f;
          for (int ch = 0; ch < 3; ch++) {
            const float c = rgb[3 * id + ch];
            accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
            last_color[ch] = c;
            dL_dalpha += (c - accum_rec[ch]) * dLp[ch];
            dL_dcolor[3 * id + ch] += dchannel_dcolor * dLp[ch];
          }
          const float cd = depths[id];
          accum_d = last_alpha * last_depth + (1.f - last_alpha) * accum_d;
          last_depth = cd;
          dL_dalpha += (cd - accum_d) * dLd;
          dL_ddepthg[id] += dchannel_dcolor * dLd;
          dL_dalpha *= T;
          last_alpha = alpha;
          float bg_dot = 0.f;
          for (int ch = 0; ch < 3; ch++) bg_dot += bg[ch] * dLp[ch];
          dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
          const float dL_dG = co[3] * dL_dalpha;
          const float gdx = G * dx, gdy = G * dy;
          const float dG_ddelx = -gdx * co[0] - gdy * co[1];
          const float dG_ddely = -gdy * co[2] - gdx * co[1];
          dL_dmean2D[2 * id + 0] += dL_dG * dG_ddelx * ddelx_dx;
          dL_dmean2D[2 * id + 1] += dL_dG * dG_ddely * ddely_dy;
          dL_dconic[3 * id + 0] += -0.5f * gdx * dx * dL_dG;
          dL_dconic[3 * id + 1] += -0.5f * gdx * dy * dL_dG;
          dL_dconic[3 * id + 2] += -0.5f * gdy * dy * dL_dG;
          dL_dopacity[id] += G * dL_dalpha;
        }
      }
  }
}

/* ---- backward: preprocess (computeCov2D bwd + projection bwd + depth + SH bwd + pose) ----
 * Inputs are the per-Gaussian gradients produced above.  Outputs (overwritten, P-sized):
 * dL_dmeans[P,3], dL_dcov6[P,6], dL_dsh[P,M,3], dL_dtau[P,6] = per-Gaussian (rho, theta) gradient of
 * the left-multiplied pose perturbation w2c' = SE3_exp(tau) w2c at tau = 0.
 * Written in double precision internally: it is the *checker* for the fp32 CUDA kernel. */
void s3r_oracle_preprocess_backward(int P, int deg, int M, const float* means, const float* cov6, const float* shs,
                                    int use_sh, const float* vm, const float* pm_raw, const float* campos, int W,
                                    int H, float tanx, float tany, const int32_t* radii, const uint8_t* clamped,
                                    const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
                                    const float* dL_ddepthg, float* dL_dmeans, float* dL_dcov6, float* dL_dsh,
                                    float* dL_dtau) {
  const double fx = W / (2.0 * (double)tanx), fy = H / (2.0 * (double)tany);
  /* rotation rows of the world->camera matrix: Rm[r][c] */
  double Rm[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Rm[r][c] = vm[4 * c + r];
  for (int i = 0; i < P; i++) {
    for (int k = 0; k < 3; k++) dL_dmeans[3 * i + k] = 0.f;
    for (int k = 0; k < 6; k++) dL_dcov6[6 * i + k] = 0.f, dL_dtau[6 * i + k] = 0.f;
    if (dL_dsh)
      for (int k = 0; k < 3 * M; k++) dL_dsh[(size_t)i * 3 * M + k] = 0.f;
    if (radii[i] <= 0) continue;
    const float* p = means + 3 * i;
    /* camera-space point t = R p + tr */
    double t[3];
    for (int r = 0; r < 3; r++) t[r] = Rm[r][0] * p[0] + Rm[r][1] * p[1] + Rm[r][2] * p[2] + vm[12 + r];
    double dL_dt[3] = {0, 0, 0};     /* gradient w.r.t. the camera-space point  */
    double dL_dR[3][3] = {{0}};      /* gradient w.r.t. the rotation (cov path) */
    double dL_dp_direct[3] = {0, 0, 0}; /* gradient w.r.t. world point not via t (SH view dir) */

    /* (1) conic -> cov2D:  conic = inverse([[a,b],[b,c]]) */
    {
      const double S[3][3] = {{cov6[6 * i + 0], cov6[6 * i + 1], cov6[6 * i + 2]},
                              {cov6[6 * i + 1], cov6[6 * i + 3], cov6[6 * i + 4]},
                              {cov6[6 * i + 2], cov6[6 * i + 4], cov6[6 * i + 5]}};
      const double limx = 1.3 * tanx, limy = 1.3 * tany;
      const double txtz = t[0] / t[2], tytz = t[1] / t[2];
      const double cx = fmin(limx, fmax(-limx, txtz)), cy = fmin(limy, fmax(-limy, tytz));
      const double x_grad_mul = (txtz < -limx || txtz > limx) ? 0.0 : 1.0;
      const double y_grad_mul = (tytz < -limy || tytz > limy) ? 0.0 : 1.0;
      const double tcx = cx * t[2], tcy = cy * t[2], tz = t[2];
      /* conventional Jacobian Jc (2x3) and M = Jc R (2x3); cov2D = M S M^T + 0.3 I */
      const double Jc[2][3] = {{fx / tz, 0, -fx * tcx / (tz * tz)}, {0, fy / tz, -fy * tcy / (tz * tz)}};
      double Mx[2][3];
      for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++) Mx[r][c] = Jc[r][0] * Rm[0][c] + Jc[r][1] * Rm[1][c] + Jc[r][2] * Rm[2][c];
      double MS[2][3];
      for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++) MS[r][c] = Mx[r][0] * S[0][c] + Mx[r][1] * S[1][c] + Mx[r][2] * S[2][c];
      const double a = MS[0][0] * Mx[0][0] + MS[0][1] * Mx[0][1] + MS[0][2] * Mx[0][2] + 0.3;
      const double b = MS[0][0] * Mx[1][0] + MS[0][1] * Mx[1][1] + MS[0][2] * Mx[1][2];
      const double c = MS[1][0] * Mx[1][0] + MS[1][1] * Mx[1][1] + MS[1][2] * Mx[1][2] + 0.3;
      const double det = a * c - b * b;
      const double gA = dL_dconic[3 * i + 0], gB = dL_dconic[3 * i + 1], gC = dL_dconic[3 * i + 2];
      /* conic = (c, -b, a)/det.  As in the upstream blend backward, dL_dconic.y carries HALF of the
       * derivative w.r.t. the off-diagonal conic parameter (gradient per symmetric matrix entry). */
      const double d2 = det * det + 1e-300;
      const double dL_da = (-c * c * gA + 2 * b * c * gB + (det - a * c) * gC) / d2;
      const double dL_dc = (-a * a * gC + 2 * a * b * gB + (det - a * c) * gA) / d2;
      const double dL_db = 2 * (b * c * gA - (det + 2 * b * b) * gB + a * b * gC) / d2;
      /* cov2D = M S M^T: dL/dS = M^T G M with G = [[da, db/2],[db/2, dc]] */
      const double G2[2][2] = {{dL_da, 0.5 * dL_db}, {0.5 * dL_db, dL_dc}};
      double GM[2][3];
      for (int r = 0; r < 2; r++)
        for (int cc = 0; cc < 3; cc++) GM[r][cc] = G2[r][0] * Mx[0][cc] + G2[r][1] * Mx[1][cc];
      double dS[3][3];
      for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 3; cc++) dS[r][cc] = Mx[0][r] * GM[0][cc] + Mx[1][r] * GM[1][cc];
      /* packed symmetric parameters: off-diagonals appear twice */
      dL_dcov6[6 * i + 0] = (float)dS[0][0];
      dL_dcov6[6 * i + 1] = (float)(dS[0][1] + dS[1][0]);
      dL_dcov6[6 * i + 2] = (float)(dS[0][2] + dS[2][0]);
      dL_dcov6[6 * i + 3] = (float)dS[1][1];
      dL_dcov6[6 * i + 4] = (float)(dS[1][2] + dS[2][1]);
      dL_dcov6[6 * i + 5] = (float)dS[2][2];
      /* dL/dM = 2 G M S  (2x3) */
      double dM[2][3];
      for (int r = 0; r < 2; r++)
        for (int cc = 0; cc < 3; cc++)
          dM[r][cc] = 2 * (GM[r][0] * S[0][cc] + GM[r][1] * S[1][cc] + GM[r][2] * S[2][cc]);
      /* M = Jc R: dL/dJc = dM R^T ; dL/dR += Jc^T dM */
      double dJ[2][3];
      for (int r = 0; r < 2; r++)
        for (int cc = 0; cc < 3; cc++) dJ[r][cc] = dM[r][0] * Rm[cc][0] + dM[r][1] * Rm[cc][1] + dM[r][2] * Rm[cc][2];
      for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 3; cc++) dL_dR[r][cc] += Jc[0][r] * dM[0][cc] + Jc[1][r] * dM[1][cc];
      /* Jc entries as functions of (tcx, tcy, tz) */
      const double tz2 = tz * tz, tz3 = tz2 * tz;
      const double dL_dtcx = -fx / tz2 * dJ[0][2];
      const double dL_dtcy = -fy / tz2 * dJ[1][2];
      double dL_dtz = -fx / tz2 * dJ[0][0] - fy / tz2 * dJ[1][1] + 2 * fx * tcx / tz3 * dJ[0][2] +
                      2 * fy * tcy / tz3 * dJ[1][2];
      /* tcx = clamp(tx/tz)*tz */
      dL_dt[0] += x_grad_mul * dL_dtcx;
      dL_dt[1] += y_grad_mul * dL_dtcy;
      dL_dtz += (1 - x_grad_mul) * cx * dL_dtcx + (1 - y_grad_mul) * cy * dL_dtcy;
      dL_dt[2] += dL_dtz;
    }
    /* (2) mean2D (NDC) through the raw projection: ndc = (P t).xy / ((P t).w + 1e-7) */
    {
      double h[4];
      for (int r = 0; r < 4; r++) h[r] = pm_raw[r] * t[0] + pm_raw[4 + r] * t[1] + pm_raw[8 + r] * t[2] + pm_raw[12 + r];
      const double w = 1.0 / (h[3] + 1e-7);
      const double gx_ = dL_dmean2D[2 * i + 0], gy_ = dL_dmean2D[2 * i + 1];
      for (int k = 0; k < 3; k++) {
        const double dh0 = pm_raw[4 * k + 0], dh1 = pm_raw[4 * k + 1], dh3 = pm_raw[4 * k + 3];
        dL_dt[k] += gx_ * (dh0 * w - h[0] * w * w * dh3) + gy_ * (dh1 * w - h[1] * w * w * dh3);
      }
    }
    /* (3) depth = t.z */
    dL_dt[2] += dL_ddepthg[i];
    /* (4) colour from SH */
    if (use_sh && dL_dsh) {
      double dir[3] = {p[0] - campos[0], p[1] - campos[1], p[2] - campos[2]};
      const double len = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
      const double x = dir[0] / len, y = dir[1] / len, z = dir[2] / len;
      double dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
      for (int c = 0; c < 3; c++) {
        const double g = clamped[3 * i + c] ? 0.0 : (double)dL_dcolor[3 * i + c];
        const float* sh = shs + (size_t)i * M * 3;
        float* dsh = dL_dsh + (size_t)i * M * 3;
        dsh[0 * 3 + c] = (float)(SH_C0 * g);
        if (deg > 0) {
          dsh[1 * 3 + c] = (float)(-SH_C1 * y * g);
          dsh[2 * 3 + c] = (float)(SH_C1 * z * g);
          dsh[3 * 3 + c] = (float)(-SH_C1 * x * g);
          dRGBdx[c] = -SH_C1 * sh[3 * 3 + c];
          dRGBdy[c] = -SH_C1 * sh[1 * 3 + c];
          dRGBdz[c] = SH_C1 * sh[2 * 3 + c];
          if (deg > 1) {
            const double xx = x * x, yy = y * y, zz = z * z, xy_ = x * y, yz = y * z, xz = x * z;
            dsh[4 * 3 + c] = (float)(SH_C2[0] * xy_ * g);
            dsh[5 * 3 + c] = (float)(SH_C2[1] * yz * g);
            dsh[6 * 3 + c] = (float)(SH_C2[2] * (2.0 * zz - xx - yy) * g);
            dsh[7 * 3 + c] = (float)(SH_C2[3] * xz * g);
            dsh[8 * 3 + c] = (float)(SH_C2[4] * (xx - yy) * g);
            dRGBdx[c] += SH_C2[0] * y * sh[4 * 3 + c] + SH_C2[2] * 2.0 * -x * sh[6 * 3 + c] +
                         SH_C2[3] * z * sh[7 * 3 + c] + SH_C2[4] * 2.0 * x * sh[8 * 3 + c];
            dRGBdy[c] += SH_C2[0] * x * sh[4 * 3 + c] + SH_C2[1] * z * sh[5 * 3 + c] +
                         SH_C2[2] * 2.0 * -y * sh[6 * 3 + c] + SH_C2[4] * 2.0 * -y * sh[8 * 3 + c];
            dRGBdz[c] += SH_C2[1] * y * sh[5 * 3 + c] + SH_C2[2] * 2.0 * 2.0 * z * sh[6 * 3 + c] +
                         SH_C2[3] * x * sh[7 * 3 + c];
            if (deg > 2) {
              dsh[9 * 3 + c] = (float)(SH_C3[0] * y * (3.0 * xx - yy) * g);
              dsh[10 * 3 + c] = (float)(SH_C3[1] * xy_ * z * g);
              dsh[11 * 3 + c] = (float)(SH_C3[2] * y * (4.0 * zz - xx - yy) * g);
              dsh[12 * 3 + c] = (float)(SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * g);
              dsh[13 * 3 + c] = (float)(SH_C3[4] * x * (4.0 * zz - xx - yy) * g);
              dsh[14 * 3 + c] = (float)(SH_C3[5] * z * (xx - yy) * g);
              dsh[15 * 3 + c] = (float)(SH_C3[6] * x * (xx - 3.0 * yy) * g);
              dRGBdx[c] += SH_C3[0] * sh[9 * 3 + c] * 3.0 * 2.0 * xy_ + SH_C3[1] * sh[10 * 3 + c] * yz +
                           SH_C3[2] * sh[11 * 3 + c] * -2.0 * xy_ + SH_C3[3] * sh[12 * 3 + c] * -3.0 * 2.0 * xz +
                           SH_C3[4] * sh[13 * 3 + c] * (-3.0 * xx + 4.0 * zz - yy) +
                           SH_C3[5] * sh[14 * 3 + c] * 2.0 * xz + SH_C3[6] * sh[15 * 3 + c] * 3.0 * (xx - yy);
              dRGBdy[c] += SH_C3[0] * sh[9 * 3 + c] * 3.0 * (xx - yy) + SH_C3[1] * sh[10 * 3 + c] * xz +
                           SH_C3[2] * sh[11 * 3 + c] * (-3.0 * yy + 4.0 * zz - xx) +
                           SH_C3[3] * sh[12 * 3 + c] * -3.0 * 2.0 * yz + SH_C3[4] * sh[13 * 3 + c] * -2.0 * xy_ +
                           SH_C3[5] * sh[14 * 3 + c] * -2.0 * yz + SH_C3[6] * sh[15 * 3 + c] * -3.0 * 2.0 * xy_;
              dRGBdz[c] += SH_C3[1] * sh[10 * 3 + c] * xy_ + SH_C3[2] * sh[11 * 3 + c] * 4.0 * 2.0 * yz +
                           SH_C3[3] * sh[12 * 3 + c] * 3.0 * (2.0 * zz - xx - yy) +
                           SH_C3[4] * sh[13 * 3 + c] * 4.0 * 2.0 * xz + SH_C3[5] * sh[14 * 3 + c] * (xx - yy);
            }
          }
        }
      }
      if (deg > 0) {
        double dLddir[3] = {0, 0, 0};
        for (int c = 0; c < 3; c++) {
          const double g = clamped[3 * i + c] ? 0.0 : (double)dL_dcolor[3 * i + c];
          dLddir[0] += dRGBdx[c] * g;
          dLddir[1] += dRGBdy[c] * g;
          dLddir[2] += dRGBdz[c] * g;
        }
        /* normalisation backward: d(v/|v|) */
        const double s2 = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
        const double il3 = 1.0 / (s2 * sqrt(s2));
        const double dot = dir[0] * dLddir[0] + dir[1] * dLddir[1] + dir[2] * dLddir[2];
        for (int k = 0; k < 3; k++) dL_dp_direct[k] = (s2 * dLddir[k] - dir[k] * dot) * il3;
      }
    }
    /* world-space mean gradient: p -> t = R p + tr */
    for (int k = 0; k < 3; k++)
      dL_dmeans[3 * i + k] =
          (float)(Rm[0][k] * dL_dt[0] + Rm[1][k] * dL_dt[1] + Rm[2][k] * dL_dt[2] + dL_dp_direct[k]);
    /* pose gradient at tau = 0 for w2c' = exp(tau) w2c: t' = t + rho + theta x t ; R' = (I + [theta]x) R.
     * campos' = -R'^T tr' moves too, but colour depends on p - campos only through the direction, whose
     * camera-frame image (R(p - campos) = t) rotates with theta; handled via dL_dp_direct below. */
    double g_rho[3] = {dL_dt[0], dL_dt[1], dL_dt[2]};
    double g_th[3] = {t[1] * dL_dt[2] - t[2] * dL_dt[1], t[2] * dL_dt[0] - t[0] * dL_dt[2],
                      t[0] * dL_dt[1] - t[1] * dL_dt[0]};
    /* rotation through the covariance: dR = [theta]x R  => dL/dtheta_k = sum_ij dL_dR_ij ([e_k]x R)_ij */
    for (int k = 0; k < 3; k++) {
      const int a1 = (k + 1) % 3, a2 = (k + 2) % 3; /* ([e_k]x R)[a2][:] = R[a1][:], [a1][:] = -R[a2][:] */
      double s = 0;
      for (int c = 0; c < 3; c++) s += dL_dR[a2][c] * Rm[a1][c] - dL_dR[a1][c] * Rm[a2][c];
      g_th[k] += s;
    }
    /* view-direction path (SH degree > 0): dir_world = p - campos with campos = -R^T tr.  Under the
     * perturbation campos' = -R'^T tr' ; d campos/d rho = -R^T, d campos/d theta_k = -R^T ([e_k]x)^T ...
     * = R^T [e_k]x tr' - ... ; in camera frame: R(p - campos') = t' - tr' - ... we use the identity
     * p - campos = R^T t (world frame), so d(p-campos) = dR^T t + R^T dt = R^T(-[theta]x t + rho + theta x t) = R^T rho. */
    {
      const double gw[3] = {dL_dp_direct[0], dL_dp_direct[1], dL_dp_direct[2]};
      for (int k = 0; k < 3; k++) g_rho[k] += Rm[k][0] * gw[0] + Rm[k][1] * gw[1] + Rm[k][2] * gw[2];
    }
    for (int k = 0; k < 3; k++) {
      dL_dtau[6 * i + k] = (float)g_rho[k];
      dL_dtau[6 * i + 3 + k] = (float)g_th[k];
    }
  }
}

void s3r_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  extern void omp_set_num_threads(int);
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int s3r_oracle_num_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
