// Rasterizer stage 4: the remaining radix digits.  After bin_emit every tile owns a contiguous segment of
// (depth_bits << 32 | gaussian) keys in ascending-gaussian order; one CTA per (view, tile) runs a stable
// LSD radix sort over the 32 depth bits (4 passes x 8 bits, passes whose digit is uniform are skipped).
// Segments of up to S3R_SORT_SMEM_CAP keys are sorted entirely in shared memory; larger ones ping-pong
// between two global-memory buffers with the same code.  The result equals upstream's
// cub::DeviceRadixSort::SortPairs over ((tile << 32) | depth_bits, gaussian) bit for bit
// (oracle/raster_oracle.c:s3r_oracle_bin_sort).
//
// The epilogue writes point_list (sorted gaussian ids), optionally the 64-bit upstream-format keys, and
// the 48-byte sorted-gathered blend records (conic pre-scaled to the log2 domain) that the blend kernels stream with TMA bulk copies:
//   (x, y, conicA, conicB | conicC, opacity, r, g | b, depth, ex, ey)
// where (ex, ey) is the half-extent in pixels of the region where alpha can reach 1/255.
#include "s3r_common.cuh"

#define SORT_THREADS 256
#define SORT_WARPS 8

template <bool kGlobal>
__device__ __forceinline__ unsigned long long ld_key(const unsigned long long* p) {
  if (kGlobal) return __ldcg(p);
  return *p;
}

// One stable counting pass on digit (key >> shift) & 255. Returns false (and leaves dst untouched) if all
// keys share the digit.  hist: [SORT_WARPS][256] in shared memory.
template <bool kGlobal>
__device__ bool radix_pass(const unsigned long long* src, unsigned long long* dst, uint32_t n, int shift,
                           uint32_t (*hist)[256], uint32_t* s_scan, int* s_flag) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) (&hist[0][0])[i] = 0u;
  if (tid == 0) *s_flag = 0;
  __syncthreads();
  uint32_t per = (n + SORT_WARPS - 1) / SORT_WARPS;
  per = (per + 31u) & ~31u;
  const uint32_t b0 = min(n, w * per), b1 = min(n, b0 + per);
  // count: fire-and-forget shared-memory reductions into the warp's private histogram (no dependency chain)
  for (uint32_t i = b0 + lane; i < b1; i += 32)
    atomicAdd(&hist[w][(uint32_t)((ld_key<kGlobal>(src + i) >> shift) & 255ull)], 1u);
  __syncthreads();
  // thread d owns digit d: total over warps -> exclusive scan over digits -> per-warp bases
  uint32_t tot = 0;
#pragma unroll
  for (int k = 0; k < SORT_WARPS; k++) tot += hist[k][tid];
  if (tot == n) *s_flag = 1;
  uint32_t incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_scan[w] = incl;
  __syncthreads();
  const int uniform = *s_flag;
  __syncthreads();  // everyone has read the flag before a following pass may reset it
  if (uniform) return false;
  uint32_t base = incl - tot;
#pragma unroll
  for (int k = 0; k < SORT_WARPS; k++)
    if (k < w) base += s_scan[k];
#pragma unroll
  for (int k = 0; k < SORT_WARPS; k++) {
    const uint32_t c = hist[k][tid];
    hist[k][tid] = base;
    base += c;
  }
  __syncthreads();
  // scatter: stable ranks inside the warp from 8 independent ballots on the digit bits (fixed latency; MATCH.ANY's
  // latency grows with the number of distinct digits in the warp, which is ~32 for mantissa bytes)
  for (uint32_t i = b0 + lane; (i - lane) < b1; i += 32) {
    const bool valid = i < b1;
    unsigned long long key = 0;
    uint32_t d = 0;
    if (valid) {
      key = ld_key<kGlobal>(src + i);
      d = (uint32_t)((key >> shift) & 255ull);
    }
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
      const bool set = (d >> bit) & 1u;
      const uint32_t m = __ballot_sync(0xffffffffu, set);
      peers &= set ? m : ~m;
    }
    uint32_t base_d = 0;
    if (valid) base_d = hist[w][d];
    __syncwarp();
    if (valid) {
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      dst[base_d + rank] = key;
      if (rank == 0) hist[w][d] = base_d + __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  return true;
}

// grid (tiles, n_views)
__global__ void __launch_bounds__(SORT_THREADS) s3r_tile_sort_kernel(
    int P, int tiles, const uint2* __restrict__ ranges, unsigned long long* __restrict__ keys_a,
    unsigned long long* __restrict__ keys_b, const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
    const float4* __restrict__ rgb, uint32_t* __restrict__ point_list, unsigned long long* __restrict__ point_keys,
    float4* __restrict__ records) {
  extern __shared__ unsigned long long s_keys[];  // [2][S3R_SORT_SMEM_CAP]
  __shared__ uint32_t s_hist[SORT_WARPS][256];
  __shared__ uint32_t s_scan[SORT_WARPS];
  __shared__ int s_flag;
  const int view = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const uint2 rg = ranges[(size_t)view * tiles + tile];
  const uint32_t n = rg.y - rg.x;
  if (n == 0) return;
  unsigned long long* ga = keys_a + rg.x;
  unsigned long long* gb = keys_b + rg.x;
  const unsigned long long* sorted;
  if (n <= S3R_SORT_SMEM_CAP) {
    unsigned long long* cur = s_keys;
    unsigned long long* nxt = s_keys + S3R_SORT_SMEM_CAP;
    for (uint32_t i = tid; i < n; i += SORT_THREADS) cur[i] = __ldcs(ga + i);
    __syncthreads();
    if (n > 1) {
#pragma unroll 1
      for (int pass = 0; pass < 4; pass++) {
        if (radix_pass<false>(cur, nxt, n, 32 + 8 * pass, s_hist, s_scan, &s_flag)) {
          unsigned long long* t = cur; cur = nxt; nxt = t;
        }
      }
    }
    sorted = cur;
  } else {
    unsigned long long* cur = ga;
    unsigned long long* nxt = gb;
#pragma unroll 1
    for (int pass = 0; pass < 4; pass++) {
      if (radix_pass<true>(cur, nxt, n, 32 + 8 * pass, s_hist, s_scan, &s_flag)) {
        unsigned long long* t = cur; cur = nxt; nxt = t;
      }
      __threadfence_block();
    }
    sorted = cur;
  }
  // epilogue: ids, keys, gathered records
  const size_t vbase = (size_t)view * P;
  const unsigned long long tile_hi = ((unsigned long long)((uint32_t)view * (uint32_t)tiles + (uint32_t)tile)) << 32;
#pragma unroll 4
  for (uint32_t i = tid; i < n; i += SORT_THREADS) {
    const unsigned long long k = (n <= S3R_SORT_SMEM_CAP) ? sorted[i] : __ldcg(sorted + i);
    const uint32_t id = (uint32_t)k, dbits = (uint32_t)(k >> 32);
    const size_t o = (size_t)rg.x + i;
    point_list[o] = id;
    if (point_keys) point_keys[o] = tile_hi | dbits;
    const float2 p = __ldg(&xy[vbase + id]);
    const float4 co = __ldg(&conic_opacity[vbase + id]);
    const float4 c = __ldg(&rgb[vbase + id]);
    // half-extent of { alpha >= 1/255 }: quadratic form q <= 2*ln(255*o); |dx| <= sqrt(q*C/det), |dy| <= sqrt(q*A/det)
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
    float4* r = records + o * 3;
    r[0] = make_float4(p.x, p.y, co.x * S3R_KA, co.y * S3R_KB);   // log2-domain conic (s3r_common.cuh)
    r[1] = make_float4(co.z * S3R_KA, co.w, c.x, c.y);
    r[2] = make_float4(c.z, __uint_as_float(dbits), ex, ey);
  }
}

int s3r_launch_sort(const s3r_raster_params& p, const s3r_raster_layout& L, char* state, cudaStream_t st) {
  const size_t smem = 2 * (size_t)S3R_SORT_SMEM_CAP * sizeof(unsigned long long);
  static bool configured = false;
  if (!configured) {
    S3R_CUDA_CHECK(cudaFuncSetAttribute(s3r_tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(L.tiles, p.n_views);
  s3r_tile_sort_kernel<<<grid, SORT_THREADS, smem, st>>>(
      p.P, L.tiles, (const uint2*)(state + L.ranges), (unsigned long long*)(state + L.keys_unsorted),
      (unsigned long long*)(state + L.keys_tmp), (const float2*)(state + L.xy),
      (const float4*)(state + L.conic_opacity), (const float4*)(state + L.rgb), (uint32_t*)(state + L.point_list),
      (unsigned long long*)(state + L.point_keys), (float4*)(state + L.records));
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
