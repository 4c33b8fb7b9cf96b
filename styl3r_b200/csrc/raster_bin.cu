// Rasterizer stages 2-3: tile binning = the most-significant radix digit (view, tile) of upstream's
// 64-bit (tile | depth) key sort, done as a stable counting sort in global memory.
//
//   bin_scan : column prefix sums over the per-chunk tile histograms written by preprocess
//              (chunk_hist[view][chunk][tile]) -> chunk_base, then (last CTA) an exclusive scan over all
//              (view, tile) totals -> ranges[view*T+tile] = (start, end)  == upstream identifyTileRanges,
//              R_total / overflow / max tile count -> status.            (replaces cub::InclusiveSum + D2H)
//   bin_emit : every Gaussian writes (depth_bits << 32 | gaussian) for each tile of its rect directly to
//              its tile-major slot  start[tile] + chunk_base[chunk][tile] + rank-in-chunk, where the rank
//              comes from the chunk's 256-bit tile incidence masks.  Within a tile the slots are in
//              ascending Gaussian index = upstream's duplicateWithKeys emission order, so the following
//              stable depth sort reproduces upstream's stable 64-bit radix sort bit for bit
//              (oracle/raster_oracle.c:s3r_oracle_bin_sort).
#include "s3r_common.cuh"

// grid (ceil(T/32), n_views), 1024 threads = 32 warps x 32 tiles; warp w scans a slice of the chunks.
#define SCAN_THREADS 1024
#define SCAN_WARPS 32
__global__ void __launch_bounds__(SCAN_THREADS) s3r_bin_scan_kernel(int tiles, int chunks, int n_views,
                                                                    const uint16_t* __restrict__ chunk_hist,
                                                                    uint32_t* __restrict__ chunk_base,
                                                                    uint32_t* __restrict__ tile_count,
                                                                    uint2* __restrict__ ranges,
                                                                    long long* __restrict__ status,
                                                                    unsigned* __restrict__ counters,
                                                                    long long capacity) {
  __shared__ uint32_t s_part[SCAN_WARPS][32];
  __shared__ unsigned long long s_wsum[SCAN_WARPS];
  __shared__ uint32_t s_wmax[SCAN_WARPS];
  __shared__ unsigned long long s_carry;
  __shared__ unsigned s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int view = blockIdx.y;
  const int t = blockIdx.x * 32 + lane;
  const int cpw = (chunks + SCAN_WARPS - 1) / SCAN_WARPS;
  const int c0 = min(chunks, w * cpw), c1 = min(chunks, c0 + cpw);
  const uint16_t* hist = chunk_hist + (size_t)view * chunks * tiles;
  uint32_t* base = chunk_base + (size_t)view * chunks * tiles;
  uint32_t sum = 0;
  if (t < tiles) {
#pragma unroll 16
    for (int c = c0; c < c1; c++) sum += hist[(size_t)c * tiles + t];
  }
  s_part[w][lane] = sum;
  __syncthreads();
  uint32_t run = 0, total = 0;
#pragma unroll
  for (int k = 0; k < SCAN_WARPS; k++) {
    const uint32_t v = s_part[k][lane];
    if (k < w) run += v;
    total += v;
  }
  if (t < tiles) {
#pragma unroll 16
    for (int c = c0; c < c1; c++) {
      const uint32_t h = hist[(size_t)c * tiles + t];
      base[(size_t)c * tiles + t] = run;
      run += h;
    }
    if (w == 0) tile_count[(size_t)view * tiles + t] = total;
  }
  // ---- last CTA: exclusive scan over all (view, tile) totals -> ranges
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = atomicAdd(&counters[0], 1u);
    s_carry = 0ull;
  }
  __syncthreads();
  if (s_last != gridDim.x * gridDim.y - 1) return;
  __threadfence();
  const int n = n_views * tiles;
  const unsigned long long cap = (unsigned long long)capacity;
  uint32_t mx = 0;
  for (int i0 = 0; i0 < n; i0 += SCAN_THREADS * 4) {
    const int i = i0 + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (i + k < n) ? __ldcg(&tile_count[i + k]) : 0u;
    const unsigned long long local = (unsigned long long)v[0] + v[1] + v[2] + v[3];
    mx = max(mx, max(max(v[0], v[1]), max(v[2], v[3])));
    unsigned long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_wsum[w] = incl;
    __syncthreads();
    unsigned long long woff = s_carry, blk = 0;
    for (int k = 0; k < SCAN_WARPS; k++) {
      const unsigned long long x = s_wsum[k];
      if (k < w) woff += x;
      blk += x;
    }
    unsigned long long runp = woff + incl - local;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i + k < n) {
        const unsigned long long sb = runp, e = runp + v[k];
        ranges[i + k] = make_uint2((uint32_t)min(sb, cap), (uint32_t)min(e, cap));
        runp = e;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += blk;
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_wmax[w] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t gmx = 0;
    for (int k = 0; k < SCAN_WARPS; k++) gmx = max(gmx, s_wmax[k]);
    const unsigned long long grand = s_carry;
    status[0] = (long long)grand;
    status[1] = grand > cap ? 1 : 0;
    status[2] = gmx;
    counters[0] = 0u;
  }
}

// grid (chunks, n_views), 256 threads. dynamic smem: tiles*8 words (masks) + tiles*8 bytes (prefix by warp)
__global__ void __launch_bounds__(S3R_CHUNK) s3r_bin_emit_kernel(int P, int tiles_x, int tiles, int chunks,
                                                                 const uint32_t* __restrict__ rect_in,
                                                                 const float* __restrict__ depths,
                                                                 const uint32_t* __restrict__ chunk_base,
                                                                 const uint2* __restrict__ ranges,
                                                                 unsigned long long* __restrict__ keys_out) {
  extern __shared__ uint32_t s_mem[];
  uint32_t* s_mask = s_mem;                                         // [tiles][8]
  unsigned char* s_pre = reinterpret_cast<unsigned char*>(s_mem + (size_t)tiles * 8);  // [tiles][8]
  const int view = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < tiles * 8; i += S3R_CHUNK) s_mask[i] = 0u;
  __syncthreads();
  const int g = chunk * S3R_CHUNK + tid;
  uint32_t rect = 0u;
  if (g < P) rect = rect_in[(size_t)view * P + g];
  const int xmin = rect & 255, ymin = (rect >> 8) & 255, xmax = (rect >> 16) & 255, ymax = rect >> 24;
  const uint32_t bit = 1u << (tid & 31);
  const int w = tid >> 5;
  for (int y = ymin; y < ymax; y++)
    for (int x = xmin; x < xmax; x++) atomicOr(&s_mask[(y * tiles_x + x) * 8 + w], bit);
  __syncthreads();
  for (int t = tid; t < tiles; t += S3R_CHUNK) {
    uint32_t run = 0;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < 4) lo |= run << (8 * k); else hi |= run << (8 * (k - 4));
      run += __popc(s_mask[t * 8 + k]);
    }
    reinterpret_cast<uint2*>(s_pre)[t] = make_uint2(lo, hi);
  }
  __syncthreads();
  if (rect == 0u) return;
  const uint32_t dbits = __float_as_uint(depths[(size_t)view * P + g]);
  const unsigned long long key = ((unsigned long long)dbits << 32) | (uint32_t)g;
  const uint32_t* cb = chunk_base + ((size_t)view * chunks + chunk) * tiles;
  const uint2* rg = ranges + (size_t)view * tiles;
  const uint32_t lt = bit - 1u;
  for (int y = ymin; y < ymax; y++)
    for (int x = xmin; x < xmax; x++) {
      const int t = y * tiles_x + x;
      const uint32_t rank = s_pre[t * 8 + w] + __popc(s_mask[t * 8 + w] & lt);
      const uint2 r = rg[t];
      const uint32_t pos = r.x + cb[t] + rank;
      if (pos < r.y) keys_out[pos] = key;
    }
}

int s3r_launch_bin(const s3r_raster_params& p, const s3r_raster_layout& L, char* state, int64_t capacity,
                   cudaStream_t st) {
  dim3 g1((L.tiles + 31) / 32, p.n_views);
  s3r_bin_scan_kernel<<<g1, SCAN_THREADS, 0, st>>>(L.tiles, L.chunks, p.n_views, (const uint16_t*)(state + L.chunk_hist),
                                          (uint32_t*)(state + L.chunk_base), (uint32_t*)(state + L.tile_count),
                                          (uint2*)(state + L.ranges), (long long*)(state + L.status),
                                          (unsigned*)(state + L.counters), (long long)capacity);
  S3R_CUDA_CHECK(cudaGetLastError());
  const size_t smem = (size_t)L.tiles * 8 * sizeof(uint32_t) + (size_t)L.tiles * 8;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    S3R_CUDA_CHECK(cudaFuncSetAttribute(s3r_bin_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 g2(L.chunks, p.n_views);
  s3r_bin_emit_kernel<<<g2, S3R_CHUNK, smem, st>>>(p.P, L.tiles_x, L.tiles, L.chunks, (const uint32_t*)(state + L.rect),
                                                  (const float*)(state + L.depths),
                                                  (const uint32_t*)(state + L.chunk_base),
                                                  (const uint2*)(state + L.ranges),
                                                  (unsigned long long*)(state + L.keys_unsorted));
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
