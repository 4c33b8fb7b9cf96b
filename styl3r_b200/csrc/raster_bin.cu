// Rasterizer stages 2-3: tile binning = the most-significant radix digit (view, tile) of upstream's
// 64-bit (tile | depth) key sort, done as a stable counting sort in global memory.
//
//   bin_scan : column prefix sums over the per-chunk tile histograms written by preprocess (one warp per (view, tile),
//              8 tiles per CTA so that the u16 histogram rows are read 16 bytes at a time)
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

// grid (ceil(T/8), n_views), 256 threads = 8 warps: the CTA owns 8 adjacent tiles of one view, warp w scans tile w
// over all chunks.  The [chunks x 8] u16 slab is read with one 16-byte load per chunk row and transposed through shared
// memory; the [chunks x 8] u32 bases are written back as two 16-byte stores per chunk row.  dynamic smem:
// chunks_pad * 8 * 4 bytes (u32 [8][chunks_pad], reused in place: counts in, exclusive prefixes out).
#define SCAN_THREADS 256
#define SCAN_TILES 8
__global__ void __launch_bounds__(SCAN_THREADS) s3r_bin_scan_kernel(int tiles, int chunks, int chunks_pad, int n_views,
                                                                    const uint16_t* __restrict__ chunk_hist,
                                                                    uint32_t* __restrict__ chunk_base,
                                                                    uint32_t* __restrict__ tile_count,
                                                                    uint2* __restrict__ ranges,
                                                                    long long* __restrict__ status,
                                                                    unsigned* __restrict__ counters,
                                                                    long long capacity) {
  extern __shared__ uint32_t s_cnt[];  // [SCAN_TILES][chunks_pad + 32]: slot of chunk c = c + c / per (one pad word per lane slice: conflict-free)
  __shared__ unsigned long long s_wsum[SCAN_THREADS / 32];
  __shared__ uint32_t s_wmax[SCAN_THREADS / 32];
  __shared__ unsigned long long s_carry;
  __shared__ unsigned s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int view = blockIdx.y;
  const int t0 = blockIdx.x * SCAN_TILES;
  const uint16_t* hist = chunk_hist + (size_t)view * chunks * tiles;
  uint32_t* base = chunk_base + (size_t)view * chunks * tiles;
  const int per = chunks_pad / 32;  // chunks scanned serially by one lane (chunks_pad is a multiple of 32)
  const int pitch = chunks_pad + 32;
  const bool vec = (tiles % SCAN_TILES) == 0;  // 16-byte rows (always true for images whose tile count is a multiple of 8)
  s3r_grid_dependency_sync();
  // ---- load + transpose
  for (int c = threadIdx.x; c < chunks_pad; c += SCAN_THREADS) {
    uint32_t v[SCAN_TILES];
#pragma unroll
    for (int k = 0; k < SCAN_TILES; k++) v[k] = 0u;
    if (c < chunks) {
      if (vec) {
        const uint4 q = *reinterpret_cast<const uint4*>(hist + (size_t)c * tiles + t0);
        v[0] = q.x & 0xffffu, v[1] = q.x >> 16, v[2] = q.y & 0xffffu, v[3] = q.y >> 16;
        v[4] = q.z & 0xffffu, v[5] = q.z >> 16, v[6] = q.w & 0xffffu, v[7] = q.w >> 16;
      } else {
#pragma unroll
        for (int k = 0; k < SCAN_TILES; k++)
          if (t0 + k < tiles) v[k] = hist[(size_t)c * tiles + t0 + k];
      }
    }
#pragma unroll
    for (int k = 0; k < SCAN_TILES; k++) s_cnt[k * pitch + c + c / per] = v[k];
  }
  __syncthreads();
  // ---- warp w: exclusive scan of tile t0 + w over the chunks (lane owns a contiguous slice)
  {
    uint32_t* row = s_cnt + w * pitch + lane * (per + 1);
    uint32_t sum = 0;
    for (int i = 0; i < per; i++) sum += row[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    uint32_t run = incl - sum;
    for (int i = 0; i < per; i++) {
      const uint32_t h = row[i];
      row[i] = run;
      run += h;
    }
    if (lane == 31 && t0 + w < tiles) tile_count[(size_t)view * tiles + t0 + w] = incl;
  }
  __syncthreads();
  // ---- write the bases back, one chunk row (8 tiles = 32 bytes) per thread
  for (int c = threadIdx.x; c < chunks; c += SCAN_THREADS) {
    uint32_t v[SCAN_TILES];
#pragma unroll
    for (int k = 0; k < SCAN_TILES; k++) v[k] = s_cnt[k * pitch + c + c / per];
    if (vec) {
      uint4* dst = reinterpret_cast<uint4*>(base + (size_t)c * tiles + t0);
      dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
      dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int k = 0; k < SCAN_TILES; k++)
        if (t0 + k < tiles) base[(size_t)c * tiles + t0 + k] = v[k];
    }
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
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; k++) {
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
    for (int k = 0; k < SCAN_THREADS / 32; k++) gmx = max(gmx, s_wmax[k]);
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
  s3r_grid_dependency_sync();
  __syncthreads();
  const int g = chunk * S3R_CHUNK + tid;
  uint32_t rect = 0u;
  uint32_t dbits = 0u;  // requested together with the rect: one global round trip instead of two
  if (g < P) {
    rect = rect_in[(size_t)view * P + g];
    dbits = __float_as_uint(depths[(size_t)view * P + g]);
  }
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
  dim3 g1((L.tiles + SCAN_TILES - 1) / SCAN_TILES, p.n_views);
  const int chunks_pad = (L.chunks + 31) / 32 * 32;
  const size_t smem1 = (size_t)(chunks_pad + 32) * SCAN_TILES * sizeof(uint32_t);
  static size_t configured1[64] = {};
  int rc = s3r_ensure_dynamic_smem(s3r_bin_scan_kernel, smem1, configured1);
  if (rc != S3R_OK) return rc;
  if (smem1 > 200 * 1024) return S3R_ERR_UNSUPPORTED;  // > 6400 chunks = 1.6 M Gaussians per set
  S3R_CUDA_CHECK(s3r_launch_pdl(s3r_bin_scan_kernel, g1, dim3(SCAN_THREADS), smem1, st, (s3r_raster_pdl_mask() >> 1) & 1, L.tiles, L.chunks, chunks_pad,
                                p.n_views, (const uint16_t*)(state + L.chunk_hist), (uint32_t*)(state + L.chunk_base),
                                (uint32_t*)(state + L.tile_count), (uint2*)(state + L.ranges),
                                (long long*)(state + L.status), (unsigned*)(state + L.counters), (long long)capacity));
  const size_t smem = (size_t)L.tiles * 8 * sizeof(uint32_t) + (size_t)L.tiles * 8;
  static size_t configured2[64] = {};
  rc = s3r_ensure_dynamic_smem(s3r_bin_emit_kernel, smem, configured2);
  if (rc != S3R_OK) return rc;
  dim3 g2(L.chunks, p.n_views);
  S3R_CUDA_CHECK(s3r_launch_pdl(s3r_bin_emit_kernel, g2, dim3(S3R_CHUNK), smem, st, (s3r_raster_pdl_mask() >> 2) & 1, p.P, L.tiles_x, L.tiles, L.chunks,
                                (const uint32_t*)(state + L.rect), (const float*)(state + L.depths),
                                (const uint32_t*)(state + L.chunk_base), (const uint2*)(state + L.ranges),
                                (unsigned long long*)(state + L.keys_unsorted)));
  return S3R_OK;
}
