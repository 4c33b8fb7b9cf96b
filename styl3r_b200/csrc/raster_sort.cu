// Rasterizer stage 4: depth order inside every tile.  After bin_emit every tile owns a contiguous segment of
// (depth_bits << 32 | gaussian) keys in ascending-gaussian order.  Upstream sorts ((tile << 32) | depth_bits, gaussian)
// with a STABLE radix sort whose input is in ascending-gaussian order, so its result inside a tile is exactly the
// ascending order of the full 64-bit key (depth_bits, gaussian) - a strict total order (gaussian ids are unique in a
// tile).  Any correct sort of those 64-bit keys therefore reproduces upstream's order bit for bit
// (oracle/raster_oracle.c:s3r_oracle_bin_sort).  One CTA per (view, tile), three paths:
//
//   bucket path  (n <= S3R_SORT_SMEM_CAP, the normal case): ONE most-significant-digit pass - 1024 buckets over the
//                tile's own depth range [min, max] of the float bits (monotone in depth for the positive depths that
//                survive the near cull) - followed by an exact rank-by-counting inside each bucket on the full 64-bit
//                key (buckets hold ~1-3 keys for pixel-aligned scenes; 5 CTA barriers in total instead of ~28 for four
//                LSD passes).  If some bucket holds more than SORT_MAX_BUCKET keys (many equal / clustered depths with far
//                outliers) the tile falls back to
//   LSD path     stable LSD radix sort over the 32 depth bits (4 passes x 8 bits, uniform digits skipped) in shared memory,
//   global path  the same LSD code ping-ponging between two global buffers for tiles above S3R_SORT_SMEM_CAP keys.
//
// The epilogue writes point_list (sorted gaussian ids), the upstream-format 64-bit keys, and gathers the 48-byte blend
// records that the preprocess stage wrote per Gaussian into sorted order (three 16-byte loads + stores per instance),
// replacing their cull half-extents by the instance's 16-bit cell mask for this tile (cell_mask).  A last pass
// (build_block_lists) compacts, per 8x4-pixel block of the tile, the sorted positions whose mask touches the block: the
// work lists of the warp-granular blend kernels (raster_blend_blocks.cu, raster_backward.cu).
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

#define SORT_NB 1024        // buckets of the most-significant-digit pass
#define SORT_MAX_BUCKET 48  // above this the rank-by-counting step could go quadratic: use the LSD path
#define SORT_KPT ((S3R_SORT_SMEM_CAP + SORT_THREADS - 1) / SORT_THREADS)  // keys per thread in the bucket path

// 16-bit incidence mask of the splat's { alpha >= 1/255 } box (centre (x, y), half-extents (ex, ey) from preprocess) over
// the 4x4-pixel cells of the 16x16 tile at (tx0, ty0): bit 4*cy + cx.  Built once per instance here, so that the blend
// kernels' per-warp cull is one bit test per record instead of four float comparisons repeated by all 8 warps.
// Conservative like the box itself (never a parity quantity); ex < 0 (alpha can never reach 1/255) gives 0.
__device__ __forceinline__ uint32_t cell_mask(float x, float y, float ex, float ey, float tx0, float ty0) {
  if (!(ex >= 0.f)) return 0u;
  const float xlo = x - ex, xhi = x + ex, ylo = y - ey, yhi = y + ey;
  uint32_t mx = 0u, my = 0u;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    mx |= (xhi >= tx0 + 4.f * i && xlo <= tx0 + 4.f * i + 3.f) ? (1u << i) : 0u;
    my |= (yhi >= ty0 + 4.f * i && ylo <= ty0 + 4.f * i + 3.f) ? (1u << (4 * i)) : 0u;
  }
  return mx * my;  // outer product: no carries (my has one bit per nibble)
}

__device__ __forceinline__ uint32_t write_sorted(uint32_t o, unsigned long long k, unsigned long long tile_hi, size_t vbase,
                                             float tx0, float ty0, const float4* __restrict__ grecords,
                                             uint32_t* __restrict__ point_list,
                                             unsigned long long* __restrict__ point_keys, float4* __restrict__ records) {
  const uint32_t id = (uint32_t)k, dbits = (uint32_t)(k >> 32);
  point_list[o] = id;
  if (point_keys) point_keys[o] = tile_hi | dbits;
  const float4* g = grecords + (vbase + id) * 3;
  const float4 a = __ldg(g), b = __ldg(g + 1);
  float4 c = __ldg(g + 2);
  const uint32_t cm = cell_mask(a.x, a.y, c.z, c.w, tx0, ty0);  // (ex, ey) -> per-instance cell mask
  c.z = __uint_as_float(cm);
  c.w = 0.f;
  float4* r = records + (size_t)o * 3;
  r[0] = a;
  r[1] = b;
  r[2] = c;
  return cm;
}

// Per-block survivor lists of one tile (second pass of the epilogue): for each of the tile's eight 8x4-pixel blocks, the
// positions (inside the tile's sorted range) of the instances whose cell mask touches the block, in sorted order.  The
// blend kernel's warps then stream exactly their own survivors: no cull, no shared ring, no coupling between the warps
// of a tile.  Warp b compacts block b: one walk over the tile's masks (shared memory in the bucket path), one ballot per
// 32 positions, the running offset in a register - no counting pass and no cross-warp scan.
template <bool kSmemMasks>
__device__ void build_block_lists(uint32_t n, uint32_t rgx, const float4* __restrict__ records,
                                  uint32_t* __restrict__ blists, uint32_t* __restrict__ bcnt, uint16_t* s_mask) {
  static_assert(SORT_WARPS == 8, "one warp per 8x4 block");
  __syncthreads();  // every write_sorted store of this CTA (records / shared-memory masks) is visible to the CTA
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int sh = 4 * (w >> 1) + 2 * (w & 1);  // the block's two cells are adjacent mask bits
  const float* mask_f = reinterpret_cast<const float*>(records + (size_t)rgx * 3) + 10;  // r2.z of record 0
  uint32_t* out = blists + (size_t)rgx * 8 + (size_t)w * n;
  uint32_t off = 0;
  // kSmemMasks: the bucket path left all n <= S3R_SORT_SMEM_CAP masks in shared memory.  Otherwise (LSD / global sort
  // paths, whose key buffers are free by now) the masks are staged from the records S3R_SORT_SMEM_CAP at a time, all
  // loads of a thread in flight together - a walk with a dependent global load per step would cost an L2 round trip
  // per 32 positions and warp.
  for (uint32_t c0 = 0; c0 < n; c0 += S3R_SORT_SMEM_CAP) {
    const uint32_t cn = min((uint32_t)S3R_SORT_SMEM_CAP, n - c0);
    if (!kSmemMasks) {
      if (c0) __syncthreads();  // the previous window has been consumed
      for (uint32_t i = tid; i < cn; i += SORT_THREADS)
        s_mask[i] = (uint16_t)__float_as_uint(__ldcg(mask_f + (size_t)(c0 + i) * S3R_REC_FLOATS));
      __syncthreads();
    }
#pragma unroll 4
    for (uint32_t p = 0; p < cn; p += 32) {
      const uint32_t q = p + lane;
      const uint32_t cm = q < cn ? (uint32_t)s_mask[q] : 0u;
      const bool hit = ((cm >> sh) & 3u) != 0u;
      const uint32_t m = __ballot_sync(0xffffffffu, hit);
      if (hit) out[off + __popc(m & lt)] = c0 + q;
      off += __popc(m);
    }
  }
  if (lane == 0) bcnt[w] = off;
}

// The blend stage's work queue: (view, tile) indices by descending instance count (256-bucket counting sort; the order
// inside a bucket is arbitrary - it only affects scheduling, never results).  Runs as ONE extra CTA of the tile-sort grid
// (block (tiles, 0)), i.e. concurrently with the sorts and off the critical path of the chain.
//
// Queue position: the blend kernel is a persistent grid of `slots` = sms x CTAs-per-SM CTAs whose FIRST unit is the queue
// entry at their own blockIdx (later units come from an atomic counter).  On an idle GPU the block scheduler places
// blocks b and b + sms on the same SM, so the first `slots` ranks are laid out boustrophedon: wave 0 in descending
// weight, wave 1 in ascending weight, ... - the SM that got the heaviest tile gets the lightest one next.  (Placement
// only changes which CTA renders which tile, never a result; under concurrency with other streams it is merely
// another valid order.)  With one 256-tile view on 148 SMs the busiest SM's load drops from n[0] + n[148] to
// n[0] + n[255].
__device__ __forceinline__ uint32_t queue_position(uint32_t rank, uint32_t n, uint32_t sms, uint32_t slots) {
  if (sms == 0u || rank >= slots) return rank;
  const uint32_t wave = rank / sms, s = rank - wave * sms;
  if ((wave & 1u) == 0u) return rank;
  const uint32_t width = min(sms, min(n, slots) - wave * sms);  // the last wave may be partial
  return wave * sms + (width - 1u - s);
}

__device__ void build_work_order(int n, const uint32_t* __restrict__ tile_count, const long long* __restrict__ status,
                                 uint32_t* __restrict__ work_order, uint32_t* s_bucket /* [256] */, uint32_t* s_w /* [8] */,
                                 uint32_t sms, uint32_t slots) {
  static_assert(SORT_THREADS == 256, "one bucket per thread");
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  s_bucket[tid] = 0u;
  __syncthreads();
  const uint32_t width = (uint32_t)status[2] / 255u + 1u;  // status[2] = largest tile count (bin_scan)
  for (int i = tid; i < n; i += SORT_THREADS) atomicAdd(&s_bucket[255u - min(255u, tile_count[i] / width)], 1u);
  __syncthreads();
  const uint32_t c = s_bucket[tid];
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  uint32_t off = incl - c;
  for (int k = 0; k < w; k++) off += s_w[k];
  s_bucket[tid] = off;
  __syncthreads();
  for (int i = tid; i < n; i += SORT_THREADS)
    work_order[queue_position(atomicAdd(&s_bucket[255u - min(255u, tile_count[i] / width)], 1u), (uint32_t)n, sms, slots)] =
        (uint32_t)i;
}

// grid (tiles + 1, n_views): block (tiles, 0) builds the blend work queue, blocks (tiles, v > 0) exit
__global__ void __launch_bounds__(SORT_THREADS) s3r_tile_sort_kernel(
    int P, int tiles, int tiles_x, const uint2* __restrict__ ranges, unsigned long long* __restrict__ keys_a,
    unsigned long long* __restrict__ keys_b, const float4* __restrict__ grecords, uint32_t* __restrict__ point_list,
    unsigned long long* __restrict__ point_keys, float4* __restrict__ records, const uint32_t* __restrict__ tile_count,
    const long long* __restrict__ status, uint32_t* __restrict__ work_order, uint32_t blend_sms, uint32_t blend_slots,
    uint32_t* __restrict__ blists, uint32_t* __restrict__ bcounts) {
  extern __shared__ unsigned long long s_keys[];  // [2][S3R_SORT_SMEM_CAP]; bucket path: [0] = grouped keys, [1] = bucket table
  __shared__ uint32_t s_hist[SORT_WARPS][256];
  __shared__ uint32_t s_scan[SORT_WARPS];
  __shared__ int s_flag;
  __shared__ uint32_t s_min, s_max, s_maxb;
  const int view = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint32_t* s_start = reinterpret_cast<uint32_t*>(s_keys + S3R_SORT_SMEM_CAP);  // [SORT_NB + 1]
  for (int i = tid; i <= SORT_NB; i += SORT_THREADS) s_start[i] = 0u;
  if (tid == 0) {
    s_min = 0xffffffffu;
    s_max = 0u;
    s_maxb = 0u;
  }
  s3r_grid_dependency_sync();
  if (tile == tiles) {
    if (view == 0) build_work_order((int)gridDim.y * tiles, tile_count, status, work_order, &s_hist[0][0], s_scan, blend_sms,
                                    blend_slots);
    return;
  }
  const uint2 rg = ranges[(size_t)view * tiles + tile];
  const uint32_t n = rg.y - rg.x;
  uint32_t* bcnt = bcounts + ((size_t)view * tiles + tile) * 8;
  // bucket path: cell masks in sorted order, in the free tail of the second key buffer (behind the bucket table)
  uint16_t* s_mask = reinterpret_cast<uint16_t*>(s_start + SORT_NB + 8);
  const uint32_t s_mask_addr = (uint32_t)__cvta_generic_to_shared(s_mask);
  static_assert((SORT_NB + 8) * 4 + S3R_SORT_SMEM_CAP * 2 <= S3R_SORT_SMEM_CAP * 8, "mask array must fit behind the bucket table");
  if (n == 0) {
    if (tid < 8) bcnt[tid] = 0u;
    return;
  }
  unsigned long long* ga = keys_a + rg.x;
  unsigned long long* gb = keys_b + rg.x;
  const size_t vbase = (size_t)view * P;
  const unsigned long long tile_hi = ((unsigned long long)((uint32_t)view * (uint32_t)tiles + (uint32_t)tile)) << 32;
  const float tx0 = (float)((tile % tiles_x) * S3R_TILE), ty0 = (float)((tile / tiles_x) * S3R_TILE);

  if (n <= S3R_SORT_SMEM_CAP) {
    // ================= bucket path
    unsigned long long k[SORT_KPT];
    uint32_t slot[SORT_KPT];
    uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
    for (int j = 0; j < SORT_KPT; j++) {
      const uint32_t i = tid + j * SORT_THREADS;
      k[j] = 0ull;
      if (i < n) {
        k[j] = __ldcs(ga + i);
        const uint32_t hi = (uint32_t)(k[j] >> 32);
        mn = min(mn, hi);
        mx = max(mx, hi);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __syncthreads();  // s_start zeroed, s_min / s_max initialised
    if (lane == 0) {
      atomicMin(&s_min, mn);
      atomicMax(&s_max, mx);
    }
    __syncthreads();
    const uint32_t lo = s_min, span = s_max - lo;
    // digit = (depth_bits - lo) >> shift < SORT_NB, monotone in depth_bits
    const int shift = max(0, (32 - __clz(span)) - 10);
    static_assert(SORT_NB == 1024, "shift assumes 10 digit bits");
#pragma unroll
    for (int j = 0; j < SORT_KPT; j++) {
      const uint32_t i = tid + j * SORT_THREADS;
      slot[j] = 0u;
      if (i < n) slot[j] = atomicAdd(&s_start[(((uint32_t)(k[j] >> 32)) - lo) >> shift], 1u);
    }
    __syncthreads();
    // exclusive scan of the SORT_NB counts (4 per thread) -> bucket starts, and the largest bucket
    {
      uint32_t c[SORT_NB / SORT_THREADS];
      uint32_t sum = 0, big = 0;
#pragma unroll
      for (int q = 0; q < SORT_NB / SORT_THREADS; q++) {
        c[q] = s_start[tid * (SORT_NB / SORT_THREADS) + q];
        sum += c[q];
        big = max(big, c[q]);
      }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) big = max(big, __shfl_xor_sync(0xffffffffu, big, o));
      if (lane == 31) s_scan[w] = incl;
      if (lane == 0) atomicMax(&s_maxb, big);
      __syncthreads();
      uint32_t run = incl - sum;
#pragma unroll
      for (int q = 0; q < SORT_WARPS; q++)
        if (q < w) run += s_scan[q];
#pragma unroll
      for (int q = 0; q < SORT_NB / SORT_THREADS; q++) {
        s_start[tid * (SORT_NB / SORT_THREADS) + q] = run;
        run += c[q];
      }
      if (tid == SORT_THREADS - 1) s_start[SORT_NB] = run;  // = n
    }
    __syncthreads();
    if (s_maxb <= SORT_MAX_BUCKET) {
      // scatter into bucket-grouped order (order inside a bucket is arbitrary: the ranks below fix it)
#pragma unroll
      for (int j = 0; j < SORT_KPT; j++) {
        const uint32_t i = tid + j * SORT_THREADS;
        if (i < n) s_keys[s_start[(((uint32_t)(k[j] >> 32)) - lo) >> shift] + slot[j]] = k[j];
      }
      __syncthreads();
      // exact rank inside the bucket on the full (depth_bits, gaussian) key, then straight to the outputs (a separate,
      // fully unrolled gather phase with 12 loads in flight per thread was measured slower: 15.3 vs 14.4 us)
#pragma unroll 2
      for (uint32_t p = tid; p < n; p += SORT_THREADS) {
        const unsigned long long key = s_keys[p];
        const uint32_t d = (((uint32_t)(key >> 32)) - lo) >> shift;
        const uint32_t b0 = s_start[d], b1 = s_start[d + 1];
        uint32_t rank = 0;
        for (uint32_t q = b0; q < b1; q++) rank += (s_keys[q] < key) ? 1u : 0u;
        const uint32_t cm =
            write_sorted(rg.x + b0 + rank, key, tile_hi, vbase, tx0, ty0, grecords, point_list, point_keys, records);
        // plain asm store: a C++ store into the same dynamic shared array would order itself against the next key's
        // s_keys / s_start reads (may-alias) and serialise the two gathers the unrolled loop keeps in flight
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(s_mask_addr + 2u * (b0 + rank)), "h"((unsigned short)cm));
      }
      build_block_lists<true>(n, rg.x, records, blists, bcnt, s_mask);
      return;
    }
    // ================= LSD path (rare): stable radix passes in shared memory, keys re-staged from registers
    __syncthreads();
    unsigned long long* cur = s_keys;
    unsigned long long* nxt = s_keys + S3R_SORT_SMEM_CAP;
#pragma unroll
    for (int j = 0; j < SORT_KPT; j++) {
      const uint32_t i = tid + j * SORT_THREADS;
      if (i < n) cur[i] = k[j];
    }
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < 4; pass++) {
      if (radix_pass<false>(cur, nxt, n, 32 + 8 * pass, s_hist, s_scan, &s_flag)) {
        unsigned long long* t = cur; cur = nxt; nxt = t;
      }
    }
    for (uint32_t i = tid; i < n; i += SORT_THREADS)
      write_sorted(rg.x + i, cur[i], tile_hi, vbase, tx0, ty0, grecords, point_list, point_keys, records);
    build_block_lists<false>(n, rg.x, records, blists, bcnt, reinterpret_cast<uint16_t*>(s_keys));
    return;
  }
  // ================= global path: tiles above the shared-memory capacity
  unsigned long long* cur = ga;
  unsigned long long* nxt = gb;
#pragma unroll 1
  for (int pass = 0; pass < 4; pass++) {
    if (radix_pass<true>(cur, nxt, n, 32 + 8 * pass, s_hist, s_scan, &s_flag)) {
      unsigned long long* t = cur; cur = nxt; nxt = t;
    }
    __threadfence_block();
  }
  for (uint32_t i = tid; i < n; i += SORT_THREADS)
    write_sorted(rg.x + i, __ldcg(cur + i), tile_hi, vbase, tx0, ty0, grecords, point_list, point_keys, records);
  build_block_lists<false>(n, rg.x, records, blists, bcnt, reinterpret_cast<uint16_t*>(s_keys));
}

int s3r_launch_sort(const s3r_raster_params& p, const s3r_raster_layout& L, char* state, cudaStream_t st) {
  const size_t smem = 2 * (size_t)S3R_SORT_SMEM_CAP * sizeof(unsigned long long);
  static_assert((size_t)S3R_SORT_SMEM_CAP * 8 >= (SORT_NB + 1) * 4, "bucket table lives in the second key buffer");
  static size_t configured[64] = {};
  int rc = s3r_ensure_dynamic_smem(s3r_tile_sort_kernel, smem, configured);
  if (rc != S3R_OK) return rc;
  dim3 grid(L.tiles + 1, p.n_views);
  int blend_sms = 0, blend_slots = 0;
  rc = s3r_blend_grid(&blend_sms, &blend_slots);
  if (rc != S3R_OK) return rc;
  S3R_CUDA_CHECK(s3r_launch_pdl(s3r_tile_sort_kernel, grid, dim3(SORT_THREADS), smem, st, (s3r_raster_pdl_mask() >> 3) & 1, p.P, L.tiles, L.tiles_x,
                                (const uint2*)(state + L.ranges), (unsigned long long*)(state + L.keys_unsorted),
                                (unsigned long long*)(state + L.keys_tmp), (const float4*)(state + L.grecords),
                                (uint32_t*)(state + L.point_list), (unsigned long long*)(state + L.point_keys),
                                (float4*)(state + L.records), (const uint32_t*)(state + L.tile_count),
                                (const long long*)(state + L.status), (uint32_t*)(state + L.work_order),
                                (uint32_t)blend_sms, (uint32_t)blend_slots, (uint32_t*)(state + L.blists),
                                (uint32_t*)(state + L.bcounts)));
  return S3R_OK;
}
