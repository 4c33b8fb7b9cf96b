// Rasterizer stage 5: per-tile front-to-back alpha blend (forward), TILE-GRANULAR kernel.  Since the end of round 2 the
// default is the warp-granular kernel of raster_blend_blocks.cu; this one is selected with S3R_TUNE_BLEND_KERNEL = 1
// (same results bit for bit) and also hosts the launcher / grid helpers of the blend stage.
//
// Persistent CTAs, one (view, 16x16 tile) unit at a time: 8 consumer warps + 1 producer warp.  Consumer warp w owns
// the 8x4 pixel block (w&1, w>>1) of the tile; each half-warp owns a 4x4 cell with its own survivor list.  The tile's
// sorted-gathered 48-byte records are streamed through a 4-stage shared-memory ring by TMA bulk copies issued by the
// producer warp (cp.async.bulk + mbarrier complete_tx; SASS: UBLKCP / SYNCS), 128 records per stage; full/empty
// mbarriers decouple the consumer warps from each other (no CTA-wide barrier in the main loop).  Per stage a warp
// tests the records' cell masks (bit per 4x4 cell of the tile, written by the tile sort) against its own cells,
// compacts the survivors with ballots into lists of 16-bit shared-memory addresses and composites only those,
// BLEND_U at a time with packed f32x2 arithmetic (skipping a record that cannot reach 1/255 is result-neutral).
//
// Semantics follow upstream renderCUDA + the "-w-pose" fork (blended depth, opacity, n_touched) as restated by
// oracle/raster_oracle.c:s3r_oracle_render.  The Gaussian exponent is evaluated
// in the log2 domain from the pre-scaled conic of the record (2 FMUL + 2 FFMA + FMUL, MUFU.EX2); the rare
// evaluations whose alpha lands within 1e-4 (relative) of the 1/255 threshold, or whose exponent is within
// 1e-5 of zero, are decided again from the exact per-Gaussian conic with the oracle's unfused fp32 operation
// order and a correctly rounded exp, so that every keep/skip decision equals the oracle's.
#include "s3r_common.cuh"

#define BLEND_CHUNK 128
#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_LO (ALPHA_MIN * (1.0f - S3R_ALPHA_BAND))
#define ALPHA_HI (ALPHA_MIN * (1.0f + S3R_ALPHA_BAND))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or `ns` have passed, so a
// waiting warp issues one instruction sequence per `ns` instead of spinning (round-2 profile: the producer's and the
// finished warps' polling loops were 16 % of all warp instructions of the kernel)
__device__ __forceinline__ bool mbar_try_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_hint(bar, parity, 4000u)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 on 64-bit register pairs)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// shared-memory loads by 32-bit shared-window address (the survivor lists hold such addresses; volatile: ordered with the
// mbarrier waits / arrivals, which are volatile asm as well)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64u(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// exact (oracle-identical) alpha for threshold-band evaluations
__device__ __noinline__ float exact_alpha(float power, float opacity) {
  const float e = (float)exp((double)power);
  return fminf(0.99f, __fmul_rn(opacity, e));
}

#ifndef BLEND_U
#define BLEND_U 4        // survivors evaluated per batch (independent alpha chains -> ILP); must stay 4 (two f32x2 pairs)
#endif
#ifndef BLEND_HALVES
// survivor lists per consumer warp: 2 = each half-warp owns a 4x4 pixel block with its own compacted list (the cull is
// twice as fine: a splat that covers only one half of the 8x4 block costs the other half nothing), 1 = one list per 8x4 block
#define BLEND_HALVES 2
#endif
#ifndef BLEND_STAGES
#define BLEND_STAGES 4   // depth of the TMA ring
#endif
#ifndef BLEND_MINB
#define BLEND_MINB 3
#endif
#ifndef BLEND_POLL_NS
#define BLEND_POLL_NS 400  // suspend-time hint of the producer's / finished warps' waits (they also watch `done_warps`)
#endif
#ifndef BLEND_SPLIT
#define BLEND_SPLIT 1     // CTAs per 16x16 tile (1: 8 consumer warps, 2: half tiles of 16x8 px with 4 consumer warps)
#endif
#ifndef BLEND_CTAS_PER_SM
#define BLEND_CTAS_PER_SM BLEND_MINB  // persistent CTAs per SM (fewer resident CTAs than work units => dynamic balancing)
#endif
#define BLEND_CWARPS (8 / BLEND_SPLIT)  // consumer warps (one 8x4 pixel block each); the last warp is the TMA producer
#define BLEND_THREADS (32 * (BLEND_CWARPS + 1))

#define BLEND_STAGE_BYTES ((BLEND_CHUNK + 1) * S3R_REC_BYTES)
#define BLEND_LIST_LEN (BLEND_CHUNK + 2 * BLEND_U + 8)
struct __align__(128) BlendSmem {
  float4 rec[BLEND_STAGES][(BLEND_CHUNK + 1) * 3];  // + one dummy record per stage (never hits): pads survivor batches
  // per-(half-)warp compacted survivors, stored as the record's 16-bit shared-memory address (the kernel is never
  // launched in a cluster, so the CTA's shared window starts at rank 0 and the whole struct lies below 2^16; checked at
  // kernel start): the compositing loop turns a list entry into an LDS address with one instruction
  uint16_t list[BLEND_CWARPS][BLEND_HALVES][BLEND_LIST_LEN];
  uint64_t full[BLEND_STAGES];                   // producer -> consumers (expect_tx / complete_tx)
  uint64_t empty[BLEND_STAGES];                  // consumers -> producer (one arrival per consumer warp)
  int done_warps;
  uint32_t unit;  // work unit fetched by thread 0 for the whole CTA
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Warp-specialised: warp 8 streams the tile's records through a BLEND_STAGES-deep ring with TMA bulk copies;
// the 8 consumer warps run decoupled from each other (no CTA-wide barrier in the main loop): each culls the
// stage against its own 8x4 pixel block, composites the survivors front to back, and releases the stage.
__device__ __forceinline__ float fast_exp2(float x) {  // MUFU.EX2 (flush-to-zero): |rel err| ~2^-22
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// oracle-identical decision for one (pixel, Gaussian): unfused fp32 power from the exact conic, correctly rounded exp
__device__ __noinline__ bool exact_decide(const float4 co, float dx, float dy, float* alpha_out) {
  const float q = __fadd_rn(__fmul_rn(__fmul_rn(co.x, dx), dx), __fmul_rn(__fmul_rn(co.z, dy), dy));
  const float power = __fsub_rn(__fmul_rn(-0.5f, q), __fmul_rn(__fmul_rn(co.y, dx), dy));
  if (power > 0.0f) return false;
  const float a = exact_alpha(power, co.w);
  *alpha_out = a;
  return a >= ALPHA_MIN;
}

template <bool kHasNT>
__global__ void __launch_bounds__(BLEND_THREADS, BLEND_MINB) s3r_blend_fwd_kernel(
    int W, int H, int P, int tiles_x, int tiles, int only_tile, uint32_t n_units, const uint32_t* __restrict__ work_order,
    unsigned* __restrict__ counters, const uint2* __restrict__ ranges, const float4* __restrict__ records,
    const uint32_t* __restrict__ point_list, const float4* __restrict__ conic_opacity,
    const float* __restrict__ background, float* __restrict__ out_color,
    float* __restrict__ out_depth, float* __restrict__ out_opacity, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, int32_t* __restrict__ n_touched) {
  __shared__ BlendSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

  if (tid < BLEND_STAGES) {  // dummy record: so far away that log2 G = -1.8e19 -> alpha = 0 (no range predicates in the loop)
    sm.rec[tid][BLEND_CHUNK * 3] = make_float4(3e9f, 3e9f, 0.0f, -1.0f);
    sm.rec[tid][BLEND_CHUNK * 3 + 1] = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
    sm.rec[tid][BLEND_CHUNK * 3 + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  if (smem_u32(&sm) + (uint32_t)sizeof(BlendSmem) > 0x10000u) __trap();  // survivor lists hold 16-bit shared addresses
  s3r_grid_dependency_sync();  // the prologue above is shared-memory only

  // Persistent CTA: work units (view, tile, part) are fetched from a device-side queue in the order bin_scan wrote to
  // `work_order` - heaviest tiles first (longest-processing-time-first list scheduling), so the SMs finish together even
  // though tiles differ 7x in instance count.
  for (bool first_unit = true;; first_unit = false) {
  if (tid == 0) {
    if (!first_unit) {
#pragma unroll
      for (int s = 0; s < BLEND_STAGES; s++) {
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sm.full[s])) : "memory");
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sm.empty[s])) : "memory");
      }
    }
#pragma unroll
    for (int s = 0; s < BLEND_STAGES; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], BLEND_CWARPS);
    }
    sm.done_warps = 0;
    // first unit: the queue entry at this CTA's own index (the tile sort laid the first gridDim.x entries out for the
    // block scheduler's placement, raster_sort.cu:queue_position); afterwards the shared counter
    sm.unit = first_unit ? blockIdx.x : gridDim.x + atomicAdd(&counters[1], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t unit = sm.unit;
  if (unit >= n_units) break;
  const uint32_t vt = only_tile >= 0 ? (uint32_t)only_tile : work_order[unit / BLEND_SPLIT];
  const int part = unit % BLEND_SPLIT;
  const int view = vt / tiles, tile = vt % tiles;
  const uint2 rg = ranges[vt];
  const uint32_t n = rg.y - rg.x;
  const uint32_t nchunks = (n + BLEND_CHUNK - 1) / BLEND_CHUNK;
  const float4* src = records + (size_t)rg.x * 3;

  if (w == BLEND_CWARPS) {
    // ===== TMA producer (one elected lane)
    if (lane == 0) {
      uint32_t issued = 0;
      for (uint32_t c = 0; c < nchunks; c++) {
        const int s = c % BLEND_STAGES;
        if (c >= BLEND_STAGES) {
          const uint32_t par = ((c / BLEND_STAGES) - 1) & 1;
          while (!mbar_try_hint(&sm.empty[s], par, BLEND_POLL_NS)) {  // parked between polls: no stolen issue slots
            if (*(volatile int*)&sm.done_warps == BLEND_CWARPS) break;
          }
        }
        if (*(volatile int*)&sm.done_warps == BLEND_CWARPS) break;
        const uint32_t cnt = min((uint32_t)BLEND_CHUNK, n - c * BLEND_CHUNK);
        mbar_expect_tx(&sm.full[s], cnt * S3R_REC_BYTES);
        bulk_g2s(sm.rec[s], src + (size_t)c * BLEND_CHUNK * 3, cnt * S3R_REC_BYTES, &sm.full[s]);
        issued = c + 1;
      }
      // drain: every issued copy must have landed before the CTA may retire
      const uint32_t first = issued > BLEND_STAGES ? issued - BLEND_STAGES : 0;
      for (uint32_t c = first; c < issued; c++) mbar_wait(&sm.full[c % BLEND_STAGES], (c / BLEND_STAGES) & 1);
    }
  } else {
  // ===== consumers
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int wt = part * BLEND_CWARPS + w;                                            // warp index inside the tile
  const int X0 = tx * S3R_TILE + (wt & 1) * 8, Y0 = ty * S3R_TILE + (wt >> 1) * 4;  // this warp's 8x4 block
#if BLEND_HALVES == 2
  const int half = lane >> 4;  // half-warp h owns the 4x4 block at x offset 4h
  const int px = X0 + 4 * half + (lane & 3), py = Y0 + ((lane >> 2) & 3);
#else
  const int half = 0;
  const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
#endif
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const uint32_t cellbit = 1u << (4 * (wt >> 1) + 2 * (wt & 1));  // left 4x4 cell of the block (the right one is the next bit)
  const uint32_t lt = (1u << lane) - 1u;
  uint16_t* list = sm.list[w][half];
  const uint32_t list_s = smem_u32(list);

  // A pixel that has terminated (test_T < 1e-4), or lies outside the image, is "dead": its lane keeps running in
  // lockstep with a poisoned pixel coordinate (+inf), so every later alpha evaluates to exactly 0 and its T, colour and
  // n_contrib stay frozen without any `done` predicate in the loop.
  // Blackwell packed fp32 (add / mul / fma .f32x2: one issue slot for two lanes of a register pair, each half rounded
  // like the scalar instruction): (dx, dy), the two (1 - alpha) / weight pairs of a batch and the (Cr, Cg) / (Cb, D)
  // accumulators - the record keeps (r, g) and (b, depth) in aligned register pairs.
  float T = 1.0f;
  uint64_t Crg = pack2(0.f, 0.f), Cbd = pack2(0.f, 0.f);
  uint64_t negpix = pack2(inside ? -pxf : -__int_as_float(0x7f800000), -pyf);  // (-x, -y) used for evaluation
  bool alive = inside;
  uint32_t last = 0;
  bool warp_done = !__any_sync(0xffffffffu, inside);
  if (warp_done && lane == 0) atomicAdd(&sm.done_warps, 1);

  for (uint32_t c = 0; c < nchunks; c++) {
    const int s = c % BLEND_STAGES;
    if (warp_done) {
      // drain mode: keep releasing stages in phase order so the producer never stalls on us
      if (*(volatile int*)&sm.done_warps == BLEND_CWARPS) break;
      if (c >= BLEND_STAGES) {
        const uint32_t par = ((c / BLEND_STAGES) - 1) & 1;
        bool all = false;
        while (!mbar_try_hint(&sm.empty[s], par, BLEND_POLL_NS)) {
          if (*(volatile int*)&sm.done_warps == BLEND_CWARPS) { all = true; break; }
        }
        if (all) break;
      }
      if (lane == 0) mbar_arrive(&sm.empty[s]);
      __syncwarp();
      continue;
    }
    const uint32_t cnt = min((uint32_t)BLEND_CHUNK, n - c * BLEND_CHUNK);
    const uint32_t stage_off = smem_u32(&sm.rec[s][0]);  // shared address of the stage's first record
    mbar_wait(&sm.full[s], (c / BLEND_STAGES) & 1);
    // ---- cull this stage against the warp's pixel block(s) and compact the survivors: one bit test per record on
    // the cell mask the sort epilogue stored in the record (bit 4*cy + cx over the tile's 4x4-pixel cells)
    int count = 0;  // survivors of this lane's list
#if BLEND_HALVES == 2
    int count_other = 0;
#endif
#pragma unroll
    for (int j = 0; j < BLEND_CHUNK / 32; j++) {
      const int i = j * 32 + lane;
      const uint16_t off = (uint16_t)(stage_off + (uint32_t)i * S3R_REC_BYTES);
      const uint32_t cm = (uint32_t)i < cnt ? __float_as_uint(sm.rec[s][i * 3 + 2].z) : 0u;
#if BLEND_HALVES == 2
      const bool hit0 = (cm & cellbit) != 0u, hit1 = (cm & (cellbit << 1)) != 0u;
      const uint32_t m0 = __ballot_sync(0xffffffffu, hit0);
      const uint32_t m1 = __ballot_sync(0xffffffffu, hit1);
      const int c0 = half ? count_other : count, c1 = half ? count : count_other;
      if (hit0) sm.list[w][0][c0 + __popc(m0 & lt)] = off;
      if (hit1) sm.list[w][1][c1 + __popc(m1 & lt)] = off;
      const int n0 = __popc(m0), n1 = __popc(m1);
      count += half ? n1 : n0;
      count_other += half ? n0 : n1;
#else
      const bool hit = (cm & (cellbit * 3u)) != 0u;
      const uint32_t m = __ballot_sync(0xffffffffu, hit);
      if (hit) list[count + __popc(m & lt)] = off;
      count += __popc(m);
#endif
    }
    // pad with the stage's dummy record up to the end of the last batch (of the longer list)
    const uint16_t dummy_off = (uint16_t)(stage_off + BLEND_CHUNK * S3R_REC_BYTES);
#if BLEND_HALVES == 2
    const int nbatch = max(count, count_other);
    for (int p = count + (lane & 15); p < nbatch + BLEND_U; p += 16) list[p] = dummy_off;
#else
    const int nbatch = count;
    if (lane < BLEND_U) list[count + lane] = dummy_off;
#endif
    __syncwarp();
    const uint32_t base_idx = c * BLEND_CHUNK;
    uint32_t lastoff = 0xffffffffu;  // list entry (record offset) of the last composited splat of this chunk
    uint2 packed_next = lds64u(list_s);  // survivor addresses of the next batch, fetched one batch ahead
    uint32_t lp = list_s;
#pragma unroll 1
    for (int k = 0; k < nbatch; k += BLEND_U) {
      const uint2 packed = packed_next;
      lp += 2 * BLEND_U;
      packed_next = lds64u(lp);
      const uint32_t off[BLEND_U] = {packed.x & 0xffffu, packed.x >> 16, packed.y & 0xffffu, packed.y >> 16};
      float alpha[BLEND_U];  // 0 unless the splat is kept
      uint64_t crg[BLEND_U], cbd[BLEND_U];
      float t[BLEND_U + 1];
      const uint32_t lastoff_in = lastoff;
      bool near_thr = false;
      // ---- BLEND_U independent alpha chains (branch-free); entries past the list's end are the dummy record
#pragma unroll
      for (int u = 0; u < BLEND_U; u++) {
        const float4 r0 = lds128(off[u]);
        const float4 r1 = lds128(off[u] + 16);
        const float2 r2 = lds64(off[u] + 32);
        float dx, dy;
        unpack2(add2(pack2(r0.x, r0.y), negpix), dx, dy);
        // log2 G = dx*(A'*dx + B'*dy) + (C'*dy)*dy on the pre-scaled conic: FMUL2 + 2 FFMA + FMUL, then MUFU.EX2
        float bdy, cdy;
        unpack2(mul2(pack2(r0.z, r0.w), pack2(dy, dy)), bdy, cdy);
        const float l2g = fmaf(dx, fmaf(r1.x, dx, bdy), cdy * dy);
        const float a = fminf(0.99f, r1.y * fast_exp2(l2g));
        const bool keep = a >= ALPHA_HI;
        // guard bands: alpha within 1e-4 (relative) below 1/255, or an exponent that is positive / so close to 0 that
        // its sign is in doubt (the oracle skips `power > 0`): those evaluations are decided exactly below
        near_thr = near_thr || ((a >= ALPHA_LO) && !keep) || (l2g > -S3R_PZERO_BAND);
        alpha[u] = keep ? a : 0.0f;
        lastoff = keep ? off[u] : lastoff;
        crg[u] = pack2(r1.z, r1.w);
        cbd[u] = pack2(r2.x, r2.y);
      }
      // ---- transmittance chain of the batch; T only decreases, so the batch contains a termination iff t[U] < 1e-4
      const uint64_t a01 = pack2(alpha[0], alpha[1]), a23 = pack2(alpha[2], alpha[3]);
      float om[BLEND_U];  // 1 - alpha: fma(alpha, -1, 1) rounds once, like the subtraction
      unpack2(fma2(a01, pack2(-1.0f, -1.0f), pack2(1.0f, 1.0f)), om[0], om[1]);
      unpack2(fma2(a23, pack2(-1.0f, -1.0f), pack2(1.0f, 1.0f)), om[2], om[3]);
      t[0] = T;
#pragma unroll
      for (int u = 0; u < BLEND_U; u++) t[u + 1] = t[u] * om[u];
      // common case: front-to-back compositing of the whole batch without a single predicate (instantiated twice, so
      // that the rare path below does not constrain the register allocation of the hot one)
      auto composite_all = [&](uint64_t p01, uint64_t p23) {
        float wgt[BLEND_U];
        unpack2(mul2(p01, pack2(t[0], t[1])), wgt[0], wgt[1]);
        unpack2(mul2(p23, pack2(t[2], t[3])), wgt[2], wgt[3]);
#pragma unroll
        for (int u = 0; u < BLEND_U; u++) {
          const uint64_t w2 = pack2(wgt[u], wgt[u]);
          Crg = fma2(crg[u], w2, Crg);
          Cbd = fma2(cbd[u], w2, Cbd);
          if (kHasNT) {
            if (alpha[u] > 0.0f && t[u + 1] > 0.5f) {
              const uint32_t i = (off[u] - stage_off) / S3R_REC_BYTES;
              atomicAdd(&n_touched[(size_t)view * P + point_list[(size_t)rg.x + base_idx + i]], 1);
            }
          }
        }
        T = t[BLEND_U];
      };
      // one vote covers both rare events of a batch: an evaluation inside a guard band, or a pixel that terminates
      if (!__any_sync(0xffffffffu, near_thr || t[BLEND_U] < 0.0001f)) {
        composite_all(a01, a23);
      } else {
        // ---- some evaluation landed in a guard band -> decide it like the oracle, from the exact conic
        if (__any_sync(0xffffffffu, near_thr)) {
          lastoff = lastoff_in;
#pragma unroll
          for (int u = 0; u < BLEND_U; u++) {
            if (k + u < count) {
              const float4 r0 = lds128(off[u]);
              const float4 r1 = lds128(off[u] + 16);
              float dx, dy;
              unpack2(add2(pack2(r0.x, r0.y), negpix), dx, dy);
              const float l2g = fmaf(dx, fmaf(r1.x, dx, r0.z * dy), (r0.w * dy) * dy);
              const float a = fminf(0.99f, r1.y * fast_exp2(l2g));
              if ((a >= ALPHA_LO && a < ALPHA_HI) || l2g > -S3R_PZERO_BAND) {
                const uint32_t i = (off[u] - stage_off) / S3R_REC_BYTES;
                const uint32_t gid = point_list[(size_t)rg.x + base_idx + i];
                float ax = a;
                const bool kx = exact_decide(conic_opacity[(size_t)view * P + gid], dx, dy, &ax);
                alpha[u] = (kx && alive) ? ax : 0.0f;
              }
            }
            lastoff = alpha[u] > 0.0f ? off[u] : lastoff;
          }
#pragma unroll
          for (int u = 0; u < BLEND_U; u++) t[u + 1] = t[u] * (1.0f - alpha[u]);
        }
        if (!__any_sync(0xffffffffu, t[BLEND_U] < 0.0001f)) {
          composite_all(pack2(alpha[0], alpha[1]), pack2(alpha[2], alpha[3]));
        } else {
          // some pixel of the warp terminates inside this batch (at most once per pixel): predicated version
          lastoff = lastoff_in;
#pragma unroll
          for (int u = 0; u < BLEND_U; u++) {
            const bool stop = t[u + 1] < 0.0001f;  // stays true for the rest of the batch
            const bool acc = !stop && alpha[u] > 0.0f;
            const float wgt = stop ? 0.0f : alpha[u] * t[u];
            const uint64_t w2 = pack2(wgt, wgt);
            Crg = fma2(crg[u], w2, Crg);
            Cbd = fma2(cbd[u], w2, Cbd);
            if (kHasNT) {
              if (acc && t[u + 1] > 0.5f) {
                const uint32_t i = (off[u] - stage_off) / S3R_REC_BYTES;
                atomicAdd(&n_touched[(size_t)view * P + point_list[(size_t)rg.x + base_idx + i]], 1);
              }
            }
            T = stop ? T : t[u + 1];
            lastoff = acc ? off[u] : lastoff;
          }
          if (t[BLEND_U] < 0.0001f) {
            alive = false;
            negpix = pack2(-__int_as_float(0x7f800000), -pyf);
          }
          if (!__any_sync(0xffffffffu, alive)) {
            warp_done = true;
            break;
          }
        }
      }
    }
    // n_contrib = 1-based index of the last composited splat
    if (lastoff != 0xffffffffu) last = base_idx + (lastoff - stage_off) / S3R_REC_BYTES + 1u;
    __syncwarp();
    if (lane == 0) {
      if (warp_done) atomicAdd(&sm.done_warps, 1);
      mbar_arrive(&sm.empty[s]);
    }
  }
  if (inside) {
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    const float* bg = background + view * 3;
    float* oc = out_color + (size_t)view * 3 * HW;
    float Cr, Cg, Cb, D;
    unpack2(Crg, Cr, Cg);
    unpack2(Cbd, Cb, D);
    oc[pix] = Cr + T * bg[0];
    oc[HW + pix] = Cg + T * bg[1];
    oc[2 * HW + pix] = Cb + T * bg[2];
    out_depth[(size_t)view * HW + pix] = D;
    out_opacity[(size_t)view * HW + pix] = 1.0f - T;
    final_T[(size_t)view * HW + pix] = T;
    n_contrib[(size_t)view * HW + pix] = last;
  }
  }  // consumers
  __syncthreads();  // the unit is finished: every bulk copy has landed, nobody touches the ring or the barriers any more
  }  // unit loop
  // the last CTA to leave rewinds the queue for the next launch on this state buffer
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&counters[2], 1u) == gridDim.x - 1) {
      counters[1] = 0u;
      counters[2] = 0u;
      counters[7] = 2u;  // this state was rendered by the tile-granular kernel
    }
  }
}

int& s3r_raster_pdl_mask() {
  // measured on B200 (scripts/raster_probe.py): PDL shortens the chain only in front of preprocess and bin_emit
  // (-1.4 us); in front of bin_scan / tile_sort / blend it costs 4-16 us
  static int v = 5;
  return v;
}

int& s3r_blend_only_tile() {
  static int v = 0;
  return v;
}

#ifndef BLEND_KERNEL_DEFAULT
#define BLEND_KERNEL_DEFAULT 0
#endif
int& s3r_blend_kernel_choice() {
  static int v = BLEND_KERNEL_DEFAULT;
  return v;
}

int s3r_blend_grid(int* sms_out, int* slots_out) {
  static int slots[64] = {}, sms[64] = {};
  int dev = 0;
  S3R_CUDA_CHECK(cudaGetDevice(&dev));
  dev &= 63;
  if (slots[dev] == 0) {
    int n = 0, per_sm = 0;
    S3R_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    S3R_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s3r_blend_fwd_kernel<false>, BLEND_THREADS, 0));
    if (per_sm > BLEND_CTAS_PER_SM) per_sm = BLEND_CTAS_PER_SM;
    sms[dev] = n;
    slots[dev] = n * (per_sm > 0 ? per_sm : 1);
  }
  *sms_out = sms[dev];
  *slots_out = slots[dev];
  return S3R_OK;
}

int s3r_launch_blend(const s3r_raster_params& p, const s3r_raster_outputs& o, const s3r_raster_layout& L,
                     char* state, cudaStream_t st) {
  if (s3r_blend_kernel_choice() == 0) {  // warp-granular kernel over the per-block survivor lists (raster_blend_blocks.cu)
    const int ot = s3r_blend_only_tile();
    return s3r_launch_blend_blocks(p, o, L, state, st, ot > 0 && ot <= L.tiles ? ot - 1 : -1);
  }
  auto kern = o.n_touched ? s3r_blend_fwd_kernel<true> : s3r_blend_fwd_kernel<false>;
  // persistent grid: as many CTAs as the device holds at once (per device, cached), never more than there are units
  int n_sms = 0, n_slots = 0;
  int rc = s3r_blend_grid(&n_sms, &n_slots);
  if (rc != S3R_OK) return rc;
  uint32_t n_units = (uint32_t)L.tiles * (uint32_t)p.n_views * BLEND_SPLIT;
  int only_tile = -1;
  if (s3r_blend_only_tile() > 0 && s3r_blend_only_tile() <= L.tiles) {  // development probe: one tile of view 0
    only_tile = s3r_blend_only_tile() - 1;
    n_units = BLEND_SPLIT;
  }
  dim3 grid(n_units < (uint32_t)n_slots ? n_units : (uint32_t)n_slots);
  S3R_CUDA_CHECK(s3r_launch_pdl(kern, grid, dim3(BLEND_THREADS), 0, st, (s3r_raster_pdl_mask() >> 4) & 1, p.width, p.height,
                                p.P, L.tiles_x, L.tiles, only_tile, n_units, (const uint32_t*)(state + L.work_order),
                                (unsigned*)(state + L.counters), (const uint2*)(state + L.ranges),
                                (const float4*)(state + L.records), (const uint32_t*)(state + L.point_list),
                                (const float4*)(state + L.conic_opacity), p.background, o.color, o.depth, o.opacity,
                                (float*)(state + L.final_T), (uint32_t*)(state + L.n_contrib), o.n_touched));
  return S3R_OK;
}
