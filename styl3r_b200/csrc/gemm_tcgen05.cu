// bf16 GEMM with fused epilogue on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   C[M,N] = act( A[M,K] . W[N,K]^T + bias[N] ) + residual[M,N]        (bf16 operands, fp32 accumulation in TMEM)
//
// This is the dense-contraction kernel behind the encoder's nn.Linear layers (qkv / proj / fc1(+GELU) / fc2 /
// projq,k,v / decoder_embed: src/model/encoder/backbone/croco/blocks.py:70-73,91-93,162-166), where the reference
// calls cuBLAS through torch (TF32).  Both operands are K-major (row-major activations, row-major [out,in] weights),
// i.e. the "TN" GEMM.
//
// Structure (one 128 x BN output tile per CTA, 6 warps, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d of a 128x64 A tile and a BNx64 W tile per k-block into a deep (6-10 stage)
//            SWIZZLE_128B shared-memory ring, completion on `full` mbarriers (expect_tx)
//   warp 1   allocates BN TMEM columns; one elected lane issues 4 x tcgen05.mma (M=128, N=BN, K=16) per k-block from
//            shared-memory matrix descriptors, releases ring slots with tcgen05.commit -> `empty` mbarriers and signals
//            the epilogue with a final commit
//   warps 2-5 epilogue phase 1: tcgen05.ld (32 lanes x 32 columns per instruction) TMEM -> registers -> fp32 staging tile
//   all 8 warps (the TMA / MMA warps have finished by then, warps 6-7 exist for this) epilogue phase 2: + bias, exact
//            GELU, + residual, RoPE -> bf16 (or fp32), 16-byte global stores.  The epilogue math is latency-bound per
//            warp (one warp per SM sub-partition issues ~0.25 IPC), so phase 2 is spread over twice the warps
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "s3r_common.cuh"

int& s3r_attn_onepass();  // attention_tcgen05.cu

#define GEMM_BM 128
#define GEMM_BK 64
#define GEMM_THREADS 256

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void g_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(g_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void g_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          g_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(g_smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void g_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          g_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(g_smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void g_tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          g_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(g_smem_u32(bar))
      : "memory");
}
// Multicast variants (thread-block clusters): one L2 read lands in the same shared-memory offset of every CTA in `mask`
// and completes bytes on each destination CTA's own mbarrier at that offset.
__device__ __forceinline__ void g_tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          g_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(g_smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void g_umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   g_smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// ---- CTA pairs (cta_group::2): two CTAs of a (2,1,1) cluster on the two SMs of a TPC run ONE tcgen05.mma of M = 256:
// each CTA holds its own 128 rows of A, HALF of the B tile (the tensor core reads the other half from the peer's shared
// memory) and 128 lanes of the accumulator in its own TMEM.  Both CTAs' TMA loads complete on the LEADER's (rank 0) full
// barrier; the leader's elected thread issues the MMAs and releases ring slots in both CTAs with a multicast commit.
__device__ __forceinline__ uint32_t g_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void g_tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          g_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void g_tma_load_4d_pair(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                                   uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          g_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void g_umma_pair(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void g_umma_commit_pair(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   g_smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void g_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t g_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 = 1 | [32,46) SBO>>4 = 1024>>4 | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t g_make_desc(const void* smem_ptr) {
  const uint32_t addr = g_smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand (the contraction index is the slow one in global memory: X^T, dY^T, W^T - the backward GEMMs) under
// SWIZZLE_128B: the tile is stored as [MN/64 atoms][64 contraction rows][64 elements = 128 B]; canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute::UMMA make_umma_desc<Major::MN>): LBO = byte offset between
// two 64-element MN atoms = 64 rows x 128 B, SBO = byte offset between two 8-row contraction groups = 1024 B.
__device__ __forceinline__ uint64_t g_make_desc_mn(const void* smem_ptr) {
  const uint32_t addr = g_smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), a/b major at bits 15/16 (0 = K-major,
// 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t g_make_idesc(int m, int n, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void g_umma(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void g_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(g_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void g_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// erf for the exact-GELU epilogues: Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16 output rounding)
// - 5 FMAs + MUFU.RCP + MUFU.EX2 instead of the ~25-instruction branchy erff
__device__ __forceinline__ float g_fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));  // MUFU.RCP
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));  // MUFU.EX2
  const float y = 1.0f - p * t * e;
  return copysignf(y, x);
}

template <int BN, int STAGES_, int PAIR = 0>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * GEMM_BK * 2;  // a CTA of a pair holds half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // ring depth: large grids use ~100 KB per CTA so that TWO CTAs are resident per SM — the epilogue of one tile (TMEM
  // -> registers, bias / erf-GELU / residual, stores) then overlaps the TMA + MMA main loop of the other (measured:
  // better than one CTA with a 200 KB ring, whose tensor pipe idles during its own epilogue).  Grids smaller than the
  // machine (the M = 257/514 shapes of cfg2) are latency-bound on the TMA round trip instead: they get an 8-stage
  // ring so that (almost) the whole K extent is in flight at once.
  static constexpr int STAGES = STAGES_;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 256 + 1024;  // + barriers + alignment slack
};

struct RopeArgs {            // RoPE-2D fused into the epilogue (qkv / projq / projk outputs: head_dim 64)
  const long long* pos;      // [M, 2] int64 (y, x) per output row
  const float2* table;       // [max_pos + 1, 16] (cos, sin) of pos * base^(-d/16)
  int cols;                  // output columns [0, cols) are rotated (q and k parts), the rest (v) is left alone
  int max_pos;
};

// Implicit-GEMM convolution (kConv): A is the NHWC activation tensor behind a 4-D tensor map {C, W, H, N}.  The M tile is
// 128 consecutive NHWC pixels = a (BW x BH x BNimg) patch, and k-block kb = (tap, 64-channel block) loads that patch
// shifted by the tap offset (kw - pad, kh - pad); the zero padding of the convolution is TMA's out-of-bounds zero fill
// (negative / past-the-edge coordinates), so there is no im2col buffer and no halo logic.  The (BW, BH, BNimg, 64ch)
// box lands in shared memory as 128 rows of 128 B under SWIZZLE_128B - exactly the K-major A tile of the plain GEMM.
struct ConvArgs {
  int H, W;      // spatial size of the input (= output: stride 1, "same" padding)
  int KW;        // kernel width (taps are enumerated kh-major)
  int cblocks;   // 64-channel blocks per tap (weights are stored [Cout][tap][cblocks*64], zero padded)
  int pad;
};

// Clusters (CM x CN CTAs = CM adjacent M tiles x CN adjacent N tiles): the CN CTAs that share an M tile each load 1/CN of
// the A tile and multicast it to the others, the CM CTAs that share an N tile do the same with the W tile, so the L2 ->
// SM traffic of a cluster is that of one (CM*128) x (CN*BN) tile while the grid keeps the parallelism of small tiles -
// the M = 257/514-row GEMMs of the batch-1 encoder are bound by exactly that traffic (every SM pulls ~45-64 B/cycle).
// Ring slots are released cluster-wide: a CTA's `empty` barrier collects one tcgen05.commit from every CTA it
// multicasts into (its cluster row and column, CM + CN - 1 CTAs).
// MAJ: bit 0 = A is MN-major (given as [K, M] row-major), bit 1 = B is MN-major (given as [K, N] row-major) - the
// dgrad (B = W as stored) and wgrad (A = dY, B = X as stored) GEMMs of the encoder's backward, no transposed copies.
// EPI: compile-time superset of the epilogue features this instantiation can execute (S3R_EPI_* bits; the runtime
// `flags` select inside it).  The epilogue is issue-bound and every compiled-in feature costs registers and branches even
// when its flag is off (measured: the three training-only features slowed the inference GEMMs by 15-25 %), so the hot
// inference shapes are instantiated per feature set and everything else uses the all-features instance.
#define S3R_EPI_ALL 0xfff
// KS > 1: cluster split-K - a cluster of KS CTAs (cluster dims (1, 1, KS), rank = blockIdx.z) shares one output tile,
// each runs 1/KS of the k-blocks into its own TMEM accumulator and parks it in its own shared-memory staging tile; rank 0
// then sums the KS tiles through distributed shared memory (ld.shared::cluster) inside its fused epilogue.  For the
// narrow-N / long-K GEMMs of the batch-1 encoder (fc2, proj: 18-80 tiles for 148 SMs, every SM bound by its ~50 B/cycle
// TMA ingress) this spreads the K stream over KS times as many SMs without the L2 workspace + ticket round trips of the
// global split-K.
// DIRECT: register epilogue straight from TMEM for the latency-bound batch-1 shapes (one tile per CTA, grid <= ~1 wave):
// warps 2-9 (the block has 320 threads) - two per TMEM lane quarter, column halves - fetch bias / residual / positions
// while the main loop runs, then each thread walks its output row in 32-column blocks (tcgen05.ld -> bias / GELU /
// residual / RoPE with the rotation pair in the same thread -> 64 contiguous bytes to global memory).  No staging tile, no
// CTA barrier, no exposed global latency after the accumulator is complete: the staged two-phase epilogue costs ~1.5 us
// of a 9 us launch at 514x3072x1024.
#define GEMM_DIRECT_THREADS 320
template <int BN, int STAGES_, bool kConv, int CM = 1, int CN = 1, int MAJ = 0, int EPI = S3R_EPI_ALL, int KS = 1, int PAIR = 0, int DIRECT = 0>
__global__ void __launch_bounds__(DIRECT ? GEMM_DIRECT_THREADS : GEMM_THREADS, (GemmSmem<BN, STAGES_, PAIR>::STAGES * GemmSmem<BN, STAGES_, PAIR>::STAGE_BYTES <= 100 * 1024) ? 2 : 1)
s3r_gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* __restrict__ residual, void* __restrict__ Cout,
                     int M, int N, int K, int ldc, int ldr, int flags, RopeArgs rope, int splits, float* __restrict__ ws,
                     unsigned* __restrict__ counters, ConvArgs conv, long long batch_stride_c) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B atoms need 1024-B alignment
  using S = GemmSmem<BN, STAGES_, PAIR>;
#define EF(X) (((EPI) & (X)) && (flags & (X)))
  constexpr int GEMM_STAGES = S::STAGES;
  uint64_t* full = (uint64_t*)(smem + GEMM_STAGES * S::STAGE_BYTES);
  uint64_t* empty = full + GEMM_STAGES;
  uint64_t* tmem_full = empty + GEMM_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
  uint32_t* s_ticket = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * BN;
  const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
  // split-K: CTA z owns k-blocks [kb0, kb0 + num_kb); partial tiles meet in an fp32 workspace and the last CTA to
  // arrive for an output tile sums them (fixed order => deterministic) and runs the fused epilogue
  // batched mode (batch_stride_c > 0, splits == 1): blockIdx.z is the batch index - 3-D tensor maps {inner, rows, batch},
  // C advances by batch_stride_c elements per batch (the per-head GEMMs of the attention backward)
  const bool batched = batch_stride_c > 0;
  const int zsplit = batched ? 0 : (int)blockIdx.z;
  const int kb0 = (int)((long long)zsplit * total_kb / splits);
  const int num_kb = (int)((long long)(zsplit + 1) * total_kb / splits) - kb0;
  if (batched)
    Cout = (flags & S3R_EPI_OUT_F32) ? (void*)((float*)Cout + (long long)blockIdx.z * batch_stride_c)
                                     : (void*)((__nv_bfloat16*)Cout + (long long)blockIdx.z * batch_stride_c);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < GEMM_STAGES; s++) {
      g_mbar_init(&full[s], 1);
      g_mbar_init(&empty[s], CM + CN - 1);
    }
    g_mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr bool kCluster = CM * CN > 1 || PAIR;
  static_assert(!PAIR || (CM * CN == 1 && KS == 1 && MAJ == 0), "CTA pairs: K-major operands, no multicast / split-K");
  const uint32_t prank = PAIR ? g_cluster_ctarank() : 0u;  // rank inside the CTA pair (0 = leader: issues the MMAs)
  static_assert(KS == 1 || (CM * CN == 1 && BN <= 128), "cluster split-K: single epilogue pass");
  static_assert(!DIRECT || (CM * CN == 1 && KS == 1 && !PAIR && MAJ == 0 && BN <= 128), "register epilogue: plain one-tile GEMM");
  uint32_t xr = 0, yr = 0;       // this CTA's position inside its cluster (x = M direction, fastest)
  uint16_t mask_a = 1, mask_b = 1, mask_rel = 1;
  if (CM * CN > 1) {
    const uint32_t crank = g_cluster_ctarank();
    xr = crank % CM, yr = crank / CM;
    mask_a = mask_b = 0;
#pragma unroll
    for (int k = 0; k < CN; k++) mask_a |= (uint16_t)(1u << (xr + k * CM));   // CTAs sharing my M tile
#pragma unroll
    for (int k = 0; k < CM; k++) mask_b |= (uint16_t)(1u << (k + yr * CM));   // CTAs sharing my N tile
    mask_rel = mask_a | mask_b;
  }
  if (warp == 1) {
    if (PAIR) {  // one warp of EACH CTA of the pair: the same columns are allocated in both CTAs' TMEM
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(BN)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(BN)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (kCluster) g_cluster_sync();  // every CTA's barriers are initialised before any remote multicast / commit
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (flags & S3R_EPI_PDL) {
    // programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped
    // the tail of the previous kernel in the stream; let OUR dependent start its own prologue now, then wait until
    // the previous kernel has completed and flushed before touching any global memory
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  if (warp == 0) {
    if (lane == 0) {
      int cw0 = 0, ch0 = 0, cn0 = 0;
      if (kConv) {  // first pixel of this CTA's patch
        cw0 = conv.W >= GEMM_BM ? m0 % conv.W : 0;
        ch0 = (m0 / conv.W) % conv.H;
        cn0 = m0 / (conv.W * conv.H);
      }
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % GEMM_STAGES;
        g_mbar_wait(&empty[s], ((kb / GEMM_STAGES) & 1) ^ 1);
        uint8_t* a_dst = smem + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + S::A_BYTES;
        if (PAIR) {
          // both CTAs' bytes land on the leader's barrier; only the leader posts the expectation (for both)
          const uint32_t lbar = g_mapa(g_smem_u32(&full[s]), 0);
          if (prank == 0) g_mbar_expect_tx(&full[s], 2 * S::STAGE_BYTES);
          if (kConv) {
            const int tap = (kb0 + kb) / conv.cblocks, cb = (kb0 + kb) - tap * conv.cblocks;
            const int kh = tap / conv.KW, kw = tap - kh * conv.KW;
            g_tma_load_4d_pair(a_dst, &tmA, cb * GEMM_BK, cw0 + kw - conv.pad, ch0 + kh - conv.pad, cn0, lbar);
          } else {
            g_tma_load_2d_pair(a_dst, &tmA, (kb0 + kb) * GEMM_BK, m0, lbar);
          }
          g_tma_load_2d_pair(b_dst, &tmB, (kb0 + kb) * GEMM_BK, n0 + (int)prank * (BN / 2), lbar);
          continue;
        }
        g_mbar_expect_tx(&full[s], S::STAGE_BYTES);
        if (kConv) {
          const int tap = (kb0 + kb) / conv.cblocks, cb = (kb0 + kb) - tap * conv.cblocks;
          const int kh = tap / conv.KW, kw = tap - kh * conv.KW;
          g_tma_load_4d(a_dst, &tmA, cb * GEMM_BK, cw0 + kw - conv.pad, ch0 + kh - conv.pad, cn0, &full[s]);
        } else if (CN > 1) {  // my 128/CN-row slice of the A tile, to every CTA that shares this M tile
          constexpr int AR = GEMM_BM / CN;
          g_tma_load_2d_mc(a_dst + yr * AR * 128, &tmA, (kb0 + kb) * GEMM_BK, m0 + yr * AR, &full[s], mask_a);
        } else if (MAJ & 1) {  // two 64-row atoms of the MN-major A tile: box = 64 contraction rows x 64 M elements
#pragma unroll
          for (int h = 0; h < GEMM_BM / 64; h++) {
            if (batched) g_tma_load_3d(a_dst + h * 8192, &tmA, m0 + h * 64, (kb0 + kb) * GEMM_BK, blockIdx.z, &full[s]);
            else g_tma_load_2d(a_dst + h * 8192, &tmA, m0 + h * 64, (kb0 + kb) * GEMM_BK, &full[s]);
          }
        } else if (batched) {
          g_tma_load_3d(a_dst, &tmA, (kb0 + kb) * GEMM_BK, m0, blockIdx.z, &full[s]);
        } else {
          g_tma_load_2d(a_dst, &tmA, (kb0 + kb) * GEMM_BK, m0, &full[s]);
        }
        if (MAJ & 2) {
#pragma unroll
          for (int h = 0; h < BN / 64; h++) {
            if (batched) g_tma_load_3d(b_dst + h * 8192, &tmB, n0 + h * 64, (kb0 + kb) * GEMM_BK, blockIdx.z, &full[s]);
            else g_tma_load_2d(b_dst + h * 8192, &tmB, n0 + h * 64, (kb0 + kb) * GEMM_BK, &full[s]);
          }
        } else if (batched) {
          g_tma_load_3d(b_dst, &tmB, (kb0 + kb) * GEMM_BK, n0, blockIdx.z, &full[s]);
        } else if (CM > 1) {         // my BN/CM-row slice of the W tile, to every CTA that shares this N tile
          constexpr int BR = BN / CM;
          g_tma_load_2d_mc(b_dst + xr * BR * 128, &tmB, (kb0 + kb) * GEMM_BK, n0 + xr * BR, &full[s], mask_b);
        } else {
          g_tma_load_2d(b_dst, &tmB, (kb0 + kb) * GEMM_BK, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && prank == 0) {
      const uint32_t idesc = g_make_idesc(PAIR ? 2 * GEMM_BM : GEMM_BM, BN, MAJ & 1, (MAJ >> 1) & 1);
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % GEMM_STAGES;
        g_mbar_wait(&full[s], (kb / GEMM_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t adesc = (MAJ & 1) ? g_make_desc_mn(smem + s * S::STAGE_BYTES) : g_make_desc(smem + s * S::STAGE_BYTES);
        const uint64_t bdesc = (MAJ & 2) ? g_make_desc_mn(smem + s * S::STAGE_BYTES + S::A_BYTES)
                                         : g_make_desc(smem + s * S::STAGE_BYTES + S::A_BYTES);
        // one k-step = 16 contraction elements: K-major +32 B inside the 128-B swizzle atom (+2 in >>4 units);
        // MN-major +16 rows of 128 B (+128)
        constexpr int AK = (MAJ & 1) ? 128 : 2, BK_ = (MAJ & 2) ? 128 : 2;
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; k++) {
          if (PAIR) g_umma_pair(tmem_base, adesc + (uint64_t)(AK * k), bdesc + (uint64_t)(BK_ * k), idesc, (kb | k) ? 1u : 0u);
          else g_umma(tmem_base, adesc + (uint64_t)(AK * k), bdesc + (uint64_t)(BK_ * k), idesc, (kb | k) ? 1u : 0u);
        }
        // frees the ring slot once these MMAs have read it (in every CTA that multicasts into it / in both CTAs of a pair)
        if (PAIR) g_umma_commit_pair(&empty[s]);
        else if (kCluster) g_umma_commit_mc(&empty[s], mask_rel);
        else g_umma_commit(&empty[s]);
      }
      if (PAIR) g_umma_commit_pair(tmem_full);  // accumulator complete (both halves)
      else g_umma_commit(tmem_full);
    }
  }
  __syncwarp();  // reconverge the single-lane producer / MMA warps before they join the epilogue's second phase
  // ===== epilogue
  // phase 1 (warps 2-5, TMEM lane quarter = warp % 4): TMEM -> registers -> fp32 staging tile in shared memory (the TMA
  //          ring is idle once tmem_full fires);
  // phase 2 (all 256 threads; only the 128 epilogue threads when split-K is on): the tile is walked row-wise, 4
  //          consecutive columns per lane, so that every global access (partials, bias, residual, output) is a fully
  //          coalesced 256/512-byte row segment.
  if constexpr (DIRECT != 0) {
    if (warp >= 2) {
      const int q = warp & 3, half = (warp - 2) >> 2;
      constexpr int NBLK = BN / 64;  // 32-column blocks per warp
      float* sb = reinterpret_cast<float*>(smem + GEMM_STAGES * S::STAGE_BYTES + 256) + (warp - 2) * (BN / 2);
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const bool has_bias = EF(S3R_EPI_BIAS), gelu = EF(S3R_EPI_GELU), has_res = EF(S3R_EPI_RESIDUAL), relu = EF(S3R_EPI_RELU),
                 out_f32 = EF(S3R_EPI_OUT_F32);
      const int cb0 = n0 + half * (BN / 2);
      // everything that does not depend on the accumulator is requested now, while the main loop runs
      int py = 0, px = 0;
      if (EF(S3R_EPI_ROPE) && row_ok) {
        const long long y = rope.pos[(size_t)row * 2], x = rope.pos[(size_t)row * 2 + 1];
        py = (int)(y < 0 ? 0 : (y > rope.max_pos ? rope.max_pos : y));
        px = (int)(x < 0 ? 0 : (x > rope.max_pos ? rope.max_pos : x));
      }
      if (has_bias) {
#pragma unroll
        for (int c = lane; c < BN / 2; c += 32) sb[c] = (cb0 + c < N) ? __bfloat162float(bias[cb0 + c]) : 0.0f;
        __syncwarp();
      }
      const bool vec_ok = (N % 8 == 0) && !(((uintptr_t)residual | (uintptr_t)Cout) & 15) && (ldr % 8 == 0);
      uint4 rq[NBLK][4];
      if (has_res) {
#pragma unroll
        for (int blk = 0; blk < NBLK; blk++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            rq[blk][j] = make_uint4(0u, 0u, 0u, 0u);
            const int c = cb0 + blk * 32 + 8 * j;
            if (row_ok && vec_ok && c + 8 <= N) rq[blk][j] = *reinterpret_cast<const uint4*>(residual + (size_t)row * ldr + c);
          }
      }
      g_mbar_wait(tmem_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int blk = 0; blk < NBLK; blk++) {
        const int c0 = half * (BN / 2) + blk * 32;
        const int col = n0 + c0;
        uint32_t v[32];
        g_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        float f[32];
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + blk * 32 + 4 * j);
            f[4 * j] = __uint_as_float(v[4 * j]) + b4.x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
        }
        if (EF(S3R_EPI_ROPE) && col < rope.cols) {
          const float2* tb = rope.table + (((col >> 5) & 1) ? px : py) * 16;
#pragma unroll
          for (int d = 0; d < 16; d++) {
            const float2 cs = __ldg(tb + d);
            const float u = f[d], w2 = f[d + 16];
            f[d] = u * cs.x - w2 * cs.y;
            f[d + 16] = w2 * cs.x + u * cs.y;
          }
        }
        if (gelu) {
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = 0.5f * f[j] * (1.0f + g_fast_erf(f[j] * 0.70710678118654752f));
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.0f);
        }
        if (has_res) {
          if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const uint32_t rw[4] = {rq[blk][j].x, rq[blk][j].y, rq[blk][j].z, rq[blk][j].w};
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const float2 rr = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[k]));
                f[8 * j + 2 * k] += rr.x;
                f[8 * j + 2 * k + 1] += rr.y;
              }
            }
          } else if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (col + j < N) f[j] += __bfloat162float(residual[(size_t)row * ldr + col + j]);
          }
        }
        if (row_ok) {
          if (out_f32) {
            float* op = (float*)Cout + (size_t)row * ldc + col;
            if (vec_ok) {
#pragma unroll
              for (int j = 0; j < 8; j++)
                if (col + 4 * j + 4 <= N)
                  *reinterpret_cast<float4*>(op + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (col + j < N) op[j] = f[j];
            }
          } else {
            __nv_bfloat16* op = (__nv_bfloat16*)Cout + (size_t)row * ldc + col;
            if (vec_ok) {
#pragma unroll
              for (int j = 0; j < 4; j++) {
                if (col + 8 * j + 8 <= N) {
                  uint4 u;
                  *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
                  *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                  *reinterpret_cast<__nv_bfloat162*>(&u.z) = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
                  *reinterpret_cast<__nv_bfloat162*>(&u.w) = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                  *reinterpret_cast<uint4*>(op + 8 * j) = u;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (col + j < N) op[j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
      }
    }
  } else {
  const bool is_p1 = warp >= 2 && warp < 6;
  const int NT2 = (splits == 1 || KS > 1) ? GEMM_THREADS : 128;   // threads of phase 2
  // phase-2 thread index: epilogue warps 0..127, then the TMA / MMA warps, then warps 6-7
  const int et = is_p1 ? (int)threadIdx.x - 64 : (warp < 2 ? 128 + (int)threadIdx.x : (int)threadIdx.x);
  if (et < NT2) {
    const int q = warp & 3;
    constexpr int CG = BN > 128 ? 128 : BN;  // columns staged per pass (BN = 256: two passes over one 68 KB staging tile)
    constexpr int LDS_ = CG + 4;      // padded row pitch (floats): conflict-free for both phases
    float* stage = reinterpret_cast<float*>(smem);
    int* spos = reinterpret_cast<int*>(smem + GEMM_STAGES * S::STAGE_BYTES - 2048);  // [128][2] row positions (RoPE)
    if (is_p1) {
      g_mbar_wait(tmem_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (is_p1 && EF(S3R_EPI_ROPE)) {
      {  // one coalesced read of this tile's 128 (y, x) positions, clamped to the table
        const int prow = m0 + q * 32 + lane;
        long long py = 0, px = 0;
        if (prow < M) {
          py = rope.pos[(size_t)prow * 2];
          px = rope.pos[(size_t)prow * 2 + 1];
        }
        spos[(q * 32 + lane) * 2] = (int)(py < 0 ? 0 : (py > rope.max_pos ? rope.max_pos : py));
        spos[(q * 32 + lane) * 2 + 1] = (int)(px < 0 ? 0 : (px > rope.max_pos ? rope.max_pos : px));
      }
    }
#pragma unroll 1
    for (int cg = 0; cg < BN / CG; cg++) {
    if (is_p1) {
      float* srow = stage + (size_t)(q * 32 + lane) * LDS_;
#pragma unroll 1
      for (int c = 0; c < CG / 32; c++) {
        uint32_t v[32];
        g_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CG + c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(srow + c * 32 + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
    if (KS > 1) g_cluster_sync();  // every rank's partial tile is parked in its staging buffer (all 256 threads take part)
    else asm volatile("bar.sync 1, %0;" ::"r"(NT2) : "memory");
    constexpr int LPR = CG / 4;        // lanes per row
    const int RPI = NT2 / LPR;         // rows per iteration of the phase-2 threads
    const int cl = (et % LPR) * 4;     // this lane's first column inside the tile
    const int col = n0 + cg * CG + cl;
    const bool has_bias = EF(S3R_EPI_BIAS), gelu = EF(S3R_EPI_GELU), has_res = EF(S3R_EPI_RESIDUAL),
               relu = EF(S3R_EPI_RELU),
               out_f32 = EF(S3R_EPI_OUT_F32), do_rope = EF(S3R_EPI_ROPE) && col < rope.cols;
    const bool full4 = col + 4 <= N;
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};
    if (has_bias) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (col + j < N) bias4[j] = __bfloat162float(bias[col + j]);
    }
    auto finish = [&](float (&f)[4], int row) {  // f = accumulators of columns col..col+3 of `row`
#pragma unroll
      for (int j = 0; j < 4; j++) f[j] += bias4[j];
      if (EF(S3R_EPI_ROPE)) {
        // pairs (d, d+16) inside each 32-column half head sit 4 lanes apart: exchange with lane ^ 4 (all lanes shuffle)
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = __shfl_xor_sync(0xffffffffu, f[j], 4);
        if (do_rope && row < M) {
          const int pp = spos[(row - m0) * 2 + ((col >> 5) & 1)];
          const int d0 = col & 15;
          const bool lower = (col & 16) == 0;  // this lane holds u (d < 16) or v (d >= 16)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float2 cs = __ldg(rope.table + pp * 16 + d0 + j);
            f[j] = lower ? (f[j] * cs.x - o[j] * cs.y) : (f[j] * cs.x + o[j] * cs.y);
          }
        }
      }
      if (EF(S3R_EPI_SAVE_PRE) && row < M && col < N) {  // training forward: keep the pre-activation for gelu'
        __nv_bfloat16* pp = reinterpret_cast<__nv_bfloat16*>(ws) + (size_t)row * ldc + col;
        if (full4) {
          uint2 u;
          *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(f[0], f[1]);
          *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(f[2], f[3]);
          *reinterpret_cast<uint2*>(pp) = u;
        } else {
          for (int j = 0; j < 4 && col + j < N; j++) pp[j] = __float2bfloat16_rn(f[j]);
        }
      }
      if (gelu) {
#pragma unroll
        for (int j = 0; j < 4; j++) f[j] = 0.5f * f[j] * (1.0f + g_fast_erf(f[j] * 0.70710678118654752f));
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 4; j++) f[j] = fmaxf(f[j], 0.0f);
      }
      if (EF(S3R_EPI_DGELU) && row < M && col < N) {
        // backward of y = gelu(h): f = dL/dy (accumulator) -> dL/dh = f * gelu'(h), h = pre-activation in `residual`
        const __nv_bfloat16* hp = residual + (size_t)row * ldr + col;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (col + j < N) {
            const float h = __bfloat162float(hp[j]);
            const float cdf = 0.5f * (1.0f + g_fast_erf(h * 0.70710678118654752f));
            f[j] *= cdf + h * 0.39894228040143268f * __expf(-0.5f * h * h);
          }
        }
      }
      if (row < M && col < N) {
        if (has_res && EF(S3R_EPI_RES_F32)) {  // fp32 residual stream (ldr in fp32 elements)
          const float* rp = reinterpret_cast<const float*>(residual) + (size_t)row * ldr + col;
          if (full4) {
            const float4 r4 = *reinterpret_cast<const float4*>(rp);
            f[0] += r4.x; f[1] += r4.y; f[2] += r4.z; f[3] += r4.w;
          } else {
            for (int j = 0; j < 4 && col + j < N; j++) f[j] += rp[j];
          }
        } else if (has_res) {
          const __nv_bfloat16* rp = residual + (size_t)row * ldr + col;
          if (full4) {
            const uint2 u = *reinterpret_cast<const uint2*>(rp);
            const float2 r0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 r1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
          } else {
            for (int j = 0; j < 4 && col + j < N; j++) f[j] += __bfloat162float(rp[j]);
          }
        }
        if (out_f32) {
          float* op = (float*)Cout + (size_t)row * ldc + col;
          if (full4) *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
          else for (int j = 0; j < 4 && col + j < N; j++) op[j] = f[j];
        } else {
          __nv_bfloat16* op = (__nv_bfloat16*)Cout + (size_t)row * ldc + col;
          if (full4) {
            uint2 u;
            *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(f[0], f[1]);
            *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(f[2], f[3]);
            *reinterpret_cast<uint2*>(op) = u;
          } else {
            for (int j = 0; j < 4 && col + j < N; j++) op[j] = __float2bfloat16_rn(f[j]);
          }
        }
      }
    };
    if (KS > 1) {
      if (blockIdx.z == 0) {  // rank 0: sum the KS partial tiles through distributed shared memory, then the fused epilogue
#pragma unroll 2
        for (int r = et / LPR; r < GEMM_BM; r += RPI) {
          const float* sp = stage + (size_t)r * LDS_ + cl;
          const float4 t = *reinterpret_cast<const float4*>(sp);
          float f[4] = {t.x, t.y, t.z, t.w};
          const uint32_t local = g_smem_u32(sp);
#pragma unroll
          for (int k = 1; k < KS; k++) {
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(k));
            float4 p;
            asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w)
                         : "r"(remote)
                         : "memory");
            f[0] += p.x; f[1] += p.y; f[2] += p.z; f[3] += p.w;
          }
          finish(f, m0 + r);
        }
      }
    } else if (splits == 1) {
#pragma unroll 2
      for (int r = et / LPR; r < GEMM_BM; r += RPI) {
        const float4 t = *reinterpret_cast<const float4*>(stage + (size_t)r * LDS_ + cl);
        float f[4] = {t.x, t.y, t.z, t.w};
        finish(f, m0 + r);
      }
    } else {
      // ---- split-K, phase A: park this CTA's partial tile in the workspace (coalesced)
      for (int r = et / LPR; r < GEMM_BM; r += RPI) {
        const int row = m0 + r;
        if (row < M && col < N) {
          const float4 t = *reinterpret_cast<const float4*>(stage + (size_t)r * LDS_ + cl);
          float* wp = ws + ((size_t)blockIdx.z * M + row) * N + col;
          if (full4) __stcg(reinterpret_cast<float4*>(wp), t);
          else {
            const float tt[4] = {t.x, t.y, t.z, t.w};
            for (int j = 0; j < 4 && col + j < N; j++) __stcg(wp + j, tt[j]);
          }
        }
      }
      // release: the CTA barrier orders the 128 threads' partial stores before ONE gpu-scope fence + ticket atomic
      // (fences are cumulative); acquire on the other side the same way; partials are read with ld.cg (L2)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        __threadfence();
        const unsigned t = atomicAdd(&counters[blockIdx.y * gridDim.x + blockIdx.x], 1u);
        __threadfence();
        if (t == (unsigned)splits - 1) counters[blockIdx.y * gridDim.x + blockIdx.x] = 0u;  // ready for the next launch
        *s_ticket = t;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*s_ticket == (uint32_t)splits - 1) {
        // ---- phase B (last CTA of this tile): sum the partials in split order, then the fused epilogue.
        // 4 rows x up to 8 splits of independent ld.cg are put in flight together (a dependent load per split and
        // row would cost one L2 round trip each: 16 x splits x ~0.3 us)
        constexpr int RB = 2;
        for (int rb = et / LPR; rb < GEMM_BM; rb += RPI * RB) {
          float4 part[RB][8];
#pragma unroll
          for (int i = 0; i < RB; i++) {
            const int row = m0 + rb + i * RPI;
#pragma unroll
            for (int z = 0; z < 8; z++) {
              part[i][z] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (z < splits && row < M && full4)
                part[i][z] = __ldcg(reinterpret_cast<const float4*>(ws + ((size_t)z * M + row) * N + col));
            }
          }
#pragma unroll
          for (int i = 0; i < RB; i++) {
            const int row = m0 + rb + i * RPI;
            float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int z = 0; z < 8; z++) {
              f[0] += part[i][z].x; f[1] += part[i][z].y; f[2] += part[i][z].z; f[3] += part[i][z].w;
            }
            if (!full4 && row < M && col < N) {  // ragged N tail: scalar path
              for (int z = 0; z < splits; z++)
                for (int j = 0; j < 4 && col + j < N; j++) f[j] += __ldcg(ws + ((size_t)z * M + row) * N + col + j);
            }
            finish(f, row);
          }
        }
      }
    }
    if (BN > CG) asm volatile("bar.sync 1, %0;" ::"r"(NT2) : "memory");  // staging tile is reused by the next column group
    }  // cg
  }
  }  // staged epilogue
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (kCluster || KS > 1) g_cluster_sync();  // no CTA may retire while a peer can still multicast into it / read its tile
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}


// ------------------------------------------------------------------------------------------ persistent CTA pairs
// The many-wave shapes (M = 4112 of cfg3, the large convolutions): ONE resident CTA pair per TPC walks the list of
// 256 x BN output tiles (static round-robin over the pairs), so that
//   * there is no wave quantisation (204 pair tiles over 74 pairs cost 3 tile times, not 2 waves of half-empty SMs),
//   * the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns): the MMAs of tile i+1 start while the epilogue warps
//     still drain tile i, and the TMA ring (the whole shared memory: no staging tile) never stops across tile boundaries,
//   * barrier set-up, TMEM allocation and the tensor-map fetch are paid once per SM instead of once per tile.
// Roles: warp 0 = TMA producer (both CTAs), warp 1 = MMA issuer (leader CTA only, cta_group::2), warps 2-9 = epilogue:
// two warps per TMEM lane quarter (column halves), each thread owns ONE output row and walks it in 32-column blocks
// straight from TMEM registers: bias / GELU / ReLU / residual / RoPE (the rotation pairs (d, d+16) sit in the same
// thread) -> 64 contiguous bytes per thread and block to global memory.  No staging tile, no CTA barrier in the loop
// (the staged two-phase epilogue of the one-tile kernel cost 11.6 k warp instructions per tile here and made the
// K = 1024 shapes epilogue-bound: ncu tensor pipe 42 %).
// tmem_full[b]: multicast commit of the leader -> both CTAs; tmem_empty[b]: lives in the leader, one arrival per
// epilogue warp of BOTH CTAs (16), remote arrivals through mapa + mbarrier.arrive.shared::cluster.
__device__ __forceinline__ void g_mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  // .relaxed: the arrival only hands back TMEM columns (ordered by tcgen05.fence::before_thread_sync); the default
  // .release at cluster scope compiles to MEMBAR.ALL.GPU, which stalls the warp until all of its earlier global stores
  // have been acknowledged - measured: the accumulator hand-back was delayed by ~2 us per tile
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

template <int BN, int STAGES_>
struct PersistSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = STAGES_;
  static constexpr int BIAS_BYTES = 8 * (BN / 2) * 4;  // per epilogue warp: its column half of the tile's bias, fp32
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BIAS_BYTES + 256 /* barriers */ + 1024 /* alignment */;
};

#define GEMM_P_EPI_WARPS 8
#define GEMM_P_THREADS (64 + 32 * GEMM_P_EPI_WARPS)
template <int BN, int STAGES_, bool kConv, int EPI>
__global__ void __launch_bounds__(GEMM_P_THREADS, 1)
s3r_gemm_pair_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* __restrict__ residual,
                                void* __restrict__ Cout, int M, int N, int K, int ldc, int ldr, int flags, RopeArgs rope,
                                ConvArgs conv, int mt_pairs, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  using S = PersistSmem<BN, STAGES_>;
#define EFP(X) (((EPI) & (X)) && (flags & (X)))
  constexpr int ST = S::STAGES;
  float* sbias_all = reinterpret_cast<float*>(smem + ST * S::STAGE_BYTES);
  uint64_t* full = (uint64_t*)(smem + ST * S::STAGE_BYTES + S::BIAS_BYTES);
  uint64_t* empty = full + ST;
  uint64_t* tmem_full = empty + ST;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2] (used in the leader CTA)
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t prank = g_cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < ST; s++) {
      g_mbar_init(&full[s], 1);
      g_mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      g_mbar_init(&tmem_full[b], 1);
      g_mbar_init(&tmem_empty[b], 2 * GEMM_P_EPI_WARPS);  // every epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  g_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (flags & S3R_EPI_PDL) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = cid; t < total_tiles; t += ncl) {
        const int m0 = ((t % mt_pairs) * 2 + (int)prank) * GEMM_BM, n0 = (t / mt_pairs) * BN;
        int cw0 = 0, ch0 = 0, cn0 = 0;
        if (kConv) {
          cw0 = conv.W >= GEMM_BM ? m0 % conv.W : 0;
          ch0 = (m0 / conv.W) % conv.H;
          cn0 = m0 / (conv.W * conv.H);
        }
        for (int kb = 0; kb < num_kb; kb++, it++) {
          const int s = it % ST;
          g_mbar_wait(&empty[s], ((it / ST) & 1) ^ 1);
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          uint8_t* b_dst = a_dst + S::A_BYTES;
          const uint32_t lbar = g_mapa(g_smem_u32(&full[s]), 0);
          if (prank == 0) g_mbar_expect_tx(&full[s], 2 * S::STAGE_BYTES);
          if (kConv) {
            const int tap = kb / conv.cblocks, cb = kb - tap * conv.cblocks;
            const int kh = tap / conv.KW, kw = tap - kh * conv.KW;
            g_tma_load_4d_pair(a_dst, &tmA, cb * GEMM_BK, cw0 + kw - conv.pad, ch0 + kh - conv.pad, cn0, lbar);
          } else {
            g_tma_load_2d_pair(a_dst, &tmA, kb * GEMM_BK, m0, lbar);
          }
          g_tma_load_2d_pair(b_dst, &tmB, kb * GEMM_BK, n0 + (int)prank * (BN / 2), lbar);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && prank == 0) {
      const uint32_t idesc = g_make_idesc(2 * GEMM_BM, BN);
      uint32_t it = 0, i = 0;
      for (int t = cid; t < total_tiles; t += ncl, i++) {
        const uint32_t b = i & 1;
        g_mbar_wait(&tmem_empty[b], ((i >> 1) & 1) ^ 1);  // both CTAs have drained this accumulator buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + b * BN;
        for (int kb = 0; kb < num_kb; kb++, it++) {
          const int s = it % ST;
          g_mbar_wait(&full[s], (it / ST) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t adesc = g_make_desc(smem + s * S::STAGE_BYTES);
          const uint64_t bdesc = g_make_desc(smem + s * S::STAGE_BYTES + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; k++)
            g_umma_pair(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          g_umma_commit_pair(&empty[s]);
        }
        g_umma_commit_pair(&tmem_full[b]);
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter q = warp % 4 (hardware rule), column half = (warp - 2) / 4
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int NB = BN / 64;  // 32-column blocks per warp and tile
    const uint32_t empty_bar0 = g_mapa(g_smem_u32(&tmem_empty[0]), 0), empty_bar1 = g_mapa(g_smem_u32(&tmem_empty[1]), 0);
    const bool has_bias = EFP(S3R_EPI_BIAS), gelu = EFP(S3R_EPI_GELU), has_res = EFP(S3R_EPI_RESIDUAL), relu = EFP(S3R_EPI_RELU),
               out_f32 = EFP(S3R_EPI_OUT_F32);
    uint32_t i = 0;
    for (int t = cid; t < total_tiles; t += ncl, i++) {
      const int m0 = ((t % mt_pairs) * 2 + (int)prank) * GEMM_BM, n0 = (t / mt_pairs) * BN;
      const uint32_t b = i & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      int py = 0, px = 0;
      if (EFP(S3R_EPI_ROPE) && row_ok) {  // requested before the accumulator wait: the latency hides behind the main loop
        const long long y = rope.pos[(size_t)row * 2], x = rope.pos[(size_t)row * 2 + 1];
        py = (int)(y < 0 ? 0 : (y > rope.max_pos ? rope.max_pos : y));
        px = (int)(x < 0 ? 0 : (x > rope.max_pos ? rope.max_pos : x));
      }
      // everything that does not depend on the accumulator is fetched BEFORE the accumulator wait, so that its global
      // latency hides behind the main loop of this tile: the warp's column half of the bias goes to its private
      // shared-memory row (fp32, read back as broadcasts), the first block's residual segment to registers
      float* sb = sbias_all + (warp - 2) * (BN / 2);
      if (has_bias) {
        const int cb0 = n0 + half * (BN / 2);
#pragma unroll
        for (int c = lane; c < BN / 2; c += 32) sb[c] = (cb0 + c < N) ? __bfloat162float(bias[cb0 + c]) : 0.0f;
        __syncwarp();
      }
      uint4 rq[4];
      auto load_res = [&](int blk_) {
        const int col_ = n0 + half * (BN / 2) + blk_ * 32;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          rq[j] = make_uint4(0u, 0u, 0u, 0u);
          if (row_ok && col_ + 8 * j + 8 <= N) rq[j] = *reinterpret_cast<const uint4*>(residual + (size_t)row * ldr + col_ + 8 * j);
        }
      };
      if (has_res) load_res(0);
      g_mbar_wait(&tmem_full[b], (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int blk = 0; blk < NB; blk++) {
        const int c0 = half * (BN / 2) + blk * 32;
        const int col = n0 + c0;
        uint32_t v[32];
        g_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + c0), v);
        if (blk == NB - 1) {  // this warp's share of the accumulator buffer is in registers: hand the buffer back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) g_mbar_arrive_cluster(b ? empty_bar1 : empty_bar0);
        }
        float f[32];
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + blk * 32 + 4 * j);  // broadcast read
            f[4 * j] = __uint_as_float(v[4 * j]) + b4.x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
        }
        if (EFP(S3R_EPI_ROPE) && col < rope.cols) {
          // 32 columns = one half head: (u, v) = (f[d], f[d + 16]), position y for the first half of a head, x for the second
          const float2* tb = rope.table + (((col >> 5) & 1) ? px : py) * 16;
#pragma unroll
          for (int d = 0; d < 16; d++) {
            const float2 cs = __ldg(tb + d);
            const float u = f[d], w2 = f[d + 16];
            f[d] = u * cs.x - w2 * cs.y;
            f[d + 16] = w2 * cs.x + u * cs.y;
          }
        }
        if (gelu) {
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = 0.5f * f[j] * (1.0f + g_fast_erf(f[j] * 0.70710678118654752f));
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.0f);
        }
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t rw[4] = {rq[j].x, rq[j].y, rq[j].z, rq[j].w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const float2 rr = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[k]));
              f[8 * j + 2 * k] += rr.x;
              f[8 * j + 2 * k + 1] += rr.y;
            }
          }
          // next block's residual segment: in flight during this block's stores and the next TMEM read.  C may alias
          // the residual (x += f(x)): the segment a thread loads here is only ever written by that same thread, later.
          if (blk + 1 < NB) load_res(blk + 1);
        }
        if (row_ok) {
          if (out_f32) {
            float* op = (float*)Cout + (size_t)row * ldc + col;
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (col + 4 * j + 4 <= N)
                *reinterpret_cast<float4*>(op + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            __nv_bfloat16* op = (__nv_bfloat16*)Cout + (size_t)row * ldc + col;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              if (col + 8 * j + 8 <= N) {
                uint4 u;
                *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
                *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                *reinterpret_cast<__nv_bfloat162*>(&u.z) = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
                *reinterpret_cast<__nv_bfloat162*>(&u.w) = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                *reinterpret_cast<uint4*>(op + 8 * j) = u;
              }
            }
          }
        }
      }
    }
  }
#undef EFP
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  g_cluster_sync();  // no CTA of the pair retires (or frees TMEM) while the other still computes
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int g_gemm_cluster = 0;   // S3R_TUNE_GEMM_CLUSTER: 0 = default, else CM*10 + CN
static int g_gemm_big_tile = 0;  // S3R_TUNE_GEMM_BIG_TILE: 0 = auto (wide many-wave grids), 1 = whenever >= 120 tiles, 2 = never
// S3R_TUNE_PDL: programmatic dependent launch for the encoder kernels (GEMM / conv / attention / LayerNorm).  Measured
// (scripts/pdl_probe.py, stream-ordered chains inside a CUDA graph): 8.4 -> 7.2 us per launch at 257x768x768, 9.7 -> 8.7
// at 514x1024x1024, neutral for grids whose shared memory fills the SMs; results identical.  On by default.
static int g_pdl = 1;
int s3r_pdl_enabled() { return g_pdl; }
static int g_gemm_ksplit = 0;    // S3R_TUNE_GEMM_KSPLIT: 0 = auto, 1 = never, 2 / 4 = force that cluster split-K factor on 64-wide tiles
static int g_gemm_shallow = 0;   // S3R_TUNE_GEMM_SHALLOW: small grids use the 4-stage (96 KB, 2 CTAs/SM) ring too
static int g_conv_cluster = 0;   // S3R_TUNE_CONV_CLUSTER: pairs of pixel tiles multicast the weight tile
static int g_gemm_direct = 0;    // S3R_TUNE_GEMM_DIRECT: register epilogue of the one-tile kernel: 0 = auto (grids <= 1.5 waves), 1 = whenever possible, 2 = never
static int g_gemm_pair = 0;      // S3R_TUNE_GEMM_PAIR: CTA pairs (cta_group::2, M = 256 per MMA): 0 = auto, 1 = 256x128 pair tiles whenever possible, 2 = never, 3 = 256x256 pair tiles whenever possible, 4 / 5 = PERSISTENT 256x128 / 256x256 pair tiles whenever possible

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] (ld elements per row), box = [box_rows, 64 cols], SWIZZLE_128B, zero OOB fill
static int make_map(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return S3R_ERR_CUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {GEMM_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3R_OK : S3R_ERR_CUDA;
}

template <int BN, int STAGES_, bool kConv = false, int CM = 1, int CN = 1, int MAJ = 0, int EPI = S3R_EPI_ALL, int KS = 1, int PAIR = 0, int DIRECT = 0>
static int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const void* bias, const void* residual, void* C, int M,
                       int N, int K, int ldc, int ldr, int flags, const RopeArgs& rope, int splits, float* ws, unsigned* counters,
                       cudaStream_t st, const ConvArgs& conv = ConvArgs{}, int batch = 0, long long batch_stride_c = 0) {
  static size_t configured[64] = {};  // per device: cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute
  const int smem = GemmSmem<BN, STAGES_, PAIR>::TOTAL + (DIRECT ? 8 * (BN / 2) * 4 : 0);  // + per-warp bias rows
  auto kern = s3r_gemm_bf16_kernel<BN, STAGES_, kConv, CM, CN, MAJ, EPI, KS, PAIR, DIRECT>;
  constexpr int CMX = PAIR ? 2 : CM;  // cluster extent along M
  if (KS > 1) splits = KS, ws = nullptr, counters = nullptr;
  {
    const int rc_ = s3r_ensure_dynamic_smem(kern, (size_t)smem, configured);
    if (rc_ != S3R_OK) return rc_;
  }
  // grid padded to whole clusters: CTAs past the last tile still take part in the multicast (their own loads are fully
  // out of bounds = zero fill, their epilogue is masked)
  const unsigned gx = ((M + GEMM_BM - 1) / GEMM_BM + CMX - 1) / CMX * CMX, gy = ((N + BN - 1) / BN + CN - 1) / CN * CN;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, gy, batch > 0 ? batch : splits);
  cfg.blockDim = dim3(DIRECT ? GEMM_DIRECT_THREADS : GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CMX * CN > 1 || KS > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CMX;
    attr[na].val.clusterDim.y = CN;
    attr[na].val.clusterDim.z = KS;
    na++;
  }
  if (g_pdl) {  // PDL: this launch may begin while the previous kernel of the stream drains (kernel waits before its loads)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
    flags |= S3R_EPI_PDL;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  S3R_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a, b, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual, C, M, N, K,
                                    ldc, ldr, flags, rope, splits, ws, counters, conv, batch > 0 ? batch_stride_c : 0LL));
  return S3R_OK;
}

// persistent CTA pairs: grid = 2 x min(#pair tiles, TPCs)
template <int BN, int STAGES_, bool kConv, int EPI>
static int launch_gemm_persist(const CUtensorMap& a, const CUtensorMap& b, const void* bias, const void* residual, void* C,
                               int M, int N, int K, int ldc, int ldr, int flags, const RopeArgs& rope, cudaStream_t st,
                               const ConvArgs& conv = ConvArgs{}) {
  static size_t configured[64] = {};
  static int pairs[64] = {};
  const int smem = PersistSmem<BN, STAGES_>::TOTAL;
  auto kern = s3r_gemm_pair_persistent_kernel<BN, STAGES_, kConv, EPI>;
  {
    const int rc_ = s3r_ensure_dynamic_smem(kern, (size_t)smem, configured);
    if (rc_ != S3R_OK) return rc_;
  }
  int dev = 0;
  S3R_CUDA_CHECK(cudaGetDevice(&dev));
  dev &= 63;
  if (pairs[dev] == 0) {
    int sms = 0;
    S3R_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    pairs[dev] = sms / 2 > 0 ? sms / 2 : 1;
  }
  const int mt_pairs = ((M + GEMM_BM - 1) / GEMM_BM + 1) / 2, nt = (N + BN - 1) / BN;
  const int total = mt_pairs * nt;
  cudaLaunchConfig_t cfg = {};
  // as many pairs as the tile list needs for its number of rounds and no more (117 tiles over 74 pairs take 2 rounds - so do
  // 59 pairs, and the other 15 TPCs stay free for the kernels of concurrent stream branches)
  const int rounds = (total + pairs[dev] - 1) / pairs[dev];
  const int use_pairs = (total + rounds - 1) / rounds;
  cfg.gridDim = dim3(2 * (unsigned)use_pairs, 1, 1);
  cfg.blockDim = dim3(GEMM_P_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  na++;
  if (g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
    flags |= S3R_EPI_PDL;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  S3R_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a, b, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual, C, M, N, K,
                                    ldc, ldr, flags, rope, conv, mt_pairs, total));
  return S3R_OK;
}

// 4-D bf16 NHWC activation map {C, W, H, N} with box {64, bw, bh, bn} (bw*bh*bn = 128 pixels), SWIZZLE_128B, zero fill
static int make_map_nhwc(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int bw, int bh, int bn) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return S3R_ERR_CUDA;
  const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  const cuuint32_t box[4] = {GEMM_BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3R_OK : S3R_ERR_CUDA;
}

// (cos, sin) table for the fused RoPE epilogue: table[pos][d] = cos/sin(pos * base^(-d/16)), pos in [0, max_pos]
__global__ void s3r_rope_table_kernel(float2* table, int n, float base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = i >> 4, d = i & 15;
  const float inv_freq = 1.0f / powf(base, (float)d / 16.0f);
  float sn, cs;
  sincosf((float)pos * inv_freq, &sn, &cs);
  table[i] = make_float2(cs, sn);
}

extern "C" int s3r_rope_table(float* table, int32_t max_pos, float base, void* stream) {
  if (!table || max_pos < 0) return S3R_ERR_INVALID_ARG;
  const int n = (max_pos + 1) * 16;
  s3r_rope_table_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>((float2*)table, n, base);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}

// set by s3r_gemm_bf16_majors around its call into s3r_gemm_bf16_rope: bf16 [M, N] (pitch ldc) buffer that receives the
// pre-activation (S3R_EPI_SAVE_PRE); travels in the kernel's (then unused) split-K workspace argument
static thread_local void* t_pre_out = nullptr;

extern "C" int s3r_gemm_bf16_rope(const void* A, const void* W, const void* bias, const void* residual, void* C, int32_t M,
                                  int32_t N, int32_t K, int32_t lda, int32_t ldw, int32_t ldc, int32_t ldr, int32_t flags,
                                  const int64_t* rope_pos, const float* rope_table, int32_t rope_cols,
                                  int32_t rope_max_pos, void* workspace, size_t workspace_bytes, void* stream) {
  if (M < 0 || N <= 0 || K <= 0) return S3R_ERR_INVALID_ARG;
  if (M == 0) return S3R_OK;
  if (!A || !W || !C) return S3R_ERR_INVALID_ARG;
  if ((flags & S3R_EPI_BIAS) && !bias) return S3R_ERR_INVALID_ARG;
  if ((flags & S3R_EPI_RESIDUAL) && !residual) return S3R_ERR_INVALID_ARG;
  if (flags & S3R_EPI_ROPE) {
    if (!rope_pos || !rope_table || rope_cols <= 0 || rope_cols % 64 || rope_cols > N || rope_max_pos < 0)
      return S3R_ERR_INVALID_ARG;
  }
  RopeArgs rope{(const long long*)rope_pos, (const float2*)rope_table, rope_cols, rope_max_pos};
  // TMA: 16-byte aligned base and row pitch; vector epilogue: 16-byte aligned output rows
  if (K % 8 || lda % 8 || ldw % 8 || ldc % 8 || ((flags & S3R_EPI_RESIDUAL) && ldr % 8)) return S3R_ERR_UNSUPPORTED;
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) & 15) return S3R_ERR_UNSUPPORTED;
  CUtensorMap ta, tb;
  int rc;
  // fewer than ~1 wave of 128x128 tiles: use 128x64 tiles to fill the 148 SMs
  const long mt = (M + 127) / 128;
  const long tiles128 = mt * ((N + 127) / 128), tiles64 = mt * ((N + 63) / 64);
  int BN = tiles128 >= 120 ? 128 : 64;  // (BN=32 measured slower: every N-tile re-reads the A tile from L2)
  // 128x256 tiles halve the A re-reads; measured faster only for wide, many-wave grids (4112x4096x1024: 44.2 vs 50.7 us;
  // 4112x3072: 41.3 vs 38.8 us; 8192^3: 825 vs 979 us) - S3R_TUNE_GEMM_BIG_TILE: 0 auto, 1 always, 2 never
  if (BN == 128 && N % 256 == 0 && g_gemm_big_tile != 2 &&
      ((g_gemm_big_tile == 1 && mt * (N / 256) >= 120) || (N >= 4096 && mt * (N / 256) >= 444)))
    BN = 256;
  // CTA pairs (cta_group::2): 256-row MMAs, every CTA ingests its own A rows and only HALF of the B tile
  {
    int pair_bn = 0;
    if (g_gemm_pair == 1 && mt >= 2 && N % 128 == 0) pair_bn = 128;
    if (g_gemm_pair == 3 && mt >= 2 && N % 256 == 0) pair_bn = 256;
    if (g_gemm_pair == 0) {
      // measured on B200 (scripts/bench_pair.py): 256x256 pair tiles win once the grid is many waves deep (8192^3: 1375 ->
      // 1547 TFLOP/s; cuBLAS 1643) and lose to wave quantisation on the 1.4-wave grids of M = 4112; 256x128 pair tiles
      // (half the B ingress per CTA, same CTA count) win on the long-K / narrow-N shapes (4112x1024x4096: 30.7 -> 27.6 us,
      // 4112x768x3072: 24.1 -> 21.6)
      if (N % 128 == 0 && N <= 1024 && K >= 3072 && mt >= 16) pair_bn = 128;  // (N % 256 != 0 cases the persistent kernel did not take)
    }
    int persist_bn = 0;
    constexpr int E_PERSIST = S3R_EPI_BIAS | S3R_EPI_OUT_F32 | S3R_EPI_PDL | S3R_EPI_GELU | S3R_EPI_RESIDUAL | S3R_EPI_ROPE | S3R_EPI_RELU;
    if (g_gemm_pair == 4 && mt >= 2 && N % 128 == 0) persist_bn = 128;
    if (g_gemm_pair == 5 && mt >= 2 && N % 256 == 0) persist_bn = 256;
    // auto: the many-row shapes (M = 4112 of cfg3).  Measured on B200 (scripts/bench_pair.py; bias / +gelu / +residual, us):
    //   4112x3072x1024  30.0/39.6/42.1 -> 24.5/29.8/31.1     4112x4096x1024  33.2/45.7/65.1 -> 28.8/36.8/37.5
    //   4112x1024x4096  30.3/33.3/36.6 -> 26.2/27.9/28.3     4112x768x3072   23.4/26.9/27.1 -> 22.1/23.9/24.1
    // and slower below ~2000 rows (1028x3072x1024: 10.7 -> 12.6: a 256-row pair tile pads 1028 rows to 1280)
    if (g_gemm_pair == 0 && M >= 2048 && N % 256 == 0 && ((mt + 1) / 2) * (long)(N / 256) >= 48) persist_bn = 256;
    if (persist_bn && !t_pre_out && !(flags & ~E_PERSIST) && N % 8 == 0 && !(((uintptr_t)bias | (uintptr_t)residual) & 15)) {
      if ((rc = make_map(&ta, A, M, K, lda, GEMM_BM)) != S3R_OK) return rc;
      if ((rc = make_map(&tb, W, N, K, ldw, persist_bn / 2)) != S3R_OK) return rc;
      cudaStream_t st = (cudaStream_t)stream;
      constexpr int E_BASE = S3R_EPI_BIAS | S3R_EPI_OUT_F32 | S3R_EPI_PDL;
      const int extra = flags & ~E_BASE;
#define S3R_GEMM_PERSIST(BN_, ST_)                                                                                       \
  do {                                                                                                                  \
    if (extra == 0)                                                                                                     \
      return launch_gemm_persist<BN_, ST_, false, E_BASE>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, st); \
    if (extra == S3R_EPI_GELU)                                                                                          \
      return launch_gemm_persist<BN_, ST_, false, E_BASE | S3R_EPI_GELU>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, st); \
    if (extra == S3R_EPI_RESIDUAL)                                                                                      \
      return launch_gemm_persist<BN_, ST_, false, E_BASE | S3R_EPI_RESIDUAL>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, st); \
    if (extra == S3R_EPI_ROPE)                                                                                          \
      return launch_gemm_persist<BN_, ST_, false, E_BASE | S3R_EPI_ROPE>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, st); \
    return launch_gemm_persist<BN_, ST_, false, E_PERSIST>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, st); \
  } while (0)
      if (persist_bn == 128) S3R_GEMM_PERSIST(128, 8);
      S3R_GEMM_PERSIST(256, 6);
#undef S3R_GEMM_PERSIST
    }
    if (pair_bn && !t_pre_out) {
      if ((rc = make_map(&ta, A, M, K, lda, GEMM_BM)) != S3R_OK) return rc;
      if ((rc = make_map(&tb, W, N, K, ldw, pair_bn / 2)) != S3R_OK) return rc;
      cudaStream_t st = (cudaStream_t)stream;
      constexpr int E_BASE = S3R_EPI_BIAS | S3R_EPI_OUT_F32 | S3R_EPI_PDL;
      const int extra = flags & ~E_BASE;
#define S3R_GEMM_PAIR(BN_, ST_)                                                                                          \
  do {                                                                                                                  \
    if (extra == 0)                                                                                                     \
      return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE, 1, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, nullptr, nullptr, st); \
    if (extra == S3R_EPI_GELU)                                                                                          \
      return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_GELU, 1, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, nullptr, nullptr, st); \
    if (extra == S3R_EPI_RESIDUAL)                                                                                      \
      return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_RESIDUAL, 1, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, nullptr, nullptr, st); \
    if (extra == S3R_EPI_ROPE)                                                                                          \
      return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_ROPE, 1, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, nullptr, nullptr, st); \
    return launch_gemm<BN_, ST_, false, 1, 1, 0, S3R_EPI_ALL, 1, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, nullptr, nullptr, st); \
  } while (0)
      if (pair_bn == 128) S3R_GEMM_PAIR(128, 4);   // 4 x 24 KB ring, 2 CTAs / SM
      S3R_GEMM_PAIR(256, 3);                        // 3 x 32 KB ring, 2 CTAs / SM (2 x 256 TMEM columns)
#undef S3R_GEMM_PAIR
    }
  }
  int cm = 1, cn = 1;  // cluster shape (multicast): tunable, else the measured default per tile class
  if (g_gemm_cluster > 0) cm = g_gemm_cluster / 10, cn = g_gemm_cluster % 10;
  if ((N + BN - 1) / BN < cn) cn = 1;
  if (mt < cm) cm = 1;
  if ((rc = make_map(&ta, A, M, K, lda, GEMM_BM / cn)) != S3R_OK) return rc;
  if ((rc = make_map(&tb, W, N, K, ldw, BN / cm)) != S3R_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // split-K for grids smaller than the machine: every SM pulls ~45 B/cycle from L2, so a 36-CTA grid streams its
  // operands at a quarter of the chip's L2 bandwidth; more CTAs with shorter K ranges fix exactly that.
  // workspace = [4096 tile counters (uint32, zero on first use, self-resetting)] [splits x M x N fp32 partials]
  int splits = 1;
  float* ws = nullptr;
  unsigned* counters = nullptr;
  if (BN == 64 && workspace && tiles64 < 120 && tiles64 <= 4096) {
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    int want = (int)((144 + tiles64 - 1) / tiles64);
    want = want > 8 ? 8 : want;
    want = want > total_kb / 2 ? total_kb / 2 : want;
    const size_t per = (size_t)M * N * sizeof(float);
    while (want > 1 && 16384 + per * want > workspace_bytes) want--;
    if (want > 1) {
      splits = want;
      counters = (unsigned*)workspace;
      ws = (float*)((char*)workspace + 16384);
    }
  }
  if (t_pre_out) {
    splits = 1, counters = nullptr;
    ws = (float*)t_pre_out;
    flags |= S3R_EPI_SAVE_PRE;
  }
#define S3R_GEMM_GO(BN_, ST_, CM_, CN_)                                                                             \
  return launch_gemm<BN_, ST_, false, CM_, CN_>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, splits, ws, \
                                                 counters, st)
  // the inference feature sets of the ViT trunks get their own instances (no cluster, no split-K)
  constexpr int E_BASE = S3R_EPI_BIAS | S3R_EPI_OUT_F32 | S3R_EPI_PDL;
  const int extra = flags & ~E_BASE;
#define S3R_GEMM_EPI_DIRECT(BN_, ST_)                                                                                   \
  do {                                                                                                                  \
    if (cm == 1 && cn == 1 && splits == 1 && g_gemm_direct != 2) {                                                      \
      if (extra == 0)                                                                                                   \
        return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE, 1, 0, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, ws, counters, st); \
      if (extra == S3R_EPI_GELU)                                                                                        \
        return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_GELU, 1, 0, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, ws, counters, st); \
      if (extra == S3R_EPI_RESIDUAL)                                                                                    \
        return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_RESIDUAL, 1, 0, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, ws, counters, st); \
      if (extra == S3R_EPI_ROPE)                                                                                        \
        return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_ROPE, 1, 0, 1>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, ws, counters, st); \
    }                                                                                                                   \
  } while (0)
#define S3R_GEMM_EPI(BN_, ST_)                                                                                          \
  do {                                                                                                                  \
    if (cm == 1 && cn == 1 && splits == 1) {                                                                            \
      if (extra == 0)                                                                                                   \
        return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, ws, counters, st); \
      if (extra == S3R_EPI_GELU)                                                                                        \
        return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_GELU>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, ws, counters, st); \
      if (extra == S3R_EPI_RESIDUAL)                                                                                    \
        return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_RESIDUAL>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, ws, counters, st); \
      if (extra == S3R_EPI_ROPE)                                                                                        \
        return launch_gemm<BN_, ST_, false, 1, 1, 0, E_BASE | S3R_EPI_ROPE>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, ws, counters, st); \
    }                                                                                                                   \
  } while (0)
#define S3R_GEMM_CLUSTERS(BN_, ST_)                  \
  do {                                               \
    if (cm == 1 && cn == 2) S3R_GEMM_GO(BN_, ST_, 1, 2); \
    if (cm == 1 && cn == 4) S3R_GEMM_GO(BN_, ST_, 1, 4); \
    if (cm == 2 && cn == 1) S3R_GEMM_GO(BN_, ST_, 2, 1); \
    if (cm == 2 && cn == 2) S3R_GEMM_GO(BN_, ST_, 2, 2); \
    if (cm == 2 && cn == 4) S3R_GEMM_GO(BN_, ST_, 2, 4); \
    S3R_GEMM_GO(BN_, ST_, 1, 1);                     \
  } while (0)
  if (BN == 64 && cm == 1 && cn == 1 && splits == 1 && !t_pre_out && g_gemm_ksplit != 1 && !(flags & S3R_EPI_ROPE)) {
    // cluster split-K for the narrow-N / long-K GEMMs of the batch-1 encoder (fc2, proj): KS CTAs per tile
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    int ks = 1;
    if (g_gemm_ksplit == 2 || g_gemm_ksplit == 4) ks = g_gemm_ksplit;
    // measured on B200 (scripts/bench_splitk.py, +residual epilogue): 257x768x3072 18.2 -> 11.9 us (KS 4), 257x1024x4096
    // 22.9 -> 13.2 (KS 4), 514x1024x4096 23.9 -> 16.5 (KS 2), 1028x768x3072 19.4 -> 14.5 (KS 2); neutral below K = 1536
    else if (tiles64 * 4 <= 200 && total_kb >= 32) ks = 4;
    else if (tiles64 * 2 <= 300 && total_kb >= 24) ks = 2;
    // 128-wide tiles under a 4-way split move a third less through L2 than 64-wide ones at the same CTA count
    // (514x1024x4096: 40 tiles x 4 = 160 CTAs, 82 MB instead of 123 MB) - measured SLOWER (19.7 vs 16.2 us; 257x768x3072
    // 14.4 vs 11.4): the DSMEM reduction of four 128-wide partial tiles costs more than the traffic saves.  Kept behind
    // S3R_TUNE_GEMM_KSPLIT = 8 for A/B runs only.
    if (g_gemm_ksplit == 8 && N % 128 == 0) {
      if ((rc = make_map(&tb, W, N, K, ldw, 128)) != S3R_OK) return rc;
      if (extra == S3R_EPI_RESIDUAL || extra == 0)
        return launch_gemm<128, 3, false, 1, 1, 0, E_BASE | S3R_EPI_RESIDUAL, 4>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, nullptr, nullptr, st);
      return launch_gemm<128, 3, false, 1, 1, 0, S3R_EPI_ALL, 4>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, nullptr, nullptr, st);
    }
    if (ks == 4 || ks == 2) {
      const bool res_only = extra == S3R_EPI_RESIDUAL || extra == 0;
#define S3R_GEMM_KS(KS_, EPI_) \
  return launch_gemm<64, 4, false, 1, 1, 0, EPI_, KS_>(ta, tb, bias, residual, C, M, N, K, ldc, ldr, flags, rope, 1, nullptr, nullptr, st)
      if (ks == 4 && res_only) S3R_GEMM_KS(4, E_BASE | S3R_EPI_RESIDUAL);
      if (ks == 4) S3R_GEMM_KS(4, S3R_EPI_ALL);
      if (res_only) S3R_GEMM_KS(2, E_BASE | S3R_EPI_RESIDUAL);
      S3R_GEMM_KS(2, S3R_EPI_ALL);
#undef S3R_GEMM_KS
    }
  }
  // register epilogue for grids of at most ~1.5 waves (the batch-1 shapes): S3R_TUNE_GEMM_DIRECT 0 = auto, 1 = whenever
  // possible, 2 = never
  const bool direct_ok = !t_pre_out && (g_gemm_direct == 1 || (g_gemm_direct == 0 && (BN == 64 ? tiles64 : tiles128) <= 222));
  if (BN == 64) {
    if (tiles64 * splits < 148 && !g_gemm_shallow) {
      if (direct_ok) S3R_GEMM_EPI_DIRECT(64, 8);
      S3R_GEMM_EPI(64, 8);
      S3R_GEMM_CLUSTERS(64, 8);
    }
    if (direct_ok) S3R_GEMM_EPI_DIRECT(64, 4);
    S3R_GEMM_EPI(64, 4);
    S3R_GEMM_CLUSTERS(64, 4);
  }
  splits = 1, counters = nullptr;
  if (!t_pre_out) ws = nullptr;
  if (BN == 256) {
    S3R_GEMM_EPI(256, 2);
    S3R_GEMM_CLUSTERS(256, 2);
  }
  // (measured: with 128-wide tiles - two 32-column blocks per warp - the register epilogue loses to the staged one:
  // 514x3072x1024 9.8 -> 11.5 us; it is used for the 64-wide tiles only unless forced)
  if (g_gemm_direct == 1 && direct_ok && BN == 128) S3R_GEMM_EPI_DIRECT(128, 3);
  S3R_GEMM_EPI(128, 3);
  S3R_GEMM_CLUSTERS(128, 3);
#undef S3R_GEMM_EPI_DIRECT
#undef S3R_GEMM_EPI
#undef S3R_GEMM_CLUSTERS
#undef S3R_GEMM_GO
}

extern "C" int s3r_gemm_bf16(const void* A, const void* W, const void* bias, const void* residual, void* C, int32_t M,
                             int32_t N, int32_t K, int32_t lda, int32_t ldw, int32_t ldc, int32_t ldr, int32_t flags,
                             void* stream) {
  if (flags & S3R_EPI_ROPE) return S3R_ERR_INVALID_ARG;
  return s3r_gemm_bf16_rope(A, W, bias, residual, C, M, N, K, lda, ldw, ldc, ldr, flags, nullptr, nullptr, 0, 0, nullptr, 0,
                            stream);
}

// Backward GEMMs of nn.Linear on the same pipeline, operands in the layouts the forward left them in (no transposes):
//   dgrad  dX[M,Kin]    = dY[M,Nout] . W[Nout,Kin]          A K-major, B MN-major (W is [contraction, N] row-major)
//   wgrad  dW[Nout,Kin] = dY[M,Nout]^T . X[M,Kin]           A and B MN-major (both are [contraction, MN] row-major)
extern "C" int s3r_gemm_bf16_majors(const void* A, const void* B, const void* bias, const void* aux, void* C, int32_t M,
                                    int32_t N, int32_t K, int32_t lda, int32_t ldb, int32_t ldc, int32_t ldaux,
                                    int32_t flags, int32_t a_mn_major, int32_t b_mn_major, void* pre_out, void* stream) {
  if (M < 0 || N <= 0 || K <= 0) return S3R_ERR_INVALID_ARG;
  if (M == 0) return S3R_OK;
  if (!A || !B || !C) return S3R_ERR_INVALID_ARG;
  if (flags & S3R_EPI_ROPE) return S3R_ERR_INVALID_ARG;
  if ((flags & S3R_EPI_BIAS) && !bias) return S3R_ERR_INVALID_ARG;
  if ((flags & (S3R_EPI_RESIDUAL | S3R_EPI_DGELU)) && !aux) return S3R_ERR_INVALID_ARG;
  if ((flags & S3R_EPI_RESIDUAL) && (flags & S3R_EPI_DGELU)) return S3R_ERR_INVALID_ARG;  // one aux operand
  if (lda % 8 || ldb % 8 || ldc % 8 || ((flags & (S3R_EPI_RESIDUAL | S3R_EPI_DGELU)) && ldaux % 8)) return S3R_ERR_UNSUPPORTED;
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return S3R_ERR_UNSUPPORTED;
  if (pre_out && (a_mn_major || b_mn_major || (flags & S3R_EPI_OUT_F32) || ((uintptr_t)pre_out & 15))) return S3R_ERR_INVALID_ARG;
  if (!a_mn_major && !b_mn_major) {
    t_pre_out = pre_out;
    const int rc0 = s3r_gemm_bf16_rope(A, B, bias, aux, C, M, N, K, lda, ldb, ldc, ldaux, flags, nullptr, nullptr, 0, 0, nullptr, 0, stream);
    t_pre_out = nullptr;
    return rc0;
  }
  CUtensorMap ta, tb;
  int rc;
  // K-major operand: rows = M (or N), inner = K, box 64 x 128; MN-major operand: rows = K, inner = M (or N), box 64 x 64
  if ((rc = a_mn_major ? make_map(&ta, A, K, M, lda, 64) : make_map(&ta, A, M, K, lda, GEMM_BM)) != S3R_OK) return rc;
  if ((rc = b_mn_major ? make_map(&tb, B, K, N, ldb, 64) : make_map(&tb, B, N, K, ldb, 128)) != S3R_OK) return rc;
  RopeArgs rope{nullptr, nullptr, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  if (a_mn_major && b_mn_major) {
    // wgrad: a small output (dW of a 768..4096-wide layer: 36-256 tiles) under a long contraction (all tokens of the batch):
    // cluster split-K spreads the token range of every tile over 2 / 4 SMs and sums the partial tiles through DSMEM
    const long tiles = (long)((M + 127) / 128) * ((N + 127) / 128);
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    if (g_gemm_ksplit != 1 && total_kb >= 16) {
      if (tiles * 4 <= 300)
        return launch_gemm<128, 3, false, 1, 1, 3, S3R_EPI_ALL, 4>(ta, tb, bias, aux, C, M, N, K, ldc, ldaux, flags, rope, 1, nullptr, nullptr, st);
      if (tiles * 2 <= 400)
        return launch_gemm<128, 3, false, 1, 1, 3, S3R_EPI_ALL, 2>(ta, tb, bias, aux, C, M, N, K, ldc, ldaux, flags, rope, 1, nullptr, nullptr, st);
    }
    return launch_gemm<128, 3, false, 1, 1, 3>(ta, tb, bias, aux, C, M, N, K, ldc, ldaux, flags, rope, 1, nullptr, nullptr, st);
  }
  if (a_mn_major)
    return launch_gemm<128, 3, false, 1, 1, 1>(ta, tb, bias, aux, C, M, N, K, ldc, ldaux, flags, rope, 1, nullptr, nullptr, st);
  return launch_gemm<128, 3, false, 1, 1, 2>(ta, tb, bias, aux, C, M, N, K, ldc, ldaux, flags, rope, 1, nullptr, nullptr, st);
}

static int make_map_3d(CUtensorMap* map, const void* ptr, int rows, int cols, long long ld, long long batch_stride, int batch,
                       int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return S3R_ERR_CUDA;
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)batch_stride * 2};
  const cuuint32_t box[3] = {GEMM_BK, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3R_OK : S3R_ERR_CUDA;
}

// Batched variant: `batch` independent GEMMs C_z = opA_z . opB_z^T with element strides between consecutive problems -
// the per-(image, head) contractions of the attention backward (S = Q K^T, dP = dO V^T, dV = P^T dO, dK = dS^T Q,
// dQ = dS K), one launch each.  No bias / activation; flags: S3R_EPI_OUT_F32 only.
extern "C" int s3r_gemm_bf16_batched(const void* A, const void* B, void* C, int32_t M, int32_t N, int32_t K, int32_t lda,
                                     int32_t ldb, int32_t ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c,
                                     int32_t batch, int32_t flags, int32_t a_mn_major, int32_t b_mn_major, void* stream) {
  if (M < 0 || N <= 0 || K <= 0 || batch < 0) return S3R_ERR_INVALID_ARG;
  if (M == 0 || batch == 0) return S3R_OK;
  if (!A || !B || !C || (flags & ~S3R_EPI_OUT_F32)) return S3R_ERR_INVALID_ARG;
  if (lda % 8 || ldb % 8 || ldc % 8 || stride_a % 8 || stride_b % 8 || stride_c % 8 || stride_c <= 0 || batch > 65535)
    return S3R_ERR_UNSUPPORTED;
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return S3R_ERR_UNSUPPORTED;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = a_mn_major ? make_map_3d(&ta, A, K, M, lda, stride_a, batch, 64) : make_map_3d(&ta, A, M, K, lda, stride_a, batch, GEMM_BM)) != S3R_OK) return rc;
  if ((rc = b_mn_major ? make_map_3d(&tb, B, K, N, ldb, stride_b, batch, 64) : make_map_3d(&tb, B, N, K, ldb, stride_b, batch, 128)) != S3R_OK) return rc;
  RopeArgs rope{nullptr, nullptr, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  const ConvArgs cv{};
#define S3R_BGEMM(MAJ_) \
  return launch_gemm<128, 3, false, 1, 1, MAJ_>(ta, tb, nullptr, nullptr, C, M, N, K, ldc, 0, flags, rope, 1, nullptr, nullptr, st, cv, batch, stride_c)
  if (a_mn_major && b_mn_major) S3R_BGEMM(3);
  if (a_mn_major) S3R_BGEMM(1);
  if (b_mn_major) S3R_BGEMM(2);
  S3R_BGEMM(0);
#undef S3R_BGEMM
}

// ---------------------------------------------------------------------------------------------------------- conv2d
// Stride-1 "same" convolution as an implicit GEMM on the same tcgen05 pipeline (DPT heads: heads/dpt_block.py:33-75,
// 121-142,189-218; dpt_head.py:35-70, dpt_gs_head.py:113-157, dpt_gs_sh_head.py:37-74 - cuDNN in the reference).
// tile variant: -1 = auto, 0 = 128x128 tiles (3-stage ring, 2 CTAs/SM), 1 = 128x256 4-stage (1 CTA/SM), 2 = 128x256
// 2-stage (2 CTAs/SM).  Measured on B200 (scripts/dev_conv.py, 3x3 256->256): at 256x256 variant 2 reaches 1071-1174
// TFLOP/s vs 882-980 for variant 0 (the A patch is read once per tap instead of twice); on the small pyramid levels
// (<= 64x64, fewer CTAs than SMs) the 128-wide tile wins because it yields twice as many CTAs.
static int g_conv_variant = -2;

extern "C" int s3r_set_tunable(int32_t key, int32_t value) {
  if (key == S3R_TUNE_GEMM_CLUSTER) {
    const int cm = value / 10, cn = value % 10;
    if (value != 0 && !((cm == 1 || cm == 2) && (cn == 1 || cn == 2 || cn == 4))) return S3R_ERR_INVALID_ARG;
    g_gemm_cluster = value;
    return S3R_OK;
  }
  if (key == S3R_TUNE_GEMM_KSPLIT) {
    if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8 && value != 9) return S3R_ERR_INVALID_ARG;
    g_gemm_ksplit = value;
    return S3R_OK;
  }
  if (key == S3R_TUNE_GEMM_PAIR) {
    if (value < 0 || value > 5) return S3R_ERR_INVALID_ARG;
    g_gemm_pair = value;
    return S3R_OK;
  }
  if (key == S3R_TUNE_ATTN_ONEPASS) {
    if (value < 0 || value > 2) return S3R_ERR_INVALID_ARG;
    s3r_attn_onepass() = value;
    return S3R_OK;
  }
  if (key == S3R_TUNE_GEMM_DIRECT) {
    if (value < 0 || value > 2) return S3R_ERR_INVALID_ARG;
    g_gemm_direct = value;
    return S3R_OK;
  }
  if (key == S3R_TUNE_GEMM_SHALLOW) {
    g_gemm_shallow = value != 0;
    return S3R_OK;
  }
  if (key == S3R_TUNE_PDL) {
    g_pdl = value != 0;
    return S3R_OK;
  }
  if (key == S3R_TUNE_CONV_CLUSTER) {
    g_conv_cluster = value != 0;
    return S3R_OK;
  }
  if (key == S3R_TUNE_GEMM_BIG_TILE) {
    if (value < 0 || value > 2) return S3R_ERR_INVALID_ARG;
    g_gemm_big_tile = value;
    return S3R_OK;
  }
  if (key == S3R_TUNE_RASTER_PDL) {
    s3r_raster_pdl_mask() = value & 31;
    return S3R_OK;
  }
  if (key == S3R_TUNE_BLEND_KERNEL) {
    if (value < 0 || value > 1) return S3R_ERR_INVALID_ARG;
    s3r_blend_kernel_choice() = value;
    return S3R_OK;
  }
  if (key == S3R_TUNE_BWD_ALL) {
    s3r_bwd_mode_override() = value != 0;
    return S3R_OK;
  }
  if (key == S3R_TUNE_BLEND_ONLY_TILE) {
    if (value < 0) return S3R_ERR_INVALID_ARG;
    s3r_blend_only_tile() = value;
    return S3R_OK;
  }
  if (key == S3R_TUNE_CONV_VARIANT) {
    if (value < -1 || value > 2) return S3R_ERR_INVALID_ARG;
    g_conv_variant = value;
    return S3R_OK;
  }
  return S3R_ERR_INVALID_ARG;
}

extern "C" int s3r_conv2d_bf16(const void* x, const void* w, const void* bias, const void* residual, void* y, int32_t n,
                               int32_t h, int32_t wd, int32_t cin, int32_t cout, int32_t kh, int32_t kw, int32_t pad,
                               int32_t flags, void* stream) {
  if (n < 0 || h <= 0 || wd <= 0 || cin <= 0 || cout <= 0 || kh <= 0 || kw <= 0 || pad < 0) return S3R_ERR_INVALID_ARG;
  if (n == 0) return S3R_OK;
  if (!x || !w || !y) return S3R_ERR_INVALID_ARG;
  if ((flags & S3R_EPI_BIAS) && !bias) return S3R_ERR_INVALID_ARG;
  if ((flags & S3R_EPI_RESIDUAL) && !residual) return S3R_ERR_INVALID_ARG;
  if (flags & (S3R_EPI_ROPE | S3R_EPI_GELU)) return S3R_ERR_INVALID_ARG;
  if (2 * pad != kh - 1 || 2 * pad != kw - 1) return S3R_ERR_UNSUPPORTED;            // "same" convolutions only
  if (cin % 8 || cout % 8 || cin < GEMM_BK) return S3R_ERR_UNSUPPORTED;              // TMA pitch / vector epilogue
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) return S3R_ERR_UNSUPPORTED;
  // the 128-pixel M tile must be a (bw x bh x bn) box of the NHWC tensor
  int bw, bh, bn;
  if (wd >= GEMM_BM) {
    if (wd % GEMM_BM) return S3R_ERR_UNSUPPORTED;
    bw = GEMM_BM, bh = 1, bn = 1;
  } else {
    if (GEMM_BM % wd) return S3R_ERR_UNSUPPORTED;
    bw = wd;
    const int rows = GEMM_BM / wd;
    if (rows <= h) {
      if (h % rows) return S3R_ERR_UNSUPPORTED;
      bh = rows, bn = 1;
    } else {
      if (rows % h) return S3R_ERR_UNSUPPORTED;
      bh = h, bn = rows / h;
    }
  }
  const long long M = (long long)n * h * wd;
  if (M > 0x7fffffffLL) return S3R_ERR_UNSUPPORTED;
  ConvArgs conv{h, wd, kw, (cin + GEMM_BK - 1) / GEMM_BK, pad};
  const int K = kh * kw * conv.cblocks * GEMM_BK;  // weights: [cout][kh*kw][cblocks*64], zero padded channels
  if (g_conv_variant == -2) {
    const char* e = getenv("S3R_CONV_VARIANT");
    g_conv_variant = e ? atoi(e) : -1;
  }
  int variant = g_conv_variant;
  if (variant < 0) variant = (cout % 256 == 0 && (M + GEMM_BM - 1) / GEMM_BM >= 256) ? 2 : 0;
  const int BN = cout <= 64 ? 64 : ((variant > 0 && cout % 256 == 0) ? 256 : 128);
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_map_nhwc(&ta, x, n, h, wd, cin, bw, bh, bn)) != S3R_OK) return rc;
  // clusters of 2 adjacent pixel tiles share the weight tile: each loads half of it and multicasts (the weights are the
  // operand every M tile re-reads; the activation patch of a 128x256 tile is already read once per tap)
  const int cm = (g_conv_cluster && (M + GEMM_BM - 1) / GEMM_BM >= 2 && BN >= 128) ? 2 : 1;
  if ((rc = make_map(&tb, w, cout, K, K, BN / cm)) != S3R_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  RopeArgs rope{nullptr, nullptr, 0, 0};
  {  // CTA pairs (cta_group::2): two adjacent pixel tiles run one 256-row MMA, each CTA stages half of the weight tile
    int pair_bn = 0;
    const bool cm_forced = g_conv_cluster != 0;
    const long mtiles = (M + GEMM_BM - 1) / GEMM_BM;
    if (g_gemm_pair == 1 && mtiles >= 2 && cout % 128 == 0) pair_bn = 128;
    if (g_gemm_pair == 3 && mtiles >= 2 && cout % 256 == 0) pair_bn = 256;
    // measured: 3x3 256->256 at 16 x 128^2 / 4 x 256^2: 1189 -> 1231 TFLOP/s, at 2 x 128^2 1005 -> 1026; below that the grid is
    // smaller than the machine and the 128-wide single-CTA tile (twice the CTAs) wins

    constexpr int E_CONVP = S3R_EPI_BIAS | S3R_EPI_OUT_F32 | S3R_EPI_PDL | S3R_EPI_RESIDUAL | S3R_EPI_RELU;
    // auto (measured, 3x3 convolutions): 256->256 at 16 x 128^2 1173 -> 1446 TFLOP/s (cuDNN bf16 NHWC 1185-1289), 16 x 64^2
    // 1100 -> 1388, 2 x 128^2 1068 -> 1262, 2 x 64^2 499 -> 535; 128->128 at 2 x 256^2 693 -> 824, 16 x 256^2 794 -> 910
    const bool auto_p = g_gemm_pair == 0 && g_conv_variant < 0 && !cm_forced &&
                        ((cout % 256 == 0 && mtiles >= 64) || (cout % 128 == 0 && mtiles >= 256));
    if ((g_gemm_pair == 4 || g_gemm_pair == 5 || auto_p) && mtiles >= 2 && !(flags & ~E_CONVP) && !(((uintptr_t)bias | (uintptr_t)residual) & 15)) {
      const int pbn = ((g_gemm_pair == 5 || auto_p) && cout % 256 == 0) ? 256 : (cout % 128 == 0 ? 128 : 0);
      if (pbn) {
        if ((rc = make_map(&tb, w, cout, K, K, pbn / 2)) != S3R_OK) return rc;
        if (pbn == 128)
          return launch_gemm_persist<128, 8, true, E_CONVP>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, st, conv);
        return launch_gemm_persist<256, 6, true, E_CONVP>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, st, conv);
      }
    }
    if (pair_bn) {
      if ((rc = make_map(&tb, w, cout, K, K, pair_bn / 2)) != S3R_OK) return rc;
      if (pair_bn == 128)
        return launch_gemm<128, 4, true, 1, 1, 0, S3R_EPI_ALL, 1, 1>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
      return launch_gemm<256, 3, true, 1, 1, 0, S3R_EPI_ALL, 1, 1>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
    }
  }
  if (cm == 2) {
    if (BN == 256 && variant == 1)
      return launch_gemm<256, 4, true, 2, 1>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
    if (BN == 256)
      return launch_gemm<256, 2, true, 2, 1>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
    return launch_gemm<128, 3, true, 2, 1>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
  }
  // small pyramid levels (8^2 .. 32^2 at batch 1-2: 2-32 CTAs under a 36-k-block loop): cluster split-K over the taps
  if (g_gemm_ksplit != 1 && BN == 128 && cm == 1) {
    const long ctas = ((M + GEMM_BM - 1) / GEMM_BM) * ((cout + 127) / 128);
    const int total_kb = K / GEMM_BK;
    if (total_kb >= 16 && ctas * 4 <= 160)
      return launch_gemm<128, 3, true, 1, 1, 0, S3R_EPI_ALL, 4>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
    if (total_kb >= 16 && ctas * 2 <= 200)
      return launch_gemm<128, 3, true, 1, 1, 0, S3R_EPI_ALL, 2>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
  }
  if (BN == 64) return launch_gemm<64, 4, true>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
  if (BN == 256 && variant == 1)
    return launch_gemm<256, 4, true>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
  if (BN == 256)
    return launch_gemm<256, 2, true>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
  return launch_gemm<128, 3, true>(ta, tb, bias, residual, y, (int)M, cout, K, cout, cout, flags, rope, 1, nullptr, nullptr, st, conv);
}
