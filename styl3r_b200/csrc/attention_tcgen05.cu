// Attention (softmax(q k^T * scale) v, no mask, no dropout) on tcgen05 / TMEM / TMA for head_dim 64 — the
// contraction the reference delegates to xformers.ops.memory_efficient_attention
// (src/model/encoder/backbone/croco/blocks.py:126-130,192-196) on [B, N, H, 64] tensors.
//
// Two kernels (S3R_TUNE_ATTN_ONEPASS): the default ONE-PASS kernel further down (s3r_attention1_kernel) and the original
// two-pass kernel described here.  Shapes in Styl3R are short (Nq, Nk in {256, 257, 514, 771, 1028}); the two-pass kernel
// favours simplicity over an online softmax: TWO passes over the keys per 128-query tile — pass 1 finds the exact row maxima (S = Q K^T only), pass 2
// recomputes S tile by tile, forms P = exp2((S - max) * scale * log2e) and accumulates O += P V.  No running-max
// correction of O is ever needed; the extra Q K^T costs one third more tensor work, which is negligible here.
//
// One CTA = one (batch, head, 128-query tile); 10 warps, warp-specialised like the GEMM:
//   warp 0    TMA producer: Q tile (128x64) once, then K tiles (64x64) for both passes and V tiles (64x64) for pass 2
//             through 4-stage rings (4-D tensor maps over the strided [B,N,H,64] views, SWIZZLE_128B)
//   warp 1    TMEM allocator (256 columns: S 2 x 64 double-buffered + O 64) and MMA issuer:
//               S = Q K_j^T : 4 x tcgen05.mma M128 N64 K16, both operands K-major
//               O += P V_j  : 4 x tcgen05.mma M128 N64 K16, A = P (K-major, written by the softmax warps),
//                             B = V_j (MN-major: head_dim contiguous)
//   warps 2-9 softmax / epilogue: TWO warps per TMEM lane quarter (warp % 4), each thread owns one query row and one
//             32-column half of the 64-key tile: tcgen05.ld S (the S buffer is released as soon as it is in registers),
//             mask the key tail, row max (pass 1) or exp2 / row sum / bf16 P into swizzled shared memory (pass 2),
//             finally O / l -> bf16.  Partial row maxima / sums of the two halves meet once per pass in shared memory.
// Software pipeline: the MMA warp issues S(j+1) = Q K_{j+1}^T *before* it waits for P(j), so the tensor core and the
// K/V TMA latency overlap the softmax of tile j, and O += P(j) V_j overlaps the first half of the softmax of tile j+1.
// The kernel is bound by the softmax warps (8192 exp2 per tile on the MUFU pipe), hence the 8 warps.
#include <cuda.h>
#include <cuda_bf16.h>

#include "s3r_common.cuh"

#define ATT_BM 128
#define ATT_BN 64
#define ATT_D 64
#define ATT_THREADS 320
#define ATT_SM_WARPS 8
#define ATT_KV_STAGES 4  // depth of the K and V TMA rings: a 64-key tile is consumed in ~0.2 us, a TMA round trip takes ~0.8 us

__device__ __forceinline__ uint32_t a_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void a_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void a_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void a_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void a_tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(a_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(a_smem_u32(bar))
      : "memory");
}
// SWIZZLE_128B descriptor, 8-row groups 1024 B apart (valid for K-major tiles and for the MN-major 64-wide V tile)
__device__ __forceinline__ uint64_t a_make_desc(const void* smem_ptr) {
  const uint32_t addr = a_smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t a_make_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void a_umma(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void a_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void a_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

__device__ __forceinline__ void a_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
      "%25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float a_ex2(float x) {  // MUFU.EX2 (flush-to-zero)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttSmem {  // offsets from a 1024-B aligned base
  static constexpr int Q = 0;                        // 128 x 64 bf16 = 16 KB
  static constexpr int K = Q + 16384;                // ATT_KV_STAGES x (64 x 64 bf16 = 8 KB)
  static constexpr int V = K + ATT_KV_STAGES * 8192; // ATT_KV_STAGES x 8 KB
  static constexpr int P = V + ATT_KV_STAGES * 8192; // 128 x 64 bf16 = 16 KB
  static constexpr int BARS = P + 16384;
  static constexpr int RED = BARS + 256;             // [2][2][128] floats: row max / row sum partials of the column halves
  static constexpr int TOTAL = RED + 2048 + 1024;
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
s3r_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ O, int H, int Nq, int Nk,
                     long long so_b, long long so_n, long long so_h, float scale_log2e, int pdl) {
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)att_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + AttSmem::BARS);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                        // [ATT_KV_STAGES]
  uint64_t* k_empty = k_full + ATT_KV_STAGES;         // [ATT_KV_STAGES]
  uint64_t* v_full = k_empty + ATT_KV_STAGES;         // [ATT_KV_STAGES]
  uint64_t* v_empty = v_full + ATT_KV_STAGES;         // [ATT_KV_STAGES]
  uint64_t* s_full = v_empty + ATT_KV_STAGES;         // [2]: S is double-buffered in TMEM
  uint64_t* s_empty = s_full + 2;                     // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint32_t* tmem_slot = (uint32_t*)(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (Nk + ATT_BN - 1) / ATT_BN;
  const int T = 2 * nkv;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    a_mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KV_STAGES; s++) {
      a_mbar_init(&k_full[s], 1);
      a_mbar_init(&k_empty[s], 1);
      a_mbar_init(&v_full[s], 1);
      a_mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      a_mbar_init(&s_full[s], 1);
      a_mbar_init(&s_empty[s], ATT_SM_WARPS);  // one elected arrival per softmax warp
    }
    a_mbar_init(p_full, ATT_SM_WARPS);
    a_mbar_init(p_empty, 1);
    a_mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;  // S0 [0,64), S1 [64,128), O [128,192)
  if (pdl) {  // programmatic dependent launch: the prologue above overlapped the previous kernel's tail (see gemm_tcgen05.cu)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  if (warp == 0) {
    if (lane == 0) {
      a_mbar_expect_tx(q_full, 16384);
      a_tma_load_4d(smem + AttSmem::Q, &tmQ, 0, h, q0, b, q_full);
      for (int it = 0; it < T; it++) {
        const int s = it % ATT_KV_STAGES, j = it % nkv;
        a_mbar_wait(&k_empty[s], ((it / ATT_KV_STAGES) & 1) ^ 1);
        a_mbar_expect_tx(&k_full[s], 8192);
        a_tma_load_4d(smem + AttSmem::K + s * 8192, &tmK, 0, h, j * ATT_BN, b, &k_full[s]);
        if (it >= nkv) {
          const int jj = it - nkv, sv = jj % ATT_KV_STAGES;
          a_mbar_wait(&v_empty[sv], ((jj / ATT_KV_STAGES) & 1) ^ 1);
          a_mbar_expect_tx(&v_full[sv], 8192);
          a_tma_load_4d(smem + AttSmem::V + sv * 8192, &tmV, 0, h, jj * ATT_BN, b, &v_full[sv]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = a_make_idesc(ATT_BM, ATT_BN, 0);
      const uint32_t idesc_o = a_make_idesc(ATT_BM, ATT_D, 1);
      const uint64_t qdesc = a_make_desc(smem + AttSmem::Q);
      const uint64_t pdesc = a_make_desc(smem + AttSmem::P);
      a_mbar_wait(q_full, 0);
      auto issue_pv = [&](int jj) {  // O += P(jj) V_jj
        const int sv = jj % ATT_KV_STAGES;
        a_mbar_wait(&v_full[sv], (jj / ATT_KV_STAGES) & 1);
        a_mbar_wait(p_full, jj & 1);  // P tile jj written
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t vdesc = a_make_desc(smem + AttSmem::V + sv * 8192);
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; k++)  // 16 key rows = 2048 B of the MN-major V tile per k-step
          a_umma(tmem_O, pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(128 * k), idesc_o, (jj | k) ? 1u : 0u);
        a_umma_commit(&v_empty[sv]);
        a_umma_commit(p_empty);
      };
      for (int it = 0; it < T; it++) {
        const int s = it & 1, ks = it % ATT_KV_STAGES;
        a_mbar_wait(&k_full[ks], (it / ATT_KV_STAGES) & 1);
        a_mbar_wait(&s_empty[s], ((it >> 1) & 1) ^ 1);  // the softmax warps hold S(it-2) in registers
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t kdesc = a_make_desc(smem + AttSmem::K + ks * 8192);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; k++)
          a_umma(tmem_S + (uint32_t)(s * 64), qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
        a_umma_commit(&k_empty[ks]);
        a_umma_commit(&s_full[s]);
        if (it - 1 >= nkv) issue_pv(it - 1 - nkv);  // one tile behind: S(it) runs while the softmax warps make P(it-1)
      }
      issue_pv(nkv - 1);
      a_umma_commit(o_full);
    }
  } else {
    // ===== softmax / epilogue: thread <-> (query row, 32-column half); TMEM lane quarter = warp % 4
    const int qd = warp & 3;
    const int ch = (warp - 2) >> 2;            // column half of the 64-key tile: warps 2-5 -> 0, warps 6-9 -> 1
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(qd * 32) << 16) + (uint32_t)(ch * 32);
    float* red = reinterpret_cast<float*>(smem + AttSmem::RED);   // [2][128]
    float m = -INFINITY, l = 0.f;
    float mb = 0.f;
    uint8_t* prow = smem + AttSmem::P + row * 128;
    for (int it = 0; it < T; it++) {
      const int j = it % nkv;
      a_mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v0[32];
      a_tmem_ld32(tmem_S + (uint32_t)((it & 1) * 64) + lane_addr, v0);
      // S is in registers: hand the buffer back so that the next Q K^T overlaps this tile's softmax
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) a_mbar_arrive(&s_empty[it & 1]);
      const int valid = Nk - j * ATT_BN - ch * 32;  // my columns >= valid are the zero-filled key tail
      if (it < nkv) {
        // pass 1: exact row maximum
        if (valid >= 32) {  // full tile (warp-uniform): no tail predicates
#pragma unroll
          for (int c = 0; c < 32; c++) m = fmaxf(m, __uint_as_float(v0[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 32; c++)
            if (c < valid) m = fmaxf(m, __uint_as_float(v0[c]));
        }
        if (it == nkv - 1) {  // combine the two column halves once
          red[ch * 128 + row] = m;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          m = fmaxf(m, red[(ch ^ 1) * 128 + row]);
          mb = m * scale_log2e;
          asm volatile("bar.sync 1, 256;" ::: "memory");  // both halves have read before `red` is reused for the sums
        }
      } else {
        // pass 2: P = exp2((S - m) * scale * log2e), row sum, bf16 P -> swizzled shared memory
        const int jj = it - nkv;
        float p[32];
        if (valid >= 32) {  // full tile (warp-uniform): FFMA + MUFU.EX2 per element, no tail predicates
#pragma unroll
          for (int c = 0; c < 32; c++) p[c] = a_ex2(fmaf(__uint_as_float(v0[c]), scale_log2e, -mb));
        } else {
#pragma unroll
          for (int c = 0; c < 32; c++) p[c] = c < valid ? a_ex2(fmaf(__uint_as_float(v0[c]), scale_log2e, -mb)) : 0.f;
        }
        // row sum in fp32 from the unrounded probabilities, four independent chains (the bf16 rounding of P is unbiased:
        // the sum of the rounded values differs by ~2^-9 / sqrt(n), far below the bf16 rounding of the output)
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
        for (int c = 0; c < 32; c += 4) l0 += p[c], l1 += p[c + 1], l2 += p[c + 2], l3 += p[c + 3];
        l += (l0 + l1) + (l2 + l3);
        a_mbar_wait(p_empty, (jj & 1) ^ 1);  // the previous P V MMA has finished reading the P buffer
#pragma unroll
        for (int c16 = 0; c16 < 4; c16++) {
          uint4 u;
          __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int t = 0; t < 4; t++) hh[t] = __floats2bfloat162_rn(p[c16 * 8 + 2 * t], p[c16 * 8 + 2 * t + 1]);
          *reinterpret_cast<uint4*>(prow + (((ch * 4 + c16) ^ (row & 7)) << 4)) = u;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
        __syncwarp();
        if (lane == 0) a_mbar_arrive(p_full);
      }
    }
    // row sums of the two halves
    red[ch * 128 + row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += red[(ch ^ 1) * 128 + row];
    // epilogue: my 32 of the 64 output columns
    a_mbar_wait(o_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t o0[32];
    a_tmem_ld32(tmem_O + lane_addr, o0);
    const int qrow = q0 + row;
    if (qrow < Nq) {
      const float inv = 1.0f / l;
      __nv_bfloat16* op = O + (long long)b * so_b + (long long)qrow * so_n + (long long)h * so_h + ch * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        uint4 u;
        __nv_bfloat162* hu = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int t = 0; t < 4; t++)
          hu[t] = __floats2bfloat162_rn(__uint_as_float(o0[c + 2 * t]) * inv, __uint_as_float(o0[c + 2 * t + 1]) * inv);
        *reinterpret_cast<uint4*>(op + c) = u;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ONE-PASS variant (longer key sequences: the 771 / 1028-key cross-view and stylizer attentions of v >= 3): every K tile is
// loaded and contracted once.  Instead of an online softmax with a per-tile exchange between the two threads that share a
// query row (column halves of the 64-key tile), EACH HALF IS ITS OWN SOFTMAX STREAM: it keeps its own reference maximum and
// row sum and accumulates into its own O accumulator in TMEM (O_a: keys 0-31 of every tile = k-steps 0,1 of the P V MMA;
// O_b: keys 32-63 = k-steps 2,3), so no cross-warp traffic exists inside the loop; the two streams are merged once in
// the epilogue: O = (O_a 2^(m_a-m) + O_b 2^(m_b-m)) / (l_a 2^(m_a-m) + l_b 2^(m_b-m)).
// The reference maximum is LAZY: P = exp2(S c - m_ref) may exceed 1 (bf16 / fp32 share the exponent range); only when a
// tile's maximum exceeds m_ref by more than 2^6 does the thread rescale its O row in TMEM (tcgen05.ld / st) - between
// the p_empty wait (P V of the previous tile has completed) and its p_full arrival (the next P V cannot start), i.e. with
// exclusive access and no extra synchronisation.  The result is the exact softmax up to rounding.
__global__ void __launch_bounds__(ATT_THREADS, 2)
s3r_attention1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ O, int H, int Nq, int Nk,
                      long long so_b, long long so_n, long long so_h, float scale_log2e, int pdl) {
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)att_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + AttSmem::BARS);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + ATT_KV_STAGES;
  uint64_t* v_full = k_empty + ATT_KV_STAGES;
  uint64_t* v_empty = v_full + ATT_KV_STAGES;
  uint64_t* s_full = v_empty + ATT_KV_STAGES;  // [2]
  uint64_t* s_empty = s_full + 2;              // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint32_t* tmem_slot = (uint32_t*)(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (Nk + ATT_BN - 1) / ATT_BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    a_mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KV_STAGES; s++) {
      a_mbar_init(&k_full[s], 1);
      a_mbar_init(&k_empty[s], 1);
      a_mbar_init(&v_full[s], 1);
      a_mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      a_mbar_init(&s_full[s], 1);
      a_mbar_init(&s_empty[s], ATT_SM_WARPS);
    }
    a_mbar_init(p_full, ATT_SM_WARPS);
    a_mbar_init(p_empty, 1);
    a_mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_Oa = tmem_base + 128, tmem_Ob = tmem_base + 192;  // S0, S1, O_a, O_b: 64 columns each
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  if (warp == 0) {
    if (lane == 0) {
      a_mbar_expect_tx(q_full, 16384);
      a_tma_load_4d(smem + AttSmem::Q, &tmQ, 0, h, q0, b, q_full);
      for (int it = 0; it < nkv; it++) {
        const int s = it % ATT_KV_STAGES;
        a_mbar_wait(&k_empty[s], ((it / ATT_KV_STAGES) & 1) ^ 1);
        a_mbar_expect_tx(&k_full[s], 8192);
        a_tma_load_4d(smem + AttSmem::K + s * 8192, &tmK, 0, h, it * ATT_BN, b, &k_full[s]);
        a_mbar_wait(&v_empty[s], ((it / ATT_KV_STAGES) & 1) ^ 1);
        a_mbar_expect_tx(&v_full[s], 8192);
        a_tma_load_4d(smem + AttSmem::V + s * 8192, &tmV, 0, h, it * ATT_BN, b, &v_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = a_make_idesc(ATT_BM, ATT_BN, 0);
      const uint32_t idesc_o = a_make_idesc(ATT_BM, ATT_D, 1);
      const uint64_t qdesc = a_make_desc(smem + AttSmem::Q);
      const uint64_t pdesc = a_make_desc(smem + AttSmem::P);
      a_mbar_wait(q_full, 0);
      auto issue_pv = [&](int jj) {  // O_a += P(jj)[:, 0:32] V_jj[0:32], O_b += P(jj)[:, 32:64] V_jj[32:64]
        const int sv = jj % ATT_KV_STAGES;
        a_mbar_wait(&v_full[sv], (jj / ATT_KV_STAGES) & 1);
        a_mbar_wait(p_full, jj & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t vdesc = a_make_desc(smem + AttSmem::V + sv * 8192);
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; k++)
          a_umma(k < 2 ? tmem_Oa : tmem_Ob, pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(128 * k), idesc_o,
                 (jj | (k & 1)) ? 1u : 0u);
        a_umma_commit(&v_empty[sv]);
        a_umma_commit(p_empty);
      };
      for (int it = 0; it < nkv; it++) {
        const int s = it & 1, ks = it % ATT_KV_STAGES;
        a_mbar_wait(&k_full[ks], (it / ATT_KV_STAGES) & 1);
        a_mbar_wait(&s_empty[s], ((it >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t kdesc = a_make_desc(smem + AttSmem::K + ks * 8192);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; k++)
          a_umma(tmem_S + (uint32_t)(s * 64), qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
        a_umma_commit(&k_empty[ks]);
        a_umma_commit(&s_full[s]);
        if (it >= 1) issue_pv(it - 1);  // one tile behind: S(it) runs while the softmax warps make P(it-1)
      }
      issue_pv(nkv - 1);
      a_umma_commit(o_full);
    }
  } else {
    const int qd = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int row = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const uint32_t tmem_Omine = ch ? tmem_Ob : tmem_Oa;
    float* red = reinterpret_cast<float*>(smem + AttSmem::RED);   // [2][2][128]: (m_ref, l) of the two halves
    float mref = -INFINITY;  // reference maximum in the scaled log2 domain (S * scale * log2e)
    float l = 0.f;
    uint8_t* prow = smem + AttSmem::P + row * 128;
    for (int it = 0; it < nkv; it++) {
      a_mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v0[32];
      a_tmem_ld32(tmem_S + (uint32_t)((it & 1) * 64) + lane_base + (uint32_t)(ch * 32), v0);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) a_mbar_arrive(&s_empty[it & 1]);
      const int valid = Nk - it * ATT_BN - ch * 32;  // my columns >= valid are the zero-filled key tail (warp-uniform)
      float p[32];
      float factor = 1.0f;     // what the accumulated l / O row must be multiplied by (!= 1: the reference moved)
      bool rescale = false;
      if (valid > 0) {
        float tm = -INFINITY;
        if (valid >= 32) {
#pragma unroll
          for (int c = 0; c < 32; c++) tm = fmaxf(tm, __uint_as_float(v0[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 32; c++)
            if (c < valid) tm = fmaxf(tm, __uint_as_float(v0[c]));
        }
        const float ts = tm * scale_log2e;
        if (ts > mref + 6.0f) {  // (always true for the first valid tile: mref = -inf)
          factor = a_ex2(mref - ts);  // exp2(-inf) = 0 for the first tile
          rescale = l > 0.f;          // nothing accumulated yet -> the O row needs no correction
          mref = ts;
        }
        if (valid >= 32) {
#pragma unroll
          for (int c = 0; c < 32; c++) p[c] = a_ex2(fmaf(__uint_as_float(v0[c]), scale_log2e, -mref));
        } else {
#pragma unroll
          for (int c = 0; c < 32; c++) p[c] = c < valid ? a_ex2(fmaf(__uint_as_float(v0[c]), scale_log2e, -mref)) : 0.f;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; c++) p[c] = 0.f;
      }
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c += 4) l0 += p[c], l1 += p[c + 1], l2 += p[c + 2], l3 += p[c + 3];
      l = l * factor + ((l0 + l1) + (l2 + l3));
      a_mbar_wait(p_empty, (it & 1) ^ 1);  // P V of the previous tile has completed: P buffer free, O rows stable
      if (__any_sync(0xffffffffu, rescale)) {
        // rare (the running maximum settles within the first tiles): scale this thread's row of its own accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float f = rescale ? factor : 1.0f;
#pragma unroll 1
        for (int hcol = 0; hcol < 2; hcol++) {
          uint32_t o[32];
          a_tmem_ld32(tmem_Omine + lane_base + (uint32_t)(hcol * 32), o);
#pragma unroll
          for (int c = 0; c < 32; c++) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
          a_tmem_st32(tmem_Omine + lane_base + (uint32_t)(hcol * 32), o);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
#pragma unroll
      for (int c16 = 0; c16 < 4; c16++) {
        uint4 u;
        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int t = 0; t < 4; t++) hh[t] = __floats2bfloat162_rn(p[c16 * 8 + 2 * t], p[c16 * 8 + 2 * t + 1]);
        *reinterpret_cast<uint4*>(prow + (((ch * 4 + c16) ^ (row & 7)) << 4)) = u;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) a_mbar_arrive(p_full);
    }
    // merge the two softmax streams of a row
    red[(ch * 2 + 0) * 128 + row] = mref;
    red[(ch * 2 + 1) * 128 + row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float mo = red[((ch ^ 1) * 2 + 0) * 128 + row], lo = red[((ch ^ 1) * 2 + 1) * 128 + row];
    const float m = fmaxf(mref, mo);
    const float w_me = a_ex2(mref - m), w_ot = a_ex2(mo - m);  // exp2(-inf) = 0: a stream that never saw a valid key
    const float inv = 1.0f / (l * w_me + lo * w_ot);
    const float wa = (ch ? w_ot : w_me) * inv, wb = (ch ? w_me : w_ot) * inv;
    a_mbar_wait(o_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t oa[32], ob[32];
    a_tmem_ld32(tmem_Oa + lane_base + (uint32_t)(ch * 32), oa);
    a_tmem_ld32(tmem_Ob + lane_base + (uint32_t)(ch * 32), ob);
    const int qrow = q0 + row;
    if (qrow < Nq) {
      __nv_bfloat16* op = O + (long long)b * so_b + (long long)qrow * so_n + (long long)h * so_h + ch * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        uint4 u;
        __nv_bfloat162* hu = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int t = 0; t < 4; t++)
          hu[t] = __floats2bfloat162_rn(__uint_as_float(oa[c + 2 * t]) * wa + __uint_as_float(ob[c + 2 * t]) * wb,
                                        __uint_as_float(oa[c + 2 * t + 1]) * wa + __uint_as_float(ob[c + 2 * t + 1]) * wb);
        *reinterpret_cast<uint4*>(op + c) = u;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// S3R_TUNE_ATTN_ONEPASS: 0 / 1 = one pass (default), 2 = the two-pass kernel
static int g_attn_onepass = 0;
int& s3r_attn_onepass() { return g_attn_onepass; }

typedef CUresult (*PFN_encodeTiledA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledA att_get_encode() {
  static PFN_encodeTiledA fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (PFN_encodeTiledA)p;
  }
  return fn;
}
// [B, N, H, 64] bf16 view with element strides (sb, sn, sh, 1): dims (64, H, N, B), box (64, 1, rows, 1)
static int att_make_map(CUtensorMap* map, const void* ptr, int B, int N, int H, long long sb, long long sn, long long sh,
                        int box_rows) {
  PFN_encodeTiledA enc = att_get_encode();
  if (!enc) return S3R_ERR_CUDA;
  const cuuint64_t dims[4] = {ATT_D, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)sn * 2, (cuuint64_t)sb * 2};
  const cuuint32_t box[4] = {ATT_D, 1, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3R_OK : S3R_ERR_CUDA;
}

extern "C" int s3r_attention_bf16(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t H, int32_t Nq,
                                  int32_t Nk, int32_t D, const int64_t* q_strides, const int64_t* k_strides,
                                  const int64_t* v_strides, const int64_t* o_strides, float scale, void* stream) {
  if (B < 0 || H <= 0 || Nq < 0 || Nk <= 0) return S3R_ERR_INVALID_ARG;
  if (D != ATT_D) return S3R_ERR_UNSUPPORTED;
  if (B == 0 || Nq == 0) return S3R_OK;
  if (!q || !k || !v || !o || !q_strides || !k_strides || !v_strides || !o_strides) return S3R_ERR_INVALID_ARG;
  for (int i = 0; i < 3; i++)
    if (q_strides[i] % 8 || k_strides[i] % 8 || v_strides[i] % 8 || o_strides[i] % 8) return S3R_ERR_UNSUPPORTED;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) & 15) return S3R_ERR_UNSUPPORTED;
  if (H > 65535 || B > 65535) return S3R_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = att_make_map(&tq, q, B, Nq, H, q_strides[0], q_strides[1], q_strides[2], ATT_BM)) != S3R_OK) return rc;
  if ((rc = att_make_map(&tk, k, B, Nk, H, k_strides[0], k_strides[1], k_strides[2], ATT_BN)) != S3R_OK) return rc;
  if ((rc = att_make_map(&tv, v, B, Nk, H, v_strides[0], v_strides[1], v_strides[2], ATT_BN)) != S3R_OK) return rc;
  static size_t configured[64] = {}, configured1[64] = {};  // per device (cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute)
  if ((rc = s3r_ensure_dynamic_smem(s3r_attention_kernel, (size_t)AttSmem::TOTAL, configured)) != S3R_OK) return rc;
  if ((rc = s3r_ensure_dynamic_smem(s3r_attention1_kernel, (size_t)AttSmem::TOTAL, configured1)) != S3R_OK) return rc;
  const int nkv_tiles = (Nk + ATT_BN - 1) / ATT_BN;
  // measured on B200 (scripts/bench_attn.py): one pass wins on every encoder shape (257^2: 9.4 -> 7.8 us, 1028^2: 45.0 -> 35.1)
  const bool one_pass = g_attn_onepass != 2;
  (void)nkv_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((Nq + ATT_BM - 1) / ATT_BM, H, B);
  cfg.blockDim = dim3(ATT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = AttSmem::TOTAL;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  const int pdl = s3r_pdl_enabled();
  if (pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  S3R_CUDA_CHECK(cudaLaunchKernelEx(&cfg, one_pass ? s3r_attention1_kernel : s3r_attention_kernel, tq, tk, tv, (__nv_bfloat16*)o, (int)H, (int)Nq, (int)Nk,
                                    (long long)o_strides[0], (long long)o_strides[1], (long long)o_strides[2],
                                    scale * 1.4426950408889634f, pdl));
  return S3R_OK;
}
