// Attention (softmax(q k^T * scale) v, no mask, no dropout) on tcgen05 / TMEM / TMA for head_dim 64 — the
// contraction the reference delegates to xformers.ops.memory_efficient_attention
// (src/model/encoder/backbone/croco/blocks.py:126-130,192-196) on [B, N, H, 64] tensors.
//
// Shapes in Styl3R are short (Nq, Nk in {256, 257, 514, 771, 1028}) so the kernel favours simplicity over an online
// softmax: TWO passes over the keys per 128-query tile — pass 1 finds the exact row maxima (S = Q K^T only), pass 2
// recomputes S tile by tile, forms P = exp2((S - max) * scale * log2e) and accumulates O += P V.  No running-max
// correction of O is ever needed; the extra Q K^T costs one third more tensor work, which is negligible here.
//
// One CTA = one (batch, head, 128-query tile); 6 warps, warp-specialised like the GEMM:
//   warp 0    TMA producer: Q tile (128x64) once, then K tiles (64x64) for both passes and V tiles (64x64) for pass 2
//             through 2-stage rings (4-D tensor maps over the strided [B,N,H,64] views, SWIZZLE_128B)
//   warp 1    TMEM allocator (128 columns: S 64 + O 64) and MMA issuer:
//               S = Q K_j^T : 4 x tcgen05.mma M128 N64 K16, both operands K-major
//               O += P V_j  : 4 x tcgen05.mma M128 N64 K16, A = P (K-major, written by the softmax warps),
//                             B = V_j (MN-major: head_dim contiguous)
//   warps 2-5 softmax / epilogue, one query row per thread (= one TMEM lane): tcgen05.ld S, mask the key tail, row max
//             (pass 1) or exp2 / row sum / bf16 P into swizzled shared memory (pass 2), finally O / l -> bf16.
#include <cuda.h>
#include <cuda_bf16.h>

#include "s3r_common.cuh"

#define ATT_BM 128
#define ATT_BN 64
#define ATT_D 64
#define ATT_THREADS 192

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

struct AttSmem {  // offsets from a 1024-B aligned base
  static constexpr int Q = 0;                        // 128 x 64 bf16 = 16 KB
  static constexpr int K = Q + 16384;                // 2 x (64 x 64 bf16 = 8 KB)
  static constexpr int V = K + 2 * 8192;             // 2 x 8 KB
  static constexpr int P = V + 2 * 8192;             // 128 x 64 bf16 = 16 KB
  static constexpr int BARS = P + 16384;
  static constexpr int TOTAL = BARS + 256 + 1024;
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
s3r_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ O, int H, int Nq, int Nk,
                     long long so_b, long long so_n, long long so_h, float scale_log2e) {
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)att_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + AttSmem::BARS);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_empty = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* p_empty = bars + 12;
  uint64_t* o_full = bars + 13;
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (Nk + ATT_BN - 1) / ATT_BN;
  const int T = 2 * nkv;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    a_mbar_init(q_full, 1);
    for (int s = 0; s < 2; s++) {
      a_mbar_init(&k_full[s], 1);
      a_mbar_init(&k_empty[s], 1);
      a_mbar_init(&v_full[s], 1);
      a_mbar_init(&v_empty[s], 1);
    }
    a_mbar_init(s_full, 1);
    a_mbar_init(s_empty, 4);  // one elected arrival per softmax warp
    a_mbar_init(p_full, 4);
    a_mbar_init(p_empty, 1);
    a_mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(tmem_slot)), "r"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 64;

  if (warp == 0) {
    if (lane == 0) {
      a_mbar_expect_tx(q_full, 16384);
      a_tma_load_4d(smem + AttSmem::Q, &tmQ, 0, h, q0, b, q_full);
      for (int it = 0; it < T; it++) {
        const int s = it & 1, j = it % nkv;
        a_mbar_wait(&k_empty[s], ((it >> 1) & 1) ^ 1);
        a_mbar_expect_tx(&k_full[s], 8192);
        a_tma_load_4d(smem + AttSmem::K + s * 8192, &tmK, 0, h, j * ATT_BN, b, &k_full[s]);
        if (it >= nkv) {
          const int jj = it - nkv, sv = jj & 1;
          a_mbar_wait(&v_empty[sv], ((jj >> 1) & 1) ^ 1);
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
      for (int it = 0; it < T; it++) {
        const int s = it & 1;
        a_mbar_wait(&k_full[s], (it >> 1) & 1);
        a_mbar_wait(s_empty, (it & 1) ^ 1);  // softmax warps are done reading the previous S
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t kdesc = a_make_desc(smem + AttSmem::K + s * 8192);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; k++) a_umma(tmem_S, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
        a_umma_commit(&k_empty[s]);
        a_umma_commit(s_full);
        if (it >= nkv) {
          const int jj = it - nkv, sv = jj & 1;
          a_mbar_wait(&v_full[sv], (jj >> 1) & 1);
          a_mbar_wait(p_full, jj & 1);  // P tile jj written (and S tile consumed)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t vdesc = a_make_desc(smem + AttSmem::V + sv * 8192);
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; k++)  // 16 key rows = 2048 B of the MN-major V tile per k-step
            a_umma(tmem_O, pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(128 * k), idesc_o, (jj | k) ? 1u : 0u);
          a_umma_commit(&v_empty[sv]);
          a_umma_commit(p_empty);
        }
      }
      a_umma_commit(o_full);
    }
  } else {
    // ===== softmax / epilogue: thread <-> query row (TMEM lane quarter = warp % 4)
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = smem + AttSmem::P + row * 128;
    for (int it = 0; it < T; it++) {
      const int j = it % nkv;
      a_mbar_wait(s_full, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v0[32], v1[32];
      a_tmem_ld32(tmem_S + lane_addr, v0);
      a_tmem_ld32(tmem_S + lane_addr + 32, v1);
      const int valid = Nk - j * ATT_BN;  // columns >= valid are the zero-filled key tail
      if (it < nkv) {
        // pass 1: exact row maximum
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) a_mbar_arrive(s_empty);
#pragma unroll
        for (int c = 0; c < 32; c++) {
          if (c < valid) m = fmaxf(m, __uint_as_float(v0[c]));
          if (c + 32 < valid) m = fmaxf(m, __uint_as_float(v1[c]));
        }
      } else {
        // pass 2: P = exp2((S - m) * scale * log2e), row sum, bf16 P -> swizzled shared memory
        const int jj = it - nkv;
        const float mb = m * scale_log2e;
        float p[64];
#pragma unroll
        for (int c = 0; c < 32; c++) {
          p[c] = c < valid ? exp2f(fmaf(__uint_as_float(v0[c]), scale_log2e, -mb)) : 0.f;
          p[c + 32] = c + 32 < valid ? exp2f(fmaf(__uint_as_float(v1[c]), scale_log2e, -mb)) : 0.f;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        a_mbar_wait(p_empty, (jj & 1) ^ 1);  // the previous P V MMA has finished reading the P buffer
#pragma unroll
        for (int c16 = 0; c16 < 8; c16++) {
          uint4 u;
          __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const __nv_bfloat162 pk = __floats2bfloat162_rn(p[c16 * 8 + 2 * t], p[c16 * 8 + 2 * t + 1]);
            hh[t] = pk;
            // accumulate the row sum from the ROUNDED probabilities so that O / l is consistent with what the MMA sees
            const float2 back = __bfloat1622float2(pk);
            l += back.x + back.y;
          }
          *reinterpret_cast<uint4*>(prow + ((c16 ^ (row & 7)) << 4)) = u;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
        __syncwarp();
        if (lane == 0) {
          a_mbar_arrive(s_empty);
          a_mbar_arrive(p_full);
        }
      }
    }
    // epilogue
    a_mbar_wait(o_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t o0[32], o1[32];
    a_tmem_ld32(tmem_O + lane_addr, o0);
    a_tmem_ld32(tmem_O + lane_addr + 32, o1);
    const int qrow = q0 + row;
    if (qrow < Nq) {
      const float inv = 1.0f / l;
      __nv_bfloat16* op = O + (long long)b * so_b + (long long)qrow * so_n + (long long)h * so_h;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        uint4 u, w;
        __nv_bfloat162* hu = reinterpret_cast<__nv_bfloat162*>(&u);
        __nv_bfloat162* hw = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
        for (int t = 0; t < 4; t++) {
          hu[t] = __floats2bfloat162_rn(__uint_as_float(o0[c + 2 * t]) * inv, __uint_as_float(o0[c + 2 * t + 1]) * inv);
          hw[t] = __floats2bfloat162_rn(__uint_as_float(o1[c + 2 * t]) * inv, __uint_as_float(o1[c + 2 * t + 1]) * inv);
        }
        *reinterpret_cast<uint4*>(op + c) = u;
        *reinterpret_cast<uint4*>(op + 32 + c) = w;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
  }
}

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
  static bool configured = false;
  if (!configured) {
    S3R_CUDA_CHECK(cudaFuncSetAttribute(s3r_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem::TOTAL));
    configured = true;
  }
  dim3 grid((Nq + ATT_BM - 1) / ATT_BM, H, B);
  s3r_attention_kernel<<<grid, ATT_THREADS, AttSmem::TOTAL, (cudaStream_t)stream>>>(
      tq, tk, tv, (__nv_bfloat16*)o, H, Nq, Nk, o_strides[0], o_strides[1], o_strides[2], scale * 1.4426950408889634f);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
