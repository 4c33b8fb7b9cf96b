/*
 * styl3r_b200 — C-ABI of the B200-native (sm_100a) Styl3R hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers / sizes / a
 * cudaStream_t (passed as void*), returns 0 or a negative S3R_ERR_* code, never
 * throws, never allocates on the caller's behalf and never synchronises the
 * stream unless the name says so.  The caller (PyTorch on the Python side) owns
 * all memory.
 *
 * Reference interfaces these entry points replace (paths relative to the
 * Styl3R repo):
 *   - s3r_raster_*      : third-party `diff_gaussian_rasterization._C.
 *                         rasterize_gaussians{,_backward}` as driven by
 *                         src/model/decoder/cuda_splatting.py:101-129 (one call
 *                         per view there; here one call renders all views).
 *   - s3r_rope2d        : src/model/encoder/backbone/croco/curope/curope.cpp:49-69
 *                         (`rope_2d`) / kernels.cu:84-108 (`rope_2d_cuda`).
 *   - s3r_se3_update_w2c: src/misc/cam_utils.py:103-137 (SE3_exp, update_pose).
 */
#ifndef STYL3R_B200_H_
#define STYL3R_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3R_ABI_VERSION 3

/* error codes */
#define S3R_OK 0
#define S3R_ERR_INVALID_ARG (-1)
#define S3R_ERR_UNSUPPORTED (-2) /* shape outside the compiled limits          */
#define S3R_ERR_STATE_TOO_SMALL (-3)
#define S3R_ERR_CUDA (-4) /* a CUDA runtime call / launch failed            */
#define S3R_ERR_NO_DEVICE (-5)

/* compiled limits */
#define S3R_TILE 16          /* tile edge in pixels (BLOCK_X = BLOCK_Y)     */
#define S3R_MAX_TILES 4096   /* tiles per view (<= 1024x1024 image)         */
#define S3R_MAX_SH_COEFFS 16 /* SH degree <= 3                              */

/* ------------------------------------------------------------------------
 * Rasterizer
 * ------------------------------------------------------------------------ */

/* Inputs of one batched rasterization: `n_views` cameras; view i renders the
 * Gaussian set `view_set[i]` (or set i when view_set == NULL, which requires
 * n_sets == n_views).  All pointers are device pointers to contiguous fp32
 * unless noted.  Matrices use the layout the reference hands to its
 * rasterizer: 16 floats, m[4*col + row] of the mathematical matrix, i.e. the
 * row-major storage of the *transposed* matrix (cuda_splatting.py:86-88). */
typedef struct s3r_raster_params {
  int32_t n_views;
  int32_t n_sets;
  int32_t P;          /* Gaussians per set                                    */
  int32_t width;      /* image width  (pixels)                                */
  int32_t height;     /* image height (pixels)                                */
  int32_t sh_degree;  /* active SH degree, 0..3                               */
  int32_t sh_coeffs;  /* stored coefficients per Gaussian M >= (deg+1)^2      */
  int32_t cov_stride; /* 6: packed (xx,xy,xz,yy,yz,zz); 9: full row-major 3x3 */
  const float* means3D;        /* [n_sets, P, 3]                              */
  const float* cov3D;          /* [n_sets, P, cov_stride]                     */
  const float* shs;            /* [n_sets, P, M, 3] or NULL                   */
  const float* colors_precomp; /* [n_sets, P, 3] or NULL (exactly one of two) */
  const float* opacities;      /* [n_sets, P]                                 */
  const int32_t* view_set;     /* [n_views] or NULL                           */
  const float* viewmatrix;     /* [n_views, 16] world->camera                 */
  const float* projmatrix;     /* [n_views, 16] full projection (view @ proj) */
  const float* projmatrix_raw; /* [n_views, 16] projection only (backward)    */
  const float* campos;         /* [n_views, 3]                                */
  const float* tanfov;         /* [n_views, 2] (tanfovx, tanfovy)             */
  const float* scales;         /* [n_views] scale-invariant factor 1/near, or
                                  NULL (= 1).  Applied in-kernel exactly like
                                  cuda_splatting.py:65-72: mean*s, cov*(s*s). */
  const float* background;     /* [n_views, 3]                                */
} s3r_raster_params;

typedef struct s3r_raster_outputs {
  float* color;       /* [n_views, 3, H, W]                                   */
  float* depth;       /* [n_views, H, W]  (alpha-blended, not normalised)     */
  float* opacity;     /* [n_views, H, W]  (1 - final transmittance)           */
  int32_t* radii;     /* [n_views, P]     screen radius in px, 0 = culled     */
  int32_t* n_touched; /* [n_views, P] or NULL; caller zero-fills              */
} s3r_raster_outputs;

/* Byte offsets of the arrays inside the opaque `state` buffer (exposed so that
 * parity tests can read tile assignment / sort order without a copy API).
 * nvP = n_views*P, nvT = n_views*tiles, cap = instance capacity.            */
typedef struct s3r_raster_layout {
  int64_t total_bytes;
  int64_t status;        /* int64[4]: R_total, overflow, max_tile_count, rsvd */
  int64_t counters;      /* uint32[8] device-side tickets                     */
  int64_t depths;        /* float   [nvP]   camera-space z                    */
  int64_t xy;            /* float2  [nvP]   pixel-space mean                  */
  int64_t conic_opacity; /* float4  [nvP]                                     */
  int64_t rgb;           /* float4  [nvP]   (r,g,b, clamp-mask as int bits)   */
  int64_t rect;          /* uint32  [nvP]   xmin|ymin<<8|xmax<<16|ymax<<24    */
  int64_t chunk_hist;    /* uint16  [n_views, chunks, tiles]                  */
  int64_t chunk_base;    /* uint32  [n_views, chunks, tiles]                  */
  int64_t tile_count;    /* uint32  [nvT]   instances per (view, tile)        */
  int64_t ranges;        /* uint2   [nvT]   (start,end) into the sorted list  */
  int64_t keys_unsorted; /* uint64  [cap]   (depth_bits<<32 | gaussian) tile-binned */
  int64_t keys_tmp;      /* uint64  [cap]   ping-pong buffer (oversized tiles)  */
  int64_t point_list;    /* uint32  [cap]   sorted gaussian index             */
  int64_t point_keys;    /* uint64  [cap]   sorted ((view*T+tile)<<32 | depth_bits) */
  int64_t records;       /* 48 B    [cap]   sorted-gathered blend records     */
  int64_t final_T;       /* float   [n_views*H*W]                             */
  int64_t n_contrib;     /* uint32  [n_views*H*W]                             */
  int64_t grecords;      /* 48 B    [nvP]   per-(view, Gaussian) blend record, written by preprocess */
  int64_t work_order;    /* uint32  [nvT]   (view*tiles + tile) by descending instance count: blend work queue */
  int64_t blists;        /* uint32  [8*cap] per (view, tile, 8x4-pixel block): indices (inside the tile's sorted range) of the
                            instances whose alpha >= 1/255 box touches the block, in sorted order; block b of a tile with
                            range [s, e) starts at 8*s + b*(e - s)                                                        */
  int64_t bcounts;       /* uint32  [8*nvT] length of every block list                                                  */
  int64_t n_contrib_blk; /* uint32  [n_views*H*W] per pixel: entries of its block list up to and including the last
                            contributor (written by the warp-granular blend kernel, read by its backward twin)           */
  int32_t tiles_x, tiles_y, tiles, chunks;
} s3r_raster_layout;

int s3r_abi_version(void);
const char* s3r_error_string(int code);

/* Fills `out`; returns S3R_OK or S3R_ERR_UNSUPPORTED/INVALID_ARG. Pure host. */
int s3r_raster_layout_query(int32_t n_views, int32_t P, int32_t width, int32_t height,
                            int64_t capacity, s3r_raster_layout* out);

/* Forward: preprocess -> tile binning (MSD radix digit, global memory) ->
 * per-tile depth radix sort -> blend.  Stream-ordered, no host sync.  If the
 * number of (tile, Gaussian) instances exceeds `capacity` the surplus is
 * dropped and status.overflow is set (read it with s3r_raster_read_status). */
int s3r_raster_forward(const s3r_raster_params* params, const s3r_raster_outputs* out, void* state,
                       size_t state_bytes, int64_t capacity, void* stream);

/* Same, but only the stages selected by `stage_mask` are launched (the others'
 * results must already be in `state` from an earlier call with identical
 * arguments).  Used by bench.py to time one stage with CUDA events. */
#define S3R_STAGE_PREPROCESS 1u
#define S3R_STAGE_BIN 2u
#define S3R_STAGE_SORT 4u
#define S3R_STAGE_BLEND 8u
#define S3R_STAGE_ALL 15u
int s3r_raster_forward_stages(const s3r_raster_params* params, const s3r_raster_outputs* out, void* state,
                              size_t state_bytes, int64_t capacity, uint32_t stage_mask, void* stream);

/* Blocking: copies status {num_instances, overflow, max_tile_count, 0}. */
int s3r_raster_read_status(const void* state, int64_t host_out[4], void* stream);

typedef struct s3r_raster_grads {
  /* inputs */
  const float* dL_dcolor; /* [n_views, 3, H, W]                              */
  const float* dL_ddepth; /* [n_views, H, W] or NULL                         */
  /* outputs — accumulated with atomicAdd, caller zero-fills; any may be NULL */
  float* dL_dmeans3D;   /* [n_sets, P, 3]                                    */
  float* dL_dcov3D;     /* [n_sets, P, cov_stride] (same packing as input; for
                           stride 9 the symmetric gradient is written to the
                           upper triangle positions only)                    */
  float* dL_dshs;       /* [n_sets, P, M, 3]                                 */
  float* dL_dcolors;    /* [n_sets, P, 3]                                    */
  float* dL_dopacities; /* [n_sets, P]                                       */
  float* dL_dmeans2D;   /* [n_views, P, 3] NDC-space mean gradient (x,y,0)   */
  float* dL_dtau;       /* [n_views, 6] (rho, theta) pose gradient           */
  /* scratch */
  void* scratch;        /* >= s3r_raster_backward_scratch_bytes              */
  size_t scratch_bytes;
} s3r_raster_grads;

size_t s3r_raster_backward_scratch_bytes(int32_t n_views, int32_t P);

int s3r_raster_backward(const s3r_raster_params* params, const void* state, size_t state_bytes,
                        int64_t capacity, const s3r_raster_grads* grads, void* stream);

/* ------------------------------------------------------------------------
 * Camera set-up of render_cuda (cuda_splatting.py:65-72,81-88; projection.py:
 * 247-261) for n views in one launch: extrinsics [n,4,4] camera-to-world
 * (row-major), intrinsics [n,3,3] normalised, near/far [n].  Outputs are the
 * tensors s3r_raster_params expects: viewmatrix / projmatrix / projmatrix_raw
 * [n,16] (transposed layout), campos [n,3], tanfov [n,2], scales [n].
 * input_is_w2c != 0: `extrinsics` already holds world-to-camera matrices (the
 * pose-align loop carries w2c, cam_utils.py:126-137) and no inverse is taken.
 * ------------------------------------------------------------------------ */
int s3r_camera_setup(const float* extrinsics, const float* intrinsics, const float* near_, const float* far_,
                     int32_t scale_invariant, int32_t input_is_w2c, int32_t n, float* viewmatrix, float* projmatrix,
                     float* projmatrix_raw, float* campos, float* tanfov, float* scales, void* stream);

/* ------------------------------------------------------------------------
 * RoPE-2D (curope replacement). tokens[B,N,H,D] modified in place.
 * dtype: 0 = fp32, 1 = fp16, 2 = bf16.  pos is int64 [B,N,2] (y,x).
 * ------------------------------------------------------------------------ */
int s3r_rope2d(void* tokens, const int64_t* pos, int32_t B, int32_t N, int32_t H, int32_t D,
               int64_t stride_b, int64_t stride_n, int64_t stride_h, float base, float fwd,
               int32_t dtype, void* stream);

/* ------------------------------------------------------------------------
 * bf16 GEMM with fused epilogue on tcgen05/TMEM (the encoder's nn.Linear:
 * croco/blocks.py:70-73,91-93,162-166):
 *   C[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) + residual[M,N]
 * A, W, bias, residual bf16 (row pitches lda/ldw/ldr in elements), C bf16 or
 * fp32 (S3R_EPI_OUT_F32), fp32 accumulation.  flags: S3R_EPI_*.
 * Requires K, lda, ldw, ldc, ldr multiples of 8 and 16-byte aligned pointers.
 * ------------------------------------------------------------------------ */
#define S3R_EPI_BIAS 1
#define S3R_EPI_GELU 2
#define S3R_EPI_RESIDUAL 4
#define S3R_EPI_OUT_F32 8
#define S3R_EPI_ROPE 16
#define S3R_EPI_RELU 32
#define S3R_EPI_DGELU 128 /* C = acc * gelu'(aux): backward of a fused GELU, aux = the saved pre-activation [M, N] bf16 */
#define S3R_EPI_RES_F32 512 /* the residual operand is fp32 [M, N] (ldr in fp32 elements): fp32 residual stream of the ViT trunks */
#define S3R_EPI_SAVE_PRE 256 /* internal: set when s3r_gemm_bf16_majors is given pre_out */
#define S3R_EPI_PDL 64 /* internal: set by the launcher when programmatic dependent launch is enabled (S3R_TUNE_PDL) */
int s3r_gemm_bf16(const void* A, const void* W, const void* bias, const void* residual, void* C, int32_t M,
                  int32_t N, int32_t K, int32_t lda, int32_t ldw, int32_t ldc, int32_t ldr, int32_t flags,
                  void* stream);

/* Same GEMM with RoPE-2D (curope.cpp:11-47 semantics, head_dim 64) applied to
 * output columns [0, rope_cols) in the epilogue, after the bias: rope_pos is
 * the int64 [M, 2] (y, x) position of every output row, rope_table the
 * (cos, sin) table filled by s3r_rope_table for positions 0..rope_max_pos.
 * Used for the qkv / projq / projk projections (blocks.py:97-106,176-182). */
int s3r_gemm_bf16_rope(const void* A, const void* W, const void* bias, const void* residual, void* C, int32_t M,
                       int32_t N, int32_t K, int32_t lda, int32_t ldw, int32_t ldc, int32_t ldr, int32_t flags,
                       const int64_t* rope_pos, const float* rope_table, int32_t rope_cols, int32_t rope_max_pos,
                       void* workspace, size_t workspace_bytes, void* stream);
/* workspace (optional, device memory, >= 16 KiB, ZERO-FILLED once by the caller, private to the stream): enables
 * split-K for grids smaller than the machine — first 16 KiB are self-resetting tile counters, the rest holds fp32
 * partial tiles.  NULL => never split. */
int s3r_rope_table(float* table /* [(max_pos+1)*16*2] */, int32_t max_pos, float base, void* stream);

/* The same GEMM with either operand MN-major, i.e. handed over as stored by the forward pass - the backward of
 * nn.Linear (autograd of blocks.py:61-82,97-134; cuBLAS in the reference) without transposed copies:
 *   C[M,N] = act( opA . opB^T + bias ) (+ aux | * gelu'(aux))
 *   a_mn_major = 0: A is [M, K] row-major (lda);  1: A is [K, M] row-major (lda)
 *   b_mn_major = 0: B is [N, K] row-major (ldb);  1: B is [K, N] row-major (ldb)
 * dgrad: dX = dY . W        -> A = dY [M, Nout] (0), B = W [Nout, Kin] (1), K = Nout, N = Kin
 * wgrad: dW = dY^T . X      -> A = dY [tokens, Nout] (1), B = X [tokens, Kin] (1), M = Nout, N = Kin, K = tokens
 * flags: BIAS | GELU | RELU | OUT_F32 | RESIDUAL (aux added) | DGELU (multiply by gelu'(aux), aux = pre-activation). */
int s3r_gemm_bf16_majors(const void* A, const void* B, const void* bias, const void* aux, void* C, int32_t M, int32_t N,
                         int32_t K, int32_t lda, int32_t ldb, int32_t ldc, int32_t ldaux, int32_t flags,
                         int32_t a_mn_major, int32_t b_mn_major, void* pre_out, void* stream);
/* `batch` independent GEMMs in one launch (grid.z = batch): operand z starts stride_* ELEMENTS after operand z-1.  No
 * epilogue besides S3R_EPI_OUT_F32.  Used by the attention backward (per image x head contractions). */
int s3r_gemm_bf16_batched(const void* A, const void* B, void* C, int32_t M, int32_t N, int32_t K, int32_t lda, int32_t ldb,
                          int32_t ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c, int32_t batch, int32_t flags,
                          int32_t a_mn_major, int32_t b_mn_major, void* stream);
/* pre_out (optional, K-major operands only): bf16 [M, N] with pitch ldc that receives the value BEFORE the activation
 * (after the bias) - what the backward of a fused GELU needs. */

/* ------------------------------------------------------------------------
 * Stride-1 "same" 2-D convolution as an implicit GEMM on tcgen05/TMEM - the
 * DPT heads' nn.Conv2d (heads/dpt_block.py:33-75,121-142,189-218;
 * dpt_head.py:35-70; dpt_gs_head.py:113-157; dpt_gs_sh_head.py:37-74), cuDNN
 * in the reference:
 *   y[n,h,w,:] = act( sum_{kh,kw} x[n,h+kh-pad,w+kw-pad,:] . W[:,kh,kw,:]^T + bias ) + residual[n,h,w,:]
 * x [n,h,w,cin] bf16 NHWC; W [cout, kh*kw, ceil(cin/64)*64] bf16 (channels zero
 * padded to a multiple of 64); bias bf16 [cout]; residual bf16 NHWC [.., cout];
 * y NHWC bf16 or fp32 (S3R_EPI_OUT_F32).  flags: BIAS | RELU | RESIDUAL |
 * OUT_F32 (ReLU is applied before the residual add).  Requires 2*pad = k-1,
 * cin, cout multiples of 8, cin >= 64 and w a divisor or multiple of 128
 * (the 128-pixel tile must be a box of the tensor) - else S3R_ERR_UNSUPPORTED.
 * ------------------------------------------------------------------------ */
int s3r_conv2d_bf16(const void* x, const void* w, const void* bias, const void* residual, void* y, int32_t n,
                    int32_t h, int32_t wd, int32_t cin, int32_t cout, int32_t kh, int32_t kw, int32_t pad,
                    int32_t flags, void* stream);

/* Tuning knob (benchmarks only; process-wide, not thread-safe).  S3R_TUNE_CONV_VARIANT: tile shape of
 * s3r_conv2d_bf16 for cout % 256 == 0 - -1 = auto (default; also env S3R_CONV_VARIANT), 0 = 128x128 tiles,
 * 2 CTAs/SM, 1 = 128x256 tiles with a 4-stage ring (1 CTA/SM), 2 = 128x256 tiles with a 2-stage ring (2 CTAs/SM). */
#define S3R_TUNE_CONV_VARIANT 1
/* S3R_TUNE_GEMM_CLUSTER: thread-block cluster shape of s3r_gemm_bf16 (TMA multicast of the shared operand tile):
 * 0 = built-in default, else CM*10 + CN with CM in {1,2} (adjacent M tiles share the W tile), CN in {1,2,4}.
 * S3R_TUNE_GEMM_BIG_TILE: 128x256 output tiles - 0 = auto (N >= 4096 and >= 3 waves), 1 = whenever >= 120 tiles, 2 = never. */
#define S3R_TUNE_GEMM_CLUSTER 2
#define S3R_TUNE_GEMM_BIG_TILE 3
#define S3R_TUNE_PDL 5 /* != 0: GEMM / conv / attention launches use programmatic dependent launch (prologue overlaps the previous kernel's tail; griddepcontrol.wait before the first global access) */
#define S3R_TUNE_GEMM_SHALLOW 6 /* != 0: GEMM grids smaller than the machine use the 4-stage 96 KB ring (2 CTAs/SM) instead of the 8-stage 192 KB one, so that kernels of concurrent stream branches can share an SM */
#define S3R_TUNE_CONV_CLUSTER 4 /* != 0: s3r_conv2d_bf16 runs clusters of 2 pixel tiles that multicast the weight tile */
#define S3R_TUNE_GEMM_KSPLIT 10 /* cluster split-K of 64-wide GEMM tiles through distributed shared memory: 0 = auto (few tiles, long K), 1 = never, 2 / 4 = forced */
#define S3R_TUNE_GEMM_PAIR 11 /* CTA pairs (tcgen05 cta_group::2: one 256-row MMA over two SMs, each CTA stages its own A rows and half of the B tile): 0 = auto, 1 = 256x128 pair tiles whenever the shape allows, 2 = never, 3 = 256x256 pair tiles whenever the shape allows */
#define S3R_TUNE_ATTN_ONEPASS 12 /* s3r_attention_bf16: 0 / 1 = single pass over the keys with two independent softmax streams per row (default), 2 = the two-pass kernel (exact row maximum first) */
#define S3R_TUNE_GEMM_DIRECT 13 /* one-tile GEMM kernel: register epilogue straight from TMEM (no staging tile): 0 = auto (grids of at most ~1.5 waves: the batch-1 shapes), 1 = whenever the epilogue feature set allows, 2 = never */
#define S3R_TUNE_RASTER_PDL 9 /* bit k != 0: raster stage k (0 preprocess, 1 bin scan, 2 bin emit, 3 tile sort, 4 blend) is launched with programmatic dependent launch */
#define S3R_TUNE_BWD_ALL 14 /* != 0: s3r_raster_backward always runs the 10-value blend-backward kernel instead of the geometry-only / colour-only variants it selects from the requested outputs (tests compare them) */
#define S3R_TUNE_BLEND_KERNEL 15 /* blend (forward) kernel: 0 = warp-granular over the per-block survivor lists (default), 1 = tile-granular (TMA ring, per-warp cull) */
#define S3R_TUNE_BLEND_ONLY_TILE 8 /* development probe: != 0 -> the blend launch renders only tile (value - 1) of every view (one CTA per view): times the critical path of one tile without contention */
int s3r_set_tunable(int32_t key, int32_t value);

/* Bilinear x2 upsampling, align_corners=True (F.interpolate in heads/dpt_block.py:209-211, dpt_head.py:57,
 * dpt_gs_head.py:138), NHWC bf16: y[n,2h,2w,c] = up(x)[...] (+ add[n,2h,2w,c] when add != NULL). c % 8 == 0. */
int s3r_upsample2x_nhwc_bf16(const void* x, const void* add, void* y, int32_t n, int32_t h, int32_t w, int32_t c,
                             void* stream);

/* Patch matrix of the 7x7 / pad 3 image-skip convolution of the Gaussian-parameter head (dpt_gs_head.py:113-118,
 * `input_merger`): img [B,3,H,W] bf16 planar -> cols [B*H*W, 152] bf16, column k < 147 = (ci, kh, kw) in F.unfold order,
 * columns 147..151 zero. */
int s3r_im2col7x7_bf16(const void* img, void* cols, int32_t B, int32_t H, int32_t W, void* stream);

/* Cross-view context of the second decoder (backbone_croco_multiview.py:170-178): x0 [b,l,c] (view 0), x1 [b,v-1,l,c]
 * (views 1..v-1) -> ctx [b, v-1, (v-1)*l, c]: for query view i >= 1 the tokens of all other views in view order.
 * bf16, c % 8 == 0, v >= 2. */
int s3r_gather_other_views_bf16(const void* x0, const void* x1, void* ctx, int32_t b, int32_t v, int32_t l, int32_t c,
                                void* stream);

/* ------------------------------------------------------------------------
 * LayerNorm over the last dimension, bf16 in / out, fp32 statistics - the
 * nn.LayerNorm(dim, eps=1e-6) of the ViT blocks (croco/blocks.py:140-147,
 * 206-217; croco.py:34).  x [M, C] with row pitch ldx (elements), y [M, C]
 * contiguous; C % 256 == 0, C <= 2048.
 * ------------------------------------------------------------------------ */
int s3r_layernorm_bf16(const void* x, const void* weight, const void* bias, void* y, int32_t M, int32_t C, int64_t ldx,
                       float eps, void* stream);
/* Same, x in fp32 (the residual stream of the inference layout is kept in fp32; y is the bf16 GEMM operand). */
int s3r_layernorm_f32_bf16(const float* x, const void* weight, const void* bias, void* y, int32_t M, int32_t C, int64_t ldx,
                           float eps, void* stream);
/* Backward of the same LayerNorm (autograd of nn.LayerNorm in the reference): dx [M, C] bf16 from x (row pitch ldx),
 * weight and dy [M, C] (contiguous); dweight / dbias [C] fp32 are ACCUMULATED with atomicAdd (caller zero-fills; either
 * may be NULL).  Statistics are recomputed from x.  C % 256 == 0, C <= 1024. */
int s3r_layernorm_bwd_bf16(const void* x, const void* weight, const void* dy, void* dx, float* dweight, float* dbias,
                           int32_t M, int32_t C, int64_t ldx, float eps, void* stream);

/* ------------------------------------------------------------------------
 * Attention softmax(q k^T * scale) v on tcgen05/TMEM, head_dim 64, bf16 —
 * replaces xformers.ops.memory_efficient_attention (croco/blocks.py:126-130,
 * 192-196).  q [B,Nq,H,64], k/v [B,Nk,H,64], o [B,Nq,H,64]; *_strides are the
 * element strides of dims (B, N, H) (unit stride on the last dim; multiples of
 * 8); no mask, no dropout.
 * ------------------------------------------------------------------------ */
int s3r_attention_bf16(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t H, int32_t Nq,
                       int32_t Nk, int32_t D, const int64_t* q_strides, const int64_t* k_strides,
                       const int64_t* v_strides, const int64_t* o_strides, float scale, void* stream);

/* Row kernels of the attention backward (styl3r_b200/attention_bwd.py; autograd of memory_efficient_attention,
 * blocks.py:126-130,192-196) between the batched GEMMs: P = softmax(scale * S) (fp32 [rows, ld] -> bf16 [rows, ldp],
 * pad columns zeroed) and dS = scale * P o (dP - rowsum(dO o O)) with O / dO [rows, 64] bf16. */
int s3r_softmax_rows_bf16(const float* S, void* P, int64_t rows, int32_t n, int32_t ld, int32_t ldp, float scale,
                          void* stream);
int s3r_attention_ds_bf16(const void* P, const float* dP, const void* O, const void* dO, void* dS, int64_t rows, int32_t n,
                          int32_t ld, int32_t ldp, float scale, void* stream);

/* ------------------------------------------------------------------------
 * Fused head epilogue -> Gaussians for one context view of a batch
 * (encoder_noposplat_multi_token_style.py:178-251, postprocess.py:45-61,
 * gaussian_adapter.py:122-153, gaussians.py:8-44).  Planar inputs pts_raw
 * [B,3,HW], params [B,8,HW], app [B,3*d_sh,HW]; outputs are written at
 * Gaussian index view*HW + pixel of the scene buffers means [B,G,3], cov
 * [B,G,3,3], harmonics [B,G,3,d_sh], opacities [B,G] (scales [B,G,3] and
 * rotations [B,G,4] optional).  exponent = 2^x of map_pdf_to_opacity.
 * ------------------------------------------------------------------------ */
int s3r_gaussian_adapter(const float* pts_raw, const float* params, const float* app, const float* sh_mask,
                         int32_t B, int32_t HW, int32_t d_sh, int32_t view, int32_t G, float exponent,
                         float* means, float* cov, float* harmonics, float* opacities, float* scales,
                         float* rotations, void* stream);
/* Same, for pixel-major head outputs (the fp32 [B*HW, ld] rows the 1x1-conv GEMM of the bf16 NHWC head path writes):
 * channel c of pixel p of batch b at (b*HW + p)*ld + c. */
int s3r_gaussian_adapter_nhwc(const float* pts_raw, const float* params, const float* app, int32_t ld_pts,
                              int32_t ld_params, int32_t ld_app, const float* sh_mask, int32_t B, int32_t HW,
                              int32_t d_sh, int32_t view, int32_t G, float exponent, float* means, float* cov,
                              float* harmonics, float* opacities, float* scales, float* rotations, void* stream);

/* ------------------------------------------------------------------------
 * Pose update: w2c_out[i] = SE3_exp([rho_i, theta_i]) @ w2c_in[i] (row-major
 * 4x4, fp32).  Mirrors cam_utils.py:103-137 without the per-view host loop.
 * ------------------------------------------------------------------------ */
int s3r_se3_update_w2c(const float* w2c_in, const float* rho, const float* theta, float* w2c_out,
                       int32_t n, void* stream);

/* ------------------------------------------------------------------------
 * .ply export packing (src/model/ply_export.py:26-74): one row of
 * 17 + (save_rest ? 3*(d_sh-1) : 0) float32 per Gaussian in the file's binary
 * layout: x y z, nx ny nz (0), f_dc_0..2, [f_rest_*], opacity, log(scale_0..2),
 * rot_0..3 (w x y z after scipy's from_quat -> as_matrix -> from_matrix ->
 * as_quat round trip, float64).  rotations are xyzw; harmonics [n,3,d_sh].
 * xform (device, optional) = {shift_x, shift_y, shift_z, scale_factor} of the
 * reference's shift_and_scale: mean' = (mean - shift)/factor, scale' = scale/factor.
 * ------------------------------------------------------------------------ */
int s3r_ply_pack(const float* means, const float* scales, const float* rotations, const float* harmonics,
                 const float* opacities, const float* xform, int32_t n, int32_t d_sh, int32_t save_rest, float* out,
                 void* stream);

/* ------------------------------------------------------------------------
 * Input staging: rescale_and_crop + normalize_image (src/dataset/shims/
 * crop_shim.py:11-79, normalize_shim.py:15-18), bit-exact with Pillow's 8-bit
 * LANCZOS resize.  img: float planes [planes, h_in, w_in] in [0,1]
 * (planes = images x channels); resized to (h_s, w_s) through a uint8
 * intermediate (scratch: planes*h_in*w_s bytes), window (crop_row, crop_col,
 * h_out, w_out) cut out, out = u8/255, then (out - mean[c])/std[c] when mean
 * and std (device, [channels]) are given.  Tap tables per output row/column:
 * bounds int32 [n,2] = (first, count), coeffs int32 [n, ksize] in Pillow's
 * 22-bit fixed point (device memory; styl3r_b200.staging builds them).
 * ------------------------------------------------------------------------ */
int s3r_rescale_crop(const float* img, int32_t planes, int32_t channels, int32_t h_in, int32_t w_in, int32_t h_s,
                     int32_t w_s, const int32_t* h_bounds, const int32_t* h_coeffs, int32_t h_ksize,
                     const int32_t* v_bounds, const int32_t* v_coeffs, int32_t v_ksize, int32_t crop_row,
                     int32_t crop_col, int32_t h_out, int32_t w_out, const float* mean, const float* stdv,
                     uint8_t* scratch, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STYL3R_B200_H_ */
